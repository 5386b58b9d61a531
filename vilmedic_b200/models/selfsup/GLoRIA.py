"""GLoRIA — mirror of vilmedic/models/selfsup/GLoRIA.py:47-130 (config/SELFSUP/gloria-mimic.yml): text tower with the last
`last_n_layers` hidden states aggregated to word level, ResNet image tower with global (average-pooled) and local (layer3 feature
map) embeddings, GLoRIA global + local contrastive loss.

Reference defect resolved to the intended semantics (SURVEY.md §8 "Reference defects" #4): the constructor hooks
`self.visual.cnn[6]`, an attribute renamed to `.model` in v1.3.2, so the reference raises at HEAD; child 6 of the truncated torchvision
ResNet is `layer3`, which is what is tapped here (VisualEncoder.tap_stage = 3, cnn.py).
Kernels: ResNet-50 tower (cnn.py), BERT tower (nn.BertTower), 1x1 local embedder and global embedder as tcgen05 GEMMs, GLoRIALoss
(blocks/losses/gloria.py).  The bilinear up-sampling of the input images to 299x299 (:64) is input preprocessing and the word-piece
aggregation (:133-191) is index bookkeeping over [B, T] tokens; both stay torch glue on the device."""
import numpy as np
import torch
import torch.nn as nn

from ...blocks.huggingface.encoder.encoder_model import EncoderModel
from ...blocks.losses import GLoRIALoss
from ...blocks.vision import *  # noqa: F401,F403
from ...cfgutil import cfg_get, to_attrdict
from ...nn import native_linear, set_arena_root


def evaluation(models, config, dl, from_training, **kwargs):
    """vilmedic/models/selfsup/GLoRIA.py:14-38."""
    model = models[0]
    losses, linguistics, visuals = [], [], []
    with torch.no_grad():
        for batch in dl:
            batch = {k: v.cuda() if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
            out = model(**batch)
            losses.append(out["loss"].mean().cpu().data.numpy())
            if not from_training:
                linguistics.append(out["sent_embeddings"].cpu().data)
                visuals.append(out["global_features"].cpu().data)
    if from_training:
        return {"loss": np.ndarray.mean(np.array(losses))}
    return {"loss": np.ndarray.mean(np.array(losses)), "linguistic": torch.cat(linguistics), "visual": torch.cat(visuals)}


def chunks(lst, n):
    for i in range(0, len(lst), n):
        yield lst[i:i + n]


class GLoRIA(nn.Module):
    def __init__(self, encoder, cnn, visual_embedder, loss, dl, forward_batch_size=12, **kwargs):
        super().__init__()
        encoder, cnn, visual_embedder, loss = to_attrdict(encoder), to_attrdict(cnn), to_attrdict(visual_embedder), to_attrdict(loss)
        self.last_n_layers = encoder.pop("last_n_layers")
        self.linguistic = EncoderModel(encoder)
        self.idxtoword = {v: k for k, v in dl.dataset.tokenizer.get_vocab().items()}
        self.visual = eval(cnn.pop("proto"))(**cnn)
        if getattr(self.visual, "_resnet", None) is None:
            raise NotImplementedError("GLoRIA needs a torchvision ResNet image tower (local features = layer3)")
        object.__setattr__(self.visual, "tap_stage", 3)
        hidden = self.linguistic.encoder.config.hidden_size
        self.global_embedder = nn.Linear(visual_embedder.feature_dim, hidden)
        self.local_embedder = nn.Conv2d(visual_embedder.interm_feature_dim, hidden, kernel_size=1, stride=1, padding=0, bias=False)
        self.up_sample = nn.Upsample(size=(299, 299), mode="bilinear", align_corners=True)
        self.loss_fn = GLoRIALoss(**loss)
        self.eval_func = evaluation
        self.fbs = forward_batch_size
        set_arena_root(self)

    def _local_embed(self, tap, shp):
        """1x1 convolution without bias on the NHWC layer3 map == one GEMM over [B*h*w, C] rows; -> fp32 [B, hidden, h, w]."""
        B, h, w, C = shp
        lin = self.local_embedder
        y = _Conv1x1Fn.apply(tap.reshape(B * h * w, C), lin.weight, self)              # [B*h*w, hidden] fp32
        return y.view(B, h, w, -1).permute(0, 3, 1, 2)

    def forward(self, input_ids, attention_mask, images, **kwargs):
        bs = images.shape[0]
        global_features, local_features, hidden_states = [], [], []
        for i in list(chunks(range(bs), min(self.fbs, bs))):
            sl = slice(i[0], i[-1] + 1)
            pooled = self.visual(self.up_sample(images[sl].cuda(non_blocking=True).float()))      # [b, 2048]
            global_features.append(native_linear(self.global_embedder, pooled.reshape(pooled.shape[0], -1), self, out_dtype=torch.float32))
            local_features.append(self._local_embed(*self.visual.tapped))
            output = self.linguistic(input_ids[sl].cuda(non_blocking=True), attention_mask[sl].cuda(non_blocking=True),
                                     output_hidden_states=True)
            hidden_states.append(torch.stack([h.float() for h in output["hidden_states"]]))
        global_features = torch.cat(global_features)
        local_features = torch.cat(local_features)
        hidden_states = torch.cat(hidden_states, dim=1)
        embeddings = hidden_states[-self.last_n_layers:]
        embeddings, sents = self.aggregate_tokens(embeddings, input_ids)
        sent_embeddings = torch.sum(torch.mean(embeddings, dim=2), dim=1)
        word_embeddings = torch.sum(embeddings, dim=1).permute(0, 2, 1)
        loss, attention_maps = self.loss_fn(global_features, local_features, word_embeddings, sent_embeddings, sents)
        return {"loss": loss, "global_features": global_features, "local_features": local_features,
                "word_embeddings": word_embeddings, "sent_embeddings": sent_embeddings}

    def aggregate_tokens(self, embeddings, input_ids):
        """Word-piece -> word aggregation of vilmedic/models/selfsup/GLoRIA.py:133-191: the embeddings of the pieces of one word are
        summed, the sentence ends at [SEP], the rest is zero padding."""
        num_layers, batch_size, num_words, dim = embeddings.shape
        embeddings = embeddings.permute(1, 2, 0, 3)
        agg_embs_batch, sentences = [], []
        ids_host = input_ids.tolist()
        for embs, caption_id in zip(embeddings, ids_host):
            agg_embs, token_bank, words, word_bank = [], [], [], []
            for word_emb, word_id in zip(embs, caption_id):
                word = self.idxtoword[word_id]
                if word == "[SEP]":
                    agg_embs.append(torch.stack(token_bank).sum(axis=0))
                    words.append("".join(word_bank))
                    agg_embs.append(word_emb)
                    words.append(word)
                    break
                if not word.startswith("##"):
                    if len(word_bank) == 0:
                        token_bank.append(word_emb)
                        word_bank.append(word)
                    else:
                        agg_embs.append(torch.stack(token_bank).sum(axis=0))
                        words.append("".join(word_bank))
                        token_bank = [word_emb]
                        word_bank = [word]
                else:
                    token_bank.append(word_emb)
                    word_bank.append(word[2:])
            agg_embs = torch.stack(agg_embs)
            padding_size = num_words - len(agg_embs)
            paddings = torch.zeros(padding_size, num_layers, dim, device=agg_embs.device, dtype=agg_embs.dtype)
            words = words + ["[PAD]"] * padding_size
            agg_embs_batch.append(torch.cat([agg_embs, paddings]))
            sentences.append(words)
        agg_embs_batch = torch.stack(agg_embs_batch).permute(0, 2, 1, 3)
        return agg_embs_batch, sentences

    def __repr__(self):
        n = sum(p.numel() for p in self.parameters())
        return "GLoRIA\n%s\n%s\n%s\n{'n_params': %d}\n" % (self.visual, self.linguistic, self.loss_fn, n)


class _Conv1x1Fn(torch.autograd.Function):
    """y = x W^T for a bias-free 1x1 Conv2d weight [Cout, Cin, 1, 1] on NHWC rows (tcgen05 GEMM; wgrad / dgrad by the same kernel)."""

    @staticmethod
    def forward(ctx, x, weight, module):
        from ... import ops
        from ...arena import get_arena
        from ...nn import _prepare, _root_of
        arena = get_arena(_root_of(module))
        _prepare(arena, module)
        Cout, Cin = weight.shape[0], weight.shape[1]
        w = arena.bf16(weight, shape=(Cout, Cin))
        x = x.contiguous()
        ctx.saved = (x, w, arena.grad(weight, shape=(Cout, Cin)))
        return ops.gemm(x, w, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, dy):
        from ... import ops
        from ...nn import _bf16_rows
        x, w, gw = ctx.saved
        dy = _bf16_rows(dy)
        ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=gw, accumulate=True)
        return ops.gemm(dy, w, b_mn_major=True), None, None
