"""ConVIRT — mirror of vilmedic/models/selfsup/conVIRT.py:46-110: text tower (EncoderModel pooler_output) and image tower
(VisualEncoder) -> Linear-ReLU-Linear projections (:58-67) -> ConVIRT / InfoNCE loss (:69,100).  Micro-batching by
`forward_batch_size` is kept (:83); negatives are the rank-local batch as in the reference."""
import numpy as np
import torch
import torch.nn as nn

from ...blocks.huggingface.encoder.encoder_model import EncoderModel
from ...blocks.losses import *  # noqa: F401,F403
from ...blocks.vision import *  # noqa: F401,F403
from ...cfgutil import cfg_get, to_attrdict
from ...nn import ReluFn, native_linear, set_arena_root


def evaluation(models, config, dl, from_training, **kwargs):
    """Mirror of vilmedic/models/selfsup/conVIRT.py:14-38: no ensembling (model 0), mean loss over the batches, and — outside
    training — the concatenated projected embeddings for post-processing."""
    model = models[0]
    losses, linguistics, visuals = [], [], []
    with torch.no_grad():
        for batch in dl:
            batch = {k: v.cuda() if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
            out = model(**batch)
            losses.append(out["loss"].mean().cpu().data.numpy())
            if not from_training:
                linguistics.append(out["linguistic"].cpu().data)
                visuals.append(out["visual"].cpu().data)
    if from_training:
        return {"loss": np.ndarray.mean(np.array(losses))}
    return {"loss": np.ndarray.mean(np.array(losses)), "linguistic": torch.cat(linguistics), "visual": torch.cat(visuals)}


def chunks(lst, n):
    for i in range(0, len(lst), n):
        yield lst[i:i + n]


class _MLP(nn.Sequential):
    """nn.Sequential(Linear, ReLU, Linear) parameter layout (keys 0.weight, 2.weight), math on the kernels."""

    def __init__(self, d_in, d_out):
        super().__init__(nn.Linear(d_in, d_out), nn.ReLU(), nn.Linear(d_out, d_out))

    def forward(self, x):
        h = native_linear(self[0], x.reshape(-1, x.shape[-1]), self, out_dtype=torch.float32)
        h = ReluFn.apply(h)
        return native_linear(self[2], h, self, out_dtype=torch.float32)


class ConVIRT(nn.Module):
    def __init__(self, encoder, cnn, projection, loss, forward_batch_size=256, **kwargs):
        super().__init__()
        cnn, loss, projection = to_attrdict(cnn), to_attrdict(loss), to_attrdict(projection)
        self.linguistic = EncoderModel(encoder)
        self.visual = eval(cnn.pop("proto"))(**cnn)
        self.vis_proj = _MLP(projection.visual_embedding_dim, projection.projection_dim)
        self.lin_proj = _MLP(projection.textual_embedding_dim, projection.projection_dim)
        self.loss_fn = eval(loss.pop("proto"))(**loss)
        self.fbs = forward_batch_size
        self.eval_func = evaluation
        set_arena_root(self)

    def forward(self, input_ids, attention_mask, images, **kwargs):
        images = images.cuda(non_blocking=True)
        input_ids = input_ids.cuda(non_blocking=True)
        attention_mask = attention_mask.cuda(non_blocking=True)
        bs = images.shape[0]
        linguistics, visuals = [], []
        for i in list(chunks(range(bs), min(self.fbs, bs))):
            sl = slice(i[0], i[-1] + 1)
            linguistic = self.linguistic(input_ids=input_ids[sl], attention_mask=attention_mask[sl])
            linguistics.append(self.lin_proj(linguistic["pooler_output"]))
            visuals.append(self.vis_proj(self.visual(images[sl])))
        linguistics = torch.cat(linguistics)
        visuals = torch.cat(visuals)
        loss, loss_l, loss_v = self.loss_fn(linguistics, visuals)
        return {"loss": loss, "loss_l": loss_l, "loss_v": loss_v, "linguistic": linguistics, "visual": visuals}
