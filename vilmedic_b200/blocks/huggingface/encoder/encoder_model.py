"""B200-native EncoderModel — mirror of vilmedic/blocks/huggingface/encoder/encoder_model.py:10-66: a bidirectional
BERT-shaped text tower (BertGenerationEncoder when `encoder.proto` is None) with an optional BertPooler
(tanh(W h[:,0] + b), :28-29,58-60).  `encoder.proto` = a LOCAL HuggingFace directory (config.json + weights of a BERT / RoBERTa /
bert-generation model) loads that checkpoint into the kernel tower (hf_loader.py) — the offline form of :20-22
`AutoModel.from_pretrained(proto)`; a hub NAME needs network access -> NotImplementedError."""
import torch
import torch.nn as nn

from ....arena import get_arena
from ....cfgutil import cfg_get, to_attrdict
from ....nn import BertTower, LinearFn, TanhFn, _lin, _root_of, bert_config


class EncoderModel(nn.Module):
    def __init__(self, encoder, **kwargs):
        super().__init__()
        from ....hf_loader import is_local_checkpoint, load_into, read_config
        encoder = to_attrdict(encoder)
        proto = cfg_get(encoder, "proto")
        d = dict(encoder)
        d.pop("proto", None)
        add_pool = bool(d.pop("add_pooling_layer", False))
        if proto is not None:
            if not is_local_checkpoint(proto):
                raise NotImplementedError("EncoderModel(proto=%r): not a local HuggingFace directory (hub access is not available)" % (proto,))
            d = read_config(proto)                          # AutoConfig.from_pretrained(path): the checkpoint's own architecture
        d["is_decoder"] = False
        d["add_cross_attention"] = False
        self.encoder = BertTower(bert_config(**d), with_lm_head=False, flat=True)
        self.config = self.encoder.config
        self.pooler = None
        if add_pool:
            self.pooler = nn.Module()
            self.pooler.dense = nn.Linear(self.config.hidden_size, self.config.hidden_size)
        if proto is not None:
            missing, unexpected = load_into(self.encoder, proto, flat=True)
            if missing or unexpected:
                raise RuntimeError("proto %r does not match the tower: missing %s, unexpected %s" % (proto, missing[:5], unexpected[:5]))

    def forward(self, input_ids, attention_mask=None, output_hidden_states=None, **kwargs):
        input_ids = input_ids.cuda(non_blocking=True)
        attention_mask = attention_mask.cuda(non_blocking=True) if attention_mask is not None else None
        states = [] if output_hidden_states else None
        x, B, T = self.encoder.hidden_states(input_ids, attention_mask, collect=states)
        D = x.shape[-1]
        out = {"last_hidden_state": x.view(B, T, D), "pooler_output": None}
        if states is not None:                       # embedding output + one entry per layer, as HF's `hidden_states` tuple
            out["hidden_states"] = tuple(h.view(B, T, D) for h in states)
        if self.pooler is not None:
            arena = get_arena(_root_of(self))
            first = x.view(B, T, D)[:, 0].contiguous()
            h = LinearFn.apply(first, self.pooler.dense.weight, _lin(arena, self.pooler.dense), torch.float32)
            out["pooler_output"] = TanhFn.apply(h)
        return out
