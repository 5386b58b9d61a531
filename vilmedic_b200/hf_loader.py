"""Loading a pretrained text tower from a LOCAL HuggingFace directory (`proto: /path/to/dir`) — the offline form of
`AutoModel.from_pretrained(proto)` / `AutoModelForCausalLM.from_pretrained(proto)` in
vilmedic/blocks/huggingface/encoder/encoder_model.py:20-22 and decoder/decoder_model.py:17-21 (every shipped RRG / SELFSUP config
names a hub checkpoint there, e.g. allenai/biomed_roberta_base; hub access is not available to this package, a directory written by
`save_pretrained` is).

Supported `model_type`s: "bert", "roberta" (BERT-shaped post-LN stacks; RoBERTa's padding-aware position ids and the one-row token-type
table are handled by the embedding kernel, the LM-head transform dense -> GELU -> LayerNorm by the GEMM / LayerNorm kernels) and
"bert-generation".  The tower keeps HF's parameter names for that family, so the checkpoint's state_dict loads key by key; as in HF,
cross-attention blocks that the checkpoint does not contain stay randomly initialised when a decoder is built from an encoder-only
checkpoint (HF prints the same "newly initialized" warning).
Host-side dictionary work only; no arithmetic."""
import json
import os

import torch

_ARCH_KEYS = ("vocab_size", "hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "hidden_act",
              "hidden_dropout_prob", "attention_probs_dropout_prob", "max_position_embeddings", "initializer_range", "layer_norm_eps",
              "pad_token_id", "bos_token_id", "eos_token_id", "type_vocab_size", "tie_word_embeddings")


def is_local_checkpoint(proto):
    return isinstance(proto, str) and os.path.isdir(proto) and os.path.exists(os.path.join(proto, "config.json"))


def read_config(path):
    """config.json -> kwargs for nn.bert_config (family + the architecture keys the kernels consume)."""
    with open(os.path.join(path, "config.json")) as f:
        raw = json.load(f)
    family = raw.get("model_type")
    if family not in ("bert", "roberta", "bert-generation"):
        raise NotImplementedError("proto %r: model_type %r has no kernel tower (bert, roberta, bert-generation)" % (path, family))
    if raw.get("position_embedding_type", "absolute") != "absolute":
        raise NotImplementedError("position_embedding_type %r is not supported" % raw["position_embedding_type"])
    kw = {k: raw[k] for k in _ARCH_KEYS if k in raw and raw[k] is not None}
    kw["family"] = family
    if family == "bert":
        kw.setdefault("pad_token_id", 0)
        kw.setdefault("type_vocab_size", 2)
    if family == "roberta":
        kw.setdefault("pad_token_id", 1)
        kw.setdefault("type_vocab_size", 1)
    return kw


def read_state_dict(path):
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return load_file(st)
    pt = os.path.join(path, "pytorch_model.bin")
    if os.path.exists(pt):
        return torch.load(pt, map_location="cpu", weights_only=True)
    raise FileNotFoundError("no model.safetensors / pytorch_model.bin under %s" % path)


def load_into(tower, path, flat):
    """Copy the checkpoint's tensors into `tower` (nn.BertTower).  flat=True: the tower has AutoModel's layout (`embeddings.…`,
    `encoder.layer.…`, `pooler.…`); a checkpoint saved from a *ForCausalLM / *ForMaskedLM class carries a `bert.` / `roberta.` prefix
    that is stripped.  flat=False: the tower has the *ForCausalLM layout; an encoder-only checkpoint gets the prefix added.
    Returns (missing, unexpected) like load_state_dict(strict=False); tied / derived entries are ignored."""
    sd = read_state_dict(path)
    fam = tower.cfg.family
    prefix = "roberta." if fam == "roberta" else "bert."
    out = {}
    for k, v in sd.items():
        if k.endswith("position_ids") or k.endswith("token_type_ids"):
            continue                                            # registered buffers of older transformers versions
        if flat:
            if k.startswith(prefix):
                k = k[len(prefix):]
            elif k.startswith(("lm_head.", "cls.")):
                continue
        else:
            if not k.startswith((prefix, "lm_head.", "cls.")):
                k = prefix + k
        out[k] = v
    own = tower.state_dict()
    missing = [k for k in own if k not in out]
    unexpected = [k for k in out if k not in own]
    for k in list(out):
        if k in own and tuple(out[k].shape) != tuple(own[k].shape):
            raise RuntimeError("size mismatch for %s: checkpoint %s vs model %s" % (k, tuple(out[k].shape), tuple(own[k].shape)))
    tower.load_state_dict({k: v for k, v in out.items() if k in own}, strict=False)
    # entries a task head of another class owns (pooler of an LM checkpoint, seq_relationship, ...) are not errors
    unexpected = [k for k in unexpected if not k.startswith(("cls.seq_relationship", "pooler.", prefix + "pooler."))]
    tied = ("lm_head.decoder.weight", "lm_head.decoder.bias", "cls.predictions.decoder.weight", "cls.predictions.decoder.bias")
    missing = [k for k in missing if k not in tied]
    return missing, unexpected
