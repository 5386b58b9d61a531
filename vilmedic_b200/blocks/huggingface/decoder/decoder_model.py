"""B200-native DecoderModel — mirror of vilmedic/blocks/huggingface/decoder/decoder_model.py:8-53.

`decoder.proto is None` -> a BertGenerationDecoder-shaped tower (is_decoder=True, add_cross_attention=True, LM head tied to
the word embeddings) built from `decoder` exactly as BertGenerationConfig(**decoder) would be (:23-26);  forward passes
labels=input_ids (:46), i.e. next-token CE where pad tokens count as targets and the last position is ignored.
`decoder.proto` = a LOCAL HuggingFace directory (BERT / RoBERTa / bert-generation checkpoint) builds the causal LM of that family
(is_decoder, add_cross_attention as :19-20; cross-attention blocks absent from the checkpoint stay randomly initialised, exactly like
AutoModelForCausalLM.from_pretrained(path, config=dec_config) :21) on the kernel tower; a hub NAME needs network access ->
NotImplementedError.
state_dict keys are identical to the reference's (`decoder.bert.…`, `decoder.lm_head.…`).
"""
import torch
import torch.nn as nn

from ....cfgutil import cfg_get, to_attrdict
from ....nn import BertTower, bert_config
from .generation import GenerationMixinB200


class BertGenerationDecoderB200(BertTower, GenerationMixinB200):
    """The object exposed as `DecoderModel.decoder` (HF-generate-capable causal LM in the reference,
    vilmedic/blocks/huggingface/decoder/evaluation.py:22,65)."""

    def __init__(self, cfg):
        BertTower.__init__(self, cfg, with_lm_head=True)

    def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                labels=None, keep_logits=None, **kwargs):
        x, B, T = self.hidden_states(input_ids, attention_mask, encoder_hidden_states, encoder_attention_mask)
        out = {"loss": None, "logits": None, "past_key_values": None, "hidden_states": None, "attentions": None,
               "cross_attentions": None}
        if labels is not None:
            if labels is not input_ids and not torch.equal(labels.to(input_ids.device), input_ids):
                raise NotImplementedError("only labels=input_ids (the reference's call, decoder_model.py:46) is wired")
            grad = torch.is_grad_enabled()
            if keep_logits is None:
                keep_logits = not (grad and self.training)
            loss, logits = self.lm_loss(x, input_ids, B, T, keep_logits)
            out["loss"] = loss
            out["logits"] = logits if logits.numel() else None
        else:
            out["logits"] = self.lm_logits(x).view(B, T, -1)
        return out


class DecoderModel(nn.Module):
    def __init__(self, decoder, **kwargs):
        super().__init__()
        from ....hf_loader import is_local_checkpoint, load_into, read_config
        decoder = to_attrdict(decoder)
        proto = cfg_get(decoder, "proto")
        d = dict(decoder)
        d.pop("proto", None)
        if proto is not None:
            if not is_local_checkpoint(proto):
                raise NotImplementedError("DecoderModel(proto=%r): not a local HuggingFace directory (hub access is not available)" % (proto,))
            d = read_config(proto)
        d["is_decoder"] = True
        d["add_cross_attention"] = True
        self.decoder = BertGenerationDecoderB200(bert_config(**d))
        if proto is not None:
            missing, unexpected = load_into(self.decoder, proto, flat=False)
            missing = [k for k in missing if "crossattention" not in k]      # newly initialised, as HF reports for the same call
            if missing or unexpected:
                raise RuntimeError("proto %r does not match the tower: missing %s, unexpected %s" % (proto, missing[:5], unexpected[:5]))
        self.generate = self.decoder.generate
        self.config = self.decoder.config

    def forward(self, input_ids, attention_mask, encoder_outputs=None, encoder_attention_mask=None, **kwargs):
        input_ids = input_ids.cuda(non_blocking=True)
        attention_mask = attention_mask.cuda(non_blocking=True)
        return self.decoder(input_ids=input_ids, attention_mask=attention_mask, encoder_hidden_states=encoder_outputs,
                            encoder_attention_mask=encoder_attention_mask, labels=input_ids, **kwargs)

    def __repr__(self):
        return "BertGenerationDecoderB200(" + str(self.decoder.config) + ")\n"
