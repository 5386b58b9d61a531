"""RRG_HF — mirror of vilmedic/models/rrg/RRG_HF.py:18-177 (VisionEncoderDecoder-style composition, multi-image aware).

Supported: `vision` = {proto_model: "vit", proto_config: "vit", proto_config_args: {...}} and
`decoder` = {proto_model: "bert-generation", proto_config: "bert-generation", proto_config_args: {...}} (:27-87) —
the ViT -> BERT-decoder path of the BASELINE configs.  `encoderdecoder=<hub name>` / string protos (:25,48,84) need the HF
hub -> NotImplementedError.  forward (:108-177): 5-D images are flattened to B*N crops, encoded at once, concatenated to
[B, N*S, D], optional enc_to_dec_proj, patch-level mask from images_mask; 4-D images pass encoder_attention_mask=None.
state_dict keys follow VisionEncoderDecoderModel: `model.encoder.…`, `model.decoder.…`, `model.enc_to_dec_proj.…`.
"""
import torch
import torch.nn as nn

from ...blocks.huggingface.decoder.decoder_model import BertGenerationDecoderB200
from ...cfgutil import cfg_get, to_attrdict
from ...nn import ViTTower, bert_config, native_linear, set_arena_root


class _ViTWithPooler(ViTTower):
    """ViTModel(config) as built by RRG_HF.py:39 keeps its (unused) pooler parameters."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.pooler = nn.Module()
        self.pooler.dense = nn.Linear(self.cfg.hidden_size, self.cfg.hidden_size)


class RRG_HF(nn.Module):
    def __init__(self, encoderdecoder=None, decoder=None, vision=None, dl=None, **kwargs):
        super().__init__()
        assert (encoderdecoder is None) ^ (decoder is None or vision is None), \
            "Either proto should be provided, or both decoder and vision should be provided."
        if encoderdecoder is not None or isinstance(vision, str) or isinstance(decoder, str):
            raise NotImplementedError("pretrained HF-hub checkpoints are not reachable offline")
        vision, decoder = to_attrdict(vision), to_attrdict(decoder)
        assert "proto_model" in vision and "proto_config" in vision
        assert "proto_model" in decoder and "proto_config" in decoder
        if vision.pop("proto_model") != "vit" or vision.pop("proto_config") != "vit":
            raise NotImplementedError("RRG_HF vision tower: only 'vit' runs on the B200 kernels")
        if decoder.pop("proto_model") != "bert-generation" or decoder.pop("proto_config") != "bert-generation":
            raise NotImplementedError("RRG_HF decoder: only 'bert-generation' runs on the B200 kernels")
        v_args = dict(vision.pop("proto_config_args")) if "proto_config_args" in vision else {}
        d_args = dict(decoder.pop("proto_config_args")) if "proto_config_args" in decoder else {}
        if dl:
            tok = dl.dataset.seq.tokenizer
            d_args.update(vocab_size=tok.vocab_size, unk_token_id=tok.unk_token_id, bos_token_id=tok.cls_token_id,
                          eos_token_id=tok.sep_token_id, pad_token_id=tok.pad_token_id)
        d_args.update(is_decoder=True, add_cross_attention=True)
        self.model = nn.Module()
        self.model.encoder = _ViTWithPooler(**v_args)
        enc_d = self.model.encoder.cfg.hidden_size
        cross = d_args.get("cross_attention_hidden_size", None)
        dec_hidden = d_args.get("hidden_size", 1024)
        if enc_d != dec_hidden and cross is None:
            self.model.enc_to_dec_proj = nn.Linear(enc_d, dec_hidden)
        elif cross is not None:
            d_args["encoder_hidden_size"] = cross
        self.model.decoder = BertGenerationDecoderB200(bert_config(**d_args))
        assert self.model.decoder.config.is_decoder and self.model.decoder.config.add_cross_attention
        self.eval_func = None
        set_arena_root(self)

    def _encode(self, flat_pixels):
        h = self.model.encoder(flat_pixels)                                    # [B', S, D] bf16
        if hasattr(self.model, "enc_to_dec_proj"):
            Bp, S, D = h.shape
            h = native_linear(self.model.enc_to_dec_proj, h.reshape(Bp * S, D), self).view(Bp, S, -1)
        return h

    def forward(self, input_ids, attention_mask, images, images_mask=None, epoch=None, iteration=None, **kwargs):
        input_ids = input_ids.cuda(non_blocking=True)
        attention_mask = attention_mask.cuda(non_blocking=True)
        images = images.cuda(non_blocking=True)
        if images.dim() == 5:
            B, N, C, H, W = images.shape
            mask = torch.ones((B, N), dtype=torch.bool, device=images.device) if images_mask is None \
                else images_mask.to(images.device).bool()
            h = self._encode(images.reshape(B * N, C, H, W))
            S, D = h.shape[1], h.shape[2]
            enc = h.reshape(B, N * S, D)
            attn_mask = mask.unsqueeze(-1).expand(B, N, S).reshape(B, N * S).long()
            return self.model.decoder(input_ids=input_ids, attention_mask=attention_mask, encoder_hidden_states=enc,
                                      encoder_attention_mask=attn_mask, labels=input_ids, **kwargs)
        if images.dim() == 4:
            enc = self._encode(images)
            return self.model.decoder(input_ids=input_ids, attention_mask=attention_mask, encoder_hidden_states=enc,
                                      encoder_attention_mask=None, labels=input_ids, **kwargs)
        raise NotImplementedError(f"Unexpected images.dim() = {images.dim()}")
