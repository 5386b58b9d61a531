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
from ...blocks.huggingface.encoder_decoder.vision_multi_evaluation import evaluation as evaluation_multi
from ...cfgutil import cfg_get, to_attrdict
from ...nn import ViTTower, bert_config, native_linear, set_arena_root


class _HFConfig:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _ViTWithPooler(ViTTower):
    """ViTModel(config) as built by RRG_HF.py:39 keeps its (unused) pooler parameters."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.pooler = nn.Module()
        self.pooler.dense = nn.Linear(self.cfg.hidden_size, self.cfg.hidden_size)


class VisionEncoderDecoderB200(nn.Module):
    """The `model` attribute of RRG_HF: what VisionEncoderDecoderModel is to the reference (RRG_HF.py:25-92) — `.encoder`,
    `.decoder`, optional `.enc_to_dec_proj`, `.config`, and a `generate` with the call shapes the reference's evaluation uses
    (vision_multi_evaluation.py:98-111): `generate(pixel_values, generation_config=...)` or
    `generate(generation_config=..., encoder_outputs=<.last_hidden_state>, attention_mask=<encoder patch mask>)`."""

    def project(self, h):
        """enc_to_dec_proj over [B, S, D] features (HF:vision_encoder_decoder/modeling_vision_encoder_decoder.py, applied when the
        encoder and decoder widths differ and cross_attention_hidden_size is None)."""
        if getattr(self, "enc_to_dec_proj", None) is None:
            return h
        Bp, S, D = h.shape
        return native_linear(self.enc_to_dec_proj, h.reshape(Bp * S, D), self).view(Bp, S, -1)

    @torch.no_grad()
    def generate(self, pixel_values=None, generation_config=None, encoder_outputs=None, attention_mask=None, **kwargs):
        if encoder_outputs is None:
            if pixel_values is None:
                raise ValueError("generate needs pixel_values or encoder_outputs")
            enc = self.project(self.encoder(pixel_values.cuda(non_blocking=True)))
        else:
            enc = encoder_outputs.last_hidden_state if hasattr(encoder_outputs, "last_hidden_state") else encoder_outputs[0]
        B = enc.shape[0]
        gc = generation_config
        start = getattr(gc, "decoder_start_token_id", None) if gc is not None else kwargs.get("decoder_start_token_id")
        if start is None:
            start = getattr(self.config, "decoder_start_token_id", None)
        if start is None:
            start = getattr(gc, "bos_token_id", None) if gc is not None else None
        if start is None:
            raise ValueError("`decoder_start_token_id` or `bos_token_id` has to be defined for encoder-decoder generation.")
        ids = torch.full((B, 1), int(start), dtype=torch.long, device=enc.device)
        return self.decoder.generate(input_ids=ids, encoder_hidden_states=enc, encoder_attention_mask=attention_mask,
                                     generation_config=generation_config, **kwargs)


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
        self.model = VisionEncoderDecoderB200()
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
        # VisionEncoderDecoderConfig fields the reference sets from the tokenizer (RRG_HF.py:93-98)
        dcfg = self.model.decoder.config
        self.model.config = _HFConfig(decoder_start_token_id=(dl.dataset.seq.tokenizer.cls_token_id if dl else dcfg.bos_token_id),
                                      pad_token_id=dcfg.pad_token_id, bos_token_id=dcfg.bos_token_id, eos_token_id=dcfg.eos_token_id,
                                      vocab_size=dcfg.vocab_size)
        self.eval_func = evaluation_multi
        set_arena_root(self)

    def _encode(self, flat_pixels):
        return self.model.project(self.model.encoder(flat_pixels))            # [B', S, D_dec] bf16

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
