"""Oracle compositions for the other models on the hot path (ConVIRT, MVQA, RRG_HF) — HF / torchvision modules wired as
the reference wires them.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  ConVIRT  vilmedic/models/selfsup/conVIRT.py:46-102 ; EncoderModel vilmedic/blocks/huggingface/encoder/encoder_model.py:16-62
  MVQA     vilmedic/models/mvqa/MVQA.py:14-54 ; Classifier vilmedic/blocks/classifier/classifier.py:4-15
  RRG_HF   vilmedic/models/rrg/RRG_HF.py:108-177 (VisionEncoderDecoderModel of ViTModel + BertGenerationDecoder)
"""
import torch
import torch.nn as nn
from transformers import BertGenerationConfig, BertGenerationDecoder, BertGenerationEncoder, ViTConfig, ViTModel
from transformers.models.bert.modeling_bert import BertEncoder, BertPooler

from .losses import convirt_loss, infonce_loss, label_smoothing_ce
from .rrg import OracleVisualEncoder


class OracleEncoderModel(nn.Module):
    def __init__(self, encoder):
        super().__init__()
        enc = dict(encoder)
        enc.pop("proto", None)
        add_pool = enc.pop("add_pooling_layer", False)
        cfg = BertGenerationConfig(**enc, is_decoder=False, add_cross_attention=False)           # encoder_model.py:23-25
        self.encoder = BertGenerationEncoder(cfg)                                                # :26
        self.encoder.config._attn_implementation = "eager"
        if add_pool:
            self.pooler = BertPooler(cfg)                                                        # :28-29

    def forward(self, input_ids, attention_mask):
        out = self.encoder(input_ids=input_ids, attention_mask=attention_mask, return_dict=True)
        res = {"last_hidden_state": out.last_hidden_state, "pooler_output": None}
        if hasattr(self, "pooler"):
            res["pooler_output"] = self.pooler(hidden_states=out.last_hidden_state)             # :58-60
        return res


class OracleConVIRT(nn.Module):
    def __init__(self, encoder, cnn, projection, loss):
        super().__init__()
        cnn = dict(cnn)
        cnn.pop("proto", None)
        cnn.pop("pretrained", None)
        self.linguistic = OracleEncoderModel(encoder)                                            # conVIRT.py:52
        self.visual = OracleVisualEncoder(**cnn)                                                 # :55
        p = projection
        self.vis_proj = nn.Sequential(nn.Linear(p["visual_embedding_dim"], p["projection_dim"]), nn.ReLU(),
                                      nn.Linear(p["projection_dim"], p["projection_dim"]))       # :58-62
        self.lin_proj = nn.Sequential(nn.Linear(p["textual_embedding_dim"], p["projection_dim"]), nn.ReLU(),
                                      nn.Linear(p["projection_dim"], p["projection_dim"]))       # :63-67
        self.loss = dict(loss)

    def forward(self, input_ids, attention_mask, images):
        l = self.lin_proj(self.linguistic(input_ids, attention_mask)["pooler_output"])           # :88-91
        v = self.vis_proj(self.visual(images))                                                   # :92
        if self.loss["proto"] == "ConVIRTLoss":
            loss, a, b = convirt_loss(l, v, self.loss["tau"], self.loss["lambda_"])             # :100
        else:
            loss, a, b = infonce_loss(l, v)
        return {"loss": loss, "loss_l": a, "loss_v": b, "linguistic": l, "visual": v}


class OracleMVQA(nn.Module):
    def __init__(self, cnn, classifier, adapter, transformer, loss):
        super().__init__()
        cnn = dict(cnn)
        cnn.pop("proto", None)
        self.cnn = OracleVisualEncoder(**cnn)                                                    # MVQA.py:22
        self.adapter = nn.Sequential(nn.Linear(adapter["input_size"], adapter["output_size"]),
                                     nn.LayerNorm(transformer["hidden_size"], eps=transformer["layer_norm_eps"]))  # :23-26
        conf = BertGenerationConfig(**transformer)                                               # :28
        conf._attn_implementation = "eager"
        self.transformer = BertEncoder(conf)                                                     # :29
        self.pooler = BertPooler(conf)                                                           # :30
        self.classifier = nn.Sequential()
        self.classifier.classifier = nn.Sequential(nn.Linear(classifier["input_size"], classifier["num_classes"]))  # classifier.py:7-9
        self.smoothing = loss.get("smoothing", 0.1)

    def forward(self, images, labels):
        out = self.cnn(images)                                                                   # :41 (forward, not encode)
        out = self.adapter(out)                                                                  # :42
        out = self.transformer(out, output_attentions=True)                                      # :43
        out = self.pooler(out.last_hidden_state)                                                 # :47
        out = self.classifier.classifier(out)                                                    # :48
        loss = label_smoothing_ce(out, labels, self.smoothing)                                   # :52
        return {"loss": loss, "output": out, "answer": torch.argmax(out, dim=-1)}


class OracleRRGHF(nn.Module):
    def __init__(self, vision_args, decoder_args):
        super().__init__()
        self.model = nn.Module()
        self.model.encoder = ViTModel(ViTConfig(**vision_args))                                  # RRG_HF.py:38-39 (pooler kept)
        self.model.encoder.config._attn_implementation = "eager"
        d = dict(decoder_args, is_decoder=True, add_cross_attention=True)                        # :76-77
        self.model.decoder = BertGenerationDecoder(BertGenerationConfig(**d))                    # :79-80
        self.model.decoder.config._attn_implementation = "eager"

    def forward(self, input_ids, attention_mask, images, images_mask=None):
        if images.dim() == 5:                                                                    # :124-152
            B, N, C, H, W = images.shape
            mask = torch.ones((B, N), dtype=torch.bool) if images_mask is None else images_mask.bool()
            h = self.model.encoder(pixel_values=images.view(B * N, C, H, W)).last_hidden_state
            S, D = h.shape[1], h.shape[2]
            enc = h.view(B, N * S, D)
            am = mask.unsqueeze(-1).expand(B, N, S).reshape(B, N * S).long()
            out = self.model.decoder(input_ids=input_ids, attention_mask=attention_mask, encoder_hidden_states=enc,
                                     encoder_attention_mask=am, labels=input_ids, use_cache=False)
        else:                                                                                    # :155-172
            enc = self.model.encoder(pixel_values=images).last_hidden_state
            out = self.model.decoder(input_ids=input_ids, attention_mask=attention_mask, encoder_hidden_states=enc,
                                     encoder_attention_mask=None, labels=input_ids, use_cache=False)
        return vars(out)
