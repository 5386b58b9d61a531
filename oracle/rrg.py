"""Oracle for the RRG path: HF ViTModel + BertGenerationDecoder composed as the reference composes them.

  get_network ViT branch ........ vilmedic/blocks/vision/visual_encoder.py:56-58
  VisualEncoder.forward/encode .. vilmedic/blocks/vision/visual_encoder.py:130-139,180-186
  DecoderModel .................. vilmedic/blocks/huggingface/decoder/decoder_model.py:14-49 (labels=input_ids :46)
  RRG.forward ................... vilmedic/models/rrg/RRG.py:25-41
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import torch
import torch.nn as nn
from transformers import BertGenerationConfig, BertGenerationDecoder, ViTConfig, ViTModel


class OracleVisualEncoder(nn.Module):
    def __init__(self, backbone="vit", permute="no_permute", dropout_out=0.0, visual_projection=None, output_layer=None,
                 **kwargs):
        super().__init__()
        self.backbone = backbone
        self.permute = permute
        if "vit" in backbone.lower():
            self.model = ViTModel(ViTConfig(return_dict=True, **kwargs), add_pooling_layer=False)   # :56-58
            self.model.config._attn_implementation = "eager"
        elif "deit" in backbone.lower():
            from transformers import DeiTConfig, DeiTModel
            self.model = DeiTModel(DeiTConfig(return_dict=True, **kwargs), add_pooling_layer=False)  # :60-61
            self.model.config._attn_implementation = "eager"
        else:
            import torchvision.models as tvm
            network = getattr(tvm, backbone)(weights=None, **kwargs)                                # :71
            if output_layer is not None and output_layer != "classifier":                           # :73-81
                sub = []
                for n, c in network.named_children():
                    sub.append(c)
                    if n == output_layer:
                        break
                network = nn.Sequential(*sub)
            self.model = network
        self.dropout_out = nn.Dropout(p=dropout_out)
        if visual_projection:
            self.visual_projection = nn.Linear(visual_projection["in_features"], visual_projection["out_features"])
        else:
            self.visual_projection = nn.Identity()                                                   # :119-122

    def forward(self, images):
        out = self.model(images)
        if hasattr(out, "last_hidden_state"):                                                        # ViTModel / DeiTModel
            return self.dropout_out(out.last_hidden_state)                                           # :182-186
        out = self.dropout_out(out)
        if self.permute == "batch_first":                                                           # :200-203
            out = out.view(*out.size()[:2], -1).permute(0, 2, 1)
            if out.shape[1] == 1:
                out = out.squeeze(1)
        return out

    def encode(self, images, images_mask=None):
        if images.dim() == 4:
            features = self(images)
            features_mask = (torch.sum(torch.abs(features), dim=-1) != 0)                           # :138
            return self.visual_projection(features), features_mask                                   # :139
        # multi-image (:159-178) with the intended num_images (SURVEY.md §8 defects #3)
        B, N = images.shape[:2]
        feats = self(images.reshape(B * N, *images.shape[2:]))
        feats = feats.view(B, N, feats.shape[-2], feats.shape[-1])
        if images_mask is not None:
            feats = feats * images_mask.unsqueeze(-1).unsqueeze(-1)
        feats = feats.reshape(B, N * feats.shape[-2], feats.shape[-1])
        features_mask = (torch.sum(torch.abs(feats), dim=-1) != 0)
        return self.visual_projection(feats), features_mask


class OracleDecoderModel(nn.Module):
    def __init__(self, decoder):
        super().__init__()
        dec_config = BertGenerationConfig(**decoder)                                                 # :23
        dec_config.is_decoder = True                                                                 # :24
        dec_config.add_cross_attention = True                                                        # :25
        self.decoder = BertGenerationDecoder(dec_config)                                             # :26
        self.decoder.config._attn_implementation = "eager"
        self.config = self.decoder.config

    def forward(self, input_ids, attention_mask, encoder_outputs=None, encoder_attention_mask=None):
        out = self.decoder(input_ids=input_ids, attention_mask=attention_mask, encoder_hidden_states=encoder_outputs,
                           encoder_attention_mask=encoder_attention_mask, labels=input_ids, use_cache=False)  # :42-47
        return vars(out)                                                                             # :48


class OracleRRG(nn.Module):
    def __init__(self, decoder, cnn):
        super().__init__()
        cnn = dict(cnn)
        cnn.pop("proto", None)
        self.dec = OracleDecoderModel(dict(decoder))                                                 # RRG.py:17
        self.enc = OracleVisualEncoder(**cnn)                                                        # RRG.py:20

    def forward(self, input_ids, attention_mask, images, images_mask=None):
        encoder_outputs, encoder_attention_mask = self.enc.encode(images, images_mask)               # RRG.py:32-33
        return self.dec(input_ids=input_ids, attention_mask=attention_mask, encoder_outputs=encoder_outputs,
                        encoder_attention_mask=encoder_attention_mask)                               # RRG.py:35-39
