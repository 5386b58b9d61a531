"""B200-native VisualEncoder — same name, constructor kwargs and encode()/forward() contract as the reference's
vilmedic/blocks/vision/visual_encoder.py:86-235, with the ViT arithmetic on sm_100a kernels (ViTTower).

Reference behaviour kept (file:line in /root/reference/vilmedic/blocks/vision/visual_encoder.py):
  * "vit" in backbone -> ViTModel(ViTConfig(return_dict=True, **kwargs), add_pooling_layer=False)          (:56-58)
  * forward: ViT -> dropout_out(last_hidden_state), no permute                                              (:180-186)
  * CNN: torchvision network truncated at `output_layer`, batch_first -> [B, HW, C] (squeeze when HW == 1)   (:71-81,196-207)
  * encode: features_mask = (sum_d |f| != 0); visual_projection Linear or Identity                          (:130-139)
  * multi-image 5-D input: flatten -> forward -> x images_mask -> concat along positions                    (:159-178)
Reference defects resolved to the intended semantics (SURVEY.md §8 "Reference defects" #3): `num_images` is read from
the 5-D input (the reference reads it after flattening, i.e. the channel count); `freeze` freezes `self.model`.
"deit" in backbone -> the same tower with HF DeiTModel's distillation token and [N+2] position table (:60-61).
Out of scope here (raise): HF-ResNet / PoolFormer / monai 3-D backbones (not in any BASELINE config).
The torchvision ResNets (ResNet-18/50 of cfg #1/#3; BasicBlock / Bottleneck, groups=1) run on the sm_100a kernels through
vilmedic_b200/cnn.py (im2col + tcgen05 GEMM convolutions, fused BatchNorm/ReLU/residual, hand-written backward); the
parameter tree stays torchvision's, so state_dict keys are the reference's.  Other CNN families (DenseNet, ResNeXt) still
go through torchvision/cuDNN (interim library path) and feed the native decoder kernels.
"""
import json

import torch
import torch.nn as nn

from ... import ops
from ...arena import get_arena
from ...cfgutil import cfg_get
from ...nn import CastBf16Fn, DropoutFn, LinearFn, ViTTower, _lin, _root_of, _prepare, set_arena_root

__all__ = ["VisualEncoder", "get_network"]


def get_network(backbone, output_layer, pretrained, **kwargs):
    if "vit" in backbone.lower():
        return ViTTower(**kwargs)
    if "deit" in backbone.lower():        # DeiTModel(DeiTConfig(**kwargs), add_pooling_layer=False), :60-61: ViT blocks + distillation token
        return ViTTower(distillation=True, **kwargs)
    for tag in ("hfresnet", "hfpoolformer", "3d"):
        if tag in backbone.lower():
            raise NotImplementedError("backbone %r is outside the B200 hot path (SURVEY.md §2 #5)" % backbone)
    import torchvision.models as tvm
    if not hasattr(tvm, backbone):
        raise NotImplementedError("unknown backbone %r" % backbone)
    if pretrained:
        raise NotImplementedError("pretrained torchvision weights need network access; use pretrained=False and load a state_dict")
    if "densenet" in backbone and output_layer == "avgpool":
        sub = get_network(backbone, "features", pretrained, **kwargs)
        sub.add_module("relu", nn.ReLU(inplace=True))
        sub.add_module("avgpool", nn.AdaptiveAvgPool2d((1, 1)))
        sub.add_module("flatten", nn.Flatten(1))
        return sub
    network = getattr(tvm, backbone)(weights=None, **kwargs)
    if output_layer is not None and output_layer != "classifier":
        layers = [n for n, _ in network.named_children()]
        assert output_layer in layers, "{} not in {}".format(output_layer, layers)
        sub = []
        for n, c in network.named_children():
            sub.append(c)
            if n == output_layer:
                break
        network = nn.Sequential(*sub)
    return network


class VisualEncoder(nn.Module):
    def __init__(self, backbone, permute, dropout_out=0.0, freeze=False, output_layer=None, pretrained=True,
                 slice_encode=None, slice_dim=None, visual_projection=None, **kwargs):
        super().__init__()
        self.backbone = backbone
        self.output_layer = output_layer
        self.permute = permute
        self.freeze = freeze
        self.pretrained = pretrained
        self.is_vit = "vit" in backbone.lower() or "deit" in backbone.lower()
        self.model = get_network(self.backbone, self.output_layer, self.pretrained and not self.is_vit, **kwargs)
        self.dropout_out = nn.Dropout(p=dropout_out)
        self._resnet = None
        if not self.is_vit:
            from ...cnn import ResNetRunner
            try:                                  # ResNet-18/34/50/101/152 run on the kernels; other CNN families stay interim
                object.__setattr__(self, "_resnet", ResNetRunner(self.model, self))
            except NotImplementedError:
                object.__setattr__(self, "_resnet", None)
        self.is3D = "3d" in backbone
        self.slice_encode = slice_encode
        self.slice_dim = slice_dim
        if self.slice_encode:
            raise NotImplementedError("slice_encode (3-D volumes) is outside the B200 hot path")
        if visual_projection:
            self.visual_projection = nn.Linear(cfg_get(visual_projection, "in_features"),
                                               cfg_get(visual_projection, "out_features"))
        else:
            self.visual_projection = nn.Identity()
        assert permute in ["batch_first", "spatial_first", "no_permute"]
        if freeze:
            for _, param in self.model.named_parameters():
                param.requires_grad = False
        # stand-alone use (no enclosing model): tower and projection share ONE parameter arena rooted here; a model that embeds this
        # block re-roots every sub-module at itself (set_arena_root in RRG / ConVIRT / MVQA)
        set_arena_root(self)

    # ------------------------------------------------------------------------------------------------ encode
    def _project(self, feats2d):
        """feats2d bf16 [R, D] -> visual_projection(feats) bf16."""
        if isinstance(self.visual_projection, nn.Identity):
            return feats2d
        arena = get_arena(_root_of(self))
        _prepare(arena, self)
        return LinearFn.apply(feats2d, self.visual_projection.weight, _lin(arena, self.visual_projection), torch.bfloat16)

    def encode(self, images, images_mask=None, **kwargs):
        images = images.cuda(non_blocking=True)
        images_mask = images_mask.cuda(non_blocking=True) if images_mask is not None else None
        if images.dim() == 4:
            features = self(images)
            if features.dim() == 2:
                features = features.unsqueeze(1)
                squeeze = True
            else:
                squeeze = False
            B, S, D = features.shape
            features = features.contiguous()
            mask = ops.features_mask(features.detach()).bool()
            out = self._project(features.view(B * S, D)).view(B, S, -1)
            if squeeze:
                return out.squeeze(1), mask.squeeze(1)
            return out, mask
        assert images.dim() == 5, "wrong images shape"
        if self.is3D:
            raise NotImplementedError("3-D encoders are outside the B200 hot path")
        Bn, N = images.shape[0], images.shape[1]
        flat = images.reshape(Bn * N, *images.shape[2:])
        features = self(flat)
        if features.dim() <= 2:
            raise Exception("The input size is too small for this model. The spatial dim has been shrunk to 1.")
        S, D = features.shape[-2], features.shape[-1]
        features = features.reshape(Bn, N * S, D).contiguous()
        if images_mask is not None:
            rows = images_mask.reshape(Bn * N).to(torch.uint8).contiguous()
            features = _MaskRowsFn.apply(features.view(Bn * N * S, D), rows, S).view(Bn, N * S, D)
        mask = ops.features_mask(features.detach()).bool()
        out = self._project(features.view(Bn * N * S, D)).view(Bn, N * S, -1)
        return out, mask

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, images, **kwargs):
        images = images.cuda(non_blocking=True)
        if self.is_vit:
            out = self.model(images)
            p = self.dropout_out.p
            if p > 0 and self.training and torch.is_grad_enabled():
                out = DropoutFn.apply(out.contiguous(), p)
            return out
        if self._resnet is not None:
            return self._forward_resnet(images)
        # interim library path for the remaining CNN families (DenseNet, ResNeXt, ...): torchvision / cuDNN, output handed
        # to the native kernels as bf16
        out = self.model(images.float())
        out = self.dropout_out(out)
        if self.permute == "no_permute":
            pass
        elif self.permute == "batch_first":
            out = out.view(*out.size()[:2], -1).permute(0, 2, 1)
            if out.shape[1] == 1:
                out = out.squeeze(1)
        elif self.permute == "spatial_first":
            out = out.view(*out.size()[:2], -1).permute(2, 0, 1)
        else:
            raise NotImplementedError()
        return CastBf16Fn.apply(out.contiguous())

    def _forward_resnet(self, images):
        """torchvision ResNet on the sm_100a kernels (vilmedic_b200/cnn.py).  The kernels work on [B, H*W, C] (NHWC), which IS
        the reference's `batch_first` layout `out.view(B, C, -1).permute(0, 2, 1)` (:200-203) — no data movement."""
        from ...cnn import resnet_forward
        tap_stage = getattr(self, "tap_stage", None)
        x, (B, H, W, C), pooled = resnet_forward(self._resnet, images, self.training, tap_stage)
        if tap_stage is not None:                 # (bf16 [B*h*w, c], (B, h, w, c)) of torchvision `layer<tap_stage>` — GLoRIA's hook
            object.__setattr__(self, "tapped", self._resnet.last_tap)
        p = self.dropout_out.p
        if p > 0 and self.training and torch.is_grad_enabled():
            pad = (-x.numel()) % 8
            if pad:
                raise NotImplementedError("dropout_out needs numel % 8 == 0")
            x = DropoutFn.apply(x.contiguous(), p)
        hw = 1 if pooled else H * W
        if self.permute == "batch_first":
            return x.view(B, C) if hw == 1 else x.view(B, hw, C)              # squeeze(1) when one position is left (:202-203)
        x4 = x.view(B, 1, 1, C) if pooled else x.view(B, H, W, C)
        if self.permute == "no_permute":
            return x4.permute(0, 3, 1, 2)                                     # NCHW view, as torchvision returns it
        return x4.reshape(B, hw, C).permute(1, 0, 2)                          # spatial_first

    def train(self, mode: bool = True):
        if self.freeze:
            mode = False
        self.training = mode
        for module in self.children():
            module.train(mode)
        return self

    def __repr__(self):
        repr_dict = {
            "type": "ViTTower(sm_100a)" if self.is_vit else ("ResNet(sm_100a)" if self._resnet is not None else None),
            "config": str(self.model.config) if self.is_vit else None,
            "dropout_out": self.dropout_out.p,
            "freeze": self.freeze,
            "output_layer": str(self.output_layer) if self.output_layer is not None else None,
            "pretrained": self.pretrained if not self.is_vit else None,
            "visual_projection": str(self.visual_projection),
        }
        repr_dict = {k: v for k, v in repr_dict.items() if v is not None}
        return f"{self.backbone}:\n{json.dumps(repr_dict, indent=2)}"


class _MaskRowsFn(torch.autograd.Function):
    """features of masked-out images are zeroed (visual_encoder.py:170-171); same op on the gradient."""

    @staticmethod
    def forward(ctx, x, rows, rows_per_mask):
        ctx.saved = (rows, rows_per_mask)
        return ops.mask_rows(x, rows, rows_per_mask)

    @staticmethod
    def backward(ctx, dy):
        rows, rpm = ctx.saved
        return ops.mask_rows(dy.contiguous(), rows, rpm), None, None
