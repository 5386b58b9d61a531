"""torchvision ResNet backbones (BasicBlock / Bottleneck, groups = 1) of the reference's CNN branch
(vilmedic/blocks/vision/visual_encoder.py:71-83: `eval(backbone)(pretrained=...)` truncated at `output_layer`) executed on
the sm_100a kernels — SURVEY.md §8 a3 / K18.

The parameter tree is torchvision's own module tree (so `state_dict()` keys / shapes / BatchNorm buffers are exactly the
reference's, including the `nn.Sequential` indices that the truncation produces); this module only *runs* it:

  * activations are NHWC bf16 matrices [B*H*W, C]; every convolution is one tcgen05 GEMM  col [M, kh*kw*Cin] x Wm^T
    (`ops.gemm`), with `col` gathered by `vlm_im2col_nhwc` (1x1 stride-1 convolutions read the activation matrix directly,
    the 7x7 stem gathers straight from the fp32 NCHW images) and Wm packed from the OIHW fp32 master every step;
  * BatchNorm runs in training mode on batch statistics (column sums), fused with the residual add and the ReLU, and updates
    `running_mean / running_var / num_batches_tracked` like torch; evaluation mode uses the running statistics;
  * the backward is written by hand: BN backward (+ReLU mask, + residual-branch gradient), wgrad GEMM (dY^T col, fp32,
    un-packed into the OIHW gradient), dgrad GEMM + col2im gather with the skip-connection gradient added in the same pass.

One `torch.autograd.Function` runs the whole backbone and keeps its own tape — no torch arithmetic on the data path.
"""
import torch
import torch.nn as nn

from . import ops
from .arena import get_arena
from .nn import _prepare, _root_of


def _ceil8(n):
    return (n + 7) // 8 * 8


class _Unit:
    """conv -> BatchNorm (-> + residual) (-> ReLU) with everything its backward needs."""

    def __init__(self, arena, conv, bn):
        if conv.groups != 1 or conv.dilation != (1, 1) or conv.bias is not None or conv.kernel_size[0] != conv.kernel_size[1] \
                or conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1] or conv.padding_mode != "zeros":
            raise NotImplementedError("convolution %r has no sm_100a path (square kernel, groups=1, dilation=1, no bias)" % (conv,))
        if bn.momentum is None or not bn.affine or not bn.track_running_stats:
            raise NotImplementedError("BatchNorm2d needs affine=True, track_running_stats=True and a numeric momentum")
        self.arena, self.conv, self.bn = arena, conv, bn
        self.k, self.s, self.p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        self.cin, self.cout = conv.in_channels, conv.out_channels
        self.direct = self.k == 1 and self.s == 1 and self.p == 0
        self.Kp = _ceil8(self.k * self.k * self.cin)

    def forward(self, x, shp, relu, training, res=None, stem_img=None, save=False):
        """x bf16 [B*H*W, Cin] (or stem_img fp32 NCHW) -> (y bf16 [B*Ho*Wo, Cout], (B, Ho, Wo, Cout))."""
        a, conv, bn = self.arena, self.conv, self.bn
        B, H, W, C = shp
        Ho, Wo = ops.conv_out_size(H, self.k, self.s, self.p), ops.conv_out_size(W, self.k, self.s, self.p)
        wm = ops.conv_weight_pack(a.fp32(conv.weight), self.Kp)
        if stem_img is not None:
            col = ops.im2col_nchw_f32(stem_img, self.k, self.k, self.s, self.p, self.Kp)
        elif self.direct:
            col = x
        else:
            col = ops.im2col_nhwc(x, B, H, W, C, self.k, self.k, self.s, self.p)
        z = ops.gemm(col, wm)
        if training:
            y, mean, rstd = ops.bn_train_fwd(z, a.fp32(bn.weight), a.fp32(bn.bias), bn.running_mean, bn.running_var,
                                             bn.num_batches_tracked, bn.eps, bn.momentum, relu, res)
        else:
            y, mean, rstd = ops.bn_eval_fwd(z, a.fp32(bn.weight), a.fp32(bn.bias), bn.running_mean, bn.running_var, bn.eps, relu, res), None, None
        if save:
            self.saved = (col, wm, z, y, mean, rstd, relu, shp, res is not None)
        return y, (B, Ho, Wo, self.cout)

    def backward(self, dy, need_dx=True, add=None):
        """dy bf16 [M, Cout] -> (dx bf16 [B*H*W, Cin] (+ add) | None, dres | None)."""
        a, conv, bn = self.arena, self.conv, self.bn
        col, wm, z, y, mean, rstd, relu, shp, has_res = self.saved
        self.saved = None
        B, H, W, C = shp
        dz, dres = ops.bn_train_bwd(dy, y, z, mean, rstd, a.fp32(bn.weight), a.grad(bn.weight), a.grad(bn.bias), relu, has_res)
        if conv.weight.requires_grad:
            # [Cout, Kp] = dZ^T col; `accumulate` into a zeroed fp32 buffer so that the GEMM may split K (= B*Ho*Wo rows)
            dwm = torch.zeros((self.cout, self.Kp), device=dz.device, dtype=torch.float32)
            ops.gemm(dz, col, a_mn_major=True, b_mn_major=True, out=dwm, accumulate=True)
            ops.conv_wgrad_unpack(dwm, a.grad(conv.weight))
        dx = None
        if need_dx:
            if self.direct:
                dx = ops.gemm(dz, wm, b_mn_major=True, residual=add)                                  # [M, Cin] (+ skip gradient)
            else:
                dcol = ops.gemm(dz, wm, b_mn_major=True)                                              # [M, k*k*Cin]
                dx = ops.col2im_nhwc(dcol.contiguous(), B, H, W, C, self.k, self.k, self.s, self.p, add=add)
        return dx, dres


class _Block:
    """torchvision BasicBlock / Bottleneck: units chained, identity (or downsample) added before the last ReLU."""

    def __init__(self, arena, block):
        names = [n for n in ("conv1", "conv2", "conv3") if hasattr(block, n)]
        self.units = [_Unit(arena, getattr(block, n), getattr(block, "bn" + n[-1])) for n in names]
        self.down = None
        if block.downsample is not None:
            ds = list(block.downsample.children())
            if len(ds) != 2 or not isinstance(ds[0], nn.Conv2d) or not isinstance(ds[1], nn.BatchNorm2d):
                raise NotImplementedError("downsample branch must be Conv2d + BatchNorm2d")
            self.down = _Unit(arena, ds[0], ds[1])

    def forward(self, x, shp, training, save):
        identity = x
        if self.down is not None:
            identity, _ = self.down.forward(x, shp, relu=False, training=training, save=save)
        out, s = x, shp
        for u in self.units[:-1]:
            out, s = u.forward(out, s, relu=True, training=training, save=save)
        return self.units[-1].forward(out, s, relu=True, training=training, res=identity, save=save)

    def backward(self, dy):
        d, d_identity = self.units[-1].backward(dy)
        add = d_identity
        if self.down is not None:
            add, _ = self.down.backward(d_identity)
        for u in reversed(self.units[1:-1]):
            d, _ = u.backward(d)
        dx, _ = self.units[0].backward(d, add=add)
        return dx


class ResNetRunner:
    """Execution plan over a torchvision ResNet (or the nn.Sequential of its leading children the reference builds)."""

    def __init__(self, model, owner):
        import torchvision.models.resnet as tvr
        children = [c for _, c in model.named_children()]
        kinds = [type(c) for c in children]
        if len(children) < 5 or kinds[:4] != [nn.Conv2d, nn.BatchNorm2d, nn.ReLU, nn.MaxPool2d]:
            raise NotImplementedError("not a torchvision ResNet stem (conv1, bn1, relu, maxpool, layer1, ...)")
        mp = children[3]
        if (mp.kernel_size, mp.stride, mp.padding, mp.dilation, mp.ceil_mode) != (3, 2, 1, 1, False):
            raise NotImplementedError("stem max-pool must be MaxPool2d(3, 2, 1)")
        self.model, self.owner = model, owner
        self.layers, self.avgpool = [], False
        for c in children[4:]:
            if isinstance(c, nn.Sequential) and all(isinstance(b, (tvr.BasicBlock, tvr.Bottleneck)) for b in c):
                if self.avgpool:
                    raise NotImplementedError("residual stage after the average pool")
                self.layers.append(list(c))
            elif isinstance(c, nn.AdaptiveAvgPool2d) and tuple(c.output_size if isinstance(c.output_size, tuple) else (c.output_size,) * 2) == (1, 1):
                self.avgpool = True
            else:
                raise NotImplementedError("ResNet child %s is outside the B200 path (use output_layer=layerN or avgpool)" % type(c).__name__)
        self.stem_conv, self.stem_bn = children[0], children[1]
        _Unit(None, self.stem_conv, self.stem_bn)                  # validate every conv / BN now (raises NotImplementedError)
        for layer in self.layers:
            for b in layer:
                _Block(None, b)

    def _build(self):
        arena = get_arena(_root_of(self.owner))
        _prepare(arena, self.owner)
        stem = _Unit(arena, self.stem_conv, self.stem_bn)
        blocks = [_Block(arena, b) for layer in self.layers for b in layer]
        return stem, blocks

    def run(self, images, training, save, tap_stage=None):
        """images fp32 [B,3,H,W] (CUDA) -> (features bf16 [B*Ho*Wo, C] or [B, C] after the average pool, shape, tape).
        tap_stage = n: the output of residual stage n (1-based, torchvision `layer<n>`) is returned as tape/aux (GLoRIA reads layer3
        through a forward hook, vilmedic/models/selfsup/GLoRIA.py:72-79)."""
        stem, blocks = self._build()
        B, Cin, H, W = images.shape
        if Cin != self.stem_conv.in_channels:
            raise RuntimeError("expected input with %d channels, got %d" % (self.stem_conv.in_channels, Cin))
        x, shp = stem.forward(None, (B, H, W, Cin), relu=True, training=training, stem_img=images, save=save)
        pool_in = shp
        x, idx = ops.maxpool3x3s2_fwd(x, *shp)
        shp = (B, ops.conv_out_size(shp[1], 3, 2, 1), ops.conv_out_size(shp[2], 3, 2, 1), shp[3])
        tap, tap_shp, tap_idx = None, None, -1
        if tap_stage is not None:
            tap_idx = sum(len(l) for l in self.layers[:tap_stage]) - 1          # index of the last block of that stage
        for bi, blk in enumerate(blocks):
            x, shp = blk.forward(x, shp, training, save)
            if bi == tap_idx:
                tap, tap_shp = x, shp
        if self.avgpool:
            x = ops.avgpool_fwd(x, B, shp[1] * shp[2], shp[3])
        tape = (stem, blocks, idx, pool_in, shp, tap_idx) if save else None
        self.last_tap = (tap, tap_shp)
        return x, shp, tape

    def backward(self, tape, dy, dtap=None):
        stem, blocks, idx, pool_in, shp, tap_idx = tape
        B = shp[0]
        d = dy.contiguous()
        if self.avgpool:
            d = ops.avgpool_bwd(d, B, shp[1] * shp[2], shp[3])
        for bi in range(len(blocks) - 1, -1, -1):
            if bi == tap_idx and dtap is not None:
                d = d + dtap.to(d.dtype).reshape(d.shape)       # gradient of the tapped activation joins the main path
            d = blocks[bi].backward(d)
        d = ops.maxpool3x3s2_bwd(d, idx, *pool_in)
        stem.backward(d, need_dx=False)


class ResNetFn(torch.autograd.Function):
    """Whole backbone as one autograd node (the images need no gradient; `anchor` is a parameter that does)."""

    @staticmethod
    def forward(ctx, images, anchor, runner, tap_stage=None):
        x, shp, tape = runner.run(images, training=True, save=True, tap_stage=tap_stage)
        ctx.runner, ctx.tape = runner, tape
        ctx.out_shape = shp
        if tap_stage is None:
            return x
        return x, runner.last_tap[0]

    @staticmethod
    def backward(ctx, dy, dtap=None):
        ctx.runner.backward(ctx.tape, dy, dtap)
        ctx.tape = None
        return None, None, None, None


def resnet_forward(runner, images, training, tap_stage=None):
    """-> (features bf16, (B, Ho, Wo, C), pooled: bool); with tap_stage also runner.last_tap = (bf16 [B*h*w, c], (B, h, w, c))."""
    images = images.contiguous().float()
    if training and torch.is_grad_enabled() and any(p.requires_grad for p in runner.model.parameters()):
        if tap_stage is None:
            x = ResNetFn.apply(images, runner.stem_bn.weight, runner)
        else:
            x, tap = ResNetFn.apply(images, runner.stem_bn.weight, runner, tap_stage)
            runner.last_tap = (tap, runner.last_tap[1])          # the autograd-connected tensor
        B, H, W = images.shape[0], images.shape[2], images.shape[3]
        shp = _out_shape(runner, B, H, W)
    else:
        with torch.no_grad():      # evaluation, or train mode without autograd (BatchNorm still uses / updates batch statistics)
            x, shp, _ = runner.run(images, training=training, save=False, tap_stage=tap_stage)
    return x, shp, runner.avgpool


def _out_shape(runner, B, H, W):
    sc = runner.stem_conv
    h, w = ops.conv_out_size(H, sc.kernel_size[0], sc.stride[0], sc.padding[0]), ops.conv_out_size(W, sc.kernel_size[0], sc.stride[0], sc.padding[0])
    h, w = ops.conv_out_size(h, 3, 2, 1), ops.conv_out_size(w, 3, 2, 1)
    c = runner.stem_conv.out_channels
    for layer in runner.layers:
        for b in layer:
            convs = [getattr(b, n) for n in ("conv1", "conv2", "conv3") if hasattr(b, n)]
            for cv in convs:
                h, w = ops.conv_out_size(h, cv.kernel_size[0], cv.stride[0], cv.padding[0]), ops.conv_out_size(w, cv.kernel_size[0], cv.stride[0], cv.padding[0])
            c = convs[-1].out_channels
    return (B, h, w, c)
