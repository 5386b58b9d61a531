"""Transformer towers of the hot path, built from the sm_100a kernels (no torch math on the data path).

* `ViTTower`  — HF ViTModel-compatible parameter tree (state_dict keys identical to
  transformers.models.vit.modeling_vit.ViTModel(add_pooling_layer=False)); pre-LN blocks.
* `BertTower` — HF BertGenerationEncoder/Decoder / BertModel-style post-LN blocks, optional causal self-attention and
  cross-attention, optional tied LM head with the fused shifted cross-entropy.

Every layer is one torch.autograd.Function whose backward is written by hand from the same kernels; weight
gradients are accumulated directly into the arena's flat fp32 gradient buffer (p.grad are views of it).
"""
import math

import torch
import torch.nn as nn

from . import ops
from .arena import get_arena

_RNG = {"seed": None, "offset": 0}

# Data-parallel hook (ddp.GradSync.attach): called by the hand-written backward of a layer / embedding block once every gradient
# of that module's parameter span is final, so that the exchange of the span can start under the rest of the backward.
GRAD_READY_HOOK = [None]


def notify_grad_ready(module):
    ops.SIDE.join()                     # bias-gradient column sums / weight-gradient GEMMs issued on the side streams
    ops.SIDE_GEMM.join()                # (ops.SideWork) are part of the span
    hook = GRAD_READY_HOOK[0]
    if hook is not None and module is not None:
        hook(module)


def _next_rng():
    if _RNG["seed"] is None:
        _RNG["seed"] = int(torch.initial_seed()) & ((1 << 62) - 1)
    _RNG["offset"] += 1
    return _RNG["seed"], _RNG["offset"]


def manual_seed(seed):
    _RNG["seed"] = int(seed) & ((1 << 62) - 1)
    _RNG["offset"] = 0


class _P:
    """Bundle of arena views for one Linear: bf16 weight, fp32 bias, fp32 grads."""

    def __init__(self, arena, weights, biases, out_features, in_features):
        self.w = arena.bf16(*weights, shape=(out_features, in_features))
        self.b = arena.fp32(*biases, shape=(out_features,)) if biases[0] is not None else None
        self.gw = arena.grad(*weights, shape=(out_features, in_features))
        self.gb = arena.grad(*biases, shape=(out_features,)) if biases[0] is not None else None


def _lin(arena, *linears):
    """Fused view over one or more adjacent nn.Linear modules (stacked along out_features)."""
    ws = [l.weight for l in linears]
    bs = [l.bias for l in linears]
    out = sum(l.weight.shape[0] for l in linears)
    return _P(arena, ws, bs, out, linears[0].weight.shape[1])


def _ln(arena, ln):
    return (arena.fp32(ln.weight), arena.fp32(ln.bias), arena.grad(ln.weight), arena.grad(ln.bias))


def _wgrad(dy, x, P, scale_t=None, bias=True, side=False):
    """P.gw += dy^T x ; P.gb += colsum(dy) (skipped when the producer of dy already accumulated it).  dy [M,N], x [M,K].
    side=True (layer backward functions, which end in notify_grad_ready -> join): the column sums run on the side stream under the
    wgrad / dgrad GEMMs that consume the same dy."""
    if bias and P.gb is not None:
        if side:
            ops.SIDE.run(lambda: ops.colsum(dy, P.gb, scale_t), dy, scale_t)
        else:
            ops.colsum(dy, P.gb, scale_t)
    if side:
        ops.SIDE_GEMM.run(lambda: ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=P.gw, accumulate=True, alpha_t=scale_t), dy, x, scale_t)
    else:
        ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=P.gw, accumulate=True, alpha_t=scale_t)


def _dgrad(dy, P, **kw):
    """dy [M,N] x W [N,K] -> [M,K] (W consumed MN-major, no transposed copy)."""
    return ops.gemm(dy, P.w, b_mn_major=True, **kw)


# =============================================================================================== generic pieces
class LinearFn(torch.autograd.Function):
    """y = x W^T + b on [M,K] bf16 rows (visual_projection, adapters, poolers, classifier heads)."""

    @staticmethod
    def forward(ctx, x, anchor, P, out_dtype):
        y = ops.gemm(x, P.w, bias=P.b, out_dtype=out_dtype)
        ctx.P = P
        ctx.x = x
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _bf16_rows(dy)
        _wgrad(dy, ctx.x, ctx.P)
        dx = _dgrad(dy, ctx.P)
        return dx, None, None, None


class LinearGeluFn(torch.autograd.Function):
    """y = GELU(x W^T + b) with GELU' stashed by the GEMM epilogue (the LM-head transform of BERT / RoBERTa checkpoints:
    HF:bert/modeling_bert.py BertPredictionHeadTransform, HF:roberta/modeling_roberta.py RobertaLMHead)."""

    @staticmethod
    def forward(ctx, x, anchor, P):
        # (grad mode is off inside Function.forward: ask the context whether a backward can follow)
        stash = torch.empty((x.shape[0], P.w.shape[0]), device=x.device, dtype=torch.bfloat16) if any(ctx.needs_input_grad) else None
        y = ops.gemm(x, P.w, bias=P.b, act=ops.ACT_GELU, aux_out=stash)
        ctx.saved = (x, stash, P)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stash, P = ctx.saved
        dy = dy.contiguous()
        # dpre = dy * GELU'(pre): the x-stash epilogue needs a GEMM in front of it; here the product is elementwise on [M, D] only
        dpre = (dy.float() * stash.float()).to(torch.bfloat16)
        _wgrad(dpre, x, P)
        return _dgrad(dpre, P), None, None


def _bf16_rows(t):
    """2-D bf16 tensor whose row pitch is a multiple of 8 elements (TMA needs 16-byte strides); zero padded."""
    N = t.shape[-1]
    if t.dtype == torch.bfloat16 and t.stride(-1) == 1 and t.stride(-2) % 8 == 0:
        return t
    if N % 8 == 0:
        t = t.contiguous()
        return t if t.dtype == torch.bfloat16 else ops.cast_bf16(t.float().contiguous())
    buf = torch.zeros((t.shape[0], (N + 7) // 8 * 8), device=t.device, dtype=torch.bfloat16)
    buf[:, :N] = t
    return buf[:, :N]


class LayerNormFn(torch.autograd.Function):
    """colsum: optional fp32 [D] gradient buffer of the bias of the Linear that produced x (its bias gradient is the
    column sum of dx, accumulated inside the LN backward kernel)."""

    @staticmethod
    def forward(ctx, x, anchor, lnp, eps, colsum=None):
        g, b, gg, gb = lnp
        y, mean, rstd = ops.layernorm_fwd(x, g, b, eps)
        ctx.saved = (x, mean, rstd, lnp, colsum)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, (g, b, gg, gb), colsum = ctx.saved
        dx = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, g, gg, gb, colsum=colsum)
        return dx, None, None, None, None


class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        seed, off = _next_rng()
        ctx.rng = (p, seed, off)
        return ops.dropout(x, p, seed, off)

    @staticmethod
    def backward(ctx, dy):
        p, seed, off = ctx.rng
        return ops.dropout(dy.contiguous(), p, seed, off), None


class TanhFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.act_fwd(x.float().contiguous(), 0)
        ctx.y = y
        return y

    @staticmethod
    def backward(ctx, dy):
        return ops.act_bwd(dy.float().contiguous(), ctx.y, 0)


class ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.act_fwd(x.float().contiguous(), 1)
        ctx.y = y
        return y

    @staticmethod
    def backward(ctx, dy):
        return ops.act_bwd(dy.float().contiguous(), ctx.y, 1)


def dropout(x, p, training):
    if p <= 0.0 or not training:
        return x
    return DropoutFn.apply(x, p)


# =============================================================================================== ViT
class ViTEmbedFn(torch.autograd.Function):
    """patchify -> patch GEMM (+bias +pos, per-image batched) -> CLS row.  HF modeling_vit.py:43-128."""

    @staticmethod
    def forward(ctx, images, anchor, mod, arena):
        cfg = mod.cfg
        B = images.shape[0]
        dist = getattr(mod.embeddings, "distillation_token", None)
        patches = ops.patchify(images, cfg.patch_size, 2 if dist is not None else 1)   # [B, S, C*P*P], token rows zero
        S, Kp = patches.shape[1], patches.shape[2]
        D = cfg.hidden_size
        proj = mod.embeddings.patch_embeddings.projection
        w = arena.bf16(proj.weight, shape=(D, Kp))
        bias = arena.fp32(proj.bias)
        pos = arena.fp32(mod.embeddings.position_embeddings, shape=(S, D))
        cls = arena.fp32(mod.embeddings.cls_token, shape=(D,))
        # per-image batch so that the position table is the (shared) residual operand
        pos16 = ops.cast_bf16(pos)
        x = torch.empty((B, S, D), device=images.device, dtype=torch.bfloat16)
        ops.gemm(patches, w.unsqueeze(0).expand(B, D, Kp), out=x, bias=bias, residual=pos16)
        ops.vit_cls_pos(x, cls, pos)
        if dist is not None:                                      # DeiT: row 1 = distillation token + its position (HF:deit/modeling_deit.py)
            ops.vit_cls_pos(x, arena.fp32(dist, shape=(D,)), pos, row=1)
        ctx.saved = (patches, mod, arena)
        return x.view(B * S, D)

    @staticmethod
    def backward(ctx, dx):
        patches, mod, arena = ctx.saved
        B, S, Kp = patches.shape
        D = mod.cfg.hidden_size
        dx = dx.contiguous()
        proj = mod.embeddings.patch_embeddings.projection
        gw = arena.grad(proj.weight, shape=(D, Kp))
        ops.gemm(dx, patches.view(B * S, Kp), a_mn_major=True, b_mn_major=True, out=gw, accumulate=True)
        dist = getattr(mod.embeddings, "distillation_token", None)
        ops.vit_embed_bwd(dx.view(B, S, D), arena.grad(mod.embeddings.position_embeddings, shape=(S, D)),
                          arena.grad(mod.embeddings.cls_token, shape=(D,)), arena.grad(proj.bias),
                          ddist=arena.grad(dist, shape=(D,)) if dist is not None else None)
        notify_grad_ready(mod.embeddings)
        return None, None, None, None


class ViTLayerFn(torch.autograd.Function):
    """One pre-LN ViT block (HF modeling_vit.py:315-346) on a [B*S, D] bf16 residual stream."""

    @staticmethod
    def forward(ctx, x, anchor, layer, arena, B, S, save, prev_gb=None, own_b2_fused=False):
        """prev_gb: gradient buffer of the bias of the Linear that produced x (previous layer's FFN-down): this layer's
        backward accumulates colsum(dx) into it.  own_b2_fused: the consumer of x2's gradient does the same for P2.gb."""
        cfg = layer.cfg
        D, H = cfg.hidden_size, cfg.num_attention_heads
        DH = D // H
        att = layer.attention.attention
        Pqkv = _lin(arena, att.query, att.key, att.value)
        Po = _lin(arena, layer.attention.output.dense)
        P1 = _lin(arena, layer.intermediate.dense)
        P2 = _lin(arena, layer.output.dense)
        ln1, ln2 = _ln(arena, layer.layernorm_before), _ln(arena, layer.layernorm_after)
        eps = cfg.layer_norm_eps
        h1, mean1, rstd1 = ops.layernorm_fwd(x, ln1[0], ln1[1], eps)
        qkv = ops.gemm(h1, Pqkv.w, bias=Pqkv.b)
        q3 = qkv.view(B, S, 3 * D)
        ctxv, lse = ops.attention_fwd(q3[:, :, :D], q3[:, :, D:2 * D], q3[:, :, 2 * D:], H, DH)
        x1 = ops.gemm(ctxv.view(B * S, D), Po.w, bias=Po.b, residual=x)
        h2, mean2, rstd2 = ops.layernorm_fwd(x1, ln2[0], ln2[1], eps)
        pre = torch.empty((B * S, cfg.intermediate_size), device=x.device, dtype=torch.bfloat16) if save else None
        hmid = ops.gemm(h2, P1.w, bias=P1.b, act=ops.ACT_GELU, aux_out=pre)   # pre <- GELU'(pre-activation)
        x2 = ops.gemm(hmid, P2.w, bias=P2.b, residual=x1)
        if save:
            ctx.saved = (x, h1, mean1, rstd1, qkv, ctxv, lse, x1, h2, mean2, rstd2, pre, hmid, Pqkv, Po, P1, P2, ln1, ln2, B, S, H, DH)
            ctx.fuse = (prev_gb, own_b2_fused)
            ctx.layer = layer
        return x2

    @staticmethod
    def backward(ctx, dx2):
        (x, h1, mean1, rstd1, qkv, ctxv, lse, x1, h2, mean2, rstd2, pre, hmid, Pqkv, Po, P1, P2, ln1, ln2, B, S, H, DH) = ctx.saved
        D = H * DH
        prev_gb, own_b2_fused = ctx.fuse
        dx2 = dx2.contiguous()
        _wgrad(dx2, hmid, P2, bias=not own_b2_fused, side=True)
        dpre = _dgrad(dx2, P2, act=ops.ACT_GELU_GRAD, aux_in=pre)
        _wgrad(dpre, h2, P1, side=True)
        dh2 = _dgrad(dpre, P1)
        dx1 = ops.layernorm_bwd(dh2, x1, mean2, rstd2, ln2[0], ln2[2], ln2[3], dres=dx2, colsum=Po.gb)
        ctx2d = ctxv.view(B * S, D)
        _wgrad(dx1, ctx2d, Po, bias=False, side=True)
        dctx = _dgrad(dx1, Po).view(B, S, D)
        dqkv = torch.empty_like(qkv)
        q3, d3 = qkv.view(B, S, 3 * D), dqkv.view(B, S, 3 * D)
        ops.attention_bwd(q3[:, :, :D], q3[:, :, D:2 * D], q3[:, :, 2 * D:], ctxv, dctx, lse,
                          d3[:, :, :D], d3[:, :, D:2 * D], d3[:, :, 2 * D:], H, DH)
        _wgrad(dqkv, h1, Pqkv, side=True)
        dh1 = _dgrad(dqkv, Pqkv)
        dx = ops.layernorm_bwd(dh1, x, mean1, rstd1, ln1[0], ln1[2], ln1[3], dres=dx1, colsum=prev_gb)
        notify_grad_ready(ctx.layer)      # (this layer's FFN-down bias gradient was completed by the layer above / the final LayerNorm)
        return dx, None, None, None, None, None, None, None, None


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to_dict(self):
        return dict(self.__dict__)

    def __repr__(self):
        return "Config(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self.__dict__.items()))


def vit_config(**kw):
    """Defaults of transformers.ViTConfig (HF configuration_vit.py) for the keys the kernels consume."""
    d = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
             hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, initializer_range=0.02, layer_norm_eps=1e-12,
             image_size=224, patch_size=16, num_channels=3, qkv_bias=True, distillation=False)
    d.update(kw)
    cfg = _Cfg(**d)
    if cfg.hidden_act != "gelu":
        raise NotImplementedError("ViT hidden_act %r: only exact-erf 'gelu' has a kernel" % cfg.hidden_act)
    if not cfg.qkv_bias:
        raise NotImplementedError("qkv_bias=False is not supported")
    if cfg.hidden_dropout_prob != 0.0 or cfg.attention_probs_dropout_prob != 0.0:
        raise NotImplementedError("ViT dropout > 0 is not wired (ViTConfig default is 0.0)")
    return cfg


class _Holder(nn.Module):
    """Parameter container that mirrors an HF sub-module name; never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder; the tower's fused kernels run the math")


def _trunc_normal_(t, std):
    nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std)


class ViTTower(nn.Module):
    """Drop-in for ViTModel(ViTConfig(**kw), add_pooling_layer=False) (vilmedic/blocks/vision/visual_encoder.py:56-58)."""

    def __init__(self, **kw):
        super().__init__()
        cfg = self.cfg = self.config = vit_config(**{k: v for k, v in kw.items() if k != "return_dict"})
        D = cfg.hidden_size
        n_patches = (cfg.image_size // cfg.patch_size) ** 2
        emb = self.embeddings = _Holder()
        emb.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        if cfg.distillation:          # DeiTModel (vilmedic/blocks/vision/visual_encoder.py:60-61): same blocks + a distillation token
            emb.distillation_token = nn.Parameter(torch.zeros(1, 1, D))
        emb.position_embeddings = nn.Parameter(torch.zeros(1, n_patches + (2 if cfg.distillation else 1), D))
        emb.patch_embeddings = _Holder()
        emb.patch_embeddings.projection = nn.Conv2d(cfg.num_channels, D, cfg.patch_size, cfg.patch_size)
        self.encoder = _Holder()
        self.encoder.layer = nn.ModuleList()
        for _ in range(cfg.num_hidden_layers):
            l = _Holder()
            l.cfg = cfg
            l.attention = _Holder()
            l.attention.attention = _Holder()
            l.attention.attention.query = nn.Linear(D, D)
            l.attention.attention.key = nn.Linear(D, D)
            l.attention.attention.value = nn.Linear(D, D)
            l.attention.output = _Holder()
            l.attention.output.dense = nn.Linear(D, D)
            l.intermediate = _Holder()
            l.intermediate.dense = nn.Linear(D, cfg.intermediate_size)
            l.output = _Holder()
            l.output.dense = nn.Linear(cfg.intermediate_size, D)
            l.layernorm_before = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
            l.layernorm_after = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
            att = l.attention.attention
            l._fused_param_groups = (lambda a=att: [[a.query.weight, a.key.weight, a.value.weight],
                                                    [a.query.bias, a.key.bias, a.value.bias]])
            self.encoder.layer.append(l)
        self.layernorm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
        self.reset_parameters()

    def reset_parameters(self):
        std = self.cfg.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                _trunc_normal_(m.weight.data, std)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.LayerNorm):
                m.weight.data.fill_(1.0)
                m.bias.data.zero_()
        _trunc_normal_(self.embeddings.cls_token.data, std)
        _trunc_normal_(self.embeddings.position_embeddings.data, std)
        if self.cfg.distillation:
            _trunc_normal_(self.embeddings.distillation_token.data, std)

    def forward(self, pixel_values):
        """pixel_values fp32 [B,C,H,W] (CUDA) -> last_hidden_state bf16 [B,S,D]."""
        cfg = self.cfg
        if pixel_values.dim() != 4 or pixel_values.shape[1] != cfg.num_channels:
            raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the configuration.")
        if pixel_values.shape[2] != cfg.image_size or pixel_values.shape[3] != cfg.image_size:
            raise ValueError("Input image size (%d*%d) doesn't match model (%d*%d)." % (
                pixel_values.shape[2], pixel_values.shape[3], cfg.image_size, cfg.image_size))
        arena = get_arena(_root_of(self))
        # nothing is saved for a backward that cannot happen (frozen tower: VisualEncoder(freeze=True), visual_encoder.py:124-128)
        grad = torch.is_grad_enabled() and (pixel_values.requires_grad or any(p.requires_grad for p in self.parameters()))
        _prepare(arena, self)
        anchor = self.layernorm.weight
        images = pixel_values.contiguous().float()
        B = images.shape[0]
        x = ViTEmbedFn.apply(images, anchor, self, arena)
        S = x.shape[0] // B
        prev_gb = None                       # layer 0's input comes from the patch embedding (its bias grad: vit_embed_bwd)
        for layer in self.encoder.layer:
            x = ViTLayerFn.apply(x, anchor, layer, arena, B, S, grad, prev_gb, True)
            prev_gb = arena.grad(layer.output.dense.bias)
        x = LayerNormFn.apply(x, anchor, _ln(arena, self.layernorm), cfg.layer_norm_eps, prev_gb)
        return x.view(B, S, cfg.hidden_size)


def _root_of(module):
    """The outermost module that registered itself as arena root (set by the model wrappers); default: the module."""
    root = getattr(module, "_vlm_root", None)
    return root if root is not None else module


def set_arena_root(root):
    """Make every kernel tower under `root` share one flat parameter arena, prepared once per root forward."""
    for m in root.modules():
        object.__setattr__(m, "_vlm_root", root)
    if getattr(root, "_vlm_root_hooks", False):
        return
    object.__setattr__(root, "_vlm_root_hooks", True)

    def pre(mod, args, kwargs=None):
        if next(mod.parameters()).is_cuda:
            arena = get_arena(mod)
            if torch.is_grad_enabled() and mod.training:
                arena.prepare_step()
            else:
                arena.refresh_mirror()
            arena.in_root_forward = True

    def post(mod, args, out):
        arena = getattr(mod, "_vlm_arena", None)
        if arena is not None:
            arena.in_root_forward = False

    root.register_forward_pre_hook(pre)
    root.register_forward_hook(post, always_call=True)


def _prepare(arena, module):
    if getattr(arena, "in_root_forward", False):
        return
    if torch.is_grad_enabled() and module.training:
        arena.prepare_step()
    else:
        arena.refresh_mirror()


# =============================================================================================== BERT-style tower
def bert_config(**kw):
    """Defaults of transformers.BertGenerationConfig (HF configuration_bert_generation.py) — note they are
    BERT-large-shaped (hidden 1024, 24 layers, 16 heads, FFN 4096, vocab 50358); configs must spell the sizes out
    (SURVEY.md §8d)."""
    d = dict(vocab_size=50358, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
             hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
             initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, bos_token_id=2, eos_token_id=1,
             is_decoder=False, add_cross_attention=False, tie_word_embeddings=True, type_vocab_size=0,
             encoder_hidden_size=None,
             # checkpoints loaded through `proto` (hf_loader.py): "bert-generation" (default, what the reference builds when proto is
             # null), "bert" (token-type table, cls.predictions head) or "roberta" (padding-aware positions, lm_head with a transform)
             family="bert-generation")
    d.update({k: v for k, v in kw.items() if k not in ("proto", "return_dict")})
    cfg = _Cfg(**d)
    if cfg.hidden_act != "gelu":
        raise NotImplementedError("hidden_act %r: only exact-erf 'gelu' has a kernel" % cfg.hidden_act)
    if cfg.family not in ("bert-generation", "bert", "roberta"):
        raise NotImplementedError("model family %r has no kernel tower (bert-generation, bert, roberta)" % cfg.family)
    if cfg.family == "bert-generation":
        cfg.type_vocab_size = 0                                   # BertGenerationEmbeddings has no token-type table
    if cfg.hidden_size % cfg.num_attention_heads != 0:
        raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)" % (
            cfg.hidden_size, cfg.num_attention_heads))
    if cfg.hidden_size // cfg.num_attention_heads not in (48, 64, 96):
        raise NotImplementedError("attention head dim %d has no kernel (48, 64, 96 supported)" % (
            cfg.hidden_size // cfg.num_attention_heads))
    return cfg


class BertEmbedFn(torch.autograd.Function):
    """word + position (+ token_type 0) -> LayerNorm.  HF modeling_bert_generation.py:395-429 / modeling_bert.py embeddings."""

    @staticmethod
    def forward(ctx, ids, anchor, emb, arena, T, pos_offset, eps, pos_ids=None):
        word = arena.fp32(emb.word_embeddings.weight)
        pos = arena.fp32(emb.position_embeddings.weight)
        tt = getattr(emb, "token_type_embeddings", None)
        tt_row = arena.fp32(tt.weight)[0].contiguous() if tt is not None else None      # token_type_ids = 0 everywhere
        z = ops.embed_fwd(ids, word, pos, T, pos_offset, pos_ids=pos_ids, tt_row=tt_row)
        lnp = _ln(arena, emb.LayerNorm)
        y, mean, rstd = ops.layernorm_fwd(z, lnp[0], lnp[1], eps)
        ctx.saved = (ids, z, mean, rstd, lnp, emb, arena, T, pos_offset, pos_ids)
        return y

    @staticmethod
    def backward(ctx, dy):
        ids, z, mean, rstd, lnp, emb, arena, T, pos_offset, pos_ids = ctx.saved
        dz = ops.layernorm_bwd(dy.contiguous(), z, mean, rstd, lnp[0], lnp[2], lnp[3])
        V = emb.word_embeddings.weight.shape[0]
        pad = emb.word_embeddings.padding_idx
        ops.embed_bwd(ids, dz, arena.grad(emb.word_embeddings.weight), arena.grad(emb.position_embeddings.weight), T, V,
                      pos_offset, -1 if pad is None else int(pad), pos_ids=pos_ids)
        tt = getattr(emb, "token_type_embeddings", None)
        if tt is not None:                                            # every token used row 0: its gradient is the column sum of dz
            ops.colsum(dz, arena.grad(tt.weight)[0])
        notify_grad_ready(emb)            # the tied LM-head weight gradient was accumulated first (LMHeadCEFn.backward)
        return None, None, None, None, None, None, None, None


class BertLayerFn(torch.autograd.Function):
    """One post-LN block: self-attention (+causal) -> [cross-attention] -> FFN.
    HF modeling_bert_generation.py:52-56,89-153 (self), :181-232 (cross), :265-293 (FFN), :296-360 (layer)."""

    @staticmethod
    def forward(ctx, x, enc, anchor, layer, arena, B, T, kmask, enc_mask, causal, training, save):
        cfg = layer.cfg
        D, H = cfg.hidden_size, cfg.num_attention_heads
        DH = D // H
        eps = cfg.layer_norm_eps
        p_h = cfg.hidden_dropout_prob if training else 0.0
        p_a = cfg.attention_probs_dropout_prob if training else 0.0
        sa = layer.attention
        Pqkv = _lin(arena, sa.self.query, sa.self.key, sa.self.value)
        Po = _lin(arena, sa.output.dense)
        ln1 = _ln(arena, sa.output.LayerNorm)
        P1, P2 = _lin(arena, layer.intermediate.dense), _lin(arena, layer.output.dense)
        ln3 = _ln(arena, layer.output.LayerNorm)
        rng = {}

        def R(name):
            rng[name] = _next_rng() if (p_h > 0 or p_a > 0) else (0, 0)
            return rng[name]

        qkv = ops.gemm(x, Pqkv.w, bias=Pqkv.b)
        q3 = qkv.view(B, T, 3 * D)
        s, o = R("a1")
        ctx1, lse1 = ops.attention_fwd(q3[:, :, :D], q3[:, :, D:2 * D], q3[:, :, 2 * D:], H, DH, kmask=kmask, causal=causal,
                                       p_drop=p_a, seed=s, offset=o)
        s, o = R("h1")
        z1 = ops.gemm(ctx1.view(B * T, D), Po.w, bias=Po.b, residual=x, p_drop=p_h, seed=s, offset=o)
        x1, m1, r1 = ops.layernorm_fwd(z1, ln1[0], ln1[1], eps)
        cross = enc is not None
        if cross:
            ca = layer.crossattention
            Pq = _lin(arena, ca.self.query)
            Pkv = _lin(arena, ca.self.key, ca.self.value)
            Poc = _lin(arena, ca.output.dense)
            ln2 = _ln(arena, ca.output.LayerNorm)
            Se = enc.shape[0] // B
            qc = ops.gemm(x1, Pq.w, bias=Pq.b)
            kvc = ops.gemm(enc, Pkv.w, bias=Pkv.b)
            kv3 = kvc.view(B, Se, 2 * D)
            s, o = R("a2")
            ctx2, lse2 = ops.attention_fwd(qc.view(B, T, D), kv3[:, :, :D], kv3[:, :, D:], H, DH, kmask=enc_mask,
                                           p_drop=p_a, seed=s, offset=o)
            s, o = R("h2")
            z2 = ops.gemm(ctx2.view(B * T, D), Poc.w, bias=Poc.b, residual=x1, p_drop=p_h, seed=s, offset=o)
            x2, m2, r2 = ops.layernorm_fwd(z2, ln2[0], ln2[1], eps)
        else:
            x2 = x1
        pre = torch.empty((B * T, cfg.intermediate_size), device=x.device, dtype=torch.bfloat16) if save else None
        hmid = ops.gemm(x2, P1.w, bias=P1.b, act=ops.ACT_GELU, aux_out=pre)   # pre <- GELU'(pre-activation)
        s, o = R("h3")
        z3 = ops.gemm(hmid, P2.w, bias=P2.b, residual=x2, p_drop=p_h, seed=s, offset=o)
        x3, m3, r3 = ops.layernorm_fwd(z3, ln3[0], ln3[1], eps)
        if save:
            ctx.c = dict(x=x, enc=enc, qkv=qkv, ctx1=ctx1, lse1=lse1, z1=z1, x1=x1, m1=m1, r1=r1, x2=x2, pre=pre, hmid=hmid,
                         z3=z3, m3=m3, r3=r3, Pqkv=Pqkv, Po=Po, ln1=ln1, P1=P1, P2=P2, ln3=ln3, B=B, T=T, H=H, DH=DH,
                         kmask=kmask, enc_mask=enc_mask, causal=causal, p_h=p_h, p_a=p_a, rng=rng, cross=cross)
            if cross:
                ctx.c.update(Pq=Pq, Pkv=Pkv, Poc=Poc, ln2=ln2, qc=qc, kvc=kvc, ctx2=ctx2, lse2=lse2, z2=z2, m2=m2, r2=r2, Se=Se)
            ctx.layer = layer
        return x3

    @staticmethod
    def backward(ctx, dx3):
        c = ctx.c
        B, T, H, DH = c["B"], c["T"], c["H"], c["DH"]
        D = H * DH
        p_h, p_a, rng = c["p_h"], c["p_a"], c["rng"]

        def drop(g, name):
            if p_h <= 0:
                return g
            s, o = rng[name]
            return ops.dropout(g, p_h, s, o)

        ln1, ln3 = c["ln1"], c["ln3"]

        def ln_bwd(dy, z, m, r, ln, name, P):
            """LN backward with the hidden-dropout mask and the bias gradient of `P` fused in; returns (dz, dropped dz)."""
            if p_h > 0:
                s_, o_ = rng[name]
                return ops.layernorm_bwd(dy, z, m, r, ln[0], ln[2], ln[3], drop=(p_h, s_, o_), colsum=P.gb)
            dz = ops.layernorm_bwd(dy, z, m, r, ln[0], ln[2], ln[3], colsum=P.gb)
            return dz, dz

        dz3, dz3d = ln_bwd(dx3.contiguous(), c["z3"], c["m3"], c["r3"], ln3, "h3", c["P2"])
        _wgrad(dz3d, c["hmid"], c["P2"], bias=False, side=True)
        dpre = _dgrad(dz3d, c["P2"], act=ops.ACT_GELU_GRAD, aux_in=c["pre"])
        _wgrad(dpre, c["x2"], c["P1"], side=True)
        dx2 = _dgrad(dpre, c["P1"], residual=dz3)
        denc = None
        if c["cross"]:
            ln2 = c["ln2"]
            Se = c["Se"]
            dz2, dz2d = ln_bwd(dx2, c["z2"], c["m2"], c["r2"], ln2, "h2", c["Poc"])
            _wgrad(dz2d, c["ctx2"].view(B * T, D), c["Poc"], bias=False, side=True)
            dctx2 = _dgrad(dz2d, c["Poc"]).view(B, T, D)
            dqc = torch.empty_like(c["qc"])
            dkvc = torch.empty_like(c["kvc"])
            kv3, dkv3 = c["kvc"].view(B, Se, 2 * D), dkvc.view(B, Se, 2 * D)
            s, o = rng["a2"]
            ops.attention_bwd(c["qc"].view(B, T, D), kv3[:, :, :D], kv3[:, :, D:], c["ctx2"], dctx2, c["lse2"],
                              dqc.view(B, T, D), dkv3[:, :, :D], dkv3[:, :, D:], H, DH, kmask=c["enc_mask"], p_drop=p_a,
                              seed=s, offset=o)
            _wgrad(dqc, c["x1"], c["Pq"], side=True)
            dx1 = _dgrad(dqc, c["Pq"], residual=dz2)
            _wgrad(dkvc, c["enc"], c["Pkv"], side=True)
            denc = _dgrad(dkvc, c["Pkv"])
        else:
            dx1 = dx2
        dz1, dz1d = ln_bwd(dx1, c["z1"], c["m1"], c["r1"], ln1, "h1", c["Po"])
        _wgrad(dz1d, c["ctx1"].view(B * T, D), c["Po"], bias=False, side=True)
        dctx1 = _dgrad(dz1d, c["Po"]).view(B, T, D)
        dqkv = torch.empty_like(c["qkv"])
        q3, d3 = c["qkv"].view(B, T, 3 * D), dqkv.view(B, T, 3 * D)
        s, o = rng["a1"]
        ops.attention_bwd(q3[:, :, :D], q3[:, :, D:2 * D], q3[:, :, 2 * D:], c["ctx1"], dctx1, c["lse1"],
                          d3[:, :, :D], d3[:, :, D:2 * D], d3[:, :, 2 * D:], H, DH, kmask=c["kmask"], causal=c["causal"],
                          p_drop=p_a, seed=s, offset=o)
        _wgrad(dqkv, c["x"], c["Pqkv"], side=True)
        dx = _dgrad(dqkv, c["Pqkv"], residual=dz1)
        notify_grad_ready(ctx.layer)
        return (dx, denc) + (None,) * 10


class LMHeadCEFn(torch.autograd.Function):
    """logits = h E^T + b (weight tied to the word embeddings) and the shifted mean cross-entropy of
    HF loss_utils.py:28-66 with labels=input_ids (vilmedic decoder_model.py:46).  dlogits overwrite the logits buffer
    unless keep_logits.  Returns (loss, logits|empty)."""

    @staticmethod
    def forward(ctx, h, anchor, ids, head, arena, B, T, keep_logits, need_grad):
        E = head.decoder.weight
        V, D = E.shape
        Vp = (V + 7) // 8 * 8
        w = arena.bf16(E)
        bias = arena.fp32(head.bias)
        if Vp != V:
            bias_p = torch.zeros(Vp, device=h.device, dtype=torch.float32)
            bias_p[:V] = bias
        else:
            bias_p = bias
        buf = torch.empty((B * T, Vp), device=h.device, dtype=torch.bfloat16)
        ops.gemm(h, w, out=buf[:, :V], bias=bias_p)  # pad columns [V,Vp) are never read as logits; CE zeroes their grads
        n_valid = B * (T - 1)
        dl = None
        if need_grad:
            dl = torch.empty_like(buf) if keep_logits else buf
        rows, _ = ops.softmax_ce(buf, ids, V, shift_T=T, grad_scale=1.0 / max(n_valid, 1), dlogits=dl)
        loss = ops.sum_scale(rows, 1.0 / max(n_valid, 1))
        ctx.saved = (h, dl, head, arena, V, Vp)
        logits = buf.view(B, T, Vp)[:, :, :V] if (keep_logits or not need_grad) else buf.new_empty(0)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        h, dl, head, arena, V, Vp = ctx.saved
        g = dloss.contiguous().float()
        E = head.decoder.weight
        gE = arena.grad(E)
        # dE += g * dlogits^T h   (rows >= V of the padded product are cut by M = V)
        gbias = arena.grad(head.bias)
        ops.SIDE.run(lambda: ops.colsum(dl[:, :V], gbias, g), dl, g)     # under the two LM-head GEMMs; joined by notify_grad_ready
        ops.SIDE_GEMM.run(lambda: ops.gemm(dl[:, :V], h, a_mn_major=True, b_mn_major=True, out=gE, accumulate=True, alpha_t=g), dl, h, g)
        dh = ops.gemm(dl[:, :V], arena.bf16(E), b_mn_major=True, alpha_t=g)
        notify_grad_ready(head)           # lm_head.bias; the tied embedding matrix is announced with the embedding block
        return dh, None, None, None, None, None, None, None, None


def make_bert_layer(cfg):
    """One HF BertLayer-shaped parameter holder (attention.self.{query,key,value}, attention.output.{dense,LayerNorm},
    [crossattention.…], intermediate.dense, output.{dense,LayerNorm})."""
    D = cfg.hidden_size
    De = cfg.encoder_hidden_size or D
    l = _Holder()
    l.cfg = cfg
    groups = []

    def attn_block(in_kv):
        a = _Holder()
        a.self = _Holder()
        a.self.query = nn.Linear(D, D)
        a.self.key = nn.Linear(in_kv, D)
        a.self.value = nn.Linear(in_kv, D)
        a.output = _Holder()
        a.output.dense = nn.Linear(D, D)
        a.output.LayerNorm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
        return a

    l.attention = attn_block(D)
    s = l.attention.self
    groups += [[s.query.weight, s.key.weight, s.value.weight], [s.query.bias, s.key.bias, s.value.bias]]
    if cfg.add_cross_attention:
        if not cfg.is_decoder:
            raise ValueError("add_cross_attention requires is_decoder=True")
        l.crossattention = attn_block(De)
        cs = l.crossattention.self
        groups += [[cs.key.weight, cs.value.weight], [cs.key.bias, cs.value.bias]]
    l.intermediate = _Holder()
    l.intermediate.dense = nn.Linear(D, cfg.intermediate_size)
    l.output = _Holder()
    l.output.dense = nn.Linear(cfg.intermediate_size, D)
    l.output.LayerNorm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
    l._fused_param_groups = (lambda g=groups: g)
    return l


def _init_bert_weights(module, std):
    for m in module.modules():
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(0.0, std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, nn.LayerNorm):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()


def run_bert_layers(layers, x, B, T, arena, anchor, cfg, kmask=None, enc=None, enc_mask=None, training=False, collect=None):
    """collect: optional list that receives the hidden states after every layer (HF output_hidden_states)."""
    grad = torch.is_grad_enabled()
    for layer in layers:
        x = BertLayerFn.apply(x, enc, anchor, layer, arena, B, T, kmask, enc_mask, bool(cfg.is_decoder), training and grad, grad)
        if collect is not None:
            collect.append(x)
    return x


class BertEncoderB200(nn.Module):
    """transformers.models.bert.modeling_bert.BertEncoder-shaped stack (state_dict keys `layer.N.…`): hidden states in,
    hidden states out — the `transformer` of vilmedic/models/mvqa/MVQA.py:28-30,43."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = self.config = cfg
        self.layer = nn.ModuleList([make_bert_layer(cfg) for _ in range(cfg.num_hidden_layers)])
        _init_bert_weights(self, cfg.initializer_range)

    def forward(self, hidden_states, attention_mask=None, **kwargs):
        """hidden_states bf16 [B,T,D] -> bf16 [B,T,D]"""
        arena = get_arena(_root_of(self))
        _prepare(arena, self)
        B, T, D = hidden_states.shape
        kmask = None
        if attention_mask is not None:
            kmask = (attention_mask.to(arena.device) != 0).to(torch.uint8).contiguous()
        x = hidden_states.reshape(B * T, D).contiguous()
        anchor = self.layer[0].output.LayerNorm.weight
        x = run_bert_layers(self.layer, x, B, T, arena, anchor, self.cfg, kmask=kmask, training=self.training)
        return x.view(B, T, D)


def native_linear(linear, x2d, module, out_dtype=torch.bfloat16):
    """y = x W^T + b through the tcgen05 GEMM for a plain nn.Linear parameter holder living in `module`'s arena."""
    arena = get_arena(_root_of(module))
    _prepare(arena, module)
    return LinearFn.apply(_to_bf16(x2d).contiguous(), linear.weight, _lin(arena, linear), out_dtype)


def native_layernorm(ln, x2d, module):
    arena = get_arena(_root_of(module))
    _prepare(arena, module)
    return LayerNormFn.apply(x2d.contiguous(), ln.weight, _ln(arena, ln), ln.eps)


class BertTower(nn.Module):
    """Post-LN transformer stack with HF BertGenerationEncoder/Decoder parameter names under `.bert` + `.lm_head`
    (decoder) — state_dict keys identical to BertGenerationDecoder (vilmedic decoder_model.py:23-26) — or bare
    encoder (`with_lm_head=False`)."""

    def __init__(self, cfg, with_lm_head, flat=False):
        """flat=True: parameters live at `embeddings.…` / `encoder.layer.…` (BertGenerationEncoder layout, the text tower of
        vilmedic/blocks/huggingface/encoder/encoder_model.py:24-26); otherwise under `bert.…` (BertGenerationDecoder)."""
        super().__init__()
        self.cfg = self.config = cfg
        D = cfg.hidden_size
        fam = cfg.family
        if flat:
            core = self
        else:                                      # BertGenerationDecoder / BertLMHeadModel keep the stack under `bert.`, Roberta under `roberta.`
            core = _Holder()
            setattr(self, "roberta" if fam == "roberta" else "bert", core)
        object.__setattr__(self, "_core", core)
        emb = core.embeddings = _Holder()
        emb.word_embeddings = nn.Embedding(cfg.vocab_size, D, padding_idx=cfg.pad_token_id)
        emb.position_embeddings = nn.Embedding(cfg.max_position_embeddings, D,
                                               padding_idx=cfg.pad_token_id if fam == "roberta" else None)
        if cfg.type_vocab_size:
            emb.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, D)
        emb.LayerNorm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
        core.encoder = _Holder()
        core.encoder.layer = nn.ModuleList([make_bert_layer(cfg) for _ in range(cfg.num_hidden_layers)])
        self.lm_head = None
        object.__setattr__(self, "_head", None)
        object.__setattr__(self, "_head_transform", None)
        if with_lm_head:
            head = _Holder()
            head.bias = nn.Parameter(torch.zeros(cfg.vocab_size))
            head.decoder = nn.Linear(D, cfg.vocab_size)
            head.decoder.bias = head.bias
            if cfg.tie_word_embeddings:
                head.decoder.weight = emb.word_embeddings.weight
            if fam == "bert-generation":           # lm_head.{decoder, bias}
                self.lm_head = head
            elif fam == "roberta":                 # lm_head.{dense, layer_norm, decoder, bias}
                head.dense = nn.Linear(D, D)
                head.layer_norm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
                self.lm_head = head
                object.__setattr__(self, "_head_transform", (head.dense, head.layer_norm))
            else:                                  # cls.predictions.{transform.{dense, LayerNorm}, decoder, bias}
                self.cls = _Holder()
                self.cls.predictions = head
                head.transform = _Holder()
                head.transform.dense = nn.Linear(D, D)
                head.transform.LayerNorm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
                object.__setattr__(self, "_head_transform", (head.transform.dense, head.transform.LayerNorm))
            object.__setattr__(self, "_head", head)
        self.reset_parameters()

    def reset_parameters(self):
        _init_bert_weights(self, self.cfg.initializer_range)

    # ---- hidden states --------------------------------------------------------------------------------------------
    def hidden_states(self, input_ids, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                      inputs_embeds=None, collect=None):
        """-> bf16 [B*T, D] after the last layer.  encoder_hidden_states: bf16 [B,S,De] (CUDA).
        collect: optional list that receives the embedding output and every layer's output (HF `output_hidden_states=True`)."""
        cfg = self.cfg
        arena = get_arena(_root_of(self))
        grad = torch.is_grad_enabled()
        training = self.training
        _prepare(arena, self)
        anchor = self._core.embeddings.LayerNorm.weight
        dev = arena.device
        if inputs_embeds is not None:
            B, T = inputs_embeds.shape[0], inputs_embeds.shape[1]
            x = LayerNormFn.apply(inputs_embeds.reshape(B * T, -1).contiguous(), anchor,
                                  _ln(arena, self._core.embeddings.LayerNorm), cfg.layer_norm_eps)
        else:
            ids = input_ids.to(dev).long().contiguous()
            B, T = ids.shape
            if T > cfg.max_position_embeddings:
                raise IndexError("sequence length %d exceeds max_position_embeddings %d" % (T, cfg.max_position_embeddings))
            pos_ids = None
            if cfg.family == "roberta":            # HF create_position_ids_from_input_ids: padding_idx + running count of non-pad tokens
                pad = cfg.pad_token_id
                m = (ids != pad).to(torch.int32)
                pos_ids = ((torch.cumsum(m, dim=1).to(torch.int32) * m) + pad).contiguous().view(-1)
            x = BertEmbedFn.apply(ids.view(-1), anchor, self._core.embeddings, arena, T, 0, cfg.layer_norm_eps, pos_ids)
        x = dropout(x, cfg.hidden_dropout_prob, training and grad)
        if collect is not None:
            collect.append(x)
        kmask = None
        if attention_mask is not None:
            kmask = (attention_mask.to(dev) != 0).to(torch.uint8).contiguous()
        enc = enc_mask = None
        if encoder_hidden_states is not None:
            if not cfg.add_cross_attention:
                raise ValueError("encoder_hidden_states given but the tower has no cross-attention layers")
            e = encoder_hidden_states
            if e.dtype != torch.bfloat16:
                e = _to_bf16(e)
            enc = e.reshape(e.shape[0] * e.shape[1], e.shape[2]).contiguous()
            if encoder_attention_mask is not None:
                enc_mask = (encoder_attention_mask.to(dev) != 0).to(torch.uint8).contiguous()
        x = run_bert_layers(self._core.encoder.layer, x, B, T, arena, anchor, cfg, kmask=kmask, enc=enc, enc_mask=enc_mask,
                            training=training, collect=collect)
        return x, B, T

    def head_input(self, x):
        """BERT / RoBERTa LM heads transform the hidden state first: LayerNorm(GELU(dense(x)))."""
        if self._head_transform is None:
            return x
        dense, ln = self._head_transform
        arena = get_arena(_root_of(self))
        h = LinearGeluFn.apply(x, dense.weight, _lin(arena, dense))
        return LayerNormFn.apply(h, ln.weight, _ln(arena, ln), ln.eps)

    def lm_loss(self, x, input_ids, B, T, keep_logits):
        arena = get_arena(_root_of(self))
        ids = input_ids.to(arena.device).long().contiguous().view(-1)
        grad = torch.is_grad_enabled()
        return LMHeadCEFn.apply(self.head_input(x), self._core.embeddings.LayerNorm.weight, ids, self._head, arena, B, T, keep_logits, grad)

    def lm_logits(self, x, padded=False):
        """bf16 [M,D] -> fp32 [M,V] logits (inference); padded=True returns the [M, Vp] buffer (row pitch Vp = V rounded up to 4)."""
        arena = get_arena(_root_of(self))
        x = self.head_input(x)
        E = self._head.decoder.weight
        V = E.shape[0]
        Vp = (V + 3) // 4 * 4
        out = torch.empty((x.shape[0], Vp), device=x.device, dtype=torch.float32)
        bias = arena.fp32(self._head.bias)
        if Vp != V:
            bp = torch.zeros(Vp, device=x.device, dtype=torch.float32)
            bp[:V] = bias
            bias = bp
        ops.gemm(x, arena.bf16(E), out=out[:, :V], bias=bias)
        return out if padded else out[:, :V]


class CastBf16Fn(torch.autograd.Function):
    """fp32 -> bf16 boundary cast (gradients flow back as fp32)."""

    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return dy.float()


def _to_bf16(t):
    if t.dtype == torch.bfloat16:
        return t
    if t.dtype != torch.float32:
        t = t.float()
    return CastBf16Fn.apply(t)
