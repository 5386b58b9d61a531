"""GLoRIALoss (global + local) on the sm_100a kernels — same constructor kwargs, call signature and return tuple as
vilmedic/blocks/losses/selfsup/GLoRIALoss.py:132-170, plus the module-level helpers the reference exports
(`cosine_similarity` :5-10, `gloria_attention_fn` :13-51, `global_loss` :54-75, `local_loss` :78-129).

The reference's local_loss loops over the B captions in Python and runs, per caption, a [B,S,D]x[D,n] bmm, two softmaxes, a
[B,D,S]x[S,n] bmm and a cosine.  Here all B x B (image, caption) pairs are evaluated together:
  A  = Xc Ww^T                  one tcgen05 GEMM   [B*S, D] x [B*L, D]^T          (csrc/gemm_tcgen05.cu)
  P1, P2                        two fused softmax kernels                          (csrc/gloria.cu)
  WC_i = P2_i^T Xc_i            one batched tcgen05 GEMM (both operands MN-major: no transposes are materialised)
  cos, sims, CE                 gloria_cos / gloria_sims / sym_lse
and the backward is four more GEMMs + three kernels.  Operands of the tensor-core products are bf16 (fp32 accumulate);
softmaxes, cosine, log-sum-exp and the cross-entropies are fp32.
"""
import torch
import torch.nn as nn

from ... import ops
from .contrastive import GLoRIAGlobalLoss


def _cap_lens_tensor(cap_lens, device):
    if torch.is_tensor(cap_lens):
        return cap_lens.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor([int(c) for c in cap_lens], device=device, dtype=torch.int32)


class _GloriaLocalFn(torch.autograd.Function):
    """(img_features [B,D,ih,iw], words_emb [B,D,Lw], cap_lens int32 [B]) -> (loss0, loss1, P2 fp32 [B,S,B*L])."""

    @staticmethod
    def forward(ctx, img, words, cap_lens, temp1, temp2, temp3):
        B, D = img.shape[0], img.shape[1]
        S = img.shape[2] * img.shape[3]
        Lw = words.shape[2]
        L = (Lw + 7) // 8 * 8                                   # NL = B*L keeps every GEMM pitch 16-byte aligned
        NL = B * L
        if words.shape[0] != B or words.shape[1] != D:
            raise ValueError("GLoRIA local loss: img_features %s and words_emb %s disagree" % (tuple(img.shape), tuple(words.shape)))
        if D % 8 != 0:
            raise ValueError("GLoRIA local loss: feature dim must be a multiple of 8")
        img3 = img.float().reshape(B, D, S).contiguous()
        w3 = words.float().contiguous()
        # forward products use the bf16 hi/lo split (x = hi + lo): hi*hi + lo*hi + hi*lo on the tensor cores, fp32 accumulate
        Xc, Xl = ops.transpose_cast(img3, torch.bfloat16, want_lo=True)                   # [B, S, D]
        Ww, Wl = ops.transpose_cast(w3, torch.bfloat16, c_out=L, row_limit=cap_lens, want_lo=True)   # [B, L, D]
        Q = ops.transpose_cast(w3, torch.float32, c_out=L, row_limit=cap_lens)            # [B, L, D] fp32 (cosine operand)
        A = ops.gemm(Xc.view(B * S, D), Ww.view(NL, D), out_dtype=torch.float32)          # [B*S, NL]
        ops.gemm(Xl.view(B * S, D), Ww.view(NL, D), out=A, accumulate=True)
        ops.gemm(Xc.view(B * S, D), Wl.view(NL, D), out=A, accumulate=True)
        P1 = ops.gloria_word_softmax(A, cap_lens, B, L)
        del A
        P2, P2h, P2l = ops.gloria_region_softmax(P1, cap_lens, B, S, B, L, temp1)
        P2h3 = P2h.view(B, S, NL)
        WC = ops.gemm(P2h3, Xc, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)  # [B, NL, D]
        ops.gemm(P2l.view(B, S, NL), Xc, a_mn_major=True, b_mn_major=True, out=WC, accumulate=True)
        ops.gemm(P2h3, Xl, a_mn_major=True, b_mn_major=True, out=WC, accumulate=True)
        del P2l, Xl, Wl
        cosv, wnorm, qnorm = ops.gloria_cos(WC, Q.view(NL, D), cap_lens, B, L)
        sims = ops.gloria_sims(cosv, cap_lens, B, L, temp2, temp3)
        lse_row, lse_col, loss_row, loss_col = ops.sym_lse(sims, 1.0)
        loss0 = ops.sum_scale(loss_row, 1.0 / B)
        loss1 = ops.sum_scale(loss_col, 1.0 / B)
        ctx.saved = (Xc, Ww, Q, P1, P2, P2h3, WC, cosv, wnorm, qnorm, sims, lse_row, lse_col, cap_lens)
        ctx.dims = (B, D, S, Lw, L, img.shape, temp1, temp2, temp3, img.dtype, words.dtype)
        P2v = P2.view(B, S, NL)
        ctx.mark_non_differentiable(P2v)
        return loss0, loss1, P2v

    @staticmethod
    def backward(ctx, g0, g1, _gp):
        Xc, Ww, Q, P1, P2, P2h3, WC, cosv, wnorm, qnorm, sims, lse_row, lse_col, cap_lens = ctx.saved
        B, D, S, Lw, L, img_shape, temp1, temp2, temp3, img_dtype, words_dtype = ctx.dims
        NL = B * L
        g0 = g0.float().contiguous()
        g1 = g1.float().contiguous()
        dWC, dQ = ops.gloria_cos_bwd(WC, Q.view(NL, D), cap_lens, cosv, wnorm, qnorm, sims, lse_row, lse_col, g0, g1, B, L,
                                     temp2, temp3)
        # dP2[i,s,c] = c_is . dWC[i,c,:]
        G = ops.gemm(Xc, dWC, out_dtype=torch.float32)                                   # [B, S, NL]
        ops.gloria_region_softmax_bwd(P2, G, B, S, temp1)                                # -> dL/dP1, in place
        dA = ops.gloria_word_softmax_bwd(P1, G.view(B * S, NL), cap_lens, B, L)          # bf16 [B*S, NL]
        # regions: dXc = dA Ww  +  P2_i dWC_i
        dXc = ops.gemm(dA, Ww.view(NL, D), b_mn_major=True, out_dtype=torch.float32)     # [B*S, D]
        ops.gemm(P2h3, dWC, b_mn_major=True, out=dXc.view(B, S, D), accumulate=True)
        # words: dWw = dA^T Xc + dQ(direct)
        dWw = ops.gemm(dA, Xc.view(B * S, D), a_mn_major=True, b_mn_major=True, out=dQ, accumulate=True)   # [NL, D]
        dimg = ops.transpose_cast(dXc.view(B, S, D), torch.float32).view(img_shape)      # [B, D, S] -> [B, D, ih, iw]
        dwords = ops.transpose_cast(dWw.view(B, L, D)[:, :Lw], torch.float32)            # [B, D, Lw]
        return dimg.to(img_dtype), dwords.to(words_dtype), None, None, None, None


def local_loss(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg="sum"):
    """GLoRIALoss.py:78-129 -> (loss0, loss1, att_maps) with att_maps[i] = [1, cap_lens[i], ih, iw]."""
    if agg != "sum":
        raise NotImplementedError("GLoRIA local_loss: only agg='sum' (the reference's default and only call, :160-167)")
    img_features = img_features.cuda()
    words_emb = words_emb.cuda()
    lens = [int(c) for c in (cap_lens.tolist() if torch.is_tensor(cap_lens) else cap_lens)]
    if max(lens) > words_emb.shape[2] or min(lens) < 1:
        raise ValueError("GLoRIA local_loss: cap_lens must lie in [1, words_emb.shape[2]]")
    cl = _cap_lens_tensor(lens, img_features.device)
    loss0, loss1, P2 = _GloriaLocalFn.apply(img_features, words_emb, cl, float(temp1), float(temp2), float(temp3))
    B, S = P2.shape[0], P2.shape[1]
    ih, iw = img_features.shape[2], img_features.shape[3]
    L = P2.shape[2] // B
    P2v = P2.view(B, S, B, L)
    att_maps = [P2v[i, :, i, :lens[i]].t().reshape(1, lens[i], ih, iw).contiguous() for i in range(B)]   # :100-102
    return loss0, loss1, att_maps


def global_loss(cnn_code, rnn_code, eps=1e-8, temp3=10.0):
    """GLoRIALoss.py:54-75 -> (loss0, loss1)."""
    return GLoRIAGlobalLoss(temp3=temp3)(cnn_code, rnn_code)


def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """GLoRIALoss.py:5-10 (kept for callers that import it; the loss itself uses the fused gloria_cos kernel)."""
    w12 = torch.sum(x1 * x2, dim)
    return (w12 / (torch.norm(x1, 2, dim) * torch.norm(x2, 2, dim)).clamp(min=eps)).squeeze()


def gloria_attention_fn(query, context, temp1):
    """GLoRIALoss.py:13-51 for ONE set of queries per image: query [B,D,Lq], context [B,D,ih,iw] ->
    (weightedContext [B,D,Lq], attn [B,Lq,ih,iw]).  Runs the same kernels as the loss with B independent 'captions' by
    evaluating the block-diagonal of the all-pairs tensors is wasteful for this helper, so it uses the batched GEMM
    directly (A_i = Xc_i Ww_i^T) and the two softmax kernels with NB = 1 per image."""
    B, D, Lq = query.shape
    ih, iw = context.shape[2], context.shape[3]
    S = ih * iw
    L = (Lq + 7) // 8 * 8
    dev = context.device
    lens = torch.full((1,), Lq, device=dev, dtype=torch.int32)
    Xc = ops.transpose_cast(context.float().reshape(B, D, S).contiguous(), torch.bfloat16)       # [B,S,D]
    Ww = ops.transpose_cast(query.float().contiguous(), torch.bfloat16, c_out=L)                 # [B,L,D]
    A = ops.gemm(Xc, Ww, out_dtype=torch.float32)                                                # [B,S,L]
    P1 = ops.gloria_word_softmax(A.view(B * S, L), lens, 1, L)
    P2, P2h, _ = ops.gloria_region_softmax(P1, lens, B, S, 1, L, float(temp1))
    WC = ops.gemm(P2h.view(B, S, L), Xc, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)   # [B,L,D]
    weighted = WC[:, :Lq].transpose(1, 2).contiguous()
    attn = P2.view(B, S, L)[:, :, :Lq].transpose(1, 2).reshape(B, Lq, ih, iw).contiguous()
    return weighted, attn


class GLoRIALoss(nn.Module):
    def __init__(self, local_loss_weight=1.0, global_loss_weight=1.0, temp1=4.0, temp2=5.0, temp3=10.0, **kwargs):
        super().__init__()
        self.local_loss_weight = local_loss_weight
        self.global_loss_weight = global_loss_weight
        self.temp1 = temp1
        self.temp2 = temp2
        self.temp3 = temp3

    def forward(self, global_features, local_features, word_embeddings, sent_embeddings, sents):
        l_loss0, l_loss1, attn_maps = self._calc_local_loss(local_features, word_embeddings, sents)
        g_loss0, g_loss1 = self._calc_global_loss(global_features, sent_embeddings)
        loss = (l_loss0 + l_loss1) * self.local_loss_weight + (g_loss0 + g_loss1) * self.global_loss_weight   # :149-151
        return loss, attn_maps

    def _calc_local_loss(self, img_emb_l, text_emb_l, sents):
        cap_lens = [len([w for w in sent if not w.startswith("[")]) + 1 for sent in sents]                   # :155-157
        return local_loss(img_emb_l, text_emb_l, cap_lens, temp1=self.temp1, temp2=self.temp2, temp3=self.temp3)

    def _calc_global_loss(self, img_emb_g, text_emb_g):
        return global_loss(img_emb_g, text_emb_g, temp3=self.temp3)
