"""Numerics-policy oracle for decoding.  TEST INFRASTRUCTURE ONLY (tests/ and __graft_entry__.smoke() may import it; the
product never does).

A plain-torch restatement of the HF BertGenerationDecoder forward that vilmedic's DecoderModel wraps
(vilmedic/blocks/huggingface/decoder/decoder_model.py:23-26,42-47; arithmetic HF:bert_generation/modeling_bert_generation.py:
embeddings 395-429, self-attention 89-153, cross-attention 181-232, self-output 52-56, FFN 265-293, layer 296-360, LM head 593-601),
parameterised by WHERE values are rounded to bf16:

  policy="fp32" : no rounding anywhere — must reproduce HF's own logits (tests/test_cpu.py pins this against the HF module),
  policy="bf16" : the storage policy of the sm_100a kernels — weights, activations between kernels and attention outputs are bf16,
                  every matmul accumulates in fp32, LayerNorm / softmax / GELU / bias / residual arithmetic is fp32 and the result
                  is rounded ONCE where a kernel writes it (gemm_epilogue.cuh, layernorm.cu, decode.cu).

north_star asks for greedy token ids bit-exact against the reference path.  An fp32 oracle and a bf16 implementation differ by the
rounding itself (1-4 % of the logit scale on the scaled random models of the tests), which no implementation detail can remove;
SURVEY.md §7 "hard parts" therefore plans exactly this: run the oracle under the same rounding policy.  What is left between this
oracle and the kernels is fp32 summation order and the transcendental approximations (A&S erfc in GELU, __expf in the softmax),
~1e-6 relative — far below the decision margins — so token ids are compared bit for bit with NO margin escape.
Full-prefix recompute (no cache): K/V of earlier positions are recomputed from identical inputs, hence identical.
"""
import math

import torch
import torch.nn.functional as F


def _rb(x, on):
    return x.to(torch.bfloat16).float() if on else x


class PolicyDecoder:
    """Callable like the HF decoder the oracle search loop drives (oracle/decode.py: next_logits): returns an object whose
    `.logits` is [rows, 1, V] — the LAST position only (the only one the search reads)."""

    def __init__(self, hf_decoder, policy="bf16", device="cpu"):
        assert policy in ("bf16", "fp32")
        self.on = policy == "bf16"
        self.config = hf_decoder.config
        self.dev = torch.device(device)
        sd = {k: v.detach().float().to(self.dev) for k, v in hf_decoder.state_dict().items()}
        self.sd = sd
        c = self.config
        self.H = c.num_attention_heads
        self.D = c.hidden_size
        self.L = c.num_hidden_layers
        self.eps = c.layer_norm_eps
        # GEMM weights are read from the bf16 mirror of the arena; biases / LayerNorm / embedding tables from the fp32 masters
        self.w = {k: _rb(v, self.on) for k, v in sd.items() if k.endswith(".weight") and v.dim() == 2 and "embeddings" not in k}
        self.E = _rb(sd["bert.embeddings.word_embeddings.weight"], self.on)          # tied LM head operand

    def _lin(self, x, name):
        return x @ self.w[name + ".weight"].t() + self.sd[name + ".bias"]

    def _ln(self, x, name):
        return F.layer_norm(x, (self.D,), self.sd[name + ".weight"], self.sd[name + ".bias"], self.eps)

    def _attn(self, q, k, v, mask):
        """q [R, Tq, D], k/v [R, S, D]; mask additive [R, 1, Tq, S] or None; fp32 softmax, probabilities NOT rounded."""
        R, Tq, _ = q.shape
        S = k.shape[1]
        H, dh = self.H, self.D // self.H
        qh = q.view(R, Tq, H, dh).transpose(1, 2)
        kh = k.view(R, S, H, dh).transpose(1, 2)
        vh = v.view(R, S, H, dh).transpose(1, 2)
        s = (qh @ kh.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
        if mask is not None:
            s = s + mask
        p = torch.softmax(s, dim=-1)
        return (p @ vh).transpose(1, 2).reshape(R, Tq, self.D)

    @torch.no_grad()
    def __call__(self, input_ids, encoder_hidden_states=None, encoder_attention_mask=None, use_cache=False):
        on = self.on
        ids = input_ids.to(self.dev)
        R, T = ids.shape
        sd = self.sd
        z = _rb(sd["bert.embeddings.word_embeddings.weight"][ids] + sd["bert.embeddings.position_embeddings.weight"][:T][None], on)
        x = _rb(self._ln(z, "bert.embeddings.LayerNorm"), on)
        causal = torch.full((T, T), float("-inf"), device=self.dev).triu(1)[None, None]
        enc = emask = None
        if encoder_hidden_states is not None:
            enc = _rb(encoder_hidden_states.float().to(self.dev), on)
            if enc.shape[0] != R:                                           # one encoder row per image, rows = images * beams
                enc = enc.repeat_interleave(R // enc.shape[0], 0)
            if encoder_attention_mask is not None:
                m = encoder_attention_mask.to(self.dev)
                if m.shape[0] != R:
                    m = m.repeat_interleave(R // m.shape[0], 0)
                emask = torch.zeros(m.shape, device=self.dev).masked_fill(m == 0, float("-inf"))[:, None, None, :]
        for i in range(self.L):
            p = "bert.encoder.layer.%d." % i
            q = _rb(self._lin(x, p + "attention.self.query"), on)
            k = _rb(self._lin(x, p + "attention.self.key"), on)
            v = _rb(self._lin(x, p + "attention.self.value"), on)
            ctx = _rb(self._attn(q, k, v, causal), on)
            z1 = _rb(self._lin(ctx, p + "attention.output.dense") + x, on)
            x1 = _rb(self._ln(z1, p + "attention.output.LayerNorm"), on)
            if enc is not None:
                qc = _rb(self._lin(x1, p + "crossattention.self.query"), on)
                kc = _rb(self._lin(enc, p + "crossattention.self.key"), on)
                vc = _rb(self._lin(enc, p + "crossattention.self.value"), on)
                ctx2 = _rb(self._attn(qc, kc, vc, emask), on)
                z2 = _rb(self._lin(ctx2, p + "crossattention.output.dense") + x1, on)
                x2 = _rb(self._ln(z2, p + "crossattention.output.LayerNorm"), on)
            else:
                x2 = x1
            h = _rb(F.gelu(self._lin(x2, p + "intermediate.dense")), on)
            z3 = _rb(self._lin(h, p + "output.dense") + x2, on)
            x = _rb(self._ln(z3, p + "output.LayerNorm"), on)
        logits = x[:, -1] @ self.E.t() + sd["lm_head.bias"]

        class _Out:
            pass
        o = _Out()
        o.logits = logits[:, None, :].float().cpu()
        return o
