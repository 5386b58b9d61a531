"""Incremental decoding for the B200 decoder tower: a KV-cached single-token step built from
the same kernels as training — per layer: q / kv projections (the kv GEMM writes straight into the cache slot of the
current position), fused attention over the cached keys (T_q = 1), cross-attention over K/V projected ONCE per
sequence, FFN; then the tied LM head on the last hidden state.  Mirrors HF's cached `generate` loop used by
vilmedic/blocks/huggingface/decoder/evaluation.py:73-78.  HBM-bound by design: every step streams the decoder
weights (275 MB bf16 at BERT-base) plus the KV cache.
"""
import torch

from .... import ops
from ....arena import get_arena
from ....nn import _ln, _lin, _prepare, _root_of


class DecodeState:
    """Per-model state of one search over `rows` = batch * beams sequences.

    HBM layout: per layer one self-attention cache bf16 [rows, max_len, 2D] ([K | V] per position) that is only ever APPENDED to —
    beam reordering permutes the shared int32 table `row_map` [rows, max_len] (which physical row holds position j of row r's
    history) instead of copying caches (beam_search.py:317-319 does an index_select per layer per step); cross-attention K/V are
    projected once per IMAGE, bf16 [batch, S_enc, 2D] per layer, and shared by the image's beams."""

    def __init__(self, tower, batch, beams, max_len, row_map, t_ptr):
        cfg = tower.cfg
        self.tower = tower
        self.batch, self.beams, self.rows, self.max_len = batch, beams, batch * beams, max_len
        self.row_map, self.t_ptr = row_map, t_ptr
        D = cfg.hidden_size
        self.arena = get_arena(_root_of(tower))
        dev = self.arena.device
        _prepare(self.arena, tower)
        self.self_kv = [torch.empty((self.rows, max_len, 2 * D), device=dev, dtype=torch.bfloat16) for _ in tower._core.encoder.layer]
        self.cross_kv = []
        self.enc_mask = None
        self.enc_len = 0

    def set_encoder(self, enc, enc_mask):
        """enc [batch, S_enc, D_enc] (one row per image): project the cross-attention K/V of every layer into static buffers."""
        tower = self.tower
        if enc is None:
            self.cross_kv, self.enc_mask, self.enc_len = [], None, 0
            return
        _prepare(self.arena, tower)
        e = enc if enc.dtype == torch.bfloat16 else ops.cast_bf16(enc.float().contiguous())
        Bn, Se = e.shape[0], e.shape[1]
        if Bn != self.batch:
            raise ValueError("encoder states for %d images, search over %d" % (Bn, self.batch))
        D = tower.cfg.hidden_size
        e2 = e.reshape(Bn * Se, e.shape[2]).contiguous()
        fresh = not self.cross_kv or self.enc_len != Se
        if fresh:
            self.cross_kv = [torch.empty((Bn, Se, 2 * D), device=e.device, dtype=torch.bfloat16) for _ in tower._core.encoder.layer]
        for li, layer in enumerate(tower._core.encoder.layer):
            ca = layer.crossattention.self
            Pkv = _lin(self.arena, ca.key, ca.value)
            ops.gemm(e2, Pkv.w, bias=Pkv.b, out=self.cross_kv[li].view(Bn * Se, 2 * D))
        self.enc_len = Se
        if enc_mask is not None:
            m = (enc_mask != 0).to(torch.uint8).contiguous()
            if self.enc_mask is None or fresh:
                self.enc_mask = m
            else:
                self.enc_mask.copy_(m)
        else:
            self.enc_mask = None
        return fresh


class GenerationMixinB200:
    @torch.no_grad()
    def decode_step(self, state, tokens):
        """tokens int64 [rows] = the token of every row at position *state.t_ptr -> fp32 next-token logits [rows, V].
        Appends the step's keys / values to the caches; every per-step scalar is read from device memory (graph-replayable)."""
        cfg = self.cfg
        arena = state.arena
        D, H = cfg.hidden_size, cfg.num_attention_heads
        DH = D // H
        eps = cfg.layer_norm_eps
        core = self._core
        emb = core.embeddings
        tt = getattr(emb, "token_type_embeddings", None)
        z = ops.embed_step(tokens, arena.fp32(emb.word_embeddings.weight), arena.fp32(emb.position_embeddings.weight), state.t_ptr,
                           pos_shift=(cfg.pad_token_id + 1) if cfg.family == "roberta" else 0,
                           tt_row=arena.fp32(tt.weight)[0].contiguous() if tt is not None else None)
        lnp = _ln(arena, emb.LayerNorm)
        x, _, _ = ops.layernorm_fwd(z, lnp[0], lnp[1], eps, save_stats=False)
        for li, layer in enumerate(core.encoder.layer):
            sa = layer.attention
            Pqkv = _lin(arena, sa.self.query, sa.self.key, sa.self.value)
            Po = _lin(arena, sa.output.dense)
            ln1 = _ln(arena, sa.output.LayerNorm)
            qkv = ops.gemm(x, Pqkv.w, bias=Pqkv.b)                                   # [rows, 3D]: q | k | v of the new token
            ctx = ops.decode_attention(qkv[:, :D], H, DH, state.self_kv[li], kv_new=qkv[:, D:], row_map=state.row_map,
                                       t_ptr=state.t_ptr, max_len=state.max_len)
            z1 = ops.gemm(ctx, Po.w, bias=Po.b, residual=x)
            x1, _, _ = ops.layernorm_fwd(z1, ln1[0], ln1[1], eps, save_stats=False)
            if state.cross_kv:
                ca = layer.crossattention
                Pqc, Poc = _lin(arena, ca.self.query), _lin(arena, ca.output.dense)
                ln2 = _ln(arena, ca.output.LayerNorm)
                qc = ops.gemm(x1, Pqc.w, bias=Pqc.b)
                ctx2 = ops.decode_attention(qc, H, DH, state.cross_kv[li], fixed_len=state.enc_len, row_div=state.beams,
                                            kmask=state.enc_mask)
                z2 = ops.gemm(ctx2, Poc.w, bias=Poc.b, residual=x1)
                x2, _, _ = ops.layernorm_fwd(z2, ln2[0], ln2[1], eps, save_stats=False)
            else:
                x2 = x1
            P1, P2 = _lin(arena, layer.intermediate.dense), _lin(arena, layer.output.dense)
            ln3 = _ln(arena, layer.output.LayerNorm)
            h = ops.gemm(x2, P1.w, bias=P1.b, act=ops.ACT_GELU)
            z3 = ops.gemm(h, P2.w, bias=P2.b, residual=x2)
            x, _, _ = ops.layernorm_fwd(z3, ln3[0], ln3[1], eps, save_stats=False)
        return self.lm_logits(x, padded=True)

    @torch.no_grad()
    def next_token_logits(self, input_ids, encoder_hidden_states=None, encoder_attention_mask=None):
        """fp32 logits [B, V] of the last position by full-prefix recompute (no cache; used by tests as a cross-check)."""
        x, B, T = self.hidden_states(input_ids, None, encoder_hidden_states, encoder_attention_mask)
        last = x.view(B, T, -1)[:, -1].contiguous()
        return self.lm_logits(last)

    @torch.no_grad()
    def generate(self, input_ids=None, encoder_hidden_states=None, encoder_attention_mask=None, max_length=None,
                 num_beams=None, bos_token_id=None, eos_token_id=None, pad_token_id=None, length_penalty=None,
                 ensemble=None, use_cache=True, generation_config=None, do_sample=None, top_k=None, temperature=None,
                 bad_words_ids=None, forced_eos_token_id=None, return_dict_in_generate=None, seed=None, **kwargs):
        """Greedy / beam decoding (ensemble = sum of next-token logits over `ensemble`).  Arguments may come individually or in a
        HF-style `generation_config` (any object or dict with max_length / num_beams / length_penalty / *_token_id attributes — the
        call of vilmedic/blocks/huggingface/decoder/evaluation.py:73-78 and vision_multi_evaluation.py:42-57); explicit keyword
        arguments win, then generation_config, then the model config.  Unknown arguments raise instead of being ignored.
        do_sample=True draws multinomial rollouts (HF sample(): NoBadWords -> TopK -> temperature-1 softmax draw) on the device —
        the call of vilmedic/blocks/rl/SCST.py:139-153; `bad_words_ids` must be single-token lists; `forced_eos_token_id` is emitted
        at the last position (an int; the reference passes True, which HF — and this code — read as token id 1).
        return_dict_in_generate=True returns an object with `.sequences` (scores are not materialised: use
        blocks.rl.sequence_log_probs for the log-probabilities of a rollout)."""
        from .beam import beam_search
        gc = generation_config

        def pick(name, explicit, default):
            if explicit is not None:
                return explicit
            if gc is not None:
                v = gc.get(name) if isinstance(gc, dict) else getattr(gc, name, None)
                if v is not None:
                    return v
            return default

        harmless = {"num_return_sequences": 1, "output_scores": None, "early_stopping": False, "decoder_start_token_id": None}
        for k_, v_ in kwargs.items():
            if k_ not in harmless:
                raise TypeError("generate() got an unsupported argument %r" % k_)
            if k_ in ("num_return_sequences", "early_stopping") and v_ not in (harmless[k_], None):
                raise NotImplementedError("generate(%s=%r) is not supported" % (k_, v_))
        do_sample = bool(pick("do_sample", do_sample, False))
        if gc is not None:
            for k_ in ("early_stopping",):
                v_ = gc.get(k_) if isinstance(gc, dict) else getattr(gc, k_, None)
                if v_:
                    raise NotImplementedError("generation_config.%s=%r is not supported" % (k_, v_))
            nrs = gc.get("num_return_sequences") if isinstance(gc, dict) else getattr(gc, "num_return_sequences", None)
            if nrs not in (None, 1):
                raise NotImplementedError("num_return_sequences=%r is not supported" % (nrs,))
        models = [self] if ensemble is None else list(ensemble)
        enc = encoder_hidden_states if isinstance(encoder_hidden_states, (list, tuple)) else [encoder_hidden_states] * len(models)
        msk = encoder_attention_mask if isinstance(encoder_attention_mask, (list, tuple)) else [encoder_attention_mask] * len(models)
        cfg = self.config
        sampling = None
        if do_sample:
            bad = []
            for w in (pick("bad_words_ids", bad_words_ids, None) or []):
                if len(w) != 1:
                    raise NotImplementedError("multi-token bad_words_ids are not supported")
                bad.append(int(w[0]))
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())           # follows torch's global generator (torch.manual_seed)
            sampling = dict(top_k=int(pick("top_k", top_k, 0) or 0), bad_ids=tuple(bad), temperature=float(pick("temperature", temperature, 1.0)),
                            seed=int(seed))
        forced = pick("forced_eos_token_id", forced_eos_token_id, None)
        out = beam_search(models, enc, msk, input_ids=input_ids, max_length=pick("max_length", max_length, 20),
                          num_beams=pick("num_beams", num_beams, 1),
                          bos_token_id=pick("bos_token_id", bos_token_id, cfg.bos_token_id),
                          eos_token_id=pick("eos_token_id", eos_token_id, cfg.eos_token_id),
                          pad_token_id=pick("pad_token_id", pad_token_id, cfg.pad_token_id),
                          length_penalty=pick("length_penalty", length_penalty, 1.0), use_cache=use_cache, sampling=sampling,
                          forced_last=-1 if forced is None else int(forced))
        if pick("return_dict_in_generate", return_dict_in_generate, False):
            class _GenerateOutput:
                pass
            o = _GenerateOutput()
            o.sequences = out
            o.scores = None
            return o
        return out
