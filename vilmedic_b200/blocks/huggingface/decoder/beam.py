"""Beam search / greedy decoding over one or several decoder towers (ensemble = sum of next-token logits).

Follows the loop the reference intends for bin/ensemble.py — vilmedic/blocks/huggingface/decoder/beam_search.py:222-342:
  per step: every model's next-token logits are SUMMED (:254), log_softmax (:260-262), + running beam score (:265),
  top-(2*num_beams) over num_beams*V (:289-294), BeamSearchScorer.process semantics (:297-304; eos candidates beyond
  rank num_beams are skipped, finished hypotheses scored sum_logprobs / len**length_penalty), reorder by beam index.
With one model and num_beams=1 this is greedy argmax decoding.

Default path (`DeviceSearch`): the whole step — 12-layer KV-cached decoder step of every ensemble member, logit sum, log-softmax,
top-2k, hypothesis bookkeeping, token / cache-indirection update — runs as kernels (csrc/decode.cu) whose per-step scalars live in
device memory, captured ONCE as a CUDA graph and replayed per generated token; the host only replays and, every few steps, reads
one int32 to see whether every batch element has finished.  Finished-hypothesis selection at the end (max over <= k stored
hypotheses per image) is host code over a few KB.  `use_cache=False` (or a prompt longer than one token) runs the plain host loop
over full-prefix recomputed logits — the cross-check the tests use; the arithmetic of both is in the same kernels.
"""
import torch

from .... import ops


class _Hyps:
    def __init__(self, k, length_penalty):
        self.k, self.lp = k, length_penalty
        self.beams = []
        self.worst = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / (hyp.shape[-1] ** self.lp)
        if len(self) < self.k or score > self.worst:
            self.beams.append((score, hyp))
            if len(self) > self.k:
                srt = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[srt[0][1]]
                self.worst = srt[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.k:
            return False
        return self.worst >= best_sum_logprobs / cur_len ** self.lp


@torch.no_grad()
def beam_search(models, encs, masks, input_ids, max_length, num_beams, bos_token_id, eos_token_id, pad_token_id,
                length_penalty=1.0, use_cache=True, sampling=None, forced_last=-1):
    """sampling: None (greedy / beam) or dict(top_k, bad_ids, temperature, seed) -> multinomial rollouts (SCST.py:139-153)."""
    if sampling is not None and (num_beams != 1 or len(models) != 1 or not use_cache or input_ids.shape[1] != 1):
        raise NotImplementedError("sampling is supported for a single model, num_beams=1, from a one-token prompt")
    if forced_last >= 0 and num_beams != 1:
        raise NotImplementedError("forced_eos_token_id is supported for num_beams=1 only")
    if use_cache and input_ids.shape[1] == 1:
        return DeviceSearch.get(models, input_ids.shape[0], num_beams, max_length).run(
            encs, masks, input_ids, eos_token_id, pad_token_id, length_penalty, sampling=sampling, forced_last=forced_last)
    if forced_last >= 0:
        raise NotImplementedError("forced_eos_token_id needs the cached device search")
    return _beam_search_host(models, encs, masks, input_ids, max_length, num_beams, bos_token_id, eos_token_id, pad_token_id,
                             length_penalty)


class SearchState:
    """Device buffers of one search (ids, scores, finished hypotheses, cache indirection, counters) + the selection step and the
    final pick.  Separate from the model step so that the selection kernels can be driven by any source of logits."""

    def __init__(self, batch, beams, max_len, vocab, device):
        if not 1 <= beams <= 8:
            raise NotImplementedError("num_beams must be in 1..8 (got %d)" % beams)
        self.B, self.k, self.L, self.V = batch, beams, max_len, vocab
        R = self.R = batch * beams
        dev = device
        i32, i64, f32 = torch.int32, torch.int64, torch.float32
        k = beams
        self.st = st = {
            "ids": torch.zeros((R, max_len), device=dev, dtype=i64), "ids_tmp": torch.zeros((R, max_len), device=dev, dtype=i64),
            "row_map": torch.zeros((R, max_len), device=dev, dtype=i32), "map_tmp": torch.zeros((R, max_len), device=dev, dtype=i32),
            "beam_scores": torch.zeros(R, device=dev, dtype=f32), "done": torch.zeros(batch, device=dev, dtype=torch.uint8),
            "next_tok": torch.zeros(R, device=dev, dtype=i64), "parent": torch.zeros(R, device=dev, dtype=i32),
            "cand_score": torch.zeros((R, 2 * k), device=dev, dtype=f32), "cand_tok": torch.zeros((R, 2 * k), device=dev, dtype=i32),
            "counters": torch.zeros(4, device=dev, dtype=i32),
        }
        if k > 1:
            st.update({"hyp_score": torch.zeros((batch, k), device=dev, dtype=torch.float64),
                       "hyp_len": torch.zeros((batch, k), device=dev, dtype=i32),
                       "hyp_tok": torch.zeros((batch, k, max_len), device=dev, dtype=i64),
                       "hyp_count": torch.zeros(batch, device=dev, dtype=i32),
                       "hyp_worst": torch.zeros(batch, device=dev, dtype=torch.float64)})
        self.consts = None

    def reset(self, bos, eos, pad, lp):
        st = self.st
        for key in ("ids", "row_map", "done", "parent", "counters", "cand_score", "cand_tok"):
            st[key].zero_()
        st["ids"][:, 0] = bos
        st["next_tok"].fill_(bos)
        bs = torch.zeros((self.B, self.k), dtype=torch.float32)
        bs[:, 1:] = -1e9                                               # only beam 0 of every image is live at the start (:232-233)
        st["beam_scores"].copy_(bs.view(-1))
        if self.k > 1:
            st["hyp_count"].zero_()
            st["hyp_worst"].fill_(1e9)
            st["hyp_score"].zero_()
        self.consts = (int(eos), int(pad), float(lp))

    def select(self, logits, sampling=None, forced_last=-1, advance=True):
        """logits: list of fp32 [rows, ld >= V] next-token logits (one per ensemble member) for the tokens in st['next_tok'].
        sampling = dict(top_k, bad_ids, temperature, seed, offset): filter + draw instead of arg-max (k must be 1)."""
        st = self.st
        eos, pad, lp = self.consts
        if sampling is None:
            ops.beam_rows(logits, self.V, st["beam_scores"], st["cand_score"], st["cand_tok"], self.k)
        else:
            lg = logits[0]
            if sampling.get("bad_ids") or sampling.get("top_k"):
                ops.logits_filter(lg, self.V, sampling.get("bad_ids") or (), sampling.get("top_k") or 0)
            ops.sample_rows(lg, self.V, st["cand_score"], st["cand_tok"], 0x5C57, 0, st["counters"], sampling.get("temperature", 1.0))
        ops.beam_select(st, self.k, self.V, self.B, self.L, eos, pad, lp, forced_last)
        if advance:
            self.advance()

    def advance(self):
        """ids / cache indirection follow the decisions in st['parent'] / st['next_tok']; the step counter advances."""
        ops.beam_advance(self.st, self.R, self.L)

    def all_done(self):
        return int(self.st["counters"][1].item()) >= self.B

    def finish(self):
        st, B, k, L = self.st, self.B, self.k, self.L
        eos, pad, lp = self.consts
        counters = st["counters"].cpu().tolist()
        ids = st["ids"].cpu()
        all_done_len = counters[2] if counters[1] >= B else 0
        if k == 1:
            return st["ids"][:, :(all_done_len or L)].clone()
        cur_len = counters[0] + 1                                          # tokens per live beam when the search stopped
        done = st["done"].cpu().tolist()
        scores = st["beam_scores"].cpu().tolist()
        h_score, h_len = st["hyp_score"].cpu().tolist(), st["hyp_len"].cpu().tolist()
        h_tok, h_cnt, h_worst = st["hyp_tok"].cpu(), st["hyp_count"].cpu().tolist(), st["hyp_worst"].cpu().tolist()
        out = []
        for b in range(B):
            hy = _Hyps(k, lp)
            hy.beams = [(h_score[b][i], h_tok[b, i, :h_len[b][i]]) for i in range(h_cnt[b])]
            hy.worst = h_worst[b]
            if not done[b]:                                                # only possible when the search ran to max_length
                for j in range(k):
                    hy.add(ids[b * k + j, :cur_len], scores[b * k + j])
            out.append(max(hy.beams, key=lambda x: x[0])[1])
        Lo = min(max(len(o) for o in out) + 1, L)
        res = torch.full((B, Lo), pad, dtype=torch.long)
        for b, o in enumerate(out):
            res[b, :len(o)] = o
            if len(o) < L:
                res[b, len(o)] = eos
        return res.to(st["ids"].device)


class DeviceSearch:
    """Device-resident greedy / beam / ensemble search: static buffers + one captured CUDA graph per (models, batch, beams, max_len)."""

    _cache = {}
    CHECK_EVERY = 4          # steps between reads of the device-side "all finished" counter

    @classmethod
    def get(cls, models, batch, beams, max_len):
        key = (tuple(id(m) for m in models), batch, beams, max_len)
        eng = cls._cache.get(key)
        if eng is None or any(not st.arena.valid() for st in eng.states):
            if len(cls._cache) >= 4:                      # bounded: each engine pins KV caches
                cls._cache.pop(next(iter(cls._cache)))
            eng = cls._cache[key] = cls(models, batch, beams, max_len)
        return eng

    def __init__(self, models, batch, beams, max_len):
        from .generation import DecodeState
        self.models = list(models)
        dev = next(models[0].parameters()).device
        V = models[0].cfg.vocab_size
        for m in models:
            if m.cfg.vocab_size != V:
                raise ValueError("ensemble members must share the vocabulary")
        self.search = SearchState(batch, beams, max_len, V, dev)
        st = self.search.st
        self.states = [DecodeState(m, batch, beams, max_len, st["row_map"], st["counters"][0:1]) for m in models]
        self.graph = None
        self.sig = None
        self.mode = (None, -1)

    def _step(self, advance=True):
        """One search step.  The ensemble members are independent until their logits are summed: each runs on its own stream, so
        that inside the captured graph they form parallel branches (a decode step is ~300 small, launch-latency-bound kernels per
        model; two branches overlap them)."""
        st = self.search.st
        cur = torch.cuda.current_stream()
        if len(self.models) == 1:
            logits = [self.models[0].decode_step(self.states[0], st["next_tok"])]
        else:
            if not hasattr(self, "_streams"):
                self._streams = [torch.cuda.Stream() for _ in self.models[1:]]
            logits = [None] * len(self.models)
            for i, side in enumerate(self._streams, start=1):
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    logits[i] = self.models[i].decode_step(self.states[i], st["next_tok"])
            logits[0] = self.models[0].decode_step(self.states[0], st["next_tok"])
            for side in self._streams:
                cur.wait_stream(side)
        self.search.select(logits, sampling=self.mode[0], forced_last=self.mode[1], advance=advance)

    @torch.no_grad()
    def run(self, encs, masks, input_ids, eos, pad, length_penalty, sampling=None, forced_last=-1):
        bos = int(input_ids[0, 0])
        self.mode = (dict(sampling) if sampling is not None else None, int(forced_last))
        if not bool((input_ids == bos).all()):
            raise NotImplementedError("per-row start tokens are not supported")
        fresh = False
        for s, e, mk in zip(self.states, encs, masks):
            fresh = bool(s.set_encoder(e, mk)) or fresh
        sig = (int(eos), int(pad), float(length_penalty), tuple(s.enc_len for s in self.states),
               tuple(s.enc_mask is not None for s in self.states), int(forced_last),
               None if sampling is None else tuple(sorted((k_, str(v_)) for k_, v_ in sampling.items() if k_ != "seed")))
        if self.graph is None or fresh or sig != self.sig:
            self.search.reset(bos, eos, pad, length_penalty)
            self._step()                                                # eager warm-up (first-use attribute setup), then reset
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(self.graph, stream=side):
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            self.sig = sig
        self.search.reset(bos, eos, pad, length_penalty)
        if sampling is not None:                     # per-rollout nonce of the draw (device memory: the captured graph is reused)
            self.search.st["counters"][3] = int(sampling.get("seed", 0)) & 0x7FFFFFFF
        for step in range(1, self.search.L):
            self.graph.replay()
            if step % self.CHECK_EVERY == 0 and self.search.all_done():
                break
        return self.search.finish()


@torch.no_grad()
def _beam_search_host(models, encs, masks, input_ids, max_length, num_beams, bos_token_id, eos_token_id, pad_token_id,
                      length_penalty=1.0):
    """Plain host loop over full-prefix recomputed logits (no cache): the reference formulation, used as the cross-check."""
    dev = input_ids.device
    B = input_ids.shape[0]
    k = num_beams
    ids = input_ids.repeat_interleave(k, dim=0)                       # [B*k, cur]
    encs_k = [e.repeat_interleave(k, dim=0) if e is not None else None for e in encs]
    masks_k = [m.repeat_interleave(k, dim=0) if m is not None else None for m in masks]
    beam_scores = torch.zeros((B, k), dtype=torch.float32, device=dev)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [_Hyps(k, length_penalty) for _ in range(B)]
    done = [False] * B
    cur_len = ids.shape[1]
    while cur_len < max_length:
        logits = None
        for m, e, mk in zip(models, encs_k, masks_k):
            l = m.next_token_logits(ids, e, mk)
            logits = l if logits is None else logits + l
        V = logits.shape[-1]
        scores = torch.log_softmax(logits.float(), dim=-1) + beam_scores[:, None]
        if k == 1:
            nxt_score, nxt_tok = scores.max(dim=-1)
            finished = torch.tensor(done, device=dev)
            nxt_tok = torch.where(finished, torch.full_like(nxt_tok, pad_token_id), nxt_tok)
            ids = torch.cat([ids, nxt_tok[:, None]], dim=1)
            beam_scores = nxt_score
            cur_len += 1
            for b, t in enumerate(nxt_tok.tolist()):
                if t == eos_token_id:
                    done[b] = True
            if all(done):
                break
            continue
        top_s, top_i = torch.topk(scores.view(B, k * V), 2 * k, dim=1, largest=True, sorted=True)
        top_beam = (top_i // V).tolist()
        top_tok = (top_i % V).tolist()
        top_sl = top_s.tolist()
        new_scores = torch.zeros((B, k), dtype=torch.float32)
        new_tok = torch.zeros((B, k), dtype=torch.long)
        new_idx = torch.zeros((B, k), dtype=torch.long)
        for b in range(B):
            if done[b]:
                new_tok[b].fill_(pad_token_id)
                new_idx[b] = torch.arange(k) + b * k
                continue
            j = 0
            for rank, (bi, tk, sc) in enumerate(zip(top_beam[b], top_tok[b], top_sl[b])):
                src = b * k + bi
                if tk == eos_token_id:
                    if rank >= k:
                        continue
                    hyps[b].add(ids[src].clone(), sc)
                else:
                    new_scores[b, j], new_tok[b, j], new_idx[b, j] = sc, tk, src
                    j += 1
                if j == k:
                    break
            done[b] = done[b] or hyps[b].is_done(max(top_sl[b]), cur_len)
        beam_scores = new_scores.view(-1).to(dev)
        idx = new_idx.view(-1).to(dev)
        ids = torch.cat([ids[idx], new_tok.view(-1, 1).to(dev)], dim=1)
        cur_len += 1
        if all(done):
            break
    if k == 1:
        return ids
    out = []
    for b in range(B):
        if not done[b]:
            for j in range(k):
                hyps[b].add(ids[b * k + j], beam_scores[b * k + j].item())
        best = max(hyps[b].beams, key=lambda x: x[0])[1]
        out.append(best)
    L = min(max(len(o) for o in out) + 1, max_length)
    res = torch.full((B, L), pad_token_id, dtype=torch.long, device=dev)
    for b, o in enumerate(out):
        res[b, :len(o)] = o
        if len(o) < max_length:
            res[b, len(o)] = eos_token_id
    return res
