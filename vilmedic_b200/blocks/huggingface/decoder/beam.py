"""Beam search / greedy decoding over one or several decoder towers (ensemble = sum of next-token logits).

Follows the loop the reference intends for bin/ensemble.py — vilmedic/blocks/huggingface/decoder/beam_search.py:222-342:
  per step: every model's next-token logits are SUMMED (:254), log_softmax (:260-262), + running beam score (:265),
  top-(2*num_beams) over num_beams*V (:289-294), BeamSearchScorer.process semantics (:297-304; eos candidates beyond
  rank num_beams are skipped, finished hypotheses scored sum_logprobs / len**length_penalty), reorder by beam index.
With one model and num_beams=1 this is greedy argmax decoding.

Round-1 status: the model math runs on the sm_100a kernels through the KV-cached single-token step
(generation.DecodeState / decode_step; `use_cache=False` falls back to full-prefix recompute as a cross-check); the
per-step selection (log_softmax / top-k / hypothesis bookkeeping over [B*k, V] fp32 logits) is still torch glue — the
fused log-softmax+top-2k kernel is the next item for this file.
"""
import torch


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
                length_penalty=1.0, use_cache=True):
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
    states = None
    if use_cache:
        from .generation import DecodeState
        states = [DecodeState(m, B * k, max_length, e, mk) for m, e, mk in zip(models, encs_k, masks_k)]
        for t in range(cur_len - 1):                 # prime the caches with the prompt (normally just BOS: nothing to do)
            for m, st in zip(models, states):
                m.decode_step(st, ids[:, t])
    while cur_len < max_length:
        logits = None
        if states is not None:
            for m, st in zip(models, states):
                l = m.decode_step(st, ids[:, -1])
                logits = l if logits is None else logits + l
        else:
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
        if states is not None:
            for st in states:
                st.reorder(idx)
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
