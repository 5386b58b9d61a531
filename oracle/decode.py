"""Oracle decoding: (i) HF `generate` on the reference decoder (what vilmedic/blocks/huggingface/decoder/evaluation.py:73-78
calls), with the transformers-5.5 cache workaround noted in SURVEY.md §8c; (ii) a restatement of the reference's ensemble
beam-search loop (vilmedic/blocks/huggingface/decoder/beam_search.py:222-342, legacy BeamSearchScorer semantics) driven by
oracle logits — full-prefix recompute, no cache.  TEST INFRASTRUCTURE ONLY."""
import torch


@torch.no_grad()
def hf_generate(decoder, enc, enc_mask, num_beams, max_length, bos, eos, pad, length_penalty=1.0):
    from transformers import GenerationConfig
    from transformers.cache_utils import DynamicCache, EncoderDecoderCache
    B = enc.shape[0]
    cfg = decoder.config
    gen = GenerationConfig(bos_token_id=bos, eos_token_id=eos, pad_token_id=pad, num_return_sequences=1, max_length=max_length,
                           use_cache=True, num_beams=num_beams, length_penalty=length_penalty, do_sample=False)
    return decoder.generate(input_ids=torch.full((B, 1), bos, dtype=torch.long), generation_config=gen,
                            encoder_hidden_states=enc, encoder_attention_mask=enc_mask,
                            past_key_values=EncoderDecoderCache(DynamicCache(config=cfg), DynamicCache(config=cfg)))


@torch.no_grad()
def next_logits(decoder, ids, enc, enc_mask):
    out = decoder(input_ids=ids, encoder_hidden_states=enc, encoder_attention_mask=enc_mask, use_cache=False)
    return out.logits[:, -1, :].float()


class _Hyps:
    def __init__(self, k, lp):
        self.k, self.lp, self.beams, self.worst = k, lp, [], 1e9

    def add(self, hyp, s):
        score = s / (hyp.shape[-1] ** self.lp)
        if len(self.beams) < self.k or score > self.worst:
            self.beams.append((score, hyp))
            if len(self.beams) > self.k:
                srt = sorted((sc, i) for i, (sc, _) in enumerate(self.beams))
                del self.beams[srt[0][1]]
                self.worst = srt[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best, cur_len):
        return len(self.beams) >= self.k and self.worst >= best / cur_len ** self.lp


@torch.no_grad()
def ensemble_beam_search(decoders, encs, masks, num_beams, max_length, bos, eos, pad, length_penalty=1.0, gaps=None,
                         trace=None, on_step=None):
    """Sum-of-logits ensemble beam search (beam_search.py:243-320); greedy when num_beams == 1.
    gaps: optional list that receives, per step, the smallest score gap between adjacent candidates among the top
    2k+1 (k>1) or top-2 (greedy) — how close the fp32 search came to a tie (tests use it to qualify bit-exactness).
    trace: optional list that receives (step, batch row, gap, max |summed logit| of that row's beams) for the same decisions:
    a bf16 implementation carries a logit error proportional to the logit magnitude, so tests qualify by gap / scale.
    on_step: optional callback(cur_len, ids [B*k, cur_len], scores [B*k], done list) called after every step's update; with
    num_beams > 1 it also receives parents=[B*k] (source row of every new beam) and tokens=[B*k] as keyword arguments."""
    B, k = encs[0].shape[0], num_beams
    ids = torch.full((B * k, 1), bos, dtype=torch.long)
    encs = [e.repeat_interleave(k, 0) for e in encs]
    masks = [m.repeat_interleave(k, 0) if m is not None else None for m in masks]
    scores = torch.zeros(B, k)
    scores[:, 1:] = -1e9
    scores = scores.view(-1)
    hyps = [_Hyps(k, length_penalty) for _ in range(B)]
    done = [False] * B
    cur = 1
    while cur < max_length:
        logits = sum(next_logits(d, ids, e, m) for d, e, m in zip(decoders, encs, masks))        # :254
        lp = torch.log_softmax(logits, -1) + scores[:, None]                                     # :260-265
        V = lp.shape[-1]
        if gaps is not None:
            # distance to a different search decision: greedy = top-1 vs top-2; beam = the k-th kept non-EOS
            # candidate vs the first one left out (a bf16 implementation may legitimately flip closer calls)
            for b in range(B):
                if done[b]:
                    continue
                if k == 1:
                    top = torch.topk(lp[b], 2).values
                    gaps.append(float(top[0] - top[1]))
                    if trace is not None:
                        trace.append((cur, b, float(top[0] - top[1]), float(logits[b].abs().max())))
                else:
                    ts_, ti_ = torch.topk(lp.view(B, k * V)[b], min(2 * k + 2, k * V))
                    keep = [float(v) for v, i in zip(ts_, ti_) if int(i) % V != eos and float(v) > -1e8]
                    if len(keep) > k:
                        gaps.append(keep[k - 1] - keep[k])
                        if trace is not None:
                            trace.append((cur, b, keep[k - 1] - keep[k], float(logits.view(B, k, V)[b].abs().max())))
        if k == 1:
            s, t = lp.max(-1)
            t = torch.where(torch.tensor(done), torch.full_like(t, pad), t)
            ids = torch.cat([ids, t[:, None]], 1)
            scores = s
            cur += 1
            for b, tok in enumerate(t.tolist()):
                done[b] = done[b] or tok == eos
            if on_step is not None:
                on_step(cur, ids.clone(), scores.clone(), list(done))
            if all(done):
                break
            continue
        ts, ti = torch.topk(lp.view(B, k * V), 2 * k, dim=1)                                     # :289-294
        nb, nt, ns = torch.zeros(B, k, dtype=torch.long), torch.zeros(B, k, dtype=torch.long), torch.zeros(B, k)
        for b in range(B):
            if done[b]:
                nt[b].fill_(pad)
                nb[b] = torch.arange(k) + b * k
                continue
            j = 0
            for rank in range(2 * k):
                bi, tk, sc = int(ti[b, rank]) // V, int(ti[b, rank]) % V, float(ts[b, rank])
                if tk == eos:
                    if rank >= k:
                        continue
                    hyps[b].add(ids[b * k + bi].clone(), sc)
                else:
                    nb[b, j], nt[b, j], ns[b, j] = b * k + bi, tk, sc
                    j += 1
                if j == k:
                    break
            done[b] = done[b] or hyps[b].is_done(float(ts[b].max()), cur)
        scores = ns.view(-1)
        ids = torch.cat([ids[nb.view(-1)], nt.view(-1, 1)], 1)
        cur += 1
        if on_step is not None:
            try:
                on_step(cur, ids.clone(), scores.clone(), list(done), parents=nb.view(-1).clone(), tokens=nt.view(-1).clone())
            except TypeError:
                on_step(cur, ids.clone(), scores.clone(), list(done))
        if all(done):
            break
    if k == 1:
        return ids
    out = []
    for b in range(B):
        if not done[b]:
            for j in range(k):
                hyps[b].add(ids[b * k + j], float(scores[b * k + j]))
        out.append(max(hyps[b].beams, key=lambda x: x[0])[1])
    L = min(max(len(o) for o in out) + 1, max_length)
    res = torch.full((B, L), pad, dtype=torch.long)
    for b, o in enumerate(out):
        res[b, :len(o)] = o
        if len(o) < max_length:
            res[b, len(o)] = eos
    return res
