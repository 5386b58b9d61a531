"""SCST rollouts and loss on the decode-step kernels (vilmedic/blocks/rl/SCST.py:12-190).

 * device sampling: the drawn tokens follow softmax(filtered logits) (frequency test over 4096 identical rows), never a bad word,
   never outside the top-k, reproducible for a seed, different across seeds; `forced_eos_token_id` lands on the last position;
 * sequence_log_probs (teacher-forced, processors applied, reward weights folded into the CE kernel): values and parameter gradients
   against the fp32 HF decoder with the same processors written in torch;
 * SCST.forward_greedy / forward_sampling run end to end with a callable scorer.
"""
import copy
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
BOS, PAD, EOS = 0, 1, 2


def _pair(seed=0, vocab=300, layers=2):
    from oracle.rrg import OracleRRG
    from vilmedic_b200 import synth
    from vilmedic_b200.models import RRG
    torch.manual_seed(seed)
    dec = synth.bert_base_decoder(vocab=vocab, layers=layers, dropout=0.0)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **dict(synth.vit_b16(), num_hidden_layers=1))
    ref = OracleRRG(dec, cnn).eval()
    with torch.no_grad():
        ref.dec.decoder.bert.embeddings.word_embeddings.weight.mul_(10.0)
        for p in ref.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    mine = RRG(copy.deepcopy(dec), copy.deepcopy(cnn))
    mine.load_state_dict(ref.state_dict())
    return ref, mine.cuda()


def _enc(ref, n, seed=3):
    from vilmedic_b200 import synth
    with torch.no_grad():
        e, m = ref.enc.encode(synth.rrg_batch(n, 8, 300, seed=seed)["images"])
    return e.to(torch.bfloat16).float(), m


def test_device_sampling_distribution_filters_and_seeds(cuda_dev):
    ref, mine = _pair()
    mine.eval()
    dec = mine.dec.decoder
    e1, m1 = _enc(ref, 1)
    N, K = 4096, 6
    enc = e1.expand(N, -1, -1).contiguous().cuda()
    mask = m1.expand(N, -1).contiguous().cuda()
    ids = torch.full((N, 1), BOS, dtype=torch.long, device="cuda")
    kw = dict(input_ids=ids, encoder_hidden_states=enc, encoder_attention_mask=mask, max_length=4, num_beams=1, do_sample=True, top_k=K,
              bad_words_ids=[[PAD], [BOS]], bos_token_id=BOS, eos_token_id=EOS, pad_token_id=PAD)
    a = dec.generate(seed=11, **kw)
    b = dec.generate(seed=11, **kw)
    c = dec.generate(seed=12, **kw)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # expected distribution of the first drawn token: softmax over the top-K of the (bad-word-free) logits
    lg = dec.next_token_logits(ids[:1], enc[:1], mask[:1]).float()[0]
    lg[PAD] = lg[BOS] = float("-inf")
    top = torch.topk(lg, K)
    probs = torch.softmax(top.values, -1).cpu()
    first = a[:, 1].cpu()
    allowed = set(top.indices.cpu().tolist())
    assert set(first.tolist()) <= allowed, "a token outside the top-k (or a bad word) was drawn"
    for tok, p in zip(top.indices.cpu().tolist(), probs.tolist()):
        f = (first == tok).float().mean().item()
        assert abs(f - p) <= 5.0 * math.sqrt(p * (1 - p) / N) + 2e-3, (tok, f, p)
    # forced token on the last position (the reference passes forced_eos_token_id=True -> token id 1)
    d = dec.generate(seed=5, forced_eos_token_id=True, **kw)
    live = d[:, 2] != EOS                                    # rows that had not finished before the last position
    assert bool((d[live][:, 3] == 1).all())
    g = dec.generate(input_ids=ids[:8], encoder_hidden_states=enc[:8], encoder_attention_mask=mask[:8], max_length=5, num_beams=1,
                     forced_eos_token_id=True, return_dict_in_generate=True, output_scores=True, bos_token_id=BOS, eos_token_id=EOS,
                     pad_token_id=PAD)
    assert g.sequences.shape == (8, 5)


def test_sequence_log_probs_and_gradients_vs_fp32_oracle(cuda_dev):
    from vilmedic_b200.blocks.rl import sequence_log_probs
    ref, mine = _pair(1)
    mine.train()                                                          # dropout 0 in these configs: train == eval numerically
    B, L, K = 5, 9, 20
    enc, mask = _enc(ref, B, seed=8)
    g = torch.Generator().manual_seed(4)
    seq = torch.randint(3, 300, (B, L), generator=g)
    seq[:, 0] = BOS
    w = torch.randn(B, L - 1, generator=g)
    # ---- oracle: fp32 HF decoder, processors written in torch (NoBadWords -> TopK -> log_softmax -> gather), SCST.py:139-158
    hf = ref.dec.decoder
    logits = hf(input_ids=seq[:, :-1], encoder_hidden_states=enc, encoder_attention_mask=mask, use_cache=False).logits.float()
    lg = logits.clone()
    lg[..., PAD] = float("-inf")
    lg[..., BOS] = float("-inf")
    kth = torch.topk(lg, K, dim=-1).values[..., -1:]
    # targets from the oracle's own top-(K-5): well inside the kept set, so a bf16 flip at the top-k boundary cannot remove them
    safe = torch.topk(lg, K - 5, dim=-1).indices
    tgt = safe.gather(-1, torch.randint(0, K - 5, (B, L - 1, 1), generator=g)).squeeze(-1)
    lg = lg.masked_fill(lg < kth, float("-inf"))
    lp_ref = torch.log_softmax(lg, -1).gather(-1, tgt[..., None]).squeeze(-1)
    (lp_ref * w).sum().backward()
    # ---- product: teacher-forced pass + vlm_logits_filter + weighted CE kernel
    lp = sequence_log_probs(mine.dec.decoder, seq, enc.cuda(), mask.cuda(), bad_ids=[PAD, BOS], top_k=K, targets=tgt)
    assert lp.shape == (B, L - 1) and torch.isfinite(lp).all()
    # log-probabilities down to -120 here (x10 embeddings): bf16 activations give ~1e-3 relative on the logits behind them
    assert (lp.float().cpu() - lp_ref.detach()).abs().max().item() <= 3e-2 + 2e-3 * lp_ref.abs().max().item()
    (lp * w.cuda()).sum().backward()
    torch.cuda.synchronize()

    def rel(a, b):
        return ((a.float().cpu() - b).norm() / (b.norm() + 1e-12)).item()

    d = mine.dec.decoder
    assert rel(d.lm_head.bias.grad, hf.lm_head.bias.grad) <= 6e-2
    assert d.lm_head.bias.grad[PAD].item() == 0.0 and d.lm_head.bias.grad[BOS].item() == 0.0        # filtered: exactly zero gradient
    l0, r0 = d.bert.encoder.layer[0], hf.bert.encoder.layer[0]
    assert rel(l0.intermediate.dense.weight.grad, r0.intermediate.dense.weight.grad) <= 8e-2
    assert rel(l0.crossattention.self.query.weight.grad, r0.crossattention.self.query.weight.grad) <= 8e-2
    assert rel(d.bert.embeddings.word_embeddings.weight.grad, hf.bert.embeddings.word_embeddings.weight.grad) <= 8e-2


def test_scst_forward_greedy_and_sampling_end_to_end(cuda_dev):
    from types import SimpleNamespace
    from vilmedic_b200.blocks.rl import SCST
    from vilmedic_b200.executors.utils import _Tokenizer
    ref, mine = _pair(2)
    mine.train()
    B, L = 6, 12
    enc, mask = _enc(ref, B, seed=6)
    tok = _Tokenizer(300)
    dl = SimpleNamespace(dataset=SimpleNamespace(tokenizer=tok, tokenizer_max_len=L))

    def scorer(refs, hyps):           # a REWARD_COMPLIANT-style scorer: per-sample rewards, a deterministic function of the hypothesis
        return [(sum(int(w[1:]) for w in h.split()) % 11) / 11.0 + 0.01 * len(h.split()) for h in hyps]

    scst = SCST(mine.dec.decoder, dl, scores=[scorer], top_k=10)
    g = torch.Generator().manual_seed(0)
    input_ids = torch.randint(5, 300, (B, L), generator=g)
    input_ids[:, 0] = BOS
    attention_mask = torch.ones(B, L, dtype=torch.long)
    with torch.no_grad():
        reward_greedy, hyps, refs = scst.forward_greedy(input_ids, enc.cuda(), mask.cuda())
    assert len(reward_greedy) == 1 and len(reward_greedy[0]) == B and len(hyps) == B
    enc_g = enc.cuda().requires_grad_(True)
    loss, delta, delta_per_metric, reward_sampling, hyps_s = scst.forward_sampling(input_ids, attention_mask, enc_g, mask.cuda(), reward_greedy)
    assert loss.dim() == 0 and torch.isfinite(loss)
    loss.backward()
    torch.cuda.synchronize()
    gb = mine.dec.decoder.lm_head.bias.grad
    assert gb is not None and torch.isfinite(gb).all() and gb.abs().sum().item() > 0
    assert gb[PAD].item() == 0.0 and gb[BOS].item() == 0.0          # filtered tokens get exactly zero gradient
    assert enc_g.grad is not None and torch.isfinite(enc_g.grad).all()
