"""Greedy / beam / ensemble decoding parity: token ids of the B200 path must equal the oracle's.

The oracle computes in fp32, the product in bf16, so a search decision closer than the bf16 logit error could
legitimately flip.  Following SURVEY.md §7 ("hard parts"): (i) weights are scaled so that decisions have non-trivial
margins and are rounded to bf16 on BOTH sides (identical parameters); (ii) the oracle reports the smallest decision margin
of its own search (top-1 vs top-2 for greedy; k-th kept vs first dropped candidate for beam) together with the logit
magnitude, and the comparison is bit-exact over every decision whose margin exceeds twice the MEASURED device-vs-oracle
logit error of that very step (itself bounded by a stated tolerance) — a mismatch on such a decision is a real failure.
The selection logic itself is verified bit-exactly on identical logits in tests/test_cpu.py."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
BOS, PAD, EOS = 0, 1, 2


def _pair(seed, vocab=300):
    from oracle.rrg import OracleRRG
    from vilmedic_b200 import synth
    from vilmedic_b200.models import RRG
    torch.manual_seed(seed)
    dec = synth.bert_base_decoder(vocab=vocab, layers=2, dropout=0.0)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **dict(synth.vit_b16(), num_hidden_layers=1))
    ref = OracleRRG(dec, cnn).eval()
    with torch.no_grad():
        for p in ref.dec.parameters():
            if p.dim() > 1:
                p.mul_(3.0)
        ref.dec.decoder.bert.embeddings.word_embeddings.weight.mul_(30.0)
        for p in ref.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    mine = RRG(copy.deepcopy(dec), copy.deepcopy(cnn))
    mine.load_state_dict(ref.state_dict())
    return ref, mine.cuda().eval()


class _MineAsOracleModel:
    """Adapter: lets the ORACLE's search loop drive the product's (uncached, full-prefix) next-token logits."""

    def __init__(self, dec, enc, mask):
        self.dec, self.enc, self.mask = dec, enc, mask
        self.config = dec.config

    def __call__(self, input_ids, encoder_hidden_states=None, encoder_attention_mask=None, use_cache=False):
        B = input_ids.shape[0]
        rep = B // self.enc.shape[0]
        enc = self.enc.repeat_interleave(rep, 0)
        mask = self.mask.repeat_interleave(rep, 0) if self.mask is not None else None
        lg = self.dec.next_token_logits(input_ids.cuda(), enc, mask).float().cpu()

        class _O:
            pass
        o = _O()
        o.logits = lg[:, None, :]
        return o


# A bf16 forward carries a logit error that scales with the logit magnitude.  It is MEASURED here (device logits vs the fp32
# oracle's, same bf16-rounded weights, same prefix): if every device logit is within e of the oracle's, a decision whose oracle
# margin exceeds 2e (+ a small cushion for cached-vs-uncached reduction order) cannot flip — such decisions are "safe" and
# must match bit for bit; the tolerance on e itself is LOGIT_TOL (measured on B200: 1.2-4.2 % of the logit scale over 66
# teacher-forced steps with these x3 / x30 scaled weights, which amplify the bf16 rounding of the activations; tests/decode_diag.py).
LOGIT_TOL = 0.08          # max |device logit - fp32 logit| / max |fp32 logit|
CUSHION = 2.0 ** -10      # x logit scale


def test_greedy_token_ids_bit_exact_vs_fp32_oracle(cuda_dev):
    """(1) the oracle restatement == HF generate; (2) teacher-forced on the oracle's prefixes, the device logits stay within
    LOGIT_TOL of the fp32 oracle's and every safe decision has the same argmax; (3) the free-running KV-cached greedy decode
    reproduces the oracle's token ids on every row up to that row's first unsafe decision."""
    from oracle import decode
    from vilmedic_b200 import synth
    ref, mine = _pair(0)
    L = 12
    NB = 6
    batch = synth.rrg_batch(NB, 8, 300, seed=9)
    enc_r, mask_r = ref.enc.encode(batch["images"])
    want = decode.ensemble_beam_search([ref.dec.decoder], [enc_r], [mask_r], 1, L, BOS, EOS, PAD)
    hf = decode.hf_generate(ref.dec.decoder, enc_r, mask_r, 1, L, BOS, EOS, PAD)
    assert torch.equal(want[:, :hf.shape[1]], hf[:, :want.shape[1]]), "oracle restatement disagrees with HF generate"
    enc, mask = mine.encode(batch["images"])
    first_unsafe = [want.shape[1]] * NB
    safe_decisions = 0
    for t in range(1, want.shape[1]):
        lr = decode.next_logits(ref.dec.decoder, want[:, :t], enc_r, mask_r)
        lm = mine.dec.decoder.next_token_logits(want[:, :t].cuda(), enc, mask).float().cpu()
        for row in range(NB):
            if want[row, t] == PAD and (want[row, :t] == EOS).any():
                continue                                                     # row already finished
            scale = lr[row].abs().max().item()
            err = (lm[row] - lr[row]).abs().max().item()
            assert err <= LOGIT_TOL * scale, "bf16 logit error %.4f of the logit scale at step %d row %d" % (err / scale, t, row)
            top = torch.topk(lr[row], 2).values
            if float(top[0] - top[1]) > 2 * err + CUSHION * scale:
                assert int(lm[row].argmax()) == int(want[row, t]), (t, row)
                safe_decisions += 1
            else:
                first_unsafe[row] = min(first_unsafe[row], t)
    assert safe_decisions >= 5 * NB, "test inputs lost their argmax margins (%d safe decisions); pick another seed" % safe_decisions
    got = mine.dec.decoder.generate(input_ids=torch.full((NB, 1), BOS, dtype=torch.long, device="cuda"), encoder_hidden_states=enc,
                                    encoder_attention_mask=mask, max_length=L, num_beams=1, bos_token_id=BOS, eos_token_id=EOS,
                                    pad_token_id=PAD).cpu()
    assert got.shape == want.shape
    compared = 0
    for row in range(NB):
        n = first_unsafe[row]                                                # tokens at positions < n come from safe decisions
        assert torch.equal(got[row, :n], want[row, :n]), (row, n, got[row].tolist(), want[row].tolist())
        compared += n
    assert compared >= 2 * NB, compared


def _rel_cache_noise(mine, images, k):
    """max |KV-cached step logits - full-prefix logits| / max |logit| over a few steps at the beam-expanded batch size."""
    from vilmedic_b200.blocks.huggingface.decoder.generation import DecodeState
    enc, mask = mine.encode(images)
    enc, mask = enc.repeat_interleave(k, 0), (mask.repeat_interleave(k, 0) if mask is not None else None)
    dec = mine.dec.decoder
    ids = torch.randint(5, 300, (enc.shape[0], 6), device="cuda")
    ids[:, 0] = BOS
    st = DecodeState(dec, enc.shape[0], 8, enc, mask)
    worst = 0.0
    for t in range(ids.shape[1]):
        step = dec.decode_step(st, ids[:, t])
        full = dec.next_token_logits(ids[:, :t + 1], enc, mask)
        worst = max(worst, ((step - full).abs().max() / full.abs().max()).item())
    return worst


@pytest.mark.parametrize("k,n_models", [(4, 1), (4, 2)])
def test_beam_ensemble_token_ids(cuda_dev, k, n_models):
    """KV-cached beam / ensemble search == the oracle's search loop (beam_search.py:222-342 restated) run over the product's
    own uncached logits, bit for bit, on the first seed whose decisions all clear the measured cached-vs-uncached noise; the
    fp32 oracle's end-to-end result is compared when its own decisions are all safe w.r.t. the bf16 logit error (beam
    pruning margins among 2k of k*V candidates are usually tighter than that: informational otherwise)."""
    from oracle import decode
    from vilmedic_b200 import synth
    pairs = [_pair(s) for s in range(n_models)]
    hf_models = [m.dec.decoder for _, m in pairs]
    checked = False
    for seed in range(35, 51):
        batch = synth.rrg_batch(2, 8, 300, seed=seed)
        encs, masks = zip(*[m.encode(batch["images"]) for _, m in pairs])
        adapters = [_MineAsOracleModel(m.dec.decoder, e, mk) for (_, m), e, mk in zip(pairs, encs, masks)]
        trace_mine = []
        want_mine = decode.ensemble_beam_search(adapters, [e.cpu() for e in encs], [mk.cpu() for mk in masks], k, 8, BOS, EOS, PAD,
                                                gaps=[], trace=trace_mine)
        noise = 2 * n_models * max(_rel_cache_noise(m, batch["images"], k) for _, m in pairs) + CUSHION / 2
        if any(gap <= noise * scale for (_, _, gap, scale) in trace_mine):
            continue                                        # a near-tie of the device logits: cached vs uncached may flip it
        got = hf_models[0].generate(input_ids=torch.full((2, 1), BOS, dtype=torch.long, device="cuda"),
                                    encoder_hidden_states=list(encs), encoder_attention_mask=list(masks), ensemble=hf_models,
                                    max_length=8, num_beams=k, bos_token_id=BOS, eos_token_id=EOS, pad_token_id=PAD).cpu()
        assert got.shape == want_mine.shape and torch.equal(got, want_mine), (seed, got.tolist(), want_mine.tolist())
        encs_r, masks_r = zip(*[r.enc.encode(batch["images"]) for r, _ in pairs])
        trace = []
        want = decode.ensemble_beam_search([r.dec.decoder for r, _ in pairs], list(encs_r), list(masks_r), k, 8, BOS, EOS, PAD,
                                           gaps=[], trace=trace)
        rel = 2 * _rel_logit_error(pairs, batch["images"])
        if all(gap > 2 * rel * scale for (_, _, gap, scale) in trace):
            assert torch.equal(got, want), (seed, got.tolist(), want.tolist())
        else:
            n = min(got.shape[1], want.shape[1])
            agree = (got[:, :n] == want[:, :n]).float().mean().item()
            print("fp32-oracle beam search has a near-tie on seed %d: token agreement %.2f (informational)" % (seed, agree))
        checked = True
        break
    assert checked, "no seed in 35..50 gave a batch whose beam decisions are all clear of the cached-vs-uncached noise"


def _rel_logit_error(pairs, images):
    """max |device logits - oracle logits| / max |oracle logits| of the (summed) first-step logits."""
    from oracle.decode import next_logits
    b = images.shape[0]
    ids = torch.full((b, 1), BOS, dtype=torch.long)
    ref_sum, mine_sum = 0.0, 0.0
    for ref, mine in pairs:
        enc_r, mask_r = ref.enc.encode(images)
        enc, mask = mine.encode(images)
        ref_sum = ref_sum + next_logits(ref.dec.decoder, ids, enc_r, mask_r)
        mine_sum = mine_sum + mine.dec.decoder.next_token_logits(ids.cuda(), enc, mask).float().cpu()
    return ((mine_sum - ref_sum).abs().max() / ref_sum.abs().max()).item()


def test_cached_step_matches_prefix_recompute(cuda_dev):
    """KV-cached single-token steps produce the same next-token logits as re-running the whole prefix, and as the oracle."""
    from oracle.decode import next_logits
    from vilmedic_b200 import synth
    from vilmedic_b200.blocks.huggingface.decoder.generation import DecodeState
    ref, mine = _pair(3)
    batch = synth.rrg_batch(3, 10, 300, seed=2)
    enc, mask = mine.encode(batch["images"])
    enc_r, mask_r = ref.enc.encode(batch["images"])
    dec = mine.dec.decoder
    ids = batch["input_ids"].cuda()[:, :9]
    st = DecodeState(dec, 3, 16, enc, mask)
    for t in range(ids.shape[1]):
        step = dec.decode_step(st, ids[:, t])
        full = dec.next_token_logits(ids[:, :t + 1], enc, mask)
        want = next_logits(ref.dec.decoder, ids[:, :t + 1].cpu(), enc_r, mask_r)
        tol = 2 ** -4 * want.abs().max().item() + 5e-2     # logits of these x30-scaled embeddings reach |150|
        assert (step - full).abs().max().item() <= tol, t
        assert (step.cpu() - want).abs().max().item() <= tol, t
    a = dec.generate(input_ids=ids[:, :1], encoder_hidden_states=enc, encoder_attention_mask=mask, max_length=12, num_beams=3,
                     bos_token_id=BOS, eos_token_id=EOS, pad_token_id=PAD, use_cache=True)
    b = dec.generate(input_ids=ids[:, :1], encoder_hidden_states=enc, encoder_attention_mask=mask, max_length=12, num_beams=3,
                     bos_token_id=BOS, eos_token_id=EOS, pad_token_id=PAD, use_cache=False)
    assert a.shape == b.shape
