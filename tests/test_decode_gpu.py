"""Greedy / beam / ensemble decoding parity: token ids of the B200 path must equal the oracle's.

The oracle computes in fp32, the product in bf16, so a search decision closer than the bf16 logit error could
legitimately flip.  Following SURVEY.md §7 ("hard parts"): (i) weights are scaled so that decisions have non-trivial
margins and are rounded to bf16 on BOTH sides (identical parameters); (ii) the oracle reports the smallest decision margin
of its own search (top-1 vs top-2 for greedy; k-th kept vs first dropped candidate for beam) together with the logit
magnitude, and the comparison is bit-exact over every decision whose margin exceeds 2^-5 x max|logit| (the bf16 logit
error scales with the logit magnitude) — a mismatch on such a decision is a real failure.
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


# A bf16 forward carries a logit error that scales with the logit magnitude.  It is MEASURED here on the first step (device
# logits vs the fp32 oracle's, same bf16-rounded weights); a search decision is "safe" when the oracle's margin exceeds
# SAFETY x that error relative to the logit scale (the error grows slowly with the prefix length, hence the cushion).
SAFETY = 6.0


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


def _first_unsafe_step(trace, row, max_length, rel):
    steps = [st for (st, b, gap, scale) in trace if b == row and gap <= rel * scale]
    return min(steps) if steps else max_length


def test_greedy_token_ids_bit_exact_vs_fp32_oracle(cuda_dev):
    """Greedy ids == fp32 oracle (== HF generate) on every row up to the first step whose oracle margin is within the bf16
    logit error; the seeds keep at least 6 safe steps per row."""
    from oracle import decode
    from vilmedic_b200 import synth
    ref, mine = _pair(0)
    batch = synth.rrg_batch(3, 8, 300, seed=9)
    enc_r, mask_r = ref.enc.encode(batch["images"])
    gaps, trace = [], []
    want = decode.ensemble_beam_search([ref.dec.decoder], [enc_r], [mask_r], 1, 12, BOS, EOS, PAD, gaps=gaps, trace=trace)
    hf = decode.hf_generate(ref.dec.decoder, enc_r, mask_r, 1, 12, BOS, EOS, PAD)
    assert torch.equal(want[:, :hf.shape[1]], hf[:, :want.shape[1]]), "oracle restatement disagrees with HF generate"
    enc, mask = mine.encode(batch["images"])
    got = mine.dec.decoder.generate(input_ids=torch.full((3, 1), BOS, dtype=torch.long, device="cuda"), encoder_hidden_states=enc,
                                    encoder_attention_mask=mask, max_length=12, num_beams=1, bos_token_id=BOS, eos_token_id=EOS,
                                    pad_token_id=PAD).cpu()
    assert got.shape == want.shape
    rel = SAFETY * _rel_logit_error([(ref, mine)], batch["images"])
    assert rel < 2.0 ** -5, "bf16 logit error unexpectedly large (%.4f of the logit scale)" % (rel / SAFETY)
    compared = 0
    for row in range(3):
        safe = _first_unsafe_step(trace, row, 12, rel)     # tokens at positions < safe come from safe decisions
        assert safe >= 6, "test inputs lost their argmax margin on row %d (first unsafe step %d); pick another seed" % (row, safe)
        assert torch.equal(got[row, :safe], want[row, :safe]), (row, safe, got[row].tolist(), want[row].tolist())
        compared += safe
    assert compared >= 24


@pytest.mark.parametrize("k,n_models", [(4, 1), (4, 2)])
def test_beam_ensemble_token_ids(cuda_dev, k, n_models):
    """KV-cached beam / ensemble search == the oracle's search loop run over the product's own (uncached) logits, bit for
    bit (on the first seed whose decisions are all safe w.r.t. cached-vs-uncached bf16 noise); and == the fp32 oracle end to
    end whenever the oracle's own decisions are all safe."""
    from oracle import decode
    from vilmedic_b200 import synth
    pairs = [_pair(s) for s in range(n_models)]
    hf_models = [m.dec.decoder for _, m in pairs]
    checked = False
    for seed in (35, 36, 37, 38, 39, 40):
        batch = synth.rrg_batch(2, 8, 300, seed=seed)
        encs, masks = zip(*[m.encode(batch["images"]) for _, m in pairs])
        adapters = [_MineAsOracleModel(m.dec.decoder, e, mk) for (_, m), e, mk in zip(pairs, encs, masks)]
        trace_mine = []
        want_mine = decode.ensemble_beam_search(adapters, [e.cpu() for e in encs], [mk.cpu() for mk in masks], k, 8, BOS, EOS, PAD,
                                                gaps=[], trace=trace_mine)
        rel = SAFETY * _rel_logit_error(pairs, batch["images"])
        if any(gap <= rel * scale for (_, _, gap, scale) in trace_mine):
            continue                                        # a near-tie of the device logits: cached vs uncached may flip it
        got = hf_models[0].generate(input_ids=torch.full((2, 1), BOS, dtype=torch.long, device="cuda"),
                                    encoder_hidden_states=list(encs), encoder_attention_mask=list(masks), ensemble=hf_models,
                                    max_length=8, num_beams=k, bos_token_id=BOS, eos_token_id=EOS, pad_token_id=PAD).cpu()
        assert got.shape == want_mine.shape and torch.equal(got, want_mine), (seed, got.tolist(), want_mine.tolist())
        encs_r, masks_r = zip(*[r.enc.encode(batch["images"]) for r, _ in pairs])
        trace = []
        want = decode.ensemble_beam_search([r.dec.decoder for r, _ in pairs], list(encs_r), list(masks_r), k, 8, BOS, EOS, PAD,
                                           gaps=[], trace=trace)
        if all(gap > 2 * rel * scale for (_, _, gap, scale) in trace):
            assert torch.equal(got, want), (seed, got.tolist(), want.tolist())
        else:
            n = min(got.shape[1], want.shape[1])
            agree = (got[:, :n] == want[:, :n]).float().mean().item()
            print("fp32-oracle beam search has a near-tie on seed %d: token agreement %.2f (informational)" % (seed, agree))
        checked = True
        break
    assert checked, "no seed in 35..40 gave a batch whose beam decisions are all clear of bf16 noise"


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
