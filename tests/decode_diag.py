"""Diagnostic for tests/test_decode_gpu.py (lives under tests/ because it drives the oracle, which only test code may import): prints the measured bf16-vs-fp32 logit error, the cached-vs-uncached logit
difference and the decision margins (relative to the logit scale) of the oracle's searches.  GPU only; not a test."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import test_decode_gpu as T  # noqa: E402
from oracle import decode  # noqa: E402
from vilmedic_b200 import synth  # noqa: E402
from vilmedic_b200.blocks.huggingface.decoder.generation import DecodeState  # noqa: E402

BOS, PAD, EOS = 0, 1, 2


def cached_vs_full(mine, images, ids):
    enc, mask = mine.encode(images)
    dec = mine.dec.decoder
    st = DecodeState(dec, ids.shape[0], 16, enc, mask)
    worst = 0.0
    for t in range(ids.shape[1]):
        step = dec.decode_step(st, ids[:, t])
        full = dec.next_token_logits(ids[:, :t + 1], enc, mask)
        worst = max(worst, ((step - full).abs().max() / full.abs().max()).item())
    return worst


def main():
    ref, mine = T._pair(0)
    batch = synth.rrg_batch(3, 8, 300, seed=9)
    rel = T._rel_logit_error([(ref, mine)], batch["images"])
    print("greedy: rel logit err (bf16 vs fp32) = %.5f" % rel)
    enc_r, mask_r = ref.enc.encode(batch["images"])
    enc, mask = mine.encode(batch["images"])
    print("  feature err rel = %.5f" % ((enc.float().cpu() - enc_r).abs().max() / enc_r.abs().max()).item())
    trace = []
    want = decode.ensemble_beam_search([ref.dec.decoder], [enc_r], [mask_r], 1, 12, BOS, EOS, PAD, gaps=[], trace=trace)
    got = mine.dec.decoder.generate(input_ids=torch.full((3, 1), BOS, dtype=torch.long, device="cuda"), encoder_hidden_states=enc,
                                    encoder_attention_mask=mask, max_length=12, num_beams=1, bos_token_id=BOS, eos_token_id=EOS,
                                    pad_token_id=PAD).cpu()
    print("  want", want.tolist())
    print("  got ", got.tolist())
    for row in range(3):
        print("  row %d gap/scale per step:" % row, " ".join("%.4f" % (g / s) for (st, b, g, s) in trace if b == row))
    print("cached vs full (rel to scale): %.6f" % cached_vs_full(mine, batch["images"], batch["input_ids"].cuda()[:, :8]))
    pairs = [(ref, mine), T._pair(1)]
    for n_models in (1, 2):
        ps = pairs[:n_models]
        for seed in range(35, 47):
            b = synth.rrg_batch(2, 8, 300, seed=seed)
            encs, masks = zip(*[m.encode(b["images"]) for _, m in ps])
            adapters = [T._MineAsOracleModel(m.dec.decoder, e, mk) for (_, m), e, mk in zip(ps, encs, masks)]
            tr = []
            decode.ensemble_beam_search(adapters, [e.cpu() for e in encs], [mk.cpu() for mk in masks], 4, 8, BOS, EOS, PAD,
                                        gaps=[], trace=tr)
            r = T._rel_logit_error(ps, b["images"])
            print("beam M=%d seed %d: rel err %.5f, min gap/scale %.5f, all: %s" % (
                n_models, seed, r, min(g / s for (_, _, g, s) in tr), " ".join("%.4f" % (g / s) for (_, _, g, s) in tr)))


if __name__ == "__main__":
    main()
