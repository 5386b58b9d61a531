"""Secondary metric (BASELINE configs[4]): tokens/s of the Sigma-logits ensemble beam search —
two independently seeded RRG(ViT-B/16 -> 12-layer decoder) models, B=32 images, beam 4, max_len 128, V=30522, KV-cached
decode steps on the kernels.  EOS is made unreachable (bias -1e4) so that every hypothesis runs to max_len: the count
B * (max_len - 1) generated tokens is then exact.  Usage: python tools/decode_bench.py [--batch 32] [--beams 4] [--models 2]
[--max-len 128]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    from vilmedic_b200 import synth
    from vilmedic_b200.models import RRG
    B, k, M, L, V = arg("--batch", 32), arg("--beams", 4), arg("--models", 2), arg("--max-len", 128), 30522
    BOS, PAD, EOS = 0, 1, 2
    models = []
    for s in range(M):
        torch.manual_seed(s)
        dec = synth.bert_base_decoder(vocab=V, layers=12, dropout=0.0)
        cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **synth.vit_b16())
        m = RRG(dec, cnn).cuda().eval()
        with torch.no_grad():
            m.dec.decoder.lm_head.bias[EOS] = -1e4
        models.append(m)
    batch = synth.rrg_batch(B, 8, V)
    images = batch["images"].cuda()
    times = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            encs, masks = zip(*[m.encode(images) for m in models])
            out = models[0].dec.decoder.generate(input_ids=torch.full((B, 1), BOS, dtype=torch.long, device="cuda"),
                                                 encoder_hidden_states=list(encs), encoder_attention_mask=list(masks),
                                                 ensemble=[m.dec.decoder for m in models], max_length=L, num_beams=k,
                                                 bos_token_id=BOS, eos_token_id=EOS, pad_token_id=PAD)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    assert out.shape[0] == B and out.shape[1] == L, tuple(out.shape)
    best = min(times[1:])
    print(json.dumps({"metric": "beam-search tokens/s (ensemble of %d RRG ViT-B/16 -> 12-layer decoder, beam %d)" % (M, k),
                      "value": B * (L - 1) / best, "unit": "tokens/s", "batch": B, "max_len": L, "seconds": best,
                      "ms_per_step": 1e3 * best / (L - 1), "includes": "image encoding by every model + the whole search (device-side step replayed as one CUDA graph per token)"}))


if __name__ == "__main__":
    main()
