"""Generate the golden fixtures that pin the oracle.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

  losses.pt  <- outputs of the REFERENCE'S OWN loss modules imported by file path from /root/reference
                (ConVIRTLoss / LabelSmoothingCrossEntropy unmodified; InfoNCELoss / GLoRIALoss with Tensor.cuda()
                patched to a no-op because this container has no GPU).
  towers.pt  <- outputs of the installed transformers (ViTModel / BertGenerationDecoder, eager attention) composed as
                the reference composes them, on seeded weights + inputs (small probes; full tensors are not stored).
Inputs are NOT stored: they are regenerated from the recorded seeds (torch CPU generators are deterministic).
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/vilmedic/blocks/losses"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def loss_inputs(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)


def gloria_inputs(b, d, hw, lw, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(b, d, hw, hw, generator=g)
    words = torch.randn(b, d, lw, generator=g)
    lens = torch.randint(3, lw + 1, (b,), generator=g).tolist()
    sents = [["w%d" % j for j in range(n - 1)] + ["[SEP]"] * 2 for n in lens]   # cap_len = #non-special + 1
    return img, words, sents


def make_losses():
    torch.Tensor.cuda = lambda self, *a, **k: self   # reference files hard-code .cuda() (InfoNCELoss.py:14, GLoRIALoss.py:56,125)
    convirt = _load(os.path.join(REF, "selfsup/ConVIRTLoss.py"), "ref_convirt")
    infonce = _load(os.path.join(REF, "selfsup/InfoNCELoss.py"), "ref_infonce")
    gloria = _load(os.path.join(REF, "selfsup/GLoRIALoss.py"), "ref_gloria")
    lsce = _load(os.path.join(REF, "mvqa/LabelSmoothingCrossEntropyLoss.py"), "ref_lsce")
    out = {"cases": []}
    for n, d, seed in [(4, 32, 0), (64, 768, 1), (512, 768, 2)]:
        l, v = loss_inputs(n, d, seed)
        l.requires_grad_(True)
        v.requires_grad_(True)
        loss, ll, lv = convirt.ConVIRTLoss(tau=0.1, lambda_=0.75)(l, v)
        gl, gv = torch.autograd.grad(loss, (l, v))
        rec = {"kind": "convirt", "n": n, "d": d, "seed": seed, "tau": 0.1, "lambda_": 0.75, "loss": loss.detach(),
               "loss_l": ll.detach(), "loss_v": lv.detach(), "grad_l_norm": gl.norm(), "grad_v_norm": gv.norm(),
               "grad_l_probe": gl[:2, :8].clone(), "grad_v_probe": gv[:2, :8].clone()}
        out["cases"].append(rec)
        l2, v2 = loss_inputs(n, d, seed)
        l2 = (l2 * 0.05).requires_grad_(True)
        v2 = (v2 * 0.05).requires_grad_(True)
        loss, lt, li = infonce.InfoNCELoss(tau=0.1)(l2, v2)
        gl, gv = torch.autograd.grad(loss, (l2, v2))
        out["cases"].append({"kind": "infonce", "n": n, "d": d, "seed": seed, "scale": 0.05, "loss": loss.detach(),
                             "loss_t": lt.detach(), "loss_i": li.detach(), "grad_l_norm": gl.norm(), "grad_v_norm": gv.norm(),
                             "grad_l_probe": gl[:2, :8].clone(), "grad_v_probe": gv[:2, :8].clone()})
    for b, d, hw, lw, seed in [(4, 32, 5, 9, 3), (8, 768, 19, 24, 4)]:
        img, words, sents = gloria_inputs(b, d, hw, lw, seed)
        g = torch.Generator().manual_seed(seed + 100)
        gi, gt = torch.randn(b, d, generator=g), torch.randn(b, d, generator=g)
        img.requires_grad_(True)
        words.requires_grad_(True)
        mod = gloria.GLoRIALoss(temp1=4.0, temp2=5.0, temp3=10.0)
        loss, attn = mod(gi, img, words, gt, sents)
        l0, l1, _ = mod._calc_local_loss(img, words, sents)
        g0, g1 = mod._calc_global_loss(gi, gt)
        dimg, dwords = torch.autograd.grad(loss, (img, words))
        out["cases"].append({"kind": "gloria", "b": b, "d": d, "hw": hw, "lw": lw, "seed": seed, "loss": loss.detach(),
                             "local0": l0.detach(), "local1": l1.detach(), "global0": g0.detach(), "global1": g1.detach(),
                             "attn0_probe": attn[0][0, :2].detach().clone(), "dimg_norm": dimg.norm(), "dwords_norm": dwords.norm()})
    for n, c, seed in [(16, 10, 5), (256, 330, 6)]:
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(n, c, generator=g) * 2
        t = torch.randint(0, c, (n,), generator=g)
        x.requires_grad_(True)
        loss = lsce.LabelSmoothingCrossEntropy(smoothing=0.1)(x, t)
        (gx,) = torch.autograd.grad(loss, (x,))
        out["cases"].append({"kind": "lsce", "n": n, "c": c, "seed": seed, "smoothing": 0.1, "loss": loss.detach(),
                             "grad_norm": gx.norm(), "grad_probe": gx[:2, :8].clone()})
    torch.save(out, os.path.join(HERE, "losses.pt"))
    print("losses.pt:", len(out["cases"]), "cases")


def make_towers():
    from vilmedic_b200 import synth
    from oracle.rrg import OracleRRG
    import transformers
    out = {"transformers": transformers.__version__, "torch": str(torch.__version__), "cases": []}
    for name, vit_layers, dec_layers, vocab, B, T in [("small", 2, 2, 1000, 2, 16), ("vitb_dec12", 12, 12, 30522, 2, 32)]:
        torch.manual_seed(0)
        dec = synth.bert_base_decoder(vocab=vocab, layers=dec_layers, dropout=0.0)
        cnn = dict(backbone="vit", permute="no_permute", **dict(synth.vit_b16(), num_hidden_layers=vit_layers))
        m = OracleRRG(dec, cnn).eval()
        batch = synth.rrg_batch(B, T, vocab)
        feats, fmask = m.enc.encode(batch["images"])
        o = m(batch["input_ids"], batch["attention_mask"], batch["images"])
        o["loss"].backward()
        gn = {n: p.grad.norm().item() for n, p in m.named_parameters()}
        keys = ["enc.model.embeddings.patch_embeddings.projection.weight", "enc.model.encoder.layer.0.attention.attention.query.weight",
                "enc.model.layernorm.weight", "dec.decoder.bert.embeddings.word_embeddings.weight",
                "dec.decoder.bert.encoder.layer.0.crossattention.self.key.weight", "dec.decoder.bert.encoder.layer.%d.output.dense.weight" % (dec_layers - 1),
                "dec.decoder.lm_head.bias"]
        out["cases"].append({"name": name, "vit_layers": vit_layers, "dec_layers": dec_layers, "vocab": vocab, "B": B, "T": T,
                             "loss": o["loss"].detach(), "feats_probe": feats[:, :3, :8].detach().clone(), "feats_norm": feats.norm().detach(),
                             "logits_probe": o["logits"][:, :3, :8].detach().clone(), "logits_norm": o["logits"].norm().detach(),
                             "grad_norms": {k: gn[k] for k in keys}})
        print(name, "loss", o["loss"].item())
    torch.save(out, os.path.join(HERE, "towers.pt"))


if __name__ == "__main__":
    make_losses()
    make_towers()
