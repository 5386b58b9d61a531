"""Parity of the other hot-path model compositions (RRG_HF multi-image, MVQA, ConVIRT) against their CPU fp32 oracles.
Tolerances as in tests/test_rrg_gpu.py (bf16 compute vs fp32 oracle): loss 2e-2 rel, gradients 1e-1 rel L2."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float().cpu() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _check_grads(mine, ref, tol=1e-1, skip=()):
    refg = {n: p.grad for n, p in ref.named_parameters()}
    gmax = max(g.norm().item() for g in refg.values() if g is not None)
    for n, p in mine.named_parameters():
        if any(s in n for s in skip) or refg.get(n) is None:
            continue
        assert p.grad is not None, n
        g = refg[n]
        # gradients that are (near) zero relative to the model's gradient scale carry no signal, only rounding noise
        small = (p.grad.cpu().float() - g).norm().item() <= 1e-5 * g.numel() ** 0.5 or g.norm().item() <= 2e-3 * gmax
        assert small or _rel(p.grad, g) <= tol, "grad %s rel err %.4f" % (n, _rel(p.grad, g))


def test_rrg_hf_multi_image(cuda_dev):
    from oracle.models import OracleRRGHF
    from vilmedic_b200 import synth
    from vilmedic_b200.models import RRG_HF
    torch.manual_seed(0)
    v_args = dict(synth.vit_b16(), num_hidden_layers=2)
    d_args = dict(synth.bert_base_decoder(vocab=800, layers=2, dropout=0.0))
    ref = OracleRRGHF(v_args, d_args).eval()
    mine = RRG_HF(vision=dict(proto_model="vit", proto_config="vit", proto_config_args=copy.deepcopy(v_args)),
                  decoder=dict(proto_model="bert-generation", proto_config="bert-generation", proto_config_args=copy.deepcopy(d_args)))
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda().train()
    batch = synth.rrg_batch(2, 16, 800, n_images=2)
    batch["images_mask"] = torch.tensor([[True, True], [True, False]])
    out_ref = ref(batch["input_ids"], batch["attention_mask"], batch["images"], batch["images_mask"])
    out_ref["loss"].backward()
    out = mine(**batch)
    out["loss"].backward()
    torch.cuda.synchronize()
    assert abs(out["loss"].item() - out_ref["loss"].item()) <= 2e-2 * abs(out_ref["loss"].item())
    _check_grads(mine, ref, skip=("pooler",))
    # single-image branch (encoder_attention_mask=None)
    b1 = synth.rrg_batch(2, 16, 800)
    with torch.no_grad():
        l1 = mine.eval()(**b1)["loss"].item()
    assert abs(l1 - ref(b1["input_ids"], b1["attention_mask"], b1["images"])["loss"].item()) <= 2e-2 * abs(l1)


def test_mvqa(cuda_dev):
    from oracle.models import OracleMVQA
    from vilmedic_b200 import synth
    from vilmedic_b200.models import MVQA
    torch.manual_seed(0)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **dict(synth.vit_b16(), num_hidden_layers=2))
    tr = dict(hidden_size=768, num_hidden_layers=2, num_attention_heads=8, intermediate_size=2048, layer_norm_eps=1e-12,
              hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)           # config/MVQA/vqa.yml:38-47 shape (dh 96)
    cl = dict(proto="Classifier", input_size=768, num_classes=330, dropout=0.0)
    ad = dict(input_size=768, output_size=768)
    ls = dict(proto="LabelSmoothingCrossEntropy", smoothing=0.1)
    ref = OracleMVQA(cnn, cl, ad, tr, ls).eval()
    mine = MVQA(copy.deepcopy(cnn), copy.deepcopy(cl), copy.deepcopy(ad), copy.deepcopy(tr), copy.deepcopy(ls))
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda().train()
    g = torch.Generator().manual_seed(5)
    images = torch.randn(4, 3, 224, 224, generator=g)
    labels = torch.randint(0, 330, (4,), generator=g)
    o_ref = ref(images, labels)
    o_ref["loss"].backward()
    o = mine(images, labels)
    o["loss"].backward()
    torch.cuda.synchronize()
    assert abs(o["loss"].item() - o_ref["loss"].item()) <= 2e-2 * abs(o_ref["loss"].item())
    assert (o["output"].float().cpu() - o_ref["output"]).abs().max().item() <= 5e-2
    _check_grads(mine, ref)


@pytest.mark.parametrize("loss_proto", ["ConVIRTLoss", "InfoNCELoss"])
def test_convirt(cuda_dev, loss_proto):
    from oracle.models import OracleConVIRT
    from vilmedic_b200 import synth
    from vilmedic_b200.models import ConVIRT
    torch.manual_seed(0)
    enc = dict(proto=None, add_pooling_layer=True, vocab_size=600, hidden_size=768, num_hidden_layers=2, num_attention_heads=12,
               intermediate_size=3072, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, max_position_embeddings=64)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **dict(synth.vit_b16(), num_hidden_layers=2))
    proj = dict(visual_embedding_dim=768, textual_embedding_dim=768, projection_dim=256)
    loss = dict(proto=loss_proto, tau=0.1, lambda_=0.75) if loss_proto == "ConVIRTLoss" else dict(proto=loss_proto, tau=0.1)

    class _CLS(torch.nn.Module):
        """ConVIRT feeds a [b, D] image vector to vis_proj; with a ViT tower we take the CLS token in both paths."""
    ref = OracleConVIRT(enc, cnn, proj, loss).eval()
    mine = ConVIRT(copy.deepcopy(enc), copy.deepcopy(cnn), copy.deepcopy(proj), copy.deepcopy(loss), forward_batch_size=4)
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda().train()
    b = synth.rrg_batch(8, 32, 600, seed=11)
    # both paths: image vector = visual(images)[:, 0] is not what the reference does for CNNs (avgpool vector); with a
    # ViT backbone `visual(im)` is [b,S,D] and vis_proj acts on every token — the contrastive loss then needs [N,D]:
    # restrict to the CLS row on both sides by monkey-patching the towers' output selection identically.
    ref_forward = ref.visual.forward
    ref.visual.forward = lambda im: ref_forward(im)[:, 0]
    mine_forward = mine.visual.forward
    mine.visual.forward = lambda im, **kw: mine_forward(im, **kw)[:, 0].contiguous()
    o_ref = ref(b["input_ids"], b["attention_mask"], b["images"])
    o_ref["loss"].backward()
    o = mine(b["input_ids"], b["attention_mask"], b["images"])
    o["loss"].backward()
    torch.cuda.synchronize()
    assert abs(o["loss"].item() - o_ref["loss"].item()) <= 2e-2 * abs(o_ref["loss"].item()), (o["loss"].item(), o_ref["loss"].item())
    _check_grads(mine, ref, tol=2.5e-1)   # 8-sample contrastive loss through two towers: bf16 noise is amplified


def test_gloria_model_forward_backward(cuda_dev):
    """`model.proto: GLoRIA` (vilmedic/models/selfsup/GLoRIA.py:47-130; config/SELFSUP/gloria-mimic.yml): built through create_model from
    the synthetic YAML config (smaller towers), one training forward + backward.  The loss must equal GLoRIALoss (golden-tested against
    the reference's own file) evaluated on the features the model returns, and gradients must reach the text tower, the layer3 tap
    (local embedder + ResNet stages up to layer3) and layer4 (global path)."""
    import os
    from vilmedic_b200 import executors
    from vilmedic_b200.blocks.losses import GLoRIALoss
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    config = executors.load_config(os.path.join(root, "config/SELFSUP/synthetic-gloria-resnet50.yml"),
                                   ["model.encoder.num_hidden_layers=4", "model.encoder.hidden_dropout_prob=0.0",
                                    "model.encoder.attention_probs_dropout_prob=0.0", "model.forward_batch_size=3",
                                    "dataset.seq.tokenizer_max_len=16", "dataset.seq.vocab_size=400", "model.encoder.vocab_size=400"])
    tcfg = executors.utils.get(config, "trainor")
    dl = executors.SyntheticLoader(tcfg, 6, n_batches=1)
    model = executors.create_model(tcfg, dl).train()
    batch = {k: v for k, v in next(iter(dl)).items() if v is not None}
    out = model(**batch)
    loss = out["loss"]
    assert loss.dim() == 0 and torch.isfinite(loss)
    assert out["local_features"].shape == (6, 768, 19, 19) and out["global_features"].shape == (6, 768)
    assert out["word_embeddings"].shape == (6, 768, 16) and out["sent_embeddings"].shape == (6, 768)
    # same loss from the returned features
    ids = batch["input_ids"]
    _, sents = model.aggregate_tokens(torch.zeros(4, 6, 16, 8), ids)
    again, _ = GLoRIALoss(**dict(config.model.loss))(out["global_features"].detach(), out["local_features"].detach(),
                                                     out["word_embeddings"].detach(), out["sent_embeddings"].detach(), sents)
    assert abs(again.item() - loss.item()) <= 1e-4 * abs(loss.item()) + 1e-5
    loss.backward()
    torch.cuda.synchronize()
    for p in (model.local_embedder.weight, model.global_embedder.weight, model.linguistic.encoder.encoder.layer[0].intermediate.dense.weight,
              model.visual.model[6][0].conv1.weight, model.visual.model[7][0].conv1.weight, model.visual.model[0].weight):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum().item() > 0
    assert callable(model.eval_func)
