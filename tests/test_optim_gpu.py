"""Fused optimizers (vilmedic_b200/optim.py, csrc/optim.cu) against torch.optim on the same GPU: the three names the
reference's configs select (vilmedic/executors/utils.py:81-86; config/: RAdam, Adam; bench: AdamW), frozen parameters,
checkpoint round trip, device-side NaN/Inf skip (vilmedic/executors/trainor.py:109-112)."""
import copy
import io

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


class _Toy(nn.Module):
    """A few oddly sized parameters so that the arena has alignment gaps; one block can be frozen."""

    def __init__(self):
        super().__init__()
        self.a = nn.Linear(37, 19)
        self.b = nn.Linear(19, 11)
        self.c = nn.Linear(11, 5)


def _pair(seed=0):
    torch.manual_seed(seed)
    ref = _Toy().cuda()
    mine = copy.deepcopy(ref)
    return ref, mine


def _set_grads(ref, mine, it, scale=1.0):
    from vilmedic_b200.arena import get_arena
    a = get_arena(mine)
    a.bind_grads()
    g = torch.Generator(device="cuda").manual_seed(100 + it)
    for pr, pm in zip(ref.parameters(), mine.parameters()):
        gr = torch.randn(pr.shape, device="cuda", generator=g) * scale
        pr.grad = gr.clone()
        if pm.requires_grad:
            pm.grad.copy_(gr)


@pytest.mark.parametrize("name,kw", [
    ("AdamW", dict(lr=1e-2, weight_decay=0.05)),
    ("Adam", dict(lr=1e-2, weight_decay=0.0)),
    ("Adam", dict(lr=3e-3, weight_decay=0.1, betas=[0.8, 0.99])),
    ("RAdam", dict(lr=1e-2, weight_decay=0.0)),          # config/RRG/biomed-roberta-baseline-mimic.yml:37-40
    ("RAdam", dict(lr=5e-3, weight_decay=0.02, eps=1e-6)),
])
def test_fused_optimizers_match_torch(cuda_dev, name, kw):
    from vilmedic_b200.optim import create_optimizer
    ref, mine = _pair()
    tkw = dict(kw)
    if "betas" in tkw:
        tkw["betas"] = tuple(tkw["betas"])
    opt_ref = getattr(torch.optim, name)(ref.parameters(), **tkw)
    opt = create_optimizer(name, mine, **kw)
    for it in range(12):                 # RAdam switches to the rectified update at step 6 (rho_t > 5)
        _set_grads(ref, mine, it, scale=1.0 + it)
        opt_ref.step()
        opt.step()
    torch.cuda.synchronize()
    assert opt.step_t.item() == 12
    for (n, pr), pm in zip(ref.named_parameters(), mine.parameters()):
        err = (pr.detach() - pm.detach()).abs().max().item()
        assert err < 2e-6 + 2e-6 * pr.abs().max().item(), (name, n, err)
        assert pm.grad.abs().sum().item() == 0          # fused zero_grad
    from vilmedic_b200.arena import get_arena
    a = get_arena(mine)
    assert torch.equal(a.flat_bf16, a.flat.to(torch.bfloat16))     # bf16 mirror written by the same kernel


def test_clip_grad_norm_matches_torch(cuda_dev):
    from vilmedic_b200.optim import FusedRAdam
    ref, mine = _pair(1)
    opt_ref = torch.optim.RAdam(ref.parameters(), lr=1e-2)
    opt = FusedRAdam(mine, lr=1e-2, max_grad_norm=0.5)
    for it in range(8):
        _set_grads(ref, mine, it, scale=3.0)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=0.5)
        opt_ref.step()
        opt.step()
    for pr, pm in zip(ref.parameters(), mine.parameters()):
        assert (pr.detach() - pm.detach()).abs().max().item() < 5e-6


def test_frozen_parameters_stay_bit_identical(cuda_dev):
    """the fused step must neither decay nor update requires_grad=False parameters, even when a backward kernel
    accumulated into their gradient slots."""
    from vilmedic_b200.arena import get_arena
    from vilmedic_b200.optim import FusedAdamW
    ref, mine = _pair(2)
    for m in (ref, mine):
        for p in m.b.parameters():
            p.requires_grad = False
    frozen_before = [p.detach().clone() for p in mine.b.parameters()]
    opt_ref = torch.optim.AdamW([p for p in ref.parameters() if p.requires_grad], lr=1e-2, weight_decay=0.1)
    opt = FusedAdamW(mine, lr=1e-2, weight_decay=0.1)
    a = get_arena(mine)
    for it in range(4):
        _set_grads(ref, mine, it)
        a.flat_grad.add_(0.25)       # junk in every slot, including the frozen ones and the alignment padding
        for p in ref.parameters():
            if p.requires_grad:
                p.grad.add_(0.25)
        opt_ref.step()
        opt.step()
        assert a.flat_grad.abs().sum().item() == 0
    for p, q in zip(mine.b.parameters(), frozen_before):
        assert torch.equal(p.detach(), q)
    for pr, pm in zip(ref.parameters(), mine.parameters()):
        assert (pr.detach() - pm.detach()).abs().max().item() < 5e-6


def test_frozen_visual_encoder_unchanged_after_steps(cuda_dev):
    """VisualEncoder(freeze=True) (vilmedic/blocks/vision/visual_encoder.py:124-128): the ViT stays bit-identical over
    training steps of the full RRG model while the decoder trains."""
    from vilmedic_b200 import synth
    from vilmedic_b200.models import RRG
    from vilmedic_b200.optim import FusedAdamW
    torch.manual_seed(0)
    dec = synth.bert_base_decoder(vocab=300, layers=1, dropout=0.0)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", freeze=True, **dict(synth.vit_b16(), num_hidden_layers=1))
    model = RRG(dec, cnn).cuda().train()
    before = {k: v.detach().clone() for k, v in model.enc.state_dict().items()}
    dec_before = model.dec.decoder.lm_head.bias.detach().clone()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=0.1)
    batch = synth.rrg_batch(2, 16, 300)
    for _ in range(3):
        model(**batch)["loss"].backward()
        opt.step()
    torch.cuda.synchronize()
    for k, v in model.enc.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert not torch.equal(model.dec.decoder.lm_head.bias.detach(), dec_before)


def test_optimizer_state_dict_round_trip(cuda_dev):
    """Resume = same trajectory (reference: trainor.py:194-199 saves optimizer.state_dict(), utils.py:90-92 reloads it)."""
    from vilmedic_b200.optim import create_optimizer
    _, mine = _pair(3)
    ref_model = copy.deepcopy(mine)
    opt = create_optimizer("RAdam", mine, lr=1e-2)
    for it in range(7):
        _set_grads(ref_model, mine, it)
        opt.step()
    buf = io.BytesIO()
    torch.save({"model": mine.state_dict(), "optimizer": opt.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf, map_location="cpu", weights_only=False)
    resumed = _Toy().cuda()
    resumed.load_state_dict(ck["model"])
    opt2 = create_optimizer("RAdam", resumed, state_dict=ck, lr=1e-2)
    assert opt2.step_t.item() == 7 and torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
    for it in range(7, 10):
        _set_grads(ref_model, mine, it)
        _set_grads(ref_model, resumed, it)
        opt.step()
        opt2.step()
    for p, q in zip(mine.parameters(), resumed.parameters()):
        assert torch.equal(p.detach(), q.detach())
    with pytest.raises(KeyError):
        opt2.load_state_dict(torch.optim.Adam(resumed.parameters()).state_dict())


def test_nan_loss_skips_step_on_device(cuda_dev):
    from vilmedic_b200.optim import FusedAdam
    ref, mine = _pair(4)
    opt = FusedAdam(mine, lr=1e-2)
    good = torch.tensor(1.5, device="cuda")
    _set_grads(ref, mine, 0)
    opt.step(loss=good)
    after_one = [p.detach().clone() for p in mine.parameters()]
    m_one = opt.m.clone()
    for bad in (float("nan"), float("inf")):
        _set_grads(ref, mine, 1)
        opt.step(loss=torch.tensor(bad, device="cuda"))
        for p, q in zip(mine.parameters(), after_one):
            assert torch.equal(p.detach(), q)
            assert p.grad.abs().sum().item() == 0        # optimizer.zero_grad() of the reference's skip branch
    # non-finite gradient with a finite loss (GradScaler found-inf semantics)
    _set_grads(ref, mine, 2)
    next(mine.parameters()).grad.view(-1)[3] = float("inf")
    opt.step(loss=good)
    assert torch.equal(opt.m, m_one)
    assert opt.step_t.item() == 1 and opt.skipped_steps.item() == 3
    _set_grads(ref, mine, 3)
    opt.step(loss=good)
    assert opt.step_t.item() == 2 and not torch.equal(opt.m, m_one)


def test_create_optimizer_errors(cuda_dev):
    from vilmedic_b200.optim import create_optimizer
    _, mine = _pair(5)
    with pytest.raises(NotImplementedError):
        create_optimizer("SGD", mine, lr=0.1)
    with pytest.raises(ValueError):
        create_optimizer("Adam", mine)
    with pytest.raises(NotImplementedError):
        create_optimizer("Adam", mine, lr=0.1, amsgrad=True)


@pytest.mark.parametrize("name", ["AdamW", "RAdam"])
def test_bf16_gradient_payload(cuda_dev, name):
    """step(grad16=...) — the exchange payload of ddp.GradSync: the gradient values come from the bf16 buffer (fp32 buffer only
    zeroed), the clip norm too.  Must equal torch.optim fed with exactly those bf16-rounded gradients."""
    from vilmedic_b200.arena import get_arena
    from vilmedic_b200.optim import create_optimizer
    ref, mine = _pair(3)
    opt_ref = getattr(torch.optim, name)(ref.parameters(), lr=1e-2, weight_decay=0.01)
    opt = create_optimizer(name, mine, lr=1e-2, weight_decay=0.01, max_grad_norm=0.7)
    a = get_arena(mine)
    g16 = torch.zeros(a.numel, device="cuda", dtype=torch.bfloat16)
    for it in range(8):
        _set_grads(ref, mine, it, scale=2.0)
        g16.copy_(a.flat_grad)                                   # what GradSync's cast kernel + all-reduce leave behind
        a.flat_grad.mul_(3.0)                                    # must NOT be read: poison the fp32 values
        for pr, pm in zip(ref.parameters(), mine.parameters()):
            off = a.offsets[id(pm)]
            pr.grad = g16[off:off + pm.numel()].float().view(pr.shape).clone()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=0.7)
        opt_ref.step()
        opt.step(grad16=g16)
    torch.cuda.synchronize()
    for (n, pr), pm in zip(ref.named_parameters(), mine.parameters()):
        err = (pr.detach() - pm.detach()).abs().max().item()
        assert err < 4e-6 + 4e-6 * pr.abs().max().item(), (name, n, err)
    assert a.flat_grad.abs().sum().item() == 0                   # fp32 accumulation buffer zeroed for the next backward
    with pytest.raises(ValueError):
        opt.step(grad16=g16[:-4])
