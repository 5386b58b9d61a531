"""CUDA-graph captured training step (vilmedic_b200/graph.py): replay == eager, and the prefetching input path (host -> device
copy of the next batch on a copy stream + device-to-device hand-over) delivers exactly the batch that was prefetched."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_step_and_prefetch(cuda_dev):
    from vilmedic_b200 import ops, synth
    from vilmedic_b200.graph import GraphedTrainStep
    from vilmedic_b200.models import RRG
    from vilmedic_b200.optim import FusedAdamW
    torch.manual_seed(0)
    dec = synth.bert_base_decoder(vocab=400, layers=1, dropout=0.0)
    cnn = dict(proto="VisualEncoder", backbone="vit", permute="no_permute", **dict(synth.vit_b16(), num_hidden_layers=1))
    model = RRG(copy.deepcopy(dec), copy.deepcopy(cnn)).cuda().train()
    opt = FusedAdamW(model, lr=0.0, weight_decay=0.0)                 # parameters stay put: the loss is a function of the batch only
    b1, b2 = synth.rrg_batch(2, 16, 400, seed=1), synth.rrg_batch(2, 16, 400, seed=2)
    pin = lambda b: {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    b1, b2 = pin(b1), pin(b2)

    def eager(b):
        out = model(**b)
        out["loss"].backward()
        opt.step()
        return out["loss"].item()

    l1, l2 = eager(b1), eager(b2)
    assert abs(l1 - l2) > 1e-4                                        # the two batches are distinguishable
    dev1 = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in b1.items()}
    g = GraphedTrainStep(model, opt, dev1, warmup=1,
                         step_fn=lambda b: (ops.rng_advance(ops.RNG_COUNTER[0], 4096), _step(model, opt, b))[1])
    assert abs(g(b1).item() - l1) < 1e-5 and abs(g(b2).item() - l2) < 1e-5
    g.prefetch(b1)
    a = g.replay_prefetched()
    g.prefetch(b2)                                                    # travels while the step above may still be running
    la = a.item()
    b = g.replay_prefetched()
    g.prefetch(b1)
    lb = b.item()
    c = g.replay_prefetched().item()
    assert abs(la - l1) < 1e-5 and abs(lb - l2) < 1e-5 and abs(c - l1) < 1e-5, (la, lb, c, l1, l2)
    # loss read-back one step behind (the host never waits for the step it has just launched): values arrive in order, none is lost
    g.prefetch(b2)
    got = [g.step_prefetched_async(b1), g.step_prefetched_async(b2), g.step_prefetched_async(b2), g.drain()]
    assert got[0] is None and g.drain() is None
    assert abs(got[1] - l2) < 1e-5 and abs(got[2] - l1) < 1e-5 and abs(got[3] - l2) < 1e-5, (got, l1, l2)


def _step(model, opt, batch):
    out = model(**batch)
    loss = out["loss"]
    loss.backward()
    opt.step()
    return loss
