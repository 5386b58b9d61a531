"""Find what breaks CUDA-graph capture of a workload's training step: runs the step eagerly, then under capture with the full
traceback printed.  Usage: python tools/capture_debug.py mvqa|convirt [batch]"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from vilmedic_b200 import executors, ops
    wl = sys.argv[1]
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    path = {"convirt": "config/SELFSUP/synthetic-convirt-resnet50.yml", "mvqa": "config/MVQA/synthetic-vit-b16.yml"}[wl]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    config = executors.load_config(os.path.join(root, path), ["model.cnn.num_hidden_layers=2"] if wl == "mvqa" else [])
    tcfg = executors.utils.get(config, "trainor")
    dl = executors.SyntheticLoader(tcfg, bs, n_batches=1)
    model = executors.create_model(tcfg, dl).train()
    opt = executors.create_optimizer(tcfg, None, model)
    batch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in next(iter(dl)).items() if v is not None}
    counter = torch.zeros(1, device="cuda", dtype=torch.int64)
    ops.RNG_COUNTER[0] = counter

    def step():
        out = model(**batch)
        out["loss"].backward()
        opt.step()
        return out["loss"]

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            print("eager loss (side stream)", float(step()))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                loss = step()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        print("capture ok, replay loss", float(loss))
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
