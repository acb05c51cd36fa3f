"""Steady-state ms/step of the train16f step under the overlap options (CUDA events over 10 steps after 3 warm-ups).
usage: python tools/train_sweep.py"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench  # noqa: E402
from mebt_b200.training import TrainState  # noqa: E402


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    cfg = bench.CONFIGS["train16f"]
    dev = torch.device("cuda", 0)
    model = bench.build_native_model(cfg, bench.synth_weights(cfg), 0.1, dev).train()
    ts = TrainState(model, n_buckets=4)
    opt = ts.make_optimizer()
    x, idx = bench.synth_batch(cfg, 6, 1)
    x, idx = x.to(dev), idx.to(dev)
    n = len(ts.modes)
    print("sequential update            : %.3f ms" % timed(lambda: ts.train_step(opt, x, idx, t=0.5), reps=20))
    if "--base" in sys.argv:
        print("sequential update (repeat)   : %.3f ms" % timed(lambda: ts.train_step(opt, x, idx, t=0.5), reps=20))
        return
    for nb in (4, 8, 12):
        edges = [round(i * n / nb) for i in range(nb + 1)]
        ts.chunks = [(edges[i], edges[i + 1]) for i in range(nb)]
        for ctas in (0, 24, 48, 96, 192):
            ts.update_ctas = ctas
            ms = timed(lambda: ts.train_step(opt, x, idx, t=0.5, overlap_update=True))
            print(f"overlap buckets={nb:2d} update_ctas={ctas:4d}: {ms:.3f} ms")


if __name__ == "__main__":
    main()
