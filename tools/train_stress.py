"""Race detector for the training engine: the same forward + loss + backward of the STL-16f model (B = 6, dropout 0.1, fixed
dropout seed) N times from unchanged weights; the loss and the gradient of every block / head parameter must be bit-identical
each time (all reductions run in a fixed order; only the embedding scatter-add uses atomics and is excluded).
usage: python tools/train_stress.py [iterations]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    dev = torch.device("cuda:0")
    cfg = bench.CONFIGS["train16f"]
    state = bench.synth_weights(cfg)
    model = bench.build_native_model(cfg, state, 0.1, dev)
    model.train()
    from mebt_b200.training import TrainState
    ts = TrainState(model, n_buckets=4)
    x, idx = bench.synth_batch(cfg, 6, 100)
    x, idx = x.to(dev), idx.to(dev)
    lo, hi = ts.block_slices[0][0], ts.head_slice[1]
    side = torch.cuda.Stream()
    big = torch.empty(2, 64 << 20, device=dev, dtype=torch.uint8)
    ref_loss = ref_grad = None
    bad = 0
    t0 = time.time()
    for it in range(iters):
        if it % 5 == 0:
            with torch.cuda.stream(side):
                big[1].copy_(big[0], non_blocking=True)
        torch.manual_seed(1234)                                  # same dropout seed / masks every time
        np.random.seed(7)                                        # ... and the same training window (MaskGen draws it with numpy)
        out = ts.loss_and_backward(x, idx, t=bench.TRAIN_T)
        g = ts.flat_grad[lo:hi]
        if ref_grad is None:
            torch.cuda.synchronize()
            ref_loss, ref_grad = out["loss"].clone(), g.clone()
            assert torch.isfinite(ref_grad).all()
        else:
            same = bool(torch.equal(out["loss"], ref_loss)) and bool(torch.equal(g, ref_grad))
            if not same:
                bad += 1
                d = (g - ref_grad).abs().max().item()
                print(f"iteration {it}: differs (loss {float(out['loss'])} vs {float(ref_loss)}, max grad diff {d:.3e})")
    torch.cuda.synchronize()
    print(f"{iters} iterations, {bad} differ from the first ({time.time() - t0:.1f} s)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
