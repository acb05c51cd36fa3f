"""Phase timing of one train16f step (CUDA events): forward, CE, backward, AdamW, operand refresh.
usage: python tools/train_probe.py [batch]"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench  # noqa: E402
from mebt_b200 import _lib, ops  # noqa: E402
from mebt_b200.training import TrainState  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    cfg = bench.CONFIGS["train16f"]
    model = bench.build_native_model(cfg, bench.synth_weights(cfg), float(sys.argv[2]) if len(sys.argv) > 2 else 0.1, torch.device("cuda", 0)).train()
    ts = TrainState(model)
    opt = ts.make_optimizer()
    x, idx = bench.synth_batch(cfg, B, 1)
    x, idx = x.cuda(), idx.cuda()
    for _ in range(3):
        ts.train_step(opt, x, idx, t=0.5)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    names = ["forward", "ce", "backward", "adamw", "refresh"]
    acc = {n: 0.0 for n in names}
    reps = 10
    m = ts.model
    for _ in range(reps):
        xi = x.reshape(B, -1)
        ctx_idx, tgt_idx, seq_len = m.mask_sampler.divide_indices(idx, torch.tensor(0.5), m.t_lengths,
                                                                  m.t_prior(m.t_lengths, 0))
        zt = torch.gather(xi, 1, tgt_idx)
        ev[0].record()
        logits = ts.forward(xi, ctx_idx, tgt_idx)
        ev[1].record()
        stats, _ = ops.masked_ce(logits, zt.reshape(-1), 0.0, dlogits=logits, grad_scale=1.0 / (B * 512))
        ev[2].record()
        ts.backward(logits)
        ev[3].record()
        opt.step()
        ev[4].record()
        ts.refresh_operands()
        ev[5].record()
        torch.cuda.synchronize()
        for i, n in enumerate(names):
            acc[n] += ev[i].elapsed_time(ev[i + 1])
    tot = sum(acc.values()) / reps
    print(f"B={B} total {tot:.2f} ms/step: " + ", ".join(f"{n} {v / reps:.2f}" for n, v in acc.items()))
    _lib.profile_enable(True)
    ts.train_step(opt, x, idx, t=0.5)
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    print({k: (round(v["ms"], 2), v["launches"]) for k, v in rep.items() if v["launches"]})


if __name__ == "__main__":
    main()
