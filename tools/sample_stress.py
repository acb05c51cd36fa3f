"""Race detector for the sampling path: the 128-frame draft-and-revise (fused Gumbel-max head, attention with two issuing
threads, split-KV at small batch) and the 16-frame maskgit loop, each run N times from the same seeds; the sampled videos must be
identical every time (counter-hash / Philox noise is a pure function of the seed and the position).
usage: python tools/sample_stress.py [repeats]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dev = torch.device("cuda:0")
    for workload, B in (("sample128f", 4), ("sample128f", 2), ("maskgit16f", 8)):
        w = bench.make_step(workload, bench.CONFIGS[workload], B, 0.0, dev, 0, 1)
        ref, bad = None, 0
        t0 = time.time()
        for r in range(reps):
            torch.manual_seed(1234)
            w.model.rng_seed = 1000
            w.model._rng_offset = 0                      # the noise stream positions (advance with every sampling call)
            w.model.mask_sampler.rng_offset = 0
            w.model.mask_sampler.rng_seed = 1001
            out = w.device()
            torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
            elif not torch.equal(out, ref):
                bad += 1
                print(f"{workload} B={B} repeat {r}: {(out != ref).float().mean().item():.4f} of the tokens differ")
        print(f"{workload} B={B}: {reps} runs, {bad} differ ({time.time() - t0:.1f} s)")
        bench.release(w)
        if bad:
            sys.exit(1)


if __name__ == "__main__":
    main()
