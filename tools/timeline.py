#!/usr/bin/env python
"""Device timeline of one bench step (CUPTI through torch.profiler; nsys is not in the image).

    python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f.json

Writes every kernel / memcpy / memset of ONE step with its stream, start and duration (microseconds, relative to the
first kernel), plus a per-kernel summary.  Unlike the ncu launch list this keeps the real concurrency between the
streams and the programmatic-dependent-launch overlap, so chain length and exposed gaps can be read from it.
`tools/timeline_report.py` turns the file into the tables kept under profiles/.
"""
from __future__ import annotations

import argparse
import json
import random
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="train16f")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--out", default="gpurun_out/timeline.json")
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    cfg = bench.CONFIGS[args.workload]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    step = bench.make_step(args.workload, cfg, args.batch, args.dropout, dev, rank=0, world=1).device
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ev.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name, getattr(e, "device_index", 0),
                       getattr(e, "stream", None)))
    # torch's FunctionEvent does not always expose the stream: read it from the raw kineto events
    kev = []
    for k in prof.profiler.kineto_results.events():
        if str(k.device_type()).endswith("CUDA"):
            kev.append((k.start_ns() / 1e3, k.duration_ns() / 1e3, k.name(), k.device_resource_id()))
    kev.sort()
    t0 = kev[0][0] if kev else 0.0
    out = {"workload": args.workload, "steps": args.steps,
           "events": [[round(s - t0, 3), round(d, 3), n[:96], int(r)] for s, d, n, r in kev]}
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out))
    span = (kev[-1][0] + kev[-1][1] - t0) if kev else 0.0
    print(f"{len(kev)} device events over {span / 1e3:.3f} ms -> {args.out}")


if __name__ == "__main__":
    main()
