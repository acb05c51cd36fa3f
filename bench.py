#!/usr/bin/env python
"""Benchmark of the MeBT hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sample128f|sample16f|train16f] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input:

  sample128f  128-frame draft-and-revise sampling (BASELINE.json configs[2]): token grid [32,16,16] = 8192 tokens,
              24-layer STL model, script defaults n_draft=8, n_revise=8, M=2 -> 24 forwards per video,
              sum(NT) = 6.5 * 8192 = 53 248 masked-token predictions per video.  Videos are sharded by batch over
              the ranks (no collective on the data path): weak scaling, B videos per GPU.
  sample16f   the same on the 16-frame model (N = 1024).

metric = masked video tokens/s = (all ranks' B * sum(NT)) / max-over-ranks device time.
`value`: inputs resident in HBM.  `e2e`: through the public API (`Net2NetTransformer.draft_and_revise`) with the
token grid coming from pinned host memory and the sampled ids copied back to the host inside the timed region.
`--impl reference` times the CPU oracle (a torch-CPU restatement of the reference; /root/reference does not exist
on the GPU box) on a bounded sample of the same workload with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

STL_MODES = ["latent_enc", "latent_self"] * 6 + ["latent_enc"] + ["latent_dec", "lt2l"] * 5 + ["latent_dec"]
CONFIGS = {
    "sample128f": dict(n_embd=1024, n_head=16, sos_emb=256, block_size=8192, shape=[32, 16, 16], n_layer=24,
                       vocab_size=16384, avg_loss=1.0, mode=STL_MODES),
    "sample16f": dict(n_embd=1024, n_head=16, sos_emb=256, block_size=1024, shape=[4, 16, 16], n_layer=24,
                      vocab_size=16384, avg_loss=1.0, mode=STL_MODES),
}
DNR = dict(n_draft=8, draft_t=1.0, n_revise=8, revise_t=1.0, M=2)


def masked_tokens_per_video(N: int) -> int:
    d, r, M = DNR["n_draft"], DNR["n_revise"], DNR["M"]
    return sum(N - i * (N // d) for i in range(d)) + M * r * (N // r)


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_cpu_state(cfg, seed=0):
    """Random-init weights in the reference's distribution (N(0,0.02), zero biases, unit LayerNorm), keyed like
    the reference state_dict.  The same dict feeds the GPU model and the CPU oracle."""
    from helpers import model_configs
    from mebt_b200.transformer import Net2NetTransformer
    torch.manual_seed(seed)
    params, vq, mask = model_configs(cfg, schedule="cosine")
    model = Net2NetTransformer(params, vq, mask)
    return model


def cpu_oracle_step(cfg, state, threads: int, reps: int):
    """One forward + sampling step of the oracle at NC = NT = N/2, B = 1.  -> (tokens/s, description)."""
    from oracle import mebt_oracle as O
    torch.set_num_threads(threads)
    N = int(np.prod(cfg["shape"]))
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, cfg["vocab_size"], (1, N), generator=g)
    perm = torch.randperm(N, generator=g).view(1, N)
    ctx, tgt = perm[:, : N // 2], perm[:, N // 2:]
    times = []
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            logits = O.reconstruct_mask(state, cfg, x, ctx, tgt)
            q = torch.empty_like(logits).exponential_()
            O.sample_from_logits(logits, 1.0, None, None, q)
            times.append(time.perf_counter() - t0)
    t = float(np.median(times[1:]))
    return (N // 2) / t, f"1 of 24 forward+sample steps per video: NC=NT={N // 2}, B=1, fp32, median of {reps} after 1 warm-up"


def run_reference(args, cfg, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    model = build_cpu_state(cfg)
    state = {k: v.detach() for k, v in model.state_dict().items()}
    from oracle import mebt_oracle as O
    torch.set_num_threads(threads)
    N = int(np.prod(cfg["shape"]))
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, cfg["vocab_size"], (1, N), generator=g)
    perm = torch.randperm(N, generator=g).view(1, N)
    ctx, tgt = perm[:, : N // 2], perm[:, N // 2:]

    def step():
        with torch.no_grad():
            logits = O.reconstruct_mask(state, cfg, x, ctx, tgt)
            O.sample_from_logits(logits, 1.0, None, None, torch.empty_like(logits).exponential_())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = (N // 2) / dt
    sample = f"each step = 1 of the 24 forward+sample steps of one video (NC=NT={N // 2}, B=1, fp32 torch-CPU oracle port)"
    print(json.dumps({
        "impl": "reference", "metric": "masked video tokens/sec", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "tokens": N, "sampler": DNR},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="sample128f", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=4, help="videos per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from mebt_b200 import _lib
    _lib.check(_lib.lib.mebt_device_check(), "mebt_device_check")
    cpu_model = build_cpu_state(cfg)                      # same weights on every rank (seed 0)
    state = {k: v.detach().clone() for k, v in cpu_model.state_dict().items()} if rank == 0 else None
    model = cpu_model.to(dev).eval()
    model.rng_mode, model.rng_seed = "philox", 1000 + rank    # per-rank noise streams: different videos per rank
    B = args.batch
    N = int(np.prod(cfg["shape"]))
    tokens_per_step = B * masked_tokens_per_video(N)
    x_host = torch.zeros(B, *cfg["shape"], dtype=torch.long).pin_memory()
    out_host = torch.empty(B, N, dtype=torch.long).pin_memory()
    x_dev = x_host.to(dev)

    def step_device():
        return model.draft_and_revise(x_dev, None, **DNR)

    def step_e2e():
        x = x_host.to(dev, non_blocking=True)
        ids = model.draft_and_revise(x, None, **DNR)
        out_host.copy_(ids, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    torch.manual_seed(1234 + rank)                        # CPU generator: the randperm draws of the gibbs masks
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = (_lib.launch_count() - launches0) // args.steps
    clock_info = clocks.stop() if rank == 0 else None

    # end to end through the public API with host buffers
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
        torch.cuda.synchronize()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps

    # per-kernel-family timing of one more step (events on the launch stream), for the roofline block
    _lib.profile_enable(True)
    step_device()
    prof = _lib.profile_report()
    _lib.profile_enable(False)

    if rank == 0:
        pk = peaks()
        gemm = prof["gemm"]
        total_ms = sum(f["ms"] for f in prof.values())
        achieved = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tf_sustained"], "traffic": None, "kernel": "gemm_bf16_kernel (tcgen05)",
                    "peak_source": f"{pk['src']} bf16 sustained (kernel timed inside a long step)",
                    "share_of_step": gemm["ms"] / total_ms if total_ms else None,
                    "families_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]},
                    "families_launches": {k: v["launches"] for k, v in prof.items() if v["launches"]}}
        samp = prof["sample"]
        if samp["ms"] > 0:
            roofline["sample_kernel_hbm"] = {"achieved_gbs": samp["work"] / (samp["ms"] * 1e-3) / 1e9,
                                             "peak_gbs": pk["hbm"],
                                             "frac": samp["work"] / (samp["ms"] * 1e-3) / 1e9 / pk["hbm"]}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, desc = cpu_oracle_step(cfg, state, threads, reps=2)
            cpu_baseline = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": desc}
        value = world * tokens_per_step / (ms * 1e-3)
        print(json.dumps({
            "metric": "masked video tokens/sec", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "tokens": N, "videos_per_gpu": B, "sampler": DNR,
                       "masked_tokens_per_video": masked_tokens_per_video(N), "noise": "in-kernel philox",
                       "weights": "random init, reference distribution", "logits": "fp32 materialised",
                       "l2": "working set (0.67 GB bf16 weights + GB-scale logits) exceeds the 126 MB L2; no flush needed",
                       "generated_tokens_per_s": world * B * N / (ms * 1e-3)},
            "e2e": {"value": world * tokens_per_step / e2e_s, "unit": "tokens/s",
                    "h2d_bytes_per_step": int(x_host.numel() * 8), "d2h_bytes_per_step": int(out_host.numel() * 8)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clock_info}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
