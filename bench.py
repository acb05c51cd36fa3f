#!/usr/bin/env python
"""Benchmark of the MeBT hot path on B200 (contract: the task statement; summary in DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train16f|sample128f|sample16f] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input:

  train16f    (default; BASELINE.json configs[1]) STL 16-frame model — 24 blocks, D=1024, 16 heads, 256 latents,
              N = 1024 tokens, 337 M parameters — bf16 training step at batch 6 per GPU: stem -> stack forward ->
              fused masked CE -> full backward -> NCCL gradient all-reduce (N > 1) -> AdamW -> bf16 operand refresh.
              metric = masked tokens/s = all ranks' B * NT / step time, with t = 0.5 (NC = NT = 512).
  sample128f  (configs[2]) 128-frame draft-and-revise sampling, token grid [32,16,16] = 8192 tokens, script defaults
              n_draft=8, n_revise=8, M=2 -> 24 forwards and 53 248 masked-token predictions per video; videos sharded
              by batch over the ranks, no collective on the data path.
  sample16f   the same on the 16-frame model.

`value`: inputs resident in HBM, CUDA events, max over ranks.  `e2e`: through the public API with the batch coming
from pinned host memory and the result (loss / sampled ids) copied back to the host inside the timed region.
`--impl reference` times the CPU oracle (a torch-CPU restatement of the reference; /root/reference does not exist on
the GPU box) on the same workload (train) or a bounded sample of it (sampling) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import random
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
_BASE = dict(n_embd=1024, n_head=16, sos_emb=256, n_layer=24, vocab_size=16384, avg_loss=1.0, mode=STL_MODES)
CONFIGS = {
    "train16f": dict(_BASE, block_size=1024, shape=[4, 16, 16]),
    "sample16f": dict(_BASE, block_size=1024, shape=[4, 16, 16]),
    "sample128f": dict(_BASE, block_size=8192, shape=[32, 16, 16]),
    "maskgit16f": dict(_BASE, block_size=1024, shape=[4, 16, 16]),
    "vq16f": dict(_BASE, block_size=1024, shape=[4, 16, 16]),
}
MASKGIT = dict(temperature=1.0, top_k=None, top_p=None, n_steps=128, strategy="maskgit", context_temperature=6.0)
DNR = dict(n_draft=8, draft_t=1.0, n_revise=8, revise_t=1.0, M=2)
TRAIN_T = 0.5


def masked_tokens_per_video(N: int) -> int:
    d, r, M = DNR["n_draft"], DNR["n_revise"], DNR["M"]
    return sum(N - i * (N // d) for i in range(d)) + M * r * (N // r)


def maskgit_masked_tokens(N: int, n_steps: int) -> int:
    """sum of NT over the forwards of Net2NetTransformer.sample with the cosine schedule (float32 arithmetic, as the
    sampler evaluates it; steps whose target count is already below the schedule are skipped)."""
    nt, total = N, 0
    for t_next in np.linspace(0, 1, n_steps + 1)[1:]:
        n_masked = int(torch.ceil(torch.cos(0.5 * np.pi * torch.full((1,), fill_value=t_next)) * N)[0])
        if n_masked > nt:
            continue
        total += nt
        if N - n_masked > N - nt:
            nt = n_masked
    return total


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


def build_cpu_model(cfg, seed=0, pdrop=0.0):
    """Random-init weights in the reference's distribution (N(0,0.02), zero biases, unit LayerNorm), built on the
    CPU so that the same values feed the GPU model and the CPU oracle.  pdrop: embd/resid/attn dropout (STL yaml: 0.1)."""
    from helpers import model_configs
    from mebt_b200.transformer import Net2NetTransformer
    torch.manual_seed(seed)
    params, vq, mask = model_configs(cfg, schedule="linear")
    params.embd_pdrop = params.resid_pdrop = params.attn_pdrop = pdrop
    return Net2NetTransformer(params, vq, mask)


def synth_batch(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    N = int(np.prod(cfg["shape"]))
    x = torch.randint(0, cfg["vocab_size"], (B, *cfg["shape"]), generator=g)
    indices = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    return x, indices


# ---- CPU oracle legs --------------------------------------------------------------------------------------------------
def oracle_sampling_step(cfg, state):
    from oracle import mebt_oracle as O
    N = int(np.prod(cfg["shape"]))
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, cfg["vocab_size"], (1, N), generator=g)
    perm = torch.randperm(N, generator=g).view(1, N)
    ctx, tgt = perm[:, : N // 2], perm[:, N // 2:]

    def step():
        with torch.no_grad():
            logits = O.reconstruct_mask(state, cfg, x, ctx, tgt)
            O.sample_from_logits(logits, 1.0, None, None, torch.empty_like(logits).exponential_())
    desc = f"1 of the 24 forward+sample steps of one video (NC=NT={N // 2}, B=1, fp32 torch-CPU oracle port)"
    return step, N // 2, desc


def oracle_vq_step(B):
    from oracle import mebt_oracle as O
    torch.manual_seed(0)
    E = torch.randn(16384, 256)
    z = torch.randn(B, 256, 4, 16, 16, generator=torch.Generator().manual_seed(4))

    def step():
        with torch.no_grad():
            out = O.codebook_quantise(z, E)
            O.codebook_decode_gather(out["encodings"], E)
    return step, B * 1024, f"Codebook.forward + decode gather on {B} videos (fp32 torch-CPU oracle port)"


def oracle_dropout_masks(cfg, B, NC, NT, p):
    """Fresh nn.Dropout keep factors for every dropout call of one training-mode forward (gpt.py:136,140,154,239-241)."""
    from oracle import mebt_oracle as O
    D, H, L = cfg["n_embd"], cfg["n_head"], cfg["sos_emb"]
    keep = lambda *shape: (torch.rand(*shape) >= p).float() / (1.0 - p)
    drop = {("stem", "lat"): keep(B, L, D), ("stem", "ctx"): keep(B, NC, D), ("stem", "tgt"): keep(B, NT, D)}
    for i, mode in enumerate(O.stack_modes(cfg)):
        nq = NT if mode == "latent_dec" else L
        nk = {"latent_enc": NC, "latent_self": L, "latent_dec": L, "lt2l": L + NT}[mode]
        drop[(i, "attn")], drop[(i, "proj")], drop[(i, "mlp")] = keep(B, H, nq, nk), keep(B, nq, D), keep(B, nq, D)
    return drop


def oracle_train_step(cfg, state, B, pdrop=0.0):
    from oracle import mebt_oracle as O
    P = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    opt = torch.optim.AdamW(list(P.values()), lr=1.08e-5, betas=(0.9, 0.95), weight_decay=0.01)
    x, indices = synth_batch(cfg, B, 1)
    N = int(np.prod(cfg["shape"]))

    def step():
        opt.zero_grad(set_to_none=True)
        drop = oracle_dropout_masks(cfg, B, N // 2, N // 2, pdrop) if pdrop > 0 else None
        r = O.shared_step(P, cfg, x, indices, TRAIN_T, "linear", drop=drop)
        r["loss"].backward()
        opt.step()
    desc = (f"full training step (fwd + CE + autograd bwd + AdamW), B={B}, t={TRAIN_T}, dropout {pdrop}, "
            "fp32 torch-CPU oracle port")
    return step, B * (N // 2), desc


def time_cpu(step, warm, reps):
    for _ in range(warm):
        step()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def run_reference(args, cfg):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_cpu_model(cfg)
    state = {k: v.detach() for k, v in model.state_dict().items()}
    if args.workload == "train16f":
        step, tokens, desc = oracle_train_step(cfg, state, args.batch or 6, args.dropout)
    elif args.workload == "vq16f":
        step, tokens, desc = oracle_vq_step(8)
    else:
        step, tokens, desc = oracle_sampling_step(cfg, state)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = tokens / dt
    print(json.dumps({
        "impl": "reference", "metric": "masked video tokens/sec", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "tokens": int(np.prod(cfg["shape"]))},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": "each step = " + desc},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---- native arm -------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="train16f", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0,
                    help="per GPU; default 6 for train16f (configs/stl/mebt_16f.yaml), 16 videos for sampling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="train16f: embd/resid/attn dropout (configs/stl/mebt_16f.yaml uses 0.1)")
    args = ap.parse_args()
    cfg = CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    training = args.workload == "train16f"
    B = args.batch or {"train16f": 6, "maskgit16f": 32, "vq16f": 64}.get(args.workload, 32)
    warmup = max(args.warmup, 3)

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
    cpu_model = build_cpu_model(cfg, pdrop=args.dropout if args.workload == "train16f" else 0.0)   # same weights on every rank (seed 0)
    state = {k: v.detach().clone() for k, v in cpu_model.state_dict().items()} if rank == 0 else None
    model = cpu_model.to(dev)
    N = int(np.prod(cfg["shape"]))
    random.seed(42)                                       # python RNG identical on every rank, as pl.seed_everything(42)

    if training:
        from mebt_b200.training import TrainState
        model.train()
        ts = TrainState(model, n_buckets=8 if world > 1 else 4)   # finer buckets shorten the exposed tail of the last all-reduce
        opt = ts.make_optimizer(lr=1.08e-5, weight_decay=0.01)
        x_cpu, idx_cpu = synth_batch(cfg, B, 100 + rank)  # each rank its own batch (DistributedSampler)
        x_host, idx_host = x_cpu.pin_memory(), idx_cpu.pin_memory()
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        x_dev, idx_dev = x_host.to(dev), idx_host.to(dev)
        tokens_per_step = B * (N // 2)

        def step_device():
            return ts.train_step(opt, x_dev, idx_dev, t=TRAIN_T, world_size=world)

        def replicas_in_sync():
            """Data-parallel invariant: after any number of steps every rank holds the same parameters (checked once,
            outside the timed region)."""
            try:
                import torch.distributed as dist
                chk = torch.stack([ts.flat.sum(dtype=torch.float64), ts.flat.abs().sum(dtype=torch.float64)])
                lo, hi = chk.clone(), chk.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                return bool(torch.equal(lo, hi))
            except Exception as exc:  # noqa: BLE001  (a diagnostic must never cost the bench line)
                return f"check failed: {type(exc).__name__}"

        def step_e2e():
            x = x_host.to(dev, non_blocking=True)
            idx = idx_host.to(dev, non_blocking=True)
            out = ts.train_step(opt, x, idx, t=TRAIN_T, world_size=world)
            loss_host.copy_(out["loss"].reshape(1), non_blocking=True)
        h2d, d2h = int(x_host.numel() * 8 + idx_host.numel() * 8), 4
    elif args.workload == "vq16f":
        # BASELINE.json configs[3]: codebook quantise + decode-side gather, 16x128x128 videos -> latents [256,4,16,16]
        from mebt_b200.vqgan import VQGAN
        torch.manual_seed(0)
        vq = VQGAN(16384, 256).to(dev).eval()
        z_host = torch.randn(B, 256, 4, 16, 16, generator=torch.Generator().manual_seed(4 + rank)).pin_memory()
        enc_host = torch.empty(B, 4, 16, 16, dtype=torch.long).pin_memory()
        z_dev = z_host.to(dev)
        tokens_per_step = B * 1024

        def step_device():
            return vq.decode(vq.encode(z_dev))

        def step_e2e():
            enc = vq.encode(z_host.to(dev, non_blocking=True))
            vq.decode(enc)
            enc_host.copy_(enc, non_blocking=True)
        h2d, d2h = int(z_host.numel() * 4), int(enc_host.numel() * 8)
    else:
        model.eval()
        model.rng_mode, model.rng_seed = "philox", 1000 + rank
        tokens_per_step = B * masked_tokens_per_video(N)
        x_host = torch.zeros(B, *cfg["shape"], dtype=torch.long).pin_memory()
        out_host = torch.empty(B, N, dtype=torch.long).pin_memory()
        x_dev = x_host.to(dev)
        torch.manual_seed(1234 + rank)                    # CPU generator: the randperm draws of the gibbs masks

        def step_device():
            return model.draft_and_revise(x_dev, None, **DNR)

        def step_e2e():
            x = x_host.to(dev, non_blocking=True)
            out_host.copy_(model.draft_and_revise(x, None, **DNR), non_blocking=True)
        h2d, d2h = int(x_host.numel() * 8), int(out_host.numel() * 8)
        if args.workload == "maskgit16f":
            # BASELINE.json configs[4]: the 24-layer model's maskgit loop (UCF recipe: 128 steps, cosine schedule,
            # context temperature 6) with the 16384-way logit head + confidence re-masking at batch 32
            model.mask_sampler.schedule = "cosine"
            tokens_per_step = B * maskgit_masked_tokens(N, MASKGIT["n_steps"])

            def step_device():                                                # noqa: F811
                return model.sample(x_dev, None, **MASKGIT)[0]

            def step_e2e():                                                   # noqa: F811
                x = x_host.to(dev, non_blocking=True)
                out_host.copy_(model.sample(x, None, **MASKGIT)[0], non_blocking=True)

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

    for _ in range(warmup):
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

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
        torch.cuda.synchronize()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    in_sync = replicas_in_sync() if (training and world > 1) else None

    # per-kernel-family timing of one more step (events on the launch stream, recorded by the library)
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
        tr = REPO / "profiles" / "r01_kernel_traffic.json"        # ncu --set full capture of one representative launch
        if tr.exists():
            t = json.loads(tr.read_text()).get("gemm_bf16_kernel")
            if t:
                roofline["traffic"] = t["dram_bytes"]
                roofline["traffic_note"] = (f"dram read+write of one launch at {t['shape']} (ncu, profiles/r01_ncu_kernels.md); "
                                            f"algorithmic bytes of that launch {t['algorithmic_bytes']}")
        for fam in ("sample", "ce", "layernorm"):
            f = prof[fam]
            if f["ms"] > 0:
                gbs = f["work"] / (f["ms"] * 1e-3) / 1e9
                roofline[f"{fam}_kernel_hbm"] = {"achieved_gbs": gbs, "peak_gbs": pk["hbm"], "frac": gbs / pk["hbm"]}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            if training:
                step, toks, desc = oracle_train_step(cfg, state, B, args.dropout)
                dt = time_cpu(step, 1, 2)
            elif args.workload == "vq16f":
                step, toks, desc = oracle_vq_step(8)
                dt = time_cpu(step, 1, 2)
            else:
                step, toks, desc = oracle_sampling_step(cfg, state)
                dt = time_cpu(step, 1, 2)
            cpu_baseline = {"value": toks / dt, "unit": "tokens/s", "cores": threads, "kind": "port",
                            "sample": desc + ", median of 2 after 1 warm-up"}
        config = {"workload": args.workload, "tokens": N, "batch_per_gpu": B,
                  "weights": "random init, reference distribution (337 M parameters)",
                  "l2": "working set (0.67 GB bf16 weights + activations/logits) exceeds the 126 MB L2; no flush needed"}
        if training:
            config.update(t=TRAIN_T, NC=N // 2, NT=N // 2, dropout=args.dropout, optimizer="AdamW fused fp32 master weights",
                          note="embd/attn/resid dropout as in configs/stl/mebt_16f.yaml; masks regenerated in backward",
                          grad_allreduce="fp32, 8 block buckets + head + embeddings, overlapped with backward" if world > 1 else "none (1 GPU)")
            if in_sync is not None:
                config["replicas_in_sync"] = in_sync
        elif args.workload == "vq16f":
            config = {"workload": "vq16f", "videos_per_gpu": B, "latent": [256, 4, 16, 16], "codebook": [16384, 256],
                      "unit_note": "a token = one quantised latent vector (fused fp32 distance+argmin, then both gathers)"}
        elif args.workload == "maskgit16f":
            config.update(sampler=MASKGIT, schedule="cosine", masked_tokens_per_video=maskgit_masked_tokens(N, 128),
                          noise="in-kernel philox", generated_tokens_per_s=world * B * N / (ms * 1e-3))
        else:
            config.update(sampler=DNR, masked_tokens_per_video=masked_tokens_per_video(N), noise="in-kernel philox (inverse CDF)",
                          logits="fp32 materialised", generated_tokens_per_s=world * B * N / (ms * 1e-3))
        print(json.dumps({
            "metric": "masked video tokens/sec", "value": world * tokens_per_step / (ms * 1e-3), "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": world * tokens_per_step / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clock_info}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
