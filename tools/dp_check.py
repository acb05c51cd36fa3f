"""Data-parallel exchange check (run under torchrun on >= 2 GPUs): K steps with the sharded exchange (reduce-scatter ->
AdamW on the 1/N shard -> bf16 all-gather) against K steps with the all-reduce + replicated AdamW, same seeds, same
per-rank batches: the fp32 masters must agree (NCCL's reduction order is the only difference), every rank must hold the
same parameters, and the steady-state step times of both forms are printed.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench  # noqa: E402
from mebt_b200.training import TrainState  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    cfg = bench.CONFIGS["train16f"]
    results = {}
    for mode in ("sharded", "allreduce"):
        model = bench.build_native_model(cfg, bench.synth_weights(cfg), 0.1, dev).train()
        ts = TrainState(model, n_buckets=8)
        ts.dropout_seed = 1234
        opt = ts.make_optimizer()
        w0 = ts.flat.clone()
        x, idx = bench.synth_batch(cfg, 6, 100 + rank)
        x, idx = x.to(dev), idx.to(dev)
        losses = []
        for _ in range(3):
            out = ts.train_step(opt, x, idx, t=0.5, world_size=world, sharded=(mode == "sharded"))
            losses.append(float(out["loss"]))
        ts.sync_masters()
        torch.cuda.synchronize()
        results[mode] = (ts.flat.clone() - w0, losses)
        # every rank holds the same masters
        chk = torch.stack([ts.flat.sum(dtype=torch.float64), ts.flat.abs().sum(dtype=torch.float64)])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
        # steady-state timing
        for _ in range(3):
            ts.train_step(opt, x, idx, t=0.5, world_size=world, sharded=(mode == "sharded"))
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ts.train_step(opt, x, idx, t=0.5, world_size=world, sharded=(mode == "sharded"))
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rep = ts.comm_report()
        if rank == 0:
            print(f"{mode:10s}: losses {['%.5f' % v for v in losses]} replicas_in_sync {in_sync} | {float(ms):.3f} ms/step, "
                  f"exposed exchange {rep['exposed_ms']:.3f} ms ({rep['exchange']})", flush=True)
        del ts, opt, model
        torch.cuda.empty_cache()
    # AdamW's first steps move every weight by ~lr * sign(g): the two exchanges differ only in fp32 summation order (NCCL
    # ring position; the embedding scatter-add uses atomics), which can flip the sign of a near-zero averaged gradient,
    # so single elements may differ by up to 2 * lr per step while the update as a whole must agree
    a, b = results["sharded"][0].double(), results["allreduce"][0].double()
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    worst = float((a - b).abs().max())
    frac = float(((a - b).abs() > 1e-7).double().mean())
    if rank == 0:
        print(f"updates after 3 steps: cosine(sharded, allreduce) = {cos:.6f}, max element difference {worst:.3e} "
              f"(bound 3 steps x 2 lr = {6 * 1.08e-5:.3e}), elements differing by > 1e-7: {100 * frac:.3f} %", flush=True)
        assert cos > 0.999 and worst <= 6 * 1.08e-5 * 1.01, (cos, worst)
        print("DP CHECK OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
