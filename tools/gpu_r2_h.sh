# round 2, call H (2 GPUs): sharded data-parallel exchange vs all-reduce; 2-GPU default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -v "^W\|warn" | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err; tail -2 gpurun_out/r02h_bench_2gpu.err
timeout 300 python -m pytest tests/test_training_gpu.py -m gpu -x -q 2>&1 | tail -3
