set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err2.log > gpurun_out/r01_bench_train16f_2gpu.json; tail -2 gpurun_out/bench_err2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload sample128f --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err2.log > gpurun_out/r01_bench_sample128f_2gpu.json; tail -2 gpurun_out/bench_err2.log
cat gpurun_out/r01_bench_train16f_2gpu.json | cut -c1-200; cat gpurun_out/r01_bench_sample128f_2gpu.json | cut -c1-200
