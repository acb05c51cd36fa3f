set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_training_gpu.py -m gpu -x -q -k overlapped 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err2.log > gpurun_out/r01_bench_train16f_2gpu.json; tail -2 gpurun_out/bench_err2.log
python -c "import json; j=json.load(open('gpurun_out/r01_bench_train16f_2gpu.json')); print(j['value'], j['ms_per_step'], j['config'].get('replicas_in_sync'))"
