# round 2, call L (8 GPUs): the driver's scaling invocation, default bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_default_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; grep -v "^W1\|warn\|\*\*\*\|OMP" gpurun_out/r02_bench_8gpu.err | tail -5
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_default_8gpu.json').read().strip().splitlines()[-1])
print('train16f x8', j['value'], j['ms_per_step'], j['detail'])
w=j['workloads']['sample128f']; print('sample128f x8', w['value'], w['ms_per_step'])
PY
