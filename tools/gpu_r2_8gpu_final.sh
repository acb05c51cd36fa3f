# round 2, final 8-GPU record: the driver's scaling invocation on the final tree
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_default_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -2 gpurun_out/r02_bench_8gpu.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_default_8gpu.json').read().strip().splitlines()[-1])
print('8gpu', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j['detail'].get('replicas_in_sync'), j['detail'].get('comm'))
w=j['workloads']['sample128f']; print('  sample128f', round(w['value']), round(w['ms_per_step'],2))
PY
