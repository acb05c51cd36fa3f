set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_default_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -2 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_default_2gpu.json').read().strip().splitlines()[-1])
print('2gpu', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j['detail'].get('replicas_in_sync'))
for n,w in j['workloads'].items(): print(' ', n, round(w.get('value',0)), round(w.get('ms_per_step',0),2), w.get('error'))
PY
