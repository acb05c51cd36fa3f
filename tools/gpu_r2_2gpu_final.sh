# round 2, final 2-GPU check: the driver's scaling invocation on the final tree + the exchange comparison
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -v "^W\|warn" | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_default_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -2 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_default_2gpu.json').read().strip().splitlines()[-1])
print('2gpu', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j.get('detail'))
w=j['workloads']['sample128f']; print('  sample128f', round(w['value']), round(w['ms_per_step'],2))
PY
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | tail -c 300
