# round 2, call A: state of the tree at round start — GPU tests, phase timing, device timeline, default bench line
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -3
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2a.json 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r02a_bench_default.json; tail -3 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>gpurun_out/bench_err.log > gpurun_out/r02a_bench_reference.json; tail -3 gpurun_out/bench_err.log
cat gpurun_out/r02a_bench_default.json | head -c 1500
