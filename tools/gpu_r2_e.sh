# round 2, call E: state check after re-entry - GPU suite, default bench line, timeline
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/r02e_bench_default.json 2> gpurun_out/r02e_bench_default.err; tail -3 gpurun_out/r02e_bench_default.err
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2e.json 2>&1 | tail -2
