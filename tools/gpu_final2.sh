set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_train16f.json; tail -2 gpurun_out/bench_err.log
timeout 900 python bench.py --workload sample128f --steps 2 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_sample128f.json; tail -2 gpurun_out/bench_err.log
timeout 600 python bench.py --workload maskgit16f --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_maskgit16f.json; tail -2 gpurun_out/bench_err.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 15400 -c 1200 --csv --log-file gpurun_out/launches_sample128f.csv python bench.py --workload sample128f --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -c 300 gpurun_out/ncu_list.log
