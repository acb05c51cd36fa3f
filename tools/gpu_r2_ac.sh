# round 2, call AC: DRAM traffic of the dominant kernel family inside the bench commands (roofline.traffic)
set -x
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:gemm -c 560 --csv --log-file gpurun_out/r02_traffic_train16f.csv python bench.py --workload train16f --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_tr_train.log 2>&1
tail -1 gpurun_out/ncu_tr_train.log | cut -c1-300
timeout 600 ncu --metrics $M --clock-control none -k regex:conv3d -c 100 --csv --log-file gpurun_out/r02_traffic_vqgan16f.csv python bench.py --workload vqgan16f --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tr_vqgan.log 2>&1
tail -1 gpurun_out/ncu_tr_vqgan.log | cut -c1-300
timeout 900 ncu --metrics $M --clock-control none -k regex:gemm -c 5400 --csv --log-file gpurun_out/r02_traffic_sample128f.csv python bench.py --workload sample128f --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tr_sample.log 2>&1
tail -1 gpurun_out/ncu_tr_sample.log | cut -c1-300
ls -la gpurun_out/r02_traffic_*.csv
