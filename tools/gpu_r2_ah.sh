set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none -k 'regex:conv3d' -o /tmp/prof_cl -f python tools/ncu_conv_last.py > gpurun_out/ncu_cl.log 2>&1
tail -3 gpurun_out/ncu_cl.log
ncu -i /tmp/prof_cl.ncu-rep --page raw --csv > gpurun_out/r02_prof_conv_last_raw.csv 2>/dev/null
ncu -i /tmp/prof_cl.ncu-rep --page details --csv > gpurun_out/r02_prof_conv_last_details.csv 2>/dev/null
ls -la gpurun_out/r02_prof_conv_last*
