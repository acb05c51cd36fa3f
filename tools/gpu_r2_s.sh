# round 2, call S: split-K for the K = 4096 latent MLP GEMMs (256-wide tiles x 3 splits), D = 1024 LayerNorm-backward kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_backward_gpu.py tests/test_dropout_gpu.py tests/test_training_gpu.py tests/test_stl_config_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 200 python tools/gemm_probe.py train 2>&1 | cut -c1-200 | tail -12
timeout 300 python tools/train_sweep.py --base 2>&1 | tail -2
