# round 2, call AO: race detectors - training engine (fwd + bwd repeatability), sampler forward repeatability
set -x
timeout 600 python tools/train_stress.py 300 2>&1 | tail -6
