set -x
timeout 900 python tools/sample_stress.py 6 2>&1 | tail -8
