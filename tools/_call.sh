set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scorer_head.py -m gpu -q -x 2>&1 | tail -25
