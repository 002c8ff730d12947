set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scorer_head.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/time_scorer_head.py --iters 2 2>&1 | tail -3
timeout 300 python tools/time_scorer_head.py --iters 2 --l2 3 2>&1 | tail -3
