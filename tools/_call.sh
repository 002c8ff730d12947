set -u
for l in 1 2 3 4; do
  ADCENSUS_LANES=$l timeout 120 python tools/time_pipeline.py --batch --iters 24 2>&1 | tail -1
done
timeout 300 python -m pytest tests/test_gpu_batch_lanes.py -m gpu -q 2>&1 | tail -2
ADCENSUS_LANES=3 timeout 300 python -m pytest tests/test_gpu_batch_lanes.py -m gpu -q 2>&1 | tail -2
