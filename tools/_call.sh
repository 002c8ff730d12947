set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cbca_tma.py tests/test_gpu_batch_lanes.py -m gpu -q -x 2>&1 | tail -3
timeout 120 python tools/time_cbca.py 2>&1 | head -1
timeout 200 python tools/cbca_accuracy.py 2>&1 | tail -1
timeout 120 python tools/time_pipeline.py --batch --iters 24 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size or middlebury" 2>&1 | tail -3
