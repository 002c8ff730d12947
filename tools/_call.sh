set -u
export ADCENSUS_CBCA_LEAN=1
timeout 300 python -m pytest tests/test_gpu_cbca_tma.py -m gpu -q -x 2>&1 | tail -2
timeout 120 python tools/time_cbca.py 2>&1 | head -1
timeout 120 python tools/time_pipeline.py --batch --iters 24 2>&1 | tail -1
timeout 120 python tools/time_pipeline.py --iters 12 2>&1 | tail -1
unset ADCENSUS_CBCA_LEAN
timeout 120 python tools/time_pipeline.py --batch --iters 24 2>&1 | tail -1
timeout 120 python tools/time_pipeline.py --iters 12 2>&1 | tail -1
