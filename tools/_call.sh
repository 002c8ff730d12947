set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sgm_dhw.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r2_c3_sgm_dhw.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_cbca_tma.py -m gpu -q -x -k "70-300-36-5-0.13--1" > gpurun_out/r2_c3_sanitizer.log 2>&1
grep -E "=========" gpurun_out/r2_c3_sanitizer.log | head -40
