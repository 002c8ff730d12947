set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rfs 2>&1 > gpurun_out/r2_final3_tests_full.log
grep -E "^FAILED|^SKIPPED|passed|failed|^E  " gpurun_out/r2_final3_tests_full.log | head -12
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_sgm_dhw.py tests/test_gpu_parity.py -m gpu -q -x -k "sgm or default_mode" 2>&1 | tail -1; done
