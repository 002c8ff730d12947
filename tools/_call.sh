set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 > gpurun_out/r2_s3_tests_full.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_s3_tests_full.log | head -40
grep -E "^E  " gpurun_out/r2_s3_tests_full.log | sort | uniq -c | sort -rn | head -20
