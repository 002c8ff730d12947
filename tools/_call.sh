set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2_c6_tests.log
timeout 120 python tools/time_pipeline.py 2>&1 | tail -1
timeout 120 python tools/time_pipeline.py --overlap 0 2>&1 | tail -1
timeout 120 python tools/time_pipeline.py --exact 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cbca_tma -c 1 -s 4 -f -o gpurun_out/r2_c6_cbca_tma python tools/time_cbca.py --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
