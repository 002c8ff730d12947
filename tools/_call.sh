set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cbca_tma.py tests/test_gpu_stereo_join_pitched.py -m gpu -q 2>&1 | tail -6
timeout 120 python tools/time_cbca.py 2>&1 | tail -2
ADCENSUS_CBCA_DCH=12 timeout 120 python tools/time_cbca.py 2>&1 | head -1
ADCENSUS_CBCA_DCH=16 timeout 120 python tools/time_cbca.py 2>&1 | head -1
timeout 120 python tools/time_pipeline.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cbca_tma -c 1 -s 2 -f -o gpurun_out/r2_c8_cbca_tma python tools/time_cbca.py --iters 1 > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c8_bench.json 2> gpurun_out/r2_c8_bench.err; tail -3 gpurun_out/r2_c8_bench.err; cut -c1-600 gpurun_out/r2_c8_bench.json
