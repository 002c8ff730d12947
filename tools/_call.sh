set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stereo_join_pitched.py -m gpu -q -x -k "not middlebury" 2>&1 | tail -6
timeout 120 python tools/time_sj.py 2>&1 | tail -2
ADCENSUS_SJ_TMA=0 timeout 120 python tools/time_sj.py 2>&1 | tail -1
timeout 120 python tools/time_pipeline.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stereo_join_tma -c 1 -s 1 -f -o gpurun_out/r2_sj_tma_v1 python tools/time_sj.py --iters 1 > /dev/null 2>&1
ls -la gpurun_out/r2_sj_tma_v1.ncu-rep
