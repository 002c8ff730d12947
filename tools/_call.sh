set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2_c9_tests.log
timeout 120 python tools/time_cbca.py 2>&1 | head -1
ADCENSUS_CBCA_WB=6 timeout 120 python tools/time_cbca.py 2>&1 | head -1
ADCENSUS_CBCA_WB=8 timeout 120 python tools/time_cbca.py 2>&1 | head -1
timeout 120 python tools/time_pipeline.py 2>&1 | tail -1
timeout 120 python tools/time_pipeline.py --preset fast --D 70 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cbca_tma -c 1 -s 2 -f -o gpurun_out/r2_c9_cbca_tma python tools/time_cbca.py --iters 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_c9_launches.csv python tools/time_pipeline.py --iters 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_c9_launches.csv | tail -24
