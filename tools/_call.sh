set -u
timeout 300 python -m pytest tests/test_gpu_feature_tower.py -m gpu -q -x 2>&1 | tail -4
timeout 120 python tools/time_feature_tower.py 2>&1 | tail -1
ADCENSUS_TOWER_RESIDENT=0 timeout 120 python tools/time_feature_tower.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none -k regex:conv3x3_resident -c 1 -f -o gpurun_out/r2_final_feature_tower_v2 python tools/time_feature_tower.py > /dev/null 2>&1
ls -la gpurun_out/r2_final_feature_tower_v2.ncu-rep
