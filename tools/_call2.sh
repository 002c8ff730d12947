set -u
mkdir -p gpurun_out
timeout 120 python tools/time_feature_tower.py 2>&1 | tail -1
ADCENSUS_TOWER_RESIDENT=0 timeout 120 python tools/time_feature_tower.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none -k regex:conv3x3_resident -c 1 -f -o gpurun_out/r2_final_feature_tower_v2 python tools/time_feature_tower.py > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
print('rowband', {k:v for k,v in d.get('rowband',{}).items() if k!='split'})
PY
