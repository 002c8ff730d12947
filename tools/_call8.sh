set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_s7_bench_n8.json 2> gpurun_out/r2_s7_bench_n8.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_s7_bench_n8.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d.get('e2e',{}).get('value'))
    print('rowband', d.get('rowband'))
except Exception as e:
    print('no json', e)
PY
grep -v "Warning\|warn" gpurun_out/r2_s7_bench_n8.err | tail -5
