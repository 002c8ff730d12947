set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n4.json 2> gpurun_out/r2_final_bench_n4.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
print('rowband', {k:v for k,v in d.get('rowband',{}).items() if k!='split'})
PY
