set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_s4_bench_n2.json 2> gpurun_out/r2_s4_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s4_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d.get('e2e'))
print('rowband', d.get('rowband'))
print('scorer_head', d.get('scorer_head'))
PY
tail -5 gpurun_out/r2_s4_bench_n2.err
