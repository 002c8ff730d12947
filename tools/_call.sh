set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch_lanes.py tests/test_gpu_parity.py -m gpu -q -x -k "not middlebury" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2_s6_bench_n1.json 2> gpurun_out/r2_s6_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s6_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['modes']['exact']['value'], d['modes']['fast']['disp_pixels_off_by_more_than_1e-4_vs_exact_frac'])
PY
tail -3 gpurun_out/r2_s6_bench_n1.err
ADCENSUS_LANES=1 timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('one lane:', d['value'], d['e2e']['value'])"
