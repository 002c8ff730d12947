set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rfs 2>&1 > gpurun_out/r2_final2_tests_full.log
grep -E "^FAILED|^SKIPPED|passed|failed|^E  " gpurun_out/r2_final2_tests_full.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['modes']['exact']['value'])"
