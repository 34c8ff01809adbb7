#!/bin/bash
# full GPU suite + smoke + the bench line
mkdir -p gpurun_out
( time timeout 800 python -m pytest tests/ -q -m gpu -x ) > gpurun_out/t_full.log 2>&1; echo "rc=$?" >> gpurun_out/t_full.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/t_full.log | tail -8
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_cfg2.json ) 2> gpurun_out/bench_cfg2.err; tail -c 400 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_cfg2.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}); print(d['e2e']); print(d['parity']); print(d['train']); print(d['roofline']['ms_per_launch'], d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('bench parse failed', e)
PY
