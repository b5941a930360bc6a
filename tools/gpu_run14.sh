#!/bin/bash
# final validation of the round: full GPU suite, smoke, default bench, order-21 (BASELINE config 3 size) check
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r1p_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 gpurun_out/r1p_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r1p_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1p_smoke.log
timeout 900 python bench.py > gpurun_out/r1p_bench.json 2> gpurun_out/r1p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r1p_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'clocks')}, d['e2e'], {k: d['roofline'][k] for k in ('achieved', 'frac', 'us_per_launch', 'traffic')}, d['roofline_tensor']['achieved'], d['cpu_baseline']['value'])
    print(d['stage_ms_per_batch'])
except Exception as e:
    print('bench json unreadable', e)
PY
timeout 600 python tools/stage_timing.py --order 21 --nb 4 --residual > gpurun_out/r1p_stage_order21.log 2>&1; echo "order21 rc=$?"; tail -14 gpurun_out/r1p_stage_order21.log
