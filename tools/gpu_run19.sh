#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1u_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 gpurun_out/r1u_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r1u_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1u_smoke.log
timeout 900 python bench.py > gpurun_out/r1u_bench.json 2> gpurun_out/r1u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r1u_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline_tensor']['achieved'], d['cpu_baseline']['value'])
except Exception as e:
    print('bench json unreadable', e)
PY
