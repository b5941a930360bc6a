#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "energy" -s > gpurun_out/r1q_pytest_energy.log 2>&1; echo "pytest rc=$?"; grep "R + T\|passed\|failed" gpurun_out/r1q_pytest_energy.log
for P in 160 192; do
  timeout 600 python bench.py --points $P --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1q_bench_p$P.json 2> gpurun_out/r1q_bench_p$P.err; echo "bench $P rc=$?"
  tail -c 300 gpurun_out/r1q_bench_p$P.err | tail -2
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r1q_bench_p$P.json').read().strip().splitlines()[-1])
    print($P, d['value'], d['ms_per_step'], d['stage_ms_per_batch']['eig_ms'])
except Exception as e:
    print('unreadable', e)
PY
done
