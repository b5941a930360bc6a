mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s > gpurun_out/r2r_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; grep -E "passed|failed|Error" gpurun_out/r2r_pytest_parity.log | tail -4; grep -E "sym=True" gpurun_out/r2r_pytest_parity.log | tail -24
for P in 128 256; do timeout 900 python bench.py --steps 2 --warmup 2 --points $P --no-cpu-baseline --no-cuda-baseline --general-points 64 > gpurun_out/r2r_bench_p$P.json 2> gpurun_out/r2r_bench_p$P.err; echo "bench P=$P rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2r_bench_p$P.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'general', d['general_path'], d['config']['symmetry_reduction'])
    print(d['kernel_time_share'])
except Exception as e:
    print('bench json unreadable', e); print(open('gpurun_out/r2r_bench_p$P.err').read()[-2500:])
PY
done
