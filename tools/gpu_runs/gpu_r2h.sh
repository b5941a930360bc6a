mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "order15_batched or batched_equals or c64_api" > gpurun_out/r2h_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/r2h_pytest_parity.log | tail -5
timeout 1500 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tc_vs_dmma" > gpurun_out/r2h_pytest_stage.log 2>&1; echo "pytest stage rc=$?"; tail -3 gpurun_out/r2h_pytest_stage.log
for pl in 1 2 3; do RCWA_B200_PIPELINE=$pl timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_bench_p$pl.json 2> gpurun_out/r2h_bench_p$pl.err; echo "bench pipeline=$pl rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2h_bench_p$pl.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], {k: round(v) for k, v in d['stage_ms_per_batch'].items() if k.endswith('_ms')})
except Exception as e:
    print('bench json unreadable', e); print(open('gpurun_out/r2h_bench_p$pl.err').read()[-1500:])
PY
done
