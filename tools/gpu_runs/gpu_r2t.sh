mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "sym_project or wrappers" > gpurun_out/r2t_pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 gpurun_out/r2t_pytest_kernels.log
timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s > gpurun_out/r2t_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; grep -E "passed|failed|Error" gpurun_out/r2t_pytest_parity.log | tail -4; grep -E "sym=True" gpurun_out/r2t_pytest_parity.log | tail -24
timeout 600 python tools/sym_profile.py --points 128 > gpurun_out/r2t_sym_profile.log 2>&1; grep -v Warn gpurun_out/r2t_sym_profile.log | tail -30
timeout 900 python bench.py --steps 2 --warmup 2 --points 128 --no-cpu-baseline --no-cuda-baseline --general-points 64 > gpurun_out/r2t_bench_p128.json 2> gpurun_out/r2t_bench_p128.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2t_bench_p128.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'general', d['general_path'], d['config']['symmetry_reduction'])
except Exception as e:
    print('bench json unreadable', e); print(open('gpurun_out/r2t_bench_p128.err').read()[-2500:])
PY
