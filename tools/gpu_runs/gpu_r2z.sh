mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 1500 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'general', d['general_path'] and d['general_path']['value'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'ms_per_batch')}, d['roofline']['kernel'][:160])
    print('general', d['roofline']['general_path'])
    print('cpu', d['cpu_baseline'] and d['cpu_baseline'].get('value'), 'cuda', d['cuda_baseline'], d['vs_reference_cuda'])
    print(d['kernel_time_share'])
except Exception as e:
    print('bench json unreadable', e); print(open('gpurun_out/r2z_bench.err').read()[-2500:])
PY
