mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3g_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r3g_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3g_smoke.log
timeout 1200 python bench.py > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3g_bench_c3.json 2> gpurun_out/r3g_bench_c3.err; echo "bench config 3 rc=$?"
python - <<PY
import json
for f in ('r3g_bench', 'r3g_bench_c3'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ('metric', 'value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['config'].get('symmetry_reduction'))
        if d.get('roofline'): print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'ms_per_batch', 'ms_per_step')})
        if d.get('general_path'): print('   general', d['general_path']['value'], d['roofline']['general_path']['frac'])
        if d.get('vs_reference_cuda'): print('   ', d['vs_reference_cuda'], d['cpu_baseline'] and d['cpu_baseline'].get('value'))
        print('   ', d['kernel_time_share'])
    except Exception as e:
        print(f, 'unreadable', e); print(open('gpurun_out/%s.err' % f).read()[-2000:])
PY
