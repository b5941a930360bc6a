mkdir -p gpurun_out
timeout 1200 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3a_bench_c3.json 2> gpurun_out/r3a_bench_c3.err; echo "bench config 3 rc=$?"
timeout 1200 python bench.py > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"
python - <<PY
import json
for f in ('r3a_bench_c3', 'r3a_bench'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ('metric', 'value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['config'].get('symmetry_reduction'))
        if 'roofline' in d and d['roofline']: print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'ms_per_batch')})
        if d.get('vs_reference_cuda'): print('   ', d['vs_reference_cuda'], d['cpu_baseline'] and d['cpu_baseline'].get('value'))
    except Exception as e:
        print(f, 'unreadable', e); print(open('gpurun_out/%s.err' % f).read()[-2000:])
PY
