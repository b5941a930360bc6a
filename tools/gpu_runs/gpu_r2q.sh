mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'steps', 'warmup')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
    print('tc', {k: (v.get('ms'), v.get('speedup_vs_dmma'), v.get('frac_of_int8_peak')) for k, v in d['roofline_tensor']['tcgen05'].items() if k.startswith('digits')})
    print('cuda', d['vs_reference_cuda'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
except Exception as e:
    print('bench json unreadable', e); print(open('gpurun_out/r2q_bench.err').read()[-2500:])
PY
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline --no-cuda-baseline > gpurun_out/r2q_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2q_launches.csv')) if len(r) > 10 and r[0].isdigit()]
per = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    nm = r[4].replace('<unnamed>::', '').replace('void ', '').split('(')[0].split('<')[0]
    per[nm][0] += 1; per[nm][1] += float(r[-1])
tot = sum(v[1] for v in per.values())
with open('gpurun_out/r2q_launches_summary.csv', 'w') as f:
    f.write('kernel,launches,total_us,share\n')
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        f.write('%s,%d,%.1f,%.4f\n' % (k, v[0], v[1] / 1e3, v[1] / tot))
print(open('gpurun_out/r2q_launches_summary.csv').read()[:1500])
PY
