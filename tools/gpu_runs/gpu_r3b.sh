mkdir -p gpurun_out
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r3b_launches.csv python bench.py --steps 1 --warmup 0 --points 96 --no-cpu-baseline --no-cuda-baseline --no-general-path > gpurun_out/r3b_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python - <<PY
import csv, collections
per = collections.defaultdict(lambda: [0, 0.0])
rows = 0
with open('gpurun_out/r3b_launches.csv') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum': continue
    nm = r['Kernel Name'].split('(')[0]
    nm = nm.replace('void ', '').replace('rcwa::', '').replace('(anonymous namespace)::', '')
    if 'zgemm_grouped_kernel' in nm: nm = 'zgemm_grouped_kernel' + nm[nm.index('<'):][:24]
    else: nm = nm.split('<')[0]
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    per[nm][0] += 1; per[nm][1] += v; rows += 1
tot = sum(v[1] for v in per.values())
with open('gpurun_out/r3b_launches_summary.csv', 'w') as f:
    f.write('kernel,launches,total_us,avg_us,share\n')
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        f.write('%s,%d,%.1f,%.2f,%.4f\n' % (k, v[0], v[1], v[1] / v[0], v[1] / tot))
print(rows, 'launches'); print(open('gpurun_out/r3b_launches_summary.csv').read()[:1500])
PY
rm -f gpurun_out/r3b_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_pass_kernel -s 3001 -c 2 -o gpurun_out/r3b_prof_qr_pass python tools/sym_chunks.py --reps 0 --points 128 --max-chunks 0 > gpurun_out/r3b_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r3b_prof_qr_pass.ncu-rep --page raw --csv > gpurun_out/r3b_prof_qr_pass_raw.csv 2>/dev/null; ls -la gpurun_out/r3b_prof_qr_pass*
