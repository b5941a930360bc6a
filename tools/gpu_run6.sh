#!/bin/bash
# full validation + bench + ncu evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1h_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/r1h_pytest_gpu.log
tail -2 gpurun_out/r1h_pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r1h_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r1h_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, d['e2e'], {k: d['roofline'][k] for k in ('achieved', 'frac', 'us_per_launch')}, d['roofline_tensor']['achieved'], d['cpu_baseline'])
    print(d['stage_ms_per_batch'])
except Exception as e:
    print('bench json unreadable', e)
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1h_bench_reference.json 2> gpurun_out/r1h_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r1h_bench_reference.json | cut -c1-400
# ncu: full captures of the named kernels
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:hb_matvec --launch-skip 300 --launch-count 2 -o gpurun_out/r1h_prof_hb_matvec -f python tools/ncu_target.py hess --nb 32 > gpurun_out/r1h_ncu_hess.log 2>&1; echo "ncu hess rc=$?"
timeout 600 $NCU -k regex:zgemm_grouped --launch-count 2 -o gpurun_out/r1h_prof_zgemm -f python tools/ncu_target.py gemm --nb 8 > gpurun_out/r1h_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 $NCU -k regex:qr_pass --launch-skip 400 --launch-count 2 -o gpurun_out/r1h_prof_qr_pass -f python tools/ncu_target.py qr --nb 8 > gpurun_out/r1h_ncu_qr.log 2>&1; echo "ncu qr rc=$?"
for f in hb_matvec zgemm qr_pass; do ncu -i gpurun_out/r1h_prof_$f.ncu-rep --page raw --csv > gpurun_out/r1h_prof_${f}_raw.csv 2>/dev/null; done
# ncu: launch list of one (small-batch) bench step: per-launch times are cold-cache and serialised; shares matter
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1h_launches.csv python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline > gpurun_out/r1h_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | grep r1h
