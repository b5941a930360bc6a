#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r1i_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r1i_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, d['e2e'], {k: d['roofline'][k] for k in ('achieved', 'frac', 'us_per_launch', 'traffic', 'bytes_per_launch')}, d['roofline']['in_step_cupti'], d['roofline_tensor']['achieved'], d['cpu_baseline'])
    print(d['stage_ms_per_batch']); print(d['kernel_time_share'])
except Exception as e:
    print('bench json unreadable', e)
PY
timeout 600 python tools/ref_cuda_timing.py --points 4 > gpurun_out/r1i_ref_cuda.json 2> gpurun_out/r1i_ref_cuda.err; echo "ref cuda rc=$?"; cat gpurun_out/r1i_ref_cuda.json; tail -c 300 gpurun_out/r1i_ref_cuda.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 22000 -c 2500 --csv --log-file gpurun_out/r1i_launches_qr_window.csv python bench.py --steps 1 --warmup 0 --points 8 --no-cpu-baseline > gpurun_out/r1i_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
