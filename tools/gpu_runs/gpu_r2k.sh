mkdir -p gpurun_out
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'])
    print('roofline', d['roofline']['frac'], d['roofline']['streaming_kernel_alone']['frac'], d['roofline']['hessenberg_phase_whole']['frac'])
    print('tensor', {k: d['roofline_tensor'][k] for k in ('achieved', 'peak', 'frac')}, d['roofline_tensor']['tcgen05'])
    print('cpu', d['cpu_baseline']); print('cuda', d['cuda_baseline']); print(d['vs_reference_cuda'])
    print({k: round(v) for k, v in d['stage_ms_per_batch'].items() if k.endswith('_ms')}); print(d['kernel_time_share'])
except Exception as e:
    print('bench json unreadable', e); print(open('gpurun_out/r2k_bench.err').read()[-2500:])
PY
timeout 900 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err; echo "bench c3 rc=$?"; cut -c1-700 gpurun_out/r2k_bench_c3.json; tail -3 gpurun_out/r2k_bench_c3.err
timeout 900 python bench.py --config 5 --order 15 --order-y 8 --points 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_bench_c5a.json 2> gpurun_out/r2k_bench_c5a.err; echo "bench c5 [15,8] rc=$?"; cut -c1-700 gpurun_out/r2k_bench_c5a.json; tail -3 gpurun_out/r2k_bench_c5a.err
timeout 1200 python bench.py --config 5 --points 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_bench_c5b.json 2> gpurun_out/r2k_bench_c5b.err; echo "bench c5 [25,25] rc=$?"; cut -c1-700 gpurun_out/r2k_bench_c5b.json; tail -3 gpurun_out/r2k_bench_c5b.err
