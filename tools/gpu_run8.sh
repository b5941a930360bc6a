#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r1j_gpus.txt
timeout 900 python -m pytest tests/test_gpu_eig.py -m gpu -x -q -k "backward or autograd" > gpurun_out/r1j_pytest_backward.log 2>&1; echo "pytest backward rc=$?"; tail -15 gpurun_out/r1j_pytest_backward.log
RCWA_B200_LIB=librcwa_b200_hb64.so timeout 600 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1j_stage_hb64.log 2>&1; echo "hb64 rc=$?"
timeout 600 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1j_stage_hb32.log 2>&1; echo "hb32 rc=$?"
grep -h "layers/s\|parity\|eig(total)\|hessenberg(alone)" gpurun_out/r1j_stage_hb*.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --points 48 --no-cpu-baseline > gpurun_out/r1j_bench_2gpu.json 2> gpurun_out/r1j_bench_2gpu.err; echo "2gpu rc=$?"
tail -c 300 gpurun_out/r1j_bench_2gpu.err; cut -c1-600 gpurun_out/r1j_bench_2gpu.json
