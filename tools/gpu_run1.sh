#!/bin/bash
# round-1 session-2 GPU call #1: kernel tests, GEMM configuration table, stage timings per tuning variant
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r1b_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r1b_pytest_kernels.log 2>&1; echo "pytest kernels rc=$?" >> gpurun_out/r1b_pytest_kernels.log
tail -3 gpurun_out/r1b_pytest_kernels.log
timeout 600 python tools/gemm_probe.py --out gpurun_out/r1b_gemm_probe.json > gpurun_out/r1b_gemm_probe.log 2>&1; echo "probe rc=$?"
for v in "default:" "nom3:0=0" "old:0=0,3=0,1=0,2=1" "m3_bigqr:1=0,2=1"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 600 python tools/stage_timing.py --nb 96 --check > gpurun_out/r1b_stage_$name.log 2>&1; echo "stage $name rc=$?"
done
RCWA_B200_LIB=librcwa_b200_hb64.so timeout 600 python tools/stage_timing.py --nb 96 --check > gpurun_out/r1b_stage_hb64.log 2>&1; echo "stage hb64 rc=$?"
grep -h "layers/s\|parity\|eig(total)\|hessenberg" gpurun_out/r1b_stage_*.log
