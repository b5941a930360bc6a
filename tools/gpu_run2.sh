#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/r1c_pytest_gpu.log
tail -3 gpurun_out/r1c_pytest_gpu.log
timeout 600 python tools/gemm_probe.py --out gpurun_out/r1c_gemm_probe.json > gpurun_out/r1c_gemm_probe.log 2>&1; echo "probe rc=$?"
for v in "default:" "shared_sm:4=0"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 600 python tools/stage_timing.py --nb 96 --check > gpurun_out/r1c_stage_$name.log 2>&1; echo "stage $name rc=$?"
done
RCWA_B200_LIB=librcwa_b200_hb64.so timeout 600 python tools/stage_timing.py --nb 96 --check > gpurun_out/r1c_stage_hb64.log 2>&1; echo "stage hb64 rc=$?"
timeout 600 python tools/eig_profile.py --nb 96 --out gpurun_out/r1c_eig_profile.json > gpurun_out/r1c_eig_profile.log 2>&1; echo "eig profile rc=$?"
for nbv in 128 148; do
  timeout 600 python tools/stage_timing.py --nb $nbv > gpurun_out/r1c_stage_nb$nbv.log 2>&1; echo "stage nb$nbv rc=$?"
done
grep -h "layers/s\|parity\|eig(total)\|hessenberg(alone)" gpurun_out/r1c_stage_*.log
cat gpurun_out/r1c_eig_profile.log | head -30
