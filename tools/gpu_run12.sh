#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1n_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r1n_pytest_gpu.log
for v in "band:" "noband:10=0"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 300 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1n_stage_$name.log 2>&1; echo "stage $name rc=$?"
  grep -h "parity\|eig(total)\|layers/s" gpurun_out/r1n_stage_$name.log
done
timeout 300 python tools/eig_profile.py --nb 128 --out gpurun_out/r1n_eig_profile.json > gpurun_out/r1n_eig_profile.log 2>&1; echo "profile rc=$?"
grep "wall\|zgemm 64x64\|qr_pass" gpurun_out/r1n_eig_profile.log | head
