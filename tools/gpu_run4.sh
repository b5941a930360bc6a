#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_eig.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1f_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/r1f_pytest_gpu.log
tail -3 gpurun_out/r1f_pytest_gpu.log
for v in "default:" "b80:5=80,6=50,7=6" "b320:5=320,6=200,7=24" "shared:4=0"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 600 python tools/eig_profile.py --nb 96 --out gpurun_out/r1f_eig_profile_$name.json > gpurun_out/r1f_eig_profile_$name.log 2>&1; echo "eig profile $name rc=$?"
  grep "wall\|qr_pass dur" gpurun_out/r1f_eig_profile_$name.log
done
grep -A7 "QR pass segments" gpurun_out/r1f_eig_profile_default.log
timeout 600 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1f_stage_nb128.log 2>&1; echo "stage rc=$?"
grep -h "layers/s\|parity\|eig(total)\|hessenberg(alone)\|stats" gpurun_out/r1f_stage_*.log
