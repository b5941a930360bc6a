#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_eig.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1g_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/r1g_pytest_gpu.log
tail -3 gpurun_out/r1g_pytest_gpu.log
for v in "default:" "g1:9=1" "g3:9=3" "g4:9=4" "us40:8=40" "us90:8=90" "shared:4=0" "g4shared:9=4,4=0"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 600 python tools/eig_profile.py --nb 128 --out gpurun_out/r1g_eig_profile_$name.json > gpurun_out/r1g_eig_profile_$name.log 2>&1; echo "eig profile $name rc=$?"
  grep "wall\|qr_pass dur" gpurun_out/r1g_eig_profile_$name.log
done
grep -A7 "QR pass segments" gpurun_out/r1g_eig_profile_default.log
