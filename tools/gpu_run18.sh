#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1t_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 gpurun_out/r1t_pytest_gpu.log
for v in "us90:" "us130:8=130" "us180:8=180" "us130g3:8=130,9=3"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 300 python tools/eig_profile.py --nb 128 > gpurun_out/r1t_eig_profile_$name.log 2>&1; echo "eig profile $name rc=$?"
  grep "wall\|qr_pass dur" gpurun_out/r1t_eig_profile_$name.log
done
