#!/bin/bash
mkdir -p gpurun_out
for v in "default:" "g1:9=1" "count:8=-1" "g1count:9=1,8=-1" "default2:"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 300 python tools/stage_timing.py --nb 64 --check > gpurun_out/r1k_stage_$name.log 2>&1; echo "stage $name rc=$?"
  grep -h "parity\|eig(total)" gpurun_out/r1k_stage_$name.log
done
