#!/bin/bash
mkdir -p gpurun_out
for v in "default:" "default2:" "g4:9=4" "g1:9=1"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 300 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1l_stage_$name.log 2>&1; echo "stage $name rc=$?"
  grep -h "parity\|eig(total)\|layers/s" gpurun_out/r1l_stage_$name.log
done
RCWA_B200_LIB=librcwa_b200_hb64.so timeout 300 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1l_stage_hb64.log 2>&1; echo "hb64 rc=$?"
grep -h "parity\|eig(total)\|layers/s\|hessenberg(alone)" gpurun_out/r1l_stage_hb64.log
