mkdir -p gpurun_out
run() { RCWA_B200_TUNE="$1" timeout 400 python tools/sym_chunks.py --reps 1 --max-chunks 1 > gpurun_out/r3e_$2.log 2>&1; echo "tune [$1]"; grep -E "rep 0|eig \(|rror" gpurun_out/r3e_$2.log | head -2; }
run "" default_aed32
run "8=160" aed32_b160
run "8=240" aed32_b240
run "15=24" aed24
run "15=24,8=160" aed24_b160
run "15=28,8=160" aed28_b160
run "15=36,8=160" aed36_b160
for T in "15=48" "15=32" "15=40"; do RCWA_B200_TUNE="$T" timeout 400 python tools/sym_profile.py --general --points 64 > gpurun_out/r3e_general.log 2>&1; echo "general path, tune [$T]"; grep "step wall" gpurun_out/r3e_general.log; done
