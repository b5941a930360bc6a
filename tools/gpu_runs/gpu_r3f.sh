mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig.py -m gpu -x -q > gpurun_out/r3f_pytest_eig.log 2>&1; echo "pytest eig rc=$?"; tail -2 gpurun_out/r3f_pytest_eig.log
run() { RCWA_B200_TUNE="$1" timeout 400 python tools/sym_chunks.py --reps 1 --max-chunks 1 > gpurun_out/r3f_$2.log 2>&1; echo "tune [$1]"; grep -E "rep 0|eig \(|rror" gpurun_out/r3f_$2.log | head -2; }
run "" default_aed24
run "15=20" aed20
run "15=16" aed16
for T in "15=24" "15=28"; do RCWA_B200_TUNE="$T" timeout 400 python tools/sym_profile.py --general --points 64 > gpurun_out/r3f_general.log 2>&1; echo "general path, tune [$T]"; grep "step wall" gpurun_out/r3f_general.log; done
