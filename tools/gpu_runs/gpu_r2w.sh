mkdir -p gpurun_out
RCWA_B200_TUNE="13=2" timeout 1500 python -m pytest tests/test_gpu_eig.py -m gpu -x -q > gpurun_out/r2w_pytest_eig_split.log 2>&1; echo "pytest eig (split forced) rc=$?"; tail -3 gpurun_out/r2w_pytest_eig_split.log
timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2w_chunks.log 2>&1; echo "auto (split)"; grep -E "rep 0|eig \(" gpurun_out/r2w_chunks.log | head -6
RCWA_B200_TUNE="13=1" timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2w_chunks_single.log 2>&1; echo "single launch"; grep -E "rep 0|eig \(" gpurun_out/r2w_chunks_single.log | head -6
RCWA_B200_TUNE="13=2,9=2" timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2w_chunks_g2.log 2>&1; echo "split, 2 groups"; grep -E "rep 0|eig \(" gpurun_out/r2w_chunks_g2.log | head -6
timeout 600 python tools/sym_profile.py --points 128 > gpurun_out/r2w_sym_profile.log 2>&1; grep -v Warn gpurun_out/r2w_sym_profile.log | head -12
