mkdir -p gpurun_out
timeout 400 python tools/sym_chunks.py --reps 1 --max-chunks 2 > gpurun_out/r3d_base.log 2>&1; echo "default (pair-sparse half spaces)"; grep -E "rep 0|eig \(|rror" gpurun_out/r3d_base.log | head -4
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r3d_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -3 gpurun_out/r3d_pytest_parity.log
for V in aed32 aed40; do RCWA_B200_LIB=librcwa_b200_$V.so timeout 400 python tools/sym_chunks.py --reps 1 --max-chunks 2 > gpurun_out/r3c_$V.log 2>&1; echo "variant $V"; grep -E "rep 0|eig \(|rror" gpurun_out/r3c_$V.log | head -4; done
for B in 120 160; do RCWA_B200_TUNE="8=$B" timeout 400 python tools/sym_chunks.py --reps 1 --max-chunks 2 > gpurun_out/r3c_b$B.log 2>&1; echo "slice budget $B us"; grep -E "rep 0|eig \(|rror" gpurun_out/r3c_b$B.log | head -4; done
