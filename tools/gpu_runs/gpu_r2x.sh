mkdir -p gpurun_out
for G in 4 6 8; do RCWA_B200_TUNE="9=$G" timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2x_chunks_g$G.log 2>&1; echo "split, $G groups"; grep -E "rep 0|eig \(" gpurun_out/r2x_chunks_g$G.log | head -5; done
RCWA_B200_TUNE="9=8,8=60" timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2x_chunks_g8_b60.log 2>&1; echo "split, 8 groups, 60 us slices"; grep -E "rep 0|eig \(" gpurun_out/r2x_chunks_g8_b60.log | head -5
