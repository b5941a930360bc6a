mkdir -p gpurun_out
RCWA_B200_TUNE="13=2,14=2" timeout 1500 python -m pytest tests/test_gpu_eig.py -m gpu -x -q > gpurun_out/r2y_pytest_eig_graph.log 2>&1; echo "pytest eig (split + graphs forced) rc=$?"; tail -3 gpurun_out/r2y_pytest_eig_graph.log
for G in 4 6 8; do RCWA_B200_TUNE="9=$G" timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2y_chunks_g$G.log 2>&1; echo "split + graphs, $G groups"; grep -E "rep 0|eig \(|Error|error" gpurun_out/r2y_chunks_g$G.log | head -5; done
RCWA_B200_TUNE="14=1" timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2y_chunks_nograph.log 2>&1; echo "split, no graphs"; grep -E "rep 0|eig \(" gpurun_out/r2y_chunks_nograph.log | head -5
