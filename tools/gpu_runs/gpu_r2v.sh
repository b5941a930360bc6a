mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig.py -m gpu -x -q -k "decoupled" > gpurun_out/r2v_pytest_eig.log 2>&1; echo "pytest eig rc=$?"; tail -5 gpurun_out/r2v_pytest_eig.log
timeout 500 python tools/sym_chunks.py --reps 1 > gpurun_out/r2v_chunks.log 2>&1; tail -14 gpurun_out/r2v_chunks.log
for G in 2 3; do timeout 500 python tools/sym_chunks.py --reps 1 --groups $G > gpurun_out/r2v_chunks_g$G.log 2>&1; echo "groups $G"; grep "rep 0" gpurun_out/r2v_chunks_g$G.log; done
timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2v_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -3 gpurun_out/r2v_pytest_parity.log
