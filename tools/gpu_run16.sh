#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1r_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 gpurun_out/r1r_pytest_gpu.log
timeout 300 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1r_stage_nb128.log 2>&1; echo "stage rc=$?"
grep -h "parity\|eig(total)\|layers/s\|inv(E)\|layer_smatrix\|redheffer" gpurun_out/r1r_stage_nb128.log
