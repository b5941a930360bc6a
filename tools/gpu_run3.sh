#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_eig.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1d_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/r1d_pytest_gpu.log
tail -3 gpurun_out/r1d_pytest_gpu.log
timeout 600 python tools/stage_timing.py --nb 96 --check > gpurun_out/r1d_stage_nb96.log 2>&1; echo "stage rc=$?"
timeout 600 python tools/stage_timing.py --nb 128 > gpurun_out/r1d_stage_nb128.log 2>&1; echo "stage rc=$?"
timeout 600 python tools/eig_profile.py --nb 96 --out gpurun_out/r1d_eig_profile.json > gpurun_out/r1d_eig_profile.log 2>&1; echo "eig profile rc=$?"
grep -h "layers/s\|parity\|eig(total)\|hessenberg(alone)\|stats" gpurun_out/r1d_stage_*.log
grep -v Warn gpurun_out/r1d_eig_profile.log | head -24
