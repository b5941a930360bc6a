#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig.py tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r1m_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r1m_pytest_gpu.log
timeout 300 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1m_stage_nb128.log 2>&1; echo "stage rc=$?"
grep -h "parity\|eig(total)\|layers/s\|hessenberg(alone)\|stats" gpurun_out/r1m_stage_nb128.log
timeout 300 python tools/eig_profile.py --nb 128 --out gpurun_out/r1m_eig_profile.json > gpurun_out/r1m_eig_profile.log 2>&1; echo "profile rc=$?"
grep -v Warn gpurun_out/r1m_eig_profile.log | grep -v "^  (anon\|cuda\|Buffer\|Activity\|Command" | head -30
