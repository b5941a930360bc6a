#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "fields or differentiable or parity_c128" > gpurun_out/r1s_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; grep "fields:\|passed\|failed\|Error" gpurun_out/r1s_pytest_gpu.log | tail -8
