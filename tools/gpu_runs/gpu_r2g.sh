mkdir -p gpurun_out
for dg in 0 5; do timeout 600 python tools/stage_kernels.py --nb 32 --digits $dg > gpurun_out/r2g_stage_kernels_d$dg.log 2>&1; echo "rc=$?"; cat gpurun_out/r2g_stage_kernels_d$dg.log; done
