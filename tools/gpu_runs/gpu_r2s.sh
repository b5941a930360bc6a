mkdir -p gpurun_out
timeout 600 python tools/sym_profile.py --points 128 > gpurun_out/r2s_sym_profile.log 2>&1; grep -v Warn gpurun_out/r2s_sym_profile.log | tail -34
