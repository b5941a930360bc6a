mkdir -p gpurun_out
timeout 900 python tools/pipe_probe.py --points 64 --pipes 1,2,3 > gpurun_out/r2i_pipe64.log 2>&1; cat gpurun_out/r2i_pipe64.log | tail -5
timeout 900 python tools/pipe_probe.py --points 96 --pipes 1,2 > gpurun_out/r2i_pipe96.log 2>&1; cat gpurun_out/r2i_pipe96.log | tail -5
