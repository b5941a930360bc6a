mkdir -p gpurun_out
echo "--- max connections 32"; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python tools/pipe_probe.py --points 64 --pipes 1,2 --steps 1 2>&1 | tail -2
echo "--- max connections 32, one QR group per eig"; CUDA_DEVICE_MAX_CONNECTIONS=32 RCWA_B200_TUNE="9=1" timeout 600 python tools/pipe_probe.py --points 64 --pipes 1,2 --steps 1 2>&1 | tail -2
echo "--- default connections, one QR group per eig"; RCWA_B200_TUNE="9=1" timeout 600 python tools/pipe_probe.py --points 64 --pipes 2 --steps 1 2>&1 | tail -1
