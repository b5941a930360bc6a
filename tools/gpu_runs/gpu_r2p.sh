mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 1 --points 32 > gpurun_out/r2p_bench_2gpu.json 2> gpurun_out/r2p_bench_2gpu.err; echo "2gpu bench rc=$?"; cut -c1-400 gpurun_out/r2p_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2p_bench_2gpu_ref.json 2> gpurun_out/r2p_bench_2gpu_ref.err; echo "2gpu ref rc=$?"; cut -c1-300 gpurun_out/r2p_bench_2gpu_ref.json
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "wrappers_follow" 2>&1 | tail -2
