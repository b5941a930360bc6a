"""Throughput of one sweep step as a function of the number of pipelined sub-batches (python tools/pipe_probe.py --points 64)."""
import argparse, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=64)
ap.add_argument("--pipes", default="1,2")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
case, mask, lams = bench.sweep_inputs(15)
grids = bench.make_grids(mask, lams).to(dev)
freq = (1.0 / lams).to(dev)
for pl in [int(x) for x in a.pipes.split(",")]:
    os.environ["RCWA_B200_PIPELINE"] = str(pl)
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    def step(s):
        sl = (torch.arange(a.points, device=dev) + s * a.points) % 512
        return bench.run_step(grids[sl], freq[sl], case, dev)
    step(0); torch.cuda.synchronize()
    r0 = torch.cuda.memory_stats().get("num_alloc_retries", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        out = step(1 + s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"pipeline": pl, "points": a.points, "ms_per_step": ms, "layers_per_s": a.points / ms * 1e3,
                      "peak_GB": torch.cuda.max_memory_allocated() / 1e9, "reserved_GB": torch.cuda.max_memory_reserved() / 1e9,
                      "alloc_retries": torch.cuda.memory_stats().get("num_alloc_retries", 0) - r0}), flush=True)
