"""Per-kernel time of one sweep step on the symmetry-reduced path (CUPTI via torch.profiler).  python tools/sym_profile.py --points 128"""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=128)
ap.add_argument("--general", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
case, mask, lams = bench.sweep_inputs(15)
grids = bench.make_grids(mask, lams).to(dev)
freq = (1.0 / lams).to(dev)
def step(s):
    sl = (torch.arange(a.points, device=dev) + s * a.points) % 512
    return bench.run_step(grids[sl], freq[sl], case, dev, False if a.general else None)
step(0); step(1); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(2); e1.record(); torch.cuda.synchronize()
print("step wall %.1f ms for %d points -> %.1f layers/s" % (e0.elapsed_time(e1), a.points, a.points / e0.elapsed_time(e1) * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(3); torch.cuda.synchronize()
per = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CUDA:
        continue
    t = float(getattr(e, "device_time", 0.0) or 0.0)
    if t <= 0: continue
    nm = e.name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").split("(")[0]
    nm = ("zgemm_grouped" + nm[nm.index("<"):][:26]) if "zgemm_grouped_kernel" in nm else nm.split("<")[0].split("::")[-1]
    per[nm][0] += 1; per[nm][1] += t
tot = sum(v[1] for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:28]:
    print("  %-40s launches %6d  total %9.2f ms  avg %8.1f us  share %.3f" % (k, v[0], v[1] / 1e3, v[1] / v[0], v[1] / tot))
print("  sum of kernel time %.1f ms (kernels on different streams overlap)" % (tot / 1e3))
