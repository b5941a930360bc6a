"""Per-kernel time of the S-matrix stage (layer S-matrix + star product) at the path's size, by engine (CUPTI via torch.profiler).
    python tools/stage_kernels.py --nb 32 --digits 5"""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torcwa_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=32)
ap.add_argument("--N", type=int, default=961)
ap.add_argument("--digits", type=int, default=5)
a = ap.parse_args()
d = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
def rnd(*sh):
    return torch.complex(torch.randn(*sh, generator=g, dtype=torch.float64), torch.randn(*sh, generator=g, dtype=torch.float64)).to(d)
nb, N = a.nb, a.N
n = 2 * N
W = (rnd(nb, n, n) / np.sqrt(n) + 2.0 * torch.eye(n, dtype=torch.complex128, device=d)).contiguous()
Q = rnd(nb, n, n) / np.sqrt(n)
kz = rnd(nb, n) * 0.3 + 1.0
kz = torch.complex(kz.real.abs() + 0.2, kz.imag.abs()).contiguous()
vfinv = (rnd(nb, 4, N) * 0.2 + torch.tensor([1.0, 0.0, 0.0, 1.0], dtype=torch.complex128, device=d)[None, :, None]).contiguous()
omega = torch.full((nb,), 2 * np.pi / 532.0, dtype=torch.float64, device=d)
thick = torch.full((nb,), 100.0, dtype=torch.float64, device=d)
bd = [rnd(nb, 4, N) * 0.3 for _ in range(4)]
def stage():
    S11, S21, _ = _lib.layer_smatrix(W, kz, Q, vfinv, omega, thick, slices=a.digits)
    out, _ = _lib.redheffer_bdleft(bd, [S11, S21, S21, S11], slices=a.digits)
    return out
stage(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); stage(); e1.record(); torch.cuda.synchronize()
print("stage wall %.1f ms for %d points (%.2f ms/point), digits %d" % (e0.elapsed_time(e1), nb, e0.elapsed_time(e1) / nb, a.digits))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    stage(); torch.cuda.synchronize()
per = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    t = float(getattr(e, "device_time", 0.0) or 0.0)
    if t <= 0: continue
    nm = e.name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    nm = nm.split("(")[0]
    if "zgemm_grouped_kernel" in nm:
        nm = "zgemm_grouped" + nm[nm.index("<"):][:24]
    else:
        nm = nm.split("<")[0].split("::")[-1]
    per[nm][0] += 1; per[nm][1] += t
tot = sum(v[1] for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:14]:
    print("  %-28s launches %5d  total %9.2f ms  avg %8.1f us  share %.3f" % (k, v[0], v[1] / 1e3, v[1] / v[0], v[1] / tot))
print("  sum of kernel time %.1f ms" % (tot / 1e3))
