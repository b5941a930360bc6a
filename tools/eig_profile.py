"""Per-kernel CUPTI profile of one rcwa_eig call on the real order-15 matrices (run under gpurun).
Prints count / total / mean per kernel and the distribution of the QR pass kernel's durations."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from oracle import cases as C
from torcwa_b200 import _lib
import torcwa_b200

ap = argparse.ArgumentParser()
ap.add_argument("--order", type=int, default=15)
ap.add_argument("--nb", type=int, default=96)
ap.add_argument("--out", default=None)
a = ap.parse_args()
d = torch.device("cuda:0")
case = dict(C.CASES["ex1_o15"]); case["order"] = [a.order, a.order]
cd = torch.complex128
thick, grid = C.build_layers(case, cd)[0]
lams = torch.linspace(400.0, 700.0, a.nb, dtype=torch.float64)
sim = torcwa_b200.rcwa(freq=1 / lams, order=case["order"], L=case["L"], dtype=cd, device=d, store_intermediates=False)
sim.add_input_layer(eps=case["eps_in"]); sim.set_incident_angle(0.0, 0.0)
E = _lib.convmat(grid.to(d), a.order, a.order, nb=a.nb)
eta, _ = _lib.inverse(E)
P, Q = _lib.pq_assemble(eta, E, sim._kx, sim._ky, mu_scalar=torch.ones(a.nb, dtype=cd, device=d))
A = _lib.zgemm(P, Q)
del E, eta, P, Q
A0 = A.clone()
_lib.eig(A0)                      # warm-up
torch.cuda.synchronize()
A0.copy_(A)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record(); lam, W, info = _lib.eig(A0); e1.record()
    torch.cuda.synchronize()
per, qr = {}, []
for e in prof.events():
    nm = e.name[:60]
    for key in ("zgemm_grouped_kernel", "qr_pass_kernel", "hb_matvec_kernel", "hb_col_kernel", "trevc_block_kernel", "Memcpy", "Memset"):
        if key in e.name:
            nm = key
            if key == "zgemm_grouped_kernel":
                # template args tell the tile: <BM, BN, WM, WN, OPA, OPB, M3, MINB>
                t = e.name[e.name.find("<") + 1:e.name.find(">")].replace(" ", "").split(",")
                nm = "zgemm %sx%s%s op%s%s" % (t[0], t[1], "+m3" if t[6] in ("true", "1") else "", t[4], t[5]) if len(t) >= 7 else key
            break
    dt = float(getattr(e, "device_time", 0.0) or 0.0)
    v = per.setdefault(nm, [0, 0.0]); v[0] += 1; v[1] += dt
    if "qr_pass_kernel" in e.name:
        qr.append(dt)
tot = e0.elapsed_time(e1)
print(f"eig batch {a.nb} order {a.order}: {tot:.1f} ms wall; info max {int(info.abs().max())}; stats {_lib.last_eig_stats[:2].tolist()}")
rows = sorted(per.items(), key=lambda kv: -kv[1][1])
for k, (c, t) in rows[:24]:
    print(f"  {k[:70]:70s} n={c:7d} total {t/1e3:9.1f} ms  mean {t/max(c,1):8.1f} us")
q = np.array(qr)
if len(q):
    print("  qr_pass durations us: " + ", ".join(f"p{p}={np.percentile(q, p):.0f}" for p in (5, 25, 50, 75, 95, 99)) + f", sum {q.sum()/1e3:.0f} ms, n={len(q)}")
pr = _lib.last_eig_profile.cpu().numpy().astype(float)      # [nb,6,3]
names = ["sweep start (scan+shifts)", "chase window", "small-block slice", "AED Schur slice", "AED scan slice", "AED finish"]
print("  QR pass segments (mean over matrices): count, total ms @1.9GHz, mean us")
for k in range(6):
    cnt, cyc = pr[:, k, 0].mean(), pr[:, k, 1].mean()
    print(f"    {names[k]:28s} n={cnt:8.1f}  total {cyc/1.9e6:8.1f} ms  mean {cyc/max(cnt,1)/1.9e3:8.1f} us   (max-matrix total {pr[:, k, 1].max()/1.9e6:.1f} ms; longest segment {pr[:, k, 2].max()/1.9e3:.0f} us, mean of per-matrix longest {pr[:, k, 2].mean()/1.9e3:.0f} us)")
if a.out:
    json.dump({"wall_ms": tot, "kernels": {k: v for k, v in rows}, "qr_pass_us_percentiles": {str(p): float(np.percentile(q, p)) for p in (5, 25, 50, 75, 95, 99)} if len(q) else None}, open(a.out, "w"), indent=1)
