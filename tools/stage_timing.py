"""Per-stage device timings of one patterned layer at a given order / batch (run under gpurun)."""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import cases as C
from torcwa_b200 import _lib
from torcwa_b200.rcwa import vf_inverse_diagonals
import torcwa_b200

ap = argparse.ArgumentParser()
ap.add_argument("--order", type=int, default=15)
ap.add_argument("--nb", type=int, default=4)
ap.add_argument("--check", action="store_true")
ap.add_argument("--digits", type=int, default=0, help="tcgen05 digits of the S-matrix stage (0 = fp64 DMMA)")
ap.add_argument("--residual", action="store_true", help="also report max |A W - W diag(lam)| / |A| of the eigendecomposition")
a = ap.parse_args()
d = torch.device("cuda:0")

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

case = dict(C.CASES["ex1_o15"]); case["order"] = [a.order, a.order]
cd = torch.complex128
thick, grid = C.build_layers(case, cd)[0]
lams = torch.linspace(400.0, 700.0, a.nb, dtype=torch.float64) if a.nb > 1 else torch.tensor([532.0], dtype=torch.float64)
if a.check: lams[0] = 532.0
sim = torcwa_b200.rcwa(freq=1 / lams if a.nb > 1 else 1 / lams[0], order=case["order"], L=case["L"], dtype=cd, device=d, store_intermediates=False)
sim.add_input_layer(eps=case["eps_in"]); sim.set_incident_angle(0.0, 0.0)
grid_d = grid.to(d)
torch.cuda.synchronize()
t = {}
for rep in range(2):
    e0 = ev(); E = _lib.convmat(grid_d, a.order, a.order, nb=a.nb)
    e1 = ev(); eta, _ = _lib.inverse(E)
    e2 = ev(); P, Q = _lib.pq_assemble(eta, E, sim._kx, sim._ky, mu_scalar=torch.ones(a.nb, dtype=cd, device=d))
    e3 = ev(); A = _lib.zgemm(P, Q)
    e4 = ev(); H = A.clone(); Z = _lib.hessenberg_(H)
    A_keep = A.clone() if (a.residual and rep == 1) else None
    e5 = ev(); del H, Z; lam, W, info = _lib.eig(A)
    e6 = ev(); kz = _lib.kz_branch(lam)
    om = sim._omega64.expand(a.nb).contiguous(); th = torch.full((a.nb,), float(thick), dtype=torch.float64, device=d)
    S11, S21, _ = _lib.layer_smatrix(W, kz, Q, sim._Vf_inv, om, th, slices=a.digits)
    e7 = ev(); S, _ = _lib.redheffer_bdleft(sim._Sin, [S11, S21, S21, S11], slices=a.digits)
    e8 = ev(); torch.cuda.synchronize()
    names = ["convmat", "inv(E)", "pq_assemble", "P@Q", "hessenberg(alone)", "eig(total)", "layer_smatrix", "redheffer(Sin*S)"]
    evs = [e0, e1, e2, e3, e4, e5, e6, e7, e8]
    t = {names[i]: evs[i].elapsed_time(evs[i + 1]) for i in range(8)}
print(torch.cuda.get_device_name(0), f"order {a.order} n={2*sim.order_N} batch {a.nb} digits {a.digits}")
tot = sum(v for k, v in t.items() if k != "hessenberg(alone)")
for k, v in t.items(): print(f"  {k:20s} {v:10.2f} ms  ({v/a.nb:9.2f} ms/point)")
print(f"  total (excl. standalone hessenberg) {tot:.1f} ms -> {a.nb/tot*1e3:.3f} layers/s")
print("  eig info max:", int(info.abs().max()), " stats [sweeps, passes, aeds, info]:", _lib.last_eig_stats.tolist()[:3])
n = 2 * sim.order_N
beig = 16.0 * (n ** 3 / 3 + 2 * n * n)
print(f"  B_eig = {beig/1e9:.2f} GB/matrix -> eig stage {a.nb*beig/t['eig(total)']/1e6:.0f} GB/s, hessenberg alone {a.nb*beig/t['hessenberg(alone)']/1e6:.0f} GB/s (algorithmic)")
if a.residual:
    R = torch.matmul(A_keep, W) - W * lam[:, None, :]
    print("  eig residual max|A W - W L| / max|A| per matrix:", float((R.abs().amax(dim=(1, 2)) / A_keep.abs().amax(dim=(1, 2))).max()),
          " cond-free check |W col norms - 1| max:", float((torch.linalg.norm(W, dim=1) - 1).abs().max()))
if a.check:
    sim._S = S; sim.S = [s[0] if a.nb == 1 else s for s in S]
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ex1_o15.npz"))
    if a.nb > 1:
        one = torcwa_b200.rcwa(freq=1 / lams[0], order=case["order"], L=case["L"], dtype=cd, device=d); one.add_input_layer(eps=case["eps_in"]); one.set_incident_angle(0., 0.)
        one._S = [s[0:1] for s in S]; sp = C.probe(one)
    else:
        sp = C.probe(sim)
    print("  order-15 parity vs reference-c128 golden: max S-param err / max|S| =", np.abs(sp - g["sparams_c128"]).max() / np.abs(g["sparams_c128"]).max())
    res = (A if False else None)
