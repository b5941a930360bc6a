"""The UNMODIFIED reference (kch3782/torcwa installed into baseline/_ref, git-ignored) on the GPU of the box:
layers/s of BASELINE config 2's unit (order 15x15, one patterned layer + SiO2 half space, complex64 and
complex128, sequential loop over wavelengths as the reference's examples do) -- the denominator of the
north-star's ">= 10x the reference PyTorch-CUDA path".  Run under gpurun; prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(1, ROOT)
import torch
import torcwa                                   # the reference
from oracle import cases as C                   # input builders only

ap = argparse.ArgumentParser()
ap.add_argument("--order", type=int, default=15)
ap.add_argument("--points", type=int, default=6)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
out = {"impl": "reference-cuda", "torcwa": getattr(torcwa, "__version__", "?"), "order": a.order, "gpu": torch.cuda.get_device_name(0)}
for name, cd in (("c64", torch.complex64), ("c128", torch.complex128)):
    rd = torch.float32 if cd == torch.complex64 else torch.float64
    mask = C.rectangle_grid(300.0, 300.0, 300, 300, 180.0, 100.0, 150.0, 150.0, 0.0, 1000.0, rd).to(dev)
    lams = torch.linspace(400.0, 700.0, a.points + 2, dtype=rd)
    ts, txx = [], None
    for i, lam in enumerate(lams):
        eps_si = complex(C.SI_EPS[532.0])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sim = torcwa.rcwa(freq=1 / lam.to(dev), order=[a.order, a.order], L=[300.0, 300.0], dtype=cd, device=dev)
        sim.add_input_layer(eps=1.46 ** 2)
        sim.set_incident_angle(inc_ang=0.0, azi_ang=0.0)
        sim.add_layer(thickness=300.0, eps=mask * eps_si + (1.0 - mask))
        sim.solve_global_smatrix()
        txx = sim.S_parameters(orders=[0, 0], direction="forward", port="transmission", polarization="xx", ref_order=[0, 0])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if i >= 2:
            ts.append(dt)
        del sim
    out[name] = {"layers_per_s": len(ts) / sum(ts), "s_per_layer": sum(ts) / len(ts), "points": len(ts), "txx_last": [float(txx.real), float(txx.imag)]}
print(json.dumps(out))
