#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (kch3782/torcwa, mounted
read-only at /root/reference) on CPU.  Build container only (the GPU box has no /root/reference;
tests there read the committed .npz files):

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py [--big] [--only NAME]

Inputs are the named cases of oracle/cases.py.  Stored per case: the reference's S-parameters
for a fixed probe set in complex128 and complex64, sorted kz^2 per layer, and for small orders the
convolution matrix, first-layer S blocks and the four global S blocks (complex128); for larger
orders only the Frobenius norms and two full columns of each global block.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import torcwa  # noqa: E402  the reference, unmodified

from oracle import cases as C  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def check_inputs_against_reference():
    """(1) our a-Si:H constants == cubic interpolation of the reference's data file;
       (2) rectangle_grid == torcwa.rcwa_geo.rectangle, bit for bit, both precisions."""
    from scipy.interpolate import interp1d
    d = np.loadtxt("/root/reference/example/Materials_data/aSiH.txt")
    for lam, val in C.SI_EPS.items():
        n = interp1d(d[:, 0], d[:, 1], kind="cubic")(lam)
        k = interp1d(d[:, 0], d[:, 2], kind="cubic")(lam)
        assert abs(complex((n + 1j * k) ** 2) - val) < 1e-13, lam
    g = torcwa.rcwa_geo
    for rd in (torch.float32, torch.float64):
        g.dtype, g.device = rd, torch.device("cpu")
        g.Lx, g.Ly, g.nx, g.ny, g.edge_sharpness = 300.0, 300.0, 300, 300, 1000.0
        g.grid()
        for th in (0.0, 0.5):
            a = g.rectangle(Wx=180.0, Wy=100.0, Cx=150.0, Cy=150.0, theta=th)
            b = C.rectangle_grid(300.0, 300.0, 300, 300, 180.0, 100.0, 150.0, 150.0, th, 1000.0, rd)
            assert torch.equal(a, b), (rd, th)


def ref_factory(freq, order, L, dtype):
    return torcwa.rcwa(freq=freq, order=order, L=L, dtype=dtype, device=torch.device("cpu"), stable_eig_grad=False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also run the order-15 case (minutes)")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    check_inputs_against_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, case in C.CASES.items():
        if args.only and name != args.only:
            continue
        if case["big"] and not (args.big or args.only == name):
            continue
        rec = {}
        for tag, cdtype in (("c128", torch.complex128), ("c64", torch.complex64)):
            if tag == "c64" and case.get("c128_only"):
                continue
            t0 = time.time()
            sim = C.run_case(ref_factory, case, cdtype)
            dt = time.time() - t0
            rec[f"sparams_{tag}"] = C.probe(sim)
            rec[f"seconds_{tag}"] = np.array(dt)
            if tag == "c128":
                if sim.kz_norm:
                    rec["kz2_sorted"] = np.stack([np.sort_complex(k.numpy() ** 2) for k in sim.kz_norm])
                rec["S_fro"] = np.array([float(torch.linalg.norm(s)) for s in sim.S])
                cols = [sim.order_N // 2, sim.order_N // 2 + sim.order_N]
                rec["S_cols_idx"] = np.array(cols)
                rec["S_cols"] = np.stack([(s if s.dim() == 2 else torch.diag(s))[:, cols].numpy() for s in sim.S])
                if case["full"]:
                    rec["S"] = np.stack([(s if s.dim() == 2 else torch.diag(s)).numpy() for s in sim.S])
                    if sim.eps_conv:
                        rec["eps_conv0"] = sim.eps_conv[0].numpy()
                        rec["layer_S0"] = np.stack([sim.layer_S11[0].numpy(), sim.layer_S21[0].numpy(),
                                                    sim.layer_S12[0].numpy(), sim.layer_S22[0].numpy()])
            print(f"{name:12s} {tag}: {dt:7.2f}s  t00_xx={complex(rec[f'sparams_{tag}'][0, 0, 0]):.6f}", flush=True)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


if __name__ == "__main__":
    main()
