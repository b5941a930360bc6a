#!/usr/bin/env python
"""tests/golden/autograd_o3.npz: figure of merit and its gradients from the UNMODIFIED reference (CPU autograd,
complex128, the reference's stabilised Eig backward) for a two-layer stack at order 3x2 under oblique incidence.
Build container only:   PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_autograd.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import torcwa  # noqa: E402  the reference, unmodified

from oracle.autograd_case import CASE, density, fom  # noqa: E402


def main():
    cd = torch.complex128
    rho = density().requires_grad_(True)
    thick = torch.tensor(CASE["thickness"], dtype=torch.float64, requires_grad=True)
    sim = torcwa.rcwa(freq=torch.tensor(1.0 / CASE["lam"], dtype=torch.float64), order=CASE["order"], L=CASE["L"], dtype=cd,
                      device=torch.device("cpu"))
    value = fom(sim, rho, thick)
    value.backward()
    out = {"rho": rho.detach().numpy(), "fom": value.detach().numpy(), "grad_rho": rho.grad.numpy(), "grad_thickness": thick.grad.numpy()}
    path = os.path.join(ROOT, "tests", "golden", "autograd_o3.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "fom", float(value), "|grad_rho|", float(rho.grad.norm()), "grad_thickness", float(thick.grad))


if __name__ == "__main__":
    main()
