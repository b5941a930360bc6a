#!/usr/bin/env python
"""tests/golden/eig_backward.npz: inputs and outputs of the UNMODIFIED reference's Eig.backward
(/root/reference/torcwa/torch_eig.py:19-44) on seeded random complex128 problems and on the order-3 RCWA matrix.
Build container only (needs /root/reference):   PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_eig_backward.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import torcwa  # noqa: E402  the reference, unmodified

from oracle import cases as C  # noqa: E402
from oracle.rcwa_oracle import OracleSim  # noqa: E402


class Ctx:
    pass


def ref_backward(A, w, V, gw, gV, broadening):
    old = torcwa.Eig.broadening_parameter
    torcwa.Eig.broadening_parameter = broadening
    try:
        ctx = Ctx()
        ctx.input, ctx.eigval, ctx.eigvec = A, w, V
        return torcwa.Eig.backward(ctx, gw, gV)
    finally:
        torcwa.Eig.broadening_parameter = old


def main():
    out = {}
    g = torch.Generator().manual_seed(20261017)
    mats = {}
    for n in (6, 24, 57):
        mats["rand%d" % n] = torch.complex(torch.randn(n, n, generator=g, dtype=torch.float64), torch.randn(n, n, generator=g, dtype=torch.float64))
    case = C.CASES["ex1_o3"]
    sim = OracleSim(freq=C.freq_of(case, torch.complex128), order=case["order"], L=case["L"], dtype=torch.complex128)
    sim.add_input_layer(eps=case["eps_in"]); sim.set_incident_angle(0.0, 0.0)
    d, e = C.build_layers(case, torch.complex128)[0]
    sim.add_layer(d, e)
    mats["rcwa_o3"] = (sim.P[0] @ sim.Q[0]).clone()
    for name, A in mats.items():
        n = A.shape[0]
        w, V = torch.linalg.eig(A)
        gw = torch.complex(torch.randn(n, generator=g, dtype=torch.float64), torch.randn(n, generator=g, dtype=torch.float64))
        gV = torch.complex(torch.randn(n, n, generator=g, dtype=torch.float64), torch.randn(n, n, generator=g, dtype=torch.float64))
        out[name + "_A"], out[name + "_w"], out[name + "_V"] = A.numpy(), w.numpy(), V.numpy()
        out[name + "_gw"], out[name + "_gV"] = gw.numpy(), gV.numpy()
        out[name + "_grad_b1e-10"] = ref_backward(A, w, V, gw, gV, 1e-10).numpy()
        out[name + "_grad_bNone"] = ref_backward(A, w, V, gw, gV, None).numpy()
    path = os.path.join(ROOT, "tests", "golden", "eig_backward.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
