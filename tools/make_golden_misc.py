#!/usr/bin/env python
"""tests/golden/misc_stack_o3.npz: outputs of the UNMODIFIED reference's small utilities rcwa.diffraction_angle
(rcwa.py:214-262) and rcwa.return_layer (:264-298) on the stack_o3 case.  Build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import torcwa  # noqa: E402

from oracle import cases as C  # noqa: E402

ORDERS = [[0, 0], [1, 0], [-1, 1], [2, -3]]

sim = C.run_case(lambda **kw: torcwa.rcwa(device=torch.device("cpu"), **kw), C.CASES["stack_o3"], torch.complex128)
out = {"orders": np.array(ORDERS)}
for layer in ("input", "output"):
    for unit in ("radian", "degree"):
        inc, azi = sim.diffraction_angle(ORDERS, layer=layer, unit=unit)
        out["inc_%s_%s" % (layer, unit)], out["azi_%s_%s" % (layer, unit)] = inc.numpy(), azi.numpy()
e, m = sim.return_layer(0, nx=20, ny=26)
out["eps_rec"], out["mu_rec"] = e.numpy(), m.numpy()
path = os.path.join(ROOT, "tests", "golden", "misc_stack_o3.npz")
np.savez_compressed(path, **out)
print("wrote", path)
