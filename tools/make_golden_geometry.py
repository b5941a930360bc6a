#!/usr/bin/env python
"""tests/golden/geometry.npz: every shape of the UNMODIFIED reference's torcwa.geometry (instance flavour) and
torcwa.rcwa_geo (class flavour) on a small grid (torcwa/geometry.py:4-290).  Build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import torcwa  # noqa: E402

from oracle.geometry_case import shapes, setup  # noqa: E402

out = {}
g = torcwa.geometry(Lx=320.0, Ly=240.0, nx=24, ny=20, edge_sharpness=35.0, dtype=torch.float64, device=torch.device("cpu"))
for name, fn in shapes().items():
    out["inst_" + name] = fn(g).numpy()
setup(torcwa.rcwa_geo)
for name, fn in shapes().items():
    out["cls_" + name] = fn(torcwa.rcwa_geo).numpy()
path = os.path.join(ROOT, "tests", "golden", "geometry.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out))
