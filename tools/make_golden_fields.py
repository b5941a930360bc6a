#!/usr/bin/env python
"""tests/golden/fields_stack_o3.npz: fields of the UNMODIFIED reference (rcwa.source_*, field_xz / field_yz /
field_xy, torcwa/rcwa.py:526-1112) on the first three layers of the stack_o3 case, forward xy plane wave and
backward ps Fourier source.  Build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import torcwa  # noqa: E402

from oracle.fields_case import build, SOURCES, planes  # noqa: E402

sim = build(lambda **kw: torcwa.rcwa(device=torch.device("cpu"), **kw))
out = {}
for sname, setter in SOURCES.items():
    setter(sim)
    for pname, getter in planes().items():
        E, H = getter(sim)
        out["%s_%s" % (sname, pname)] = np.stack([t.numpy() for t in E + H])
path = os.path.join(ROOT, "tests", "golden", "fields_stack_o3.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()})
