"""torcwa_b200.geometry / rcwa_geo (drop-in names for the reference's input rasterisers, torcwa/geometry.py:4-290)
against the unmodified reference's outputs (tests/golden/geometry.npz, tools/make_golden_geometry.py)."""
import os

import numpy as np
import pytest
import torch

from oracle.geometry_case import shapes, setup


@pytest.mark.parametrize("flavour", ["inst", "cls"])
def test_shapes_match_reference(flavour, golden_dir):
    import torcwa_b200
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    if flavour == "inst":
        geo = torcwa_b200.geometry(Lx=320.0, Ly=240.0, nx=24, ny=20, edge_sharpness=35.0, dtype=torch.float64, device=torch.device("cpu"))
    else:
        geo = torcwa_b200.rcwa_geo
        setup(geo)
    for name, fn in shapes().items():
        got = fn(geo).numpy()
        ref = g["%s_%s" % (flavour, name)]
        assert got.shape == ref.shape, name
        assert np.abs(got - ref).max() <= 1e-12, name
