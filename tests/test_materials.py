"""torcwa_b200.materials.NKTable against scipy's interp1d(kind='cubic') -- the interpolant the reference's
example/Materials.py:19-20 builds per call -- including the clamping outside the table (:24-29)."""
import numpy as np
import torch
from scipy.interpolate import interp1d


def test_nk_table_matches_scipy_cubic_and_clamps(tmp_path):
    from torcwa_b200.materials import NKTable
    rng = np.random.default_rng(3)
    lam = np.sort(rng.uniform(300.0, 900.0, 40))
    n = 3.5 + 0.8 * np.sin(lam / 90.0) + 0.05 * rng.standard_normal(40)
    k = 0.3 * np.exp(-(lam - 300.0) / 150.0)
    path = tmp_path / "nk.txt"
    np.savetxt(path, np.stack([lam, n, k], 1))
    tab = NKTable.from_file(str(path))
    q = torch.tensor(rng.uniform(lam[0], lam[-1], 500), dtype=torch.float64)
    got = tab.apply(q).numpy()
    ref = interp1d(lam, n, kind='cubic')(q.numpy()) + 1j * interp1d(lam, k, kind='cubic')(q.numpy())
    assert got.dtype == np.complex128 and np.abs(got - ref).max() <= 1e-12
    assert tab.apply(q[:3].to(torch.float32)).dtype == torch.complex64
    lo, hi = tab.apply(torch.tensor(100.0, dtype=torch.float64)), tab.apply(torch.tensor(2000.0, dtype=torch.float64))
    assert abs(complex(lo) - (n[0] + 1j * k[0])) <= 1e-13 and abs(complex(hi) - (n[-1] + 1j * k[-1])) <= 1e-13
    # differentiable: matches the reference's central difference (example/Materials.py:43-49)
    x = torch.tensor(555.5, dtype=torch.float64, requires_grad=True)
    tab.permittivity(x).real.backward()
    dl = 1e-3
    fd = (interp1d(lam, n, kind='cubic')(555.5 + dl) + 1j * interp1d(lam, k, kind='cubic')(555.5 + dl)) ** 2 \
        - (interp1d(lam, n, kind='cubic')(555.5 - dl) + 1j * interp1d(lam, k, kind='cubic')(555.5 - dl)) ** 2
    assert abs(float(x.grad) - (fd / (2 * dl)).real) <= 1e-6 * abs((fd / (2 * dl)).real)
