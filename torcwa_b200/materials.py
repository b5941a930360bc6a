"""Tabulated dispersive materials evaluated on the device for a whole sweep (scope row f2, second half).

The reference's examples re-read an (wavelength, n, k) text file and build two scipy cubic interpolants for EVERY
wavelength of a sweep, on the host, one scalar at a time (example/Materials.py:5-50).  Here the not-a-knot cubic
splines are built once (scipy, host) and kept as piecewise-polynomial coefficient tables on the device; `apply`
evaluates them for a tensor of wavelengths with torch ops (searchsorted + Horner), clamps outside the table like the
reference does (:24-29), and is differentiable through torch itself (the reference uses a central finite difference,
:43-49).  Not part of the hot path."""
import numpy as np
import torch


class NKTable:
    def __init__(self, wavelength, n, k, *, device=None):
        from scipy.interpolate import CubicSpline            # same spline as interp1d(kind='cubic'): not-a-knot
        lam = np.asarray(wavelength, dtype=np.float64)
        order = np.argsort(lam)
        lam, n, k = lam[order], np.asarray(n, dtype=np.float64)[order], np.asarray(k, dtype=np.float64)[order]
        self.device = torch.device(device) if device is not None else torch.device('cpu')
        self._x = torch.tensor(lam, dtype=torch.float64, device=self.device)
        # PPoly coefficients c[m, i]: sum_m c[m, i] (x - x_i)^(3 - m) on [x_i, x_{i+1}]
        self._cn = torch.tensor(CubicSpline(lam, n, bc_type='not-a-knot').c, dtype=torch.float64, device=self.device)
        self._ck = torch.tensor(CubicSpline(lam, k, bc_type='not-a-knot').c, dtype=torch.float64, device=self.device)

    @classmethod
    def from_file(cls, path, *, device=None):
        """Whitespace-separated rows `wavelength n k` (the format of example/Materials_data/*.txt)."""
        data = np.loadtxt(path)
        return cls(data[:, 0], data[:, 1], data[:, 2], device=device)

    def to(self, device):
        self.device = torch.device(device)
        self._x, self._cn, self._ck = self._x.to(device), self._cn.to(device), self._ck.to(device)
        return self

    def apply(self, wavelength):
        """Complex refractive index n + i k at the given wavelength(s); complex128 for float64 input, else complex64."""
        lam = torch.as_tensor(wavelength, device=self.device)
        out_dtype = torch.complex128 if lam.dtype in (torch.float64, torch.complex128) else torch.complex64
        x = torch.clamp(lam.real.to(torch.float64) if torch.is_complex(lam) else lam.to(torch.float64), self._x[0], self._x[-1])
        i = torch.clamp(torch.searchsorted(self._x, x.detach(), right=True) - 1, 0, len(self._x) - 2)
        t = x - self._x[i]

        def horner(c):
            return ((c[0][i] * t + c[1][i]) * t + c[2][i]) * t + c[3][i]
        return torch.complex(horner(self._cn), horner(self._ck)).to(out_dtype)

    def permittivity(self, wavelength):
        return self.apply(wavelength) ** 2
