"""CPU oracle for the RCWA hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (its ``cpu_baseline`` leg and
``--impl reference`` arm) may import this module.  The product package ``torcwa_b200`` never
imports it and never falls back to it.

What it is
----------
A plain restatement, on the CPU, of the algorithm the reference solver (kch3782/torcwa @ 51c0d24,
``torcwa/rcwa.py``) runs for one design point of the path

    Fourier factorisation -> per-layer eigendecomposition -> Redheffer cascade -> S-parameters.

It keeps the reference's *dense* formulation on purpose (two dense inverses of the 4N x 4N coupling
matrix per layer, dense products against diagonal matrices, two inverses per star product), so
that (a) it is an independent check of the new minimal-algebra CUDA path and (b) timing it is a
fair stand-in for "the reference's own CPU path" (same torch/MKL LAPACK calls, same flop count).

Where the arithmetic lives
--------------------------
The reference has no native code: every number is produced by PyTorch (``torch>=1.10.1``,
unpinned; here torch 2.11.0+cu128 / MKL 2024.2) -> LAPACK ``?geev`` (``torch.linalg.eig``),
``?getrf/?getri`` (``torch.linalg.inv``), ``?gemm`` and the MKL FFT.  This file calls the same
torch CPU entry points.

Parity pinning
--------------
The reference ships no tests and no golden vectors ("parity unpinned" by the reference itself).
The oracle is therefore pinned against the *live reference imported in the build container*:
``tools/make_golden.py`` runs the unmodified reference and this oracle on the same inputs and
stores the reference outputs under ``tests/golden/``; ``tests/test_oracle.py`` asserts
oracle == stored reference outputs (<=1e-12 in complex128).  Soft pins from the reference's
notebooks (Fresnel identity of Example0, ``Delta: 0.287`` of Example5) are in the same test file.

Citations are ``file:line`` in /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

# torcwa/rcwa.py:5 -- the reference's pi is mistyped; parity at 1e-10 needs the same constant.
PI_REF = 3.141592652589793


def _cdtype_real(cdtype):
    return torch.float32 if cdtype == torch.complex64 else torch.float64


def _sqrt_upper(z):
    """sqrt with the branch the reference uses for half-space / homogeneous kz:
    conjugate when Im < 0 (rcwa.py:1143-1144, :1152-1153, :1169-1170, :1217-1218)."""
    r = torch.sqrt(z)
    return torch.where(r.imag < 0, r.conj(), r)


def _diag_blocks_to_dense(d11, d12, d21, d22):
    """[[diag d11, diag d12],[diag d21, diag d22]] as a dense 2N x 2N matrix."""
    top = torch.hstack((torch.diag(d11), torch.diag(d12)))
    bot = torch.hstack((torch.diag(d21), torch.diag(d22)))
    return torch.vstack((top, bot))


def _v_matrix(kx, ky, kz):
    """E->H matrix of a homogeneous medium with normal wavevector kz (rcwa.py:1145-1147):
    [[-kx ky/kz, -(kz + ky^2/kz)], [kz + kx^2/kz, kx ky/kz]] as four diagonal blocks."""
    return _diag_blocks_to_dense(-ky * kx / kz, -kz - ky * ky / kz,
                                 kz + kx * kx / kz, kx * ky / kz)


@dataclass
class OracleSim:
    """One design point, dense CPU algebra.  Attribute names follow the reference's public
    attributes so tests can compare field by field."""
    freq: float
    order: Sequence[int]
    L: Sequence[float]
    dtype: torch.dtype = torch.complex128
    eps_in: complex = 1.0
    mu_in: complex = 1.0
    eps_out: complex = 1.0
    mu_out: complex = 1.0
    has_in: bool = False
    has_out: bool = False
    # filled while solving
    layer_S: List[List[torch.Tensor]] = field(default_factory=list)
    kz_norm: List[torch.Tensor] = field(default_factory=list)
    E_eigvec: List[torch.Tensor] = field(default_factory=list)
    eps_conv: List[torch.Tensor] = field(default_factory=list)
    P: List[torch.Tensor] = field(default_factory=list)
    Q: List[torch.Tensor] = field(default_factory=list)

    def __post_init__(self):
        ct = self.dtype
        self.ox, self.oy = int(self.order[0]), int(self.order[1])
        self.mx = torch.arange(-self.ox, self.ox + 1, dtype=torch.int64)
        self.my = torch.arange(-self.oy, self.oy + 1, dtype=torch.int64)
        self.order_N = len(self.mx) * len(self.my)          # rcwa.py:68
        self._freq = torch.as_tensor(self.freq, dtype=ct)   # rcwa.py:60
        self.omega = 2 * PI_REF * self.freq                 # rcwa.py:61 (raw argument, not cast)
        self.Gx = 1 / (self.L[0] * self._freq)              # rcwa.py:72
        self.Gy = 1 / (self.L[1] * self._freq)
        self.eps_in = torch.as_tensor(self.eps_in, dtype=ct)
        self.mu_in = torch.as_tensor(self.mu_in, dtype=ct)
        self.eps_out = torch.as_tensor(self.eps_out, dtype=ct)
        self.mu_out = torch.as_tensor(self.mu_out, dtype=ct)
        self.thickness = []

    # ------------------------------------------------------------------ k-vectors, half spaces
    def set_incident_angle(self, inc_ang=0.0, azi_ang=0.0, angle_layer="input"):
        """rcwa.py:123-144 -> _kvectors rcwa.py:1124-1181."""
        ct = self.dtype
        inc = torch.as_tensor(inc_ang, dtype=ct)
        azi = torch.as_tensor(azi_ang, dtype=ct)
        self.inc_ang, self.azi_ang = inc, azi
        if angle_layer in ("i", "in", "input"):
            n_ref = torch.sqrt(self.eps_in * self.mu_in).real
        else:
            n_ref = torch.sqrt(self.eps_out * self.mu_out).real
        kx0 = n_ref * torch.sin(inc) * torch.cos(azi)          # rcwa.py:1125-1130
        ky0 = n_ref * torch.sin(inc) * torch.sin(azi)
        kx_line = kx0 + self.mx * self.Gx                      # rcwa.py:1133-1134
        ky_line = ky0 + self.my * self.Gy
        gx, gy = torch.meshgrid(kx_line, ky_line, indexing="ij")
        self.Kx_norm_dn = gx.reshape(-1)                       # x-major flatten, rcwa.py:1138-1139
        self.Ky_norm_dn = gy.reshape(-1)
        kx, ky = self.Kx_norm_dn, self.Ky_norm_dn
        self.Kx_norm = torch.diag(kx)                          # dense diag, rcwa.py:1140-1141
        self.Ky_norm = torch.diag(ky)
        self.Vf = _v_matrix(kx, ky, _sqrt_upper(1.0 - kx * kx - ky * ky))
        if self.has_in:                                        # rcwa.py:1149-1164
            self.Vi = _v_matrix(kx, ky, _sqrt_upper(self.eps_in * self.mu_in - kx * kx - ky * ky))
            t = torch.linalg.inv(self.Vf + self.Vi)
            d = self.Vf - self.Vi
            self.Sin = [2 * (t @ self.Vi), -(t @ d), t @ d, 2 * (t @ self.Vf)]
        if self.has_out:                                       # rcwa.py:1166-1181
            self.Vo = _v_matrix(kx, ky, _sqrt_upper(self.eps_out * self.mu_out - kx * kx - ky * ky))
            t = torch.linalg.inv(self.Vf + self.Vo)
            d = self.Vf - self.Vo
            self.Sout = [2 * (t @ self.Vf), t @ d, -(t @ d), 2 * (t @ self.Vo)]

    def add_input_layer(self, eps=1.0, mu=1.0):
        self.eps_in = torch.as_tensor(eps, dtype=self.dtype)
        self.mu_in = torch.as_tensor(mu, dtype=self.dtype)
        self.has_in = True

    def add_output_layer(self, eps=1.0, mu=1.0):
        self.eps_out = torch.as_tensor(eps, dtype=self.dtype)
        self.mu_out = torch.as_tensor(mu, dtype=self.dtype)
        self.has_out = True

    # ------------------------------------------------------------------ stage 1: Fourier factorisation
    def material_conv(self, grid: torch.Tensor) -> torch.Tensor:
        """Laurent/Toeplitz convolution matrix of a sampled unit cell (rcwa.py:1183-1204):
        F = fft2(grid)/(nx*ny);  E[i,j] = F[mx_i - mx_j, my_i - my_j] with python-style
        negative index wrap-around; real and imaginary parts gathered separately, so the
        result carries the *material's* precision (rcwa.py:1196-1202)."""
        nx, ny = grid.shape
        spec = torch.fft.fft2(grid) / (nx * ny)
        gx, gy = torch.meshgrid(self.mx, self.my, indexing="ij")
        px, py = gx.reshape(-1), gy.reshape(-1)
        dx = px[:, None] - px[None, :]
        dy = py[:, None] - py[None, :]
        return torch.complex(spec.real[dx, dy], spec.imag[dx, dy])

    @staticmethod
    def _is_homogeneous(v):
        """rcwa.py:156-157."""
        return isinstance(v, (float, complex)) or v.dim() == 0 or (v.dim() == 1 and v.shape[0] == 1)

    # ------------------------------------------------------------------ stage 2: layer eigenproblem
    def add_layer(self, thickness, eps=1.0, mu=1.0):
        """rcwa.py:146-170."""
        ct, N = self.dtype, self.order_N
        eye = torch.eye(N, dtype=ct)
        he, hm = self._is_homogeneous(eps), self._is_homogeneous(mu)
        E = eps * eye if he else self.material_conv(eps)
        M = mu * eye if hm else self.material_conv(mu)
        self.eps_conv.append(E)
        self.thickness.append(thickness)
        Kx, Ky = self.Kx_norm, self.Ky_norm
        K_col = torch.vstack((Kx, Ky))
        zero = torch.zeros_like(M)
        P_base = torch.vstack((torch.hstack((zero, M)), torch.hstack((-M, zero))))     # rcwa.py:1227
        Q_base = torch.vstack((torch.hstack((zero, -E)), torch.hstack((E, zero))))     # rcwa.py:1231
        if he and hm:
            # rcwa.py:1206-1222: analytic modes, W = I
            P = P_base + (1 / eps) * (K_col @ torch.hstack((Ky, -Kx)))
            Q = Q_base + (1 / mu) * (K_col @ torch.hstack((-Ky, Kx)))
            W = torch.eye(2 * N, dtype=ct)
            kz = _sqrt_upper(eps * mu - self.Kx_norm_dn ** 2 - self.Ky_norm_dn ** 2)
            kz = torch.cat((kz, kz))
        else:
            # rcwa.py:1224-1242
            P = P_base + (K_col @ torch.linalg.inv(E)) @ torch.hstack((Ky, -Kx))
            Q = Q_base + (K_col @ torch.linalg.inv(M)) @ torch.hstack((-Ky, Kx))
            lam, W = torch.linalg.eig(P @ Q)                 # torch_eig.py:14 / rcwa.py:1238
            kz = torch.sqrt(lam)
            kz = torch.where(kz.imag < 0, -kz, kz)           # negate (not conj), rcwa.py:1241
        self.P.append(P); self.Q.append(Q)
        self.kz_norm.append(kz); self.E_eigvec.append(W)
        self._layer_smatrix(P, W, kz, thickness)

    # ------------------------------------------------------------------ stage 3a: layer S-matrix
    def _layer_smatrix(self, P, W, kz, thickness):
        """rcwa.py:1244-1281, dense as written (inv(P), inv(Vf) x4, inv(Ctmp) x2)."""
        ct, n = self.dtype, 2 * self.order_N
        Kz = torch.diag(kz)
        X = torch.diag(torch.exp(1.0j * self.omega * kz * thickness))
        V = torch.linalg.inv(P) @ (W @ Kz)                               # rcwa.py:1248,1264
        a = W + torch.linalg.inv(self.Vf) @ V
        b = (W - torch.linalg.inv(self.Vf) @ V) @ X
        C = torch.vstack((torch.hstack((a, b)), torch.hstack((b, a))))   # rcwa.py:1266-1268
        eye, zero = torch.eye(n, dtype=ct), torch.zeros((n, n), dtype=ct)
        Cf = torch.linalg.inv(C) @ torch.vstack((2 * eye, zero))         # rcwa.py:1271-1272
        Cb = torch.linalg.inv(C) @ torch.vstack((zero, 2 * eye))         # rcwa.py:1273-1274
        WX = W @ X
        S11 = WX @ Cf[:n] + W @ Cf[n:]                                   # rcwa.py:1276-1281
        S21 = W @ Cf[:n] + WX @ Cf[n:] - eye
        S12 = WX @ Cb[:n] + W @ Cb[n:] - eye
        S22 = W @ Cb[:n] + WX @ Cb[n:]
        self.layer_S.append([S11, S21, S12, S22])

    # ------------------------------------------------------------------ stage 3b: Redheffer cascade
    def _star(self, Sm, Sn):
        """rcwa.py:1283-1294 (two inverses, as written)."""
        eye = torch.eye(2 * self.order_N, dtype=self.dtype)
        t1 = torch.linalg.inv(eye - Sm[2] @ Sn[1])
        t2 = torch.linalg.inv(eye - Sn[1] @ Sm[2])
        return [Sn[0] @ (t1 @ Sm[0]),
                Sm[1] + Sm[3] @ (t2 @ (Sn[1] @ Sm[0])),
                Sn[2] + Sn[0] @ (t1 @ (Sm[2] @ Sn[3])),
                Sm[3] @ (t2 @ Sn[3])]

    def solve_global_smatrix(self):
        """rcwa.py:173-211."""
        n = 2 * self.order_N
        if self.layer_S:
            S = list(self.layer_S[0])
            for nxt in self.layer_S[1:]:
                S = self._star(S, nxt)
        else:
            # rcwa.py:186-189 uses 1-D zero vectors and broadcasting; dense zeros are equivalent
            eye = torch.eye(n, dtype=self.dtype)
            zero = torch.zeros((n, n), dtype=self.dtype)
            S = [eye, zero, zero.clone(), eye.clone()]
        if self.has_in:
            S = self._star(self.Sin, S)
        if self.has_out:
            S = self._star(S, self.Sout)
        self.S = S
        return S

    # ------------------------------------------------------------------ readout
    def _order_index(self, orders):
        """rcwa.py:1115-1122 (clamp to the truncation, x-major flat index)."""
        o = torch.as_tensor(orders, dtype=torch.int64).reshape(-1, 2).clone()
        o[:, 0].clamp_(-self.ox, self.ox)
        o[:, 1].clamp_(-self.oy, self.oy)
        return len(self.my) * (o[:, 0] + self.ox) + o[:, 1] + self.oy

    def _kz_power(self, eps, mu, evanescent, abs_when_evanescent=False):
        kzc = torch.sqrt(eps * mu - self.Kx_norm_dn ** 2 - self.Ky_norm_dn ** 2)
        ev = torch.abs(kzc.real / kzc.imag) < evanescent
        repl = torch.abs(kzc.real) if abs_when_evanescent else torch.zeros_like(kzc.real)
        k = torch.where(ev, repl, kzc.real)
        return torch.hstack((k, k))

    def S_parameters(self, orders, *, direction="forward", port="transmission", polarization="xx",
                     ref_order=(0, 0), power_norm=True, evanscent=1e-3):
        """rcwa.py:300-524."""
        N = self.order_N
        oi = self._order_index(orders)
        ri = self._order_index(ref_order)
        block = {("forward", "transmission"): 0, ("forward", "reflection"): 1,
                 ("backward", "reflection"): 2, ("backward", "transmission"): 3}[(direction, port)]
        S = self.S[block]
        kz_pairs = {0: ("out", "in"), 1: ("in", "in"), 2: ("out", "out"), 3: ("in", "out")}[block]
        if polarization in ("xx", "yx", "xy", "yy"):
            oi2 = oi + N if polarization[0] == "y" else oi
            ri2 = ri + N if polarization[1] == "y" else ri
            norm = 1.0
            if power_norm:                                              # rcwa.py:354-391
                kzs = {"in": self._kz_power(self.eps_in, self.mu_in, evanscent),
                       "out": self._kz_power(self.eps_out, self.mu_out, evanscent)}
                kx2 = torch.hstack((self.Kx_norm_dn.real, self.Kx_norm_dn.real))
                ky2 = torch.hstack((self.Ky_norm_dn.real, self.Ky_norm_dn.real))
                num_pol = kx2 if polarization[0] == "x" else ky2
                den_pol = kx2 if polarization[1] == "x" else ky2
                num_kz, den_kz = kzs[kz_pairs[0]], kzs[kz_pairs[1]]
                norm = torch.sqrt((1 + (num_pol[oi2] / num_kz[oi2]) ** 2) /
                                  (1 + (den_pol[ri2] / den_kz[ri2]) ** 2))
                norm = norm * torch.sqrt(num_kz[oi2] / den_kz[ri2])
            out = S[oi2, ri2] * norm
            out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
            return torch.where(torch.isnan(out), torch.zeros_like(out), out)
        # ---- ps polarisation, rcwa.py:410-521
        osign, rsign = {0: (1, 1), 1: (-1, 1), 2: (1, -1), 3: (-1, -1)}[block]
        em = {"in": self.eps_in * self.mu_in, "out": self.eps_out * self.mu_out}
        ok2, rk2 = em[kz_pairs[0]], em[kz_pairs[1]]

        def angles(idx, k2, sign):
            kx, ky = self.Kx_norm_dn[idx], self.Ky_norm_dn[idx]
            kt = torch.sqrt(kx ** 2 + ky ** 2)
            kzc = torch.sqrt(k2 - kx ** 2 - ky ** 2)
            kz = sign * torch.abs(kzc.real)
            ev = torch.abs(kzc.real / kzc.imag) < evanscent
            return torch.atan2(kt.real, kz), torch.atan2(ky.real, kx.real), ev

        o_inc, o_azi, o_ev = angles(oi, ok2, osign)
        r_inc, r_azi, r_ev = angles(ri, rk2, rsign)
        z = lambda t: torch.where(o_ev, torch.zeros_like(t), t)
        xx, xy = z(S[oi, ri]), z(S[oi, ri + N])
        yx, yy = z(S[oi + N, ri]), z(S[oi + N, ri + N])
        if bool(r_ev):
            return torch.zeros_like(xx)
        co, so, ci = torch.cos(o_azi), torch.sin(o_azi), torch.cos(o_inc)
        cr, sr, cri = torch.cos(r_azi), torch.sin(r_azi), torch.cos(r_inc)
        if polarization == "pp":
            out = (co / ci) * cri * cr * xx + (so / ci) * cri * cr * yx + (co / ci) * cri * sr * xy + (so / ci) * cri * sr * yy
        elif polarization == "ps":
            out = (co / ci) * (-sr) * xx + (so / ci) * (-sr) * yx + (co / ci) * cr * xy + (so / ci) * cr * yy
        elif polarization == "sp":
            out = -so * cri * cr * xx + co * cri * cr * yx - so * cri * sr * xy + co * cri * sr * yy
        else:  # ss
            out = -so * (-sr) * xx + co * (-sr) * yx - so * cr * xy + co * cr * yy
        norm = 1.0
        if power_norm:                                                   # rcwa.py:487-516
            kzs = {"in": self._kz_power(self.eps_in, self.mu_in, evanscent),
                   "out": self._kz_power(self.eps_out, self.mu_out, evanscent, abs_when_evanescent=True)}
            norm = torch.sqrt(kzs[kz_pairs[0]][oi] / kzs[kz_pairs[1]][ri])
        out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
        out = torch.where(torch.isnan(out), torch.zeros_like(out), out)
        return out * norm


# ---------------------------------------------------------------------- input builders (shared by
# tests, bench and the golden generator; geometry itself is out of scope)
def rectangle_grid(Lx, Ly, nx, ny, Wx, Wy, Cx, Cy, theta=0.0, edge_sharpness=1000.0,
                   dtype=torch.float64):
    """Sigmoid-edged rectangle on the cell-centred grid (torcwa/geometry.py:42-47, :86-100)."""
    x = (Lx / nx) * (torch.arange(nx, dtype=dtype) + 0.5)
    y = (Ly / ny) * (torch.arange(ny, dtype=dtype) + 0.5)
    X, Y = torch.meshgrid(x, y, indexing="ij")
    th = torch.as_tensor(theta, dtype=dtype)
    u = ((X - Cx) * torch.cos(th) + (Y - Cy) * torch.sin(th)) / (Wx / 2.0)
    v = (-(X - Cx) * torch.sin(th) + (Y - Cy) * torch.cos(th)) / (Wy / 2.0)
    return torch.sigmoid(edge_sharpness * (1.0 - torch.maximum(u.abs(), v.abs())))


def solve_point(freq, order, L, layers, *, dtype=torch.complex128, eps_in=None, eps_out=None,
                inc_ang=0.0, azi_ang=0.0):
    """Convenience: one design point through the whole path.  ``layers`` = [(thickness, eps), ...]."""
    sim = OracleSim(freq=freq, order=order, L=L, dtype=dtype)
    if eps_in is not None:
        sim.add_input_layer(eps=eps_in)
    if eps_out is not None:
        sim.add_output_layer(eps=eps_out)
    sim.set_incident_angle(inc_ang, azi_ang)
    for d, e in layers:
        sim.add_layer(d, e)
    sim.solve_global_smatrix()
    return sim


def eig_backward(eigval, eigvec, grad_eigval, grad_eigvec, broadening_parameter=1e-10):
    """Restatement of Eig.backward (torcwa/torch_eig.py:19-44) for one matrix, complex input:
    s_ij = lambda_j - lambda_i (:25); F = conj(s) / (|s|^2 + delta), delta = broadening parameter or the
    smallest denormal of the precision (:28-33); diag(F) = 0 (:35-36);
    grad = X^-H (diag(g_lambda) + conj(F) o (X^H g_X)) X^H (:37-40).  TEST INFRASTRUCTURE (see module header)."""
    s = eigval.unsqueeze(-2) - eigval.unsqueeze(-1)
    if broadening_parameter is not None:
        delta = broadening_parameter
    else:
        delta = 1.4e-45 if s.dtype == torch.complex64 else 4.9e-324
    F = torch.conj(s) / (torch.abs(s) ** 2 + delta)
    idx = torch.arange(F.shape[-1])
    F[idx, idx] = 0.0
    XH = torch.transpose(torch.conj(eigvec), -2, -1)
    tmp = torch.conj(F) * torch.matmul(XH, grad_eigvec)
    return torch.matmul(torch.matmul(torch.inverse(XH), torch.diag(grad_eigval) + tmp), XH)
