"""Sources, mode coefficients and field reconstruction (scope row f1; reference torcwa/rcwa.py:526-1112,
mode coefficients :1266-1274, their propagation through the cascade :1296-1304).

The fused forward path (rcwa_layer_smatrix / rcwa_redheffer) produces S-matrices only.  When a field is
asked for, the mode coefficients are built once, lazily, from what the simulation stored per layer
(eigenvectors W, Q, kz, convolution matrices -- unbatched simulations keep them by default):

  layer:    V = Q W Kz^-1 (H-field modes),  M+- = W (1 +- X) + Vf^-1 V (1 -+ X),  X = exp(i w kz d);
            with Ip = M+^-1, Im = M-^-1 the reference's Cf = Ctmp^-1 [2I; 0], Cb = Ctmp^-1 [0; 2I] are
            Cf = [Ip + Im; Ip - Im],  Cb = [Ip - Im; Ip + Im]            
  cascade:  the same left fold as solve_global_smatrix, carrying the coefficient matrices (star_with_modes).

Dense O(n^3) work (products, inverses) runs on the CUDA GEMM / LU kernels through `_lib`; the rest is O(n^2)
torch glue.  Fields are physical quantities: they do not depend on the normalisation or order of the
eigenvectors, so they agree with the reference although W comes from a different eigensolver.

Batched simulations (needs store_intermediates=True): every design point is processed through the same,
unbatched code on a one-point view of the simulation (`_PointView`), and the planes are stacked to [B, ...].
"""
import warnings

import torch

from . import _lib
from ._bd import sqrt_upper

_C = torch.complex128


def _mm(A, B, opb="N"):
    """Dense product on the CUDA GEMM ([m,k] x [k,n], complex128)."""
    return _lib.zgemm(A[None].contiguous(), B[None].contiguous(), "N", opb)[0]


def _inv(A):
    X, info = _lib.inverse(A[None].contiguous())
    if int(info.abs().max()) != 0:
        raise torch.linalg.LinAlgError('singular matrix while building mode coefficients')
    return X[0]


def _dense_bd(d4):
    """[4,N] four diagonals -> dense [2N,2N]."""
    a, b, c, d = (torch.diag(d4[k]) for k in range(4))
    return torch.cat((torch.cat((a, b), 1), torch.cat((c, d), 1)), 0)


# ------------------------------------------------------------------------------------------ batched simulations
class _PointView:
    """One design point of a batched simulation, presented as an unbatched simulation: every per-point tensor is
    sliced to [b:b+1], everything else is delegated.  The field code below only ever sees unbatched objects."""
    _SLICED = ('_kx', '_ky', '_Vf', '_Vf_inv', '_Vi', '_Vo', '_omega64')

    def __init__(self, sim, b):
        object.__setattr__(self, '_sim', sim)
        object.__setattr__(self, '_pt', b)
        self._batched = False
        self._modes_ready = False
        self._modes_src = [{k: (v if v is None else v[b:b + 1]) for k, v in rec.items()} for rec in sim._modes_src]
        self._layers = [[x[b:b + 1] for x in lay] for lay in sim._layer_pairs_dense()]     # symmetry-block layers in the original basis
        self._S = [x[b:b + 1] for x in sim._S]
        if hasattr(sim, '_Sin'):
            self._Sin = [x[b:b + 1] for x in sim._Sin]
        if hasattr(sim, '_Sout'):
            self._Sout = [x[b:b + 1] for x in sim._Sout]
        self.thickness = [t.reshape(-1)[b] if (isinstance(t, torch.Tensor) and t.numel() == sim._B and sim._B > 1) else t
                          for t in sim.thickness]

    def __getattr__(self, name):
        sim, b = object.__getattribute__(self, '_sim'), object.__getattribute__(self, '_pt')
        v = getattr(sim, name)
        if name in _PointView._SLICED:
            return v[b:b + 1]
        return v

    def _b(self, v):
        return self._sim._b(v)[self._pt:self._pt + 1]

    def _layer_pairs_dense(self):
        return self._layers                     # this point's slices (already in the original basis)

    def _pub(self, t):
        return t.to(self._sim._dtype)[0]

    def _matching_indices(self, orders):
        return self._sim._matching_indices(orders)


def _stack_points(sim, fn):
    """Run fn on every point's view and stack the six field planes to [B, ...]."""
    if not hasattr(sim, '_source_spec'):
        raise RuntimeError('define a source first (source_planewave / source_fourier)')
    outs = []
    for b in range(sim._B):
        view = sim._views.get(b)
        if view is None or view._sim_solve_id != id(sim._S):
            view = _PointView(sim, b)
            view._sim_solve_id = id(sim._S)
            source_fourier(view, *sim._source_spec)
            sim._views[b] = view
        outs.append(fn(view))
    E = [torch.stack([o[0][k] for o in outs]) for k in range(3)]
    H = [torch.stack([o[1][k] for o in outs]) for k in range(3)]
    return E, H


# ------------------------------------------------------------------------------------------ sources
def source_fourier(sim, amplitude, orders, direction, notation):
    amplitude = torch.as_tensor(amplitude, dtype=sim._dtype, device=sim._device).reshape([-1, 2])
    orders = torch.as_tensor(orders, dtype=torch.int64, device=sim._device).reshape([-1, 2])
    if direction in ['f', 'forward']:
        direction = 'forward'
    elif direction in ['b', 'backward']:
        direction = 'backward'
    else:
        warnings.warn('Invalid source direction. Set as forward.', UserWarning)
        direction = 'forward'
    if notation not in ['xy', 'ps']:
        warnings.warn('Invalid amplitude notation. Set as xy notation.', UserWarning)
        notation = 'xy'
    if sim._batched and not isinstance(sim, _PointView):
        # one source specification for the whole sweep; the ps -> xy rotation depends on each point's wavevectors,
        # so the amplitude vector is built per point (on its view) when a field is asked for
        sim.source_direction = direction
        sim._source_spec = (amplitude, orders, direction, notation)
        sim._views = {}
        return
    idx = sim._matching_indices(orders)
    N = sim.order_N
    sim.source_direction = direction
    E_i = torch.zeros(2 * N, dtype=_C, device=sim._device)
    E_i[idx] = amplitude[:, 0].to(_C)
    E_i[idx + N] = amplitude[:, 1].to(_C)
    if notation == 'ps':
        # (p, s) amplitudes of every order -> (x, y): rotation by the order's own propagation angles (rcwa.py:575-594)
        eps, mu, sign = (sim.eps_in, sim.mu_in, 1.0) if direction == 'forward' else (sim.eps_out, sim.mu_out, -1.0)
        kx, ky = sim._kx[0], sim._ky[0]
        kt = torch.sqrt(kx ** 2 + ky ** 2)
        kz = sign * torch.abs(torch.sqrt(sim._b(eps)[0] * sim._b(mu)[0] - kx ** 2 - ky ** 2).real)
        inc = torch.atan2(kt.real, kz)
        azi = torch.atan2(ky.real, kx.real)
        p, s_ = E_i[:N].clone(), E_i[N:].clone()
        E_i = torch.cat((torch.cos(inc) * torch.cos(azi) * p - torch.sin(azi) * s_,
                         torch.cos(inc) * torch.sin(azi) * p + torch.cos(azi) * s_))
    sim._E_i = E_i
    sim.E_i = E_i.to(sim._dtype).reshape(-1, 1)


# ------------------------------------------------------------------------------------------ mode coefficients
def _layer_modes(sim, rec):
    """(W, V, kz, Cf, Cb, eps_inv, mu_inv) of one stored layer; matrices [n,n] / [2n,n], complex128."""
    N = sim.order_N
    n = 2 * N
    dev = sim._device
    kz = rec['kz'][0]
    eye = torch.eye(n, dtype=_C, device=dev)
    if rec['W'] is None:                       # homogeneous layer: W = I, Q four diagonals (rcwa.py:1206-1222)
        W = eye
        V = _dense_bd(rec['Q'][0]) / kz[None, :]
        eps_inv = (1 / rec['E'][0]) * torch.eye(N, dtype=_C, device=dev)
        mu_inv = (1 / rec['M'][0]) * torch.eye(N, dtype=_C, device=dev)
    else:
        W = rec['W'][0]
        V = _mm(rec['Q'][0], W) / kz[None, :]
        eps_inv = _inv(rec['E'][0])
        M = rec['M']
        mu_inv = _inv(M[0]) if M.dim() == 3 else (1 / M[0]) * torch.eye(N, dtype=_C, device=dev)
    a, b, c, d = (sim._Vf_inv[0, k][:, None] for k in range(4))
    Bm = torch.cat((a * V[:N] + b * V[N:], c * V[:N] + d * V[N:]), 0)                  # Vf^-1 V
    X = torch.exp(1j * (rec['omega'][0] * rec['thick'][0]) * kz)[None, :]
    Ip = _inv(W * (1 + X) + Bm * (1 - X))
    Im = _inv(W * (1 - X) + Bm * (1 + X))
    Cf = torch.cat((Ip + Im, Ip - Im), 0)
    Cb = torch.cat((Ip - Im, Ip + Im), 0)
    return W, V, kz, Cf, Cb, eps_inv, mu_inv


def _star_with_modes(Sm, Sn, Cm, Cn):
    """Redheffer star product carrying the mode-coefficient matrices of the layers on either side
    (rcwa.py:1283-1306): S = Sm (*) Sn; every C of the left stack picks up the reflection from the right stack
    and vice versa."""
    n = Sm[0].shape[0]
    eye = torch.eye(n, dtype=_C, device=Sm[0].device)
    t1 = _inv(eye - _mm(Sm[2], Sn[1]))
    t2 = _inv(eye - _mm(Sn[1], Sm[2]))
    A1 = _mm(t1, Sm[0])                        # forward wave entering the right stack
    A2 = _mm(t1, _mm(Sm[2], Sn[3]))
    B1 = _mm(t2, _mm(Sn[1], Sm[0]))            # wave reflected back into the left stack
    B2 = _mm(t2, Sn[3])
    S = [_mm(Sn[0], A1), Sm[1] + _mm(Sm[3], B1), Sn[2] + _mm(Sn[0], A2), _mm(Sm[3], B2)]
    C = [[], []]
    for cf, cb in zip(Cm[0], Cm[1]):
        C[0].append(cf + _mm(cb, B1))
        C[1].append(_mm(cb, B2))
    for cf, cb in zip(Cn[0], Cn[1]):
        C[0].append(_mm(cf, A1))
        C[1].append(cb + _mm(cf, A2))
    return S, C


def ensure_modes(sim):
    """Build H_eigvec, Cf, Cb per layer and the propagated coefficient lists sim.C (once per solve)."""
    if sim._modes_ready:
        return
    if len(sim._modes_src) != sim.layer_N:
        raise RuntimeError('fields need the per-layer intermediates: construct the simulation with store_intermediates=True')
    if not hasattr(sim, '_S'):
        raise RuntimeError('call solve_global_smatrix() before asking for fields')
    n = 2 * sim.order_N
    dev = sim._device
    sim._modes = [_layer_modes(sim, rec) for rec in sim._modes_src]
    sim.H_eigvec = [sim._pub(m[1][None]) for m in sim._modes]
    sim.Cf = [sim._pub(m[3][None]) for m in sim._modes]
    sim.Cb = [sim._pub(m[4][None]) for m in sim._modes]
    layers = sim._layer_pairs_dense()       # the mode coefficients are carried in the original basis (symmetry blocks unprojected)
    if sim.layer_N > 0:
        s11, s21 = (x[0] for x in layers[0])
        S = [s11, s21, s21, s11]
        C = [[sim._modes[0][3]], [sim._modes[0][4]]]
        for l in range(1, sim.layer_N):
            n11, n21 = (x[0] for x in layers[l])
            S, C = _star_with_modes(S, [n11, n21, n21, n11], C, [[sim._modes[l][3]], [sim._modes[l][4]]])
    else:
        eye = torch.eye(n, dtype=_C, device=dev)
        zero = torch.zeros((n, n), dtype=_C, device=dev)
        S, C = [eye, zero, zero.clone(), eye.clone()], [[], []]
    if hasattr(sim, 'Sin'):
        S, C = _star_with_modes([_dense_bd(s[0]) for s in sim._Sin], S, [[], []], C)
    if hasattr(sim, 'Sout'):
        S, C = _star_with_modes(S, [_dense_bd(s[0]) for s in sim._Sout], C, [[], []])
    sim._C = C
    sim.C = [[sim._pub(c[None]) for c in C[0]], [sim._pub(c[None]) for c in C[1]]]
    sim._modes_ready = True


# ------------------------------------------------------------------------------------------ Fourier-domain fields
def _half_space_coefficients(sim, which, z_prop):
    """Fourier coefficients [6, N, nz] (Ex, Ey, Ez, Hx, Hy, Hz) in the input (which = -1) or output half space at
    distances z_prop [nz] from its boundary (rcwa.py:640-696)."""
    N = sim.order_N
    kx, ky = sim._kx[0], sim._ky[0]
    E_i = sim._E_i
    S = [s[0] for s in sim._S]
    fwd = sim.source_direction == 'forward'
    if which == -1:
        eps, mu = sim._b(sim.eps_in)[0], sim._b(sim.mu_in)[0]
        Vh = sim._Vi[0] if hasattr(sim, '_Vi') else sim._Vf[0]
        kz = torch.sqrt(eps * mu - kx ** 2 - ky ** 2)
        kz = torch.where(kz.imag > 0, kz.conj(), kz)                # the reference's branch for z < 0 (:651)
        z_prop = torch.clamp(z_prop, max=0.0)
        up = E_i if fwd else torch.zeros_like(E_i)
        down = S[1] @ E_i if fwd else S[3] @ E_i
    else:
        # the reference tests hasattr(self, 'eps_in') here, which is always true (:658-659): eps_out is used
        eps, mu = sim._b(sim.eps_out)[0], sim._b(sim.mu_out)[0]
        Vh = sim._Vo[0] if hasattr(sim, '_Vo') else sim._Vf[0]
        kz = sqrt_upper(eps * mu - kx ** 2 - ky ** 2)
        z_prop = torch.clamp(z_prop, min=0.0)
        up = S[0] @ E_i if fwd else S[2] @ E_i
        down = torch.zeros_like(E_i) if fwd else E_i
    kz2 = torch.cat((kz, kz))
    ph = torch.exp(1j * sim._omega64[0] * kz2[:, None] * z_prop[None, :])          # [2N, nz]
    Ep = up[:, None] * ph
    Em = down[:, None] * ph.conj()
    a, b, c, d = (Vh[k][:, None] for k in range(4))

    def applyV(F):
        return torch.cat((a * F[:N] + b * F[N:], c * F[:N] + d * F[N:]), 0)
    Hp, Hm = applyV(Ep), -applyV(Em)
    Ex, Ey = Ep[:N] + Em[:N], Ep[N:] + Em[N:]
    Hx, Hy = Hp[:N] + Hm[:N], Hp[N:] + Hm[N:]
    Hz = (kx[:, None] * Ey - ky[:, None] * Ex) / mu
    Ez = (ky[:, None] * Hx - kx[:, None] * Hy) / eps
    return torch.stack((Ex, Ey, Ez, Hx, Hy, Hz))


def _layer_coefficients(sim, l, z_prop):
    """Fourier coefficients [6, N, nz] inside layer l at heights z_prop [nz] above its lower boundary (rcwa.py:712-760):
    E = W (e^{i w kz z} c+) + W (e^{i w kz (d - z)} c-),  H = V (...) - V (...),  Ez, Hz from the curl equations."""
    N = sim.order_N
    n = 2 * N
    W, V, kz, _, _, eps_inv, mu_inv = sim._modes[l]
    Cl = sim._C[0][l] if sim.source_direction == 'forward' else sim._C[1][l]
    c = Cl @ sim._E_i
    cp, cm = c[:n], c[n:]
    rec = sim._modes_src[l]
    om, d = rec['omega'][0], rec['thick'][0]
    Ap = torch.exp(1j * om * kz[:, None] * z_prop[None, :]) * cp[:, None]           # [n, nz]
    Am = torch.exp(1j * om * kz[:, None] * (d - z_prop)[None, :]) * cm[:, None]
    Exy = _mm(W, Ap + Am)
    Hxy = _mm(V, Ap - Am)
    kx, ky = sim._kx[0][:, None], sim._ky[0][:, None]
    Ex, Ey, Hx, Hy = Exy[:N], Exy[N:], Hxy[:N], Hxy[N:]
    Hz = mu_inv @ (kx * Ey - ky * Ex)
    Ez = eps_inv @ (ky * Hx - kx * Hy)
    return torch.stack((Ex, Ey, Ez, Hx, Hy, Hz))


def _layer_of(sim, z_axis):
    """Layer index of every z (rcwa.py:622-634): -1 below 0, l for zm[l] <= z <= zp[l] (a boundary belongs to the
    layer below it), layer_N above the stack."""
    if sim.layer_N > 0:
        th = torch.stack([torch.as_tensor(t, dtype=torch.float64, device=sim._device).reshape(()) for t in sim.thickness])
    else:
        th = torch.zeros(0, dtype=torch.float64, device=sim._device)
    zp = torch.cumsum(th, 0)
    zm = torch.cat((torch.zeros(1, dtype=torch.float64, device=sim._device), zp[:-1])) if sim.layer_N > 0 else zp
    num = torch.zeros(len(z_axis), dtype=torch.int64, device=sim._device)
    num[z_axis < 0.] = -1
    for b in range(len(zp)):
        num[z_axis > zp[b]] += 1
    return num, zm, zp


def _coefficients_along_z(sim, z_axis):
    """[6, N, nz] for an arbitrary set of z, grouped by layer."""
    ensure_modes(sim)
    if not hasattr(sim, '_E_i'):
        raise RuntimeError('define a source first (source_planewave / source_fourier)')
    z = z_axis.to(device=sim._device, dtype=torch.float64).reshape(-1)
    num, zm, zp = _layer_of(sim, z)
    out = torch.zeros((6, sim.order_N, len(z)), dtype=_C, device=sim._device)
    for l in torch.unique(num).tolist():
        sel = torch.nonzero(num == l).reshape(-1)
        if l == -1:
            out[:, :, sel] = _half_space_coefficients(sim, -1, z[sel])
        elif l == sim.layer_N:
            top = zp[-1] if len(zp) > 0 else torch.zeros((), dtype=torch.float64, device=sim._device)
            out[:, :, sel] = _half_space_coefficients(sim, sim.layer_N, z[sel] - top)
        else:
            out[:, :, sel] = _layer_coefficients(sim, l, z[sel] - zm[l])
    return out


# ------------------------------------------------------------------------------------------ spatial synthesis
def _finish(sim, F):
    F = [f.to(sim._dtype) for f in F]
    return [F[0], F[1], F[2]], [F[3], F[4], F[5]]


def field_plane(sim, plane, t_axis, z_axis, other):
    """xz (other = y) or yz (other = x) cut: sum over orders of coefficient(z) * exp(i w (Kx x + Ky y))."""
    if type(t_axis) != torch.Tensor or type(z_axis) != torch.Tensor:
        warnings.warn('%s and z axis must be torch.Tensor type. Return None.' % plane[0], UserWarning)
        return None
    if sim._batched and not isinstance(sim, _PointView):
        return _stack_points(sim, lambda v: field_plane(v, plane, t_axis, z_axis, other))
    coef = _coefficients_along_z(sim, z_axis)                                         # [6, N, nz]
    t = t_axis.to(device=sim._device, dtype=torch.float64).reshape(-1, 1)
    om = sim._omega64[0]
    kx, ky = sim._kx[0][None, :], sim._ky[0][None, :]
    o = float(other)
    phase = torch.exp(1j * om * (kx * t + ky * o)) if plane == 'xz' else torch.exp(1j * om * (kx * o + ky * t))   # [nt, N]
    return _finish(sim, [_mm(phase, coef[k]) for k in range(6)])


def field_xy(sim, layer_num, x_axis, y_axis, z_prop):
    if type(layer_num) != int:
        warnings.warn('Parameter "layer_num" must be int type. Return None.', UserWarning)
        return None
    if layer_num < -1 or layer_num > sim.layer_N:
        warnings.warn('Layer number is out of range. Return None.', UserWarning)
        return None
    if type(x_axis) != torch.Tensor or type(y_axis) != torch.Tensor:
        warnings.warn('x and y axis must be torch.Tensor type. Return None.', UserWarning)
        return None
    if sim._batched and not isinstance(sim, _PointView):
        return _stack_points(sim, lambda v: field_xy(v, layer_num, x_axis, y_axis, z_prop))
    ensure_modes(sim)
    if not hasattr(sim, '_E_i'):
        raise RuntimeError('define a source first (source_planewave / source_fourier)')
    zp = torch.as_tensor(z_prop, dtype=torch.float64, device=sim._device).reshape(1)
    if layer_num == -1 or layer_num == sim.layer_N:
        coef = _half_space_coefficients(sim, layer_num, zp)[:, :, 0]                  # [6, N]
    else:
        coef = _layer_coefficients(sim, layer_num, zp)[:, :, 0]
    x = x_axis.to(device=sim._device, dtype=torch.float64).reshape(-1, 1)
    y = y_axis.to(device=sim._device, dtype=torch.float64).reshape(-1, 1)
    om = sim._omega64[0]
    ex = torch.exp(1j * om * sim._kx[0][None, :] * x)                                  # [nx, N]
    ey = torch.exp(1j * om * sim._ky[0][None, :] * y)                                  # [ny, N]
    return _finish(sim, [_mm(ex * coef[k][None, :], ey, "T") for k in range(6)])
