"""Host side of the B200-native RCWA solver: the reference's public object, batched.

Mirrors ``torcwa.rcwa`` (torcwa/rcwa.py:7-524 of the reference) -- same constructor, same call
order (add_input_layer -> set_incident_angle -> add_layer... -> solve_global_smatrix ->
S_parameters), same attribute names, same warnings -- but every heavy step is a call into
librcwa_b200.so (C ABI, include/rcwa_b200.h) and every tensor carries a leading batch dimension of
independent design points (wavelength x geometry), which the reference does not have
(SURVEY.md 7.2):

  * ``freq`` may be a [B] tensor, ``eps`` of a layer may be [nx,ny] (shared) or [B,nx,ny],
    ``thickness`` / angles / half-space permittivities may be scalars or [B];
  * with scalar ``freq`` every public attribute has exactly the reference's shape.

Arithmetic: all device work is complex128 ("fp64 internals behind the c64 API", SURVEY.md finding
5); for ``dtype=torch.complex64`` inputs are widened on entry and public results rounded on exit.
torch is used for device memory, streams and O(N) per-order bookkeeping only; there is no torch
fallback for the dense stages -- if the CUDA library is missing the constructor raises.
"""
import os
import warnings
import weakref

import torch

from . import _lib
from . import autodiff
from . import symmetry
from ._bd import bd, bd_diag, bd_inv, bd_mul, sqrt_upper, v_matrix

# The reference's pi is mistyped (torcwa/rcwa.py:5); omega = 2*pi*freq enters every layer phase, so
# parity at 1e-10 requires the same constant.
pi = 3.141592652589793

_C = torch.complex128

# Test hook only: tests/fake_lib.py swaps `_lib` for a torch-on-CPU double of the C ABI to exercise
# this file's host logic without a GPU, and sets this flag so the constructor accepts a CPU device.
# With the real `_lib` a non-CUDA tensor is rejected by every wrapper, so there is no CPU fallback.
_TEST_ALLOW_NON_CUDA = False


def vf_inverse_diagonals(kx, ky):
    """Four diagonals of Vf^-1 (free-space E->H matrix, torcwa/rcwa.py:1143-1147) as [B,4,N]."""
    return bd_inv(v_matrix(kx, ky, sqrt_upper(1.0 - kx * kx - ky * ky))).contiguous()


class rcwa:
    def __init__(self, freq, order, L, *,
                 dtype=torch.complex64,
                 device=None,
                 stable_eig_grad=True,
                 avoid_Pinv_instability=False,
                 max_Pinv_instability=0.005,
                 store_intermediates=None,
                 gemm_digits=None,
                 pipeline=None,
                 symmetry_reduction=None):
        """Same parameters as the reference (torcwa/rcwa.py:9-35).  ``store_intermediates``
        (new): keep per-layer P, Q, eigenvectors, convolution matrices as attributes
        (default: only for unbatched sims, where the reference keeps them).
        ``gemm_digits`` (new): engine of the dense products of the S-matrix stage (layer S-matrix, star
        products, their triangular solves): 0 = fp64 tensor pipe (DMMA), 2..8 = tcgen05 int8-digit GEMM with
        that many 8-bit digits per number (include/rcwa_b200.h: rcwa_zgemm_tc_batched).  Default: 7 for
        complex64 simulations, 0 for complex128.  Why 7 and not fewer: the stage amplifies a product's error by up
        to ~1e6 into the far-evanescent entries of the S blocks (measured on config 4's sweep corners at order 15:
        5 digits -> 2e-4 on a block column, 6 -> 1e-6, 7 and 8 -> the float32 input noise 8e-8), so 7 keeps the
        complex64 gate of 1e-4 with four orders of margin at 1.9x the fp64 kernel's speed (5 digits: 2.6x);
        the environment variable RCWA_B200_GEMM_DIGITS overrides the complex64 default.  The eigensolver always
        runs in fp64 (SURVEY.md finding 5).
        ``symmetry_reduction`` (new): True / False / None (= environment RCWA_B200_SYMMETRY, default on).  When a
        patterned layer's cell has a mirror plane in x and / or y, or an inversion centre, and the incidence respects
        it (kx0 = 0 and / or ky0 = 0), the layer eigenproblem and the S-matrix cascade are solved in the
        symmetry-adapted basis as 2 or 4 independent blocks of ~n/2 or ~n/4 (torcwa_b200/symmetry.py): the same
        numbers to round-off at 1/4 ... 1/16 of the dense work.  Detected from the convolution matrix itself
        (tolerance 1e-11); any layer without the symmetry sends the whole stack through the general path.
        ``pipeline`` (new): number of sub-batches a batched simulation is run as (default 1 = off; environment
        RCWA_B200_PIPELINE; needs >= 8 points per sub-batch).  Sub-batches run on their own CUDA streams, driven by
        their own host threads, staggered so that the Hessenberg reduction and the S-matrix stage of one sub-batch run
        under the QR iteration of another.  Measured on B200 (profiles/r2_pipeline.md): a LOSS at every batch size --
        the QR phase is bound by the fp64 tensor pipe (its K = 64 update GEMMs), not idle, so two concurrent
        eigensolvers only share it -- hence off by default.  Results do not depend on it: every design point's
        arithmetic is independent of the batch it is solved in (tests/test_gpu_parity.py checks bit identity)."""
        if dtype != torch.complex64 and dtype != torch.complex128:
            warnings.warn('Invalid simulation data type. Set as torch.complex64.', UserWarning)
            dtype = torch.complex64
        self._dtype = dtype
        self._rdtype = torch.float32 if dtype == torch.complex64 else torch.float64
        if device is None:
            device = torch.device('cuda')
        self._device = torch.device(device)
        if gemm_digits is None:
            gemm_digits = int(os.environ.get('RCWA_B200_GEMM_DIGITS', '7')) if dtype == torch.complex64 else 0
        self._digits = int(gemm_digits) if 2 <= int(gemm_digits) <= 8 else 0
        if self._device.type != 'cuda' and not _TEST_ALLOW_NON_CUDA:
            raise RuntimeError('torcwa_b200 runs on CUDA devices only (no CPU path); got device=%s' % device)
        _lib.load()   # fail loudly, now, if the CUDA library is missing

        self.stable_eig_grad = True if stable_eig_grad else False
        if not stable_eig_grad:
            # accepted for drop-in compatibility; the broadened backward (torch_eig.py:28-33) is the only one implemented
            warnings.warn('torcwa_b200: stable_eig_grad=False is not honoured -- gradients of the eigendecomposition always use '
                          'the Lorentzian-broadened formula (torcwa.Eig.broadening_parameter)', UserWarning)
        if avoid_Pinv_instability is True:
            self.avoid_Pinv_instability = True
            self.max_Pinv_instability = max_Pinv_instability
            self.Pinv_instability, self.Qinv_instability = [], []
        else:
            self.avoid_Pinv_instability = False
            self.max_Pinv_instability = None
            self.Pinv_instability = self.Qinv_instability = None

        f = torch.as_tensor(freq)
        self._batched = f.dim() >= 1 and f.numel() > 1
        self._B = f.numel() if self._batched else 1
        self.freq = torch.as_tensor(freq, dtype=self._dtype, device=self._device)
        self.omega = 2 * pi * freq                     # raw argument, as the reference (rcwa.py:61)
        om = self.omega
        om = om.to(self._device).real.to(torch.float64) if isinstance(om, torch.Tensor) else torch.tensor(om, dtype=torch.float64, device=self._device)
        self._omega64 = om.reshape(-1)
        self._freq128 = self.freq.to(_C).reshape(-1)   # widened *after* the cast to the sim dtype (rcwa.py:60)
        self.L = L
        self.order = order
        self.order_x = torch.arange(-order[0], order[0] + 1, dtype=torch.int64, device=self._device)
        self.order_y = torch.arange(-order[1], order[1] + 1, dtype=torch.int64, device=self._device)
        self.order_N = len(self.order_x) * len(self.order_y)
        self.Gx_norm, self.Gy_norm = 1 / (L[0] * self.freq), 1 / (L[1] * self.freq)
        self._Gx = 1 / (L[0] * self._freq128)
        self._Gy = 1 / (L[1] * self._freq128)

        one = torch.tensor(1., dtype=self._dtype, device=self._device)
        self.eps_in, self.mu_in, self.eps_out, self.mu_out = one, one.clone(), one.clone(), one.clone()
        self._store = (not self._batched) if store_intermediates is None else bool(store_intermediates)

        self.layer_N = 0
        self.thickness = []
        self.eps_conv, self.mu_conv = [], []
        self.P, self.Q = [], []
        self.kz_norm, self.E_eigvec, self.H_eigvec = [], [], []
        self.Cf, self.Cb = [], []
        self.layer_S11, self.layer_S21, self.layer_S12, self.layer_S22 = [], [], [], []
        self._modes_src = []       # per layer (only when intermediates are stored): what the mode coefficients / fields need
        self._modes_ready = False
        self._diff = False         # a layer asked for gradients: the cascade runs on the differentiable primitives
        self._layers = []          # internal: per layer [S11, S21] complex128 [B,n,n]
        self.eig_info = []         # per patterned layer: int32 [B] status of the eigensolver
        self._status = []          # (what, int32 [B] device tensor) of every factorisation / eigensolve: checked lazily
        self._gate = None          # (wait flag, wait event, signal flag, signal event): stagger of pipelined sub-batches
        if symmetry_reduction is None:
            symmetry_reduction = os.environ.get('RCWA_B200_SYMMETRY', '1') != '0'
        self._sym_on = bool(symmetry_reduction)
        self._sym = None           # symmetry.Basis shared by the block layers of this stack
        self._sym_G = None         # per block: Vf^-1 in adapted coordinates
        self._pad_err = []         # per extended block: max |coupling| to the decoupled extension entries; must be exactly 0
        self._kz_min = []          # per patterned layer: min |kz| / max |kz| over the batch (device scalar), read with the status words

        # ---- pipelined sub-batches (children); the parent keeps the O(N) per-order state and delegates the dense stages
        self._children = None
        if pipeline is None:
            pipeline = int(os.environ.get('RCWA_B200_PIPELINE', '1'))
        k = int(pipeline) if (self._batched and self._B >= 16 and not self._store and self._device.type == 'cuda') else 1
        if k > 1:
            k = min(k, self._B // 8)
            bounds = [self._B * i // k for i in range(k + 1)]
            self._slices = [slice(bounds[i], bounds[i + 1]) for i in range(k)]
            fr = torch.as_tensor(freq).reshape(-1)
            with torch.cuda.device(self._device):
                self._streams = [torch.cuda.Stream(device=self._device) for _ in range(k)]
            self._children = [rcwa(fr[sl], order, L, dtype=dtype, device=self._device, stable_eig_grad=stable_eig_grad,
                                   avoid_Pinv_instability=avoid_Pinv_instability, max_Pinv_instability=max_Pinv_instability,
                                   store_intermediates=False, gemm_digits=self._digits, pipeline=1,
                                   symmetry_reduction=self._sym_on) for sl in self._slices]

    # ------------------------------------------------------------------ pipelined sub-batches
    def _part(self, v, sl):
        """the slice of a per-point argument that belongs to one sub-batch (anything else is passed through)"""
        # per-point scalars are [B], per-point material grids [B,nx,ny]; a 2-D tensor is always a shared grid
        if isinstance(v, torch.Tensor) and v.dim() in (1, 3) and v.shape[0] == self._B and v.numel() > 1:
            return v[sl]
        return v

    def _fanout(self, method, args=(), kwargs=None, gates=None):
        """Run `method` of every sub-batch simulation on its own stream from its own host thread (the C ABI calls release
        the GIL; rcwa_eig polls its convergence flag from the host, so each sub-batch needs its own thread to keep its
        queue fed).  Joined -- on the host and on the caller's stream -- before returning."""
        import threading
        kwargs = kwargs or {}
        cur = torch.cuda.current_stream(self._device)
        start = torch.cuda.Event()
        start.record(cur)
        results, errors = [None] * len(self._children), []

        def work(i):
            child, sl = self._children[i], self._slices[i]
            child._gate = gates[i] if gates else None
            try:
                with torch.cuda.device(self._device), torch.cuda.stream(self._streams[i]):
                    self._streams[i].wait_event(start)
                    results[i] = getattr(child, method)(*[self._part(a, sl) for a in args],
                                                        **{k: self._part(v, sl) for k, v in kwargs.items()})
            except BaseException as e:      # re-raised in the caller's thread
                errors.append(e)
            finally:
                if child._gate is not None and child._gate[2] is not None:
                    child._gate[2].set()    # never leave the next sub-batch waiting
                child._gate = None
        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(self._children))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for st in self._streams:
            cur.wait_stream(st)
        if errors:
            raise errors[0]
        return results

    def _gather(self, parts):
        """per-sub-batch public tensors -> one tensor on the caller's stream"""
        cur = torch.cuda.current_stream(self._device)
        for p in parts:
            p.record_stream(cur)
        return torch.cat(parts, dim=0)

    # ------------------------------------------------------------------ helpers
    def _b(self, v):
        """scalar or [B] -> complex128 [B] on the device."""
        # python scalars must be widened directly: torch.as_tensor(0.3) would round to float32 first
        t = v.to(device=self._device, dtype=_C) if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=_C, device=self._device)
        t = t.reshape(-1)
        if t.numel() == 1:
            t = t.expand(self._B)
        elif t.numel() != self._B:
            raise ValueError('expected a scalar or %d values' % self._B)
        return t

    def _pub(self, t):
        """internal complex128 [B,...] -> public tensor (sim dtype, batch dim dropped if unbatched)."""
        t = t.to(self._dtype)
        return t if self._batched else t[0]

    # ------------------------------------------------------------------ setup (rcwa.py:95-144)
    def add_input_layer(self, eps=1., mu=1.):
        self.eps_in = torch.as_tensor(eps, dtype=self._dtype, device=self._device)
        self.mu_in = torch.as_tensor(mu, dtype=self._dtype, device=self._device)
        self.Sin = []
        if self._children:
            self._fanout('add_input_layer', kwargs=dict(eps=eps, mu=mu))

    def add_output_layer(self, eps=1., mu=1.):
        self.eps_out = torch.as_tensor(eps, dtype=self._dtype, device=self._device)
        self.mu_out = torch.as_tensor(mu, dtype=self._dtype, device=self._device)
        self.Sout = []
        if self._children:
            self._fanout('add_output_layer', kwargs=dict(eps=eps, mu=mu))

    def set_incident_angle(self, inc_ang, azi_ang, angle_layer='input'):
        self.inc_ang = torch.as_tensor(inc_ang, dtype=self._dtype, device=self._device)
        self.azi_ang = torch.as_tensor(azi_ang, dtype=self._dtype, device=self._device)
        if angle_layer in ['i', 'in', 'input']:
            self.angle_layer = 'input'
        elif angle_layer in ['o', 'out', 'output']:
            self.angle_layer = 'output'
        else:
            warnings.warn('Invalid angle layer. Set as input layer.', UserWarning)
            self.angle_layer = 'input'
        self._kvectors()
        if self._children:
            self._fanout('set_incident_angle', args=(inc_ang, azi_ang, self.angle_layer))

    def _kvectors(self):
        """Per-order wavevectors and half-space S-matrices (rcwa.py:1124-1181), O(N)."""
        B, N = self._B, self.order_N
        e_in, m_in = self._b(self.eps_in), self._b(self.mu_in)
        e_out, m_out = self._b(self.eps_out), self._b(self.mu_out)
        inc, azi = self._b(self.inc_ang), self._b(self.azi_ang)
        nref = torch.sqrt(e_in * m_in).real if self.angle_layer == 'input' else torch.sqrt(e_out * m_out).real
        kx0 = nref * torch.sin(inc) * torch.cos(azi)
        ky0 = nref * torch.sin(inc) * torch.sin(azi)
        kxl = kx0[:, None] + self.order_x[None, :] * self._Gx[:, None]      # [B, 2ox+1]
        kyl = ky0[:, None] + self.order_y[None, :] * self._Gy[:, None]      # [B, 2oy+1]
        kx = kxl[:, :, None].expand(B, len(self.order_x), len(self.order_y)).reshape(B, N).contiguous()
        ky = kyl[:, None, :].expand(B, len(self.order_x), len(self.order_y)).reshape(B, N).contiguous()
        self._kx, self._ky = kx, ky
        self._k0_zero = (bool((kx0.abs() < 1e-14).all()), bool((ky0.abs() < 1e-14).all()))
        self.kx0_norm, self.ky0_norm = self._pub(kx0), self._pub(ky0)
        self.kx_norm, self.ky_norm = self._pub(kxl), self._pub(kyl)
        self.Kx_norm_dn, self.Ky_norm_dn = self._pub(kx), self._pub(ky)
        self._Vf = v_matrix(kx, ky, sqrt_upper(1.0 - kx * kx - ky * ky))
        self._Vf_inv = bd_inv(self._Vf).contiguous()
        if hasattr(self, 'Sin'):
            Vi = v_matrix(kx, ky, sqrt_upper((e_in * m_in)[:, None] - kx * kx - ky * ky))
            T = bd_inv(self._Vf + Vi)
            D = bd_mul(T, self._Vf - Vi)
            self._Vi = Vi
            self._Sin = [2 * bd_mul(T, Vi), -D, D, 2 * bd_mul(T, self._Vf)]
            self.Sin = _LazyDense(self, self._Sin)
        if hasattr(self, 'Sout'):
            Vo = v_matrix(kx, ky, sqrt_upper((e_out * m_out)[:, None] - kx * kx - ky * ky))
            T = bd_inv(self._Vf + Vo)
            D = bd_mul(T, self._Vf - Vo)
            self._Vo = Vo
            self._Sout = [2 * bd_mul(T, self._Vf), D, -D, 2 * bd_mul(T, Vo)]
            self.Sout = _LazyDense(self, self._Sout)

    # dense views of the 2x2-block-diagonal matrices, built on demand (drop-in attributes)
    @property
    def Kx_norm(self):
        return torch.diag_embed(self.Kx_norm_dn)

    @property
    def Ky_norm(self):
        return torch.diag_embed(self.Ky_norm_dn)

    @property
    def Vf(self):
        return self._pub(_lib.blockdiag_dense(self._Vf.contiguous()))

    @property
    def Vi(self):
        return self._pub(_lib.blockdiag_dense(self._Vi.contiguous()))

    @property
    def Vo(self):
        return self._pub(_lib.blockdiag_dense(self._Vo.contiguous()))

    # ------------------------------------------------------------------ layers (rcwa.py:146-170)
    def _is_homogeneous(self, v):
        # same acceptance rule as the reference (rcwa.py:156-157): float, complex, 0-d tensor,
        # one-element 1-d tensor; an int has no .dim() and raises AttributeError before any state changes.
        if type(v) == float or type(v) == complex:
            return True
        if v.dim() == 0 or (v.dim() == 1 and v.shape[0] == 1):
            return True
        return self._batched and v.dim() == 1 and v.shape[0] == self._B     # per-point scalar (new)

    def _conv(self, grid, differentiable=False):
        """Convolution matrix of a sampled cell, [B,N,N] complex128 (CUDA stage 1)."""
        grid = torch.as_tensor(grid, device=self._device)
        want_real = torch.float32 if self._dtype == torch.complex64 else torch.float64
        if grid.dtype not in (want_real, self._dtype):
            # the reference fails inside torch.matmul on mixed precision (SURVEY.md finding 9)
            raise RuntimeError('expected material of dtype %s or %s but found %s' % (want_real, self._dtype, grid.dtype))
        if grid.dim() == 3 and grid.shape[0] != self._B:
            raise ValueError('batched material must have leading dimension %d' % self._B)
        if grid.shape[-2] < 4 * self.order[0] + 1 or grid.shape[-1] < 4 * self.order[1] + 1:
            raise IndexError('material grid too coarse for the Fourier order (needs >= 4*order+1 samples)')
        if differentiable:
            return autodiff.ConvMat.apply(grid, self.order[0], self.order[1], self._B)
        return _lib.convmat(grid, self.order[0], self.order[1], nb=self._B)

    def add_layer(self, thickness, eps=1., mu=1.):
        he, hm = self._is_homogeneous(eps), self._is_homogeneous(mu)
        if self._children:
            return self._add_layer_pipelined(thickness, eps, mu, patterned=not (he and hm))
        B, N = self._B, self.order_N
        kx, ky = self._kx, self._ky
        thick = (thickness.to(device=self._device, dtype=torch.float64) if isinstance(thickness, torch.Tensor)
                 else torch.tensor(thickness, dtype=torch.float64, device=self._device)).reshape(-1)
        thick = (thick.expand(B) if thick.numel() == 1 else thick).contiguous()
        omega = (self._omega64.expand(B) if self._omega64.numel() == 1 else self._omega64).contiguous()

        diff = any(isinstance(v, torch.Tensor) and v.requires_grad for v in (eps, mu, thickness))
        if diff:
            self._diff = True
        if (diff and not (he and hm)) or not hm:
            # layers whose symmetry is not analysed (differentiable pipeline, patterned permeability): the stack is cascaded in
            # the original basis; block layers added so far are returned to it in solve_global_smatrix
            self._sym = False
        blocks = None
        if he and hm:
            S11, S21, kz, Qbd = self._homogeneous_layer(self._b(eps), self._b(mu), omega, thick, diff)
            if self._store:
                self._modes_src.append(dict(W=None, Q=Qbd.detach(), kz=kz.detach(), E=self._b(eps).detach(), M=self._b(mu).detach(),
                                            thick=thick.detach(), omega=omega))
                self.eps_conv.append(self._pub(self._b(eps)[:, None, None] * torch.eye(N, dtype=_C, device=self._device)))
                self.mu_conv.append(self._pub(self._b(mu)[:, None, None] * torch.eye(N, dtype=_C, device=self._device)))
                self.E_eigvec.append(self._pub(torch.eye(2 * N, dtype=_C, device=self._device).expand(B, -1, -1)))
        elif diff:
            # gradients requested: same algebra on the differentiable primitives (torcwa_b200/autodiff.py)
            eyeN = torch.eye(N, dtype=_C, device=self._device)
            E = self._b(eps)[:, None, None] * eyeN if he else self._conv(eps, differentiable=True)
            if hm:
                mu_s = self._b(mu)
                Mc, nu = mu_s[:, None, None] * eyeN, (1 / mu_s)[:, None, None] * eyeN
            else:
                Mc = self._conv(mu, differentiable=True)
                nu = autodiff.inverse(Mc)
            S11, S21, kz, W, P, Q = autodiff.patterned_layer(E, Mc, nu, kx, ky, self._Vf_inv, omega, thick)
            self.eig_info.append(torch.zeros(B, dtype=torch.int32, device=self._device))   # Eig raises on non-convergence
            if self._store:
                self.eps_conv.append(self._pub(E)); self.mu_conv.append(self._pub(Mc))
                self.P.append(self._pub(P)); self.Q.append(self._pub(Q))
                self.E_eigvec.append(self._pub(W))
                self._modes_src.append(dict(W=W.detach(), Q=Q.detach(), kz=kz.detach(), E=E.detach(), M=Mc.detach(),
                                            thick=thick.detach(), omega=omega))
        else:
            E = self._b(eps)[:, None, None] * torch.eye(N, dtype=_C, device=self._device) if he else self._conv(eps)
            E = E.contiguous()
            basis = self._symmetry_of(E) if (hm and not he) else None
            eta, info_e = _lib.inverse(E)
            self._status.append(('inverse of the permittivity convolution matrix (layer %d)' % self.layer_N, info_e))
            if hm:
                P, Q = _lib.pq_assemble(eta, E, kx, ky, mu_scalar=self._b(mu).contiguous())
                M = None
            else:
                M = self._conv(mu)
                nu, _ = _lib.inverse(M)
                P, Q = _lib.pq_assemble(eta, E, kx, ky, Mc=M, nu=nu)
                del nu
            del eta
            if self._store:
                self.eps_conv.append(self._pub(E))
                self.mu_conv.append(self._pub(M if M is not None else self._b(mu)[:, None, None] * torch.eye(N, dtype=_C, device=self._device)))
                self.P.append(self._pub(P)); self.Q.append(self._pub(Q))
            E_keep, M_keep = (E, M if M is not None else self._b(mu)) if self._store else (None, None)
            del E, M
            if self.avoid_Pinv_instability:
                self._pinv_metrics(P, Q)
            if basis is not None:
                blocks, kz, W, S11, S21 = self._patterned_layer_blocks(basis, P, Q, omega, thick)
                del P
            else:
                blocks = None
                A = _lib.zgemm(P, Q)
                del P                                  # free early: a batch chunk is sized by its peak footprint
                lam, W, info = self._eig(A)
                del A
                self.eig_info.append(info)
                self._status.append(('eigendecomposition (layer %d): QR iteration did not converge' % self.layer_N, info))
                kz = _lib.kz_branch(lam)
                ka = kz.abs()
                self._kz_min.append((ka.amin(dim=1) / ka.amax(dim=1)).min())
                del ka
                S11, S21, info_s = _lib.layer_smatrix(W, kz, Q, self._Vf_inv, omega, thick, slices=self._digits)
                self._status.append(('layer S-matrix (layer %d): singular coupling matrix' % self.layer_N, info_s))
            if self._store:
                self.E_eigvec.append(self._pub(W))
                self._modes_src.append(dict(W=W, Q=Q, kz=kz, E=E_keep, M=M_keep, thick=thick, omega=omega))
            del Q, W
        self.kz_norm.append(self._pub(kz))
        self.layer_N += 1
        self.thickness.append(thickness)
        self._layers.append(_BlockLayer(blocks, self._sym) if blocks is not None else [S11, S21])
        if self._store:
            s11, s21 = self._pub(S11), self._pub(S21)
            self.layer_S11.append(s11); self.layer_S21.append(s21)
            self.layer_S12.append(s21); self.layer_S22.append(s11)      # single-layer symmetry

    # ------------------------------------------------------------------ symmetry-reduced layers (torcwa_b200/symmetry.py)
    def _proj(self, basis, X, chi, left='E', right='E'):
        return _lib.sym_project(X.contiguous(), *basis.tables(chi, left, right))

    def _symmetry_of(self, E):
        """Basis of the symmetry this layer shares with the stack so far (None: general path).  A layer with less symmetry
        than its predecessors (a C2 layer on top of a doubly mirror-symmetric one, ...) moves the whole stack to the common
        subgroup: the earlier block layers are re-expressed in its basis (O(n^2) per layer); no common element at all
        sends the stack to the general path."""
        if not self._sym_on or not hasattr(self, '_k0_zero') or self._sym is False:
            return None
        ox, oy = int(self.order[0]), int(self.order[1])
        found = symmetry.detect(E, ox, oy, self._k0_zero[0], self._k0_zero[1])
        if found is None:
            self._sym = False                     # one layer without it: the whole stack runs the general path
            return None
        gens, thx, thy = found
        if self._sym is None:
            self._set_basis(symmetry.Basis(ox, oy, gens, thx, thy, self._device))
            return self._sym
        common = symmetry.common_subgroup(self._sym.gens, (self._sym.thx, self._sym.thy), gens, (thx, thy))
        if common is None:
            self._sym = False
            return None
        if set(common) != set(self._sym.gens):
            old = self._sym
            self._set_basis(symmetry.Basis(ox, oy, common, old.thx, old.thy, self._device))
            for layer in self._layers:
                if isinstance(layer, _BlockLayer):
                    dense = [old.unproject({c: v[k] for c, v in layer.blocks.items()}) for k in range(2)]
                    layer.blocks = {chi: [self._proj(self._sym, d, chi) for d in dense] for chi in self._sym.chars}
                    layer.basis = self._sym
                    del dense
        return self._sym

    def _set_basis(self, basis):
        self._sym = basis
        Vfi = _lib.blockdiag_dense(self._Vf_inv.contiguous())
        self._sym_G = {chi: self._proj(basis, Vfi, chi, 'E', 'H') for chi in basis.chars}

    def _patterned_layer_blocks(self, basis, P, Q, omega, thick):
        """One patterned layer solved block by block in the symmetry-adapted basis: the algebra of rcwa_layer_smatrix
        (V = Q W Kz^-1, T+- = R+- M+-^-1, S11 = T+ + T-, S21 = T+ - T- - I) on matrices of ~n/2 or ~n/4.
        All blocks go through the C-ABI calls as ONE batch (the QR phase of rcwa_eig is a chain of short launches whose
        latency hides only behind other matrices): a block one or two rows smaller than the largest is extended at the FRONT
        by an exactly decoupled diagonal entry (P' = diag(1, P), Q' = diag(1, Q), Vf^-1' = diag(1, G)).  Row / column 0 is
        then never touched by a Householder reflector of the Hessenberg reduction (its first column is already zero), the
        QR iteration sees a split at the top, and every later operator stays diag(t, T) with exact zeros in the couplings
        -- the block's results are the trailing sub-matrices, bit for bit what the unextended block would give up to the
        order of floating-point summation."""
        B = self._B
        out, kzs, Ws, infos = {}, {}, {}, []
        chars = basis.chars
        nk = max(basis.sizes.values())
        pad = {chi: nk - basis.sizes[chi] for chi in chars}

        def stacked(fn, dummy):
            """the blocks of all characters as ONE batch [len(chars) B, nk, nk]: a block smaller than the largest is
            extended at the FRONT by decoupled diagonal entries `dummy` (see _patterned_layer_blocks.__doc__)"""
            if not any(pad.values()):
                return torch.cat([fn(chi) for chi in chars], dim=0)
            X = torch.zeros((len(chars) * B, nk, nk), dtype=_C, device=self._device)
            for j, chi in enumerate(chars):
                d = pad[chi]
                X[j * B:(j + 1) * B, d:, d:] = fn(chi)
                if d:
                    X[j * B:(j + 1) * B, range(d), range(d)] = dummy
            return X
        Pk = stacked(lambda chi: self._proj(basis, P, chi, 'E', 'H'), 1.0)
        Qk = stacked(lambda chi: self._proj(basis, Q, chi, 'H', 'E'), 1.0)
        Gk = stacked(lambda chi: self._sym_G[chi], 1.0)
        om, th = omega.repeat(len(chars)), thick.repeat(len(chars))
        A = _lib.zgemm(Pk, Qk)
        del Pk
        lam, W, info = self._eig(A)
        del A
        infos.append(info)
        kz = _lib.kz_branch(lam)
        ka = kz.abs()
        self._kz_min.append((ka.amin(dim=1) / ka.amax(dim=1)).min())
        V = _lib.zgemm(Qk, W) / kz[:, None, :]
        del Qk
        Bm = _lib.zgemm(Gk, V)
        del V, Gk
        X = torch.exp(1j * (om * th)[:, None] * kz)[:, None, :]
        Rp, Rm = W * (1 + X), W * (X - 1)
        Mp, Mm = Rp + Bm * (1 - X), W * (1 - X) + Bm * (1 + X)
        del Bm
        Tp, i1 = _lib.right_solve(Rp, Mp)
        del Rp, Mp
        Tm, i2 = _lib.right_solve(Rm, Mm)
        del Rm, Mm
        infos += [i1, i2]
        eye = torch.eye(nk, dtype=_C, device=self._device)
        S11, S21 = Tp + Tm, Tp - Tm - eye
        del Tp, Tm
        for j, chi in enumerate(chars):
            sl, d = slice(j * B, (j + 1) * B), pad[chi]
            out[chi] = [S11[sl, d:, d:].contiguous(), S21[sl, d:, d:].contiguous()]
            kzs[chi], Ws[chi] = kz[sl, d:], W[sl, d:, d:]
            if d:       # the extension must have stayed exactly decoupled (checked with the other status words, _check_status)
                self._pad_err.append(torch.stack([W[sl, :d, d:].abs().amax(), W[sl, d:, :d].abs().amax(), (lam[sl, :d] - 1).abs().amax(),
                                                  S11[sl, :d, d:].abs().amax(), S11[sl, d:, :d].abs().amax()]).amax())
        info_all = torch.stack([i.reshape(len(i) // B, B).abs().amax(dim=0) for i in infos]).amax(dim=0).to(torch.int32)
        self.eig_info.append(info_all)
        self._status.append(('symmetry-reduced layer %d: eigendecomposition or coupling-matrix factorisation failed' % self.layer_N, info_all))
        kz_full = torch.cat([kzs[chi] for chi in basis.chars], dim=1)
        W_full = S11_full = S21_full = None
        if self._store:
            # drop-in attributes in the original basis: eigenvectors W = [T_chi W_chi], dense layer S blocks
            W_full = torch.zeros((B, basis.n, basis.n), dtype=_C, device=self._device)
            c0 = 0
            for chi in basis.chars:
                idx, cf = basis.E[chi]
                for t in range(idx.shape[0]):
                    W_full[:, :, c0:c0 + basis.sizes[chi]].index_add_(1, idx[t], Ws[chi] * cf[t][None, :, None])
                c0 += basis.sizes[chi]
            S11_full = basis.unproject({chi: v[0] for chi, v in out.items()})
            S21_full = basis.unproject({chi: v[1] for chi, v in out.items()})
        return out, kz_full, W_full, S11_full, S21_full

    def _eig(self, A):
        """rcwa_eig; as a pipelined sub-batch, wait for the previous sub-batch's Hessenberg phase and announce our own."""
        gate = self._gate
        if gate is None:
            return _lib.eig(A)
        wait_flag, wait_event, signal_flag, signal_event = gate
        if wait_flag is not None:
            wait_flag.wait()
            torch.cuda.current_stream().wait_event(wait_event)

        def announce():
            if signal_flag is not None:
                signal_event.record(torch.cuda.current_stream())
                signal_flag.set()
        return _lib.eig(A, after_reduction=announce)

    def _add_layer_pipelined(self, thickness, eps, mu, patterned):
        import threading
        k = len(self._children)
        diff = any(isinstance(v, torch.Tensor) and v.requires_grad for v in (eps, mu, thickness))
        gates = None
        if patterned and not diff:
            flags = [threading.Event() for _ in range(k)]
            events = [torch.cuda.Event() for _ in range(k)]
            gates = [(flags[i - 1] if i > 0 else None, events[i - 1] if i > 0 else None,
                      flags[i] if i < k - 1 else None, events[i] if i < k - 1 else None) for i in range(k)]
        self._fanout('add_layer', args=(thickness,), kwargs=dict(eps=eps, mu=mu), gates=gates)
        self._diff = self._diff or any(c._diff for c in self._children)
        self.kz_norm.append(self._gather([c.kz_norm[-1] for c in self._children]))
        if patterned and not diff:
            self.eig_info.append(self._gather([c.eig_info[-1] for c in self._children]))
        if self.avoid_Pinv_instability and patterned and not diff:
            self.Pinv_instability.append(self._gather([c.Pinv_instability[-1] for c in self._children]))
            self.Qinv_instability.append(self._gather([c.Qinv_instability[-1] for c in self._children]))
        self.layer_N += 1
        self.thickness.append(thickness)

    def _pinv_metrics(self, P, Q):
        """`avoid_Pinv_instability=True` (rcwa.py:1249-1262): the reference measures how badly P (and Q) invert,
        max|P P^-1 - I| and max|P^-1 P - I|, to choose between H = P^-1 W Kz and H = Q W Kz^-1.  This path always
        uses the second, inverse-free form, so the numbers are reported for the user only.  (The reference's Q
        metric evaluates Q Q^-1 twice, :1253-1254; one evaluation is the same number.)"""
        eye = torch.eye(P.shape[-1], dtype=_C, device=self._device)
        Pi, _ = _lib.inverse(P)
        m = torch.maximum((_lib.zgemm(P, Pi) - eye).abs().amax(dim=(1, 2)), (_lib.zgemm(Pi, P) - eye).abs().amax(dim=(1, 2)))
        del Pi
        Qi, _ = _lib.inverse(Q)
        q = (_lib.zgemm(Q, Qi) - eye).abs().amax(dim=(1, 2))
        del Qi
        self.Pinv_instability.append((m if self._batched else m[0]).to(self._rdtype))
        self.Qinv_instability.append((q if self._batched else q[0]).to(self._rdtype))

    def _homogeneous_layer(self, eps, mu, omega, thick, diff=False):
        """Analytic modes (rcwa.py:1206-1222: W = I, kz = conj-branch sqrt) pushed through the
        minimal layer-S algebra in 2x2-block form; returns dense S11, S21 and kz [B,2N]."""
        kx, ky = self._kx, self._ky
        kz1 = sqrt_upper((eps * mu)[:, None] - kx * kx - ky * ky)
        im = 1 / mu[:, None]
        Q = bd(-kx * ky * im, kx * kx * im - eps[:, None], eps[:, None] - ky * ky * im, ky * kx * im)
        ikz = bd_diag(1 / kz1)
        Bm = bd_mul(self._Vf_inv, bd_mul(Q, ikz))                 # Vf^-1 Q Kz^-1  (W = I)
        X = torch.exp(1j * (omega * thick)[:, None] * kz1)
        onep, onem = bd_diag(1 + X), bd_diag(1 - X)
        Tp = bd_mul(onep, bd_inv(onep + bd_mul(Bm, onem)))        # R+ M+^-1
        Tm = bd_mul(-onem, bd_inv(onem + bd_mul(Bm, onep)))       # R- M-^-1
        eye = bd_diag(torch.ones_like(kz1))
        dense = autodiff.blockdiag_dense if diff else _lib.blockdiag_dense
        S11 = dense((Tp + Tm).contiguous())
        S21 = dense((Tp - Tm - eye).contiguous())
        return S11, S21, torch.cat((kz1, kz1), dim=1), Q

    # ------------------------------------------------------------------ cascade (rcwa.py:173-211)
    def solve_global_smatrix(self):
        B, n = self._B, 2 * self.order_N
        if self._children:
            self._fanout('solve_global_smatrix')
            self._S = _CatList(self, '_S')
            self.S = _CatList(self, 'S')
            self.C = [[], []]
            return
        if self._diff:
            self._dense_layers()
            return self._solve_global_smatrix_differentiable()
        if any(isinstance(l, _BlockLayer) for l in self._layers):
            if self._sym not in (None, False):
                return self._solve_global_smatrix_blocks()
            self._dense_layers()            # a later layer broke the symmetry
        if self.layer_N > 0:
            s11, s21 = self._layers[0]
            S = [s11, s21, s21, s11]
            for i in range(1, self.layer_N):
                n11, n21 = self._layers[i]
                S, info_r = _lib.redheffer(S, [n11, n21, n21, n11], slices=self._digits)
                self._status.append(('star product with layer %d' % i, info_r))
        else:
            eye = torch.eye(n, dtype=_C, device=self._device).expand(B, -1, -1).contiguous()
            zero = torch.zeros((B, n, n), dtype=_C, device=self._device)
            S = [eye, zero, zero.clone(), eye.clone()]
        if hasattr(self, 'Sin'):
            S, info_r = _lib.redheffer_bdleft(self._Sin, S, slices=self._digits)     # Sin is four-diagonal: O(n^2) instead of 6 GEMMs
            self._status.append(('star product with the input half space', info_r))
        if hasattr(self, 'Sout'):
            S, info_r = _lib.redheffer(S, [_lib.blockdiag_dense(s.contiguous()) for s in self._Sout], slices=self._digits)
            self._status.append(('star product with the output half space', info_r))
        self._check_status()
        self._S = S
        self.S = [self._pub(s) for s in S]
        self.C = [[], []]      # filled on demand (fields.ensure_modes): the fused cascade does not carry mode coefficients
        self._modes_ready = False

    def _solve_global_smatrix_blocks(self):
        """The left fold of solve_global_smatrix (rcwa.py:173-211) block by block in the symmetry-adapted basis; layers and
        half spaces that were built in the original basis (homogeneous layers: four diagonals) are projected first.  The
        global S-matrix is returned to the original basis at the end (O(n^2))."""
        basis = self._sym
        Sin_d = [_lib.blockdiag_dense(s.contiguous()) for s in self._Sin] if hasattr(self, 'Sin') else None
        Sout_d = [_lib.blockdiag_dense(s.contiguous()) for s in self._Sout] if hasattr(self, 'Sout') else None
        Sblocks = {}
        for chi in basis.chars:
            def part(layer):
                if isinstance(layer, _BlockLayer):
                    a, b = layer.blocks[chi]
                else:
                    a, b = self._proj(basis, layer[0], chi), self._proj(basis, layer[1], chi)
                return [a, b, b, a]
            S = part(self._layers[0])
            for i in range(1, self.layer_N):
                nxt = part(self._layers[i])
                if isinstance(self._layers[i], _BlockLayer):
                    S, info_r = _lib.redheffer(S, nxt, slices=self._digits)
                else:       # a homogeneous layer: four diagonals in the original basis, like a half space
                    S, info_r = self._four_diagonal_product(nxt, S, left=False)
                self._status.append(('star product with layer %d (block %s)' % (i, chi), info_r))
            if hasattr(self, 'Sin'):
                S, info_r = self._four_diagonal_product([self._proj(basis, s, chi) for s in Sin_d], S, left=True)
                self._status.append(('star product with the input half space (block %s)' % (chi,), info_r))
            if hasattr(self, 'Sout'):
                S, info_r = self._four_diagonal_product([self._proj(basis, s, chi) for s in Sout_d], S, left=False)
                self._status.append(('star product with the output half space (block %s)' % (chi,), info_r))
            Sblocks[chi] = S
        del Sin_d, Sout_d
        self._check_status()
        # the dense blocks in the original basis are formed on first access; S_parameters reads its entries off the blocks
        self._S = _BlockS(basis, Sblocks)
        self.S = self._S.view(self)
        self.C = [[], []]
        self._modes_ready = False

    def _four_diagonal_product(self, blocks, S, left):
        """blocks (x) S (left) or S (x) blocks, where `blocks` are the projections of four-diagonal matrices (a half space, a
        homogeneous layer): a diagonal plus one partner entry per row in the adapted basis (symmetry.PairSparse), so six
        (left) or four (right) of the eight dense products of the star product become row / column combinations.  Falls back
        to the dense routine if the pattern is not found."""
        uniq = {}
        for b in blocks:                                     # a layer passes [S11, S21, S21, S11]: analyse each tensor once
            if id(b) not in uniq:
                uniq[id(b)] = symmetry.PairSparse.from_dense(b)
        if not bool(torch.stack([x.ok for x in uniq.values()]).all()):            # one host read for all blocks
            return _lib.redheffer(blocks, S, slices=self._digits) if left else _lib.redheffer(S, blocks, slices=self._digits)
        sparse = [uniq[id(b)] for b in blocks]
        if left:
            return symmetry.redheffer_sparse_left(_lib, sparse, S)
        return symmetry.redheffer_sparse_right(_lib, S, sparse)

    def _layer_pairs_dense(self):
        """[S11, S21] of every layer in the original basis (layers held as symmetry blocks are unprojected, O(n^2) each; cached
        until the layer list changes -- the field code asks once per design point)"""
        key = tuple(id(l.blocks) if isinstance(l, _BlockLayer) else id(l) for l in self._layers)
        cached = getattr(self, '_pairs_cache', None)
        if cached is None or cached[0] != key:
            pairs = [[l.basis.unproject({c: v[k] for c, v in l.blocks.items()}) for k in range(2)] if isinstance(l, _BlockLayer) else l
                     for l in self._layers]
            if not any(isinstance(l, _BlockLayer) for l in self._layers):
                return pairs                          # nothing was computed: nothing to cache
            self._pairs_cache = (key, pairs, list(self._layers))      # the layer objects are kept alive so that the ids stay theirs
        return self._pairs_cache[1]

    def _dense_layers(self):
        """bring layers held as symmetry blocks back to the original basis (the stack is not cascaded in blocks after all)"""
        self._layers = self._layer_pairs_dense()

    def _check_status(self):
        """Numerical status of everything enqueued so far: ONE device-to-host read of the per-matrix info words
        (the reference's torch.linalg.eig / inv raise LinAlgError eagerly; here the kernels only record a status
        and the host looks at it once, when the global S-matrix is complete)."""
        if not self._status:
            return
        # Wood-anomaly guard (scope row f3).  The layer kernel forms the H modes as V = Q W Kz^-1 (the reference's fallback
        # branch, rcwa.py:1260); the reference's DEFAULT, P^-1 W Kz (:1264), stays finite when a mode's kz vanishes, this
        # form does not -- say so instead of returning a silently degraded S-matrix.
        if self._kz_min:
            ratio = float(torch.stack(self._kz_min).min())
            if ratio < 1e-9:
                warnings.warn('torcwa_b200: a layer mode has |kz| / max|kz| = %.1e (Wood anomaly / cut-off): the H modes are '
                              'formed as Q W Kz^-1 and lose accuracy there; the reference\'s default P^-1 W Kz stays finite' % ratio,
                              UserWarning)
            self._kz_min = []
        if self._pad_err:
            leak = float(torch.stack(self._pad_err).max())
            self._pad_err = []
            if leak != 0.0:
                raise RuntimeError('torcwa_b200 internal error: the decoupled extension of a symmetry block picked up a coupling of %.3e' % leak)
        worst = torch.stack([i.to(self._device).abs().max() for _, i in self._status])
        if int(worst.max()) != 0:
            k = int(torch.nonzero(worst)[0])
            what, info = self._status[k]
            raise torch.linalg.LinAlgError('torcwa_b200: %s, batch entries %s (info = %s)'
                                           % (what, torch.nonzero(info).flatten().tolist(), info[info != 0].tolist()))
        self._status = []

    def _solve_global_smatrix_differentiable(self):
        """The same left fold (rcwa.py:173-211) on the differentiable star product."""
        B, n = self._B, 2 * self.order_N
        if self.layer_N > 0:
            s11, s21 = self._layers[0]
            S = [s11, s21, s21, s11]
            for i in range(1, self.layer_N):
                n11, n21 = self._layers[i]
                S = autodiff.redheffer(S, [n11, n21, n21, n11])
        else:
            eye = torch.eye(n, dtype=_C, device=self._device).expand(B, -1, -1).contiguous()
            zero = torch.zeros((B, n, n), dtype=_C, device=self._device)
            S = [eye, zero, zero.clone(), eye.clone()]
        if hasattr(self, 'Sin'):
            S = autodiff.redheffer([autodiff.blockdiag_dense(s) for s in self._Sin], S)
        if hasattr(self, 'Sout'):
            S = autodiff.redheffer(S, [autodiff.blockdiag_dense(s) for s in self._Sout])
        self._check_status()         # info words of the fused (non-differentiable) layers of a mixed stack
        self._S = S
        self.S = [self._pub(s) for s in S]
        self.C = [[], []]
        self._modes_ready = False

    # ------------------------------------------------------------------ small utilities (rcwa.py:214-298)
    def diffraction_angle(self, orders, *, layer='output', unit='radian'):
        """Inclination and azimuth of the selected diffraction orders in the input or output half space
        (rcwa.py:214-262).  Unbatched: tensors of shape [len(orders)]; batched: [B, len(orders)]."""
        orders = torch.as_tensor(orders, dtype=torch.int64, device=self._device).reshape([-1, 2])
        if layer in ['i', 'in', 'input']:
            layer = 'input'
        elif layer in ['o', 'out', 'output']:
            layer = 'output'
        else:
            warnings.warn('Invalid layer. Set as output layer.', UserWarning)
            layer = 'output'
        if unit in ['r', 'rad', 'radian']:
            unit = 'radian'
        elif unit in ['d', 'deg', 'degree']:
            unit = 'degree'
        else:
            warnings.warn('Invalid unit. Set as radian.', UserWarning)
            unit = 'radian'
        idx = self._matching_indices(orders)
        eps, mu = (self.eps_in, self.mu_in) if layer == 'input' else (self.eps_out, self.mu_out)
        kx, ky = self._kx[:, idx], self._ky[:, idx]
        kt = torch.sqrt(kx ** 2 + ky ** 2)
        kz = torch.sqrt((self._b(eps) * self._b(mu))[:, None] - kx ** 2 - ky ** 2)
        inc = torch.atan2(kt.real, kz.real)
        azi = torch.atan2(ky.real, kx.real)
        if unit == 'degree':
            inc, azi = (180. / pi) * inc, (180. / pi) * azi
        inc, azi = inc.to(self._rdtype), azi.to(self._rdtype)
        return (inc, azi) if self._batched else (inc[0], azi[0])

    def return_layer(self, layer_num, nx=100, ny=100):
        """eps and mu of a layer on an nx x ny grid, recovered from the truncated Fourier series held in the
        convolution matrices (rcwa.py:264-298): coefficient (i, j), |i| <= 2ox, |j| <= 2oy, is read from the first
        column (non-negative index) or first row (negative index) of the Toeplitz matrix.  Needs the stored
        convolution matrices (store_intermediates=True for batched simulations)."""
        if not self.eps_conv:
            raise RuntimeError('return_layer needs the stored convolution matrices: construct with store_intermediates=True')
        ox, oy = int(self.order[0]), int(self.order[1])
        wy = 2 * oy + 1
        i = torch.arange(-2 * ox, 2 * ox + 1, device=self._device)[:, None].expand(4 * ox + 1, 4 * oy + 1)
        j = torch.arange(-2 * oy, 2 * oy + 1, device=self._device)[None, :].expand(4 * ox + 1, 4 * oy + 1)
        ip, jp = torch.clamp(i, min=0), torch.clamp(j, min=0)         # contribution to the row index
        im, jm = torch.clamp(-i, min=0), torch.clamp(-j, min=0)       # contribution to the column index
        row, col = ip * wy + jp, im * wy + jm
        out = []
        for conv in (self.eps_conv[layer_num], self.mu_conv[layer_num]):
            c = conv if self._batched else conv[None]
            fftgrid = torch.zeros((c.shape[0], nx, ny), dtype=self._dtype, device=self._device)
            fftgrid[:, i % nx, j % ny] = c[:, row, col]
            rec = torch.fft.ifftn(fftgrid, dim=(-2, -1)) * nx * ny
            out.append(rec if self._batched else rec[0])
        return out[0], out[1]

    # ------------------------------------------------------------------ sources and fields (rcwa.py:526-1112)
    def source_planewave(self, *, amplitude=[1., 0.], direction='forward', notation='xy'):
        """Plane-wave source in the zeroth order (rcwa.py:526-537)."""
        self.source_fourier(amplitude=amplitude, orders=[0, 0], direction=direction, notation=notation)

    def source_fourier(self, *, amplitude, orders, direction='forward', notation='xy'):
        """Source given by amplitudes at selected diffraction orders (rcwa.py:539-596)."""
        from . import fields
        fields.source_fourier(self, amplitude, orders, direction, notation)

    def field_xz(self, x_axis, z_axis, y):
        """[Ex, Ey, Ez], [Hx, Hy, Hz] on the plane y = const (rcwa.py:598-775); tensors of shape [len(x), len(z)]."""
        from . import fields
        return fields.field_plane(self, 'xz', x_axis, z_axis, y)

    def field_yz(self, y_axis, z_axis, x):
        """[Ex, Ey, Ez], [Hx, Hy, Hz] on the plane x = const (rcwa.py:777-957); tensors of shape [len(y), len(z)]."""
        from . import fields
        return fields.field_plane(self, 'yz', y_axis, z_axis, x)

    def field_xy(self, layer_num, x_axis, y_axis, z_prop=0.):
        """[Ex, Ey, Ez], [Hx, Hy, Hz] on a plane z = const inside layer `layer_num` (-1: input, layer_N: output half
        space), z_prop from the layer's lower boundary (rcwa.py:959-1112); tensors of shape [len(x), len(y)]."""
        from . import fields
        return fields.field_xy(self, layer_num, x_axis, y_axis, z_prop)

    # ------------------------------------------------------------------ readout (rcwa.py:300-524)
    def _matching_indices(self, orders):
        # clamps out-of-range orders to the truncation edge, in place like the reference (rcwa.py:1115-1122)
        orders[orders[:, 0] < -self.order[0], 0] = int(-self.order[0])
        orders[orders[:, 0] > self.order[0], 0] = int(self.order[0])
        orders[orders[:, 1] < -self.order[1], 1] = int(-self.order[1])
        orders[orders[:, 1] > self.order[1], 1] = int(self.order[1])
        return len(self.order_y) * (orders[:, 0] + int(self.order[0])) + orders[:, 1] + int(self.order[1])

    def _kz_power(self, eps, mu, evanscent, abs_when_evanescent=False):
        kzc = torch.sqrt((self._b(eps) * self._b(mu))[:, None] - self._kx ** 2 - self._ky ** 2)
        ev = torch.abs(kzc.real / kzc.imag) < evanscent
        repl = torch.abs(kzc.real) if abs_when_evanescent else torch.zeros_like(kzc.real)
        k = torch.where(ev, repl, kzc.real)
        return torch.cat((k, k), dim=1)

    def S_parameters(self, orders, *, direction='forward', port='transmission', polarization='xx',
                     ref_order=[0, 0], power_norm=True, evanscent=1e-3):
        if self._children:
            cur = torch.cuda.current_stream(self._device)
            parts = []
            for c, st in zip(self._children, self._streams):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    parts.append(c.S_parameters(orders, direction=direction, port=port, polarization=polarization,
                                                ref_order=ref_order, power_norm=power_norm, evanscent=evanscent))
                cur.wait_stream(st)
            return self._gather(parts)
        orders = torch.as_tensor(orders, dtype=torch.int64, device=self._device).reshape([-1, 2])
        if direction in ['f', 'forward']:
            direction = 'forward'
        elif direction in ['b', 'backward']:
            direction = 'backward'
        else:
            warnings.warn('Invalid propagation direction. Set as forward.', UserWarning)
            direction = 'forward'
        if port in ['t', 'transmission']:
            port = 'transmission'
        elif port in ['r', 'reflection']:
            port = 'reflection'
        else:
            warnings.warn('Invalid port. Set as tramsmission.', UserWarning)
            port = 'transmission'
        if polarization not in ['xx', 'yx', 'xy', 'yy', 'pp', 'sp', 'ps', 'ss']:
            warnings.warn('Invalid polarization. Set as xx.', UserWarning)
            polarization = 'xx'
        ref_order = torch.as_tensor(ref_order, dtype=torch.int64, device=self._device).reshape([1, 2])
        oi = self._matching_indices(orders)
        ri = self._matching_indices(ref_order)
        N = self.order_N
        blk = {('forward', 'transmission'): 0, ('forward', 'reflection'): 1,
               ('backward', 'reflection'): 2, ('backward', 'transmission'): 3}[(direction, port)]
        if isinstance(self._S, _BlockS):
            entry = lambda a, b: self._S.entries(blk, a, b)
        else:
            entry = lambda a, b: self._S[blk][:, a, b]
        side = {0: ('out', 'in'), 1: ('in', 'in'), 2: ('out', 'out'), 3: ('in', 'out')}[blk]
        kx, ky = self._kx, self._ky

        if polarization in ['xx', 'yx', 'xy', 'yy']:
            oi2 = oi + N if polarization[0] == 'y' else oi
            ri2 = ri + N if polarization[1] == 'y' else ri
            out = entry(oi2, ri2)
            if power_norm:
                kzs = {'in': self._kz_power(self.eps_in, self.mu_in, evanscent),
                       'out': self._kz_power(self.eps_out, self.mu_out, evanscent)}
                kx2 = torch.cat((kx.real, kx.real), dim=1)
                ky2 = torch.cat((ky.real, ky.real), dim=1)
                num_pol = kx2 if polarization[0] == 'x' else ky2
                den_pol = kx2 if polarization[1] == 'x' else ky2
                num_kz, den_kz = kzs[side[0]], kzs[side[1]]
                norm = torch.sqrt((1 + (num_pol[:, oi2] / num_kz[:, oi2]) ** 2) / (1 + (den_pol[:, ri2] / den_kz[:, ri2]) ** 2))
                norm = norm * torch.sqrt(num_kz[:, oi2] / den_kz[:, ri2])
                out = out * norm
            out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
            out = torch.where(torch.isnan(out), torch.zeros_like(out), out)
            return self._pub(out)

        osign, rsign = {0: (1, 1), 1: (-1, 1), 2: (1, -1), 3: (-1, -1)}[blk]
        em = {'in': self._b(self.eps_in) * self._b(self.mu_in), 'out': self._b(self.eps_out) * self._b(self.mu_out)}

        def angles(idx, k2, sign):
            kxs, kys = kx[:, idx], ky[:, idx]
            kt = torch.sqrt(kxs ** 2 + kys ** 2)
            kzc = torch.sqrt(k2[:, None] - kxs ** 2 - kys ** 2)
            kz = sign * torch.abs(kzc.real)
            ev = torch.abs(kzc.real / kzc.imag) < evanscent
            return torch.atan2(kt.real, kz), torch.atan2(kys.real, kxs.real), ev

        o_inc, o_azi, o_ev = angles(oi, em[side[0]], osign)
        r_inc, r_azi, r_ev = angles(ri, em[side[1]], rsign)

        def pick(a, b):
            v = entry(a, b)
            return torch.where(o_ev, torch.zeros_like(v), v)

        xx, xy = pick(oi, ri), pick(oi, ri + N)
        yx, yy = pick(oi + N, ri), pick(oi + N, ri + N)
        co, so, ci = torch.cos(o_azi), torch.sin(o_azi), torch.cos(o_inc)
        cr, sr, cri = torch.cos(r_azi), torch.sin(r_azi), torch.cos(r_inc)
        if polarization == 'pp':
            out = (co / ci) * cri * cr * xx + (so / ci) * cri * cr * yx + (co / ci) * cri * sr * xy + (so / ci) * cri * sr * yy
        elif polarization == 'ps':
            out = (co / ci) * (-sr) * xx + (so / ci) * (-sr) * yx + (co / ci) * cr * xy + (so / ci) * cr * yy
        elif polarization == 'sp':
            out = -so * cri * cr * xx + co * cri * cr * yx - so * cri * sr * xy + co * cri * sr * yy
        else:
            out = -so * (-sr) * xx + co * (-sr) * yx - so * cr * xy + co * cr * yy
        out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
        out = torch.where(torch.isnan(out), torch.zeros_like(out), out)
        if power_norm:
            kzs = {'in': self._kz_power(self.eps_in, self.mu_in, evanscent),
                   'out': self._kz_power(self.eps_out, self.mu_out, evanscent, abs_when_evanescent=True)}
            out = out * torch.sqrt(kzs[side[0]][:, oi] / kzs[side[1]][:, ri])
        # an evanescent reference order gives zeros (rcwa.py:447-449), here per design point
        out = torch.where(r_ev, torch.zeros_like(out), out)
        return self._pub(out)


class _BlockLayer:
    """Layer S-matrix held as blocks in a symmetry-adapted basis: {character: [S11, S21]} (torcwa_b200/symmetry.py)."""

    def __init__(self, blocks, basis=None):
        self.blocks, self.basis = blocks, basis


class _BlockS:
    """The four global S-matrix blocks held as symmetry blocks {character: [S11, S21, S12, S22]}; list-like over the dense
    [B, n, n] matrices in the original basis, each formed on first access (S_parameters never needs them)."""

    def __init__(self, basis, blocks, cache=None, sim=None):
        # `sim` (public view only): weak, like _LazyDense -- a strong reference would make sim <-> sim.S a cycle and keep a
        # finished simulation's multi-GB blocks alive until the cyclic garbage collector happens to run
        self.basis, self.blocks, self._cache = basis, blocks, ({} if cache is None else cache)
        self._sim = weakref.ref(sim) if sim is not None else None

    def view(self, sim):
        """the public list `sim.S`: same blocks and cache, entries passed through sim._pub (dtype / batch squeeze)"""
        return _BlockS(self.basis, self.blocks, self._cache, sim)

    def entries(self, k, a, b):
        return self.basis.entries({chi: v[k] for chi, v in self.blocks.items()}, a, b)

    def __len__(self):
        return 4

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(4)[k]]
        k = range(4)[k]
        if k not in self._cache:
            self._cache[k] = self.basis.unproject({chi: v[k] for chi, v in self.blocks.items()})
        return self._sim()._pub(self._cache[k]) if self._sim is not None else self._cache[k]

    def __iter__(self):
        return (self[k] for k in range(4))


class _CatList:
    """The four global S-matrix blocks of a pipelined simulation: concatenated from the sub-batches on first access
    (a [B,n,n] copy per block -- the readout, S_parameters, never needs it)."""

    def __init__(self, sim, attr):
        self._sim, self._attr, self._val = weakref.ref(sim), attr, None

    def _get(self):
        if self._val is None:
            sim = self._sim()
            for st in sim._streams:
                torch.cuda.current_stream(sim._device).wait_stream(st)
            self._val = [torch.cat([getattr(c, self._attr)[k] for c in sim._children], dim=0) for k in range(4)]
        return self._val

    def __len__(self):
        return 4

    def __getitem__(self, k):
        return self._get()[k]

    def __iter__(self):
        return iter(self._get())


class _LazyDense:
    """List-like view that materialises dense [2N,2N] tensors from 2x2-block-diagonal storage on
    access (the reference's ``Sin`` / ``Sout`` are lists of four dense matrices)."""

    def __init__(self, sim, blocks):
        # weak back-reference: a strong one makes sim <-> sim.Sin a reference cycle, and a cycle keeps a finished
        # simulation's multi-GB S-matrices alive until the cyclic garbage collector happens to run
        self._sim, self._blocks = weakref.ref(sim), blocks

    def __len__(self):
        return len(self._blocks)

    def __getitem__(self, k):
        return self._sim()._pub(_lib.blockdiag_dense(self._blocks[k].contiguous()))
