"""ctypes binding of librcwa_b200.so (C ABI: include/rcwa_b200.h).

The library is the product: there is no Python/torch fallback.  If it is missing or a call
fails, this module raises -- it never reroutes work to another implementation.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("RCWA_B200_LIB", "librcwa_b200.so"))   # env: experimental build variants

EXPORTS = [
    "rcwa_b200_abi_version", "rcwa_gemm_scratch_bytes", "rcwa_convmat_workspace_bytes", "rcwa_convmat",
    "rcwa_zgemm_batched", "rcwa_zgemm_batched_cfg", "rcwa_zgemm_tc_workspace_bytes", "rcwa_zgemm_tc_batched", "rcwa_tc_split", "rcwa_tc_schedule", "rcwa_tc_issue_entry", "rcwa_set_tuning", "rcwa_get_tuning", "rcwa_lu_tinv_bytes", "rcwa_lu_factor", "rcwa_lu_solve_right", "rcwa_pq_assemble",
    "rcwa_eig_workspace_bytes", "rcwa_eig", "rcwa_eig_phases", "rcwa_eig_stats", "rcwa_eig_profile", "rcwa_hessenberg", "rcwa_hessenberg_matvec_probe", "rcwa_hessenberg_panel_width", "rcwa_kz_branch", "rcwa_eig_backward_workspace_bytes", "rcwa_eig_backward", "rcwa_layer_smatrix_workspace_bytes",
    "rcwa_layer_smatrix", "rcwa_redheffer_workspace_bytes", "rcwa_redheffer", "rcwa_redheffer_bdleft", "rcwa_blockdiag_dense", "rcwa_sym_project",
]

_vp, _i, _ll, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_size_t
_SIGS = {
    "rcwa_b200_abi_version": (_i, []),
    "rcwa_gemm_scratch_bytes": (_sz, [_i]),
    "rcwa_convmat_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "rcwa_convmat": (_i, [_vp, _i, _ll, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rcwa_zgemm_batched": (_i, [_i, _i, _i, _i, _i, _d, _d, _vp, _i, _ll, _vp, _i, _ll, _d, _d, _vp, _i, _ll, _i, _vp, _vp]),
    "rcwa_zgemm_batched_cfg": (_i, [_i, _i, _i, _i, _i, _i, _d, _d, _vp, _i, _ll, _vp, _i, _ll, _d, _d, _vp, _i, _ll, _i, _vp, _vp]),
    "rcwa_zgemm_tc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "rcwa_zgemm_tc_batched": (_i, [_i, _i, _i, _i, _i, _i, _d, _vp, _i, _ll, _vp, _i, _ll, _d, _d, _vp, _i, _ll, _i, _vp, _sz, _vp]),
    "rcwa_tc_split": (_i, [_vp, _i, _ll, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "rcwa_tc_schedule": (_i, [_i, _i, _vp, _vp, _vp, _vp]),
    "rcwa_tc_issue_entry": (_i, [_i, _i, _i, _i, _i, _vp, _vp]),
    "rcwa_set_tuning": (_i, [_i, _i]),
    "rcwa_get_tuning": (_i, [_i]),
    "rcwa_lu_tinv_bytes": (_sz, [_i, _i]),
    "rcwa_lu_factor": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rcwa_lu_solve_right": (_i, [_vp, _ll, _i, _i, _vp, _vp, _vp, _ll, _i, _i, _vp, _ll, _i, _vp, _i, _vp, _vp]),
    "rcwa_pq_assemble": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "rcwa_eig_workspace_bytes": (_sz, [_i, _i]),
    "rcwa_eig": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "rcwa_eig_phases": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp, _vp, _i, _vp]),
    "rcwa_eig_stats": (_i, [_vp, _i, _i, _vp, _vp]),
    "rcwa_eig_profile": (_i, [_vp, _i, _i, _vp, _vp]),
    "rcwa_hessenberg": (_i, [_vp, _i, _i, _vp, _vp, _sz, _vp]),
    "rcwa_hessenberg_matvec_probe": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp]),
    "rcwa_hessenberg_panel_width": (_i, []),
    "rcwa_kz_branch": (_i, [_vp, _vp, _ll, _vp]),
    "rcwa_eig_backward_workspace_bytes": (_sz, [_i, _i]),
    "rcwa_eig_backward": (_i, [_vp, _vp, _vp, _vp, _d, _i, _i, _vp, _vp, _vp, _vp]),
    "rcwa_layer_smatrix_workspace_bytes": (_sz, [_i, _i, _i]),
    "rcwa_layer_smatrix": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "rcwa_redheffer_workspace_bytes": (_sz, [_i, _i, _i]),
    "rcwa_redheffer": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "rcwa_redheffer_bdleft": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "rcwa_blockdiag_dense": (_i, [_vp, _i, _i, _vp, _vp]),
    "rcwa_sym_project": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
}

_lib = None


class RcwaB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (no CUDA context is created by loading)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "torcwa_b200: %s is missing. Build it with `python -m torcwa_b200.build` "
                "(needs nvcc, sm_100a). There is no fallback implementation." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.rcwa_b200_abi_version() != 2:
            raise ImportError("torcwa_b200: ABI version mismatch")
        _lib = lib
        # development knob: RCWA_B200_TUNE="0=1,1=2" -> rcwa_set_tuning(key, value) (see include/rcwa_b200.h)
        for kv in filter(None, os.environ.get("RCWA_B200_TUNE", "").split(",")):
            k, v = kv.split("=")
            lib.rcwa_set_tuning(int(k), int(v))
    return _lib


def _check(rc, what):
    if rc != 0:
        if rc <= -1000:
            raise RcwaB200Error("%s: CUDA error %d (%s)" % (what, -1000 - rc, "see cudaError_t"))
        raise RcwaB200Error("%s: invalid argument #%d" % (what, -rc))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(fn):
    """The library launches on the CURRENT device and stream: run every wrapper with the device of its first CUDA tensor
    argument current, so that a simulation on cuda:1 works whatever device the caller left current (the reference accepts
    any device index)."""
    import functools

    def first_cuda(args):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                return a.device
            if isinstance(a, (list, tuple)):
                d = first_cuda(a)
                if d is not None:
                    return d
        return None

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        dev = first_cuda(args) or first_cuda(list(kw.values()))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapped


def _c128(t, name):
    if not (t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA complex128 tensor" % name)
    return t


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


GRID_TYPES = {torch.float32: 0, torch.float64: 1, torch.complex64: 2, torch.complex128: 3}


# ------------------------------------------------------------------------------------------ wrappers
@_on_device
def convmat(grid, ox, oy, nb=None):
    """grid: [nx,ny] (shared) or [B,nx,ny]; -> E [B,N,N] complex128."""
    lib = load()
    if not grid.is_cuda:
        raise TypeError("grid must be a CUDA tensor")
    grid = grid.contiguous()
    if grid.dim() == 2:
        nbatch, stride = (nb or 1), 0
    else:
        nbatch, stride = grid.shape[0], grid.shape[-2] * grid.shape[-1]
    nx, ny = grid.shape[-2], grid.shape[-1]
    N = (2 * ox + 1) * (2 * oy + 1)
    E = torch.empty((nbatch, N, N), dtype=torch.complex128, device=grid.device)
    ws = _ws(lib.rcwa_convmat_workspace_bytes(nx, ny, nbatch, ox, oy), grid.device)
    _check(lib.rcwa_convmat(_ptr(grid), GRID_TYPES[grid.dtype], stride, nx, ny, nbatch, ox, oy, _ptr(E), _ptr(ws), _stream()),
           "rcwa_convmat")
    return E


_OPS = {"N": 0, "T": 1, "H": 2}


@_on_device
def zgemm(A, B, opa="N", opb="N", alpha=1.0, beta=0.0, out=None, cfg=None):
    """Batched C = alpha op(A) op(B) + beta C on [nb,*,*] complex128 tensors (cfg: explicit kernel configuration)."""
    lib = load()
    _c128(A, "A"); _c128(B, "B")
    nb = A.shape[0]
    M = A.shape[1] if opa == "N" else A.shape[2]
    K = A.shape[2] if opa == "N" else A.shape[1]
    N = B.shape[2] if opb == "N" else B.shape[1]
    if out is None:
        out = torch.empty((nb, M, N), dtype=torch.complex128, device=A.device)
        beta = 0.0
    _c128(out, "out")
    gs = _ws(lib.rcwa_gemm_scratch_bytes(nb), A.device)
    al, be = complex(alpha), complex(beta)
    if cfg is not None:
        _check(lib.rcwa_zgemm_batched_cfg(int(cfg), _OPS[opa], _OPS[opb], M, N, K, al.real, al.imag,
                                          _ptr(A), A.shape[2], A.shape[1] * A.shape[2], _ptr(B), B.shape[2], B.shape[1] * B.shape[2],
                                          be.real, be.imag, _ptr(out), N, M * N, nb, _ptr(gs), _stream()), "rcwa_zgemm_batched_cfg")
        return out
    _check(lib.rcwa_zgemm_batched(_OPS[opa], _OPS[opb], M, N, K, al.real, al.imag,
                                  _ptr(A), A.shape[2], A.shape[1] * A.shape[2], _ptr(B), B.shape[2], B.shape[1] * B.shape[2],
                                  be.real, be.imag, _ptr(out), N, M * N, nb, _ptr(gs), _stream()), "rcwa_zgemm_batched")
    return out


@_on_device
def zgemm_tc(A, B, opa="N", opb="N", alpha=1.0, beta=0.0, out=None, slices=7, ws_bytes=None):
    """The same product on tcgen05 (int8 digit products, include/rcwa_b200.h: rcwa_zgemm_tc_batched); alpha real."""
    lib = load()
    _c128(A, "A"); _c128(B, "B")
    nb = A.shape[0]
    M = A.shape[1] if opa == "N" else A.shape[2]
    K = A.shape[2] if opa == "N" else A.shape[1]
    N = B.shape[2] if opb == "N" else B.shape[1]
    if out is None:
        out = torch.empty((nb, M, N), dtype=torch.complex128, device=A.device)
        beta = 0.0
    _c128(out, "out")
    nbytes = int(ws_bytes) if ws_bytes else lib.rcwa_zgemm_tc_workspace_bytes(M, N, K, nb, int(slices))
    ws = _ws(nbytes, A.device)
    be = complex(beta)
    _check(lib.rcwa_zgemm_tc_batched(int(slices), _OPS[opa], _OPS[opb], M, N, K, float(alpha),
                                     _ptr(A), A.shape[2], A.shape[1] * A.shape[2], _ptr(B), B.shape[2], B.shape[1] * B.shape[2],
                                     be.real, be.imag, _ptr(out), N, M * N, nb, _ptr(ws), nbytes, _stream()), "rcwa_zgemm_tc_batched")
    return out


@_on_device
def tc_split(X, rows_contiguous, slices, conj=False):
    """Digit split of the tcgen05 GEMM on its own (tests): X [nb,a,b] -> (planes int8 [nb,3,slices,R,Kp], ex int32 [nb,R])."""
    lib = load()
    _c128(X, "X")
    nb = X.shape[0]
    R, Kc = (X.shape[1], X.shape[2]) if rows_contiguous else (X.shape[2], X.shape[1])
    Kp = (Kc + 127) // 128 * 128
    planes = torch.empty((nb, 3, slices, R, Kp), dtype=torch.int8, device=X.device)
    ex = torch.empty((nb, R), dtype=torch.int32, device=X.device)
    _check(lib.rcwa_tc_split(_ptr(X), X.shape[2], X.shape[1] * X.shape[2], int(bool(rows_contiguous)), R, Kc, int(slices), int(bool(conj)),
                             _ptr(planes), _ptr(ex), nb, _stream()), "rcwa_tc_split")
    return planes, ex


def tc_schedule(slices, levels=4):
    """Host-only: the per-K-chunk schedule of the tcgen05 GEMM -> (ops list, groups list of dicts incl. the compact
    per-role tables `loads` / `mmas` the kernel walks)."""
    lib = load()
    ops = (ctypes.c_uint * 128)()
    meta = (ctypes.c_int * 64)()
    loads = (ctypes.c_uint * (8 * 16))()
    mmas = (ctypes.c_uint * (8 * 32))()
    _check(lib.rcwa_tc_schedule(int(slices), int(levels), ops, meta, loads, mmas), "rcwa_tc_schedule")
    groups = []
    for g in range(meta[0]):
        d = dict(d0=meta[2 + 6 * g], nl=meta[3 + 6 * g], op0=meta[4 + 6 * g], nops=meta[5 + 6 * g], nloads=meta[6 + 6 * g], nmma=meta[7 + 6 * g])
        d["loads"] = list(loads)[16 * g:16 * g + d["nloads"]]
        d["mmas"] = list(mmas)[32 * g:32 * g + d["nmma"]]
        groups.append(d)
    return list(ops)[:meta[1]], groups


def tc_issue_entries(slices, group, ring_pos, levels=4):
    """Host-only: the decoded issue-table entries of every step of one level group at one ring position."""
    lib = load()
    out = []
    ns = ctypes.c_int(0)
    e = (ctypes.c_uint * 16)()
    step = 0
    while True:
        rc = lib.rcwa_tc_issue_entry(int(slices), int(levels), int(group), int(ring_pos), step, e, ctypes.byref(ns))
        if rc == -5:
            break
        _check(rc, "rcwa_tc_issue_entry")
        ng = e[1] & 15
        out.append(dict(a_slot=e[0], nacq=(e[1] >> 4) & 15, rel_b=((e[1] >> 17) & 255) if (e[1] >> 16) & 1 else None,
                        groups=[dict(b_slot=e[2 + 2 * j], col=e[3 + 2 * j] & 511, first=(e[3 + 2 * j] >> 9) & 1, double=(e[3 + 2 * j] >> 10) & 1)
                                for j in range(ng)]))
        step += 1
    return out


@_on_device
def lu_factor_(A):
    """In place on A [nb,n,n]; returns (perm, info, tinv)."""
    lib = load()
    _c128(A, "A")
    nb, n = A.shape[0], A.shape[1]
    ipiv = torch.empty((nb, n), dtype=torch.int32, device=A.device)
    perm = torch.empty((nb, n), dtype=torch.int32, device=A.device)
    info = torch.zeros((nb,), dtype=torch.int32, device=A.device)
    tinv = _ws(lib.rcwa_lu_tinv_bytes(n, nb), A.device)
    gs = _ws(lib.rcwa_gemm_scratch_bytes(nb), A.device)
    _check(lib.rcwa_lu_factor(_ptr(A), n * n, n, n, nb, _ptr(ipiv), _ptr(perm), _ptr(info), _ptr(tinv), _ptr(gs), _stream()), "rcwa_lu_factor")
    return perm, info, tinv


@_on_device
def lu_solve_right(LU, perm, tinv, Bm):
    """X = Bm @ inv(A) for factored LU [nb,n,n], Bm [nb,r,n]."""
    lib = load()
    _c128(LU, "LU"); _c128(Bm, "B")
    nb, n = LU.shape[0], LU.shape[1]
    r = Bm.shape[1]
    X = torch.empty_like(Bm)
    work = torch.empty_like(Bm)
    gs = _ws(lib.rcwa_gemm_scratch_bytes(nb), LU.device)
    _check(lib.rcwa_lu_solve_right(_ptr(LU), n * n, n, n, _ptr(perm), _ptr(tinv), _ptr(Bm), r * n, n, r, _ptr(X), r * n, n, _ptr(work),
                                   nb, _ptr(gs), _stream()), "rcwa_lu_solve_right")
    return X


def right_solve(Bm, A):
    """X = Bm @ inv(A) ([nb,r,n], [nb,n,n]); returns (X, info). A is preserved."""
    LU = A.clone()
    perm, info, tinv = lu_factor_(LU)
    return lu_solve_right(LU, perm, tinv, Bm), info


def inverse(A):
    """inv(A) for [nb,n,n] via X*A = I; returns (inv, info). A is preserved."""
    LU = A.clone()
    perm, info, tinv = lu_factor_(LU)
    eye = torch.eye(A.shape[1], dtype=A.dtype, device=A.device).expand(A.shape[0], -1, -1).contiguous()
    return lu_solve_right(LU, perm, tinv, eye), info


@_on_device
def pq_assemble(eta, E, kx, ky, mu_scalar=None, Mc=None, nu=None):
    lib = load()
    nb, N = E.shape[0], E.shape[1]
    P = torch.empty((nb, 2 * N, 2 * N), dtype=torch.complex128, device=E.device)
    Q = torch.empty_like(P)
    _check(lib.rcwa_pq_assemble(_ptr(_c128(eta, "eta")), _ptr(_c128(E, "E")), _ptr(Mc), _ptr(nu), _ptr(mu_scalar),
                                _ptr(_c128(kx, "kx")), _ptr(_c128(ky, "ky")), nb, N, _ptr(P), _ptr(Q), _stream()), "rcwa_pq_assemble")
    return P, Q


_tls = threading.local()       # pinned polling flags of rcwa_eig: one buffer per host thread (calls may run concurrently)


@_on_device
def eig(A, after_reduction=None):
    """A [nb,n,n] (destroyed) -> (w [nb,n], V [nb,n,n], info [nb]).  after_reduction: optional callable invoked on the host
    once the Hessenberg phase has been ENQUEUED on the current stream (the routine then runs as two calls, rcwa_eig_phases):
    the hook a pipelining host uses to release the next sub-batch (torcwa_b200/rcwa.py)."""
    lib = load()
    _c128(A, "A")
    nb, n = A.shape[0], A.shape[1]
    w = torch.empty((nb, n), dtype=torch.complex128, device=A.device)
    V = torch.empty((nb, n, n), dtype=torch.complex128, device=A.device)
    info = torch.zeros((nb,), dtype=torch.int32, device=A.device)
    nbytes = lib.rcwa_eig_workspace_bytes(n, nb)
    ws = _ws(nbytes, A.device)
    if getattr(_tls, "host_flag", None) is None:
        _tls.host_flag = torch.zeros(16, dtype=torch.int32).pin_memory()
    hf = ctypes.c_void_p(_tls.host_flag.data_ptr())
    if after_reduction is None:
        _check(lib.rcwa_eig(_ptr(A), n, nb, _ptr(w), _ptr(V), _ptr(ws), nbytes, _ptr(info), hf, _stream()), "rcwa_eig")
    else:
        try:
            _check(lib.rcwa_eig_phases(_ptr(A), n, nb, _ptr(w), _ptr(V), _ptr(ws), nbytes, _ptr(info), hf, 1, _stream()), "rcwa_eig_phases(1)")
        finally:
            after_reduction()
        _check(lib.rcwa_eig_phases(_ptr(A), n, nb, _ptr(w), _ptr(V), _ptr(ws), nbytes, _ptr(info), hf, 2, _stream()), "rcwa_eig_phases(2)")
    global last_eig_stats
    stats = torch.empty((nb, 4), dtype=torch.int32, device=A.device)
    _check(lib.rcwa_eig_stats(_ptr(ws), n, nb, _ptr(stats), _stream()), "rcwa_eig_stats")
    last_eig_stats = stats
    prof = torch.empty((nb, 6, 3), dtype=torch.int64, device=A.device)
    _check(lib.rcwa_eig_profile(_ptr(ws), n, nb, _ptr(prof), _stream()), "rcwa_eig_profile")
    global last_eig_profile
    last_eig_profile = prof
    return w, V, info


last_eig_stats = None
last_eig_profile = None      # [nb,6,3] int64: {launches, SM cycles, longest segment} per QR pass segment (rcwa_eig_profile)


@_on_device
def hessenberg_(A):
    """A [nb,n,n] -> Hessenberg in place; returns Z [nb,n,n] with A_in = Z H Z^H."""
    lib = load()
    _c128(A, "A")
    nb, n = A.shape[0], A.shape[1]
    Z = torch.empty_like(A)
    nbytes = lib.rcwa_eig_workspace_bytes(n, nb)
    ws = _ws(nbytes, A.device)
    _check(lib.rcwa_hessenberg(_ptr(A), n, nb, _ptr(Z), _ptr(ws), nbytes, _stream()), "rcwa_hessenberg")
    return Z


@_on_device
def matvec_probe(A, ws, j):
    """One launch of the Hessenberg streaming mat-vec for column j (profiling; A is only read)."""
    lib = load()
    nb, n = A.shape[0], A.shape[1]
    _check(lib.rcwa_hessenberg_matvec_probe(_ptr(_c128(A, "A")), n, nb, int(j), _ptr(ws), ws.numel(), _stream()), "rcwa_hessenberg_matvec_probe")


def eig_workspace(n, nb, device):
    lib = load()
    return _ws(lib.rcwa_eig_workspace_bytes(n, nb), device)


@_on_device
def eig_backward(lam, X, glam, gX, delta):
    """Gradient of the eigendecomposition (Eig.backward): lam [nb,n], X [nb,n,n], glam / gX or None -> (grad [nb,n,n], info)."""
    lib = load()
    nb, n = X.shape[0], X.shape[1]
    grad = torch.empty_like(X)
    info = torch.zeros((nb,), dtype=torch.int32, device=X.device)
    ws = _ws(lib.rcwa_eig_backward_workspace_bytes(n, nb), X.device)
    _check(lib.rcwa_eig_backward(_ptr(_c128(lam, "lam")), _ptr(_c128(X, "X")), _ptr(None if glam is None else _c128(glam, "glam")),
                                 _ptr(None if gX is None else _c128(gX, "gX")), float(delta), nb, n, _ptr(grad), _ptr(ws), _ptr(info), _stream()),
           "rcwa_eig_backward")
    return grad, info


@_on_device
def kz_branch(lam):
    lib = load()
    kz = torch.empty_like(lam)
    _check(lib.rcwa_kz_branch(_ptr(_c128(lam, "lam")), _ptr(kz), lam.numel(), _stream()), "rcwa_kz_branch")
    return kz


@_on_device
def layer_smatrix(W, kz, Q, vfinv, omega, thickness, slices=0):
    """-> (S11, S21, info) for the single layer (S22 = S11, S12 = S21).  slices: digits of the tcgen05 GEMM (0 = fp64 DMMA)."""
    lib = load()
    nb, n = W.shape[0], W.shape[1]
    N = n // 2
    S11 = torch.empty_like(W)
    S21 = torch.empty_like(W)
    info = torch.zeros((nb,), dtype=torch.int32, device=W.device)
    ws = _ws(lib.rcwa_layer_smatrix_workspace_bytes(N, nb, int(slices)), W.device)
    omega = omega.to(torch.float64).contiguous()
    thickness = thickness.to(torch.float64).contiguous()
    _check(lib.rcwa_layer_smatrix(_ptr(_c128(W, "W")), _ptr(_c128(kz, "kz")), _ptr(_c128(Q, "Q")), _ptr(_c128(vfinv, "vfinv")),
                                  _ptr(omega), _ptr(thickness), nb, N, _ptr(S11), _ptr(S21), _ptr(ws), _ptr(info), int(slices), _stream()),
           "rcwa_layer_smatrix")
    return S11, S21, info


@_on_device
def redheffer(Sm, Sn, slices=0):
    """Star product of two S-matrices given as lists [S11,S21,S12,S22] of [nb,n,n]; -> (list, info)."""
    lib = load()
    nb, n = Sm[0].shape[0], Sm[0].shape[1]
    out = [torch.empty_like(Sm[0]) for _ in range(4)]
    info = torch.zeros((nb,), dtype=torch.int32, device=Sm[0].device)
    ws = _ws(lib.rcwa_redheffer_workspace_bytes(n, nb, int(slices)), Sm[0].device)
    arr = ctypes.c_void_p * 4
    a_m = arr(*[t.data_ptr() for t in (_c128(x, "Sm") for x in Sm)])
    a_n = arr(*[t.data_ptr() for t in (_c128(x, "Sn") for x in Sn)])
    a_o = arr(*[t.data_ptr() for t in out])
    _check(lib.rcwa_redheffer(a_m, a_n, a_o, nb, n, _ptr(ws), _ptr(info), int(slices), _stream()), "rcwa_redheffer")
    return out, info


@_on_device
def redheffer_bdleft(Sm_bd, Sn, slices=0):
    """Star product with a 2x2-block-diagonal left factor: Sm_bd = four [nb,4,N] tensors, Sn dense."""
    lib = load()
    nb, n = Sn[0].shape[0], Sn[0].shape[1]
    out = [torch.empty_like(Sn[0]) for _ in range(4)]
    info = torch.zeros((nb,), dtype=torch.int32, device=Sn[0].device)
    ws = _ws(lib.rcwa_redheffer_workspace_bytes(n, nb, int(slices)), Sn[0].device)
    arr = ctypes.c_void_p * 4
    keep = [_c128(x.contiguous(), "Sm_bd") for x in Sm_bd]
    a_m = arr(*[t.data_ptr() for t in keep])
    a_n = arr(*[t.data_ptr() for t in (_c128(x, "Sn") for x in Sn)])
    a_o = arr(*[t.data_ptr() for t in out])
    _check(lib.rcwa_redheffer_bdleft(a_m, a_n, a_o, nb, n // 2, _ptr(ws), _ptr(info), int(slices), _stream()), "rcwa_redheffer_bdleft")
    return out, info


@_on_device
def sym_project(X, il, cl, ir, cr):
    """T_L^H X T_R for symmetry-adapted bases given as (index int32 [G,nk], coefficient complex128 [G,nk]) tables."""
    lib = load()
    nb, n = X.shape[0], X.shape[1]
    G, nkl, nkr = il.shape[0], il.shape[1], ir.shape[1]
    for t in (il, ir):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise ValueError("sym_project: index tables must be contiguous int32")
    out = torch.empty((nb, nkl, nkr), dtype=torch.complex128, device=X.device)
    _check(lib.rcwa_sym_project(_ptr(_c128(X, "X")), nb, n, _ptr(il), _ptr(_c128(cl, "cl")), _ptr(ir), _ptr(_c128(cr, "cr")),
                                G, nkl, nkr, _ptr(out), _stream()), "rcwa_sym_project")
    return out


@_on_device
def blockdiag_dense(d4):
    """d4 [nb,4,N] -> dense [nb,2N,2N]."""
    lib = load()
    nb, N = d4.shape[0], d4.shape[2]
    D = torch.empty((nb, 2 * N, 2 * N), dtype=torch.complex128, device=d4.device)
    _check(lib.rcwa_blockdiag_dense(_ptr(_c128(d4, "d4")), nb, N, _ptr(D), _stream()), "rcwa_blockdiag_dense")
    return D
