"""CPU test double of torcwa_b200._lib (TEST INFRASTRUCTURE).

Same function signatures as the ctypes wrappers, implemented with torch fp64 on the CPU following
the *new* minimal algebra (right-solves, symmetric layer S, one-LU Redheffer), so that the host
logic of torcwa_b200/rcwa.py (batching, k-vectors, 2x2-block half spaces, homogeneous layers,
cascade order, S-parameter readout) can be checked against the oracle without a GPU.  Installed
only by the `cpu_double` fixture below; the product never imports this file."""
import pytest
import torch


def load():
    return None


def convmat(grid, ox, oy, nb=None):
    grid = grid.to(torch.complex128)
    if grid.dim() == 2:
        grid = grid[None].expand(nb or 1, -1, -1)
    nx, ny = grid.shape[-2:]
    F = torch.fft.fft2(grid) / (nx * ny)
    mx = torch.arange(-ox, ox + 1).repeat_interleave(2 * oy + 1)
    my = torch.arange(-oy, oy + 1).repeat(2 * ox + 1)
    return F[:, (mx[:, None] - mx[None, :]) % nx, (my[:, None] - my[None, :]) % ny].contiguous()


def inverse(A):
    return torch.linalg.inv(A), torch.zeros(A.shape[0], dtype=torch.int32)


def right_solve(Bm, A):
    return Bm @ torch.linalg.inv(A), torch.zeros(A.shape[0], dtype=torch.int32)


def eig_backward(lam, X, glam, gX, delta):
    """Batched restatement of torch_eig.py:19-44 (what rcwa_eig_backward computes)."""
    s = lam.unsqueeze(-2) - lam.unsqueeze(-1)
    Fc = s / (torch.abs(s) ** 2 + delta)                        # conj(F)
    n = lam.shape[-1]
    Fc = Fc * (1 - torch.eye(n, dtype=Fc.dtype))
    XH = X.transpose(-2, -1).conj()
    M = torch.zeros_like(X)
    if gX is not None:
        M = Fc * (XH @ gX)
    if glam is not None:
        M = M + torch.diag_embed(glam)
    return torch.linalg.inv(XH) @ M @ XH, torch.zeros(X.shape[0], dtype=torch.int32)


def zgemm(A, B, opa="N", opb="N", alpha=1.0, beta=0.0, out=None):
    f = {"N": lambda x: x, "T": lambda x: x.transpose(1, 2), "H": lambda x: x.transpose(1, 2).conj()}
    r = alpha * (f[opa](A) @ f[opb](B))
    if out is not None:
        out.copy_(r + beta * out)
        return out
    return r


def pq_assemble(eta, E, kx, ky, mu_scalar=None, Mc=None, nu=None):
    N = E.shape[1]
    if Mc is None:
        eye = torch.eye(N, dtype=E.dtype)
        Mc, nu = mu_scalar[:, None, None] * eye, (1 / mu_scalar)[:, None, None] * eye
    kxr, kxc, kyr, kyc = kx[:, :, None], kx[:, None, :], ky[:, :, None], ky[:, None, :]
    P = torch.cat((torch.cat((kxr * eta * kyc, Mc - kxr * eta * kxc), 2), torch.cat((kyr * eta * kyc - Mc, -kyr * eta * kxc), 2)), 1)
    Q = torch.cat((torch.cat((-kxr * nu * kyc, kxr * nu * kxc - E), 2), torch.cat((E - kyr * nu * kyc, kyr * nu * kxc), 2)), 1)
    return P.contiguous(), Q.contiguous()


def eig(A, after_reduction=None):
    """torch.linalg.eig per matrix, with one documented property of rcwa_eig that the host relies on: leading rows / columns
    that are exactly decoupled (the extension of a symmetry block, rcwa._patterned_layer_blocks) stay in place -- their
    eigenvalues come first, the eigenvector matrix is diag(I, W)."""
    nb, n = A.shape[0], A.shape[1]
    w = torch.zeros((nb, n), dtype=A.dtype)
    V = torch.zeros_like(A)
    for b in range(nb):
        d = 0
        while d < n - 1 and not bool(A[b, d, d + 1:].any()) and not bool(A[b, d + 1:, d].any()):
            d += 1
        wt, Vt = torch.linalg.eig(A[b, d:, d:])
        w[b, :d], w[b, d:] = torch.diagonal(A[b])[:d], wt
        V[b, range(d), range(d)] = 1.0
        V[b, d:, d:] = Vt
    if after_reduction is not None:
        after_reduction()
    return w, V, torch.zeros(nb, dtype=torch.int32)


def kz_branch(lam):
    r = torch.sqrt(lam)
    return torch.where(r.imag < 0, -r, r)


def sym_project(X, il, cl, ir, cr):
    """T_L^H X T_R from the (index, coefficient) tables, in plain torch (what rcwa_sym_project computes in one pass)."""
    Y = sum(X.index_select(1, il[t].long()) * cl[t].conj()[None, :, None] for t in range(il.shape[0]))
    return sum(Y.index_select(2, ir[t].long()) * cr[t][None, None, :] for t in range(ir.shape[0])).contiguous()


def blockdiag_dense(d4):
    a, b, c, d = (torch.diag_embed(d4[:, k]) for k in range(4))
    return torch.cat((torch.cat((a, b), 2), torch.cat((c, d), 2)), 1)


def layer_smatrix(W, kz, Q, vfinv, omega, thickness, slices=0):
    n = W.shape[1]
    V = (Q @ W) / kz[:, None, :]
    Bm = blockdiag_dense(vfinv) @ V
    X = torch.exp(1j * (omega * thickness)[:, None] * kz)[:, None, :]
    Rp, Rm = W * (1 + X), W * (X - 1)
    Mp, Mm = Rp + Bm * (1 - X), W * (1 - X) + Bm * (1 + X)
    Tp, Tm = Rp @ torch.linalg.inv(Mp), Rm @ torch.linalg.inv(Mm)
    return Tp + Tm, Tp - Tm - torch.eye(n, dtype=W.dtype), torch.zeros(W.shape[0], dtype=torch.int32)


def redheffer(Sm, Sn, slices=0):
    n = Sm[0].shape[1]
    Di = torch.linalg.inv(torch.eye(n, dtype=Sm[0].dtype) - Sm[2] @ Sn[1])
    Y1, Y2, G = Sn[0] @ Di, Sn[1] @ Di, Sm[2] @ Sn[3]
    return [Y1 @ Sm[0], Sm[1] + Sm[3] @ (Y2 @ Sm[0]), Sn[2] + Y1 @ G, Sm[3] @ (Sn[3] + Y2 @ G)], torch.zeros(Sm[0].shape[0], dtype=torch.int32)


def redheffer_bdleft(Sm_bd, Sn, slices=0):
    return redheffer([blockdiag_dense(x) for x in Sm_bd], Sn)


@pytest.fixture
def cpu_double(monkeypatch):
    import sys
    import torcwa_b200  # noqa: F401
    host = sys.modules['torcwa_b200.rcwa']      # the module (the package attribute `rcwa` is the class)
    monkeypatch.setattr(host, "_lib", sys.modules[__name__])
    monkeypatch.setattr(sys.modules['torcwa_b200.autodiff'], "_lib", sys.modules[__name__])
    monkeypatch.setattr(sys.modules['torcwa_b200.torch_eig'], "_lib", sys.modules[__name__])
    monkeypatch.setattr(host, "_TEST_ALLOW_NON_CUDA", True)
    return host
