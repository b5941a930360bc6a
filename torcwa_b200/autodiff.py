"""Differentiable layer pipeline: autograd wrappers around the C-ABI kernels.

The reference gets gradients for free because every step of `rcwa.add_layer` /
`solve_global_smatrix` is a PyTorch op (torcwa/rcwa.py:1183-1294) and only the eigendecomposition
has a hand-written backward (torcwa/torch_eig.py:19-44).  Here the forward path is fused CUDA
kernels without autograd, so when a layer's material (or thickness) requires a gradient the same
algebra is composed from differentiable primitives instead:

  * dense O(n^3) work -- products, right-solves X = B A^-1, the eigendecomposition and its gradient --
    stays on the CUDA kernels (`rcwa_zgemm_batched`, `rcwa_lu_factor/solve_right`, `rcwa_eig`,
    `rcwa_eig_backward`), each wrapped in a `torch.autograd.Function` whose backward is again
    those kernels;
  * O(n^2) elementwise assembly (P/Q scalings, the X = exp(i w kz d) factors, the four-diagonal
    half-space algebra) is written with torch ops, which differentiate themselves;
  * the convolution matrix uses `rcwa_convmat` forward and its adjoint (scatter-add of the Toeplitz
    gather + inverse DFT) backward.

Arithmetic is complex128 throughout, as in the fused path.  Gradients follow
PyTorch's convention for complex tensors (grad = conj Wirtinger derivative), so they can be compared
with the reference's autograd directly.
"""
import torch

from . import _lib
from .torch_eig import Eig

_C = torch.complex128


class ZGemm(torch.autograd.Function):
    """C = A @ B on [nb,m,k] x [nb,k,n] complex128."""

    @staticmethod
    def forward(ctx, A, B):
        A, B = A.contiguous(), B.contiguous()
        ctx.save_for_backward(A, B)
        return _lib.zgemm(A, B)

    @staticmethod
    def backward(ctx, gC):
        A, B = ctx.saved_tensors
        gC = gC.contiguous()
        gA = _lib.zgemm(gC, B, "N", "H") if ctx.needs_input_grad[0] else None      # gC B^H
        gB = _lib.zgemm(A, gC, "H", "N") if ctx.needs_input_grad[1] else None      # A^H gC
        return gA, gB


class RightSolve(torch.autograd.Function):
    """X = Bm @ inv(A) (Bm [nb,r,n], A [nb,n,n]) by the row-major right-solve LU.
    dX = dB A^-1 - X dA A^-1   =>   gB = gX A^-H,  gA = -X^H gB."""

    @staticmethod
    def forward(ctx, Bm, A):
        X, info = _lib.right_solve(Bm.contiguous(), A.contiguous())
        if int(info.abs().max()) != 0:
            raise torch.linalg.LinAlgError('right-solve: singular matrix for batch entries %s' % torch.nonzero(info).flatten().tolist())
        ctx.save_for_backward(X, A)
        return X

    @staticmethod
    def backward(ctx, gX):
        X, A = ctx.saved_tensors
        AH = A.transpose(1, 2).conj().contiguous()
        gB, _ = _lib.right_solve(gX.contiguous(), AH)                                # gX A^-H
        gA = -_lib.zgemm(X, gB, "H", "N") if ctx.needs_input_grad[1] else None
        return (gB if ctx.needs_input_grad[0] else None), gA


class ConvMat(torch.autograd.Function):
    """E[b,i,j] = F[b, (mi-mj) mod nx, (ni-nj) mod ny], F = fft2(grid)/(nx ny) (rcwa.py:1183-1204).
    The map is linear; its adjoint is scatter-add of the gather followed by the inverse DFT."""

    @staticmethod
    def forward(ctx, grid, ox, oy, nb):
        ctx.shape, ctx.ox, ctx.oy, ctx.nb = tuple(grid.shape), ox, oy, nb
        ctx.real_in = not torch.is_complex(grid)
        ctx.in_dtype = grid.dtype
        return _lib.convmat(grid.detach(), ox, oy, nb=nb)

    @staticmethod
    def backward(ctx, gE):
        nx, ny = ctx.shape[-2], ctx.shape[-1]
        ox, oy = ctx.ox, ctx.oy
        dev = gE.device
        mx = torch.arange(-ox, ox + 1, device=dev).repeat_interleave(2 * oy + 1)
        my = torch.arange(-oy, oy + 1, device=dev).repeat(2 * ox + 1)
        ix = ((mx[:, None] - mx[None, :]) % nx).reshape(-1)
        iy = ((my[:, None] - my[None, :]) % ny).reshape(-1)
        B = gE.shape[0]
        gF = torch.zeros((B, nx * ny), dtype=_C, device=dev)
        gF.index_add_(1, ix * ny + iy, gE.reshape(B, -1).to(_C))
        g = torch.fft.ifft2(gF.reshape(B, nx, ny))            # (fft2 / (nx ny))^H = ifft2
        if len(ctx.shape) == 2:                               # one grid shared by the batch
            g = g.sum(dim=0)
        if ctx.real_in:
            g = g.real
        return g.to(ctx.in_dtype), None, None, None


def zgemm(A, B):
    return ZGemm.apply(A, B)


def right_solve(Bm, A):
    return RightSolve.apply(Bm, A)


def inverse(A):
    eye = torch.eye(A.shape[-1], dtype=_C, device=A.device).expand(A.shape[0], -1, -1).contiguous()
    return RightSolve.apply(eye, A)


def blockdiag_dense(d4):
    """[nb,4,N] four diagonals -> dense [nb,2N,2N] with torch ops (differentiable twin of rcwa_blockdiag_dense)."""
    a, b, c, d = (torch.diag_embed(d4[:, k]) for k in range(4))
    return torch.cat((torch.cat((a, b), 2), torch.cat((c, d), 2)), 1)


def bd_left_mul(d4, X):
    """(four-diagonal matrix) @ X without forming it: rows of the two halves are scaled and combined."""
    N = d4.shape[-1]
    a, b, c, d = (d4[:, k][:, :, None] for k in range(4))
    top, bot = X[:, :N], X[:, N:]
    return torch.cat((a * top + b * bot, c * top + d * bot), dim=1)


def pq_assemble(eta, E, kx, ky, Mc, nu):
    """P, Q of rcwa.py:1224-1232 (Mc = mu convolution matrix, nu = its inverse): the reference's dense
    diag(K) products are row / column scalings."""
    kxr, kxc, kyr, kyc = kx[:, :, None], kx[:, None, :], ky[:, :, None], ky[:, None, :]
    P = torch.cat((torch.cat((kxr * eta * kyc, Mc - kxr * eta * kxc), 2), torch.cat((kyr * eta * kyc - Mc, -kyr * eta * kxc), 2)), 1)
    Q = torch.cat((torch.cat((-kxr * nu * kyc, kxr * nu * kxc - E), 2), torch.cat((E - kyr * nu * kyc, kyr * nu * kxc), 2)), 1)
    return P, Q


def patterned_layer(E, Mc, nu, kx, ky, vfinv, omega, thick):
    """One patterned layer from its convolution matrices (E, Mc = mu, nu = Mc^-1; [nb,N,N]):
    returns (S11, S21, kz, W, P, Q).  Same minimal algebra as rcwa_layer_smatrix:
    V = Q W Kz^-1, two right-solves."""
    eta = inverse(E)
    P, Q = pq_assemble(eta, E, kx, ky, Mc, nu)
    A = zgemm(P, Q)
    lam, W = Eig.apply(A)
    kz = torch.sqrt(lam)
    kz = torch.where(kz.imag < 0, -kz, kz)                                   # rcwa.py:1240-1241
    n = W.shape[1]
    V = zgemm(Q, W) / kz[:, None, :]
    Bm = bd_left_mul(vfinv, V)
    X = torch.exp(1j * (omega * thick)[:, None] * kz)[:, None, :]
    Rp, Rm = W * (1 + X), W * (X - 1)
    Mp, Mm = Rp + Bm * (1 - X), W * (1 - X) + Bm * (1 + X)
    Tp, Tm = right_solve(Rp, Mp), right_solve(Rm, Mm)
    eye = torch.eye(n, dtype=_C, device=W.device)
    return Tp + Tm, Tp - Tm - eye, kz, W, P, Q


def redheffer(Sm, Sn):
    """Star product (rcwa.py:1283-1294) with one LU: D = I - Sm12 Sn21, [Y1; Y2] = [Sn11; Sn21] D^-1."""
    n = Sm[0].shape[1]
    eye = torch.eye(n, dtype=_C, device=Sm[0].device)
    D = eye - zgemm(Sm[2], Sn[1])
    Y = right_solve(torch.cat((Sn[0], Sn[1]), dim=1), D)
    Y1, Y2 = Y[:, :n], Y[:, n:]
    G = zgemm(Sm[2], Sn[3])
    return [zgemm(Y1, Sm[0]), Sm[1] + zgemm(Sm[3], zgemm(Y2, Sm[0])), Sn[2] + zgemm(Y1, G), zgemm(Sm[3], Sn[3] + zgemm(Y2, G))]
