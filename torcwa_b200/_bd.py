"""Per-order 2x2 algebra for matrices made of four diagonal N x N blocks.

Half-space E->H matrices (Vf, Vi, Vo), their S-matrices (Sin, Sout) and everything about a
homogeneous layer have the form [[diag a, diag b],[diag c, diag d]] (reference:
torcwa/rcwa.py:1143-1181 builds them densely and inverts them with LU).  Here such a matrix is a
tensor [..., 4, N] holding (a, b, c, d); products and inverses are O(N) elementwise work, done with
torch on the device as host-side plumbing.  `_lib.blockdiag_dense` scatters one into a dense
[2N, 2N] matrix when a dense operand is needed.
"""
import torch


def bd(a, b, c, d):
    return torch.stack((a, b, c, d), dim=-2)


def bd_mul(x, y):
    a, b, c, d = x.unbind(-2)
    e, f, g, h = y.unbind(-2)
    return bd(a * e + b * g, a * f + b * h, c * e + d * g, c * f + d * h)


def bd_inv(x):
    a, b, c, d = x.unbind(-2)
    det = a * d - b * c
    return bd(d / det, -b / det, -c / det, a / det)


def bd_diag(v_top, v_bot=None):
    """diag(v) as a block-diagonal matrix (v of length 2N split in halves, or top==bottom)."""
    if v_bot is None:
        v_bot = v_top
    z = torch.zeros_like(v_top)
    return bd(v_top, z, z, v_bot)


def bd_eye_like(x):
    one = torch.ones_like(x[..., 0, :])
    return bd_diag(one)


def v_matrix(kx, ky, kz):
    """E->H matrix of a homogeneous medium (torcwa/rcwa.py:1145-1147)."""
    return bd(-ky * kx / kz, -kz - ky * ky / kz, kz + kx * kx / kz, kx * ky / kz)


def sqrt_upper(z):
    """sqrt, conjugated where Im < 0 (the reference's branch for half spaces and homogeneous
    layers, torcwa/rcwa.py:1143-1144, :1217-1218)."""
    r = torch.sqrt(z)
    return torch.where(r.imag < 0, r.conj(), r)
