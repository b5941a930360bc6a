"""`Eig`: the reference's eigendecomposition entry point (torcwa/torch_eig.py:8-44), forward and backward.

`Eig.apply(A)` returns (eigenvalues, eigenvectors) of a complex matrix (or a batch) computed by the CUDA
eigensolver (C ABI rcwa_eig).  The backward is the reference's Lorentzian-broadened formula
(torch_eig.py:19-44) on the CUDA GEMM / LU kernels (C ABI rcwa_eig_backward), with the same global knob
`Eig.broadening_parameter` (torch_eig.py:9; None = smallest denormal of the input precision, :30-33).
Arithmetic is complex128 also for complex64 inputs; gradients are rounded to the input dtype,
and cast to real for a real input as the reference does (:41-42).
"""
import torch

from . import _lib


class Eig(torch.autograd.Function):
    broadening_parameter = 1e-10

    @staticmethod
    def forward(ctx, x):
        batched = x.dim() == 3
        A = (x if batched else x[None]).to(torch.complex128).contiguous().clone()
        w, V, info = _lib.eig(A)
        if int(info.abs().max()) != 0:
            raise torch.linalg.LinAlgError('rcwa_eig: QR iteration did not converge for batch entries %s'
                                           % torch.nonzero(info).flatten().tolist())
        ctx.batched, ctx.in_dtype, ctx.in_complex = batched, x.dtype, torch.is_complex(x)
        ctx.save_for_backward(w, V)                      # complex128, on the device (the reference parks them on the CPU, :15-16)
        out_dtype = x.dtype if torch.is_complex(x) else (torch.complex64 if x.dtype == torch.float32 else torch.complex128)
        w, V = w.to(out_dtype), V.to(out_dtype)
        return (w, V) if batched else (w[0], V[0])

    @staticmethod
    def backward(ctx, grad_eigval, grad_eigvec):
        w, V = ctx.saved_tensors
        if Eig.broadening_parameter is not None:
            delta = float(Eig.broadening_parameter)
        else:                                            # torch_eig.py:30-33
            delta = 1.4e-45 if ctx.in_dtype in (torch.complex64, torch.float32) else 4.9e-324

        def widen(g):
            if g is None:
                return None
            g = g if ctx.batched else g[None]
            return g.to(torch.complex128).contiguous()
        grad, info = _lib.eig_backward(w, V, widen(grad_eigval), widen(grad_eigvec), delta)
        if int(info.abs().max()) != 0:
            raise torch.linalg.LinAlgError('rcwa_eig_backward: singular eigenvector matrix for batch entries %s'
                                           % torch.nonzero(info).flatten().tolist())
        if not ctx.in_complex:
            grad = grad.real
        grad = grad.to(ctx.in_dtype)
        return grad if ctx.batched else grad[0]
