"""`Eig`: the reference's eigendecomposition entry point (torcwa/torch_eig.py:8-44), forward only.

`Eig.apply(A)` returns (eigenvalues, eigenvectors) of a complex matrix (or a batch) computed by
the CUDA eigensolver (C ABI rcwa_eig).  `Eig.broadening_parameter` is kept as the global knob the
reference exposes (torch_eig.py:9); the Lorentzian-broadened backward that uses it belongs to the
autograd row of the scope table and is not built yet -- calling backward raises.
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
        w, V = w.to(x.dtype), V.to(x.dtype)
        return (w, V) if batched else (w[0], V[0])

    @staticmethod
    def backward(ctx, grad_eigval, grad_eigvec):
        raise NotImplementedError('torcwa_b200.Eig backward (SURVEY.md 8a11) is not implemented in this round')
