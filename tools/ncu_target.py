"""Small, fixed workloads for `ncu --set full -k regex:<kernel>` captures (run under gpurun).
    python tools/ncu_target.py hess|gemm|qr [--n 1922] [--nb 4]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torcwa_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["hess", "gemm", "qr"])
ap.add_argument("--n", type=int, default=1922)
ap.add_argument("--nb", type=int, default=4)
a = ap.parse_args()
d = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
A = torch.complex(torch.randn(a.nb, a.n, a.n, generator=g, dtype=torch.float64), torch.randn(a.nb, a.n, a.n, generator=g, dtype=torch.float64)).to(d)
if a.what == "hess":
    _lib.hessenberg_(A)
elif a.what == "gemm":
    B = A.clone()
    _lib.zgemm(A, B)                                        # full n^3 product (layer-S / Redheffer / P*Q shape)
    U = A[:, :64, :64].contiguous()
    P = A[:, :, :64].contiguous()
    _lib.zgemm(P, U)                                        # K = 64 panel update (QR sweep shape)
else:
    _lib.eig(A)
torch.cuda.synchronize()
print("done", a.what)
