"""ncu target: a few launches of the tcgen05 GEMM at the path's size (python tools/tc_ncu_target.py [slices] [n] [nb])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torcwa_b200 import _lib
s = int(sys.argv[1]) if len(sys.argv) > 1 else 7
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1922
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g = torch.Generator().manual_seed(0)
A = torch.complex(torch.randn(nb, n, n, generator=g, dtype=torch.float64), torch.randn(nb, n, n, generator=g, dtype=torch.float64)).cuda()
B = A.transpose(1, 2).contiguous()
C = torch.empty_like(A)
for _ in range(3):
    _lib.zgemm_tc(A, B, slices=s, out=C)
torch.cuda.synchronize()
