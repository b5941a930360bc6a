"""Quick device timings of the dense building blocks (run under gpurun)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torcwa_b200 import _lib

def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

d = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
res = {}
for n, nb in ((1922, 8), (1922, 1), (512, 32)):
    A = torch.randn(nb, n, n, dtype=torch.complex128, device=d); B = torch.randn(nb, n, n, dtype=torch.complex128, device=d)
    t = timeit(lambda: _lib.zgemm(A, B))
    t2 = timeit(lambda: torch.matmul(A, B))
    fl = 8.0 * n ** 3 * nb
    res[f"zgemm_n{n}_b{nb}"] = dict(ms=t, tflops=fl / t / 1e9, cublas_ms=t2, cublas_tflops=fl / t2 / 1e9)
    print(f"zgemm n={n} nb={nb}: ours {t:.2f} ms {fl/t/1e9:.1f} TF | cuBLAS {t2:.2f} ms {fl/t2/1e9:.1f} TF", flush=True)
    # K = 32 / 64 rank updates (LU trailing / QR sweeps)
    for K in (32, 64):
        A2 = A[:, :, :K].contiguous(); B2 = B[:, :K, :].contiguous(); C = torch.zeros(nb, n, n, dtype=torch.complex128, device=d)
        t = timeit(lambda: _lib.zgemm(A2, B2, beta=1.0, alpha=-1.0, out=C))
        print(f"  rank-{K} update: {t:.3f} ms {8.0*n*n*K*nb/t/1e9:.1f} TF, {nb*n*n*32/t/1e6:.0f} GB/s C traffic", flush=True)
for n, nb in ((1922, 8), (1922, 1), (961, 8)):
    A = torch.randn(nb, n, n, dtype=torch.complex128, device=d) + 3 * torch.eye(n, dtype=torch.complex128, device=d)
    def f():
        LU = A.clone(); perm, info, tinv = _lib.lu_factor_(LU); return LU, perm, tinv
    t = timeit(f, n=2)
    LU, perm, tinv = f()
    Bm = torch.randn(nb, n, n, dtype=torch.complex128, device=d)
    t2 = timeit(lambda: _lib.lu_solve_right(LU, perm, tinv, Bm), n=2)
    t3 = timeit(lambda: torch.linalg.inv(A), n=2)
    print(f"LU n={n} nb={nb}: factor {t:.1f} ms, solve(n rhs) {t2:.1f} ms | torch inv {t3:.1f} ms", flush=True)
# fp64 real GEMM peak via cuBLAS for the roofline denominator
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=d); b = torch.randn(n, n, dtype=torch.float64, device=d)
t = timeit(lambda: torch.matmul(a, b), n=3)
print(f"cuBLAS dgemm {n}: {t:.1f} ms {2.0*n**3/t/1e9:.1f} TF")
