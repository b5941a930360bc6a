"""Kernel-configuration table of the DMMA complex GEMM on the shapes the RCWA path issues (run under gpurun).

For every (shape, cfg) it checks the result against torch.matmul (cuBLAS) and times it with CUDA events.
cfg = tile | 8*m3 (include/rcwa_b200.h).  Output: one JSON line per shape to stdout (and --out file)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torcwa_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--n", type=int, default=1922)
a = ap.parse_args()
d = torch.device("cuda:0")
n = a.n
g = torch.Generator(device="cpu").manual_seed(1)


def rnd(*shape):
    return torch.complex(torch.randn(*shape, generator=g, dtype=torch.float64), torch.randn(*shape, generator=g, dtype=torch.float64)).to(d)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# (label, opa, opb, M, N, K, batch, beta)
SHAPES = [
    ("qr_z_update   Z[:,win]*U", "N", "N", n, 64, 64, 192, 0.0),
    ("qr_row_update U^H*H[win,:]", "H", "N", 64, n // 2, 64, 96, 0.0),
    ("rank32_update (LU/Hess trailing)", "N", "N", n, n, 32, 16, 1.0),
    ("rank64_update", "N", "N", n, n, 64, 16, 1.0),
    ("rank128_update (solve blocks)", "N", "N", n, n, 128, 16, 1.0),
    ("hess_right NB=32  A-=Y*V^H", "N", "H", n, n, 32, 16, 1.0),
    ("hess_right NB=64  A-=Y*V^H", "N", "H", n, n, 64, 16, 1.0),
    ("skinny N=32  A*V", "N", "N", n, 32, n, 32, 0.0),
    ("skinny N=64  A*V", "N", "N", n, 64, n, 32, 0.0),
    ("skinny M=32  V^H*A", "H", "N", 32, n, n, 32, 0.0),
    ("skinny M=64  V^H*A", "H", "N", 64, n, n, 32, 0.0),
    ("full n^3", "N", "N", n, n, n, 8, 0.0),
]
CFGS = [0, 1, 2, 3, 4, 8, 9, 10, 11, 12]
NAMES = {0: "64x128", 1: "128x64", 2: "64x64", 3: "128x32", 4: "32x128"}

print(torch.cuda.get_device_name(0), flush=True)
rows = []
for label, opa, opb, M, N, K, nb, beta in SHAPES:
    A = rnd(nb, M, K) if opa == "N" else rnd(nb, K, M)
    B = rnd(nb, K, N) if opb == "N" else rnd(nb, N, K)
    C0 = rnd(nb, M, N)
    opA = A if opa == "N" else A.conj().transpose(1, 2)
    opB = B if opb == "N" else B.conj().transpose(1, 2)
    ref = torch.matmul(opA, opB) + beta * C0
    t_cublas = timeit(lambda: torch.matmul(opA, opB))
    fl = 8.0 * M * N * K * nb
    row = {"shape": label, "M": M, "N": N, "K": K, "batch": nb, "cublas_tflops": round(fl / t_cublas / 1e9, 2), "cfg": {}}
    scale = float(ref.abs().max())
    for cfg in CFGS:
        out = C0.clone()
        try:
            _lib.zgemm(A, B, opa=opa, opb=opb, beta=beta, out=out, cfg=cfg)
            torch.cuda.synchronize()
        except Exception as ex:
            row["cfg"]["%s%s" % (NAMES[cfg & 7], "+m3" if cfg & 8 else "")] = {"error": str(ex)[:60]}
            continue
        err = float((out - ref).abs().max()) / scale
        t = timeit(lambda: _lib.zgemm(A, B, opa=opa, opb=opb, beta=beta, out=out, cfg=cfg))
        row["cfg"]["%s%s" % (NAMES[cfg & 7], "+m3" if cfg & 8 else "")] = {"tflops": round(fl / t / 1e9, 2), "ms": round(t, 4), "err": float("%.2e" % err)}
    auto = C0.clone()
    t = timeit(lambda: _lib.zgemm(A, B, opa=opa, opb=opb, beta=beta, out=auto))
    row["auto_tflops"] = round(fl / t / 1e9, 2)
    best = max(((k, v["tflops"]) for k, v in row["cfg"].items() if "tflops" in v), key=lambda kv: kv[1])
    row["best"] = best
    rows.append(row)
    print(json.dumps(row), flush=True)
    del A, B, C0, ref, out, auto
if a.out:
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)
