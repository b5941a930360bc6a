"""Bring-up diagnostics of the tcgen05 GEMM: tiny structured cases whose wrong answers show WHAT is wrong
(swizzle / descriptor / level weights).  Each case in its own process with a timeout.

    python tools/tc_debug.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r"""
import sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
from torcwa_b200 import _lib
np.set_printoptions(linewidth=220, precision=4, suppress=True)
M, N, K, s, kind = %(M)d, %(N)d, %(K)d, %(s)d, %(kind)r
g = np.random.default_rng(0)
if kind == 'eyeB':            # C = A : shows a K permutation (swizzle / descriptor advance) directly
    A = (np.arange(M)[:, None] * 1.0 + np.arange(K)[None, :] / 1024.0) + 0j
    B = np.eye(K, N) + 0j
elif kind == 'ones':          # C[i,j] = K
    A = np.ones((M, K)) + 0j
    B = np.ones((K, N)) + 0j
elif kind == 'real':
    A = g.integers(-3, 4, (M, K)) + 0j
    B = g.integers(-3, 4, (K, N)) + 0j
else:
    A = g.standard_normal((M, K)) + 1j * g.standard_normal((M, K))
    B = g.standard_normal((K, N)) + 1j * g.standard_normal((K, N))
ref = A @ B
At = torch.from_numpy(A[None]).to('cuda:0'); Bt = torch.from_numpy(B[None]).to('cuda:0')
C = _lib.zgemm_tc(At, Bt, slices=s)
torch.cuda.synchronize()
C = C.cpu().numpy()[0]
err = np.max(np.abs(C - ref)) / max(np.max(np.abs(ref)), 1e-300)
print('CASE', kind, M, N, K, 's=%%d' %% s, 'relerr %%.3e' %% err, 'OK' if err < 1e-6 else 'WRONG')
if err >= 1e-6:
    print('got  re[:4,:8]\n', C.real[:4, :8]); print('want re[:4,:8]\n', ref.real[:4, :8])
    print('got  im[:4,:8]\n', C.imag[:4, :8]); print('want im[:4,:8]\n', ref.imag[:4, :8])
    bad = np.argwhere(np.abs(C - ref) > 1e-6 * np.max(np.abs(ref)))
    print('bad entries', len(bad), 'of', C.size, 'first', bad[:6].tolist(), 'rows', sorted(set(bad[:, 0].tolist()))[:12], 'cols', sorted(set(bad[:, 1].tolist()))[:12])
"""


def main():
    cases = [(128, 128, 32, 2, 'ones'), (128, 128, 32, 2, 'eyeB'), (128, 128, 128, 2, 'eyeB'), (128, 128, 128, 3, 'real'),
             (128, 128, 128, 7, 'rand'), (128, 16, 32, 4, 'rand'), (130, 140, 150, 7, 'rand'), (300, 300, 300, 8, 'rand'),
             (128, 128, 2048, 5, 'rand')]
    for (M, N, K, s, kind) in cases:
        code = CASE % dict(root=ROOT, M=M, N=N, K=K, s=s, kind=kind)
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
            print(r.stdout[-3000:], flush=True)
            if r.returncode != 0:
                print("rc", r.returncode, r.stderr[-800:], flush=True)
        except subprocess.TimeoutExpired:
            print("CASE", kind, M, N, K, s, "TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
