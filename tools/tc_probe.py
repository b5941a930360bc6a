"""Probe of the tcgen05 (int8 digit) complex GEMM on a B200: correctness against the DMMA kernel and timing at the
path's shapes.  Each case runs in its own process with a timeout, so that a faulting kernel cannot take the run down.

    python tools/tc_probe.py [--out gpurun_out/tc_probe.json]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r"""
import json, sys, time
sys.path.insert(0, %(root)r)
import torch
from torcwa_b200 import _lib
M, N, K, nb, s, reps = %(M)d, %(N)d, %(K)d, %(nb)d, %(s)d, %(reps)d
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(1)
def rnd(*sh):
    return torch.complex(torch.randn(*sh, generator=g, dtype=torch.float64), torch.randn(*sh, generator=g, dtype=torch.float64)).to(dev)
A, B = rnd(nb, M, K), rnd(nb, K, N)
out = {'M': M, 'N': N, 'K': K, 'nb': nb, 's': s}
ref = _lib.zgemm(A, B)
torch.cuda.synchronize()
if s > 0:
    C = _lib.zgemm_tc(A, B, slices=s)
    torch.cuda.synchronize()
    out['err_vs_dmma'] = float((C - ref).abs().max() / ref.abs().max())
    fn = lambda: _lib.zgemm_tc(A, B, slices=s, out=C)
else:
    C = ref
    fn = lambda: _lib.zgemm(A, B, out=C)
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
out['ms'] = ms
out['tflops_8mnk'] = 8.0 * M * N * K * nb / ms / 1e9
if s > 0:
    pairs = s * (s + 1) // 2
    out['int8_tops'] = 2.0 * 3 * pairs * M * N * K * nb / ms / 1e9
print(json.dumps(out))
"""


def run_case(**kw):
    code = CASE % dict(root=ROOT, **kw)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=kw.get("timeout", 180))
    except subprocess.TimeoutExpired:
        return dict(kw, error="timeout")
    if r.returncode != 0:
        return dict(kw, error=(r.stderr or r.stdout)[-600:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tc_probe.json"))
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    res = []
    shapes = [(256, 256, 256, 2), (1922, 1922, 1922, 8)]
    if not a.quick:
        shapes += [(1922, 1922, 961, 8), (1922, 128, 1922, 16), (3698, 3698, 3698, 2)]
    for (M, N, K, nb) in shapes:
        for s in ([0, 4, 5, 7, 8] if M > 256 else [7]):
            r = run_case(M=M, N=N, K=K, nb=nb, s=s, reps=3 if M > 256 else 1)
            print(json.dumps(r), flush=True)
            res.append(r)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
