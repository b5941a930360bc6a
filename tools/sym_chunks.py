"""Step time and QR statistics of every 128-wavelength chunk of the sweep on the symmetry-reduced path (which chunk is slow, and why).
python tools/sym_chunks.py [--points 128]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch
import bench
from torcwa_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=128)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--max-chunks", type=int, default=4)
ap.add_argument("--groups", type=int, default=0, help="QR groups per rcwa_eig call (tuning key 9; 0 = automatic)")
a = ap.parse_args()
dev = torch.device("cuda:0")
if a.groups:
    _lib.load().rcwa_set_tuning(9, a.groups)
case, mask, lams = bench.sweep_inputs(15)
grids = bench.make_grids(mask, lams).to(dev)
freq = (1.0 / lams).to(dev)
log = []
real_eig = _lib.eig
def eig_logged(A, after_reduction=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = real_eig(A, after_reduction)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    st = _lib.last_eig_stats.cpu()
    log.append((tuple(A.shape), dt, st[:, 0].float().mean().item(), int(st[:, 0].max()), st[:, 1].float().mean().item(), int(st[:, 1].max()), int(st[:, 1].argmax())))
    return out
import torcwa_b200
host = sys.modules['torcwa_b200.rcwa']
def step(s):
    sl = (torch.arange(a.points, device=dev) + s * a.points) % 512
    return bench.run_step(grids[sl], freq[sl], case, dev, None)
step(0); torch.cuda.synchronize()
nchunks = min(512 // a.points, a.max_chunks)
for rep in range(a.reps):
    for s in range(nchunks):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(s); e1.record(); torch.cuda.synchronize()
        print("rep %d chunk %d: %.1f ms (%.1f layers/s)  reserved %.1f GB" % (rep, s, e0.elapsed_time(e1), a.points / e0.elapsed_time(e1) * 1e3, torch.cuda.memory_reserved() / 1e9), flush=True)
host._lib.eig = eig_logged
for s in range(nchunks):
    log.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(s); e1.record(); torch.cuda.synchronize()
    print("chunk %d with eig calls synchronised: %.1f ms" % (s, e0.elapsed_time(e1)))
    for shp, dt, sw_mean, sw_max, p_mean, p_max, p_arg in log:
        print("    eig %s: %.1f ms; sweeps mean %.1f max %d; passes mean %.0f max %d (matrix %d)" % (shp, dt, sw_mean, sw_max, p_mean, p_max, p_arg))
