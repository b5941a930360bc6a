mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "c64_api and sweep_o15_b" > gpurun_out/r2o_c4v.log 2>&1; tail -30 gpurun_out/r2o_c4v.log
python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import cases as C
import torcwa_b200
g = np.load('tests/golden/sweep_o15_b.npz')
for dg in (0, 5, 6, 8):
    sim = C.run_case(lambda freq, order, L, dtype: torcwa_b200.rcwa(freq=freq, order=order, L=L, dtype=dtype, device=torch.device('cuda:0'), gemm_digits=dg), C.CASES['sweep_o15_b'], torch.complex64)
    for k in range(4):
        a = sim.S[k][:, g['S_cols_idx']].cpu().numpy().astype(np.complex128); b = g['S_cols'][k]
        print(dg, k, 'relfro', np.linalg.norm(a - b) / np.linalg.norm(b), 'norm', np.linalg.norm(b), 'per col', [float(np.linalg.norm(a[:, j] - b[:, j]) / np.linalg.norm(b[:, j])) for j in range(2)])
PY
