"""N>1 path on CPU: world_size 2, gloo.  The per-slice solver is the oracle (checker) at a tiny
order; what is under test is the sharding + gather logic of torcwa_b200/sweep.py."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from torcwa_b200.sweep import shard_bounds, solve_sweep
    from oracle.rcwa_oracle import OracleSim
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lams = torch.linspace(400.0, 700.0, 5, dtype=torch.float64)      # 5 points over 2 ranks: uneven shards
    def one(lam):
        sim = OracleSim(freq=1 / float(lam), order=[1, 1], L=[300.0, 300.0], dtype=torch.complex128)
        sim.add_input_layer(eps=2.1316); sim.set_incident_angle(0.1, 0.0)
        sim.add_layer(120.0, 6.0 + 0.1j); sim.solve_global_smatrix()
        return torch.stack([sim.S_parameters([0, 0])[0], sim.S_parameters([0, 0], port="reflection")[0]])
    def solve_slice(lo, hi):
        return torch.stack([one(l) for l in lams[lo:hi]]) if hi > lo else torch.zeros((0, 2), dtype=torch.complex128)
    full = solve_sweep(solve_slice, len(lams))
    ref = torch.stack([one(l) for l in lams])
    assert full.shape == (5, 2)
    assert float((full - ref).abs().max()) == 0.0, (rank, full, ref)
    assert shard_bounds(5, 0, 2) == (0, 3) and shard_bounds(5, 1, 2) == (3, 5)
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_sweep_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_shard_bounds_cover_everything():
    from torcwa_b200.sweep import shard_bounds
    for n in (1, 7, 512, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_chunked_and_points_per_call():
    import torch
    from torcwa_b200 import sweep
    calls = []

    def solve(lo, hi):
        calls.append((lo, hi))
        return torch.arange(lo, hi, dtype=torch.float64)[:, None].to(torch.complex128)
    out = sweep.chunked(solve, 128)(10, 400)
    assert calls == [(10, 138), (138, 266), (266, 394), (394, 400)]
    assert torch.equal(out[:, 0].real, torch.arange(10, 400, dtype=torch.float64))
    assert sweep.points_per_call(15, free_bytes=178 * 2 ** 30) == 128           # a whole B200: capped at 128
    assert sweep.points_per_call(15, free_bytes=40 * 2 ** 30) == 45             # 40 GB free
    assert sweep.points_per_call([21, 21], free_bytes=178 * 2 ** 30) == 55      # BASELINE config 3 (n = 3698)
    assert sweep.points_per_call(25, free_bytes=2 ** 30) == 1
