"""The tcgen05 (int8 digit) complex GEMM: host-side schedule and arithmetic model on the CPU, kernels on the GPU.

CPU part: the load / MMA / release table the kernel walks (rcwa_tc_schedule, host code of the product library) covers
every digit pair exactly once and cannot deadlock its 12-slot ring; the digit arithmetic (tests/tc_model.py, a numpy
mirror of the kernels) reaches the accuracy the header states.
GPU part (-m gpu): the split kernels agree with the model digit for digit; the GEMM agrees with torch fp64.
"""
import numpy as np
import pytest
import torch

import tc_model


def _rnd(shape, seed, spread=0.0):
    g = np.random.default_rng(seed)
    x = g.standard_normal(shape) + 1j * g.standard_normal(shape)
    if spread:
        x = x * np.exp(spread * g.standard_normal(shape))
    return x


@pytest.mark.parametrize("s", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("nl", [1, 2, 3, 4])
def test_schedule_covers_all_pairs_and_ring_is_deadlock_free(s, nl):
    from torcwa_b200 import _lib
    ops, groups = _lib.tc_schedule(s, nl)
    assert sum(g["nl"] for g in groups) == s
    pairs = tc_model.simulate_schedule(ops, groups, s, ring=12)
    assert sorted(tc_model.simulate_compact(groups, ring=12)) == sorted(pairs)        # the tables the kernel actually walks
    want = sorted((p, q) for p in range(s) for q in range(s) if p + q <= s - 1)
    got = sorted((p, q) for (_, p, q, _, _) in pairs)
    assert got == want
    for d0, p, q, lvl, first in pairs:
        assert lvl == p + q - d0 and 0 <= lvl < nl
    # exactly one "first" MMA per level and group
    for g in groups:
        firsts = [lvl for (d0, _, _, lvl, first) in pairs if d0 == g["d0"] and first]
        assert sorted(firsts) == list(range(g["nl"]))


@pytest.mark.parametrize("s", [2, 4, 5, 7, 8])
def test_issue_table_entries_cover_every_pair_once(s):
    """The issue table the kernel's MMA issuer reads (tc_issue_entry, shared host/device code): for every level group and
    every ring position, the MMA groups of a step -- single planes (N = 128) and doubles (two digit planes in adjacent ring
    slots, one N = 256 MMA over two adjacent accumulator levels) -- cover exactly the (A digit, B digit) pairs of the
    schedule, each into the accumulator column of its level."""
    from torcwa_b200 import _lib
    RING = 12
    ops, groups = _lib.tc_schedule(s, 4)
    n_double = 0
    for gi, g in enumerate(groups):
        nl = g["nl"]
        for bs in range(RING):
            entries = _lib.tc_issue_entries(s, gi, bs)
            # expected pairs per step from the compact MMA table: consecutive ops sharing their A load
            steps, cur = [], None
            for w in g["mmas"]:
                ia, ib, lvl, first = w & 31, (w >> 5) & 31, (w >> 10) & 3, (w >> 12) & 1
                if cur is None or cur[0] != ia:
                    cur = (ia, [])
                    steps.append(cur)
                cur[1].append((ib, lvl, first))
            assert len(entries) == len(steps)
            for e, (ia, pairs) in zip(entries, steps):
                assert e["a_slot"] == (bs + ia) % RING
                got = []
                for grp in e["groups"]:
                    lvl = nl - 1 - grp["col"] // 128
                    got.append((grp["b_slot"], lvl, grp["first"]))
                    if grp["double"]:
                        n_double += 1
                        assert grp["b_slot"] + 1 < RING and lvl >= 1
                        got.append((grp["b_slot"] + 1, lvl - 1, grp["first"]))       # second plane: next slot, next column
                want = [((bs + ib) % RING, lvl, first) for (ib, lvl, first) in pairs]
                assert got == want
    assert n_double > 0 or s < 3


@pytest.mark.parametrize("s,tol", [(4, 2e-6), (5, 1e-8), (6, 3e-11), (7, 2e-13), (8, 2e-15)])
def test_model_accuracy(s, tol):
    A = _rnd((37, 150), 1, spread=1.0)
    B = _rnd((150, 29), 2, spread=1.0)
    ref = A @ B
    got = tc_model.gemm(A, B, s)
    # norm-wise per row / column scale: compare against |A| |B|
    bound = np.abs(A) @ np.abs(B)
    assert np.max(np.abs(got - ref) / bound) < tol


def test_model_ops_and_reconstruction():
    A = _rnd((20, 33), 3)
    B = _rnd((21, 33), 4)
    assert np.max(np.abs(tc_model.gemm(A, B, 7, "N", "H") - A @ B.conj().T)) < 1e-12
    assert np.max(np.abs(tc_model.gemm(A.T.copy(), B, 7, "T", "T") - A @ B.T)) < 1e-12
    dig, ex = tc_model.split_vectors(A, 7)
    rec = tc_model.reconstruct(dig, ex, 7)
    scale = 2.0 ** ex[:, None]
    assert np.max(np.abs(rec[0] - A.real) / scale) < 2.0 ** -53
    assert np.max(np.abs(rec[2] - (rec[0] + rec[1])) / scale) < 2.0 ** -60       # re+im digits are the exact integer sum


# ------------------------------------------------------------------------------------------------ GPU
def _dev():
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("rows_contiguous", [True, False])
@pytest.mark.parametrize("s", [4, 7, 8])
@pytest.mark.parametrize("shape", [(5, 7), (130, 200), (64, 128), (33, 257)])
def test_split_kernels_match_model(rows_contiguous, s, shape):
    from torcwa_b200 import _lib
    X = _rnd((2,) + shape, 5, spread=2.0)
    X[1, 1, :] = 0.0
    X[1, :, 2] = 0.0
    Xt = torch.from_numpy(X).to(_dev())
    planes, ex = _lib.tc_split(Xt, rows_contiguous, s, conj=True)
    planes, ex = planes.cpu().numpy(), ex.cpu().numpy()
    for b in range(2):
        V = X[b] if rows_contiguous else np.ascontiguousarray(X[b].T)
        dig, e = tc_model.split_vectors(V, s, conj=True)
        K = V.shape[1]
        assert np.array_equal(ex[b], e)
        assert np.array_equal(planes[b][:, :, :, :K], dig)
        assert not planes[b][:, :, :, K:].any()


@pytest.mark.gpu
@pytest.mark.parametrize("opa", ["N", "T", "H"])
@pytest.mark.parametrize("opb", ["N", "T", "H"])
@pytest.mark.parametrize("mnk", [(128, 128, 128), (97, 130, 37), (1, 1, 1), (200, 33, 64), (300, 260, 300), (128, 16, 129)])
def test_zgemm_tc_vs_torch(opa, opb, mnk):
    from torcwa_b200 import _lib
    M, N, K = mnk
    nb = 3
    A = torch.from_numpy(_rnd((nb,) + ((M, K) if opa == "N" else (K, M)), 1)).to(_dev())
    B = torch.from_numpy(_rnd((nb,) + ((K, N) if opb == "N" else (N, K)), 2)).to(_dev())
    Cin = torch.from_numpy(_rnd((nb, M, N), 3)).to(_dev())
    f = {"N": lambda x: x, "T": lambda x: x.transpose(1, 2), "H": lambda x: x.transpose(1, 2).conj()}
    prod = f[opa](A) @ f[opb](B)
    bound = f[opa](A).abs() @ f[opb](B).abs()
    out0 = _lib.zgemm_tc(A, B, opa, opb, slices=8)
    assert float(((out0 - prod).abs() / bound).max()) < 4e-15
    out = Cin.clone()
    _lib.zgemm_tc(A, B, opa, opb, alpha=-0.5, beta=2.0 + 1.0j, out=out, slices=8)
    ref = -0.5 * prod + (2.0 + 1.0j) * Cin
    assert float(((out - ref).abs() / (bound + Cin.abs())).max()) < 4e-15


@pytest.mark.gpu
@pytest.mark.parametrize("s,tol", [(3, 5e-4), (4, 2e-6), (5, 1e-8), (6, 3e-11), (7, 2e-13), (8, 4e-15)])
def test_zgemm_tc_slices(s, tol):
    from torcwa_b200 import _lib
    A = torch.from_numpy(_rnd((2, 260, 515), 7, spread=1.0)).to(_dev())
    B = torch.from_numpy(_rnd((2, 515, 140), 8, spread=1.0)).to(_dev())
    out = _lib.zgemm_tc(A, B, slices=s)
    bound = A.abs() @ B.abs()
    assert float(((out - A @ B).abs() / bound).max()) < tol


@pytest.mark.gpu
def test_zgemm_tc_chunked_workspace_and_model_bit_agreement():
    """A workspace for one matrix only (the routine then runs the batch in chunks), and the result equals the numpy
    model of the same digit arithmetic exactly (the int32 level sums are exact; only the last fp64 roundings may differ)."""
    from torcwa_b200 import _lib
    lib = _lib.load()
    A = _rnd((3, 140, 200), 9)
    B = _rnd((3, 200, 150), 10)
    At, Bt = torch.from_numpy(A).to(_dev()), torch.from_numpy(B).to(_dev())
    one = lib.rcwa_zgemm_tc_workspace_bytes(140, 150, 200, 1, 5)
    out = _lib.zgemm_tc(At, Bt, slices=5, ws_bytes=one).cpu().numpy()
    full = _lib.zgemm_tc(At, Bt, slices=5).cpu().numpy()
    assert np.array_equal(out, full)
    for b in range(3):
        model = tc_model.gemm(A[b], B[b], 5)
        assert np.max(np.abs(out[b] - model)) <= 1e-13 * np.max(np.abs(model))


@pytest.mark.gpu
def test_zgemm_tc_path_size():
    """One product at the path's size (order 15: n = 1922) against the DMMA kernel."""
    from torcwa_b200 import _lib
    n = 1922
    A = torch.from_numpy(_rnd((2, n, n), 11)).to(_dev())
    B = torch.from_numpy(_rnd((2, n, n), 12)).to(_dev())
    ref = _lib.zgemm(A, B)
    out = _lib.zgemm_tc(A, B, slices=8)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-14
    out = _lib.zgemm_tc(A, B, slices=5)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-8
