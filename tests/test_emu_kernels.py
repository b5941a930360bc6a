"""CPU emulation of the phase-structured single-CTA kernels (tests/_emu/librcwa_emu.so, built from
the same .cu sources with -DRCWA_EMU): control flow and index arithmetic of the LU panel, the
windowed multishift QR (deflation, shifts, bulge chains, small-block solves) and the triangular
eigenvector solve, checked against numpy/scipy LAPACK.  The emulation library is test
infrastructure only; the product never loads it."""
import ctypes

import numpy as np
import pytest
import scipy.linalg as sl
import torch

from oracle import cases as C
from oracle.rcwa_oracle import OracleSim


@pytest.fixture(scope="module")
def emu():
    from torcwa_b200 import build
    return ctypes.CDLL(build.build_emu())


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize("n", [1, 5, 32, 33, 97])
def test_lu_right_solve(emu, n):
    rng = np.random.default_rng(n)
    A = crand(rng, n, n)
    LU = A.copy()
    ipiv, perm, info = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(1, np.int32)
    emu.emu_lu_factor(P(LU), n, n, P(ipiv), P(perm), P(info))
    assert info[0] == 0
    B = crand(rng, 6, n)
    X = np.zeros_like(B)
    emu.emu_lu_solve(P(LU), n, n, P(perm), P(B), 6, n, P(X), n)
    assert np.abs(X @ A - B).max() < 1e-11 * max(1, n)


def schur_and_vectors(emu, A):
    n = A.shape[0]
    if n > 2:
        H, Z = sl.hessenberg(A, calc_q=True)
    else:
        H, Z = A.copy(), np.eye(n, dtype=complex)
    H, Z = np.ascontiguousarray(H), np.ascontiguousarray(Z)
    stats = np.zeros(8, np.int32)
    info = emu.emu_qr(P(H), P(Z), n, 10 ** 6, P(stats))
    assert info == 0
    T = np.triu(H)
    assert np.abs(np.tril(H, -1)).max() == 0.0
    X = np.zeros((n, n), complex)
    emu.emu_trevc(P(np.ascontiguousarray(T)), n, P(X))
    V = Z @ X
    V /= np.linalg.norm(V, axis=0)
    return T, Z, V, stats


@pytest.mark.parametrize("n", [2, 3, 16, 17, 48, 49, 64, 65, 130])
def test_qr_and_eigenvectors_random(emu, n):
    rng = np.random.default_rng(1000 + n)
    A = crand(rng, n, n)
    T, Z, V, stats = schur_and_vectors(emu, A)
    assert np.abs(Z @ T @ Z.conj().T - A).max() < 1e-12 * n
    assert np.abs(Z.conj().T @ Z - np.eye(n)).max() < 1e-12 * n
    w = np.diag(T)
    assert np.abs(A @ V - V * w).max() < 1e-11 * n


@pytest.mark.parametrize("n", [49, 97, 130, 200])
@pytest.mark.parametrize("aed_w", [0, 32, 24, 16])
def test_qr_two_launch_mode_and_aed_windows(emu, n, aed_w):
    """The large-batch device path runs every iteration as two launches (bulge-chase windows, then the small dense solves on
    48-row buffers) and picks the AED window by matrix size.  Per matrix the two-launch mode does exactly the same arithmetic
    as the single launch: bit-identical T and Z; every AED window gives a valid Schur decomposition."""
    rng = np.random.default_rng(4000 + n)
    A = crand(rng, n, n)
    H0, Z0 = sl.hessenberg(A, calc_q=True)
    out = []
    for split in (0, 1):
        H, Z = np.ascontiguousarray(H0.copy()), np.ascontiguousarray(Z0.copy())
        stats = np.zeros(8, np.int32)
        info = emu.emu_qr_opts(P(H), P(Z), n, 10 ** 6, P(stats), split, aed_w)
        assert info == 0
        out.append((H, Z, stats.copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert out[0][2][0] == out[1][2][0] and out[0][2][4] == out[1][2][4]          # same sweeps and AEDs
    T, Z = np.triu(out[1][0]), out[1][1]
    assert np.abs(np.tril(out[1][0], -1)).max() == 0.0
    assert np.abs(Z @ T @ Z.conj().T - A).max() < 1e-12 * n
    assert np.abs(Z.conj().T @ Z - np.eye(n)).max() < 1e-12 * n
    w_ref = np.linalg.eigvals(A)
    assert max(np.abs(w_ref - x).min() for x in np.diag(T)) < 1e-10 * np.abs(w_ref).max()


def test_qr_on_rcwa_matrix_and_degenerate_cell(emu):
    for name in ("ex1_o3", "square_o4"):
        case = C.CASES[name]
        sim = OracleSim(freq=C.freq_of(case, torch.complex128), order=case["order"], L=case["L"], dtype=torch.complex128)
        sim.add_input_layer(eps=case["eps_in"])
        sim.set_incident_angle(0.0, 0.0)
        d, e = C.build_layers(case, torch.complex128)[0]
        sim.add_layer(d, e)
        A = (sim.P[0] @ sim.Q[0]).numpy()
        T, Z, V, stats = schur_and_vectors(emu, A)
        w = np.diag(T)
        assert np.abs(A @ V - V * w).max() < 1e-10 * np.abs(A).max()
        ref = np.sort_complex((sim.kz_norm[0].numpy()) ** 2)
        assert np.abs(np.sort_complex(w) - ref).max() < 1e-9 * np.abs(ref).max()
        assert np.linalg.cond(V) < 1e8
        assert stats[0] < 3 * A.shape[0] / 16 + 10      # sweeps stay ~ n/16 * small constant


def test_tiny_shift_solver(emu):
    rng = np.random.default_rng(5)
    for m in (1, 2, 3, 7, 16):
        T = np.triu(crand(rng, m, m), -1).copy()
        w = np.zeros(m, complex)
        assert emu.emu_tiny_eigs(P(T.copy()), m, P(w)) == 0
        ref = np.linalg.eigvals(T)
        assert max(np.min(np.abs(ref - x)) for x in w) < 1e-12


def test_qr_robustness_battery(emu):
    """The QR state machine (deflation scan, AED with time-sliced Schur / scan / restore, bulge chains, small-block
    solves) on matrices that stress it: graded, defective, non-normal (Grcar, companion), unitary, Hermitian, skew,
    rank one, tightly clustered, and scaled to the ends of the fp64 range.  Every case must converge with a backward
    error at round-off level."""
    rng = np.random.default_rng(7)
    n = 96
    J = np.diag(np.ones(n - 1), 1) + 2 * np.eye(n)
    comp = np.zeros((n, n)); comp[0, :] = -rng.standard_normal(n); comp[np.arange(1, n), np.arange(n - 1)] = 1
    Hm = rng.standard_normal((n, n))
    cases = {
        "graded": np.diag(10.0 ** np.linspace(-6, 6, n)) @ rng.standard_normal((n, n)) @ np.diag(10.0 ** np.linspace(6, -6, n)),
        "jordan+eps": J + 1e-10 * rng.standard_normal((n, n)),
        "grcar": np.triu(np.ones((n, n))) - np.triu(np.ones((n, n)), 4) - np.diag(np.ones(n - 1), -1),
        "companion": comp,
        "unitary": np.linalg.qr(crand(rng, n, n))[0],
        "hermitian": Hm + Hm.T,
        "skew": Hm - Hm.T,
        "rank_one": np.outer(rng.standard_normal(n), rng.standard_normal(n)),
        "clustered": np.diag(1 + 1e-9 * rng.standard_normal(n)) + 1e-9 * rng.standard_normal((n, n)),
        "huge": 1e150 * rng.standard_normal((n, n)),
        "tiny": 1e-150 * rng.standard_normal((n, n)),
        "zero": np.zeros((n, n)),
    }
    for name, A in cases.items():
        A = np.ascontiguousarray(A.astype(complex))
        H, Z = sl.hessenberg(A, calc_q=True)
        H, Z = np.ascontiguousarray(H), np.ascontiguousarray(Z)
        stats = np.zeros(8, np.int32)
        info = emu.emu_qr(P(H), P(Z), n, 10 ** 6, P(stats))
        assert info == 0, name
        T = np.triu(H)
        scale = max(np.abs(A).max(), 1e-300)
        assert np.abs(Z @ T @ Z.conj().T - A).max() <= 1e-12 * scale, name
        assert np.abs(Z.conj().T @ Z - np.eye(n)).max() <= 1e-12, name
