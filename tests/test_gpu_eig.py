"""CUDA eigensolver (C ABI rcwa_hessenberg / rcwa_eig) against fp64 LAPACK through torch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.complex(torch.randn(*shape, generator=g, dtype=torch.float64),
                         torch.randn(*shape, generator=g, dtype=torch.float64)).to(dev())


def rel(a, b):
    return float(torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1)))


def match_sorted(a, b):
    """max distance between two multisets of complex numbers after greedy nearest matching."""
    a, b = list(a), list(b)
    worst = 0.0
    for x in a:
        d = [abs(x - y) for y in b]
        k = int(np.argmin(d))
        worst = max(worst, d[k])
        b.pop(k)
    return worst


@pytest.mark.parametrize("n", [3, 4, 17, 63, 64, 65, 130, 200])
def test_hessenberg(n):
    from torcwa_b200 import _lib
    A = rnd(3, n, n, seed=n)
    H = A.clone()
    Z = _lib.hessenberg_(H)
    torch.cuda.synchronize()
    eye = torch.eye(n, dtype=torch.complex128, device=dev())
    assert float((Z.conj().transpose(1, 2) @ Z - eye).abs().max()) < 1e-12
    assert float(torch.tril(H, -2).abs().max()) == 0.0
    back = Z @ H @ Z.conj().transpose(1, 2)
    assert float((back - A).abs().max() / A.abs().max()) < 1e-12


@pytest.mark.parametrize("n", [2, 3, 5, 16, 17, 48, 49, 64, 65, 100, 150, 242])
def test_eig_random(n):
    from torcwa_b200 import _lib
    nb = 3
    A = rnd(nb, n, n, seed=100 + n)
    w, V, info = _lib.eig(A.clone())
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0, info
    res = (A @ V - V * w[:, None, :]).abs().max() / A.abs().max()
    assert float(res) < 1e-11
    assert float(((V.abs() ** 2).sum(dim=1) - 1).abs().max()) < 1e-12
    ref = torch.linalg.eigvals(A.cpu())
    for b in range(nb):
        assert match_sorted(w[b].cpu().numpy(), ref[b].numpy()) < 1e-10 * float(ref[b].abs().max())


@pytest.mark.parametrize("n,d", [(97, 1), (200, 2), (482, 1)])
def test_eig_keeps_an_exactly_decoupled_leading_block_in_place(n, d):
    """A = diag(1 ... 1, A1) with exact zeros in the couplings (how the host extends a symmetry block to the common batch
    size, rcwa._patterned_layer_blocks): the first d eigenvalues are exactly 1, the eigenvector matrix is exactly
    diag(I, W1), and (lam1, W1) solve A1 as well as a direct call does.  Mixed with unextended matrices in one batch."""
    from torcwa_b200 import _lib
    A1 = rnd(2, n - d, n - d, seed=n)
    A = torch.zeros((3, n, n), dtype=torch.complex128, device=dev())
    A[:2, d:, d:] = A1
    A[:2, range(d), range(d)] = 1.0
    A[2] = rnd(1, n, n, seed=n + 1)[0]
    A0 = A.clone()
    w, V, info = _lib.eig(A)
    assert int(info.abs().max()) == 0
    assert float((w[:2, :d] - 1).abs().max()) == 0.0
    assert float(V[:2, :d, d:].abs().max()) == 0.0 and float(V[:2, d:, :d].abs().max()) == 0.0
    assert float((V[:2, range(d), range(d)] - 1).abs().max()) == 0.0
    R = A0 @ V - V * w[:, None, :]
    assert float(R.abs().max() / A0.abs().max()) <= 1e-11 * n
    w1, _, _ = _lib.eig(A1.clone())
    for b in range(2):
        assert match_sorted(w[b, d:].cpu().numpy(), w1[b].cpu().numpy()) <= 1e-9 * float(w1[b].abs().max())


def test_eig_batch_entries_independent():
    """A batch of different sizes of difficulty must give the same answer as one at a time."""
    from torcwa_b200 import _lib
    A = rnd(4, 90, 90, seed=7)
    A[1] = torch.diag(torch.arange(1, 91, dtype=torch.float64, device=dev()).to(torch.complex128))   # already triangular
    A[2] = A[2] * 1e-3
    w, V, info = _lib.eig(A.clone())
    for b in range(4):
        w1, V1, i1 = _lib.eig(A[b:b + 1].clone())
        assert int(i1[0]) == 0 and int(info[b]) == 0
        assert match_sorted(w[b].cpu().numpy(), w1[0].cpu().numpy()) < 1e-10 * float(w1.abs().max())


@pytest.mark.parametrize("name", ["ex1_o3", "ex1_o5", "square_o4"])
def test_eig_rcwa_matrix(name, golden_dir):
    """The real thing: P*Q of a patterned layer (incl. the C4v cell with degenerate pairs);
    eigenvalues against the reference's kz^2 (golden) and residual of the eigenpairs."""
    from oracle import cases as C
    from oracle.rcwa_oracle import OracleSim
    from torcwa_b200 import _lib
    case = C.CASES[name]
    sim = OracleSim(freq=C.freq_of(case, torch.complex128), order=case["order"], L=case["L"], dtype=torch.complex128)
    sim.add_input_layer(eps=case["eps_in"])
    sim.set_incident_angle(0.0, 0.0)
    d, e = C.build_layers(case, torch.complex128)[0]
    sim.add_layer(d, e)
    A = (sim.P[0] @ sim.Q[0]).to(dev())[None].contiguous()
    w, V, info = _lib.eig(A.clone())
    assert int(info[0]) == 0
    res = (A @ V - V * w[:, None, :]).abs().max() / A.abs().max()
    assert float(res) < 1e-11
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    mine = np.sort_complex(w[0].cpu().numpy())
    assert np.abs(mine - g["kz2_sorted"][0]).max() < 1e-9 * np.abs(mine).max()
    assert float(torch.linalg.cond(V[0].cpu())) < 1e8


# ------------------------------------------------------------------ Eig.backward
@pytest.mark.parametrize("name", ["rand6", "rand24", "rand57", "rcwa_o3"])
@pytest.mark.parametrize("broadening", [1e-10, None])
def test_eig_backward_kernel_vs_reference_golden(name, broadening, golden_dir):
    """rcwa_eig_backward on the reference's own (eigval, eigvec, incoming gradients) == the reference's Eig.backward."""
    import os
    from torcwa_b200 import _lib
    g = np.load(os.path.join(golden_dir, "eig_backward.npz"))
    t = lambda k: torch.from_numpy(g[name + "_" + k]).to(dev())[None].contiguous()
    ref = g[name + "_grad_b" + ("1e-10" if broadening is not None else "None")]
    delta = 1e-10 if broadening is not None else 4.9e-324
    grad, info = _lib.eig_backward(t("w"), t("V"), t("gw"), t("gV"), delta)
    assert int(info.abs().max()) == 0
    got = grad[0].cpu().numpy()
    assert np.linalg.norm(got - ref) <= 1e-10 * np.linalg.norm(ref)
    # either incoming gradient may be absent
    from oracle.rcwa_oracle import eig_backward
    w, V, gw, gV = (torch.from_numpy(g[name + "_" + k]) for k in ("w", "V", "gw", "gV"))
    only_w, _ = _lib.eig_backward(t("w"), t("V"), t("gw"), None, 1e-10)
    only_V, _ = _lib.eig_backward(t("w"), t("V"), None, t("gV"), 1e-10)
    rw = eig_backward(w, V, gw, torch.zeros_like(gV), 1e-10).numpy()
    rV = eig_backward(w, V, torch.zeros_like(gw), gV, 1e-10).numpy()
    assert np.linalg.norm(only_w[0].cpu().numpy() - rw) <= 1e-10 * np.linalg.norm(rw)
    assert np.linalg.norm(only_V[0].cpu().numpy() - rV) <= 1e-10 * np.linalg.norm(rV)


def test_eig_autograd_gauge_invariant_loss_matches_native():
    """End to end through torcwa_b200.Eig (CUDA forward + CUDA backward) against PyTorch's native eig autograd on the
    CPU, on a gauge-invariant loss |sum M o (V exp(0.1 i L) V^-1)|^2; batched and unbatched, c128 and c64."""
    import torcwa_b200
    n = 40
    A0 = rnd(2, n, n, seed=77).cpu()
    Mw = rnd(n, n, seed=78).cpu()

    def loss(w, V, Mw):
        f = V @ torch.diag_embed(torch.exp(0.1j * w)) @ torch.linalg.inv(V)
        return ((Mw * f).sum(dim=(-2, -1)).abs() ** 2).sum()

    old = torcwa_b200.Eig.broadening_parameter
    try:
        torcwa_b200.Eig.broadening_parameter = None
        Ac = A0.clone().requires_grad_(True)
        loss(*torch.linalg.eig(Ac), Mw).backward()
        Ag = A0.clone().to(dev()).requires_grad_(True)
        loss(*torcwa_b200.Eig.apply(Ag), Mw.to(dev())).backward()
        assert rel(Ag.grad.cpu(), Ac.grad) < 1e-9
        A1 = A0[0].clone().to(dev()).requires_grad_(True)            # unbatched
        loss(*torcwa_b200.Eig.apply(A1), Mw.to(dev())).backward()
        A1c = A0[0].clone().requires_grad_(True)
        loss(*torch.linalg.eig(A1c), Mw).backward()
        assert rel(A1.grad.cpu(), A1c.grad) < 1e-9
        A64 = A0.to(torch.complex64).to(dev()).requires_grad_(True)   # complex64 API, complex128 arithmetic
        loss(*torcwa_b200.Eig.apply(A64), Mw.to(torch.complex64).to(dev())).backward()
        assert A64.grad.dtype == torch.complex64
        assert rel(A64.grad.cpu().to(torch.complex128), Ac.grad) < 1e-4
        Ar = A0[0].real.clone().to(dev()).requires_grad_(True)        # real input -> real gradient (torch_eig.py:41-42)
        loss(*torcwa_b200.Eig.apply(Ar), Mw.to(dev())).backward()
        assert Ar.grad.dtype == torch.float64 and not torch.is_complex(Ar.grad)
    finally:
        torcwa_b200.Eig.broadening_parameter = old


@pytest.mark.parametrize("keys", [{13: 2}, {13: 2, 14: 2}, {14: 2}, {13: 2, 9: 4}])
def test_eig_launch_modes_give_identical_results(keys):
    """Tuning keys 13 (QR pass as two launches per iteration: windows / small dense solves at two CTAs per SM), 14 (CUDA
    graph replay of the QR loop) and 9 (matrix groups) only change how the same per-matrix work is scheduled: eigenvalues
    and eigenvectors are bit-identical to the default mode, on matrices large enough for sweeps, AED and small blocks."""
    from torcwa_b200 import _lib
    lib = _lib.load()
    A = rnd(24, 230, 230, seed=77)
    A[3] = torch.diag(torch.arange(1.0, 231.0, dtype=torch.float64, device=dev()).to(torch.complex128))      # converged at once
    try:
        for k in (9, 13, 14):
            lib.rcwa_set_tuning(k, 1 if k != 9 else 0)
        w0, V0, i0 = _lib.eig(A.clone())
        for k, v in keys.items():
            lib.rcwa_set_tuning(k, v)
        w1, V1, i1 = _lib.eig(A.clone())
    finally:
        for k in (9, 13, 14):
            lib.rcwa_set_tuning(k, 0)
    assert int(i0.abs().max()) == 0 and int(i1.abs().max()) == 0
    assert torch.equal(w0, w1) and torch.equal(V0, V1)
    R = A @ V1 - V1 * w1[:, None, :]
    assert float(R.abs().max() / A.abs().max()) <= 1e-11 * 230


def test_eig_is_reentrant_across_host_threads_and_bitwise_reproducible():
    """Boundary contract (include/rcwa_b200.h): rcwa_eig may run concurrently from different host threads on their own
    streams and workspaces (its internal streams are per thread), twice in a row from the same thread (the cached streams
    are reused), and gives bit-identical results however the calls interleave.  rcwa_eig_phases(1) + (2) == rcwa_eig."""
    import threading
    from torcwa_b200 import _lib
    n, nb = 150, 4
    A0, A1 = rnd(nb, n, n, seed=7), rnd(nb, n, n, seed=8)
    ref0 = _lib.eig(A0.clone())
    ref1 = _lib.eig(A1.clone())
    again = _lib.eig(A0.clone())                       # same thread, cached internal streams reused
    torch.cuda.synchronize()
    assert torch.equal(again[0], ref0[0]) and torch.equal(again[1], ref0[1])
    out, errs = {}, []

    def work(k, A):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(3):
                    out[k] = _lib.eig(A.clone())
            st.synchronize()
        except BaseException as e:
            errs.append(e)
    ts = [threading.Thread(target=work, args=(0, A0)), threading.Thread(target=work, args=(1, A1))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for k, ref in ((0, ref0), (1, ref1)):
        assert int(out[k][2].abs().max()) == 0
        assert torch.equal(out[k][0], ref[0]) and torch.equal(out[k][1], ref[1])
    hooks = []
    two = _lib.eig(A0.clone(), after_reduction=lambda: hooks.append(1))
    torch.cuda.synchronize()
    assert hooks == [1] and torch.equal(two[0], ref0[0]) and torch.equal(two[1], ref0[1])


@pytest.mark.parametrize("n", [1054, 1922])
def test_eig_backward_at_path_size(n):
    """Eig.backward at the sizes of the path (n = 1054: Example6's order [15,8]; n = 1922: order 15x15): rcwa_eig_backward on
    our own eigen-decomposition against the reference's formula (torch_eig.py:25-40) evaluated with torch fp64 ops on the GPU:
    grad = X^-H (diag(g_lambda) + conj(F) o (X^H g_X)) X^H,  F_ij = conj(s) / (|s|^2 + delta), s = lambda_j - lambda_i."""
    from torcwa_b200 import _lib
    A = rnd(1, n, n, seed=5) / np.sqrt(n) + torch.diag(torch.linspace(-3.0, 1.0, n, dtype=torch.float64)).to(dev())[None]
    w, V, info = _lib.eig(A.clone())
    assert int(info.abs().max()) == 0
    gw, gV = rnd(1, n, seed=6), rnd(1, n, n, seed=7) / np.sqrt(n)
    delta = 1e-10
    grad, info = _lib.eig_backward(w, V, gw, gV, delta)
    assert int(info.abs().max()) == 0
    s = w[0][None, :] - w[0][:, None]
    F = s.conj() / (s.abs() ** 2 + delta)
    F.fill_diagonal_(0.0)
    Xh = V[0].conj().T
    inner = torch.diag(gw[0]) + F.conj() * (Xh @ gV[0])
    ref = torch.linalg.solve(Xh, inner @ Xh)
    err = rel(grad[0], ref)
    print("n = %d: eig backward vs torch fp64 formula %.2e" % (n, err))
    assert err < 1e-9
