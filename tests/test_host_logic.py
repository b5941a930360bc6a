"""Host logic of torcwa_b200.rcwa on the CPU, with the C ABI replaced by the torch test double
(tests/fake_lib.py), against the reference's stored outputs and the oracle."""
import os

import numpy as np
import pytest
import torch

from fake_lib import cpu_double  # noqa: F401  (fixture)
from oracle import cases as C
from oracle.rcwa_oracle import OracleSim

SMALL = ["ex1_o3", "ex1_o5", "stack_o3", "stack_o4x2", "fresnel_o2", "square_o4", "c2_o3", "ymirror_o3", "xmirror_o3", "offcentre_o3"]
# which symmetry the reduction (torcwa_b200/symmetry.py) must find in each case (None: general path)
SYMMETRY = {"ex1_o3": ("x", "y"), "ex1_o5": ("x", "y"), "square_o4": ("x", "y"), "offcentre_o3": ("x", "y"), "c2_o3": ("c2",),
            "ymirror_o3": ("y",), "xmirror_o3": ("x",), "stack_o3": None, "stack_o4x2": None, "fresnel_o2": None}
CPU = torch.device("cpu")


def relfro(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name", SMALL)
def test_host_path_matches_reference_c128(cpu_double, name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sim = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), C.CASES[name], torch.complex128)
    sp = C.probe(sim)
    assert np.abs(sp - g["sparams_c128"]).max() <= 1e-10 * np.abs(g["sparams_c128"]).max()
    for k in range(4):
        assert relfro(sim.S[k][:, g["S_cols_idx"]].numpy(), g["S_cols"][k]) <= 1e-10
    if "S" in g:
        for k in range(4):
            assert relfro(sim.S[k].numpy(), g["S"][k]) <= 1e-10
    found = sim._sym.gens if sim._sym not in (None, False) else None
    assert found == SYMMETRY[name]


@pytest.mark.parametrize("name", [k for k, v in SYMMETRY.items() if v])
def test_symmetry_reduced_path_equals_general_path(cpu_double, name, golden_dir):
    """The block path (2 or 4 blocks in the symmetry-adapted basis) and the general path give the same global S-matrix,
    layer S-matrices, kz multiset and fields-relevant attributes; both match the reference's stored output."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    mk = lambda sym: C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=sym, **kw), C.CASES[name], torch.complex128)
    a, b = mk(True), mk(False)
    assert a._sym not in (None, False) and b._sym is None
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-11
        if "S" in g:
            assert relfro(a.S[k].numpy(), g["S"][k]) <= 1e-10
        assert relfro(a.S[k][:, g["S_cols_idx"]].numpy(), g["S_cols"][k]) <= 1e-10
    for li in range(a.layer_N):
        assert relfro(a.layer_S11[li].numpy(), b.layer_S11[li].numpy()) <= 1e-11
        assert relfro(a.layer_S21[li].numpy(), b.layer_S21[li].numpy()) <= 1e-11
        ka, kb = np.sort_complex(a.kz_norm[li].numpy() ** 2), np.sort_complex(b.kz_norm[li].numpy() ** 2)
        assert np.abs(ka - kb).max() <= 1e-10 * np.abs(kb).max()
    # the eigenvectors returned to the original basis still diagonalise P Q
    li = 0
    W, kz = a.E_eigvec[li], a.kz_norm[li]
    A = a.P[li] @ a.Q[li]
    assert relfro((A @ W).numpy(), (W * (kz ** 2)[None, :]).numpy()) <= 1e-10


def test_symmetry_is_dropped_when_a_later_layer_breaks_it(cpu_double):
    """First layer symmetric (solved in blocks), second layer an off-axis rotated bar without the mirrors: the stack is
    cascaded in the original basis, and equals the general path."""
    case = dict(C.CASES["ex1_o3"])
    case["layers"] = [C._rect(), dict(C._rect(theta=0.4, d=120.0), Cx=140.0)]

    def run(sym):
        return C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=sym, **kw), case, torch.complex128)
    a, b = run(True), run(False)
    assert a._sym is False
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-11


@pytest.mark.parametrize("thetas,expect", [((0.0, 0.5), ("c2",)), ((0.5, 0.0), ("c2",)), ((0.0, 0.5, 0.0), ("c2",))])
def test_stack_moves_to_the_common_subgroup(cpu_double, thetas, expect):
    """Mirror-symmetric bars and rotated bars about the same centre share C2 only: the stack is solved in the C2 blocks,
    the earlier layers re-expressed in that basis, and equals the general path."""
    case = dict(C.CASES["ex1_o3"])
    case["layers"] = [C._rect(theta=t, d=100.0 + 40.0 * i) for i, t in enumerate(thetas)]

    def run(sym):
        return C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=sym, **kw), case, torch.complex128)
    a, b = run(True), run(False)
    assert a._sym.gens == expect and set(a._S.blocks) == set(a._sym.chars)
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-11
    for li in range(a.layer_N):
        assert relfro(a.layer_S11[li].numpy(), b.layer_S11[li].numpy()) <= 1e-11


def test_mirror_subgroup_when_only_one_centre_agrees(cpu_double):
    """Two centred bars, the second shifted in y: the x mirror survives, the y mirror and C2 do not."""
    case = dict(C.CASES["ex1_o3"])
    case["layers"] = [C._rect(), dict(C._rect(d=120.0), Cy=110.0)]

    def run(sym):
        return C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=sym, **kw), case, torch.complex128)
    a, b = run(True), run(False)
    assert a._sym.gens == ("x",)
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-11


def test_block_s_entries_and_lazy_dense(cpu_double):
    """S_parameters reads its entries off the symmetry blocks; the dense blocks appear only when asked for."""
    sim = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), C.CASES["ex1_o3"], torch.complex128)
    assert sim._S._cache == {}
    sp = C.probe(sim)
    assert sim._S._cache == {}                       # the readout did not densify anything
    n = 2 * sim.order_N
    a = torch.arange(n)[:, None]
    b = torch.arange(n)[None, :]
    for k in range(4):
        assert float((sim._S.entries(k, a, b)[0] - sim.S[k]).abs().max()) <= 1e-15
    assert len(sim.S) == 4 and len(list(sim.S)) == 4 and sim.S[-1].shape == (n, n) and set(sim._S._cache) == {0, 1, 2, 3}
    gen = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=False, **kw), C.CASES["ex1_o3"], torch.complex128)
    assert np.abs(sp - C.probe(gen)).max() <= 1e-12


@pytest.mark.parametrize("sym", [None, False])
def test_finished_simulation_is_freed_without_the_cycle_collector(cpu_double, sym):
    """No reference cycle through the lazy S / Sin / Sout views: a sweep loop must release each step's multi-GB blocks when
    the simulation object goes out of scope, not when the cyclic GC happens to run (measured: the allocator's reserved pool
    grew 13 GB per step and the step time doubled)."""
    import gc
    import weakref
    gc.collect()
    gc.disable()
    try:
        sim = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=sym, **kw), C.CASES["ex1_o3"], torch.complex128)
        C.probe(sim)
        _ = sim.S[0], sim.Sin[0]
        ref = weakref.ref(sim)
        del sim
        assert ref() is None
    finally:
        gc.enable()


@pytest.mark.parametrize("order,inc,azi,expect", [([4, 2], 0.0, 0.0, ("x", "y")), ([2, 3], 0.35, 0.0, ("y",)), ([3, 2], 0.35, np.pi / 2, ("x",)),
                                                   ([3, 3], 0.35, 0.6, None)])
def test_symmetry_follows_the_illumination_and_rectangular_truncations(cpu_double, order, inc, azi, expect):
    """A doubly mirror-symmetric cell: normal incidence keeps both mirrors, incidence in the xz plane (ky0 = 0) only the
    mirror in y, in the yz plane only the mirror in x, a skew azimuth none; rectangular truncations (ox != oy).  Whatever
    is used, the result equals the general path."""
    case = dict(C.CASES["ex1_o3"], order=order, inc=inc, azi=azi, eps_out=2.1)          # with an output half space (Sout on the right)
    mk = lambda sym: C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=sym, **kw), case, torch.complex128)
    a, b = mk(None), mk(False)
    found = a._sym.gens if a._sym not in (None, False) else None
    assert found == expect
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-11
    assert np.abs(C.probe(a) - C.probe(b)).max() <= 1e-12


@pytest.mark.parametrize("name", ["ex1_o3", "c2_o3", "ymirror_o3", "offcentre_o3", "square_o4"])
def test_half_spaces_are_pair_sparse_in_the_adapted_basis(cpu_double, name, monkeypatch):
    """Every half-space block, projected, is a diagonal plus one partner entry per row (also with an off-centre cell, where the
    basis carries translation phases): the cheap star products are the ones that run, and PairSparse reproduces the block."""
    from torcwa_b200 import symmetry
    seen = []
    real = symmetry.PairSparse.from_dense

    def spy(M, tol=1e-12):
        out = real(M, tol)
        seen.append((M, out))
        return out
    monkeypatch.setattr(symmetry.PairSparse, "from_dense", staticmethod(spy))
    sim = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), C.CASES[name], torch.complex128)
    assert sim._sym not in (None, False) and len(seen) >= 4 * len(sim._sym.chars)      # Sin at least; Sout and homogeneous layers if present
    eye = None
    for M, sp in seen:
        assert bool(sp.ok)
        eye = torch.eye(M.shape[1], dtype=M.dtype).expand(M.shape[0], -1, -1)
        assert float((sp.left(eye) - M).abs().max()) <= 1e-12 * float(M.abs().max())
        assert float((sp.right(eye) - M).abs().max()) <= 1e-12 * float(M.abs().max())
        assert float((sp.add_to(torch.zeros_like(M)) - M).abs().max()) <= 1e-12 * float(M.abs().max())


@pytest.mark.parametrize("first_homogeneous", [False, True])
def test_symmetric_stack_with_homogeneous_layers(cpu_double, first_homogeneous, monkeypatch):
    """Config 3's shape at small size, normal incidence: rotated bars (C2) separated by homogeneous spacers, lossy one
    included, output half space.  The spacers and half spaces go through the pair-sparse star products (never the dense
    fallback); the result equals the general path."""
    from torcwa_b200 import symmetry
    case = dict(C.CASES["ex1_o3"], lam=650.0, eps_out=2.1)
    case["layers"] = C._stack()[1:] if first_homogeneous else C._stack()
    oks = []
    real = symmetry.PairSparse.from_dense

    def spy(M, tol=1e-12):
        out = real(M, tol)
        oks.append(bool(out.ok))
        return out
    monkeypatch.setattr(symmetry.PairSparse, "from_dense", staticmethod(spy))
    a = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), case, torch.complex128)
    n_spacers_after_first = 4 - (1 if first_homogeneous else 0)
    assert a._sym.gens == ("c2",) and all(oks) and len(oks) == 2 * (2 * n_spacers_after_first + 8)
    b = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, symmetry_reduction=False, **kw), case, torch.complex128)
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-11
    assert np.abs(C.probe(a) - C.probe(b)).max() <= 1e-12


def test_pair_sparse_star_products_equal_the_dense_routine(cpu_double):
    import fake_lib
    from torcwa_b200 import symmetry
    g = torch.Generator().manual_seed(3)
    n, B = 14, 2
    rnd = lambda *s: torch.complex(torch.randn(*s, generator=g, dtype=torch.float64), torch.randn(*s, generator=g, dtype=torch.float64))
    p = torch.tensor([1, 0, 2, 5, 6, 3, 4, 7, 9, 8, 13, 12, 11, 10])          # an involution with fixed points 2 and 7

    def sparse_dense():
        M = torch.diag_embed(rnd(B, n))
        o = rnd(B, n) * (p != torch.arange(n))
        M[:, torch.arange(n), p] += o
        return M
    half = [0.3 * sparse_dense() for _ in range(4)]
    S = [0.3 * rnd(B, n, n) for _ in range(4)]
    sp = [symmetry.PairSparse.from_dense(h) for h in half]
    assert all(bool(x.ok) for x in sp)
    for got, want in ((symmetry.redheffer_sparse_left(fake_lib, sp, S)[0], fake_lib.redheffer(half, S)[0]),
                      (symmetry.redheffer_sparse_right(fake_lib, S, sp)[0], fake_lib.redheffer(S, half)[0])):
        for k in range(4):
            assert relfro(got[k].numpy(), want[k].numpy()) <= 1e-13
    assert not bool(symmetry.PairSparse.from_dense(S[0]).ok)                         # a dense matrix is not mistaken for one


@pytest.mark.parametrize("bump,expect", [(0.0, ("x", "y")), (1e-7, None), (1e-3, None)])
def test_nearly_symmetric_cells_are_not_reduced(cpu_double, bump, expect):
    """The symmetry must hold to 1e-11 of the largest Fourier coefficient to be used: a cell that is symmetric only to 1e-7
    would otherwise pick up an error of that order, far above the 1e-10 parity gate.  Either way the result equals the
    general path."""
    case = C.CASES["ex1_o3"]
    cd = torch.complex128
    d0, grid0 = C.build_layers(case, cd)[0]
    grid = grid0.clone()
    grid[40:60, 200:230] += bump                      # an off-centre patch: breaks both mirrors and C2

    def run(sym):
        sim = cpu_double.rcwa(freq=torch.tensor(1.0 / case["lam"], dtype=torch.float64), order=case["order"], L=case["L"], dtype=cd, device=CPU,
                              symmetry_reduction=sym)
        sim.add_input_layer(eps=case["eps_in"])
        sim.set_incident_angle(0.0, 0.0)
        sim.add_layer(thickness=d0, eps=grid)
        sim.solve_global_smatrix()
        return sim
    a, b = run(None), run(False)
    assert (a._sym.gens if a._sym not in (None, False) else None) == expect
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-11


@pytest.mark.parametrize("batched", [False, True])
def test_fields_and_resolve_after_a_symmetry_reduced_solve(cpu_double, batched):
    """Solve, add layers, solve again, then ask for fields: the mode coefficients are carried in the original basis from the
    unprojected block layers, the global S-matrix densifies on demand.  Everything equals the general path."""
    import sys
    import torcwa_b200.fields as F
    F._lib = sys.modules["fake_lib"]
    try:
        case = C.CASES["ex1_o3"]
        cd = torch.complex128
        d0, grid0 = C.build_layers(case, cd)[0]
        x = torch.linspace(0, 300, 7, dtype=torch.float64)
        z = torch.linspace(-50, 500, 9, dtype=torch.float64)

        def run(sym):
            lam = torch.tensor([case["lam"], 600.0], dtype=torch.float64) if batched else torch.tensor(case["lam"], dtype=torch.float64)
            sim = cpu_double.rcwa(freq=1.0 / lam, order=case["order"], L=case["L"], dtype=cd, device=CPU, symmetry_reduction=sym,
                                  **({"store_intermediates": True} if batched else {}))
            sim.add_input_layer(eps=case["eps_in"])
            sim.set_incident_angle(0.0, 0.0)
            sim.add_layer(thickness=d0, eps=grid0)
            sim.solve_global_smatrix()
            out = [sim.S_parameters(orders=[0, 0], polarization="xx")]
            sim.add_layer(thickness=70.0, eps=2.3)
            sim.add_layer(thickness=50.0, eps=grid0 * 0.7 + 0.3)
            sim.solve_global_smatrix()
            out.append(sim.S_parameters(orders=[0, 0], polarization="yy"))
            sim.source_planewave(amplitude=[1.0, 0.0], direction="forward")
            (Ex, Ey, Ez), (Hx, Hy, Hz) = sim.field_xz(x, z, 150.0)
            (Ex2, _, _), _ = sim.field_xy(0, x, x, z_prop=20.0)
            return sim, out + [Ex, Ez, Hy, Ex2]
        a, fa = run(None)
        b, fb = run(False)
        assert a._sym not in (None, False) and b._sym is None
        for u, v in zip(fa, fb):
            assert float((u - v).abs().max()) <= 1e-11 * max(float(v.abs().max()), 1.0)
    finally:
        F._lib = sys.modules["torcwa_b200._lib"]


@pytest.mark.parametrize("seed", range(24))
def test_random_symmetric_cells_equal_the_general_path(cpu_double, seed):
    """Randomised cells built to have a given symmetry about a random centre (two mirrors, one mirror, or C2 only, from pairs of
    tilted bars), random rectangular truncation, lossy materials, incidence chosen to keep or to break the symmetry, spacer
    layer and output half space at random: whatever the detection decides, the result equals the general path, and when the
    illumination allows it the expected group is found."""
    import math
    from oracle.cases import rectangle_grid
    g = torch.Generator().manual_seed(9000 + seed)
    u = lambda a, b: float(a + (b - a) * torch.rand((), generator=g))
    L = [300.0, 360.0]
    nxy = [90, 108]
    px, py = L[0] / nxy[0], L[1] / nxy[1]
    kind = ("xy", "x", "y", "c2")[seed % 4]
    # centres on the half-pixel lattice: the sampled grid is then symmetric to round-off, like a designed cell
    x0, y0 = 0.5 * px * round(u(130.0, 170.0) / (0.5 * px)), 0.5 * py * round(u(150.0, 210.0) / (0.5 * py))     # everything stays inside the cell
    order = [int(torch.randint(1, 4, (), generator=g)), int(torch.randint(1, 4, (), generator=g))]
    bar = lambda cx, cy, wx, wy, th: rectangle_grid(L[0], L[1], nxy[0], nxy[1], wx, wy, cx, cy, th, 1000.0, torch.float64)
    wx, wy, dx, dy, th = u(30.0, 60.0), u(30.0, 60.0), u(10.0, 25.0), u(15.0, 35.0), u(0.2, 1.2)
    if kind == "xy":
        mask = bar(x0, y0, wx, wy, 0.0)
    elif kind == "x":          # mirror x -> 2 x0 - x only: two bars at x0 -+ dx with opposite tilt, and a small bar off to one side in y
        mask = torch.clamp(bar(x0 - dx - wx / 2, y0, wx, wy, th) + bar(x0 + dx + wx / 2, y0, wx, wy, -th) + bar(x0, y0 + 1.5 * dy + wy, 30.0, 20.0, 0.0), max=1.0)
    elif kind == "y":
        mask = torch.clamp(bar(x0, y0 - dy - wy / 2, wx, wy, th) + bar(x0, y0 + dy + wy / 2, wx, wy, -th) + bar(x0 + 1.5 * dx + wx, y0, 20.0, 30.0, 0.0), max=1.0)
    else:                      # inversion centre only: one tilted bar
        mask = bar(x0, y0, wx, wy, th)
    eps = complex(u(6.0, 14.0), u(0.0, 0.8))
    grid = (mask * eps + (1.0 - mask) * 1.0).to(torch.complex128)
    keep = seed % 3 != 2       # every third trial breaks the symmetry with a skew incidence
    inc = 0.0 if (kind in ("xy", "c2") and keep) else u(0.1, 0.5)
    azi = (math.pi / 2 if kind == "x" else 0.0) if keep else u(0.3, 1.2)
    if kind == "xy" and keep and seed % 8 == 4:
        inc, azi = u(0.1, 0.5), 0.0                                  # incidence in the xz plane: only the y mirror survives
    expect = None if not keep else {"xy": ("y",) if inc else ("x", "y"), "x": ("x",), "y": ("y",), "c2": ("c2",)}[kind]
    d = 173.0 + 10.0 * seed
    sims = []
    for sym in (None, False):
        sim = cpu_double.rcwa(freq=torch.tensor(1.0 / 560.0, dtype=torch.float64), order=order, L=L, dtype=torch.complex128, device=CPU,
                              symmetry_reduction=sym)
        sim.add_input_layer(eps=2.1)
        if seed % 2:
            sim.add_output_layer(eps=complex(1.8, 0.0))
        sim.set_incident_angle(inc, azi)
        sim.add_layer(thickness=d, eps=grid)
        if seed % 3 == 1:
            sim.add_layer(thickness=60.0, eps=complex(2.4, 0.1))
            sim.add_layer(thickness=d * 0.5, eps=grid * 0.6 + 0.4)
        sim.solve_global_smatrix()
        sims.append(sim)
    a, b = sims
    found = a._sym.gens if a._sym not in (None, False) else None
    assert found == expect, (kind, inc, azi, found, expect)
    for k in range(4):
        assert relfro(a.S[k].numpy(), b.S[k].numpy()) <= 1e-10


def test_unanalysed_layers_send_the_stack_to_the_general_path(cpu_double):
    """A symmetric patterned layer (solved in blocks) followed by (a) a layer with patterned permeability, (b) a layer on the
    differentiable pipeline: neither is analysed for symmetry, so the stack is cascaded in the original basis -- same
    S-matrix as with the reduction switched off, and the gradient still flows."""
    case = C.CASES["ex1_o3"]
    cd = torch.complex128
    d0, grid0 = C.build_layers(case, cd)[0]
    g = torch.Generator().manual_seed(11)
    mu_grid = torch.complex(1.0 + 0.4 * torch.rand(grid0.shape, generator=g, dtype=torch.float64), torch.zeros(grid0.shape, dtype=torch.float64))
    rho0 = torch.rand(grid0.shape, generator=g, dtype=torch.float64)

    def run(sym, kind):
        sim = cpu_double.rcwa(freq=torch.tensor(1.0 / case["lam"], dtype=torch.float64), order=case["order"], L=case["L"], dtype=cd, device=CPU,
                              symmetry_reduction=sym)
        sim.add_input_layer(eps=case["eps_in"])
        sim.set_incident_angle(0.0, 0.0)
        sim.add_layer(thickness=d0, eps=grid0)
        first_in_blocks = sim._sym not in (None, False)
        rho = rho0.clone().requires_grad_(kind == "diff")
        if kind == "mu":
            sim.add_layer(thickness=80.0, eps=grid0 * 0.5 + 1.0, mu=mu_grid)
        else:
            sim.add_layer(thickness=80.0, eps=(1.0 + 3.0 * rho).to(cd))
        sim.solve_global_smatrix()
        t = torch.cat([sim.S_parameters(orders=[0, 0], polarization=p) for p in ("xx", "yy", "yx")])    # (1, 0) is evanescent here
        grad = None
        if kind == "diff":
            (t.abs() ** 2).sum().backward()
            grad = rho.grad.clone()
        return sim, t.detach(), grad, first_in_blocks
    for kind in ("mu", "diff"):
        a, ta, ga, blocks_a = run(None, kind)
        b, tb, gb, blocks_b = run(False, kind)
        assert blocks_a and not blocks_b and a._sym is False
        assert float((ta - tb).abs().max()) <= 1e-12
        for k in range(4):
            assert relfro(a.S[k].detach().numpy(), b.S[k].detach().numpy()) <= 1e-11
        if kind == "diff":
            assert float((ga - gb).abs().max()) <= 1e-10 * float(gb.abs().max())


def test_symmetry_reduced_batched_sweep(cpu_double):
    """Batched (no stored intermediates): per-point results of the block path == the general path."""
    case = C.CASES["ex1_o3"]
    cd = torch.complex128
    lams = torch.tensor([500.0, 532.0, 610.0], dtype=torch.float64)
    d, grid = C.build_layers(case, cd)[0]
    grids = torch.stack([grid, grid * 0.9 + 0.1, grid])
    out = []
    for sym in (True, False):
        sim = cpu_double.rcwa(freq=1 / lams, order=case["order"], L=case["L"], dtype=cd, device=CPU, symmetry_reduction=sym)
        sim.add_input_layer(eps=case["eps_in"])
        sim.set_incident_angle(0.0, 0.0)
        sim.add_layer(thickness=torch.tensor([300.0, 250.0, 300.0]), eps=grids)
        sim.add_layer(thickness=50.0, eps=2.25)
        sim.solve_global_smatrix()
        out.append(torch.stack([sim.S_parameters(orders=[[0, 0], [1, 0], [-1, 1]], polarization=p, port=q)
                                for p in ("xx", "yx", "pp") for q in ("transmission", "reflection")]))
        assert (sim._sym not in (None, False)) == sym
    assert float((out[0] - out[1]).abs().max()) <= 1e-12


def test_half_space_blocks_match_oracle_dense(cpu_double):
    case = C.CASES["stack_o3"]
    sim = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), case, torch.complex128)
    ref = C.run_case(lambda **kw: OracleSim(**kw), case, torch.complex128)
    for k in range(4):
        assert relfro(sim.Sin[k].numpy(), ref.Sin[k].numpy()) < 1e-12
        assert relfro(sim.Sout[k].numpy(), ref.Sout[k].numpy()) < 1e-12
    assert relfro(sim.Vf.numpy(), ref.Vf.numpy()) < 1e-13
    assert relfro(sim.Kx_norm.numpy(), ref.Kx_norm.numpy()) < 1e-15
    # homogeneous lossy layer (index 3 of the stack): analytic 2x2-block path vs dense oracle
    for li in (1, 3):
        for k, blk in enumerate((sim.layer_S11[li], sim.layer_S21[li], sim.layer_S12[li], sim.layer_S22[li])):
            assert relfro(blk.numpy(), ref.layer_S[li][k].numpy()) < 1e-11, (li, k)
        assert relfro(sim.kz_norm[li].numpy(), ref.kz_norm[li].numpy()) < 1e-13


def test_batched_shapes_and_values(cpu_double):
    case = C.CASES["ex1_o3"]
    cd = torch.complex128
    lams = torch.tensor([500.0, 532.0], dtype=torch.float64)
    d, grid = C.build_layers(case, cd)[0]
    sim = cpu_double.rcwa(freq=1 / lams, order=case["order"], L=case["L"], dtype=cd, device=CPU)
    sim.add_input_layer(eps=case["eps_in"])
    sim.set_incident_angle(0.0, 0.0)
    sim.add_layer(thickness=d, eps=grid)
    sim.solve_global_smatrix()
    t = sim.S_parameters(orders=[[0, 0], [1, 0], [0, 1]], polarization="xx")
    assert t.shape == (2, 3) and sim.S[0].shape == (2, 98, 98)
    ref = C.run_case(lambda **kw: OracleSim(**kw), case, cd)
    assert abs(complex(t[1, 0]) - complex(ref.S_parameters([0, 0])[0])) < 1e-11


def test_warnings_match_reference_messages(cpu_double):
    sim = cpu_double.rcwa(freq=1 / 532.0, order=[1, 1], L=[300.0, 300.0], dtype=torch.complex128, device=CPU)
    with pytest.warns(UserWarning, match="Invalid angle layer"):
        sim.set_incident_angle(0.0, 0.0, angle_layer="sideways")
    sim.solve_global_smatrix()
    with pytest.warns(UserWarning, match="Invalid polarization"):
        sim.S_parameters([0, 0], polarization="zz")
    with pytest.warns(UserWarning, match="Invalid port"):
        sim.S_parameters([0, 0], port="sideways")
    with pytest.warns(UserWarning, match="Invalid propagation direction"):
        sim.S_parameters([0, 0], direction="up")
    o = torch.tensor([[9, 0]])
    sim.S_parameters(o)
    assert o.tolist() == [[1, 0]]           # out-of-range orders are clamped in place (rcwa.py:1115-1122)


def test_finished_simulation_is_freed_by_refcount(cpu_double):
    """A batched sweep allocates tens of GB per simulation object; it must die with its last reference, not
    whenever the cyclic garbage collector runs (a sim <-> sim.Sin cycle once kept two steps' S-matrices alive)."""
    import gc
    import weakref
    gc.disable()
    try:
        sim = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), C.CASES["stack_o3"], torch.complex128)
        _ = sim.Sin[0], sim.Sout[0]
        ref = weakref.ref(sim)
        del sim
        assert ref() is None
    finally:
        gc.enable()


def test_differentiable_pipeline_matches_reference_autograd(cpu_double, golden_dir):
    """Gradients of a topology-optimisation style figure of merit w.r.t. the density grid and the layer thickness:
    torcwa_b200's differentiable path (autograd wrappers around the C-ABI ops, here the CPU double) == the unmodified
    reference's autograd (tests/golden/autograd_o3.npz, tools/make_golden_autograd.py)."""
    from oracle.autograd_case import CASE, fom
    g = np.load(os.path.join(golden_dir, "autograd_o3.npz"))
    rho = torch.from_numpy(g["rho"]).clone().requires_grad_(True)
    thick = torch.tensor(CASE["thickness"], dtype=torch.float64, requires_grad=True)
    sim = cpu_double.rcwa(freq=torch.tensor(1.0 / CASE["lam"], dtype=torch.float64), order=CASE["order"], L=CASE["L"],
                          dtype=torch.complex128, device=CPU)
    value = fom(sim, rho, thick)
    value.backward()
    assert abs(float(value.detach()) - float(g["fom"])) <= 1e-10 * abs(float(g["fom"]))
    assert np.linalg.norm(rho.grad.numpy() - g["grad_rho"]) <= 1e-8 * np.linalg.norm(g["grad_rho"])
    assert abs(float(thick.grad) - float(g["grad_thickness"])) <= 1e-8 * abs(float(g["grad_thickness"]))


def test_diffraction_angle_and_return_layer_match_reference(cpu_double, golden_dir):
    g = np.load(os.path.join(golden_dir, "misc_stack_o3.npz"))
    sim = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), C.CASES["stack_o3"], torch.complex128)
    orders = g["orders"].tolist()
    for layer in ("input", "output"):
        for unit in ("radian", "degree"):
            inc, azi = sim.diffraction_angle(orders, layer=layer, unit=unit)
            np.testing.assert_allclose(inc.numpy(), g["inc_%s_%s" % (layer, unit)], rtol=0, atol=1e-13)
            np.testing.assert_allclose(azi.numpy(), g["azi_%s_%s" % (layer, unit)], rtol=0, atol=1e-13)
    e, m = sim.return_layer(0, nx=20, ny=26)
    assert np.abs(e.numpy() - g["eps_rec"]).max() <= 1e-12 * np.abs(g["eps_rec"]).max()
    assert np.abs(m.numpy() - g["mu_rec"]).max() <= 1e-12
    with pytest.warns(UserWarning):
        sim.diffraction_angle(orders, layer="sideways")


def test_sources_and_fields_match_reference(cpu_double, golden_dir):
    """source_planewave / source_fourier (xy and ps notation, forward and backward) and field_xz / field_yz / field_xy
    in the half spaces and inside patterned and homogeneous layers == the unmodified reference
    (tests/golden/fields_stack_o3.npz, tools/make_golden_fields.py).  The mode coefficients are built lazily from the
    stored layer intermediates; the fused forward path is untouched."""
    from oracle.fields_case import build, SOURCES, planes
    import sys
    import torcwa_b200.fields as F
    F._lib = sys.modules["fake_lib"]
    try:
        g = np.load(os.path.join(golden_dir, "fields_stack_o3.npz"))
        sim = build(lambda **kw: cpu_double.rcwa(device=CPU, **kw))
        for sname, setter in SOURCES.items():
            setter(sim)
            for pname, getter in planes().items():
                E, H = getter(sim)
                got = np.stack([t.numpy() for t in E + H])
                ref = g["%s_%s" % (sname, pname)]
                err = np.abs(got - ref).max() / np.abs(ref).max()
                assert err <= 1e-9, (sname, pname, err)
        assert len(sim.Cf) == 3 and len(sim.C[0]) == 3 and sim.H_eigvec[0].shape == (98, 98)
    finally:
        from torcwa_b200 import _lib as real
        F._lib = real


def test_pinv_instability_metrics_are_reported(cpu_double):
    """avoid_Pinv_instability=True (rcwa.py:1249-1262): one metric per patterned layer, at round-off level for a
    well-conditioned cell; the S-parameters are the same as without the flag (this path never forms P^-1)."""
    case = C.CASES["ex1_o3"]
    a = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, avoid_Pinv_instability=True, **kw), case, torch.complex128)
    b = C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), case, torch.complex128)
    assert len(a.Pinv_instability) == 1 and len(a.Qinv_instability) == 1
    assert 0.0 <= float(a.Pinv_instability[0]) < 1e-8 and 0.0 <= float(a.Qinv_instability[0]) < 1e-8
    assert b.Pinv_instability is None
    assert np.abs(C.probe(a) - C.probe(b)).max() == 0.0


def test_numerical_failure_is_raised_once_at_solve(cpu_double, monkeypatch):
    """A non-zero info word of any stage surfaces as torch.linalg.LinAlgError when the global S-matrix is complete."""
    import fake_lib
    real_eig = fake_lib.eig

    def failing_eig(A):
        w, V, info = real_eig(A)
        info = info.clone(); info[0] = 7
        return w, V, info
    monkeypatch.setattr(fake_lib, "eig", failing_eig)
    with pytest.raises(torch.linalg.LinAlgError, match="eigendecomposition"):
        C.run_case(lambda **kw: cpu_double.rcwa(device=CPU, **kw), C.CASES["ex1_o3"], torch.complex128)


@pytest.mark.parametrize("seed", range(12))
def test_random_stacks_match_oracle(cpu_double, seed):
    """Randomised host-logic parity: random truncation orders, periods, wavelengths, incidence angles (referred to the
    input or the output half space), lossy half spaces, mixed patterned / homogeneous layers with complex permittivity,
    one layer with patterned permeability -- torcwa_b200.rcwa (CPU double of the C ABI) against the oracle's dense
    restatement of the reference, on S-parameters in all polarisation notations.  (On these same twelve stacks the
    oracle equals the live reference bit for bit -- checked in the build container when the test was written.)"""
    g = torch.Generator().manual_seed(1000 + seed)

    def u(a, b):
        return float(a + (b - a) * torch.rand((), generator=g, dtype=torch.float64))
    cd = torch.complex128
    order = [int(torch.randint(1, 3, (), generator=g)), int(torch.randint(1, 3, (), generator=g))]
    L = [u(250.0, 700.0), u(250.0, 700.0)]
    lam = u(400.0, 900.0)
    nx, ny = 4 * order[0] + 1 + int(torch.randint(0, 6, (), generator=g)), 4 * order[1] + 1 + int(torch.randint(0, 6, (), generator=g))
    eps_in = None if seed % 4 == 0 else complex(u(1.0, 3.0), u(0.0, 0.2) if seed % 3 == 0 else 0.0)
    eps_out = None if seed % 4 == 1 else complex(u(1.0, 4.0), u(0.0, 0.3) if seed % 5 == 0 else 0.0)
    inc, azi = u(0.0, 0.6), u(-1.5, 1.5)
    angle_layer = "output" if (seed % 2 == 1 and eps_out is not None) else "input"
    layers = []
    for k in range(int(torch.randint(0, 4, (), generator=g))):
        d = u(20.0, 300.0)
        if torch.rand((), generator=g) < 0.4:
            layers.append((d, complex(u(1.0, 6.0), u(0.0, 0.5)), 1.0))
        else:
            grid = torch.complex(1.0 + 8.0 * torch.rand(nx, ny, generator=g, dtype=torch.float64), 0.3 * torch.rand(nx, ny, generator=g, dtype=torch.float64))
            mu = 1.0
            if k == 1:
                mu = torch.complex(1.0 + 0.5 * torch.rand(nx, ny, generator=g, dtype=torch.float64), torch.zeros(nx, ny, dtype=torch.float64))
            layers.append((d, grid, mu))

    def run(factory):
        sim = factory(freq=torch.tensor(1.0 / lam, dtype=torch.float64), order=order, L=L, dtype=cd)
        if eps_in is not None:
            sim.add_input_layer(eps=eps_in)
        if eps_out is not None:
            sim.add_output_layer(eps=eps_out)
        sim.set_incident_angle(inc_ang=inc, azi_ang=azi, angle_layer=angle_layer)
        for d, e, m in layers:
            sim.add_layer(thickness=d, eps=e, mu=m)
        sim.solve_global_smatrix()
        return C.probe(sim)
    mine = run(lambda **kw: cpu_double.rcwa(device=CPU, **kw))
    ref = run(lambda **kw: OracleSim(**kw))
    scale = max(np.abs(ref).max(), 1e-30)
    assert np.abs(mine - ref).max() <= 1e-9 * scale, np.abs(mine - ref).max() / scale


def test_batched_fields_equal_per_point_fields(cpu_double):
    """Fields of a batched sweep (three wavelengths, per-point thickness) == the fields of three separate unbatched
    simulations, plane by plane."""
    import sys
    import torcwa_b200.fields as F
    F._lib = sys.modules["fake_lib"]
    try:
        case = C.CASES["stack_o3"]
        cd = torch.complex128
        lams = torch.tensor([600.0, 650.0, 710.0], dtype=torch.float64)
        layers = C.build_layers(case, cd)[:2]
        thick0 = torch.tensor([200.0, 180.0, 230.0], dtype=torch.float64)
        x = torch.linspace(0.0, 300.0, 5, dtype=torch.float64)
        z = torch.tensor([-30.0, 50.0, 190.0, 250.0, 400.0], dtype=torch.float64)

        def solve(freq, t0):
            sim = cpu_double.rcwa(freq=freq, order=case["order"], L=case["L"], dtype=cd, device=CPU, store_intermediates=True)
            sim.add_input_layer(eps=case["eps_in"]); sim.add_output_layer(eps=case["eps_out"])
            sim.set_incident_angle(inc_ang=case["inc"], azi_ang=case["azi"])
            sim.add_layer(thickness=t0, eps=layers[0][1])
            sim.add_layer(thickness=layers[1][0], eps=layers[1][1])
            sim.solve_global_smatrix()
            sim.source_fourier(amplitude=[[0.7, 0.2], [0.1, -0.4j]], orders=[[0, 0], [1, -1]], direction="forward", notation="ps")
            return sim
        sb = solve(1 / lams, thick0)
        Eb, Hb = sb.field_xz(x, z, 40.0)
        Exy, Hxy = sb.field_xy(0, x, x, 60.0)
        assert Eb[0].shape == (3, 5, 5)
        for b in range(3):
            s1 = solve(1 / lams[b], float(thick0[b]))
            E1, H1 = s1.field_xz(x, z, 40.0)
            E2, H2 = s1.field_xy(0, x, x, 60.0)
            for k in range(3):
                assert float((Eb[k][b] - E1[k]).abs().max()) <= 1e-11 * float(E1[k].abs().max() + 1e-30)
                assert float((Hb[k][b] - H1[k]).abs().max()) <= 1e-11 * float(H1[k].abs().max() + 1e-30)
                assert float((Exy[k][b] - E2[k]).abs().max()) <= 1e-11 * float(E2[k].abs().max() + 1e-30)
                assert float((Hxy[k][b] - H2[k]).abs().max()) <= 1e-11 * float(H2[k].abs().max() + 1e-30)
    finally:
        from torcwa_b200 import _lib as real
        F._lib = real
