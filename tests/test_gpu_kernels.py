"""Stage-level parity of the CUDA kernels (through the C ABI) against torch fp64 / the oracle.
All tests need a GPU:  python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cases as C  # noqa: E402
from oracle.rcwa_oracle import OracleSim  # noqa: E402


def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.complex(torch.randn(*shape, generator=g, dtype=torch.float64),
                         torch.randn(*shape, generator=g, dtype=torch.float64)).to(dev())


def rel(a, b):
    return float(torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1)))


@pytest.mark.parametrize("opa", ["N", "T", "H"])
@pytest.mark.parametrize("opb", ["N", "T", "H"])
@pytest.mark.parametrize("mnk", [(64, 128, 8), (97, 130, 37), (1, 1, 1), (200, 33, 64), (33, 300, 5), (128, 64, 129)])
def test_zgemm(opa, opb, mnk):
    from torcwa_b200 import _lib
    M, N, K = mnk
    nb = 3
    A = rnd(nb, *((M, K) if opa == "N" else (K, M)), seed=1)
    B = rnd(nb, *((K, N) if opb == "N" else (N, K)), seed=2)
    Cin = rnd(nb, M, N, seed=3)
    f = {"N": lambda x: x, "T": lambda x: x.transpose(1, 2), "H": lambda x: x.transpose(1, 2).conj()}
    ref = (0.5 - 0.25j) * (f[opa](A) @ f[opb](B)) + (2.0 + 1.0j) * Cin
    out = Cin.clone()
    _lib.zgemm(A, B, opa, opb, alpha=0.5 - 0.25j, beta=2.0 + 1.0j, out=out)
    assert rel(out, ref) < 1e-14
    out0 = _lib.zgemm(A, B, opa, opb)
    assert rel(out0, f[opa](A) @ f[opb](B)) < 1e-14


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 8, 9, 10, 11, 12])
@pytest.mark.parametrize("ops", [("N", "N"), ("N", "H"), ("H", "N")])
@pytest.mark.parametrize("mnk", [(97, 130, 37), (200, 33, 64), (33, 300, 5), (130, 70, 129), (64, 64, 64)])
def test_zgemm_explicit_configs(cfg, ops, mnk):
    """Every tile / 3-multiplication configuration of the DMMA GEMM (rcwa_zgemm_batched_cfg), ragged edges included.
    The 3M product is accurate norm-wise (|err| <= c eps |A||B|), not component-wise: the bound is on the Frobenius norm."""
    from torcwa_b200 import _lib
    opa, opb = ops
    M, N, K = mnk
    nb = 3
    A = rnd(nb, *((M, K) if opa == "N" else (K, M)), seed=11)
    B = rnd(nb, *((K, N) if opb == "N" else (N, K)), seed=12)
    Cin = rnd(nb, M, N, seed=13)
    f = {"N": lambda x: x, "H": lambda x: x.transpose(1, 2).conj()}
    ref = (0.5 - 0.25j) * (f[opa](A) @ f[opb](B)) + (2.0 + 1.0j) * Cin
    out = Cin.clone()
    _lib.zgemm(A, B, opa, opb, alpha=0.5 - 0.25j, beta=2.0 + 1.0j, out=out, cfg=cfg)
    assert rel(out, ref) < 2e-14


def test_zgemm_inplace_window_updates():
    """The QR phase applies window unitaries in place: C aliases B (rows, M <= 64) or A (columns, N <= 64)."""
    from torcwa_b200 import _lib
    U = torch.linalg.qr(rnd(2, 64, 64, seed=21))[0].contiguous()
    Hrow = rnd(2, 64, 700, seed=22)
    ref = U.transpose(1, 2).conj() @ Hrow
    for cfg in (0, 2, 8, 10):
        X = Hrow.clone()
        _lib.zgemm(U, X, "H", "N", out=X, cfg=cfg)
        assert rel(X, ref) < 2e-14
    Z = rnd(2, 900, 64, seed=23)
    ref = Z @ U
    for cfg in (1, 2, 9, 10):
        X = Z.clone()
        _lib.zgemm(X, U, "N", "N", out=X, cfg=cfg)
        assert rel(X, ref) < 2e-14


def test_zgemm_large():
    from torcwa_b200 import _lib
    A, B = rnd(2, 500, 700, seed=4), rnd(2, 700, 450, seed=5)
    assert rel(_lib.zgemm(A, B), A @ B) < 1e-14


@pytest.mark.parametrize("n", [5, 32, 33, 100, 128, 129, 242, 500])
def test_lu_right_solve(n):
    from torcwa_b200 import _lib
    nb = 3
    A = rnd(nb, n, n, seed=n) + 0.5 * torch.eye(n, dtype=torch.complex128, device=dev())
    Bm = rnd(nb, 17, n, seed=n + 1)
    LU = A.clone()
    perm, info, tinv = _lib.lu_factor_(LU)
    assert int(info.abs().max()) == 0
    X = _lib.lu_solve_right(LU, perm, tinv, Bm)
    assert rel(X @ A, Bm) < 1e-11
    Ai, info = _lib.inverse(A)
    assert rel(Ai, torch.linalg.inv(A)) < 1e-10


def test_lu_singular_reports_info():
    from torcwa_b200 import _lib
    A = rnd(2, 40, 40, seed=9)
    A[1, 7, :] = 0
    _, info, _t = _lib.lu_factor_(A.clone())
    assert int(info[0]) == 0 and int(info[1]) > 0


@pytest.mark.parametrize("name,gdtype", [("ex1_o3", torch.float64), ("ex1_o3", torch.complex128), ("stack_o4x2", torch.complex64),
                                         ("stack_o4x2", torch.float32)])
def test_convmat_vs_oracle(name, gdtype):
    from torcwa_b200 import _lib
    case = C.CASES[name]
    cd = torch.complex128
    grid = C.build_layers(case, cd)[0][1]
    if not gdtype.is_complex:
        grid = grid.real.contiguous()
    grid = grid.to(gdtype)
    sim = OracleSim(freq=1 / case["lam"], order=case["order"], L=case["L"], dtype=cd)
    ref = sim.material_conv(grid.to(torch.complex128) if gdtype.is_complex else grid.to(torch.float64))
    E = _lib.convmat(grid.to(dev()), case["order"][0], case["order"][1])
    tol = 1e-13 if gdtype in (torch.float64, torch.complex128) else 1e-13   # inputs are exactly representable
    assert rel(E[0].cpu(), ref.to(torch.complex128)) < tol
    # batched with distinct grids
    g2 = torch.stack([grid, 2 * grid]).to(dev())
    E2 = _lib.convmat(g2, case["order"][0], case["order"][1])
    assert rel(E2[1].cpu(), 2 * ref.to(torch.complex128)) < tol


def _oracle_layer(name):
    case = C.CASES[name]
    sim = C.run_case(lambda **kw: OracleSim(**kw), case, torch.complex128)
    return case, sim


@pytest.mark.parametrize("name", ["ex1_o3", "stack_o3"])
def test_pq_and_layer_smatrix_vs_oracle(name):
    """Feed the oracle's eigenpairs to the CUDA layer-S stage: isolates stage 3a from stage 2."""
    from torcwa_b200 import _lib
    from torcwa_b200.rcwa import vf_inverse_diagonals
    case, sim = _oracle_layer(name)
    d = dev()
    N = sim.order_N
    E = sim.eps_conv[0].to(d)[None]
    eta, info = _lib.inverse(E)
    kx, ky = sim.Kx_norm_dn.to(d)[None].contiguous(), sim.Ky_norm_dn.to(d)[None].contiguous()
    mu = torch.ones(1, dtype=torch.complex128, device=d)
    P, Q = _lib.pq_assemble(eta, E, kx, ky, mu_scalar=mu)
    assert rel(P[0].cpu(), sim.P[0]) < 1e-12
    assert rel(Q[0].cpu(), sim.Q[0]) < 1e-12
    W, kz = sim.E_eigvec[0].to(d)[None].contiguous(), sim.kz_norm[0].to(d)[None].contiguous()
    vfinv = vf_inverse_diagonals(kx, ky)
    omega = torch.tensor([float(sim.omega)], dtype=torch.float64, device=d)
    thick = torch.tensor([float(sim.thickness[0])], dtype=torch.float64, device=d)
    S11, S21, info = _lib.layer_smatrix(W, kz, Q, vfinv, omega, thick)
    assert int(info.abs().max()) == 0
    ref = sim.layer_S[0]
    assert rel(S11[0].cpu(), ref[0]) < 1e-10 and rel(S21[0].cpu(), ref[1]) < 1e-10
    assert rel(S11[0].cpu(), ref[3]) < 1e-10 and rel(S21[0].cpu(), ref[2]) < 1e-10   # single-layer symmetry


def test_redheffer_vs_oracle():
    from torcwa_b200 import _lib
    case, sim = _oracle_layer("stack_o3")
    d = dev()
    Sm = [s.to(d)[None].contiguous() for s in sim.layer_S[0]]
    Sn = [s.to(d)[None].contiguous() for s in sim.layer_S[1]]
    out, info = _lib.redheffer(Sm, Sn)
    ref = sim._star(sim.layer_S[0], sim.layer_S[1])
    for k in range(4):
        assert rel(out[k][0].cpu(), ref[k]) < 1e-11
    # batched: two different pairs at once
    Sm2 = [torch.cat([a, b]) for a, b in zip(Sm, Sn)]
    Sn2 = [torch.cat([b, a]) for a, b in zip(Sm, Sn)]
    out2, _ = _lib.redheffer(Sm2, Sn2)
    ref2 = sim._star(sim.layer_S[1], sim.layer_S[0])
    for k in range(4):
        assert rel(out2[k][0].cpu(), ref[k]) < 1e-11
        assert rel(out2[k][1].cpu(), ref2[k]) < 1e-11


def test_redheffer_bdleft_vs_dense():
    from torcwa_b200 import _lib
    case, sim = _oracle_layer("stack_o3")
    d = dev()
    N = sim.order_N
    bd = [rnd(2, 4, N, seed=20 + k) * 0.3 for k in range(4)]
    Sn = [torch.stack([s, 0.5 * s]).to(d).contiguous() for s in sim.layer_S[0]]
    out, info = _lib.redheffer_bdleft(bd, Sn)
    ref, _ = _lib.redheffer([_lib.blockdiag_dense(x) for x in bd], Sn)
    for k in range(4):
        assert rel(out[k], ref[k]) < 1e-12


def test_blockdiag_dense():
    from torcwa_b200 import _lib
    d4 = rnd(2, 4, 9, seed=11)
    D = _lib.blockdiag_dense(d4)
    for b in range(2):
        ref = torch.cat([torch.cat([torch.diag(d4[b, 0]), torch.diag(d4[b, 1])], 1),
                         torch.cat([torch.diag(d4[b, 2]), torch.diag(d4[b, 3])], 1)], 0)
        assert torch.equal(D[b], ref)


# ------------------------------------------------------------------------------------------------ tcgen05 engine of the S-matrix stage
def _well_conditioned(nb, n, seed):
    """random matrices with a dominant diagonal (like the coupling matrices of the path: cond ~ 1e2-1e3)"""
    A = rnd(nb, n, n, seed=seed) / np.sqrt(n)
    return (A + 2.0 * torch.eye(n, dtype=torch.complex128, device=dev())).contiguous()


@pytest.mark.parametrize("order,gens", [((3, 3), ("x", "y")), ((4, 2), ("x",)), ((2, 5), ("y",)), ((3, 4), ("c2",)), ((15, 15), ("x", "y"))])
def test_sym_project_kernel_vs_index_arithmetic(order, gens):
    """rcwa_sym_project (one gather pass) == T_L^H X T_R by torch index_select / multiply / add (symmetry.Basis.project);
    the blocks of a random matrix carry no structure, so every gathered term matters.  Also the round trip through
    unproject for an operator that commutes with the group (built by symmetrising)."""
    from torcwa_b200 import _lib, symmetry
    basis = symmetry.Basis(order[0], order[1], gens, 0.3, -1.1, dev())
    nb = 2 if order[0] > 10 else 3
    X = rnd(nb, basis.n, basis.n, seed=5)
    for chi in basis.chars:
        for left, right in (("E", "E"), ("E", "H"), ("H", "E")):
            got = _lib.sym_project(X, *basis.tables(chi, left, right))
            want = basis.project(X, chi, left, right)
            assert got.shape == want.shape
            assert rel(got, want) <= 1e-14
    # symmetrised operator: sum_chi T S_chi T^H gives it back
    Xs = basis.unproject({chi: _lib.sym_project(X, *basis.tables(chi)) for chi in basis.chars})
    back = {chi: _lib.sym_project(Xs, *basis.tables(chi)) for chi in basis.chars}
    assert rel(basis.unproject(back), Xs) <= 1e-13
    a = torch.arange(basis.n, device=dev())
    assert rel(basis.entries(back, a[:, None], a[None, :]), Xs) <= 1e-13


@pytest.mark.parametrize("N,nb", [(169, 3), (961, 2)])
@pytest.mark.parametrize("slices,tol", [(5, 2e-7), (8, 2e-11)])
def test_layer_smatrix_and_redheffer_tc_vs_dmma(N, nb, slices, tol):
    """The S-matrix stage with its dense products on the tcgen05 int8-digit GEMM (gemm_slices = 5 / 8: products, and the
    512-wide right-looking triangular solves) against the same stage on the fp64 DMMA kernels (gemm_slices = 0)."""
    from torcwa_b200 import _lib
    d = dev()
    n = 2 * N
    W = _well_conditioned(nb, n, 31)
    Q = rnd(nb, n, n, seed=32) / np.sqrt(n)
    kz = (rnd(nb, n, seed=33) * 0.3 + 1.0).contiguous()
    kz = torch.complex(kz.real.abs() + 0.2, kz.imag.abs())
    vfinv = (rnd(nb, 4, N, seed=34) * 0.2 + torch.tensor([1.0, 0.0, 0.0, 1.0], dtype=torch.complex128, device=d)[None, :, None]).contiguous()
    omega = torch.full((nb,), 2 * np.pi / 532.0, dtype=torch.float64, device=d)
    thick = torch.full((nb,), 100.0, dtype=torch.float64, device=d)
    r11, r21, info0 = _lib.layer_smatrix(W, kz, Q, vfinv, omega, thick, slices=0)
    s11, s21, info1 = _lib.layer_smatrix(W, kz, Q, vfinv, omega, thick, slices=slices)
    assert int(info0.abs().max()) == 0 and int(info1.abs().max()) == 0
    e = max(rel(s11, r11), rel(s21, r21))
    print("layer S-matrix, %d digits vs fp64: %.2e" % (slices, e))
    assert e < tol
    Sm = [(0.4 * rnd(nb, n, n, seed=40 + k) / np.sqrt(n)).contiguous() for k in range(4)]
    Sn = [(0.4 * rnd(nb, n, n, seed=50 + k) / np.sqrt(n)).contiguous() for k in range(4)]
    ref, i0 = _lib.redheffer(Sm, Sn, slices=0)
    out, i1 = _lib.redheffer(Sm, Sn, slices=slices)
    assert int(i0.abs().max()) == 0 and int(i1.abs().max()) == 0
    e = max(rel(out[k], ref[k]) for k in range(4))
    print("star product, %d digits vs fp64: %.2e" % (slices, e))
    assert e < tol
    bd = [rnd(nb, 4, N, seed=60 + k) * 0.3 for k in range(4)]
    refb, _ = _lib.redheffer_bdleft(bd, Sn, slices=0)
    outb, _ = _lib.redheffer_bdleft(bd, Sn, slices=slices)
    assert max(rel(outb[k], refb[k]) for k in range(4)) < tol


def test_wrappers_follow_the_tensors_device():
    """A second GPU (when the box has one): the wrappers make the tensors' device current."""
    from torcwa_b200 import _lib
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    A = rnd(2, 70, 50, seed=1).to("cuda:1")
    B = rnd(2, 50, 40, seed=2).to("cuda:1")
    with torch.cuda.device(0):
        out = _lib.zgemm(A, B)
    assert out.device.index == 1 and rel(out, A @ B) < 1e-14
