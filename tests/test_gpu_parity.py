"""End-to-end parity of the B200 path (torcwa_b200.rcwa -> C ABI -> CUDA) against the reference's
stored outputs (tests/golden, generated from the unmodified reference by tools/make_golden.py).

Gates: complex128 vs reference-complex128 <= 1e-10; complex64 API vs
reference-complex128 <= 1e-4 (the reference's own complex64 run is ~3e-4 away from its complex128
run at order 15; both distances are printed)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cases as C  # noqa: E402

SMALL = ["ex1_o3", "ex1_o5", "stack_o3", "stack_o4x2", "fresnel_o2", "square_o4", "c2_o3", "ymirror_o3", "offcentre_o3"]
# BASELINE configs at their real sizes: corners / C4v-symmetric centre of config 4's (Wx, Wy, lambda) sweep at order 15, and
# config 3's 8-layer stack at order 21 (complex128 only)
SWEEP = ["sweep_o15_a", "sweep_o15_b", "sweep_o15_c"]


def b200_factory(freq, order, L, dtype, **kw):
    import torcwa_b200
    return torcwa_b200.rcwa(freq=freq, order=order, L=L, dtype=dtype, device=torch.device("cuda:0"), **kw)


def to_dev(case, cdtype):
    return case


def run(name, cdtype, **kw):
    case = C.CASES[name]
    # same driver as the golden generator, tensors moved to the GPU by the solver
    return C.run_case(lambda freq, order, L, dtype: b200_factory(freq, order, L, dtype, **kw), case, cdtype)


# cases whose cell + incidence have a symmetry the reduction (torcwa_b200/symmetry.py) finds: they run on BOTH paths
SYMMETRIC = {"ex1_o3", "ex1_o5", "square_o4", "c2_o3", "ymirror_o3", "offcentre_o3", "ex1_o15", "sweep_o15_a", "sweep_o15_b", "sweep_o15_c", "stack_o21"}


def paths(names):
    return [(n, s) for n in names for s in ((True, False) if n in SYMMETRIC else (False,))]


def relfro(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name,sym", paths(SMALL + SWEEP + ["stack_o21"]))
def test_parity_c128(name, sym, golden_dir):
    """sym = True: the symmetry-reduced block path (2 or 4 blocks); False: the general path."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sim = run(name, torch.complex128, symmetry_reduction=sym)
    assert (sim._sym not in (None, False)) == sym
    for info in sim.eig_info:
        assert int(info.abs().max()) == 0
    sp = C.probe(sim)
    scale = np.abs(g["sparams_c128"]).max()
    err = np.abs(sp - g["sparams_c128"]).max() / scale
    print(name, "sym=%s" % sym, "S-parameter max err / max|S|:", err)
    assert err <= 1e-10
    cols = g["S_cols_idx"]
    for k in range(4):
        assert relfro(sim.S[k][:, cols].cpu().numpy(), g["S_cols"][k]) <= 1e-10
    fro = np.array([float(torch.linalg.norm(s)) for s in sim.S])
    np.testing.assert_allclose(fro, g["S_fro"], rtol=1e-10)
    if "S" in g:
        for k in range(4):
            assert relfro(sim.S[k].cpu().numpy(), g["S"][k]) <= 1e-10
    if "eps_conv0" in g and sim.eps_conv:
        assert relfro(sim.eps_conv[0].cpu().numpy(), g["eps_conv0"]) <= 1e-13
        ls = [sim.layer_S11[0], sim.layer_S21[0], sim.layer_S12[0], sim.layer_S22[0]]
        for k in range(4):
            assert relfro(ls[k].cpu().numpy(), g["layer_S0"][k]) <= 1e-10
    if "kz2_sorted" in g:
        for l, kz in enumerate(sim.kz_norm):
            mine = np.sort_complex(kz.cpu().numpy().astype(np.complex128) ** 2)
            assert np.abs(mine - g["kz2_sorted"][l]).max() <= 1e-9 * np.abs(mine).max()


@pytest.mark.parametrize("name,sym", paths(SMALL + ["ex1_o15"] + SWEEP))
def test_parity_c64_api(name, sym, golden_dir):
    """complex64 API (fp64 eigensolver, S-matrix stage on the tcgen05 7-digit GEMM where the matrices are large enough
    -- at order 15 every dense product and triangular-solve update of the stage) against the reference's complex128 run."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sim = run(name, torch.complex64, symmetry_reduction=sym)
    assert sim._digits == 7 and (sim._sym not in (None, False)) == sym
    assert sim.S[0].dtype == torch.complex64
    sp = C.probe(sim)
    scale = np.abs(g["sparams_c128"]).max()
    new_vs_ref128 = np.abs(sp - g["sparams_c128"]).max() / scale
    ref64_vs_ref128 = np.abs(g["sparams_c64"] - g["sparams_c128"]).max() / scale
    new_vs_ref64 = np.abs(sp - g["sparams_c64"]).max() / scale
    print(f"{name} sym={sym}: new-c64 vs ref-c128 {new_vs_ref128:.2e} | ref-c64 vs ref-c128 {ref64_vs_ref128:.2e} | new-c64 vs ref-c64 {new_vs_ref64:.2e}")
    assert new_vs_ref128 <= 1e-4
    for k in range(4):
        assert relfro(sim.S[k][:, g["S_cols_idx"]].cpu().numpy().astype(np.complex128), g["S_cols"][k]) <= 1e-4


@pytest.mark.parametrize("digits,gate", [(0, 1e-10), (8, 1e-10), (7, 1e-9), (5, 1e-6), (4, 1e-4)])
def test_order15_parity_vs_digits_of_the_tcgen05_engine(digits, gate, golden_dir):
    """BASELINE config 2's size through the complex128 API with the S-matrix stage on the tcgen05 GEMM at 8 / 7 / 5 / 4
    digits (0 = fp64 DMMA): distance to the reference's complex128 run.  8 digits keep the complex128 gate; 4-5 are the
    complex64 API's engine."""
    import torcwa_b200
    g = np.load(os.path.join(golden_dir, "ex1_o15.npz"))
    sim = C.run_case(lambda freq, order, L, dtype: torcwa_b200.rcwa(freq=freq, order=order, L=L, dtype=dtype, device=torch.device("cuda:0"),
                                                                  gemm_digits=digits, symmetry_reduction=False), C.CASES["ex1_o15"], torch.complex128)
    sp = C.probe(sim)
    err = np.abs(sp - g["sparams_c128"]).max() / np.abs(g["sparams_c128"]).max()
    print("order 15, %d digits: S-parameter max err / max|S| = %.2e" % (digits, err))
    assert err <= gate


def test_pinv_metrics_and_dispersion_table_on_the_gpu(tmp_path):
    """Scope rows f3 / f2 on the device: avoid_Pinv_instability=True reports one round-off-level metric per patterned layer
    without changing the result (rcwa.py:1249-1262), and NKTable evaluates a sweep of wavelengths on the GPU exactly as
    scipy's cubic interpolant does on the host (example/Materials.py:5-50)."""
    import torcwa_b200
    from scipy.interpolate import interp1d
    from torcwa_b200.materials import NKTable
    case = C.CASES["ex1_o5"]
    mk = lambda **kw: (lambda freq, order, L, dtype: torcwa_b200.rcwa(freq=freq, order=order, L=L, dtype=dtype, device=torch.device("cuda:0"), **kw))
    a = C.run_case(mk(avoid_Pinv_instability=True), case, torch.complex128)
    b = C.run_case(mk(), case, torch.complex128)
    assert len(a.Pinv_instability) == 1 and len(a.Qinv_instability) == 1
    assert 0.0 <= float(a.Pinv_instability[0]) < 1e-8 and 0.0 <= float(a.Qinv_instability[0]) < 1e-8
    assert b.Pinv_instability is None
    assert np.abs(C.probe(a) - C.probe(b)).max() == 0.0
    rng = np.random.default_rng(5)
    lam = np.sort(rng.uniform(300.0, 900.0, 60))
    n, k = 3.5 + 0.8 * np.sin(lam / 90.0), 0.3 * np.exp(-(lam - 300.0) / 150.0)
    tab = NKTable(lam, n, k, device="cuda:0")
    q = torch.linspace(float(lam[0]), float(lam[-1]), 512, dtype=torch.float64, device="cuda:0")
    got = tab.apply(q)
    assert got.is_cuda
    ref = interp1d(lam, n, kind="cubic")(q.cpu().numpy()) + 1j * interp1d(lam, k, kind="cubic")(q.cpu().numpy())
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-12


def test_batched_equals_unbatched():
    """New API surface: a [B] frequency batch with per-point grids gives, entry by
    entry, what B separate sims give."""
    import torcwa_b200
    case = C.CASES["ex1_o3"]
    cd = torch.complex128
    lams = torch.tensor([500.0, 532.0, 610.0], dtype=torch.float64)
    d, grid = C.build_layers(case, cd)[0]
    grids = torch.stack([grid, grid * 0.9 + 0.1, grid]).to("cuda:0")
    sim = torcwa_b200.rcwa(freq=1 / lams, order=case["order"], L=case["L"], dtype=cd, device=torch.device("cuda:0"))
    sim.add_input_layer(eps=case["eps_in"])
    sim.set_incident_angle(inc_ang=0.1, azi_ang=0.2)
    sim.add_layer(thickness=torch.tensor([300.0, 250.0, 300.0]), eps=grids)
    sim.add_layer(thickness=50.0, eps=2.25)
    sim.solve_global_smatrix()
    tb = sim.S_parameters(orders=[[0, 0], [1, 0]], polarization="xx")
    rb = sim.S_parameters(orders=[[0, 0], [1, 0]], polarization="pp", port="reflection")
    assert tb.shape == (3, 2)
    for b in range(3):
        one = torcwa_b200.rcwa(freq=1 / lams[b], order=case["order"], L=case["L"], dtype=cd, device=torch.device("cuda:0"))
        one.add_input_layer(eps=case["eps_in"])
        one.set_incident_angle(inc_ang=0.1, azi_ang=0.2)
        one.add_layer(thickness=[300.0, 250.0, 300.0][b], eps=grids[b])
        one.add_layer(thickness=50.0, eps=2.25)
        one.solve_global_smatrix()
        t1 = one.S_parameters(orders=[[0, 0], [1, 0]], polarization="xx")
        r1 = one.S_parameters(orders=[[0, 0], [1, 0]], polarization="pp", port="reflection")
        assert t1.shape == (2,)
        assert float((tb[b] - t1).abs().max()) < 1e-11
        assert float((rb[b] - r1).abs().max()) < 1e-11


def test_dtype_and_device_rules():
    import torcwa_b200
    with pytest.raises(RuntimeError):
        torcwa_b200.rcwa(freq=1 / 532.0, order=[1, 1], L=[300.0, 300.0], device=torch.device("cpu"))
    sim = torcwa_b200.rcwa(freq=1 / 532.0, order=[1, 1], L=[300.0, 300.0], dtype=torch.complex128, device=torch.device("cuda:0"))
    sim.set_incident_angle(0.0, 0.0)
    with pytest.raises(RuntimeError):     # complex64 material in a complex128 sim (SURVEY.md finding 9)
        sim.add_layer(10.0, torch.ones(16, 16, dtype=torch.complex64, device="cuda:0"))
    with pytest.raises(AttributeError):   # python int has no .dim() -- same failure mode as the reference
        sim.add_layer(10.0, 2)
    with pytest.warns(UserWarning):
        torcwa_b200.rcwa(freq=1 / 532.0, order=[1, 1], L=[300.0, 300.0], dtype=torch.float32, device=torch.device("cuda:0"))


@pytest.mark.parametrize("sym", [False, True])
def test_order15_batched_sweep_parity_and_batch_independence(sym, golden_dir):
    """BASELINE config 2 at its real size (order 15x15, n = 1922), batch 24: the QR phase runs as two pipelined
    matrix groups with time-sliced passes and side-stream updates.  (a) the 532 nm point matches the reference's
    complex128 run to 1e-10; (b) every point is BIT-IDENTICAL to what a different batch composition gives -- the
    per-matrix arithmetic must not depend on how launches interleave (a cross-stream ordering race once broke this
    only at large batch)."""
    import torcwa_b200
    g = np.load(os.path.join(golden_dir, "ex1_o15.npz"))
    case = dict(C.CASES["ex1_o15"])
    cd = torch.complex128
    dev = torch.device("cuda:0")
    d, grid = C.build_layers(case, cd)[0]
    lams = torch.linspace(500.0, 640.0, 24, dtype=torch.float64)
    lams[0] = 532.0

    def solve(lam_vec, pipeline=1):
        sim = torcwa_b200.rcwa(freq=1 / lam_vec, order=case["order"], L=case["L"], dtype=cd, device=dev, store_intermediates=False, pipeline=pipeline,
                               symmetry_reduction=sym)
        sim.add_input_layer(eps=case["eps_in"])
        sim.set_incident_angle(0.0, 0.0)
        sim.add_layer(thickness=float(d), eps=grid.to(dev))
        sim.solve_global_smatrix()
        assert int(sim.eig_info[0].abs().max()) == 0
        return sim

    sim = solve(lams)
    one = torcwa_b200.rcwa(freq=1 / lams[0], order=case["order"], L=case["L"], dtype=cd, device=dev)
    one.add_input_layer(eps=case["eps_in"]); one.set_incident_angle(0.0, 0.0)
    one._S = [s[0:1].clone() for s in sim._S]
    sp = C.probe(one)
    scale = np.abs(g["sparams_c128"]).max()
    err = np.abs(sp - g["sparams_c128"]).max() / scale
    print("order-15 batched S-parameter max err / max|S|:", err)
    assert err <= 1e-10
    t_all = sim.S_parameters(orders=[[0, 0], [1, 0], [0, -1]], polarization="xx")
    sub = solve(lams[:5].clone())                                  # one group, other slice boundaries
    t_sub = sub.S_parameters(orders=[[0, 0], [1, 0], [0, -1]], polarization="xx")
    assert torch.equal(t_all[:5], t_sub)
    del sim, sub
    # the same sweep as two pipelined sub-batches (own streams, own host threads, staggered eigensolvers)
    piped = solve(lams, pipeline=2)
    assert piped._children is not None and len(piped._children) == 2
    assert int(piped.eig_info[0].abs().max()) == 0
    assert torch.equal(piped.S_parameters(orders=[[0, 0], [1, 0], [0, -1]], polarization="xx"), t_all)


def test_order15_energy_conservation_lossless_cell():
    """Size-independent property at the full BASELINE size (order 15x15, n = 1922): for a lossless cell the power
    carried by all propagating transmitted and reflected orders equals the incident power.  Batch of 3 wavelengths,
    complex64 API (complex128 arithmetic): |1 - (R + T)| <= 1e-6."""
    import torcwa_b200
    dev = torch.device("cuda:0")
    case = dict(C.CASES["ex1_o15"])
    rd = torch.float32
    mask = C.rectangle_grid(300.0, 300.0, 300, 300, 180.0, 100.0, 150.0, 150.0, 0.0, 1000.0, rd).to(dev)
    lams = torch.tensor([450.0, 532.0, 640.0], dtype=rd)
    sim = torcwa_b200.rcwa(freq=1 / lams, order=case["order"], L=case["L"], dtype=torch.complex64, device=dev)
    sim.add_input_layer(eps=1.46 ** 2)
    sim.set_incident_angle(0.0, 0.0)
    sim.add_layer(thickness=300.0, eps=mask * 6.25 + (1.0 - mask))          # real permittivity: no absorption
    sim.solve_global_smatrix()
    o = case["order"][0]
    orders = [[i, j] for i in range(-2, 3) for j in range(-2, 3)]           # every order that can propagate at these wavelengths
    tot = torch.zeros(3, dtype=torch.float64, device=dev)
    for port in ("transmission", "reflection"):
        for pol in ("xx", "yx"):
            s = sim.S_parameters(orders=orders, direction="forward", port=port, polarization=pol, ref_order=[0, 0], power_norm=True, evanscent=1e-3)
            tot += (s.abs().to(torch.float64) ** 2).sum(dim=1)
    print("R + T per wavelength:", tot.tolist())
    assert float((tot - 1.0).abs().max()) <= 1e-6


def test_differentiable_pipeline_gradients_vs_reference_autograd(golden_dir):
    """BASELINE config 5's shape of work (forward + autograd backward through the whole layer pipeline) on the CUDA
    kernels: figure of merit and its gradients w.r.t. the density grid and the thickness == the unmodified
    reference's CPU autograd (tests/golden/autograd_o3.npz).  Also as a batch of two identical points."""
    import torcwa_b200
    from oracle.autograd_case import CASE, fom
    g = np.load(os.path.join(golden_dir, "autograd_o3.npz"))
    dev = torch.device("cuda:0")
    rho = torch.from_numpy(g["rho"]).to(dev).requires_grad_(True)
    thick = torch.tensor(CASE["thickness"], dtype=torch.float64, device=dev, requires_grad=True)
    sim = torcwa_b200.rcwa(freq=torch.tensor(1.0 / CASE["lam"], dtype=torch.float64), order=CASE["order"], L=CASE["L"],
                           dtype=torch.complex128, device=dev)
    value = fom(sim, rho, thick)
    value.backward()
    assert abs(float(value.detach()) - float(g["fom"])) <= 1e-10 * abs(float(g["fom"]))
    gr = rho.grad.cpu().numpy()
    print("grad_rho rel err:", np.linalg.norm(gr - g["grad_rho"]) / np.linalg.norm(g["grad_rho"]))
    assert np.linalg.norm(gr - g["grad_rho"]) <= 1e-8 * np.linalg.norm(g["grad_rho"])
    assert abs(float(thick.grad) - float(g["grad_thickness"])) <= 1e-8 * abs(float(g["grad_thickness"]))
    # batched: two design points, each with its own density -> per-point gradients
    rho2 = torch.from_numpy(g["rho"]).to(dev)[None].repeat(2, 1, 1).requires_grad_(True)
    sim2 = torcwa_b200.rcwa(freq=torch.full((2,), 1.0 / CASE["lam"], dtype=torch.float64), order=CASE["order"], L=CASE["L"],
                            dtype=torch.complex128, device=dev)
    v2 = fom(sim2, rho2, torch.tensor(CASE["thickness"], dtype=torch.float64, device=dev))
    v2.backward()
    assert abs(float(v2.detach()) - 2 * float(g["fom"])) <= 1e-9 * abs(float(g["fom"]))
    for b in range(2):
        assert np.linalg.norm(rho2.grad[b].cpu().numpy() - g["grad_rho"]) <= 1e-8 * np.linalg.norm(g["grad_rho"])


def test_sources_and_fields_vs_reference(golden_dir):
    """Field reconstruction (scope row f1) on the CUDA kernels: lazily built mode coefficients + field_xz / field_yz /
    field_xy in the half spaces, patterned and homogeneous layers, forward xy and backward ps sources == the unmodified
    reference (tests/golden/fields_stack_o3.npz).  Fields do not depend on the eigenvector basis, so this also checks
    that our eigenvectors (different order / normalisation than LAPACK's) are used consistently."""
    import torcwa_b200
    from oracle.fields_case import build, SOURCES, planes
    g = np.load(os.path.join(golden_dir, "fields_stack_o3.npz"))
    sim = build(lambda **kw: torcwa_b200.rcwa(device=torch.device("cuda:0"), **kw))
    worst = 0.0
    for sname, setter in SOURCES.items():
        setter(sim)
        for pname, getter in planes().items():
            E, H = getter(sim)
            got = np.stack([t.cpu().numpy() for t in E + H])
            ref = g["%s_%s" % (sname, pname)]
            err = np.abs(got - ref).max() / np.abs(ref).max()
            worst = max(worst, err)
            assert err <= 1e-8, (sname, pname, err)
    print("fields: worst relative error vs reference", worst)
