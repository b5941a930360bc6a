"""The oracle (oracle/rcwa_oracle.py) against the reference's stored outputs (tests/golden/*.npz,
written by tools/make_golden.py from the unmodified reference) and against the soft pins the
reference's notebooks hold."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import cases as C
from oracle.rcwa_oracle import OracleSim, PI_REF

SMALL = ["ex1_o3", "ex1_o5", "stack_o3", "stack_o4x2", "fresnel_o2", "square_o4", "c2_o3", "ymirror_o3", "xmirror_o3", "offcentre_o3"]


def oracle_factory(freq, order, L, dtype):
    return OracleSim(freq=freq, order=order, L=L, dtype=dtype)


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_c128(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sim = C.run_case(oracle_factory, C.CASES[name], torch.complex128)
    sp = C.probe(sim)
    scale = np.abs(g["sparams_c128"]).max()
    assert np.abs(sp - g["sparams_c128"]).max() <= 1e-11 * scale
    fro = np.array([float(torch.linalg.norm(s)) for s in sim.S])
    np.testing.assert_allclose(fro, g["S_fro"], rtol=1e-11)
    cols = g["S_cols_idx"]
    for k in range(4):
        assert relerr(sim.S[k][:, cols].numpy(), g["S_cols"][k]) <= 1e-11
    if "S" in g:
        for k in range(4):
            assert relerr(sim.S[k].numpy(), g["S"][k]) <= 1e-11
    if "eps_conv0" in g:
        assert relerr(sim.eps_conv[0].numpy(), g["eps_conv0"]) <= 1e-14
        for k in range(4):
            assert relerr(sim.layer_S[0][k].numpy(), g["layer_S0"][k]) <= 1e-11
    if "kz2_sorted" in g:
        for l, kz in enumerate(sim.kz_norm):
            mine = np.sort_complex(kz.numpy() ** 2)
            assert np.abs(mine - g["kz2_sorted"][l]).max() <= 1e-9 * np.abs(mine).max()


@pytest.mark.parametrize("name", ["ex1_o3", "stack_o4x2"])
def test_oracle_matches_reference_c64(name, golden_dir):
    """complex64: same LAPACK calls in the same order -> agreement far below the c64-vs-c128 gap."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sim = C.run_case(oracle_factory, C.CASES[name], torch.complex64)
    sp = C.probe(sim)
    assert np.abs(sp - g["sparams_c64"]).max() <= 2e-4 * np.abs(g["sparams_c64"]).max()


def test_pi_is_the_references_constant():
    assert PI_REF != math.pi and abs(PI_REF / math.pi - 1) < 4e-10     # torcwa/rcwa.py:5


def test_example1_order5_known_answer(golden_dir):
    """Example1 @ order 5 -> txx = -0.6182-0.0593j."""
    g = np.load(os.path.join(golden_dir, "ex1_o5.npz"))
    assert abs(complex(g["sparams_c128"][0, 0, 0]) - complex(-0.6182, -0.0593)) < 1e-4


def test_fresnel_interface_example0():
    """Example0.ipynb:59-76 (sim) vs :94-97 (analytic Fresnel): n1=1.46 -> n2=1, incl. TIR."""
    n1, n2 = 1.46, 1.0
    for deg in (0.0, 20.0, 40.0, 60.0):
        th = math.radians(deg)
        sim = OracleSim(freq=1 / 532.0, order=[1, 1], L=[300.0, 300.0], dtype=torch.complex128)
        sim.add_input_layer(eps=n1 ** 2)
        sim.add_output_layer(eps=n2 ** 2)
        sim.set_incident_angle(th, 0.0)
        sim.solve_global_smatrix()
        rpp = sim.S_parameters([0, 0], direction="forward", port="reflection", polarization="pp")
        rss = sim.S_parameters([0, 0], direction="forward", port="reflection", polarization="ss")
        s = n1 * math.sin(th) / n2
        if s < 1:
            ct = math.sqrt(1 - s * s)
            r_te = ((n1 * math.cos(th) - n2 * ct) / (n1 * math.cos(th) + n2 * ct)) ** 2
            r_tm = ((n2 * math.cos(th) - n1 * ct) / (n2 * math.cos(th) + n1 * ct)) ** 2
        else:
            r_te = r_tm = 1.0
        assert abs(float(abs(rss[0]) ** 2) - r_te) < 1e-9
        assert abs(float(abs(rpp[0]) ** 2) - r_tm) < 1e-9


def test_energy_conservation_lossless():
    """Physics invariant: sum T + sum R = 1 for lossless eps."""
    case = dict(C.CASES["ex1_o3"])
    lay = dict(case["layers"][0]); lay["eps_in"] = 6.25
    case["layers"] = [lay]
    sim = C.run_case(oracle_factory, case, torch.complex128)
    allo = [[i, j] for i in range(-3, 4) for j in range(-3, 4)]
    tot = 0.0
    for port in ("transmission", "reflection"):
        for pol in ("xx", "yx"):
            v = sim.S_parameters(allo, direction="forward", port=port, polarization=pol)
            tot += float((v.abs() ** 2).sum())
    assert abs(tot - 1.0) < 1e-10


@pytest.mark.parametrize("name", ["rand6", "rand24", "rand57", "rcwa_o3"])
@pytest.mark.parametrize("broadening", [1e-10, None])
def test_oracle_eig_backward_matches_reference(name, broadening, golden_dir):
    """oracle.eig_backward == the unmodified reference's Eig.backward (tools/make_golden_eig_backward.py)."""
    from oracle.rcwa_oracle import eig_backward
    g = np.load(os.path.join(golden_dir, "eig_backward.npz"))
    t = lambda k: torch.from_numpy(g[name + "_" + k])
    ref = g[name + "_grad_b" + ("1e-10" if broadening is not None else "None")]
    got = eig_backward(t("w"), t("V"), t("gw"), t("gV"), broadening).numpy()
    assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)
