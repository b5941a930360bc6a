"""Named parity cases (inputs only) -- TEST INFRASTRUCTURE, shared by tools/make_golden.py (which
runs the unmodified reference on them), tests/ and bench.py.  No reference code is needed to
rebuild the inputs: geometry comes from oracle.rcwa_oracle.rectangle_grid (asserted bit-identical
to torcwa.rcwa_geo.rectangle by the golden generator) and the a-Si:H permittivities are the
constants below (cubic interpolation of example/Materials_data/aSiH.txt, example/Materials.py:9-28,
re-derived and asserted by the generator)."""
import math

import torch

from .rcwa_oracle import rectangle_grid

SI_EPS = {532.0: complex(12.011610263133004, 0.5259120147560001),
          650.0: complex(10.362267239174999, 0.15362360819199997),
          400.0: complex(16.24464604339499, 3.9697033465479983),
          550.0: complex(11.646077602684997, 0.41213333041199984),
          700.0: complex(9.985966439994998, 0.11010441325199999)}
SU8 = 1.6 ** 2

PROBE_ORDERS = [[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1], [1, 1]]
PROBE_POLS = ["xx", "yx", "xy", "yy", "pp", "sp", "ps", "ss"]
PROBE_PORTS = [("forward", "transmission"), ("forward", "reflection"),
               ("backward", "transmission"), ("backward", "reflection")]


def _rect(d=300.0, Wx=180.0, Wy=100.0, theta=0.0, eps_in=SI_EPS[532.0], eps_bg=1.0):
    return dict(kind="rect", d=d, Wx=Wx, Wy=Wy, Cx=150.0, Cy=150.0, theta=theta, eps_in=eps_in, eps_bg=eps_bg)


def _base(**kw):
    c = dict(L=[300.0, 300.0], nxy=[300, 300], edge_sharpness=1000.0, lam=532.0, eps_in=1.46 ** 2,
             eps_out=None, inc=0.0, azi=0.0, full=False, big=False, c128_only=False)
    c.update(kw)
    return c


def _stack():
    out = []
    for k, th in enumerate([0.0, math.pi / 6, math.pi / 3, math.pi / 2]):
        out.append(_rect(d=200.0, theta=th, eps_in=SI_EPS[650.0], eps_bg=SU8))
        out.append(dict(kind="homogeneous", d=100.0, eps=SU8 if k != 1 else complex(2.0, 0.3)))
    return out


def _stack_config3():
    """BASELINE.json configs[2] (Example1-1.ipynb:58-69,159-177): four a-Si:H bars in SU-8, rotated by 0 / 30 / 60 / 90
    degrees, 200 nm each, separated by 100 nm homogeneous SU-8 spacers: 8 layers."""
    out = []
    for th in (0.0, math.pi / 6, math.pi / 3, math.pi / 2):
        out.append(_rect(d=200.0, theta=th, eps_in=SI_EPS[650.0], eps_bg=SU8))
        out.append(dict(kind="homogeneous", d=100.0, eps=SU8))
    return out


CASES = {
    # BASELINE.json configs[0] (Example1.ipynb:40-57,164-175) and siblings
    "ex1_o3": _base(order=[3, 3], layers=[_rect()], full=True),
    "ex1_o5": _base(order=[5, 5], layers=[_rect()]),
    # BASELINE.json configs[1] unit (one wavelength of the 512 batch)
    "ex1_o15": _base(order=[15, 15], layers=[_rect()], big=True),
    # Example1-1 style stack, oblique incidence, lossy homogeneous layer, output half space
    "stack_o3": _base(order=[3, 3], layers=_stack(), lam=650.0, eps_out=2.1, inc=0.3, azi=0.4, full=True),
    "stack_o4x2": _base(order=[4, 2], layers=_stack()[:3], lam=650.0, eps_in=None, eps_out=1.46 ** 2,
                        inc=0.2, azi=-0.7, nxy=[64, 48], full=True),
    # zero-layer Fresnel interface (Example0.ipynb:59-76)
    "fresnel_o2": _base(order=[2, 2], layers=[], eps_out=1.0, inc=0.5, full=True),
    # BASELINE.json configs[3] (Example3.ipynb:85-101): corners and centre of the (Wx, Wy, lambda) sweep at order 15
    "sweep_o15_a": _base(order=[15, 15], layers=[_rect(Wx=50.0, Wy=250.0, eps_in=SI_EPS[400.0])], lam=400.0, big=True),
    "sweep_o15_b": _base(order=[15, 15], layers=[_rect(Wx=250.0, Wy=50.0, eps_in=SI_EPS[700.0])], lam=700.0, big=True),
    "sweep_o15_c": _base(order=[15, 15], layers=[_rect(Wx=150.0, Wy=150.0, eps_in=SI_EPS[550.0])], lam=550.0, big=True),
    # BASELINE.json configs[2]: order 21 x 21, 8 stacked layers, complex128 (the reference's complex64 run is skipped:
    # its MKL cgetri slow path would take hours at 4N = 7396)
    "stack_o21": _base(order=[21, 21], layers=_stack_config3(), lam=650.0, big=True, c128_only=True),
    # symmetry-reduction cases (torcwa_b200/symmetry.py): inversion centre only (rotated bars + spacer + output half space);
    # one mirror only (incidence in the x-z plane keeps the y mirror); off-centre mirror planes
    "c2_o3": _base(order=[3, 3], layers=[_rect(theta=math.pi / 6, eps_in=SI_EPS[650.0], eps_bg=SU8, d=200.0),
                                         dict(kind="homogeneous", d=100.0, eps=SU8),
                                         _rect(theta=math.pi / 3, eps_in=SI_EPS[650.0], eps_bg=SU8, d=150.0)], lam=650.0, eps_out=2.1, full=True),
    "ymirror_o3": _base(order=[3, 2], layers=[_rect()], inc=0.35, azi=0.0, nxy=[96, 80], full=True),
    # the other single mirror: incidence in the y-z plane keeps the x mirror; spacer layer and output half space on top
    "xmirror_o3": _base(order=[2, 3], layers=[_rect(), dict(kind="homogeneous", d=80.0, eps=complex(2.0, 0.2)), _rect(Wx=120.0, Wy=200.0, d=150.0)],
                        inc=0.3, azi=math.pi / 2, nxy=[80, 96], eps_out=2.1, full=True),
    "offcentre_o3": _base(order=[3, 3], layers=[dict(kind="rect", d=250.0, Wx=140.0, Wy=90.0, Cx=101.0, Cy=187.5, theta=0.0,
                                                     eps_in=SI_EPS[532.0], eps_bg=1.0)], full=True),
    # C4v-symmetric cell: exactly degenerate eigenpairs
    "square_o4": _base(order=[4, 4], layers=[_rect(Wx=150.0, Wy=150.0)]),
}


def real_dtype(cdtype):
    return torch.float32 if cdtype == torch.complex64 else torch.float64


def build_layers(case, cdtype):
    """-> [(thickness, eps)], eps a python scalar (homogeneous) or an [nx,ny] tensor of cdtype."""
    rd = real_dtype(cdtype)
    layers = []
    for lay in case["layers"]:
        if lay["kind"] == "homogeneous":
            layers.append((lay["d"], lay["eps"]))
            continue
        mask = rectangle_grid(case["L"][0], case["L"][1], case["nxy"][0], case["nxy"][1], lay["Wx"], lay["Wy"],
                              lay["Cx"], lay["Cy"], lay["theta"], case["edge_sharpness"], rd)
        e_in = torch.as_tensor(lay["eps_in"], dtype=cdtype)
        e_bg = torch.as_tensor(lay["eps_bg"], dtype=cdtype)
        layers.append((lay["d"], mask * e_in + (1.0 - mask) * e_bg))
    return layers


def freq_of(case, cdtype):
    """The reference examples pass freq = 1/lam with lam a real tensor of the sim's precision."""
    return 1 / torch.tensor(case["lam"], dtype=real_dtype(cdtype))


def run_case(sim_factory, case, cdtype):
    """Drive any solver object exposing the reference's call sequence."""
    sim = sim_factory(freq=freq_of(case, cdtype), order=case["order"], L=case["L"], dtype=cdtype)
    if case["eps_in"] is not None:
        sim.add_input_layer(eps=case["eps_in"])
    if case["eps_out"] is not None:
        sim.add_output_layer(eps=case["eps_out"])
    sim.set_incident_angle(inc_ang=case["inc"], azi_ang=case["azi"])
    for d, e in build_layers(case, cdtype):
        sim.add_layer(thickness=d, eps=e)
    sim.solve_global_smatrix()
    return sim


def probe(sim):
    """[ports, pols, orders] complex128 numpy array of S-parameters."""
    import numpy as np
    rows = []
    for d, p in PROBE_PORTS:
        for pol in PROBE_POLS:
            v = sim.S_parameters(orders=[list(o) for o in PROBE_ORDERS], direction=d, port=p,
                                 polarization=pol, ref_order=[0, 0])
            rows.append(v.detach().cpu().to(torch.complex128).numpy())
    return np.stack(rows).reshape(len(PROBE_PORTS), len(PROBE_POLS), len(PROBE_ORDERS))
