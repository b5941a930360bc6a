"""Shared definition of the gradient parity case (TEST INFRASTRUCTURE): a topology-optimisation style
figure of merit through the public API, written once so that the reference (tools/make_golden_autograd.py),
the CPU double and the CUDA path run literally the same driver.  Mirrors example/Example6.ipynb's use of the
API (density -> permittivity grid -> add_layer -> S_parameters -> |t|^2)."""
import torch

CASE = {"lam": 560.0, "order": [3, 2], "L": [600.0, 520.0], "nx": 36, "ny": 30, "thickness": 210.0,
        "eps_si": complex(12.011610263133004, 0.5259120147560001), "inc": 0.2, "azi": 0.1}


def density():
    """Smooth seeded density in (0, 1) on the nx x ny grid (float64)."""
    g = torch.Generator().manual_seed(333)
    r = torch.rand(CASE["nx"], CASE["ny"], generator=g, dtype=torch.float64)
    k = torch.fft.fft2(r)
    fx = torch.fft.fftfreq(CASE["nx"], dtype=torch.float64)[:, None]
    fy = torch.fft.fftfreq(CASE["ny"], dtype=torch.float64)[None, :]
    blur = torch.exp(-((fx * 6.0) ** 2 + (fy * 6.0) ** 2))
    s = torch.fft.ifft2(k * blur).real
    s = (s - s.min()) / (s.max() - s.min())
    return (0.1 + 0.8 * s).clone()


def fom(sim, rho, thick):
    """sum of |t_xx(0,0)|^2 + |t_yx(0,0)|^2 + |r_xx(1,0)|^2 for a (patterned + homogeneous) stack on glass."""
    dev = rho.device
    eps = rho.to(torch.complex128) * CASE["eps_si"] + (1.0 - rho)
    sim.add_input_layer(eps=1.46 ** 2)
    sim.set_incident_angle(inc_ang=CASE["inc"], azi_ang=CASE["azi"])
    sim.add_layer(thickness=thick, eps=eps)
    sim.add_layer(thickness=50.0, eps=2.25)
    sim.solve_global_smatrix()
    txx = sim.S_parameters(orders=[0, 0], direction="forward", port="transmission", polarization="xx", ref_order=[0, 0])
    tyx = sim.S_parameters(orders=[0, 0], direction="forward", port="transmission", polarization="yx", ref_order=[0, 0])
    rxx = sim.S_parameters(orders=[1, 0], direction="forward", port="reflection", polarization="xx", ref_order=[0, 0])
    return (txx.abs() ** 2).sum() + (tyx.abs() ** 2).sum() + (rxx.abs() ** 2).sum()
