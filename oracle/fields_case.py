"""Shared driver of the field-reconstruction parity case (TEST INFRASTRUCTURE): the same calls run against the
reference (tools/make_golden_fields.py), the CPU double and the CUDA path."""
import torch

from . import cases as C


def build(factory):
    case = dict(C.CASES["stack_o3"])
    case["layers"] = case["layers"][:3]              # patterned, lossy homogeneous, patterned (rotated)
    return C.run_case(factory, case, torch.complex128)


def _fwd_xy(sim):
    sim.source_planewave(amplitude=[1.0, 0.5j], direction="forward", notation="xy")


def _bwd_ps(sim):
    sim.source_fourier(amplitude=[[0.7, 0.2], [0.1, -0.4j]], orders=[[0, 0], [1, -1]], direction="backward", notation="ps")


SOURCES = {"fwd_xy": _fwd_xy, "bwd_ps": _bwd_ps}


def planes():
    x = torch.linspace(0.0, 300.0, 7, dtype=torch.float64)
    y = torch.linspace(-20.0, 280.0, 5, dtype=torch.float64)
    z = torch.tensor([-80.0, -1.0, 0.0, 60.0, 200.0, 230.0, 300.0, 301.0, 450.0, 500.0, 620.0], dtype=torch.float64)
    return {
        "xz": lambda s: s.field_xz(x, z, 40.0),
        "yz": lambda s: s.field_yz(y, z, 110.0),
        "xy_in": lambda s: s.field_xy(-1, x, y, -35.0),
        "xy_l0": lambda s: s.field_xy(0, x, y, 120.0),
        "xy_l1": lambda s: s.field_xy(1, x, y, 30.0),
        "xy_l2": lambda s: s.field_xy(2, x, y, 0.0),
        "xy_out": lambda s: s.field_xy(3, x, y, 75.0),
    }
