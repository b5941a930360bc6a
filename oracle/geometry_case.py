"""Shared calls of the geometry parity case (TEST INFRASTRUCTURE)."""
import torch


def setup(cls):
    cls.dtype, cls.device = torch.float64, torch.device("cpu")
    cls.Lx, cls.Ly, cls.nx, cls.ny, cls.edge_sharpness = 320.0, 240.0, 24, 20, 35.0
    cls.grid()


def shapes():
    return {
        "circle": lambda g: g.circle(R=70.0, Cx=150.0, Cy=110.0),
        "ellipse": lambda g: g.ellipse(Rx=90.0, Ry=40.0, Cx=160.0, Cy=120.0, theta=0.4),
        "square": lambda g: g.square(W=100.0, Cx=140.0, Cy=100.0, theta=0.2),
        "rectangle": lambda g: g.rectangle(Wx=180.0, Wy=60.0, Cx=170.0, Cy=130.0, theta=-0.7),
        "rhombus": lambda g: g.rhombus(Wx=150.0, Wy=90.0, Cx=160.0, Cy=120.0, theta=0.3),
        "super_ellipse": lambda g: g.super_ellipse(Wx=160.0, Wy=100.0, Cx=160.0, Cy=120.0, theta=0.1, power=4.0),
        "union": lambda g: g.union(g.circle(R=50.0, Cx=100.0, Cy=100.0), g.square(W=80.0, Cx=200.0, Cy=140.0)),
        "intersection": lambda g: g.intersection(g.circle(R=80.0, Cx=150.0, Cy=110.0), g.rectangle(Wx=200.0, Wy=50.0, Cx=160.0, Cy=120.0)),
        "difference": lambda g: g.difference(g.circle(R=80.0, Cx=150.0, Cy=110.0), g.circle(R=40.0, Cx=150.0, Cy=110.0)),
    }
