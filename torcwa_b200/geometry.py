"""Sigmoid-edged shape rasterisers with the reference's names and signatures
(torcwa/geometry.py:4-290).  Input generation only -- outside the hot path (SURVEY.md section 2
row 10); plain torch elementwise code, differentiable, any device."""
import torch


def _frame(x_grid, y_grid, Cx, Cy, theta, dtype, device):
    th = torch.as_tensor(theta, dtype=dtype, device=device)
    dx, dy = x_grid - Cx, y_grid - Cy
    c, s = torch.cos(th), torch.sin(th)
    return dx * c + dy * s, -dx * s + dy * c


class _Shapes:
    """Shape methods shared by the instance flavour (`geometry`) and the class flavour (`rcwa_geo`)."""

    def _grid(self):
        self.grid()
        return self.x_grid, self.y_grid

    def _edge(self, level):
        return torch.sigmoid(self.edge_sharpness * level)

    def _uv(self, Cx, Cy, theta):
        X, Y = self._grid()
        return _frame(X, Y, Cx, Cy, theta, self.dtype, self.device)

    def circle(self, R, Cx, Cy):
        u, v = self._uv(Cx, Cy, 0.)
        return self._edge(1. - torch.sqrt((u / R) ** 2 + (v / R) ** 2))

    def ellipse(self, Rx, Ry, Cx, Cy, theta=0.):
        u, v = self._uv(Cx, Cy, theta)
        return self._edge(1. - torch.sqrt((u / Rx) ** 2 + (v / Ry) ** 2))

    def square(self, W, Cx, Cy, theta=0.):
        return self.rectangle(W, W, Cx, Cy, theta)

    def rectangle(self, Wx, Wy, Cx, Cy, theta=0.):
        u, v = self._uv(Cx, Cy, theta)
        return self._edge(1. - torch.maximum(torch.abs(u / (Wx / 2.)), torch.abs(v / (Wy / 2.))))

    def rhombus(self, Wx, Wy, Cx, Cy, theta=0.):
        u, v = self._uv(Cx, Cy, theta)
        return self._edge(1. - (torch.abs(u / (Wx / 2.)) + torch.abs(v / (Wy / 2.))))

    def super_ellipse(self, Wx, Wy, Cx, Cy, theta=0., power=2.):
        u, v = self._uv(Cx, Cy, theta)
        return self._edge(1. - (torch.abs(u / (Wx / 2.)) ** power + torch.abs(v / (Wy / 2.)) ** power) ** (1 / power))

    @staticmethod
    def _union(A, B):
        return torch.maximum(A, B)

    @staticmethod
    def _intersection(A, B):
        return torch.minimum(A, B)

    @staticmethod
    def _difference(A, B):
        return torch.minimum(A, 1. - B)


class geometry(_Shapes):
    def __init__(self, Lx: float = 1., Ly: float = 1., nx: int = 100, ny: int = 100, edge_sharpness: float = 1000., *,
                 dtype=torch.float32, device=torch.device('cuda' if torch.cuda.is_available() else 'cpu')):
        self.Lx, self.Ly, self.nx, self.ny = Lx, Ly, nx, ny
        self.edge_sharpness = edge_sharpness
        self.dtype, self.device = dtype, device

    def grid(self):
        self.x = (self.Lx / self.nx) * (torch.arange(self.nx, dtype=self.dtype, device=self.device) + 0.5)
        self.y = (self.Ly / self.ny) * (torch.arange(self.ny, dtype=self.dtype, device=self.device) + 0.5)
        self.x_grid, self.y_grid = torch.meshgrid(self.x, self.y, indexing='ij')

    def union(self, A, B):
        return self._union(A, B)

    def intersection(self, A, B):
        return self._intersection(A, B)

    def difference(self, A, B):
        return self._difference(A, B)


class _ClassProxy(_Shapes):
    """Binds the shape methods to class attributes so that `rcwa_geo.rectangle(...)` works without
    an instance, as in the reference (torcwa/geometry.py:155-290)."""

    def __init__(self, cls):
        object.__setattr__(self, '_cls', cls)

    def __getattr__(self, k):
        return getattr(self._cls, k)

    def __setattr__(self, k, v):
        setattr(self._cls, k, v)

    def grid(self):
        self._cls.grid()


class _GeoMeta(type):
    _shape_names = ('circle', 'ellipse', 'square', 'rectangle', 'rhombus', 'super_ellipse')

    def __getattr__(cls, name):
        if name in _GeoMeta._shape_names:
            return getattr(_ClassProxy(cls), name)
        raise AttributeError(name)


class rcwa_geo(metaclass=_GeoMeta):
    edge_sharpness = 1000.
    Lx, Ly = 1., 1.
    nx, ny = 100, 100
    dtype = torch.float32
    device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')

    @classmethod
    def grid(cls):
        cls.x = (cls.Lx / cls.nx) * (torch.arange(cls.nx, dtype=cls.dtype, device=cls.device) + 0.5)
        cls.y = (cls.Ly / cls.ny) * (torch.arange(cls.ny, dtype=cls.dtype, device=cls.device) + 0.5)
        cls.x_grid, cls.y_grid = torch.meshgrid(cls.x, cls.y, indexing='ij')

    union = staticmethod(_Shapes._union)
    intersection = staticmethod(_Shapes._intersection)
    difference = staticmethod(_Shapes._difference)
