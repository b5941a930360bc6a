"""torcwa_b200 -- B200-native RCWA inner loop with the torcwa API (drop-in for the hot path).

    import torcwa_b200 as torcwa
    sim = torcwa.rcwa(freq=1/532., order=[15, 15], L=[300., 300.], dtype=torch.complex64)

Same public names as the reference package (torcwa/__init__.py:1-6)."""
from .torch_eig import Eig
from .geometry import geometry, rcwa_geo
from .rcwa import rcwa

__version__ = '0.1.4.2+b200.r1'
