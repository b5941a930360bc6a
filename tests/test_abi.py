"""The C-ABI library loads without a GPU and exports every symbol include/rcwa_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rcwa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcwa_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_everything():
    from torcwa_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == names          # the Python binding covers exactly the header
    assert _lib.load().rcwa_b200_abi_version() == 2


def test_argument_checks_do_not_touch_the_gpu():
    """LAPACK-style negative return codes for bad arguments, before any CUDA call."""
    from torcwa_b200 import _lib
    lib = _lib.load()
    assert lib.rcwa_convmat(None, 0, 0, 16, 16, 1, 1, 1, None, None, None) == -1
    assert lib.rcwa_eig(None, 4, 1, None, None, None, 0, None, None, None) == -1
    assert lib.rcwa_lu_factor(None, 0, 4, 4, 1, None, None, None, None, None, None) == -1
    assert lib.rcwa_eig_workspace_bytes(1922, 1) > 2 * 1922 * 1922 * 16
    assert lib.rcwa_sym_project(None, 1, 8, None, None, None, None, 2, 4, 4, None, None) == -1
    one = ctypes.c_void_p(16)          # a non-null placeholder: the checks return before anything is dereferenced
    assert lib.rcwa_sym_project(one, 1, 8, one, one, one, one, 5, 4, 4, one, None) == -8      # G > 4
    assert lib.rcwa_sym_project(one, 0, 8, one, one, one, one, 2, 4, 4, one, None) == -2


def test_product_refuses_cpu_devices():
    import pytest
    import torch
    import torcwa_b200
    with pytest.raises(RuntimeError):
        torcwa_b200.rcwa(freq=1 / 532.0, order=[1, 1], L=[300.0, 300.0], device=torch.device("cpu"))
