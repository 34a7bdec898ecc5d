"""CPU: the C-ABI library builds, loads, and exports every symbol include/fi_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    from field_interpolation_b200 import build as fib
    return fib.build()


def test_header_symbols_exported(lib_path):
    header = open(os.path.join(ROOT, "include", "fi_b200.h")).read()
    declared = set(re.findall(r"^FI_API\s+[\w\s\*]+?\b(fi_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 24
    dll = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(dll, name), f"{name} declared in fi_b200.h but not exported"
    from field_interpolation_b200 import _lib
    assert set(_lib.SIGNATURES) == declared


def test_struct_layouts_match_reference_types():
    from field_interpolation_b200 import _lib
    assert ctypes.sizeof(_lib.fi_triplet) == 12          # Triplet, sparse_linear.hpp:8-15
    assert ctypes.sizeof(_lib.fi_weights) == 8 * 4 + 2 * 4  # Weights, field_interpolation.hpp:75-95


def test_no_silent_cpu_fallback(lib_path):
    """Without a GPU every compute entry point must fail loudly (FI_ERR_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import field_interpolation_b200 as fi
    with pytest.raises(fi.FiError) as e:
        fi.LatticeField([8, 8])
    assert e.value.code == 2


def test_defaults_match_reference(lib_path):
    from field_interpolation_b200 import _lib
    w = _lib.fi_weights()
    _lib.lib().fi_weights_default(ctypes.byref(w))
    assert (w.data_pos, w.data_gradient, w.model_0, w.model_1, w.model_2, w.model_3, w.model_4,
            w.gradient_smoothness, w.value_kernel, w.gradient_kernel) == (1.0, 1.0, 0.0, 0.0, 0.5, 0.0, 0.0, 0.0, 1, 1)
    o = _lib.fi_solve_options()
    _lib.lib().fi_solve_options_default(ctypes.byref(o))
    assert abs(o.tolerance - 1e-3) < 1e-15 and o.max_iterations == 0  # SolveOptions, sparse_linear.hpp:66-73
