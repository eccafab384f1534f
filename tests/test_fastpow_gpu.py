"""pow_pos() (csrc/fastpow.cuh), the inlined pow of the RHS kernels, against
libdevice pow(): bitwise on the fast path, and identical through the fallback."""
import numpy as np
import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib

pytestmark = pytest.mark.gpu


def run(x, y):
    L = lib.load_library()
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    fast = np.empty_like(x); ref = np.empty_like(x)
    assert L.pihm_b200_test_pow(len(x), x.ctypes.data, y.ctypes.data, fast.ctypes.data, ref.ctypes.data) == 0
    return fast, ref


def test_domain_of_the_rhs_bitwise():
    rng = np.random.default_rng(0)
    n = 1 << 20
    # saturations, 1 - s^m, (1/s)^m - 1, ponding depths; van Genuchten / Manning exponents
    x = np.concatenate([rng.uniform(0.1, 1.0, n // 4), 10.0 ** rng.uniform(-12, 0, n // 4),
                        rng.uniform(1.0, 1.0e5, n // 4), 10.0 ** rng.uniform(-8, 1, n // 4)])
    y = np.concatenate([rng.uniform(1.0, 10.0, n // 4), rng.uniform(0.05, 1.0, n // 4),
                        rng.uniform(0.1, 1.0, n // 4), np.full(n // 4, 0.6666667)])
    fast, ref = run(x, y)
    assert np.array_equal(fast, ref), f"{(fast != ref).sum()} of {n} differ, max rel {np.abs(fast / ref - 1).max():.2e}"


def test_edge_cases_take_the_fallback():
    x = np.array([0.0, 1.0, 1.0, 5e-324, 1e-310, np.inf, 2.0, 0.5, 1e300, 1e-300, np.nan, 3.0])
    y = np.array([0.5, 3.3, 0.0, 0.5, 2.0, 0.5, 2000.0, 2000.0, 2.0, 2.0, 1.0, np.nan])
    fast, ref = run(x, y)
    assert np.array_equal(fast, ref, equal_nan=True), (fast, ref)
    assert fast[0] == 0.0 and fast[1] == 1.0 and fast[2] == 1.0
