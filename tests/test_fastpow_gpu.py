"""pow_pos() (csrc/fastpow.cuh), the inlined pow of the RHS kernels, against
libdevice pow(): on the fast path bitwise in the bit-for-bit build (RHS_RELAX=0) and, in the
default build (plain-double u^3 c(u^2) term of the logarithm, PB_RELAX & 8), equal in all but a
few results per thousand and never more than two representable numbers apart (libdevice's own
bound against the exact power is 2 ulp; the RHS contract is 1e-12); identical through the fallback."""
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


def test_domain_of_the_rhs_two_ulp():
    rng = np.random.default_rng(0)
    n = 1 << 20
    # saturations, 1 - s^m, (1/s)^m - 1, ponding depths; van Genuchten / Manning exponents
    x = np.concatenate([rng.uniform(0.1, 1.0, n // 4), 10.0 ** rng.uniform(-12, 0, n // 4),
                        rng.uniform(1.0, 1.0e5, n // 4), 10.0 ** rng.uniform(-8, 1, n // 4)])
    y = np.concatenate([rng.uniform(1.0, 10.0, n // 4), rng.uniform(0.05, 1.0, n // 4),
                        rng.uniform(0.1, 1.0, n // 4), np.full(n // 4, 0.6666667)])
    fast, ref = run(x, y)
    differ = fast != ref
    # distance in representable numbers (positive finite results: the bit patterns are ordered), so a
    # pair that straddles a power of two counts as the one step it is
    ulp = np.abs(fast.view(np.int64) - ref.view(np.int64)).astype(np.float64)
    print(f"pow_pos vs libdevice pow: {differ.sum()} of {n} differ ({differ.mean():.2e}), max {ulp.max():.0f} ulp")
    assert ulp.max() <= 2.0, f"max {ulp.max()} ulp, max rel {np.abs(fast / ref - 1).max():.2e}"
    assert differ.mean() <= 5e-3, f"{differ.sum()} of {n} differ"


def test_edge_cases_take_the_fallback():
    x = np.array([0.0, 1.0, 1.0, 5e-324, 1e-310, np.inf, 2.0, 0.5, 1e300, 1e-300, np.nan, 3.0])
    y = np.array([0.5, 3.3, 0.0, 0.5, 2.0, 0.5, 2000.0, 2000.0, 2.0, 2.0, 1.0, np.nan])
    fast, ref = run(x, y)
    assert np.array_equal(fast, ref, equal_nan=True), (fast, ref)
    assert fast[0] == 0.0 and fast[1] == 1.0 and fast[2] == 1.0


# ---- Arith<true> division (csrc/fdiv.cuh) against the hardware `/` ------------------------

def run_div(a, b):
    L = lib.load_library()
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    fast = np.empty_like(a); ok = np.empty_like(a); ref = np.empty_like(a)
    assert L.pihm_b200_test_div(len(a), a.ctypes.data, b.ctypes.data, fast.ctypes.data,
                                ok.ctypes.data, ref.ctypes.data) == 0
    return fast, ok, ref


def test_division_bitwise_on_the_rhs_domain():
    rng = np.random.default_rng(1)
    n = 1 << 21
    sgn = rng.choice([-1.0, 1.0], n)
    # fluxes / head differences over areas, distances, porosities, conductivities ...
    a = sgn * 10.0 ** rng.uniform(-30, 12, n)
    a[rng.random(n) < 0.1] = 0.0
    a[rng.random(n) < 0.02] = -0.0
    b = 10.0 ** rng.uniform(-12, 8, n)
    b[: n // 8] = rng.uniform(0.5, 2.0, n // 8)               # mantissa sweep near 1
    b[n // 8: n // 4] *= -1.0                                 # quo() takes either sign
    fast, ok, ref = run_div(a, b)
    assert (ok >= 0).all(), "shared-reciprocal form disagrees with `/`"
    assert (ok == 1).all(), f"{(ok != 1).sum()} operand pairs left the branch-free domain"
    # bitwise the hardware quotient; a zero numerator gives a zero (of either sign, fdiv.cuh)
    same = (fast == ref) & ((np.signbit(fast) == np.signbit(ref)) | (a == 0.0))
    assert same.all(), f"{(~same).sum()} of {n} quotients differ from the hardware division"
    # and the hardware division is IEEE: identical to the host's
    with np.errstate(all="ignore"):
        assert np.array_equal(ref, a / b)


def test_division_flags_everything_outside_its_domain():
    # divisor zero / denormal / huge / inf / nan, numerator inf / nan, overflowing quotient
    a = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 3.0, np.inf, np.nan, 1e300, 1e308])
    b = np.array([0.0, 5e-324, 1e-310, 1e308, np.inf, 0.0, np.inf, np.nan, np.nan, 2.0, 2.0, 1e-10, 1e-10])
    fast, ok, ref = run_div(a, b)
    assert (ok >= 0).all()
    assert not (ok == 1).any(), ok            # every one of these elements is recomputed with `/`
    # inside the divisor's domain numerators below 2^-969 are not flagged: within one ulp
    a = np.array([1e-300, 5e-324, 1e-200, 3e-295, -2e-300, 1e-300, 1e-310, 7e-292])
    b = np.array([3.0, 3.0, 1e200, 7e-296, 3e-302, 1e-5, 1e-300, 1e-290])
    fast, ok, ref = run_div(a, b)
    assert (ok == 1).all(), ok
    assert (np.abs(fast - ref) <= np.spacing(np.abs(ref))).all(), (fast, ref)
