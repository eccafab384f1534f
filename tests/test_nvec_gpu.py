"""Device N_Vector ops against nvector_serial.c (golden outputs of the
reference library) and against the oracle's serial arithmetic."""
import numpy as np
import pytest

import oraclelib
from helpers import load_golden
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    tb = W.make_watershed(20, 8, river=True, trib_every=5)     # nsv >= 1000
    m = lib.Model(tb)
    yield m
    m.close()


def _pad(model, a, fill=0.0):
    out = np.full(model.nsv, fill)
    out[:len(a)] = a
    return out


def test_streaming_ops_bit_exact(model):
    g = load_golden("nvec_serial.npz")
    n = len(g["x"])
    assert model.nsv >= n
    # reorder=0: device order == host order, so elementwise results compare 1:1
    x = model.N_VNew(_pad(model, g["x"])); y = model.N_VNew(_pad(model, g["y"]))
    w = model.N_VNew(_pad(model, g["w"], 1.0)); z = model.N_VNew()
    for (a, b), zref in zip(g["ls_coeffs"], g["ls_z"]):
        model.N_VLinearSum(a, x, b, y, z)
        assert np.array_equal(z.download()[:n], zref), (a, b)
        # in-place forms (Vaxpy paths, nvector_serial.c:431-439)
        t = model.N_VNew(_pad(model, g["y"]))
        model.N_VLinearSum(a, x, b, t, t)
        assert np.array_equal(t.download()[:n], oraclelib.nv_linearsum(a, g["x"], b, g["y"], inplace="y"))
        t.free()
    for c, zref in zip(g["sc_c"], g["sc_z"]):
        model.N_VScale(c, x, z)
        assert np.array_equal(z.download()[:n], zref)
    model.N_VProd(x, y, z); assert np.array_equal(z.download()[:n], g["prod"])
    model.N_VDiv(x, w, z); assert np.array_equal(z.download()[:n], g["div"])
    model.N_VAbs(x, z); assert np.array_equal(z.download()[:n], g["abs"])
    model.N_VInv(w, z); assert np.array_equal(z.download()[:n], g["inv"])
    model.N_VAddConst(x, 0.125, z); assert np.array_equal(z.download()[:n], g["addconst"])
    model.N_VConst(3.25, z); assert (z.download() == 3.25).all()
    for v in (x, y, w, z):
        v.free()


def test_reductions(model):
    """tree sums differ from the serial left-to-right sum in the last bits only"""
    g = load_golden("nvec_serial.npz")
    n = len(g["x"])
    x = model.N_VNew(_pad(model, g["x"])); y = model.N_VNew(_pad(model, g["y"]))
    w = model.N_VNew(_pad(model, g["w"], 1.0))
    sumabs = np.abs(g["x"] * g["y"]).sum()
    assert abs(model.N_VDotProd(x, y) - float(g["dot"])) <= 1e-14 * sumabs
    assert model.N_VMaxNorm(x) == float(g["maxnorm"])
    assert model.N_VMin(x) == min(float(g["min"]), 0.0)            # padding zeros
    # WRMS over nsv components (padding contributes 0 to the sum)
    ref = np.sqrt(((g["x"] * g["w"]) ** 2).sum() / model.nsv)
    assert abs(model.N_VWrmsNorm(x, w) - ref) <= 1e-14 * ref
    for v in (x, y, w):
        v.free()


@pytest.mark.parametrize("n_target", [1, 255, 256, 257, 100_003])
def test_reductions_ragged_lengths(n_target):
    """lengths around the block size and a large odd one; repeated calls are bitwise stable"""
    nx = max(1, int(np.ceil(n_target / 6)))
    tb = W.make_watershed(nx, 1, river=False)
    m = lib.Model(tb)
    rng = np.random.default_rng(n_target)
    a = rng.standard_normal(m.nsv); b = rng.uniform(0.5, 2.0, m.nsv)
    x = m.N_VNew(a); w = m.N_VNew(b)
    d1 = m.N_VDotProd(x, w)
    assert d1 == m.N_VDotProd(x, w)                      # deterministic
    assert abs(d1 - oraclelib.nv_dotprod(a, b)) <= 1e-13 * np.abs(a * b).sum()
    r = oraclelib.nv_wrmsnorm(a, b)
    assert abs(m.N_VWrmsNorm(x, w) - r) <= 1e-13 * r
    assert m.N_VMaxNorm(x) == oraclelib.nv_maxnorm(a)
    assert m.N_VMin(x) == oraclelib.nv_min(a)
    m.close()


def test_upload_download_roundtrip_with_reordering():
    tb = W.make_named("small", fbr=True)
    m = lib.Model(tb, reorder=1)
    perm = m.permutation()
    assert sorted(perm.tolist()) == list(range(tb["nelem"]))
    assert (perm != np.arange(tb["nelem"])).any()
    y = np.random.default_rng(0).standard_normal(m.nsv)
    v = m.N_VNew(y)
    assert np.array_equal(v.download(), y)
    m.close()
