"""The drop-in boundary: the UNMODIFIED reference CVODE
(oracle/_ref, compiled from /root/reference) integrates the watershed using
our device-resident N_Vector (N_VNew_PihmB200) and our RHS callback
(PihmB200_ODE) -- the SetCVodeParam/SolveCVode call sequence of src/ode.c with
two pointers swapped.

Because the fused kernels of our own integrator perform, per component, the
same operations in the same order and the reductions share one summation tree,
our integrator must reproduce this run BIT FOR BIT (state and counters): that
pins the restatement of CVODE/SPGMR in cvode_b200.cu to the reference's
control flow far more sharply than the tolerance-based trajectory tests."""
import ctypes as C

import numpy as np
import pytest

import reflib
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu


class NVContent(C.Structure):
    """struct _N_VectorContent_PihmB200 (include/pihm_b200_sundials.h)"""
    _fields_ = [("length", C.c_long), ("own_data", C.c_int), ("data", C.POINTER(C.c_double)),
                ("dev", C.c_void_p), ("ctx", C.c_void_p)]


class GenericNV(C.Structure):
    _fields_ = [("content", C.POINTER(NVContent)), ("ops", C.c_void_p)]


def host_mirror(nv_ptr, n):
    nv = C.cast(nv_ptr, C.POINTER(GenericNV)).contents
    assert nv.content.contents.length == n
    return np.ctypeslib.as_array(nv.content.contents.data, shape=(n,))


@pytest.mark.parametrize("size,steps,t0", [("small", 40, 0.0), ("small", 30, 3 * 3600.0), ("10k", 12, 2 * 3600.0)])
def test_reference_cvode_on_device_nvector_matches_our_integrator_bitwise(size, steps, t0):
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    tb = W.make_named(size, dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    L = lib.load_library()

    # --- A: reference CVODE + our N_Vector + our ODE ------------------------------
    mA = lib.Model(tb, reorder=1)
    nv = L.N_VNew_PihmB200(mA.h)
    assert nv
    y_host = host_mirror(nv, mA.nsv)              # NV_DATA_S view of the pinned mirror
    y_host[:] = tb["y0"]                          # InitVar writes through NV_Ith (initialize.c:572-610)
    assert L.N_VPihmB200_Push(nv) == 0
    rhs_ptr = C.cast(L.PihmB200_ODE, C.c_void_p).value
    ext = reflib.RefCvodeExternal(nv, rhs_ptr, mA.h)
    yA_vec = lib.Vec(mA, handle=L.N_VPihmB200_Device(nv))

    # --- B: our integrator ----------------------------------------------------------
    mB = lib.Model(tb, reorder=1)
    cv = lib.Cvode(mB)
    yB = mB.N_VNew(tb["y0"])
    cv.SetCVodeParam(yB)

    for k in range(steps):
        if k % 15 == 0:
            f = W.storm_forcing(tb, t0 + k * 60.0)
            mA.set_forcing(f, np.zeros(nr)); mB.set_forcing(f, np.zeros(nr))
        mA.Summary(yA_vec); mB.Summary(yB)
        tA = ext.solve((k + 1) * 60.0)
        tB = cv.SolveCVode((k + 1) * 60.0, yB)
        assert tA == tB == (k + 1) * 60.0
        assert L.N_VPihmB200_Pull(nv) == 0
        a, b = np.array(y_host), yB.download()
        sA, sB = ext.stats(), cv.stats()
        assert {k_: sB[k_] for k_ in sA} == sA, f"step {k + 1}: counters differ {sA} vs {sB}"
        assert np.array_equal(a, b), f"step {k + 1}: max |diff| {np.abs(a - b).max():.3e}"
    print(f"{size}: {steps} model steps, {sA['nst']} CVODE steps, {sA['nfe'] + sA['nfeLS']} RHS evals -- bit-identical")
    assert mA.check_nan() == 0
    ext.free()
    L.N_VDestroy_PihmB200(nv)
    cv.close(); mA.close(); mB.close()


@pytest.mark.parametrize("size,steps,t0", [("small", 40, 0.0), ("small", 30, 3 * 3600.0)])
def test_device_spgmr_plugged_into_reference_cvode(size, steps, t0):
    """SURVEY 8(b) 'Linear-solver plug-in': the reference CVODE keeps its BDF/Newton stepper and gets
    N_VNew_PihmB200, PihmB200_ODE and -- through cv_mem->cv_lsolve, attached the way CVSpgmr attaches
    itself -- pihm_b200_spgmr_solve.  Same bits and counters as the all-device integrator."""
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    tb = W.make_named(size, dirichlet_edges=True)
    nr = tb["nriver"]
    L = lib.load_library()
    mA = lib.Model(tb, reorder=1)
    nv = L.N_VNew_PihmB200(mA.h)
    y_host = host_mirror(nv, mA.nsv)
    y_host[:] = tb["y0"]
    assert L.N_VPihmB200_Push(nv) == 0
    ext = reflib.RefCvodeExternal(nv, C.cast(L.PihmB200_ODE, C.c_void_p).value, mA.h)
    engine = lib.Cvode(mA)                                  # vectors + scalar buffers of the solver
    ext.attach_lsolve(C.cast(L.pihm_b200_spgmr_solve, C.c_void_p).value, engine.h)
    yA_vec = lib.Vec(mA, handle=L.N_VPihmB200_Device(nv))

    mB = lib.Model(tb, reorder=1)
    cv = lib.Cvode(mB)
    yB = mB.N_VNew(tb["y0"])
    cv.SetCVodeParam(yB)
    for k in range(steps):
        if k % 15 == 0:
            f = W.storm_forcing(tb, t0 + k * 60.0)
            mA.set_forcing(f, np.zeros(nr)); mB.set_forcing(f, np.zeros(nr))
        mA.Summary(yA_vec); mB.Summary(yB)
        assert ext.solve((k + 1) * 60.0) == cv.SolveCVode((k + 1) * 60.0, yB) == (k + 1) * 60.0
        assert L.N_VPihmB200_Pull(nv) == 0
        sA, sE, sB = ext.stepper_stats(), engine.stats(), cv.stats()
        assert {k_: sB[k_] for k_ in sA} == sA, f"step {k + 1}: stepper counters {sA} vs {sB}"
        assert [sE[k_] for k_ in ("nli", "ncfl", "nfeLS", "njtimes")] == [sB[k_] for k_ in ("nli", "ncfl", "nfeLS", "njtimes")]
        a, b = np.array(y_host), yB.download()
        assert np.array_equal(a, b), f"step {k + 1}: max |diff| {np.abs(a - b).max():.3e}"
    print(f"{size}: {steps} model steps with the plugged-in device SPGMR, {sA['nst']} CVODE steps, "
          f"{sE['nli']} Krylov iterations -- bit-identical")
    ext.free()
    engine.close()
    L.N_VDestroy_PihmB200(nv)
    cv.close(); mA.close(); mB.close()


def dropin_run(tb, steps, t0, diagnostics=False):
    """Route 1 model steps (reference CVODE on our N_Vector / RHS) -> per step dict(y[, xf, sr]);
    with diagnostics the device Summary()/MassBalance() follows every solve (tests/test_summary_gpu.py)."""
    nr = tb["nriver"]
    L = lib.load_library()
    m = lib.Model(tb, reorder=1)
    nv = L.N_VNew_PihmB200(m.h)
    y_host = host_mirror(nv, m.nsv)
    y_host[:] = tb["y0"]
    assert L.N_VPihmB200_Push(nv) == 0
    ext = reflib.RefCvodeExternal(nv, C.cast(L.PihmB200_ODE, C.c_void_p).value, m.h)
    yv = lib.Vec(m, handle=L.N_VPihmB200_Device(nv))
    if diagnostics:
        m.set_diagnostics(True)
        m.set_ws0(yv)
    out = []
    for k in range(steps):
        if k % 15 == 0:
            m.set_forcing(W.storm_forcing(tb, t0 + k * 60.0), np.zeros(nr))
        m.Summary(yv)
        assert ext.solve((k + 1) * 60.0) == (k + 1) * 60.0
        r = {}
        if diagnostics:
            m.SummaryMB(yv, tb["stepsize"])
            r["xf"] = m.get_fluxes()[0]
            r["sr"] = m.get_summary()[0]
        assert L.N_VPihmB200_Pull(nv) == 0
        r["y"] = np.array(y_host)
        out.append(r)
    ext.free()
    L.N_VDestroy_PihmB200(nv)
    m.close()
    return out


def test_nvector_ops_table_and_mirror():
    """clone/destroy through the ops table and Serial-compatible content prefix"""
    tb = W.make_named("tiny")
    m = lib.Model(tb)
    L = lib.load_library()
    nv = L.N_VNew_PihmB200(m.h)
    host = host_mirror(nv, m.nsv)
    host[:] = np.arange(m.nsv) * 0.5
    assert L.N_VPihmB200_Push(nv) == 0
    host[:] = -1.0
    assert L.N_VPihmB200_Pull(nv) == 0
    assert np.array_equal(host, np.arange(m.nsv) * 0.5)
    L.N_VDestroy_PihmB200(nv)
    m.close()
