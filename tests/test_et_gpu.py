"""IntcpSnowEt + forcing scatter on the device (pihm_b200_et_create /
pihm_b200_intcp_snow_et) through the C ABI against the golden
vectors written by the reference on input/example.

Bound: every output within 1e-12 of max(|reference|, largest entry of the
column in that call) -- the arithmetic is IEEE and in the reference's order,
only exp / log / pow / cos differ from glibc in the last ulp, and differences
like 1 - (cmc/max)^cfactr lose relative but not absolute accuracy.  The two
storages ws.sneqv / ws.cmc and the snow/rain split are bit exact."""
import numpy as np
import pytest

import oraclelib
from helpers import et_cases, golden_tables, load_golden
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu
NAMES = ("pcpdrp", "edir", "ett", "ec", "drip", "sneqv", "cmc")


def check(out, ref, tag):
    worst = 0.0
    for c in range(W.EO_NCOL):
        scale = max(np.abs(ref[c]).max(), 1e-300)
        e = np.abs(out[c] - ref[c]) / np.maximum(np.abs(ref[c]), scale)
        worst = max(worst, e.max())
        assert e.max() <= 1e-12, f"{tag}: {NAMES[c]} differs by {e.max():.2e} at {np.argmax(e)}"
    return worst


@pytest.mark.parametrize("reorder", [0, 1])
def test_et_matches_golden(reorder):
    g = load_golden("et_example.npz")
    tb = golden_tables(g)
    ne, nr = tb["nelem"], tb["nriver"]
    model = lib.Model(tb, reorder=reorder)
    model.et_create(g["et_f64"], g["et_i32"])
    om = oraclelib.OracleModel(tb)
    yv = model.N_VNew()
    worst, exact = 0.0, []
    for k, c in enumerate(et_cases(g)):
        st, keep = lib.make_et_step(c["stepsize"], g["cal"], c["meltf"], c["meteo"], c["lai"], c["lai_lc"], c["z0_lc"])
        model.set_forcing(np.zeros((W.F_NCOL, ne)), np.zeros(nr))     # the kernel must fill the columns itself
        model.et_set_state(c["state_in"][0], c["state_in"][1])
        yv.upload(c["y"])
        model.IntcpSnowEt(st, yv)
        out = model.et_get()
        worst = max(worst, check(out, c["out"], f"case {k}"))
        exact.append(np.mean(out == c["out"]))
        # the RHS now sees wf.pcpdrp / edir / ett: same dy as the oracle fed with the reference's columns
        f = np.zeros((W.F_NCOL, ne))
        f[W.F_PCPDRP], f[W.F_EDIR], f[W.F_ETT] = c["out"][W.EO_PCPDRP], c["out"][W.EO_EDIR], c["out"][W.EO_ETT]
        om.set_forcing(f, np.zeros(nr)); om.set_stale_ovlflow(np.zeros((3, ne)))
        model.set_stale_ovlflow(np.zeros((3, ne)))
        y = np.maximum(c["y"], 0.0)
        dy, dyo = model.ODE(0.0, y), om.ode(y)
        scale = np.abs(dyo).max()
        assert np.abs(dy - dyo).max() <= 1e-11 * scale, f"case {k}: RHS after the device ET step differs"
    print(f"reorder={reorder}: {len(exact)} calls, worst {worst:.2e}, bit-exact share {np.mean(exact):.3f}")
    model.close(); om.close()


def test_et_sequence_keeps_storages_on_device():
    """the reference's own four etsteps of the first hour, storages carried on the device"""
    g = load_golden("et_example.npz")
    tb = golden_tables(g)
    model = lib.Model(tb, reorder=1)
    model.et_create(g["et_f64"], g["et_i32"])
    yv = model.N_VNew()
    cases = et_cases(g)[:4]
    model.et_set_state(cases[0]["state_in"][0], cases[0]["state_in"][1])
    for k, c in enumerate(cases):
        st, keep = lib.make_et_step(c["stepsize"], g["cal"], c["meltf"], c["meteo"], c["lai"], c["lai_lc"], c["z0_lc"])
        yv.upload(c["y"])
        model.IntcpSnowEt(st, yv)
        check(model.et_get(), c["out"], f"etstep {k}")
    model.close()


def test_et_argument_checks():
    g = load_golden("et_example.npz")
    tb = golden_tables(g)
    model = lib.Model(tb)
    yv = model.N_VNew()
    c = et_cases(g)[0]
    st, keep = lib.make_et_step(c["stepsize"], g["cal"], c["meltf"], c["meteo"], c["lai"], c["lai_lc"], c["z0_lc"])
    with pytest.raises(RuntimeError):
        model.IntcpSnowEt(st, yv)                      # et_create not called
    model.et_create(g["et_f64"], g["et_i32"])
    st2, keep2 = lib.make_et_step(c["stepsize"], g["cal"], c["meltf"], c["meteo"], c["lai"], c["lai_lc"][:3], c["z0_lc"][:3])
    with pytest.raises(RuntimeError):
        model.IntcpSnowEt(st2, yv)                     # lc table shorter than the largest lc_type
    bad = g["et_i32"].copy(); bad[W.ETI_METEO_TYPE, 0] = 0
    with pytest.raises(RuntimeError):
        model.et_create(g["et_f64"], bad)
    model.close()
