"""IntcpSnowEt + the per-element forcing assignments (src/is_sm_et.c:4-225,
src/forcing.c:134-160,242-258): the oracle port against the
golden vectors written by the reference (tests/golden/make_golden.py et) and,
where oracle/_ref exists, against the live reference.  CPU only."""
import numpy as np
import pytest

import oraclelib
import reflib
from helpers import et_cases, golden_tables, load_golden
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W


def test_oracle_et_matches_golden():
    g = load_golden("et_example.npz")
    tb = golden_tables(g)
    om = oraclelib.OracleModel(tb)
    cases = et_cases(g)
    assert len(cases) == 18
    seen = dict(snow=False, melt=False, nolai=False, full=False, dry=False)
    for c in cases:
        st, keep = lib.make_et_step(c["stepsize"], g["cal"], c["meltf"], c["meteo"], c["lai"], c["lai_lc"], c["z0_lc"])
        state = np.zeros((W.EO_NCOL, tb["nelem"]))
        state[[W.EO_SNEQV, W.EO_CMC]] = c["state_in"]
        out = om.intcp_snow_et(st, g["et_f64"], g["et_i32"], c["y"], state)
        assert np.array_equal(out, c["out"])                    # bit exact
        seen["snow"] |= bool((out[W.EO_SNEQV] > c["state_in"][0]).any())
        seen["melt"] |= bool((out[W.EO_SNEQV] < c["state_in"][0]).any())
        seen["nolai"] |= bool(c["lai"][0] == 0.0)
        seen["full"] |= bool((out[W.EO_DRIP] > 0).any())
        seen["dry"] |= bool((out[W.EO_ETT] == 0).any() and (out[W.EO_ETT] > 0).any())
    assert all(seen.values()), seen
    om.close()


def test_oracle_et_matches_golden_mixed_types():
    """lai_type 0 (monthly LAI table by land-cover class) and 40 roughness classes, forcing.c:249-257"""
    g = load_golden("et_example.npz")
    tb = golden_tables(g)
    om = oraclelib.OracleModel(tb)
    eti = g["et_i32_b"]
    assert (eti[W.ETI_LAI_TYPE] == 0).any() and (eti[W.ETI_LAI_TYPE] > 0).any() and len(set(eti[W.ETI_LC_TYPE])) >= 10
    cases = et_cases(g, "etb")
    assert len(cases) == 6
    nolai = False
    for c in cases:
        st, keep = lib.make_et_step(c["stepsize"], g["cal"], c["meltf"], c["meteo"], c["lai"], c["lai_lc"], c["z0_lc"])
        state = np.zeros((W.EO_NCOL, tb["nelem"]))
        state[[W.EO_SNEQV, W.EO_CMC]] = c["state_in"]
        out = om.intcp_snow_et(st, g["et_f64"], eti, c["y"], state)
        assert np.array_equal(out, c["out"])
        nolai |= bool((c["lai_lc"][eti[W.ETI_LC_TYPE][eti[W.ETI_LAI_TYPE] == 0] - 1] == 0).any())
    assert nolai                                  # a class whose monthly LAI is zero: the lai <= 0 branch per element
    om.close()


def test_oracle_et_matches_live_reference():
    if not reflib.available(False):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import os
    if not os.path.isdir("/root/reference/input/example"):
        pytest.skip("the example project is not on this machine")
    m = reflib.RefModel(fbr=False).open_project("/root/reference", "example")
    tb = m.pack_tables()
    etf, eti = m.pack_et_tables()
    d = m.et_dims()
    om = oraclelib.OracleModel(tb)
    rng = np.random.default_rng(4)
    ne = m.nelem
    y = m.get_ws()
    for k in range(6):
        t = 1230768000 + 86400 * 45 * k + 3600 * k
        meteo = np.array([[rng.uniform(0, 5), rng.uniform(262, 303), rng.uniform(30, 100), rng.uniform(0.5, 8),
                           rng.uniform(-10, 800), 300.0, 98000.0]])
        lai = [rng.choice([0.0, 1.0, 4.0])]
        sneqv = rng.uniform(0, 0.01, ne) * (rng.random(ne) < 0.5)
        cmc = rng.uniform(-1e-6, 6e-4, ne)
        yk = y * (1 + 0.3 * rng.standard_normal(y.shape))
        m.et_set_state(sneqv, cmc)
        m.et_run(t, 900.0, meteo, lai, yk)
        lai_lc, z0_lc, meltf = m.et_monthly(t, d["nlc"])
        st, keep = lib.make_et_step(900.0, m.et_get_cal(), meltf, meteo, lai, lai_lc, z0_lc)
        state = np.zeros((W.EO_NCOL, ne)); state[W.EO_SNEQV] = sneqv; state[W.EO_CMC] = cmc
        assert np.array_equal(om.intcp_snow_et(st, etf, eti, yk, state), m.et_get())
    m.close(); om.close()
