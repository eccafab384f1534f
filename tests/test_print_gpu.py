"""Print accumulation on the device (pihm_b200_print_add / print_update /
print_data): running sums of states, element fluxes, river
fluxes and ET outputs stay in HBM between print records.  Bit-exact against the
numpy restatement of UpdPrintVar / PrintData (pinned to the reference in
tests/test_print.py) fed with the per-step values, and against the reference's
own print records along its CVODE run while the integrators are in lock step."""
import numpy as np
import pytest

import oraclelib
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu
FIELDS = [(W.PS_STATE, 0), (W.PS_STATE, 1), (W.PS_STATE, 2), (W.PS_STATE, 3), (W.PS_STATE, 4),
          (W.PS_ELEM_FLUX, W.X_INFIL), (W.PS_ELEM_FLUX, W.X_RECHG), (W.PS_ELEM_FLUX, W.X_SUB0),
          (W.PS_ELEM_FLUX, W.X_OVL0 + 2), (W.PS_RIV_FLUX, 1), (W.PS_RIV_FLUX, 6)]


def is_river(s, c):
    return s == W.PS_RIV_FLUX or (s == W.PS_STATE and c in (3, 4))


def step_values(tb, s, c, y, xf, rf):
    ne, nr = tb["nelem"], tb["nriver"]
    if s == W.PS_STATE:
        return [y[:ne], y[ne:2 * ne], y[2 * ne:3 * ne], y[3 * ne:3 * ne + nr], y[3 * ne + nr:3 * ne + 2 * nr]][c]
    return xf[c] if s == W.PS_ELEM_FLUX else rf[c]


@pytest.mark.parametrize("reorder", [0, 1])
def test_print_accumulation_bit_exact(reorder):
    tb = W.make_named("small", dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    model = lib.Model(tb, reorder=reorder)
    model.set_diagnostics(True)
    y = model.N_VNew(tb["y0"])
    model.set_ws0(y)
    cv = lib.Cvode(model)
    cv.SetCVodeParam(y)
    ids = [model.print_add(s, c) for s, c in FIELDS]
    orc = [oraclelib.PrintVarOracle(nr if is_river(s, c) else ne) for s, c in FIELDS]
    records = 0
    for k in range(12):
        if k % 15 == 0:
            model.set_forcing(W.storm_forcing(tb, 3600.0 + k * 60.0), np.zeros(nr))
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
        model.SummaryMB(y, tb["stepsize"])
        model.UpdPrintVar(ids, y)                      # UpdPrintVar(.., HYDROL_STEP), nothing leaves the device
        xf, rf = model.get_fluxes(); yh = y.download()
        for (s, c), o in zip(FIELDS, orc):
            o.update(step_values(tb, s, c, yh, xf, rf))
        if (k + 1) % 4 == 0:                           # PrintNow: a record every 240 s
            for vid, (s, c), o in zip(ids, FIELDS, orc):
                out, n = model.PrintData(vid, river=is_river(s, c))
                ref, nref = o.data()
                assert n == nref == 4
                assert np.array_equal(out, ref), (k, s, c)
                records += 1
    assert records == 3 * len(FIELDS)
    out, n = model.PrintData(ids[0])                   # nothing accumulated: counter 0 -> zeros
    assert n == 0 and not out.any()
    cv.close(); model.close()


def test_print_records_along_live_reference():
    import reflib
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    tb = W.make_named("small", dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    ref = reflib.RefModel(fbr=False).create_from_tables(tb)
    ref.init_state(tb["y0"]); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
    ref.print_reset()
    rids = [ref.print_add(s, c, upd_intvl=0, intvl=240) for s, c in FIELDS]
    model = lib.Model(tb, reorder=1)
    model.set_diagnostics(True)
    y = model.N_VNew(tb["y0"]); model.set_ws0(y)
    cv = lib.Cvode(model); cv.SetCVodeParam(y)
    ids = [model.print_add(s, c) for s, c in FIELDS]
    for k in range(8):
        if k % 15 == 0:
            f = W.storm_forcing(tb, 3600.0 + k * 60.0)
            model.set_forcing(f, np.zeros(nr))
        fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(fr, np.zeros(nr))
        ref.model_step(k)                               # SolveCVode + Summary
        ref.print_update(0)
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
        model.SummaryMB(y, tb["stepsize"])
        model.UpdPrintVar(ids, y)
        lapse = 60 * (k + 1)
        for rid, vid, (s, c) in zip(rids, ids, FIELDS):
            n = nr if is_river(s, c) else ne
            rec = ref.print_data(rid, lapse, lapse, n)
            if lapse % 240 == 0:
                assert rec is not None
                out, cnt = model.PrintData(vid, river=is_river(s, c))
                assert cnt == 4
                scale = max(np.abs(rec).max(), 1e-30)
                e = np.abs(out - rec).max() / scale
                assert e <= 1e-5, f"record at {lapse} s, field {(s, c)}: {e:.2e} of the column scale"
            else:
                assert rec is None
    ref.print_reset(); ref.close(); cv.close(); model.close()
