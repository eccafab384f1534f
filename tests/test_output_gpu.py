"""Output files written from the device columns (csrc/output.cu; SURVEY 8(f) f3, second half): the .dat / .txt
print records and the restart .ic file have the bytes the reference's InitOutputFile / PrintData / PrintInit
(src/print.c:72-313) would write for the same values -- record = { (double)t ; value[nvar] } with value =
buffer / counter, text lines "<time>"\\t%lf..., restart record { cmc, sneqv, surf, unsat, gw [, fbr_unsat, fbr_gw] }
per element then { stage, gw } per river -- and all the variables due at a print time cross PCIe in ONE copy.
The values are those of the numpy restatement of UpdPrintVar / PrintData (oraclelib.PrintVarOracle, pinned to the
reference in tests/test_print.py) fed with the per-step values of the same run.  The reference's own files are
compared with the files of the unchanged driver in tests/test_driver_gpu.py."""
import numpy as np
import pytest

import oraclelib
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu
FIELDS = [("surf", W.PS_STATE, 0), ("unsat", W.PS_STATE, 1), ("gw", W.PS_STATE, 2), ("stage", W.PS_STATE, 3),
          ("rivgw", W.PS_STATE, 4), ("infil", W.PS_ELEM_FLUX, W.X_INFIL), ("recharge", W.PS_ELEM_FLUX, W.X_RECHG),
          ("subflx0", W.PS_ELEM_FLUX, W.X_SUB0), ("rivflx1", W.PS_RIV_FLUX, 1)]


def is_river(s, c):
    return s == W.PS_RIV_FLUX or (s == W.PS_STATE and c in (3, 4))


def step_values(tb, s, c, y, xf, rf):
    ne, nr = tb["nelem"], tb["nriver"]
    if s == W.PS_STATE:
        return [y[:ne], y[ne:2 * ne], y[2 * ne:3 * ne], y[3 * ne:3 * ne + nr], y[3 * ne + nr:3 * ne + 2 * nr]][c]
    return xf[c] if s == W.PS_ELEM_FLUX else rf[c]


@pytest.mark.parametrize("reorder", [0, 1])
def test_dat_txt_records_have_the_reference_bytes(tmp_path, reorder):
    tb = W.make_named("small", dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    model = lib.Model(tb, reorder=reorder)
    model.set_diagnostics(True)
    y = model.N_VNew(tb["y0"])
    model.set_ws0(y)
    cv = lib.Cvode(model)
    cv.SetCVodeParam(y)
    ids = [model.print_add(s, c) for _, s, c in FIELDS]
    for vid, (name, _, _) in zip(ids, FIELDS):
        model.print_open(vid, str(tmp_path / f"proj.{name}"), ascii=True)
    orc = [oraclelib.PrintVarOracle(nr if is_river(s, c) else ne) for _, s, c in FIELDS]
    want_dat = {name: [] for name, _, _ in FIELDS}
    want_txt = {name: [] for name, _, _ in FIELDS}
    nrec = 0
    for k in range(9):
        if k % 15 == 0:
            model.set_forcing(W.storm_forcing(tb, 3600.0 + k * 60.0), np.zeros(nr))
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
        model.SummaryMB(y, tb["stepsize"])
        model.UpdPrintVar(ids, y)
        xf, rf = model.get_fluxes(); yh = y.download()
        for (_, s, c), o in zip(FIELDS, orc):
            o.update(step_values(tb, s, c, yh, xf, rf))
        t = 1230768000 + 60 * (k + 1)
        if (k + 1) % 3 == 0:                      # PrintNow: every 180 s all variables are due
            timestr = f"2009-01-01 00:{k + 1:02d}"
            model.print_write(ids, t, timestr)
            nrec += 1
            for (name, _, _), o in zip(FIELDS, orc):
                vals, cnt = o.data()
                assert cnt == 3
                want_dat[name].append(np.concatenate([[float(t)], vals]))
                want_txt[name].append('"%s"' % timestr + "".join("\t%f" % v for v in vals) + "\n")
    copies, nbytes = model.print_io_stats()
    assert copies == nrec, "one device -> host copy per print time"
    assert nbytes == nrec * 8 * sum(nr if is_river(s, c) else ne for _, s, c in FIELDS)
    model.print_close()
    for name, _, _ in FIELDS:
        got = np.fromfile(tmp_path / f"proj.{name}.dat", dtype=np.float64)
        want = np.concatenate(want_dat[name])
        assert got.tobytes() == want.tobytes(), name
        assert (tmp_path / f"proj.{name}.txt").read_text() == "".join(want_txt[name]), name
    cv.close(); model.close()


@pytest.mark.parametrize("fbr", [False, True])
def test_restart_ic_file(tmp_path, fbr):
    tb = W.make_named("small", fbr=fbr)
    ne, nr = tb["nelem"], tb["nriver"]
    model = lib.Model(tb, reorder=1)
    rng = np.random.default_rng(3)
    yh = tb["y0"] * (1.0 + 0.1 * rng.random(len(tb["y0"])))
    y = model.N_VNew(yh)
    cmc, sneqv = rng.random(ne) * 1e-3, rng.random(ne) * 1e-2
    path = tmp_path / "proj.200901020000.ic"
    model.write_ic(str(path), y, cmc=cmc, sneqv=sneqv)
    # PrintInit, src/print.c:272-305
    cols = [cmc, sneqv, yh[:ne], yh[ne:2 * ne], yh[2 * ne:3 * ne]]
    if fbr:
        o = 3 * ne + 2 * nr
        cols += [yh[o:o + ne], yh[o + ne:o + 2 * ne]]
    want = np.concatenate([np.stack(cols, axis=1).ravel(),
                           np.stack([yh[3 * ne:3 * ne + nr], yh[3 * ne + nr:3 * ne + 2 * nr]], axis=1).ravel()])
    got = np.fromfile(path, dtype=np.float64)
    assert got.tobytes() == want.tobytes()
    # without host arrays and without the device ET state: zero storages, like Initialize() without an .ic file
    model.write_ic(str(path), y)
    got = np.fromfile(path, dtype=np.float64)
    per = 7 if fbr else 5
    assert not got[:ne * per].reshape(ne, per)[:, :2].any()
    assert np.array_equal(got[:ne * per].reshape(ne, per)[:, 2], yh[:ne])
    model.close()
