"""Print accumulation (UpdPrintVar / PrintData averaging, src/print.c:171-251;
SURVEY 8(f) f3): the numpy restatement (oraclelib.PrintVarOracle) against the
reference's own functions run on its structs through oracle/_ref.  CPU only."""
import numpy as np
import pytest

import oraclelib
import reflib
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import watershed as W


def test_print_oracle_matches_live_reference():
    if not reflib.available(False):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    tb = W.make_watershed(24, 16, dirichlet_edges=True, trib_every=8)
    ne, nr = tb["nelem"], tb["nriver"]
    ref = reflib.RefModel(fbr=False).create_from_tables(tb)
    rng = np.random.default_rng(8)
    y = W.wet_state(tb, seed=2)
    ref.init_state(y); ref.set_ovlflow(np.zeros((3, ne)))
    fields = [(W.PS_STATE, 0), (W.PS_STATE, 2), (W.PS_STATE, 3), (W.PS_ELEM_FLUX, W.X_INFIL),
              (W.PS_ELEM_FLUX, W.X_SUB0 + 1), (W.PS_ELEM_FLUX, W.X_OVL0), (W.PS_RIV_FLUX, 1), (W.PS_RIV_FLUX, 9)]
    ref.print_reset()
    ids = [ref.print_add(s, c, upd_intvl=0, intvl=180) for s, c in fields]
    orc = [oraclelib.PrintVarOracle(nr if (s == W.PS_RIV_FLUX or (s == W.PS_STATE and c in (3, 4))) else ne)
           for s, c in fields]
    records = 0
    for step in range(7):
        forc = W.storm_forcing(tb, 3600.0 * (1 + step), ws0_surf=np.maximum(y[:ne], 0))
        ref.set_forcing(forc, np.zeros(nr))
        dy = ref.ode(y * (1 + 1e-3 * rng.standard_normal(y.shape)))
        y = y + 60.0 * dy
        ref.summary(y)
        ref.print_update(0)                              # UpdPrintVar(.., HYDROL_STEP)
        ref.print_update(1)                              # a land-surface update must not touch these
        xf, rf = ref.get_fluxes(); ws = ref.get_ws()
        for (s, c), o in zip(fields, orc):
            if s == W.PS_STATE:
                blk = [ws[:ne], ws[ne:2 * ne], ws[2 * ne:3 * ne], ws[3 * ne:3 * ne + nr], ws[3 * ne + nr:3 * ne + 2 * nr]][c]
                o.update(blk)
            elif s == W.PS_ELEM_FLUX:
                o.update(xf[c])
            else:
                o.update(rf[c])
        lapse = 60 * (step + 1)
        for vid, o in zip(ids, orc):
            rec = ref.print_data(vid, 1230768000 + lapse, lapse, len(o.buffer))
            if lapse % 180 == 0:                         # PrintNow (print.c:612-616)
                assert rec is not None
                out, n = o.data()
                assert n == 3 and np.array_equal(rec, out)     # bit exact
                records += 1
            else:
                assert rec is None
    assert records == 2 * len(fields)
    ref.print_reset(); ref.close()
