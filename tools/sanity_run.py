"""Small end-to-end run for compute-sanitizer (RHS pihm + fbr on ragged meshes, a few model steps)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W
for fbr in (False, True):
    for nx, ny, river in ((7, 5, False), (40, 30, True), (13, 9, True)):
        tb = W.make_watershed(nx, ny, fbr=fbr, river=river, dirichlet_edges=True)
        m = lib.Model(tb)
        y = W.wet_state(tb, seed=3)
        m.set_forcing(W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:tb["nelem"]], 0)), np.zeros(tb["nriver"]))
        dy = m.ODE(0.0, y)
        assert np.isfinite(dy).all()
        cv = lib.Cvode(m)
        yv = m.N_VNew(tb["y0"])
        cv.SetCVodeParam(yv)
        for k in range(3):
            cv.SolveCVode((k + 1) * 60.0, yv)
        assert np.isfinite(yv.download()).all()
        cv.close(); m.close()
        print("ok", fbr, nx, ny, river)
