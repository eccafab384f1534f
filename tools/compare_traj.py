"""Diagnostic (test tooling, not product): run the reference CVODE (oracle/_ref)
and the device integrator side by side on the same watershed and print the
counters / state differences per model step."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reflib
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W

name = sys.argv[1] if len(sys.argv) > 1 else "small"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fbr = len(sys.argv) > 3 and sys.argv[3] == "fbr"
tb = W.make_named(name, fbr=fbr, dirichlet_edges=True)
ne, nr = tb["nelem"], tb["nriver"]
ref = reflib.RefModel(fbr=fbr).create_from_tables(tb)
ref.init_state(tb["y0"]); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
model = lib.Model(tb, reorder=0); cv = lib.Cvode(model)
y = model.N_VNew(tb["y0"]); cv.SetCVodeParam(y)
keys = ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")
t0 = float(os.environ.get("T0", "0"))
for k in range(nsteps):
    if k % 15 == 0:
        f = W.storm_forcing(tb, t0 + k * 60.0)
        model.set_forcing(f, np.zeros(nr))
    fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
    ref.set_forcing(fr, np.zeros(nr))
    ref.model_step(k)
    model.Summary(y); cv.SolveCVode((k + 1) * 60.0, y)
    yr, yg = ref.get_y(), y.download()
    sr, sg = ref.stats(), cv.stats()
    bound = 10 * (1e-3 * np.abs(yr) + 1e-4)
    print(k + 1, "err/bound %.3e maxabs %.3e |" % ((np.abs(yg - yr) / bound).max(), np.abs(yg - yr).max()),
          " ".join(f"{kk}:{sg[kk]}/{sr[kk]}" for kk in keys), "q %d/%d h %.4g/%.4g" % (sg["qlast"], sr["qlast"], sg["hlast"], sr["hlast"]))
