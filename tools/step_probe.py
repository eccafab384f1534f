"""profiling helper: RHS evaluations per model step of the bench workload under variations of the step protocol"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W
size = sys.argv[1] if len(sys.argv) > 1 else "1M"
mode = sys.argv[2] if len(sys.argv) > 2 else "old"
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
T0 = 2 * 3600.0
tb = W.make_named(size)
ne, nr = tb["nelem"], tb["nriver"]
m = lib.Model(tb, reorder=1)
cv = lib.Cvode(m)
y = m.N_VNew(tb["y0"] * (1.0 + float(os.environ.get("PERTURB", "0"))))
if mode != "old":
    m.set_diagnostics(True)
m.set_stale_ovlflow(np.zeros((3, ne)))
m.set_forcing(W.storm_forcing(tb, T0), np.zeros(nr))
if mode != "old":
    m.set_ws0(y)
cv.SetCVodeParam(y)
t0 = time.perf_counter()
for k in range(nsteps):
    if k % 15 == 0:
        m.set_forcing(W.storm_forcing(tb, T0 + k * 60.0), np.zeros(nr))
        if mode != "old":
            m.Summary(y)
    if mode == "old":
        m.Summary(y)
    cv.SolveCVode((k + 1) * 60.0, y)
    if mode == "new":
        m.SummaryMB(y, 60.0)
    elif mode == "new_nomb":
        m.Summary(y)
    st = cv.stats()
    if k % 5 == 4:
        print(f"[{mode}] step {k+1}: nst {st['nst']} nfe+nfeLS {st['nfe'] + st['nfeLS']} nni {st['nni']} ncfn {st['ncfn']} netf {st['netf']}", flush=True)
print(f"[{mode}] {size}: {(st['nfe'] + st['nfeLS']) / nsteps:.2f} evals/step, {(time.perf_counter() - t0) / nsteps * 1e3:.2f} ms/step wall")
