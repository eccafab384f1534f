"""Profiling target: a few back-to-back RHS evaluations at a given size
(used under ncu; not part of the product)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W
size = sys.argv[1] if len(sys.argv) > 1 else "1M"
fbr = len(sys.argv) > 2 and sys.argv[2] == "fbr"
n = int(os.environ.get("NREP", "5"))
tb = W.make_named(size, fbr=fbr) if os.environ.get("RIVER", "1") == "1" else W.make_watershed(*W.SIZES[size], fbr=fbr, river=False)
m = lib.Model(tb, reorder=int(os.environ.get("REORDER", "1")))
y = W.wet_state(tb, seed=11) if os.environ.get("STATE", "wet") == "wet" else tb["y0"]
m.set_forcing(W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:tb["nelem"]], 0)), np.zeros(tb["nriver"]))
yv = m.N_VNew(y); yd = m.N_VNew()
import time
for _ in range(3):
    m.ode_dev(0.0, yv, yd)
m.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    m.ode_dev(0.0, yv, yd)
m.synchronize()
print("rhs us/eval %.1f" % ((time.perf_counter() - t0) / n * 1e6), os.environ.get("PIHM_B200_LIB", "default"))
print("rhs_probe ok", size, fbr, np.isfinite(yd.download()).all(), "exact-path elements:", m.slow_path_count(), "of", (n + 3) * tb["nelem"] * 2)
