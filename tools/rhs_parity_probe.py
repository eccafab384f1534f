"""Profiling / tuning helper (not part of the product): back-to-back RHS time of one library
variant (PIHM_B200_LIB) at a given size and its parity against the C oracle on the same state,
in the metric of tests/test_rhs_gpu.py.   usage: rhs_parity_probe.py <size> [fbr] [state...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W
import oraclelib
from helpers import dy_scale, rel_err

size = sys.argv[1] if len(sys.argv) > 1 else "1M"
fbr = "fbr" in sys.argv[2:]
nrep = int(os.environ.get("NREP", "50"))
tag = os.environ.get("PIHM_B200_LIB", "default")
tb = W.make_named(size, fbr=fbr)
ne, nr = tb["nelem"], tb["nriver"]
m = lib.Model(tb, reorder=1)
om = oraclelib.OracleModel(tb) if os.environ.get("PARITY", "1") == "1" else None
for seed in (11, 12):
    y = W.wet_state(tb, seed=seed)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    m.set_forcing(forc, np.zeros(nr))
    yv = m.N_VNew(y); yd = m.N_VNew()
    if seed == 11:
        for _ in range(3):
            m.ode_dev(0.0, yv, yd)
        m.synchronize()
        t0 = time.perf_counter()
        for _ in range(nrep):
            m.ode_dev(0.0, yv, yd)
        m.synchronize()
        us = (time.perf_counter() - t0) / nrep * 1e6
    if om is None:
        print(f"[{tag}] {size} fbr={fbr}: rhs {us:.1f} us"); break
    m.set_stale_ovlflow(np.zeros((3, ne)))
    m.ode_dev(0.0, yv, yd)
    dy = yd.download()
    om.set_forcing(forc, np.zeros(nr)); om.set_stale_ovlflow(np.zeros((3, ne)))
    ref = om.ode(y)
    xf, rf = om.get_fluxes()
    err = rel_err(dy, ref, dy_scale(tb, forc, xf, rf))
    k = int(err.argmax())
    srt = np.sort(err)[::-1]
    print(f"[{tag}] {size} fbr={fbr} seed={seed}: rhs {us:.1f} us  max err {err.max():.2e} at {k} (top5 {srt[:5]})"
          f"  >1e-13: {(err > 1e-13).sum()}  bit-exact {np.mean(dy == ref):.4f}  slow-path {m.slow_path_count()}", flush=True)
