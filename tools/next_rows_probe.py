"""Device time of the SURVEY 8(f) rows at a given size (profiling helper; not part of the product):
Summary/MassBalance (replay of the last RHS call + k_summary_mb + ws0 copy), forcing scatter +
IntcpSnowEt, print accumulation.  ET columns: the first element of tests/golden/et_example.npz
repeated (the kernel's cost does not depend on the values)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W

size = sys.argv[1] if len(sys.argv) > 1 else "1M"
tb = W.make_named(size)
ne, nr = tb["nelem"], tb["nriver"]
g = np.load(os.path.join(ROOT, "tests", "golden", "et_example.npz"))
m = lib.Model(tb, reorder=1)
m.set_diagnostics(True)
m.et_create(np.repeat(g["et_f64"][:, :1], ne, axis=1), np.repeat(g["et_i32"][:, :1], ne, axis=1))
y = W.wet_state(tb, seed=11)
m.set_forcing(W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0)), np.zeros(nr))
yv = m.N_VNew(y); yd = m.N_VNew()
m.set_ws0(yv)
st, keep = lib.make_et_step(900.0, g["cal"], g["et0_meltf"], g["et0_meteo"], g["et0_lai"], g["et0_lai_lc"], g["et0_z0_lc"])
ids = [m.print_add(W.PS_STATE, c) for c in range(5)] + [m.print_add(W.PS_ELEM_FLUX, c) for c in (W.X_INFIL, W.X_RECHG, W.X_SUB0)]


def timed(fn, n=50):
    for _ in range(3):
        fn()
    m.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    m.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def summary():
    m.ode_dev(0.0, yv, yd)          # a fresh "last call": the next SummaryMB has to re-evaluate it
    m.SummaryMB(yv, 60.0)


t_rhs = timed(lambda: m.ode_dev(0.0, yv, yd))
t_sum = timed(summary) - t_rhs
t_et = timed(lambda: m.IntcpSnowEt(st, yv))
t_pr = timed(lambda: m.UpdPrintVar(ids, yv))
print(f"{size}: RHS {t_rhs:.1f} us; Summary/MassBalance incl. the re-evaluation {t_sum:.1f} us; "
      f"IntcpSnowEt {t_et:.1f} us (incl. the by-type table upload); UpdPrintVar of {len(ids)} variables {t_pr:.1f} us")
