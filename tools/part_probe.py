"""profiling helper: RHS time of one rank's share of a partitioned mesh WITHOUT communication (ghost records set
once), against the unpartitioned mesh of the same size -- isolates the cost of the ghost-aware kernel variants"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_pihm_b200  # noqa
from mm_pihm_b200 import lib, watershed as W, partition as PT
nrep = int(os.environ.get("NREP", "100"))
fbr = "fbr" in sys.argv[1:]
def t_rhs(m, yv, yd):
    for _ in range(5): m.ode_dev(0.0, yv, yd)
    m.synchronize(); t0 = time.perf_counter()
    for _ in range(nrep): m.ode_dev(0.0, yv, yd)
    m.synchronize(); return (time.perf_counter() - t0) / nrep * 1e6
tb = W.make_named("2M", fbr=fbr)
y = W.wet_state(tb, seed=5)
part = PT.partition(tb, 2, parts=[1])[0]
forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:tb["nelem"]], 0))[:, part["elem_gid"]]
m = lib.Model(part)
m.set_forcing(forc, np.zeros(part["nriver"]))
m.set_ghosts(*PT.ghost_records(part, y))
yv = m.N_VNew(PT.local_state(part, y)); yd = m.N_VNew()
print(f"partition 1 of 2 ({part['nown_elem']} owned + {part['nelem'] - part['nown_elem']} ghosts), fbr={fbr}: rhs {t_rhs(m, yv, yd):.1f} us")
print("slow-path elements:", m.slow_path_count())
m.close()
# the same local mesh as a plain (unpartitioned) mesh: ghosts become ordinary elements
plain = {k: v for k, v in part.items() if k not in ("nown_elem", "nown_riv")}
m = lib.Model(plain, reorder=0)
m.set_forcing(forc, np.zeros(part["nriver"]))
yv = m.N_VNew(PT.local_state(part, y, extended=True)); yd = m.N_VNew()
print(f"same local mesh, no ghost distinction ({part['nelem']} elements): rhs {t_rhs(m, yv, yd):.1f} us; slow-path {m.slow_path_count()}")
m.close()
tb1 = W.make_named("1M", fbr=fbr)
y1 = W.wet_state(tb1, seed=5)
m = lib.Model(tb1, reorder=1)
m.set_forcing(W.storm_forcing(tb1, 3 * 3600.0, ws0_surf=np.maximum(y1[:tb1["nelem"]], 0)), np.zeros(tb1["nriver"]))
yv = m.N_VNew(y1); yd = m.N_VNew()
print(f"unpartitioned 1M, fbr={fbr}: rhs {t_rhs(m, yv, yd):.1f} us")
