"""One simulated day of a synthetic watershed through the REFERENCE (oracle/_ref: the unmodified MM-PIHM +
CVODE compiled from /root/reference), twice: from the RelaxIc state and from that state perturbed by 1e-15
relative (the reference's own sensitivity, which calibrates the long-horizon tolerance).  Writes
tests/golden/day_<size>.npz: states and counters at the snapshots (6 h = end of the storm, 12 h, 24 h).

    python tests/golden/make_day_golden.py [size] [nsteps]        (build container; needs oracle/_ref)

The GPU test tests/test_cvode_gpu.py::test_one_simulated_day replays the same steps on the device.
Cost: the day is far more expensive per model step than its first half hour -- from t = 30 min on the RelaxIc
state (unsat + gw == depth, the threshold of Infil(), src/vert_flow.c:52) saturates from below along the valley
and CVODE needs 20-80 internal steps per 60 s model step (100k triangles: 5 s of 8 host cores per model step,
i.e. two hours for the day; 2400 triangles: minutes)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mm_pihm_b200  # noqa: E402,F401
from mm_pihm_b200 import watershed as W  # noqa: E402
import reflib  # noqa: E402

KEYS = ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")


def run(tb, y0, nsteps, snaps, tag):
    ne, nr = tb["nelem"], tb["nriver"]
    big = ne >= 30000
    ref = reflib.RefModel(fbr=False, cvode_omp=big, threads=(os.cpu_count() if big else 2)).create_from_tables(tb)
    ref.init_state(y0); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
    ys, sts = [], []
    t0 = time.time()
    for k in range(nsteps):
        if k % 15 == 0:
            f = W.storm_forcing(tb, k * 60.0)
        fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(fr, np.zeros(nr))
        ref.model_step(k)
        if k + 1 in snaps:
            ys.append(ref.get_y()); s = ref.stats(); sts.append([s[q] for q in KEYS])
            print(f"[{tag}] step {k + 1}: {time.time() - t0:.0f} s, nst {s['nst']}, rhs evals {s['nfe'] + s['nfeLS']}", flush=True)
    ref.close()
    return np.array(ys), np.array(sts)


if __name__ == "__main__":
    size = sys.argv[1] if len(sys.argv) > 1 else "small"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1440
    snaps = tuple(s for s in (360, 720, 1440) if s <= nsteps) or (nsteps,)
    tb = W.make_named(size, dirichlet_edges=(size == "small"))
    y, st = run(tb, tb["y0"], nsteps, snaps, "reference")
    yp, stp = run(tb, tb["y0"] * (1.0 + 1e-15), nsteps, snaps, "perturbed")
    unit = 1e-3 * np.abs(y) + 1e-4
    print("self-sensitivity (x (reltol|y|+abstol)) at the snapshots:", (np.abs(yp - y) / unit).max(axis=1))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"day_{size}.npz"), steps=np.array(snaps),
                        y=y, y_pert=yp, stats=st, stats_pert=stp, stat_keys=np.array(KEYS))
