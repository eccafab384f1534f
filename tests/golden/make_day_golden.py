"""One simulated day of the synthetic 100k-triangle watershed through the REFERENCE (oracle/_ref:
the unmodified MM-PIHM + CVODE compiled from /root/reference), twice: from the RelaxIc state and
from that state perturbed by 1e-15 relative (the reference's own sensitivity, which calibrates
the long-horizon tolerance).  Writes tests/golden/day_100k.npz (states after 1/2 and 1 day,
counters).  Run in the build container (needs oracle/_ref):  python tests/golden/make_day_golden.py
The GPU test tests/test_cvode_gpu.py::test_100k_one_day replays the same steps on the device."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mm_pihm_b200  # noqa
from mm_pihm_b200 import watershed as W
import reflib

SIZE, NSTEPS, SNAPS = "100k", 1440, (720, 1440)
KEYS = ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")


def run(tb, y0):
    ne, nr = tb["nelem"], tb["nriver"]
    ref = reflib.RefModel(fbr=False, cvode_omp=True, threads=os.cpu_count()).create_from_tables(tb)
    ref.init_state(y0); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
    ys, sts = [], []
    for k in range(NSTEPS):
        if k % 15 == 0:
            f = W.storm_forcing(tb, k * 60.0)
        fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(fr, np.zeros(nr))
        ref.model_step(k)
        if k + 1 in SNAPS:
            ys.append(ref.get_y()); s = ref.stats(); sts.append([s[q] for q in KEYS])
    ref.close()
    return np.array(ys), np.array(sts)


if __name__ == "__main__":
    tb = W.make_named(SIZE)
    t0 = time.time()
    y, st = run(tb, tb["y0"])
    print("reference run", time.time() - t0, "s", dict(zip(KEYS, st[-1])))
    yp, stp = run(tb, tb["y0"] * (1.0 + 1e-15))
    unit = 1e-3 * np.abs(y) + 1e-4
    print("self-sensitivity (x (reltol|y|+abstol)) at the snapshots:", (np.abs(yp - y) / unit).max(axis=1))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "day_100k.npz"), steps=np.array(SNAPS),
                        y=y, y_pert=yp, stats=st, stats_pert=stp, stat_keys=np.array(KEYS))
