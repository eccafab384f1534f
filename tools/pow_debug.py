import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mm_pihm_b200
from mm_pihm_b200 import lib
L = lib.load_library()
def run(x, y):
    fast = np.empty_like(x); ref = np.empty_like(x)
    L.pihm_b200_test_pow(len(x), x.ctypes.data, y.ctypes.data, fast.ctypes.data, ref.ctypes.data)
    return fast, ref
rng = np.random.default_rng(0); n = 1 << 18
groups = {"s^m": (rng.uniform(0.1, 1.0, n), rng.uniform(1.0, 10.0, n)),
          "small^frac": (10.0 ** rng.uniform(-12, 0, n), rng.uniform(0.05, 1.0, n)),
          "big^frac": (rng.uniform(1.0, 1.0e5, n), rng.uniform(0.1, 1.0, n)),
          "manning": (10.0 ** rng.uniform(-8, 1, n), np.full(n, 0.6666667))}
for k, (x, y) in groups.items():
    f, r = run(x, y)
    bad = np.nonzero(f != r)[0]
    print(k, len(bad), "of", n)
    for i in bad[:4]:
        z = y[i] * np.log(x[i])
        print("   x=%.17g y=%.17g fast=%.17g ref=%.17g z=%.6f frac(z/ln2)=%.6f" % (x[i], y[i], f[i], r[i], z, (z / np.log(2)) % 1))
