"""Synthetic watershed generator: geometric and topological invariants the
reference's Initialize() would also produce (InitTopo/InitSurfL/InitRiver)."""
import numpy as np
import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import watershed as W


@pytest.mark.parametrize("fbr", [False, True])
def test_geometry_and_topology(fbr):
    tb = W.make_named("small", fbr=fbr)
    ne, nr = tb["nelem"], tb["nriver"]
    ef, ei, rf, ri = tb["elem_f64"], tb["elem_i32"], tb["riv_f64"], tb["riv_i32"]
    assert ne == 2 * 40 * 30 and nr > 40
    assert (ef[W.E_AREA] > 0).all() and np.allclose(ef[W.E_AREA], 200.0)      # CCW, 20 m cells
    assert (ef[W.E_DEPTH] > 0.5).all() and (ef[W.E_DMAC] <= ef[W.E_DEPTH]).all()
    assert (ef[W.E_NABRDIST0:W.E_NABRDIST0 + 3] > 0).all()
    nab = ei[W.EI_NABR0:W.EI_NABR0 + 3]
    # neighbour symmetry: if n is neighbour of e then e is neighbour of n
    for j in range(3):
        sel = np.nonzero(nab[j] > 0)[0]
        n = nab[j, sel] - 1
        assert ((nab[:, n] == sel + 1).sum(axis=0) == 1).all()
    # river edges: both banks point back at the segment; down links are in range
    for r in range(nr):
        l, rt = ri[W.RI_LEFTELE, r] - 1, ri[W.RI_RIGHTELE, r] - 1
        assert (nab[:, l] == -(r + 1)).sum() == 1 and (nab[:, rt] == -(r + 1)).sum() == 1
        d = ri[W.RI_DOWN, r]
        assert (1 <= d <= nr) or d == -3
    assert (ri[W.RI_DOWN] == -3).sum() == 1
    assert (rf[W.R_ZBED] > rf[W.R_ZMIN]).all() and (rf[W.R_AREA] > 0).all()
    # the whole network drains to the outlet
    down = ri[W.RI_DOWN]
    for r in range(nr):
        k, hops = r, 0
        while down[k] > 0:
            k = down[k] - 1; hops += 1
            assert hops <= nr
    assert tb["y0"].shape == ((5 if fbr else 3) * ne + 2 * nr,)


def test_sizes_of_named_configs():
    assert W.SIZES["100k"][0] * W.SIZES["100k"][1] * 2 == 100_000
    assert W.SIZES["1M"][0] * W.SIZES["1M"][1] * 2 == 1_000_000
    assert W.SIZES["8M"][0] * W.SIZES["8M"][1] * 2 == 8_000_000
