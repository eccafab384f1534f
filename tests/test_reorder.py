"""The locality ordering of the element graph (csrc/reorder.h: greedy graph growing into patches of 128
elements; BASELINE.json north_star 'RCM/METIS-style mesh reordering').  The synthetic watersheds are structured
meshes whose reference order is already local, so the ordering is exercised here on a mesh whose elements were
renumbered at random -- the case of a real .mesh file written by a mesh generator: afterwards nine neighbour pairs
in ten sit in the same 128-element patch (one CTA's tiles), from one in four hundred before.  Host code only
(the single-part partition exposes the device order as `elem_gid`)."""
import numpy as np
import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import partition as PT, watershed as W


def shuffle_elements(tb, seed=3):
    """the same watershed with its elements renumbered at random (tables, neighbour codes, river banks)"""
    rng = np.random.default_rng(seed)
    ne = tb["nelem"]
    new_of_old = rng.permutation(ne)
    old_of_new = np.argsort(new_of_old)
    out = dict(tb)
    out["elem_f64"] = tb["elem_f64"][:, old_of_new].copy()
    ei = tb["elem_i32"][:, old_of_new].copy()
    for j in range(3):
        v = ei[W.EI_NABR0 + j]
        pos = v > 0
        v[pos] = new_of_old[v[pos] - 1] + 1
    out["elem_i32"] = ei
    ri = tb["riv_i32"].copy()
    for c in (W.RI_LEFTELE, W.RI_RIGHTELE):
        ri[c] = new_of_old[ri[c] - 1] + 1
    out["riv_i32"] = ri
    out["xc"], out["yc"] = tb["xc"][old_of_new], tb["yc"][old_of_new]
    out["bc_head"] = tb["bc_head"][:, old_of_new]
    y0 = tb["y0"].copy()
    for b in range(5 if tb["fbr"] else 3):
        off = b * ne if b < 3 else 3 * ne + 2 * tb["nriver"] + (b - 3) * ne
        y0[off:off + ne] = tb["y0"][off:off + ne][old_of_new]
    out["y0"] = y0
    return out, new_of_old


def pair_distances(pos, nabr):
    d = []
    for j in range(3):
        n = nabr[j]
        sel = np.nonzero(n > 0)[0]
        d.append(np.abs(pos[sel] - pos[n[sel] - 1]))
    return np.concatenate(d)


@pytest.mark.parametrize("size", ["10k", "100k"])
def test_patch_order_restores_locality_of_a_shuffled_mesh(size):
    tb = W.make_named(size)
    sh, _ = shuffle_elements(tb)
    ne = tb["nelem"]
    before = pair_distances(np.arange(ne), sh["elem_i32"][:3])
    assert (before < 128).mean() < 0.05                       # the shuffle destroyed it
    part = PT.partition(sh, 1)[0]
    order = part["elem_gid"]
    assert part["nown_elem"] == ne and np.array_equal(np.sort(order), np.arange(ne))     # a permutation
    pos = np.empty(ne, np.int64)
    pos[order] = np.arange(ne)
    after = pair_distances(pos, sh["elem_i32"][:3])
    in_patch = (after < 128).mean()
    print(f"{size}: neighbour pairs within 128 positions {100 * (before < 128).mean():.2f} % -> {100 * in_patch:.1f} %, "
          f"median distance {np.median(before):.0f} -> {np.median(after):.0f}")
    assert in_patch >= 0.85 and np.median(after) <= 32
    # as good as on the structured numbering of the same mesh
    p0 = PT.partition(tb, 1)[0]
    pos0 = np.empty(ne, np.int64)
    pos0[p0["elem_gid"]] = np.arange(ne)
    assert in_patch >= (pair_distances(pos0, tb["elem_i32"][:3]) < 128).mean() - 0.02
