"""RHS parity on the GPU: libpihm_b200's ODE() through the C ABI (host
buffers in, host buffers out) against
  * the committed golden vectors produced by the reference itself,
  * the C oracle port on fresh seeded states (sizes it finishes in seconds),
  * size-independent properties at the 1M-triangle benchmark size.
Tolerance (BASELINE.json north_star / SURVEY 8(d)): 1e-12 relative per
component, measured against max(|dy_ref|, sum |flux terms|)."""
import numpy as np
import pytest

import oraclelib
from helpers import dy_scale, golden_cases, golden_tables, load_golden, record, rel_err
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def check_case(model, tb, c, tag=""):
    model.set_forcing(c["forc"], c["rivbc"])
    model.set_stale_ovlflow(c["stale"])
    model.set_flux_recording(True)
    dy = model.ODE(0.0, c["y"])
    xf, rf = model.get_fluxes()
    scale = dy_scale(tb, c["forc"], c["xflux"], c["rivflow"])
    err = rel_err(dy, c["dy"], scale)
    assert model.nan_flag == 0
    assert err.max() <= RTOL, f"{tag}: dy rel err {err.max():.3e} at {err.argmax()}"
    # fluxes: relative to their own magnitude (no cancellation inside a flux)
    fe = rel_err(xf, c["xflux"], np.abs(c["xflux"]).max(axis=1, keepdims=True) * 1e-6)
    assert fe.max() <= 1e-10, f"{tag}: flux rel err {fe.max():.3e} col {np.unravel_index(fe.argmax(), fe.shape)}"
    if rf.size:
        re_ = rel_err(rf, c["rivflow"], np.abs(c["rivflow"]).max(axis=1, keepdims=True) * 1e-6 + 1e-300)
        assert re_.max() <= 1e-10, f"{tag}: rivflow rel err {re_.max():.3e}"
    return err.max(), float((dy == c["dy"]).mean())


@pytest.mark.parametrize("name", ["example_pihm.npz", "example_fbr.npz"])
@pytest.mark.parametrize("reorder", [0, 1])
def test_rhs_golden_example(name, reorder):
    g = load_golden(name)
    tb = golden_tables(g)
    model = lib.Model(tb, reorder=reorder)
    for k, c in enumerate(golden_cases(g)):
        worst, exact = check_case(model, tb, c, f"{name}[{k}]")
        print(f"{name} case {k} reorder {reorder}: max rel {worst:.2e}, bit-exact fraction {exact:.4f}")
    model.close()


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("reorder", [0, 1])
def test_rhs_golden_synthetic(fbr, reorder):
    g = load_golden("synth_small_fbr.npz" if fbr else "synth_small_pihm.npz")
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    model = lib.Model(tb, reorder=reorder)
    for k, c in enumerate(golden_cases(g)):
        check_case(model, tb, c, f"synth fbr={fbr}[{k}]")
    model.close()


def _oracle_case(tb, om, y, forc, rivbc, stale):
    om.set_forcing(forc, rivbc); om.set_stale_ovlflow(stale)
    dy = om.ode(y)
    xf, rf = om.get_fluxes()
    return dict(y=y, forc=forc, rivbc=rivbc, stale=stale, dy=dy, xflux=xf, rivflow=rf)


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("modes", [(2, 2), (1, 1)])
@pytest.mark.parametrize("riv_order", [1, 2, 3, 4])
def test_rhs_vs_oracle_variants(fbr, modes, riv_order):
    """every river cross-section shape, both routing modes, all outlet codes,
    Dirichlet/Neumann element and river boundary conditions"""
    tb = W.make_watershed(24, 16, fbr=fbr, dirichlet_edges=True, surf_mode=modes[0],
                          riv_mode=modes[1], riv_order=riv_order, trib_every=8)
    ne, nr = tb["nelem"], tb["nriver"]
    tb["elem_i32"][W.EI_BC2, 0] = -1
    if fbr:     # FbrBoundFluxElem's Neumann branch (lat_flow.c:392-424): a prescribed bedrock flux on the same edge
        tb["elem_i32"][W.EI_FBRBC2, 0] = -1
    tb["riv_i32"][W.RI_BCTYPE, 3] = 1
    tb["riv_i32"][W.RI_BCTYPE, 5] = -1
    rng = np.random.default_rng(riv_order + 10 * fbr)
    for outlet in (-1, -2, -3, -4):
        tb["riv_i32"][W.RI_DOWN, 23] = outlet
        om = oraclelib.OracleModel(tb)
        model = lib.Model(tb, reorder=outlet % 2)
        y = W.wet_state(tb, seed=outlet + 20)
        forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
        forc[W.F_BC2, 0] = 1e-5
        if fbr:
            forc[W.F_FBRBC2, 0] = 3e-6
        rivbc = rng.uniform(0.0, 1e-3, nr)
        rivbc[3] = tb["riv_f64"][W.R_ZBED, 3] - 0.3
        rivbc[23] = tb["riv_f64"][W.R_ZBED, 23] - 0.4 if outlet == -1 else 2e-3
        stale = rng.standard_normal((3, ne)) * 1e-5
        c = _oracle_case(tb, om, y, forc, rivbc, stale)
        check_case(model, tb, c, f"variant ord={riv_order} outlet={outlet}")
        # second call, no explicit stale: the device must carry the hidden state itself
        model.set_forcing(forc, rivbc)
        dy2 = model.ODE(0.0, y)
        d2 = om.ode(y)
        xf, rf = om.get_fluxes()
        err = rel_err(dy2, d2, dy_scale(tb, forc, xf, rf))
        assert err.max() <= RTOL
        model.close()


def test_rhs_no_river_and_ragged_sizes():
    """meshes without rivers, and element counts that are not multiples of the CTA size"""
    for nx, ny in ((3, 2), (7, 5), (13, 9)):
        tb = W.make_watershed(nx, ny, river=False)
        assert tb["nriver"] == 0
        om = oraclelib.OracleModel(tb)
        model = lib.Model(tb)
        y = W.wet_state(tb, seed=nx)
        forc = W.storm_forcing(tb, 4 * 3600.0, ws0_surf=np.maximum(y[:tb["nelem"]], 0))
        c = _oracle_case(tb, om, y, forc, np.zeros(0), np.zeros((3, tb["nelem"])))
        check_case(model, tb, c, f"noriver {nx}x{ny}")
        model.close()


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("mode", ["unique", "single"])
def test_rhs_class_dictionary_extremes(fbr, mode):
    """The soil / land-cover / geology columns travel as a class dictionary (rhs.cuh CC_*).
    Degenerate tables must give the same answers: every element its own class (nothing
    repeats), and one class for the whole mesh."""
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    ef = tb["elem_f64"]
    ne = tb["nelem"]
    cls_cols = [W.E_KSATH, W.E_KSATV, W.E_KINFV, W.E_ALPHA, W.E_BETA, W.E_POROSITY, W.E_KMACH,
                W.E_KMACV, W.E_AREAFV, W.E_AREAFH, W.E_ROUGH, W.E_RZD]
    if fbr:
        cls_cols += [W.E_GKSATH, W.E_GKSATV, W.E_GALPHA, W.E_GBETA, W.E_GPOROSITY]
    if mode == "unique":
        wiggle = 1.0 + 1e-7 * np.arange(ne)
        for c in (W.E_KSATV, W.E_BETA, W.E_ROUGH):
            ef[c] = ef[c] * wiggle
    else:
        for c in cls_cols:
            ef[c] = ef[c][0]
    om = oraclelib.OracleModel(tb)
    model = lib.Model(tb)
    y = W.wet_state(tb, seed=21)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    c = _oracle_case(tb, om, y, forc, np.zeros(tb["nriver"]), np.zeros((3, ne)))
    worst, exact = check_case(model, tb, c, f"dictionary {mode} fbr={fbr}")
    print(f"dictionary {mode} fbr={fbr}: max rel err {worst:.2e}, bit-exact {exact:.4f}")
    model.close()


def test_rhs_nan_flag():
    """CheckDy: a NaN in dy must raise the device flag (the reference exits,
    ode.c:305-310).  A NaN *state* is clamped to 0 by `y >= 0 ? y : 0` in both
    codes (ode.c:31), so the NaN is injected through the forcing."""
    tb = W.make_named("tiny")
    model = lib.Model(tb)
    y = tb["y0"].copy()
    f = W.storm_forcing(tb, 0.0)
    f[W.F_PCPDRP, 5] = np.nan
    model.set_forcing(f)
    dy = model.ODE(0.0, y)
    assert model.nan_flag == 1 and np.isnan(dy[5])
    model.set_forcing(W.storm_forcing(tb, 0.0))
    model.ODE(0.0, y)
    assert model.nan_flag == 0
    model.close()


def test_rhs_bad_mesh_rejected():
    tb = W.make_named("tiny")
    tb["riv_i32"][W.RI_DOWN, 0] = -7          # OutletFlux: unknown code -> reference exits
    with pytest.raises(RuntimeError):
        lib.Model(tb)


@pytest.mark.parametrize("fbr", [False, True])
def test_rhs_100k_vs_oracle(fbr):
    """BASELINE config[1] size (100k triangles) against the oracle port."""
    tb = W.make_named("100k", fbr=fbr)
    ne = tb["nelem"]
    om = oraclelib.OracleModel(tb)
    model = lib.Model(tb, reorder=1)
    y = W.wet_state(tb, seed=3)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    c = _oracle_case(tb, om, y, forc, np.zeros(tb["nriver"]), np.zeros((3, ne)))
    worst, exact = check_case(model, tb, c, "100k")
    print(f"100k fbr={fbr}: max rel err {worst:.2e}; bit-exact fraction {exact:.4f}")
    record(f"RHS vs oracle, 100k fbr={fbr}", max_rel_err=worst, bit_exact_fraction=exact, bound=RTOL)
    model.close()


@pytest.mark.parametrize("fbr", [False, True])
def test_rhs_1m_vs_oracle(fbr):
    """BASELINE configs [2] and [3]: the 1M-triangle benchmark mesh (pihm and pihm-fbr) against the oracle
    port on two seeded states."""
    tb = W.make_named("1M", fbr=fbr)
    ne, nr = tb["nelem"], tb["nriver"]
    om = oraclelib.OracleModel(tb)
    model = lib.Model(tb, reorder=1)
    worst = 0.0
    for seed in (11, 12):
        y = W.wet_state(tb, seed=seed)
        forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
        c = _oracle_case(tb, om, y, forc, np.zeros(nr), np.zeros((3, ne)))
        w, exact = check_case(model, tb, c, f"1M fbr={fbr} seed {seed}")
        worst = max(worst, w)
        print(f"1M fbr={fbr} seed {seed}: max rel err {w:.2e}; bit-exact fraction {exact:.4f}; "
              f"slow-path elements {model.slow_path_count()}")
    record(f"RHS vs oracle, 1M fbr={fbr}", max_rel_err=worst, bound=RTOL)
    assert worst <= RTOL
    model.close()


def test_rhs_8m_partition_vs_oracle():
    """BASELINE config[4]: one rank's share of the 8M-triangle mesh split over 8 GPUs (1M owned triangles +
    two rings of ghosts), ghost records as the halo exchange delivers them, against the oracle port run on
    the same local mesh (owned + ghosts as a plain mesh: its owned components are those of the global RHS)."""
    from mm_pihm_b200 import partition as PT
    tb = W.make_named("8M")
    ne, nr = tb["nelem"], tb["nriver"]
    y = W.wet_state(tb, seed=5)
    part = PT.partition(tb, 8, parts=[3])[0]
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))[:, part["elem_gid"]]
    del tb
    nl, rl = part["nelem"], part["nriver"]
    own = PT.owned_in_local(part)
    yl = PT.local_state(part, y, extended=True)
    om = oraclelib.OracleModel(part)
    c = _oracle_case(part, om, yl, forc, np.zeros(rl), np.zeros((3, nl)))
    scale = dy_scale(part, forc, c["xflux"], c["rivflow"])[own]
    model = lib.Model(part)
    model.set_forcing(forc, np.zeros(rl))
    model.set_ghosts(*PT.ghost_records(part, y))
    yv = model.N_VNew(PT.local_state(part, y)); dv = model.N_VNew()
    model.ode_dev(0.0, yv, dv)
    dy = dv.download()
    assert model.check_nan() == 0
    err = rel_err(dy, c["dy"][own], scale)
    print(f"8M/8 partition ({part['nown_elem']} owned + {nl - part['nown_elem']} ghost elements): max rel err "
          f"{err.max():.2e}; bit-exact fraction {np.mean(dy == c['dy'][own]):.4f}")
    record("RHS vs oracle, one rank's share of the 8M mesh", max_rel_err=err.max(), bound=RTOL)
    assert err.max() <= RTOL, f"rel err {err.max():.3e} at {err.argmax()}"
    model.close()


def test_rhs_1m_properties():
    """1M triangles (BASELINE config[2]): properties that need no oracle run.
    (a) reordering is transparent: reorder=0 and reorder=1 give identical bits;
    (b) mass balance: sum over elements of area*(lateral part of dy) cancels
        between element pairs -> total lateral exchange ~ 0 up to rounding;
    (c) locality: changing one element's state only changes dy within 2 rings."""
    tb = W.make_named("1M")
    ne, nr = tb["nelem"], tb["nriver"]
    y = W.wet_state(tb, seed=11)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    outs = []
    for reorder in (0, 1):
        model = lib.Model(tb, reorder=reorder)
        model.set_forcing(forc, np.zeros(nr))
        model.set_flux_recording(True)
        dy = model.ODE(0.0, y)
        xf, rf = model.get_fluxes()
        outs.append((dy, xf, rf))
        if reorder == 0:
            y2 = y.copy()
            k = ne // 2 + 17
            y2[k] += 0.01
            model.set_stale_ovlflow(np.zeros((3, ne)))
            dyb = model.ODE(0.0, y2)
            model.set_stale_ovlflow(np.zeros((3, ne)))
            dya = model.ODE(0.0, y)
            changed = np.nonzero(dya[:ne] != dyb[:ne])[0]
            assert 1 <= len(changed) <= 10            # self + <=3 + <=6
            nab = tb["elem_i32"][:3]
            ring1 = set(int(v) - 1 for v in nab[:, k] if v > 0)
            ring2 = set(int(v) - 1 for e in ring1 for v in nab[:, e] if v > 0)
            assert set(changed.tolist()) <= ({k} | ring1 | ring2)
        model.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])
    dy, xf, rf = outs[0]
    assert not np.isnan(dy).any()
    # subsurface exchange between element pairs is antisymmetric: Q_ij = -Q_ji
    nab = tb["elem_i32"][:3]
    for j in range(3):
        sel = np.nonzero(nab[j] > 0)[0]
        n = nab[j, sel] - 1
        # find the edge slot of i in n
        back = np.full(len(sel), -1)
        for jj in range(3):
            back = np.where(nab[jj, n] == sel + 1, jj, back)
        q_ij = xf[W.X_SUB0 + j, sel]
        q_ji = xf[W.X_SUB0 + back, n]
        assert np.allclose(q_ij, -q_ji, rtol=1e-13, atol=0)


@pytest.mark.parametrize("fbr", [False, True])
def test_rhs_of_a_randomly_numbered_mesh_is_bitwise_the_same(fbr):
    """A real .mesh file numbers its triangles in the order its mesh generator left them.  The 100k watershed with
    its elements renumbered at random (tests/test_reorder.py: the patch ordering brings nine neighbour pairs in ten
    back into one 128-element patch) must give, element for element, the bits of the structured numbering -- with
    the locality ordering on and off, first and second call (river-edge flows of the previous call)."""
    from test_reorder import shuffle_elements
    tb = W.make_named("100k", fbr=fbr)
    ne, nr = tb["nelem"], tb["nriver"]
    sh, new_of_old = shuffle_elements(tb, seed=5)
    y = W.wet_state(tb, seed=21)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    # state and forcing of the renumbered mesh: element blocks permuted, river blocks as they are
    old_of_new = np.argsort(new_of_old)
    ys = y.copy()
    for b in range(5 if fbr else 3):
        off = b * ne if b < 3 else 3 * ne + 2 * nr + (b - 3) * ne
        ys[off:off + ne] = y[off:off + ne][old_of_new]
    forc_s = forc[:, old_of_new]
    ref_model = lib.Model(tb, reorder=1)
    ref_model.set_forcing(forc, np.zeros(nr))
    want = [ref_model.ODE(0.0, y), ref_model.ODE(0.0, y)]
    ref_model.close()
    for reorder in (1, 0):
        model = lib.Model(sh, reorder=reorder)
        model.set_forcing(np.ascontiguousarray(forc_s), np.zeros(nr))
        for call in range(2):
            dy = model.ODE(0.0, ys)
            back = dy.copy()
            for b in range(5 if fbr else 3):
                off = b * ne if b < 3 else 3 * ne + 2 * nr + (b - 3) * ne
                back[off:off + ne] = dy[off:off + ne][new_of_old]
            assert np.array_equal(back, want[call]), \
                f"reorder={reorder} call {call}: {np.abs(back - want[call]).max():.3e}"
        assert model.nan_flag == 0
        model.close()
