"""Summary() + MassBalance() on the device (src/update.c:3-160)
through the C ABI: pihm_b200_set_diagnostics / set_ws0 / summary_mb /
get_fluxes / get_summary.

  * known-answer cases written by the reference (tests/golden/summary_small_*.npz):
    flux columns within the RHS bound (1e-12 of the flux scale), the
    mass-balance wf.infil / wf.fbr_infil within 1e-12 of the sum of their
    terms, ws0 bit for bit, subrunoff against the oracle port;
  * the fluxes "of the last ODE() call"  are produced by
    evaluating that call once more: bitwise equal to recording every call, the
    hidden state of the RHS untouched (identical trajectories), on both
    integrator routes (own integrator / reference CVODE on our N_Vector);
  * along the reference's own CVODE run (golden, and live through oracle/_ref
    where it travelled to the box) while the integrators walk in lock step.
"""
import numpy as np
import pytest

import oraclelib
from helpers import load_golden, summary_cases
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu
RELTOL, ABSTOL = 1e-3, 1e-4
STAT_KEYS = ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")
FLUX_COLS = [c for c in range(W.X_NCOL) if c not in (W.X_INFIL, W.X_FBR_INFIL)]


def mb_scale(tb, xf, y0, y1):
    """sum of |terms| of update.c:135 (and :148 for fbr): the 1e-12 bound is relative to it"""
    ne, nr = tb["nelem"], tb["nriver"]
    ef = tb["elem_f64"]
    area, depth = ef[W.E_AREA], ef[W.E_DEPTH]
    sw = lambda y: np.clip(y[2 * ne:3 * ne] + y[ne:2 * ne], 0.0, depth)
    s = (np.abs(sw(y1)) + np.abs(sw(y0))) * ef[W.E_POROSITY] / tb["stepsize"]
    s = s + np.abs(xf[W.X_SUB0:W.X_SUB0 + 3]).sum(0) / area
    s = s + np.abs(xf[[W.X_EDIR_UNSAT, W.X_EDIR_GW, W.X_ETT_UNSAT, W.X_ETT_GW]]).sum(0)
    sf = np.zeros(ne)
    if tb["fbr"]:
        o = 3 * ne + 2 * nr
        fw = lambda y: np.abs(y[o + ne:o + 2 * ne]) + np.abs(y[o:o + ne])
        sf = (fw(y1) + fw(y0)) * ef[W.E_GPOROSITY] / tb["stepsize"]
        sf = sf + np.abs(xf[W.X_FBRFLOW0:W.X_FBRFLOW0 + 3]).sum(0) / area
    return s + sf, sf


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("eager", [False, True])
def test_summary_matches_golden(fbr, eager):
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    cases = summary_cases(load_golden("summary_small_fbr.npz" if fbr else "summary_small_pihm.npz"))
    for ci, c in enumerate(cases):
        model = lib.Model(tb, reorder=1)
        model.set_diagnostics(True)
        if eager:
            model.set_flux_recording(True)
        om = oraclelib.OracleModel(tb)
        v = model.N_VNew(c["ws0"])
        model.set_ws0(v)
        om.set_ws0(c["ws0"])
        y_prev = c["ws0"]
        for si, s in enumerate(c["steps"]):
            model.set_forcing(s["forc"], np.zeros(nr))
            model.set_stale_ovlflow(s["stale"])
            model.ODE(0.0, s["y_rhs"])
            v.upload(s["y_new"])
            model.SummaryMB(v, tb["stepsize"])
            xf, _ = model.get_fluxes()
            sr, ws0 = model.get_summary()
            ref = s["xflux_sum"]
            # flux columns of the last RHS call: the RHS bound
            for col in FLUX_COLS:
                scale = np.abs(ref[col]).max()
                if scale > 0:
                    assert np.abs(xf[col] - ref[col]).max() <= 1e-12 * scale, (ci, si, col)
            # mass balance
            scale, scale_f = mb_scale(tb, ref, y_prev, s["y_new"])
            e = np.abs(xf[W.X_INFIL] - ref[W.X_INFIL]) / scale
            print(f"fbr={fbr} eager={eager} case {ci} step {si}: infil err {e.max():.2e} of the term sum, "
                  f"bit-exact {np.mean(xf[W.X_INFIL] == ref[W.X_INFIL]):.3f}")
            assert e.max() <= 1e-12
            if fbr:
                assert (np.abs(xf[W.X_FBR_INFIL] - ref[W.X_FBR_INFIL]) <= 1e-12 * scale_f).all()
            assert np.array_equal(ws0, s["ws0"])            # ws0 = y, bit for bit
            # subrunoff is a local of MassBalance: the oracle port has it
            om.set_forcing(s["forc"], np.zeros(nr)); om.set_stale_ovlflow(s["stale"])
            om.ode(s["y_rhs"])
            sro = om.summary(s["y_new"], tb["stepsize"])
            assert (np.abs(sr - sro) <= 1e-12 * scale).all()
            y_prev = s["y_new"]
            # the ws0.surf column of Infil() follows: next RHS sees it (golden step 1 depends on it)
        model.close(); om.close()


def run_trajectory(tb, mode, nsteps, t_start, y0=None, route=2):
    """model steps of src/pihm.c:3-134 (forcing, SolveCVode, Summary) -> per step
    (y, xflux after Summary, subrunoff, ws0, counters).  mode: 'lazy' (diagnostics only: the
    last call is evaluated again at Summary time) or 'eager' (every RHS call records)."""
    nr = tb["nriver"]
    model = lib.Model(tb, reorder=1)
    model.set_diagnostics(True)
    if mode == "eager":
        model.set_flux_recording(True)
    y = model.N_VNew(tb["y0"] if y0 is None else y0)
    model.set_ws0(y)
    cv = lib.Cvode(model)
    cv.SetCVodeParam(y)
    out = []
    for k in range(nsteps):
        if k % 15 == 0:
            model.set_forcing(W.storm_forcing(tb, t_start + k * 60.0), np.zeros(nr))
        model.Summary(y)                # ws0.surf column (set_forcing rewrites the whole table)
        cv.SolveCVode((k + 1) * 60.0, y)
        model.SummaryMB(y, tb["stepsize"])
        xf, rf = model.get_fluxes()
        sr, ws0 = model.get_summary()
        out.append(dict(y=y.download(), xf=xf, rf=rf, sr=sr, ws0=ws0, st=cv.stats()))
    cv.close(); model.close()
    return out


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("t_start", [0.0, 3600.0])
def test_lazy_replay_equals_recording(fbr, t_start):
    """quiet start (Newton often converges without a Krylov iteration: the last RHS input is
    zn[0] / y and gets overwritten before Summary) and the rain pulse (last call = DQ call)"""
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    a = run_trajectory(tb, "lazy", 40, t_start)
    b = run_trajectory(tb, "eager", 40, t_start)
    for k, (u, v) in enumerate(zip(a, b)):
        assert np.array_equal(u["y"], v["y"]), k            # the re-evaluation leaves the hidden state alone
        assert u["st"] == v["st"], k
        assert np.array_equal(u["xf"], v["xf"]), k
        assert np.array_equal(u["rf"], v["rf"]), k
        assert np.array_equal(u["sr"], v["sr"]) and np.array_equal(u["ws0"], v["ws0"]), k
        assert np.array_equal(u["ws0"], u["y"])
    assert np.abs(a[-1]["xf"][W.X_SUB0:W.X_SUB0 + 3]).max() > 0


@pytest.mark.parametrize("fbr", [False, True])
def test_summary_along_reference_trajectory_golden(fbr):
    """wf.* after SolveCVode + Summary of the reference's own run, steps 1 and 15 (lock step)."""
    g = load_golden("summary_small_fbr.npz" if fbr else "summary_small_pihm.npz")
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    run = run_trajectory(tb, "lazy", 15, 0.0)
    for step, xref, yref, sref in zip(g["traj_steps"], g["traj_xflux"], g["traj_y"], g["traj_stats"]):
        r = run[int(step) - 1]
        assert [int(r["st"][k]) for k in STAT_KEYS] == [int(v) for v in sref], f"step {step}: not in lock step"
        check_flux_lockstep(tb, r, xref, yref, f"golden fbr={fbr} step {step}")


def check_flux_lockstep(tb, r, xref, yref, tag):
    """Same step sequence => the states agree to <= 1e-6 (reltol|y|+abstol) (test_cvode_gpu) and
    so does the input of the last RHS call; fluxes are Lipschitz in it: 1e-5 of the column's
    largest entry, the storage term of the mass balance 4 state errors * porosity / stepsize."""
    unit = RELTOL * np.abs(yref) + ABSTOL
    yerr = (np.abs(r["y"] - yref) / unit).max()
    assert yerr <= 1e-6, f"{tag}: state error {yerr:.2e}"
    worst = 0.0
    for col in FLUX_COLS:
        scale = np.abs(xref[col]).max()
        if scale > 0:
            e = np.abs(r["xf"][col] - xref[col]).max() / scale
            worst = max(worst, e)
            assert e <= 1e-5, f"{tag}: flux column {col} differs by {e:.2e} of its scale"
    tol = 1e-5 * max(np.abs(xref[W.X_INFIL]).max(), 1e-12) + 4 * 1e-6 * unit.max() / tb["stepsize"]
    e_inf = np.abs(r["xf"][W.X_INFIL] - xref[W.X_INFIL]).max()
    print(f"{tag}: state err {yerr:.2e} units, flux columns {worst:.2e} of scale, infil {e_inf:.2e} (tol {tol:.2e})")
    assert e_inf <= tol


def test_summary_along_live_reference_rain():
    """rain pulse, 100 s of CPU at most: the reference (oracle/_ref) and the device integrator side
    by side from a wet state; compared while their counters coincide (at least the first steps)."""
    import reflib
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    tb = W.make_named("small", dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    ref = reflib.RefModel(fbr=False).create_from_tables(tb)
    ref.init_state(tb["y0"]); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
    run = run_trajectory(tb, "lazy", 12, 3600.0)
    nlock = 0
    for k in range(12):
        if k % 15 == 0:
            f = W.storm_forcing(tb, 3600.0 + k * 60.0)
        fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(fr, np.zeros(nr))
        ref.model_step(k)
        sref = ref.stats()
        if [run[k]["st"][s] for s in STAT_KEYS] != [sref[s] for s in STAT_KEYS]:
            break
        check_flux_lockstep(tb, run[k], ref.get_fluxes()[0], ref.get_y(), f"live rain step {k + 1}")
        assert np.array_equal(run[k]["ws0"], run[k]["y"])
        nlock += 1
    ref.close()
    print(f"live reference: {nlock} of 12 rain steps in lock step")
    assert nlock >= 3


def test_dropin_route_gives_same_fluxes():
    """Route 1 of INTEGRATION.md (the reference CVODE on N_VNew_PihmB200 + PihmB200_ODE): the
    write hooks of the N_Vector ops keep the last RHS input alive, same fluxes as route 2."""
    import reflib
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    from test_dropin_gpu import dropin_run
    tb = W.make_named("small", dirichlet_edges=True)
    a = run_trajectory(tb, "lazy", 10, 3600.0)
    b = dropin_run(tb, 10, 3600.0, diagnostics=True)
    for k in range(10):
        assert np.array_equal(a[k]["y"], b[k]["y"]), k
        assert np.array_equal(a[k]["xf"], b[k]["xf"]), k
        assert np.array_equal(a[k]["sr"], b[k]["sr"]), k


@pytest.mark.parametrize("fbr", [False, True])
def test_summary_ragged_sizes_and_no_river(fbr):
    """element counts that are not multiples of a tile, meshes without rivers: Summary/MassBalance,
    print accumulation of states and fluxes against the oracle port"""
    for nx, ny in ((3, 2), (7, 5), (13, 9)):
        tb = W.make_watershed(nx, ny, river=False, fbr=fbr)
        ne = tb["nelem"]
        assert tb["nriver"] == 0 and ne % 32 != 0
        om = oraclelib.OracleModel(tb)
        model = lib.Model(tb, reorder=1)
        model.set_diagnostics(True)
        rng = np.random.default_rng(nx)
        y0 = W.wet_state(tb, seed=nx)
        forc = W.storm_forcing(tb, 4 * 3600.0, ws0_surf=np.maximum(y0[:ne], 0))
        model.set_forcing(forc, np.zeros(0)); om.set_forcing(forc, np.zeros(0))
        v = model.N_VNew(y0)
        model.set_ws0(v); om.set_ws0(y0)
        pid = [model.print_add(W.PS_STATE, 2), model.print_add(W.PS_ELEM_FLUX, W.X_INFIL)]
        acc = [oraclelib.PrintVarOracle(ne), oraclelib.PrintVarOracle(ne)]
        y = y0
        for step in range(3):
            y_rhs = y * (1 + 1e-3 * rng.standard_normal(y.shape))
            dy = model.ODE(0.0, y_rhs)
            om.ode(y_rhs)
            y = y_rhs + 60.0 * dy
            v.upload(y)
            model.SummaryMB(v, tb["stepsize"])
            sro = om.summary(y, tb["stepsize"])
            xf, _ = model.get_fluxes(); xo, _ = om.get_fluxes()
            sr, ws0 = model.get_summary()
            scale, scale_f = mb_scale(tb, xo, y_rhs, y)
            for col in range(W.X_NCOL):
                cs = max(np.abs(xo[col]).max(), 1e-300)
                tol = 1e-12 * (np.maximum(scale, cs) if col in (W.X_INFIL, W.X_FBR_INFIL) else cs)
                assert (np.abs(xf[col] - xo[col]) <= tol).all(), (nx, ny, step, col)
            assert (np.abs(sr - sro) <= 1e-12 * np.maximum(scale, 1e-300)).all()
            assert np.array_equal(ws0, y) and np.array_equal(om.get_ws0(), y)
            model.UpdPrintVar(pid, v)
            acc[0].update(y[2 * ne:3 * ne]); acc[1].update(xf[W.X_INFIL])
        for p_, a_ in zip(pid, acc):
            out, n = model.PrintData(p_)
            ref, nref = a_.data()
            assert n == nref == 3 and np.array_equal(out, ref)
        model.close(); om.close()


def test_summary_1m_properties():
    """BASELINE's full size (1M triangles): size-independent properties of the device Summary --
    re-evaluating the last RHS call gives the bits of recording every call; ws0 = y; the mass balance
    closes: infil = max(0, d(storage) * porosity / dt + sum subsurf / area + ET terms) recomputed from
    the outputs; subrunoff = sum subsurf / area + max(0, -that sum)."""
    tb = W.make_named("1M")
    ne, nr = tb["nelem"], tb["nriver"]
    ef = tb["elem_f64"]
    runs = {}
    for mode in ("lazy", "eager"):
        model = lib.Model(tb, reorder=1)
        model.set_diagnostics(True)
        if mode == "eager":
            model.set_flux_recording(True)
        y = model.N_VNew(tb["y0"]); model.set_ws0(y)
        cv = lib.Cvode(model); cv.SetCVodeParam(y)
        model.set_forcing(W.storm_forcing(tb, 2 * 3600.0), np.zeros(nr))
        out = []
        y_prev = tb["y0"]
        for k in range(2):
            model.Summary(y)
            cv.SolveCVode((k + 1) * 60.0, y)
            model.SummaryMB(y, tb["stepsize"])
            xf, _ = model.get_fluxes(); sr, ws0 = model.get_summary(); yh = y.download()
            out.append((yh, xf, sr, ws0, y_prev))
            y_prev = yh
        runs[mode] = out
        cv.close(); model.close()
    for a, b in zip(runs["lazy"], runs["eager"]):
        for u, v in zip(a[:4], b[:4]):
            assert np.array_equal(u, v)
    area, depth, por = ef[W.E_AREA], ef[W.E_DEPTH], ef[W.E_POROSITY]
    for yh, xf, sr, ws0, y_prev in runs["lazy"]:
        assert np.array_equal(ws0, yh)
        sw = lambda y: np.clip(y[2 * ne:3 * ne] + y[ne:2 * ne], 0.0, depth)
        sub = xf[W.X_SUB0] / area + xf[W.X_SUB1] / area + xf[W.X_SUB2] / area
        raw = (sw(yh) - sw(y_prev)) * por / tb["stepsize"] + sub + xf[W.X_EDIR_UNSAT] + xf[W.X_EDIR_GW] \
            + xf[W.X_ETT_UNSAT] + xf[W.X_ETT_GW]
        assert np.array_equal(xf[W.X_INFIL], np.where(raw < 0.0, 0.0, raw))      # same IEEE operations in numpy
        assert np.array_equal(sr, np.where(raw < 0.0, sub - raw, sub))
        assert (xf[W.X_INFIL] > 0).any()
