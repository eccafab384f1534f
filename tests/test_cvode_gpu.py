"""Integrated-state parity: the device BDF/Newton/SPGMR integrator driven
through SetCVodeParam / SolveCVode / Summary against trajectories produced by
the reference's own CVODE run (tests/golden/*.npz, made by make_golden.py).

Acceptance.  While the two integrators walk in lock step (identical step /
iteration counters -- the first model steps) the states must agree to
round-off: <= 1e-6 * (reltol*|y| + abstol).  Over the whole simulated period
the bound is MULT * (reltol*|y_ref| + abstol) with MULT = 30: PIHM's switches
(depression storage, dinf/dmac thresholds, weir regimes) amplify last-bit
differences, and the REFERENCE ITSELF, restarted from an initial state
perturbed by 1e-15 relative, drifts from its own trajectory by up to ~14 *
(reltol*|y|+abstol) over these runs (stored as traj_y_pert by make_golden.py;
printed next to our error below).  libdevice pow differs from glibc in the
last ulp and reductions are tree sums, so bit-identical step sequences cannot
be expected beyond the lock-step phase; work counters must stay within 25 % at the end of the run."""
import numpy as np
import pytest

from helpers import golden_tables, load_golden, record
import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

pytestmark = pytest.mark.gpu
RELTOL, ABSTOL, MULT, MULT_LOCKSTEP = 1e-3, 1e-4, 30.0, 1e-6
DAY_MULT = 30.0          # one simulated day: see test_one_simulated_day
STAT_KEYS = ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")


def check_state(y, yref, tag, ypert=None, lockstep=False):
    unit = RELTOL * np.abs(yref) + ABSTOL
    err = np.abs(y - yref) / unit
    worst = err.max()
    self_sens = (np.abs(ypert - yref) / unit).max() if ypert is not None else float("nan")
    print(f"{tag}: max err {worst:.3e} x (reltol|y|+abstol); reference's own 1e-15 sensitivity {self_sens:.3e}")
    lim = MULT_LOCKSTEP if lockstep else MULT
    record("trajectory " + tag, multiple_of_reltol_y_plus_abstol=worst, reference_self_sensitivity=self_sens, bound=lim)
    assert worst <= lim, f"{tag}: state error {worst:.3e} x (reltol|y|+abstol) at {np.argmax(err)}"
    return worst


def check_stats(st, ref_row, tag, tol=0.25, lockstep=False, final=True):
    """lock-step phase: identical counters.  End of the run: work counters
    (nst, nfe, nni, nli) within 25 %.  Failure counters (ncfn, ncfl) are rare
    events of a chaotic transient: printed, not asserted."""
    if lockstep:
        assert [int(st[k]) for k in STAT_KEYS] == [int(r) for r in ref_row], f"{tag}: counters differ in lock step"
    print(tag, "counters", {k: (int(st[k]), int(r)) for k, r in zip(STAT_KEYS, ref_row)})
    if not final:
        return
    for k, r in zip(STAT_KEYS, ref_row):
        v = st[k]
        if k in ("nst", "nfe", "nni", "nli", "nfeLS") and r >= 50:
            assert abs(v - r) <= tol * r, f"{tag}: counter {k} = {v}, reference {r}"


@pytest.mark.parametrize("name", ["example_pihm.npz", "example_fbr.npz"])
@pytest.mark.parametrize("reorder", [0, 1])
def test_example_trajectory(name, reorder):
    """BASELINE config[0]: the bundled input/example project, first simulated hour,
    forcing tables taken from the reference's ApplyForc/IntcpSnowEt."""
    g = load_golden(name)
    tb = golden_tables(g)
    model = lib.Model(tb, reorder=reorder)
    cv = lib.Cvode(model)
    y = model.N_VNew(g["y0"])
    forc = {int(k): f for k, f in zip(g["forc_steps"], g["forc_tabs"])}
    cv.SetCVodeParam(y, reltol=float(g["ctrl_reltol"]), abstol=float(g["ctrl_abstol"]),
                     initstep=float(g["ctrl_initstep"]), stepsize=tb["stepsize"])
    snaps = {int(s): (yy, st, yp) for s, yy, st, yp in
             zip(g["traj_steps"], g["traj_y"], g["traj_stats"], g["traj_y_pert"])}
    for k in range(60):
        if k in forc:
            f = forc[k].copy()
            model.set_forcing(f, np.zeros(tb["nriver"]))
        model.Summary(y)                              # ws0.surf of this step = y_surf now
        t = cv.SolveCVode((k + 1) * 60.0, y)
        assert t == (k + 1) * 60.0
        if k + 1 in snaps:
            yref, sref, ypert = snaps[k + 1]
            check_state(y.download(), yref, f"{name} reorder={reorder} step {k + 1}", ypert)
            check_stats(cv.stats(), sref, f"{name} step {k + 1}", final=(k + 1 == 60))
    cv.close(); model.close()


@pytest.mark.parametrize("fbr", [False, True])
def test_synthetic_trajectory(fbr):
    """2400-triangle synthetic watershed, 2 simulated hours through the rain pulse."""
    g = load_golden("synth_small_fbr.npz" if fbr else "synth_small_pihm.npz")
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    model = lib.Model(tb, reorder=1)
    cv = lib.Cvode(model)
    y = model.N_VNew(tb["y0"])
    cv.SetCVodeParam(y)
    snaps = {int(s): (yy, st, yp) for s, yy, st, yp in
             zip(g["traj_steps"], g["traj_y"], g["traj_stats"], g["traj_y_pert"])}
    for k in range(120):
        if k % 15 == 0:
            model.set_forcing(W.storm_forcing(tb, k * 60.0), np.zeros(tb["nriver"]))
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
        if k + 1 in snaps:
            yref, sref, ypert = snaps[k + 1]
            lock = (k + 1) <= 15          # both builds still take identical steps here
            check_state(y.download(), yref, f"synth fbr={fbr} step {k + 1}", ypert, lockstep=lock)
            check_stats(cv.stats(), sref, f"synth fbr={fbr} step {k + 1}", lockstep=lock, final=(k + 1 == 120))
    cv.close(); model.close()


def test_reinit_and_max_step_controller():
    """SetCVodeParam twice (spin-up re-init, ode.c:351-360) and AdjCVodeMaxStep (ode.c:500-560)"""
    tb = W.make_named("tiny")
    model = lib.Model(tb)
    cv = lib.Cvode(model)
    y = model.N_VNew(tb["y0"])
    model.set_forcing(W.storm_forcing(tb, 3 * 3600.0), np.zeros(tb["nriver"]))
    outs = []
    for rep in range(2):
        y.upload(tb["y0"])
        model.set_stale_ovlflow(np.zeros((3, tb["nelem"])))
        model.set_forcing_col(W.F_WS0SURF, np.zeros(tb["nelem"]))
        cv.SetCVodeParam(y)
        for k in range(5):
            model.Summary(y)
            cv.SolveCVode((k + 1) * 60.0, y)
            hmax = cv.AdjCVodeMaxStep()
            assert 1.0 <= hmax <= 60.0
        outs.append((y.download(), cv.stats()))
    assert np.array_equal(outs[0][0], outs[1][0])          # deterministic restart
    assert outs[0][1]["nst"] == outs[1][1]["nst"] > 0
    cv.close(); model.close()


@pytest.mark.parametrize("ctl", [dict(stmin=1.0, nncfn=0.0, nnimax=3.0, nnimin=1.0, decr=1.2, incr=1.2),
                                 dict(stmin=5.0, nncfn=0.0, nnimax=1.5, nnimin=1.2, decr=1.5, incr=1.1)])
def test_adj_cvode_max_step_vs_reference(ctl):
    """AdjCVodeMaxStep (src/ode.c:500-560) against the reference's own controller: ref_model_step(adj_max_step=1)
    calls the unmodified function after every model step; in the lock-step phase of the 2400-triangle run
    (identical counters) the two maxstep sequences must be identical, value for value.  The second parameter
    set makes the controller move in both directions."""
    import reflib
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    tb = W.make_named("small", dirichlet_edges=True)
    ne, nr = tb["nelem"], tb["nriver"]
    ref = reflib.RefModel(fbr=False).create_from_tables(tb)
    ref.set_maxstep_ctrl(**ctl)
    ref.init_state(tb["y0"]); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
    model = lib.Model(tb, reorder=1)
    cv = lib.Cvode(model)
    y = model.N_VNew(tb["y0"])
    cv.SetCVodeParam(y, **ctl)
    seq = []
    for k in range(15):
        if k % 15 == 0:
            f = W.storm_forcing(tb, k * 60.0)
            model.set_forcing(f, np.zeros(nr))
        fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(fr, np.zeros(nr))
        ref.model_step(k, adj_max_step=True)
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
        hmax = cv.AdjCVodeMaxStep()
        sr, sg = ref.stats(), cv.stats()
        assert [sg[q] for q in ("nst", "nni", "ncfn", "nfe")] == [sr[q] for q in ("nst", "nni", "ncfn", "nfe")], \
            f"step {k}: counters left lock step"
        assert hmax == ref.ctrl()["maxstep"], f"step {k}: maxstep {hmax} vs reference {ref.ctrl()['maxstep']}"
        seq.append(hmax)
    print("maxstep sequence:", seq)
    if ctl["nnimax"] < 3.0:
        assert min(seq) < 60.0, "the controller never moved"
    ref.close(); cv.close(); model.close()


def test_100k_lockstep_with_live_reference():
    """BASELINE config[1]: synthetic 100k-triangle watershed with river network on 1 B200,
    RHS + CVODE correctness against the reference run on the same box (oracle/_ref)."""
    import reflib
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    tb = W.make_named("100k")
    ne, nr = tb["nelem"], tb["nriver"]
    ref = reflib.RefModel(fbr=False, cvode_omp=False).create_from_tables(tb)
    ref.init_state(tb["y0"]); ref.set_ovlflow(np.zeros((3, ne))); ref.set_cvode_param()
    model = lib.Model(tb, reorder=1)
    cv = lib.Cvode(model)
    y = model.N_VNew(tb["y0"])
    cv.SetCVodeParam(y)
    nsteps = 12
    for k in range(nsteps):
        if k % 15 == 0:
            f = W.storm_forcing(tb, 2 * 3600.0 + k * 60.0)
            model.set_forcing(f, np.zeros(nr))
        fr = f.copy(); fr[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(fr, np.zeros(nr))
        ref.model_step(k)
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
    yr, yg = ref.get_y(), y.download()
    sr, sg = ref.stats(), cv.stats()
    unit = RELTOL * np.abs(yr) + ABSTOL
    err = (np.abs(yg - yr) / unit).max()
    print(f"100k: {nsteps} model steps, err {err:.3e} x (reltol|y|+abstol); nst {sg['nst']}/{sr['nst']} "
          f"nfe {sg['nfe']}/{sr['nfe']} nli {sg['nli']}/{sr['nli']}")
    record("100k live reference, 12 model steps", multiple_of_reltol_y_plus_abstol=err, bound=MULT,
           nst=int(sg["nst"]), nst_reference=int(sr["nst"]))
    assert err <= MULT
    assert abs(sg["nst"] - sr["nst"]) <= 0.25 * sr["nst"]
    ref.close(); cv.close(); model.close()


@pytest.mark.parametrize("size", ["small", "10k"])
def test_one_simulated_day(size):
    """BASELINE.md section 4 / SURVEY 8(d) 'Parity acceptance': the integrated state after ONE SIMULATED DAY (1440
    model steps: dry hour, 6 h storm, 17 h of recession) against the reference's own run of the same day
    (tests/golden/day_<size>.npz, made by tests/golden/make_day_golden.py from oracle/_ref in the build container),
    at the end of the storm (6 h), after 12 h and after 24 h.  2400 and 10 000 triangles: the day of the 100k
    mesh costs the reference two hours of 8 cores (make_day_golden.py); that mesh is compared with the live
    reference over its first model steps (test_100k_lockstep_with_live_reference).
    Bounds.  BASELINE.md asks for 10 x (reltol |y| + abstol).  The reference ITSELF, restarted from an initial
    state perturbed by 1e-15 relative, ends up 9.8 (2400 triangles) and 13 / 24 / 17 (10k, at 6 / 12 / 24 h) of
    these units away from its own trajectory in its worst component, 3-5.5 in the 99.9th percentile (a handful of
    elements sitting on a regime switch; stored in the golden file, recorded next to our figures).  So the 10 x
    bound is asserted for all but one component in a thousand (99.9th percentile <= 10), and the worst component
    must stay within max(DAY_MULT, 3 x the reference's own worst drift at that snapshot)  (BASELINE.md section 4,
    builder's note)."""
    import os
    from helpers import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, f"day_{size}.npz")):
        pytest.skip(f"tests/golden/day_{size}.npz not generated")
    g = load_golden(f"day_{size}.npz")
    tb = W.make_named(size, dirichlet_edges=(size == "small"))
    nr = tb["nriver"]
    model = lib.Model(tb, reorder=1)
    cv = lib.Cvode(model)
    y = model.N_VNew(tb["y0"])
    cv.SetCVodeParam(y)
    snaps = {int(s): i for i, s in enumerate(g["steps"])}
    keys = [str(k) for k in g["stat_keys"]]
    for k in range(int(max(snaps))):
        if k % 15 == 0:
            model.set_forcing(W.storm_forcing(tb, k * 60.0), np.zeros(nr))
        model.Summary(y)
        cv.SolveCVode((k + 1) * 60.0, y)
        if k + 1 in snaps:
            i = snaps[k + 1]
            yref, ypert = g["y"][i], g["y_pert"][i]
            unit = RELTOL * np.abs(yref) + ABSTOL
            err = np.abs(y.download() - yref) / unit
            sens = np.abs(ypert - yref) / unit
            st = cv.stats()
            sref = dict(zip(keys, g["stats"][i]))
            evals, evals_ref = st["nfe"] + st["nfeLS"], int(sref["nfe"] + sref["nfeLS"])
            print(f"{size}, {k + 1} model steps: max err {err.max():.3e} x (reltol|y|+abstol) (99.9th percentile "
                  f"{np.percentile(err, 99.9):.3e}, {int((err > 10.0).sum())} of {err.size} components above 10); "
                  f"reference's own 1e-15 sensitivity {sens.max():.3e}; nst {st['nst']}/{int(sref['nst'])} "
                  f"rhs evals {evals}/{evals_ref}")
            record(f"one simulated day, {size}, step {k + 1}", multiple_of_reltol_y_plus_abstol=err.max(),
                   percentile_99_9=float(np.percentile(err, 99.9)), components_above_10=int((err > 10.0).sum()),
                   reference_self_sensitivity=sens.max(), reference_self_sensitivity_99_9=float(np.percentile(sens, 99.9)),
                   bound=max(DAY_MULT, 3.0 * sens.max()), bound_99_9=10.0, nst=int(st["nst"]),
                   nst_reference=int(sref["nst"]), rhs_evals=int(evals), rhs_evals_reference=evals_ref)
            assert np.percentile(err, 99.9) <= 10.0, f"step {k + 1}: 99.9th percentile {np.percentile(err, 99.9):.3e}"
            assert err.max() <= max(DAY_MULT, 3.0 * sens.max()), \
                f"step {k + 1}: {err.max():.3e} x (reltol|y|+abstol) at {np.argmax(err)}"
            assert abs(st["nst"] - sref["nst"]) <= 0.25 * sref["nst"]
    assert model.check_nan() == 0
    cv.close(); model.close()


@pytest.mark.parametrize("fbr", [False, True])
def test_mgs_chain_cooperative_launch_is_bit_identical(fbr, monkeypatch):
    """PIHM_B200_MGS_CHAIN=1: the Gram-Schmidt chain of a Krylov iteration as one cooperative launch
    (k_mgs_chain, device-side ticket gate between the steps) against one k_mgs_step launch per step."""
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    outs = []
    for chain in ("0", "1"):
        monkeypatch.setenv("PIHM_B200_MGS_CHAIN", chain)
        model = lib.Model(tb, reorder=1)
        cv = lib.Cvode(model)                    # reads the switch
        y = model.N_VNew(tb["y0"])
        cv.SetCVodeParam(y)
        launches0 = model.launches
        for k in range(25):
            if k % 15 == 0:
                model.set_forcing(W.storm_forcing(tb, 3600.0 + k * 60.0), np.zeros(tb["nriver"]))
            model.Summary(y)
            cv.SolveCVode((k + 1) * 60.0, y)
        outs.append((y.download(), cv.stats(), model.launches - launches0))
        cv.close(); model.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1] and outs[0][1]["nli"] > 50
    assert outs[1][2] < outs[0][2]               # the chain really ran: fewer launches
