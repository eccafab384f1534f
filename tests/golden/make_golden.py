"""Generate the golden vectors in tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Every array written here comes out of the unmodified reference code
(oracle/_ref/libpihm*_ref.so): the reference's ReadAlloc/Initialize on its own
input/example project, its ODE() (src/ode.c:3) and its CVODE run
(SetCVodeParam/SolveCVode/Summary, src/pihm.c:3-134).  The files are small and
committed, so that the oracle port and the CUDA path stay pinned to the
reference on machines where /root/reference does not exist (the GPU box).

  example_pihm.npz / example_fbr.npz
      tables of input/example after Initialize(); K RHS known-answer cases
      (y, forcing, river bc, stale wf.ovlflow, dy, element fluxes, rivflow);
      a CVODE trajectory: y after model steps 1, 5, 15, 30, 45, 60 with the
      forcing table of each etstep and the CVODE counters.
  synth_small_{pihm,fbr}.npz
      same for watershed.make_named('small', dirichlet_edges=True) driven by
      watershed.storm_forcing: RHS cases on wet/branch-rich states and a
      2-hour trajectory through the rain pulse.
  summary_small_{pihm,fbr}.npz
      Summary() + MassBalance() (src/update.c:3-160) known-answer cases on the
      same synthetic watershed: ws0, the RHS call that leaves the wf.* fields
      behind, the new state y -> wf.infil / wf.fbr_infil after Summary and the
      new ws0 (two consecutive model steps per case).
  et_example.npz
      IntcpSnowEt (src/is_sm_et.c) with the per-element forcing assignments of
      ApplyMeteoForc/ApplyLai on input/example: the static columns, the
      reference's own ApplyForc + IntcpSnowEt calls of the first simulated hour,
      and calls on made-up station values / canopy and snow storages / months
      that reach the other branches (rain, melt, lai = 0, full canopy ...).
  nvec_serial.npz
      outputs of nvector_serial.c for the special-case table of N_VLinearSum /
      N_VScale and the four reductions.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import reflib  # noqa: E402
import mm_pihm_b200  # noqa: E402,F401
from mm_pihm_b200 import watershed as W  # noqa: E402

REF_ROOT = "/root/reference"
TABLE_KEYS = ("nelem", "nriver", "fbr", "surf_mode", "riv_mode", "stepsize",
              "elem_f64", "elem_i32", "riv_f64", "riv_i32")


def rhs_case(m, y, forc, rivbc, stale):
    """Run the reference ODE() on (y, forcing, hidden state) and collect outputs."""
    m.set_forcing(forc, rivbc)
    m.set_ovlflow(stale)
    dy = m.ode(y)
    xf, rf = m.get_fluxes()
    return dict(y=y.copy(), forc=forc.copy(), rivbc=np.array(rivbc, float), stale=stale.copy(),
                dy=dy, xflux=xf, rivflow=rf)


def pack_cases(cases, prefix="rhs"):
    out = {f"{prefix}_n": np.int64(len(cases))}
    for k, c in enumerate(cases):
        for key, v in c.items():
            out[f"{prefix}{k}_{key}"] = v
    return out


def example(fbr: bool):
    m = reflib.RefModel(fbr=fbr).open_project(REF_ROOT, "example")
    tb = m.pack_tables()
    ctrl = m.ctrl()
    out = {k: tb[k] for k in TABLE_KEYS}
    out["ctrl_reltol"], out["ctrl_abstol"], out["ctrl_initstep"] = ctrl["reltol"], ctrl["abstol"], ctrl["initstep"]
    out["y0"] = m.get_y()
    ne = m.nelem
    rng = np.random.default_rng(2024)
    cases = []
    # trajectory with the reference's own forcing (ApplyForc + IntcpSnowEt)
    m.set_cvode_param()
    snaps = {1, 5, 15, 30, 45, 60}
    traj_steps, traj_y, traj_stats, forc_steps, forc_tabs = [], [], [], [], []
    for k in range(60):
        m.apply_forcing(k)
        f, rb = m.get_forcing()
        if k % 15 == 0:
            forc_steps.append(k)
            forc_tabs.append(f.copy())
        if k in (0, 7, 31, 59):
            # RHS known-answer case on a perturbed copy of the current state;
            # restore the hidden state afterwards so the trajectory is untouched
            stale = m.get_ovlflow()
            y = m.get_y()
            yp = y * (1.0 + 0.02 * rng.standard_normal(y.shape))
            if k == 7:      # ponded surface / wet river case
                yp[:ne] = np.abs(yp[:ne]) + rng.uniform(0, 3e-4, ne)
                yp[3 * ne:3 * ne + m.nriver] += rng.uniform(0, 0.3, m.nriver)
            st = stale + (rng.standard_normal(stale.shape) * 1e-6 if k else 0.0)
            cases.append(rhs_case(m, yp, f, rb, st))
            m.set_forcing(f, rb)
            m.set_ovlflow(stale)
            # ODE() overwrote elem.ws and river fluxes; CVODE re-evaluates them
        m.model_step(k, skip_forcing=True)
        if k + 1 in snaps:
            traj_steps.append(k + 1)
            traj_y.append(m.get_y())
            s = m.stats()
            traj_stats.append([s[key] for key in ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")])
    out.update(pack_cases(cases))
    out["traj_steps"] = np.array(traj_steps)
    out["traj_y"] = np.array(traj_y)
    out["traj_stats"] = np.array(traj_stats)
    out["forc_steps"] = np.array(forc_steps)
    out["forc_tabs"] = np.array(forc_tabs)
    # the reference's own sensitivity to round-off: same run from an initial
    # state perturbed by 1e-15 relative (calibrates the long-horizon tolerance)
    m2 = reflib.RefModel(fbr=fbr).open_project(REF_ROOT, "example")
    m2.init_state(out["y0"] * (1.0 + 1e-15 * np.random.default_rng(1).standard_normal(out["y0"].shape)))
    m2.set_cvode_param()
    pert = []
    for k in range(60):
        m2.model_step(k)
        if k + 1 in snaps:
            pert.append(m2.get_y())
    out["traj_y_pert"] = np.array(pert)
    m2.close()
    name = "example_fbr.npz" if fbr else "example_pihm.npz"
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "cases", len(cases), "traj", traj_steps, traj_stats[-1])
    m.close()


def synth(fbr: bool):
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    m = reflib.RefModel(fbr=fbr).create_from_tables(tb)
    ne, nr = tb["nelem"], tb["nriver"]
    rng = np.random.default_rng(99)
    cases = []
    for k, (seed, t) in enumerate([(7, 3 * 3600.0), (9, 0.0)]):
        y = W.wet_state(tb, seed=seed, ponded_frac=0.2 + 0.3 * k)
        forc = W.storm_forcing(tb, t, ws0_surf=np.maximum(y[:ne], 0.0) * rng.uniform(0.5, 1.5, ne))
        rivbc = np.zeros(nr)
        stale = np.zeros((3, ne))
        cases.append(rhs_case(m, y, forc, rivbc, stale))
        # second call on the same state: exercises the stale river-edge ovlflow
        stale2 = m.get_ovlflow()
        cases.append(rhs_case(m, y * (1 + 1e-3 * rng.standard_normal(y.shape)), forc, rivbc, stale2))
    out = pack_cases(cases)
    # trajectory: RelaxIc start, 2 h through the rain pulse (storm starts at 1 h)
    m.init_state(tb["y0"])
    m.set_ovlflow(np.zeros((3, ne)))
    m.set_cvode_param()
    traj_steps, traj_y, traj_stats = [], [], []
    f = None
    for k in range(120):
        ws = m.get_ws()
        if k % 15 == 0:
            f = W.storm_forcing(tb, k * 60.0)
        f[W.F_WS0SURF] = ws[:ne]
        m.set_forcing(f, np.zeros(nr))
        m.model_step(k)
        if (k + 1) in (1, 15, 60, 90, 120):
            traj_steps.append(k + 1)
            traj_y.append(m.get_y())
            s = m.stats()
            traj_stats.append([s[key] for key in ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")])
    out["traj_steps"] = np.array(traj_steps)
    out["traj_y"] = np.array(traj_y)
    out["traj_stats"] = np.array(traj_stats)
    m2 = reflib.RefModel(fbr=fbr).create_from_tables(tb)
    m2.init_state(tb["y0"] * (1.0 + 1e-15 * np.random.default_rng(1).standard_normal(tb["y0"].shape)))
    m2.set_ovlflow(np.zeros((3, ne)))
    m2.set_cvode_param()
    pert = []
    for k in range(120):
        ws = m2.get_ws()
        if k % 15 == 0:
            f = W.storm_forcing(tb, k * 60.0)
        f[W.F_WS0SURF] = ws[:ne]
        m2.set_forcing(f, np.zeros(nr))
        m2.model_step(k)
        if (k + 1) in (1, 15, 60, 90, 120):
            pert.append(m2.get_y())
    out["traj_y_pert"] = np.array(pert)
    m2.close()
    name = "synth_small_fbr.npz" if fbr else "synth_small_pihm.npz"
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "cases", len(cases), "traj", traj_steps, traj_stats[-1])
    m.close()


def summary(fbr: bool):
    tb = W.make_named("small", fbr=fbr, dirichlet_edges=True)
    m = reflib.RefModel(fbr=fbr).create_from_tables(tb)
    ne, nr = tb["nelem"], tb["nriver"]
    depth = tb["elem_f64"][W.E_DEPTH]
    rng = np.random.default_rng(123 + fbr)
    cases = []
    for k, (seed, t) in enumerate([(11, 3 * 3600.0), (13, 600.0)]):
        y0 = W.wet_state(tb, seed=seed, ponded_frac=0.3)
        m.init_state(y0)                       # ws = ws0 = y0 (initialize.c:598,612)
        m.set_ovlflow(np.zeros((3, ne)))
        c = dict(ws0=y0.copy())
        y_prev = y0
        for step in range(2):
            # the RHS call whose fluxes Summary sees: a state near the step's end state
            y_rhs = y_prev * (1 + 1e-3 * rng.standard_normal(y_prev.shape))
            forc = W.storm_forcing(tb, t + 60.0 * step, ws0_surf=np.maximum(y_prev[:ne], 0.0))
            stale = m.get_ovlflow()
            m.set_forcing(forc, np.zeros(nr))
            dy = m.ode(y_rhs)
            xf_rhs, rf = m.get_fluxes()
            # end-of-step state: drift by dy*60 s, plus cases that hit the clamps of
            # MassBalance (storage above the soil depth, negative storage, infil < 0)
            y_new = y_rhs + 60.0 * dy
            sel = rng.permutation(ne)
            y_new[2 * ne + sel[:ne // 10]] = depth[sel[:ne // 10]] * 1.05       # gw + unsat > depth
            y_new[ne + sel[ne // 10:ne // 5]] = -1e-3                            # negative unsat
            y_new[2 * ne + sel[ne // 5:3 * ne // 10]] *= 0.9                     # storage drops: infil < 0
            m.summary(y_new)
            xf_sum, _ = m.get_fluxes()
            c.update({f"s{step}_y_rhs": y_rhs, f"s{step}_forc": forc, f"s{step}_stale": stale,
                      f"s{step}_infil_rhs": xf_rhs[[W.X_INFIL, W.X_FBR_INFIL]], f"s{step}_y_new": y_new,
                      f"s{step}_xflux_sum": xf_sum, f"s{step}_ws0": m.get_ws0()})
            y_prev = y_new
        cases.append(c)
    out = pack_cases(cases, prefix="sum")
    # the wf.* fields after SolveCVode + Summary along the reference's own run (the quiet first
    # 15 model steps of synth()'s trajectory, where both integrators walk in lock step)
    m.init_state(tb["y0"])
    m.set_ovlflow(np.zeros((3, ne)))
    m.set_cvode_param()
    tx, ty, ts = [], [], []
    for k in range(15):
        ws = m.get_ws()
        if k % 15 == 0:
            f = W.storm_forcing(tb, k * 60.0)
        f[W.F_WS0SURF] = ws[:ne]
        m.set_forcing(f, np.zeros(nr))
        m.model_step(k)
        if (k + 1) in (1, 15):
            tx.append(m.get_fluxes()[0]); ty.append(m.get_y())
            s = m.stats()
            ts.append([s[key] for key in ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS")])
    out["traj_steps"] = np.array([1, 15]); out["traj_xflux"] = np.array(tx)
    out["traj_y"] = np.array(ty); out["traj_stats"] = np.array(ts)
    name = "summary_small_fbr.npz" if fbr else "summary_small_pihm.npz"
    np.savez_compressed(os.path.join(HERE, name), **out)
    neg = [(c[f"s{s_}_xflux_sum"][W.X_INFIL] == 0).mean() for c in cases for s_ in range(2)]
    print(name, "cases", len(cases), "share of infil clipped at 0:", np.round(neg, 3))
    m.close()


def et():
    m = reflib.RefModel(fbr=False).open_project(REF_ROOT, "example")
    ne = m.nelem
    d = m.et_dims()
    etf, eti = m.pack_et_tables()
    tb = m.pack_tables()
    out = {k: tb[k] for k in TABLE_KEYS}
    out.update(et_f64=etf, et_i32=eti, cal=m.et_get_cal(), etstep=np.float64(d["etstep"]))
    cases = []

    def record(t, state_in, y, meteo, lai, res):
        lai_lc, z0_lc, meltf = m.et_monthly(t, d["nlc"])
        cases.append(dict(t=np.int64(t), state_in=state_in, y=y.copy(), meteo=meteo.copy(), lai=np.array(lai, float),
                          lai_lc=lai_lc, z0_lc=z0_lc, meltf=np.float64(meltf), out=res))

    # (a) the reference's own calls: ApplyForc + IntcpSnowEt at the etsteps of the first hour,
    #     with the model advancing in between (elem.ws changes, snow accumulates)
    m.set_cvode_param()
    for k in range(60):
        if k % 15 == 0:
            state_in = m.et_get()[[W.EO_SNEQV, W.EO_CMC]]
            y = m.get_ws()
            m.apply_forcing(k)
            meteo, lai = m.et_get_forc()
            record(m.tout(k), state_in, y, meteo, lai, m.et_get())
        m.model_step(k, skip_forcing=True)
    # (b) made-up inputs through the same IntcpSnowEt
    rng = np.random.default_rng(77)
    y0 = m.get_ws()
    depth = tb["elem_f64"][W.E_DEPTH]
    month_t = [1230768000 + int(86400 * 30.5 * k) for k in range(12)]
    for k in range(14):
        t = month_t[k % 12] + 3600 * int(rng.integers(0, 24))
        temp_c = [-8.0, -2.0, 0.5, 3.0, 12.0, 25.0, 33.0][k % 7] + rng.uniform(-0.4, 0.4)
        meteo = np.array([[rng.choice([0.0, 0.8, 6.0]), temp_c + 273.15, rng.uniform(25, 100), rng.uniform(0.3, 9.0),
                           rng.choice([-5.0, 0.0, 150.0, 700.0]), 300.0, 98000.0]])
        lai = [[0.0, 0.4, 2.5, 5.0][k % 4]]
        intcp_max = etf[W.ET_CMCFACTR] * lai[0] * etf[W.ET_SHDFAC]
        cmc = rng.choice([0.0, 0.3, 0.9, 1.0, 1.4], ne) * np.maximum(intcp_max, 1e-5) - 1e-6 * (rng.random(ne) < 0.05)
        sneqv = rng.choice([0.0, 1e-5, 2e-3, 0.05], ne)
        y = y0.copy()
        y[ne:2 * ne] = rng.choice([-1e-3, 0.0, 0.05, 0.3, 1.5], ne) * rng.random(ne)          # unsat
        y[2 * ne:3 * ne] = depth * rng.choice([0.0, 0.2, 0.7, 0.97, 1.02], ne)               # gw
        m.et_set_state(sneqv, cmc)
        m.et_run(t, [900.0, 3600.0][k % 2], meteo, lai, y)
        cases.append(dict(t=np.int64(t), state_in=np.array([sneqv, cmc]), y=y, meteo=meteo, lai=np.array(lai[:1]),
                          stepsize=np.float64([900.0, 3600.0][k % 2]), out=m.et_get(),
                          **dict(zip(("lai_lc", "z0_lc", "meltf"), m.et_monthly(t, d["nlc"])))))
        cases[-1]["meltf"] = np.float64(cases[-1]["meltf"])
    for c in cases:
        c.setdefault("stepsize", np.float64(d["etstep"]))
    out.update(pack_cases(cases, prefix="et"))
    # (c) mixed types: half of the elements take their LAI from the monthly table of a random
    #     land-cover class (lai_type 0), every element gets a random class for its roughness length
    eti_b = eti.copy()
    eti_b[W.ETI_LAI_TYPE] = np.where(rng.random(ne) < 0.5, 0, eti[W.ETI_LAI_TYPE])
    # classes with a roughness length in every month (the table rows of unused classes are zero and
    # make the reference itself divide by zero)
    valid = np.array([c + 1 for c in range(d["nlc"]) if all(m.et_monthly(tt, d["nlc"])[1][c] > 0 for tt in month_t)])
    eti_b[W.ETI_LC_TYPE] = rng.choice(valid, ne)
    m.et_set_types(eti_b)
    out["et_i32_b"] = eti_b
    cases_b = []
    for k in range(6):
        t = month_t[(2 * k + 1) % 12] + 3600 * int(rng.integers(0, 24))
        temp_c = [-5.0, 2.0, 9.0, 18.0, 27.0, 0.8][k] + rng.uniform(-0.4, 0.4)
        meteo = np.array([[rng.choice([0.0, 1.5, 9.0]), temp_c + 273.15, rng.uniform(25, 100), rng.uniform(0.3, 9.0),
                           rng.choice([0.0, 250.0, 800.0]), 300.0, 98000.0]])
        lai = [[0.0, 1.2, 3.5][k % 3]]
        cmc = rng.uniform(-1e-6, 9e-4, ne)
        sneqv = rng.choice([0.0, 2e-3, 0.05], ne)
        y = y0.copy()
        y[ne:2 * ne] = rng.choice([-1e-3, 0.0, 0.05, 0.3, 1.5], ne) * rng.random(ne)
        y[2 * ne:3 * ne] = depth * rng.choice([0.0, 0.2, 0.7, 0.97, 1.02], ne)
        m.et_set_state(sneqv, cmc)
        m.et_run(t, 900.0, meteo, lai, y)
        lai_lc, z0_lc, meltf = m.et_monthly(t, d["nlc"])
        cases_b.append(dict(t=np.int64(t), state_in=np.array([sneqv, cmc]), y=y, meteo=meteo, lai=np.array(lai[:1]),
                            stepsize=np.float64(900.0), out=m.et_get(), lai_lc=lai_lc, z0_lc=z0_lc,
                            meltf=np.float64(meltf)))
    out.update(pack_cases(cases_b, prefix="etb"))
    m.et_set_types(eti)
    np.savez_compressed(os.path.join(HERE, "et_example.npz"), **out)
    o = np.array([c["out"] for c in cases])
    print("et_example.npz cases", len(cases), "pcpdrp>0:", (o[:, W.EO_PCPDRP] > 0).mean().round(3),
          "ett>0:", (o[:, W.EO_ETT] > 0).mean().round(3), "snow>0:", (o[:, W.EO_SNEQV] > 0).mean().round(3),
          "cmc at max:", np.mean([np.mean(c["out"][W.EO_CMC] >= etf[W.ET_CMCFACTR] * c["lai"][0] * etf[W.ET_SHDFAC])
                                  for c in cases[4:]]).round(3))
    m.close()


def nvec():
    m = reflib.RefModel(fbr=False)
    rng = np.random.default_rng(5)
    n = 1000
    x = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
    y = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
    w = np.abs(rng.standard_normal(n)) + 0.1
    out = dict(x=x, y=y, w=w)
    coeffs = [(1.0, 1.0), (1.0, -1.0), (-1.0, 1.0), (1.0, 0.37), (2.5, 1.0), (-1.0, 0.41),
              (3.3, -1.0), (0.77, 0.77), (0.77, -0.77), (1.3e-3, -4.7), (-1.0, -1.0)]
    out["ls_coeffs"] = np.array(coeffs)
    out["ls_z"] = np.array([m.nvec_op(0, a=a, x=x, b=b, y=y)[1] for a, b in coeffs])
    scales = [1.0, -1.0, 0.3, -2.5e-4]
    out["sc_c"] = np.array(scales)
    out["sc_z"] = np.array([m.nvec_op(4, a=c, x=x, y=y)[1] for c in scales])
    out["prod"] = m.nvec_op(2, x=x, y=y)[1]
    out["div"] = m.nvec_op(3, x=x, y=w)[1]
    out["abs"] = m.nvec_op(5, x=x, y=y)[1]
    out["inv"] = m.nvec_op(6, x=w, y=y)[1]
    out["addconst"] = m.nvec_op(7, x=x, b=0.125, y=y)[1]
    out["dot"] = m.nvec_op(8, x=x, y=y)[0]
    out["maxnorm"] = m.nvec_op(9, x=x, y=y)[0]
    out["wrms"] = m.nvec_op(10, x=x, y=w)[0]
    out["min"] = m.nvec_op(11, x=x, y=y)[0]
    np.savez_compressed(os.path.join(HERE, "nvec_serial.npz"), **out)
    print("nvec_serial.npz", out["dot"], out["wrms"])


if __name__ == "__main__":
    if sys.argv[1:] == ["et"]:          # only the file added for SURVEY 8(f) f2
        et()
        sys.exit(0)
    if sys.argv[1:] == ["summary"]:     # only the files added for SURVEY 8(f) f1
        summary(False)
        summary(True)
        sys.exit(0)
    example(False)
    example(True)
    synth(False)
    synth(True)
    nvec()
    summary(False)
    summary(True)
    et()
