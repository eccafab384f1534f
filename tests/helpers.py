"""Shared helpers of the test-suite (golden loading, parity metrics)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TABLE_KEYS = ("nelem", "nriver", "fbr", "surf_mode", "riv_mode", "stepsize",
              "elem_f64", "elem_i32", "riv_f64", "riv_i32")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_tables(g):
    tb = {k: g[k] for k in TABLE_KEYS}
    for k in ("nelem", "nriver", "fbr", "surf_mode", "riv_mode"):
        tb[k] = int(tb[k])
    tb["stepsize"] = float(tb["stepsize"])
    return tb


def golden_cases(g, prefix="rhs"):
    n = int(g[f"{prefix}_n"])
    keys = ("y", "forc", "rivbc", "stale", "dy", "xflux", "rivflow")
    return [{k: g[f"{prefix}{i}_{k}"] for k in keys} for i in range(n)]


def summary_cases(g):
    """Summary()/MassBalance() known-answer cases (make_golden.py summary)."""
    out = []
    for i in range(int(g["sum_n"])):
        steps = []
        for s in range(2):
            keys = ("y_rhs", "forc", "stale", "infil_rhs", "y_new", "xflux_sum", "ws0")
            steps.append({k: g[f"sum{i}_s{s}_{k}"] for k in keys})
        out.append(dict(ws0=g[f"sum{i}_ws0"], steps=steps))
    return out


def et_cases(g, prefix="et"):
    """IntcpSnowEt known-answer cases (make_golden.py et); prefix 'etb': the mixed-type table et_i32_b"""
    keys = ("t", "state_in", "y", "meteo", "lai", "lai_lc", "z0_lc", "meltf", "stepsize", "out")
    return [{k: g[f"{prefix}{i}_{k}"] for k in keys} for i in range(int(g[f"{prefix}_n"]))]


def dy_scale(tables, case_forc, xflux, rivflow):
    """Per-component magnitude of the flux terms that are summed into dy
    (SURVEY 8(d) 'Parity acceptance'): the 1e-12 bound is relative to
    max(|dy|, sum |terms|) because dy is a difference of fluxes."""
    from mm_pihm_b200 import watershed as W
    ne, nr, fbr = tables["nelem"], tables["nriver"], bool(tables["fbr"])
    ef, rf = tables["elem_f64"], tables["riv_f64"]
    area, por = ef[W.E_AREA], ef[W.E_POROSITY]
    X = np.abs(xflux)
    s_surf = np.abs(case_forc[W.F_PCPDRP]) + X[W.X_INFIL] + X[W.X_EDIR_SURF] + X[W.X_OVL0:W.X_OVL0 + 3].sum(0) / area
    s_unsat = (X[W.X_INFIL] + X[W.X_RECHG] + X[W.X_EDIR_UNSAT] + X[W.X_ETT_UNSAT]) / por
    s_gw = (X[W.X_RECHG] + X[W.X_EDIR_GW] + X[W.X_ETT_GW] + X[W.X_FBR_INFIL]
            + X[W.X_SUB0:W.X_SUB0 + 3].sum(0) / area) / por
    parts = [s_surf, s_unsat, s_gw]
    if nr:
        R = np.abs(rivflow)
        parts += [R[0:7].sum(0) / rf[W.R_AREA],
                  (R[7] + R[8] + R[9] + R[10] + R[6]) / (rf[W.R_POROSITY] * rf[W.R_AREA])]
    else:
        parts += [np.zeros(0), np.zeros(0)]
    if fbr:
        gp = ef[W.E_GPOROSITY]
        parts += [(X[W.X_FBR_INFIL] + X[W.X_FBR_RECHG]) / gp,
                  (X[W.X_FBR_RECHG] + X[W.X_FBRFLOW0:W.X_FBRFLOW0 + 3].sum(0) / area) / gp]
    return np.concatenate(parts)


def rel_err(a, ref, scale=None):
    den = np.abs(ref) if scale is None else np.maximum(np.abs(ref), scale)
    den = np.where(den == 0.0, 1.0, den)
    return np.abs(a - ref) / den


def record(name, **values):
    """append a measured figure (parity multiple, worst error ...) to gpurun_out/parity_record.jsonl so that
    the numbers the tolerance tests observed are kept next to the pass / fail verdict (copied to profiles/)"""
    import json
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_record.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass
