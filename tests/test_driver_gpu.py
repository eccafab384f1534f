"""The drop-in boundary, end to end: the reference's own programs -- main.c, pihm.c,
the readers, Initialize(), Summary(), the print path and CVODE, all UNCHANGED -- linked with
glue/pihm_b200_glue.c in place of src/ode.c (oracle/Makefile `drivers`) run the bundled
input/example project on the GPU; the unmodified reference program runs the same project on the
CPU; the binary output files (.dat records of print.c:230-246) of the two are compared.

Hydrologic states (surf, unsat, gw, river stage, river gw) must agree within
10 x (reltol |y| + abstol) (reltol 1e-3, abstol 1e-4 of example.para), the
hourly-averaged fluxes within 1e-6 of the column's largest magnitude (they are sums of RHS terms:
1e-12 per evaluation, amplified by the integrator's step choices)."""
import glob
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
RUN = os.path.join(REFDIR, "run")
RELTOL, ABSTOL = 1e-3, 1e-4
STATES = ("surf", "unsat", "gw", "stage", "rivgw", "fbr_unsat", "fbr_gw", "deep_unsat", "deep_gw")


def run_program(exe, outdir, env_extra=None, rundir=RUN, project="example"):
    exe = os.path.join(REFDIR, exe)
    assert os.path.exists(exe), f"{exe} missing: make -C oracle drivers (needs /root/reference)"
    out = os.path.join(rundir, "output", outdir)
    shutil.rmtree(out, ignore_errors=True)
    env = dict(os.environ, OMP_NUM_THREADS=str(min(os.cpu_count() or 1, 8)))
    env.update(env_extra or {})
    p = subprocess.run([exe, "-o", outdir, project], cwd=rundir, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, f"{exe} failed:\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}"
    assert "Simulation completed." in p.stdout
    return out


def read_dat(path):
    raw = np.fromfile(path, dtype=np.float64)
    return raw


def compare(ref_out, our_out, label, state_bound=10.0, flux_bound=1e-6, what="input/example, 3 simulated hours"):
    files = sorted(glob.glob(os.path.join(ref_out, "*.dat")))
    assert len(files) >= 25, files
    worst_state, worst_flux, bad = 0.0, 0.0, []
    for f in files:
        name = os.path.basename(f)
        var = name.split(".")[1]
        a = read_dat(f)
        b = read_dat(os.path.join(our_out, name))
        assert a.shape == b.shape and a.size > 0, f"{name}: {a.shape} vs {b.shape}"
        assert np.isfinite(b).all(), name
        if var in STATES:
            mult = float((np.abs(a - b) / (RELTOL * np.abs(a) + ABSTOL)).max())
            worst_state = max(worst_state, mult)
            if mult > state_bound:
                bad.append(f"{name}: {mult:.3g} x (reltol|y|+abstol)")
        else:
            scale = max(np.abs(a).max(), 1e-300)
            rel = float(np.abs(a - b).max() / scale)
            worst_flux = max(worst_flux, rel)
            if rel > flux_bound:
                bad.append(f"{name}: {rel:.3g} of the column's magnitude")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import record
    record(f"unchanged driver, {what}: " + label, files=len(files),
           states_multiple_of_reltol_y_plus_abstol=worst_state, bound=state_bound, fluxes_rel_to_column_magnitude=worst_flux,
           flux_bound=flux_bound)
    print(f"[{label}] {len(files)} output files: states within {worst_state:.3g} x (reltol|y|+abstol), "
          f"fluxes within {worst_flux:.3g} of their magnitude")
    assert not bad, f"{label}: " + "; ".join(bad)
    return worst_state, worst_flux


@pytest.fixture(scope="module")
def ref_run():
    return run_program("pihm_ref", "ref_out")


@pytest.mark.parametrize("route", [2, 1])
def test_unchanged_driver_on_the_gpu(ref_run, route):
    ours = run_program("pihm_b200", f"b200_route{route}", {"PIHM_B200_ROUTE": str(route)})
    compare(ref_run, ours, f"pihm route {route}")


def test_unchanged_fbr_driver_on_the_gpu():
    ref = run_program("pihm_fbr_ref", "fbr_ref_out")
    ours = run_program("pihm_fbr_b200", "fbr_b200")
    compare(ref, ours, "pihm-fbr route 2")


def test_routes_agree_bitwise():
    """CVODE on the device N_Vector (route 1) and the device integrator (route 2) write the same bytes"""
    a = os.path.join(RUN, "output", "b200_route1")
    b = os.path.join(RUN, "output", "b200_route2")
    if not (os.path.isdir(a) and os.path.isdir(b)):
        pytest.skip("needs the outputs of test_unchanged_driver_on_the_gpu")
    for f in sorted(glob.glob(os.path.join(a, "*.dat"))):
        assert np.array_equal(read_dat(f), read_dat(os.path.join(b, os.path.basename(f)))), os.path.basename(f)


@pytest.mark.parametrize("fbr", [False, True])
def test_unchanged_driver_on_a_synthetic_project(tmp_path, fbr):
    """The same two programs on a synthetic watershed that reaches them as FILES (project_files.write_project:
    2400 triangles, 5 soil and 4 land-cover classes, main stem + tributaries, one meteo station; 3 simulated hours
    with the rain pulse): the reference's readers, Initialize(), forcing interpolation, IntcpSnowEt, Summary and
    print path run unchanged in both, glue/pihm_b200_glue.c packs what Initialize() built.  State bound 30 x
    (reltol |y| + abstol), not 10: on this case the reference program itself moves by 8.6 of these units (hourly
    mean of unsat) when the node elevations of the .mesh file are perturbed by one unit in the last place
    (measured in the build container, tests/golden/README of this case is this docstring); fluxes 1e-6 as above."""
    sys.path.insert(0, ROOT)
    import mm_pihm_b200  # noqa: F401
    from mm_pihm_b200 import project_files as PF, watershed as W
    tb = W.make_named("small", fbr=fbr, keep_mesh=True)
    PF.write_project(tb, str(tmp_path), "synth", hours=3)
    ref = run_program("pihm_fbr_ref" if fbr else "pihm_ref", "ref_out", rundir=str(tmp_path), project="synth")
    ours = run_program("pihm_fbr_b200" if fbr else "pihm_b200", "b200_out", rundir=str(tmp_path), project="synth")
    compare(ref, ours, "pihm-fbr" if fbr else "pihm", state_bound=30.0,
            what="synthetic 2400-triangle project from files, 3 simulated hours")
