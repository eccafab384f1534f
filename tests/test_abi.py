"""CPU-side checks of the drop-in boundary: the C ABI library loads, exports
every symbol include/pihm_b200.h declares, the Python mirror types every one
of them, the column enums agree, and without a GPU the product refuses to run
(no CPU fallback)."""
import ctypes
import os
import re

import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import lib, watershed as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pihm_b200.h")


def header_text():
    with open(HEADER) as f:
        txt = f.read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def declared_symbols():
    return sorted(set(re.findall(r"\b(pihm_b200_[a-z0-9_]+)\s*\(", header_text())))


def test_sundials_header_symbols_exported():
    with open(os.path.join(ROOT, "include", "pihm_b200_sundials.h")) as f:
        txt = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(N_V[A-Za-z0-9_]*PihmB200[A-Za-z0-9_]*|PihmB200_ODE)\s*\(", txt)))
    assert syms == sorted(lib._SUNDIALS_SIGS), syms
    L = lib.load_library()
    for s in syms:
        assert hasattr(L, s)


def test_library_exports_every_declared_symbol():
    L = lib.load_library()
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), f"libpihm_b200.so does not export {s}"
    assert sorted(lib._SIGS) == syms, "lib.py mirror and header disagree"
    assert L.pihm_b200_abi_version() == 1


def test_enums_match_python_mirror():
    txt = header_text()
    for enum, prefix, pyprefix in (("pihm_b200_elem_col", "PB_E_", "E_"),
                                   ("pihm_b200_elem_icol", "PB_EI_", "EI_"),
                                   ("pihm_b200_elem_forc_col", "PB_F_", "F_"),
                                   ("pihm_b200_riv_col", "PB_R_", "R_"),
                                   ("pihm_b200_riv_icol", "PB_RI_", "RI_"),
                                   ("pihm_b200_elem_flux_col", "PB_X_", "X_")):
        body = re.search(r"enum\s+%s\s*\{(.*?)\}" % enum, txt, flags=re.S).group(1)
        names = [n.strip().split("=")[0].strip() for n in body.split(",") if n.strip()]
        for idx, name in enumerate(names):
            assert getattr(W, pyprefix + name[len(prefix):]) == idx, name


def test_new_enums_match_python_mirror():
    txt = header_text()
    for enum, prefix, pyprefix in (("pihm_b200_et_col", "PB_ET_", "ET_"), ("pihm_b200_et_icol", "PB_ETI_", "ETI_"),
                                   ("pihm_b200_et_out_col", "PB_EO_", "EO_"), ("pihm_b200_print_src", "PB_PS_", "PS_")):
        body = re.search(r"enum\s+%s\s*\{(.*?)\}" % enum, txt, flags=re.S).group(1)
        names = [n.strip().split("=")[0].strip() for n in body.split(",") if n.strip()]
        for idx, name in enumerate(names):
            assert getattr(W, pyprefix + name[len(prefix):]) == idx, name
    assert int(re.search(r"PIHM_B200_NUM_METEO_VAR\s+(\d+)", txt).group(1)) == W.NUM_METEO_VAR


def test_null_handles_are_refused():
    """every entry point added for SURVEY 8(f) / the lsolve hook fails cleanly on a null handle (no GPU needed)"""
    L = lib.load_library()
    assert L.pihm_b200_set_diagnostics(None, 1) < 0
    assert L.pihm_b200_set_ws0(None, None) < 0
    assert L.pihm_b200_summary_mb(None, None, 60.0) < 0
    assert L.pihm_b200_get_summary(None, None, None) < 0
    assert L.pihm_b200_et_create(None, None, None) < 0
    assert L.pihm_b200_intcp_snow_et(None, None, None) < 0
    assert L.pihm_b200_et_set_state(None, None, None) < 0
    assert L.pihm_b200_et_get(None, None) < 0
    assert L.pihm_b200_print_add(None, 0, 0) < 0
    assert L.pihm_b200_print_update(None, None, 0, None) < 0
    assert L.pihm_b200_print_data(None, 0, None, None) < 0
    assert L.pihm_b200_spgmr_solve(None, 0.0, 0.0, 0.0, 0, None, None, None, None) < 0
    # pipelined transfers (csrc/transfer.cu)
    assert L.pihm_b200_forcing_prefetch(None, 1, None, None) < 0
    assert L.pihm_b200_forcing_commit(None) < 0
    assert L.pihm_b200_vec_download_async(None, None) < 0
    assert L.pihm_b200_transfer_wait(None) < 0
    assert L.pihm_b200_transfer_release(None) == 0          # nothing to release


def test_struct_layouts():
    assert ctypes.sizeof(lib.EtStep) == 5 * 8 + 4 * 4 + 4 * 8
    assert ctypes.sizeof(lib.MeshStruct) == 64
    assert ctypes.sizeof(lib.CvodeParam) == 48
    assert ctypes.sizeof(lib.MaxStepCtrl) == 64
    assert ctypes.sizeof(lib.CvodeStats) == 11 * 8 + 2 * 4 + 3 * 8


def test_no_cpu_fallback():
    L = lib.load_library()
    if L.pihm_b200_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        lib.Model(W.make_named("tiny"))
    # the raw C entry point refuses too
    m = lib.MeshStruct()
    assert not L.pihm_b200_create(ctypes.byref(m), 0, 0)
    assert L.pihm_b200_last_error()


def test_unchanged_driver_with_the_glue_stops_without_a_gpu():
    """oracle/_ref/pihm_b200 = the reference's main.c / pihm.c / readers / CVODE with glue/pihm_b200_glue.c for
    src/ode.c (make -C oracle drivers): reads the project, initialises, and then fails loudly at
    pihm_b200_create -- the glue has no CPU path either"""
    import subprocess
    L = lib.load_library()
    exe = os.path.join(ROOT, "oracle", "_ref", "pihm_b200")
    run = os.path.join(ROOT, "oracle", "_ref", "run")
    if L.pihm_b200_device_count() > 0 or not os.path.exists(exe):
        pytest.skip("GPU present, or the drivers are not built (needs /root/reference)")
    p = subprocess.run([exe, "-o", "nogpu", "example"], cwd=run, capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert "libpihm_b200" in (p.stdout + p.stderr) and "Simulation completed" not in p.stdout


def test_glue_defines_the_reference_symbols():
    """the five external symbols of src/ode.c (pihm_func.h:104,231,294,296) and the lifecycle hooks"""
    import re
    src = open(os.path.join(ROOT, "glue", "pihm_b200_glue.c")).read()
    for sig in (r"int ODE\(realtype t, N_Vector y, N_Vector ydot, void \*pihm_data\)", r"int NumStateVar\(void\)",
                r"void SetCVodeParam\(pihm_struct pihm, void \*cvode_mem, N_Vector CV_Y\)",
                r"void SolveCVode\(int starttime, int \*t, int nextptr, double cputime, void \*cvode_mem, N_Vector CV_Y\)",
                r"void AdjCVodeMaxStep\(void \*cvode_mem, ctrl_struct \*ctrl\)",
                r"void PihmB200Init\(pihm_struct pihm\)", r"void PihmB200PushForcing\(pihm_struct pihm\)",
                r"void PihmB200PullState\(pihm_struct pihm, N_Vector CV_Y\)", r"void PihmB200Free\(void\)"):
        assert re.search(sig, src), sig
    assert "oracle" not in src.replace("oracle/Makefile", "")


def test_product_never_imports_oracle():
    """the product package must not reference oracle/ anywhere"""
    pkg = os.path.join(ROOT, "mm-pihm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert "oraclelib" not in src and "reflib" not in src and "pihm_oracle" not in src, f
