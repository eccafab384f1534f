"""glue/pihm_b200_glue.c's pihm_struct -> column-table packer, checked on the CPU:
the unchanged pihm / pihm-fbr programs linked with the glue (oracle/_ref/pihm_b200, pihm_fbr_b200) read a
project, run the reference's ReadAlloc() + Initialize(), and call the C ABI; an LD_PRELOAD interposer
(tests/glue_capture_stub.c, test infrastructure, computes nothing) takes the place of pihm_b200_create(), writes
the tables the glue handed over and ends the program.  They must be, bit for bit, what the test shim's own packer
(oracle/ref_shim.c ref_pack_tables -- written independently, pinned to the generator's tables by
tests/test_project_files.py) extracts from the same structures: input/example and a synthetic project written as
files, pihm and pihm-fbr.  (What the programs then compute on the device is tests/test_driver_gpu.py.)"""
import os
import shutil
import subprocess

import numpy as np
import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import project_files as PF, watershed as W

reflib = pytest.importorskip("reflib")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
NE, NEI, NR, NRI = W.E_NCOL, W.EI_NCOL, W.R_NCOL, W.RI_NCOL


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("stub") / "glue_capture_stub.so")
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", so,
                           os.path.join(ROOT, "tests", "glue_capture_stub.c")])
    return so


def capture(stub, exe, rundir, project, tmp_path):
    exe = os.path.join(REFDIR, exe)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} missing (make -C oracle drivers needs /root/reference)")
    cap = str(tmp_path / "capture.bin")
    env = dict(os.environ, LD_PRELOAD=stub, PIHM_B200_CAPTURE=cap, OMP_NUM_THREADS="1")
    p = subprocess.run([exe, "-o", "cap_out", project], cwd=rundir, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and os.path.exists(cap), p.stdout[-1500:] + p.stderr[-1500:]
    raw = open(cap, "rb").read()
    head = np.frombuffer(raw, np.int32, 8)
    ne, nr = int(head[0]), int(head[1])
    off = 32
    step = float(np.frombuffer(raw, np.float64, 1, off)[0]); off += 8
    ef = np.frombuffer(raw, np.float64, NE * ne, off).reshape(NE, ne); off += 8 * NE * ne
    ei = np.frombuffer(raw, np.int32, NEI * ne, off).reshape(NEI, ne); off += 4 * NEI * ne
    rf = np.frombuffer(raw, np.float64, NR * nr, off).reshape(NR, nr); off += 8 * NR * nr
    ri = np.frombuffer(raw, np.int32, NRI * nr, off).reshape(NRI, nr); off += 4 * NRI * nr
    assert off == len(raw)
    return dict(nelem=ne, nriver=nr, fbr=int(head[2]), surf_mode=int(head[3]), riv_mode=int(head[4]),
                device=int(head[5]), reorder=int(head[6]), stepsize=step, elem_f64=ef, elem_i32=ei, riv_f64=rf, riv_i32=ri)


def check(got, rundir, project, fbr):
    ref = reflib.RefModel(fbr=fbr).open_project(rundir, project)
    want = ref.pack_tables()
    ref.close()
    for k in ("nelem", "nriver", "fbr", "surf_mode", "riv_mode", "stepsize"):
        assert got[k] == want[k], k
    assert got["reorder"] == 1 and got["device"] == 0
    for k in ("elem_f64", "elem_i32", "riv_f64", "riv_i32"):
        a, b = got[k], want[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        bad = [c for c in range(a.shape[0]) if not np.array_equal(a[c], b[c])]
        assert not bad, f"{k}: columns {bad} differ from the shim's packer"
    assert got["nelem"] > 0 and np.abs(got["elem_f64"][W.E_AREA]).min() > 0


@pytest.mark.parametrize("fbr", [False, True])
def test_glue_packs_the_example_project(stub, tmp_path, fbr):
    rundir = os.path.join(REFDIR, "run")
    if not os.path.isdir(os.path.join(rundir, "input", "example")):
        pytest.skip("oracle/_ref/run not prepared")
    work = tmp_path / "run"
    shutil.copytree(os.path.join(rundir, "input"), work / "input")          # outputs stay out of oracle/_ref/run
    got = capture(stub, "pihm_fbr_b200" if fbr else "pihm_b200", str(work), "example", tmp_path)
    check(got, str(work), "example", fbr)


@pytest.mark.parametrize("fbr", [False, True])
def test_glue_packs_a_synthetic_project(stub, tmp_path, fbr):
    tb = W.make_watershed(40, 30, fbr=fbr, riv_order=3, keep_mesh=True)
    PF.write_project(tb, str(tmp_path), "synth", hours=3)
    got = capture(stub, "pihm_fbr_b200" if fbr else "pihm_b200", str(tmp_path), "synth", tmp_path)
    check(got, str(tmp_path), "synth", fbr)
    # and, through test_project_files' chain, the generator's own tables
    assert np.array_equal(got["elem_i32"], tb["elem_i32"][:NEI]) and np.array_equal(got["riv_f64"], tb["riv_f64"][:NR])
