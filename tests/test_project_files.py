"""The synthetic watershed through the reference's OWN front end: project_files.write_project() emits the
text inputs, oracle/_ref -- the unmodified MM-PIHM compiled from /root/reference -- reads
them with ReadAlloc() and builds its structures with Initialize() (InitTopo src/init_topo.c:14-44, InitSurfL
src/initialize.c:365-438, InitRiver src/init_river.c:10-116, InitSoil, InitLc, InitGeol, RelaxIc), and the
tables packed from those structures must be the ones watershed.make_watershed() states directly -- which is
what every synthetic parity case, the bench and the golden files hand to both arms.  A mistake in the
generator's geometry (neighbour convention, river banks, distances, initial state) shows up here.
VERDICT r1 'missing 5'.  CPU only; needs oracle/_ref (skipped on a box without it)."""
import numpy as np
import pytest

import mm_pihm_b200  # noqa: F401
from mm_pihm_b200 import project_files as PF, watershed as W

reflib = pytest.importorskip("reflib")

ULP2 = 4.5e-16          # two units in the last place: the re-rounded products named below


def _open(tmp_path, tb, fbr):
    if not reflib.available(fbr):
        pytest.skip("oracle/_ref not present")
    PF.write_project(tb, str(tmp_path), "synth", hours=24)
    return reflib.RefModel(fbr=fbr).open_project(str(tmp_path), "synth")


@pytest.mark.parametrize("fbr", [False, True])
@pytest.mark.parametrize("riv_order", [1, 3])
def test_reference_front_end_builds_the_same_tables(tmp_path, fbr, riv_order):
    tb = W.make_watershed(40, 30, fbr=fbr, riv_order=riv_order, keep_mesh=True)
    ref = _open(tmp_path, tb, fbr)
    got = ref.pack_tables()
    y0 = ref.get_y()
    ctrl = ref.ctrl()
    ref.close()
    assert (got["nelem"], got["nriver"]) == (tb["nelem"], tb["nriver"])
    assert (got["surf_mode"], got["riv_mode"], got["stepsize"]) == (tb["surf_mode"], tb["riv_mode"], tb["stepsize"])
    assert (ctrl["reltol"], ctrl["abstol"], ctrl["initstep"], ctrl["etstep"]) == (1e-3, 1e-4, 5e-5, 900)
    # topology: neighbours with the river edges rewritten by InitRiver, bc types, banks, downstream links
    assert np.array_equal(got["elem_i32"], tb["elem_i32"][:got["elem_i32"].shape[0]])
    assert np.array_equal(got["riv_i32"], tb["riv_i32"][:got["riv_i32"].shape[0]])
    # geometry and parameters: bit for bit, except products the reference forms from other file columns
    loose_e = {W.E_KMACH: "kmach = KMACH_RO * ksath", W.E_KMACV: "kmacv = KMACV_RO * kinfv"}
    names = {v: k for k, v in vars(W).items() if k.startswith("E_") and isinstance(v, int)}
    worst = {}
    for col in range(got["elem_f64"].shape[0]):
        a, b = got["elem_f64"][col], tb["elem_f64"][col]
        if not fbr and col in (W.E_ZBED, W.E_GDEPTH, W.E_GKSATH, W.E_GKSATV, W.E_GALPHA, W.E_GBETA, W.E_GPOROSITY):
            continue
        if col in loose_e:
            assert np.allclose(a, b, rtol=ULP2, atol=0), f"{names[col]} ({loose_e[col]})"
        else:
            if not np.array_equal(a, b):
                worst[names[col]] = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    assert not worst, f"element columns differ from the generator's: {worst}"
    rnames = {v: k for k, v in vars(W).items() if k.startswith("R_") and isinstance(v, int)}
    for col in range(got["riv_f64"].shape[0]):
        a, b = got["riv_f64"][col], tb["riv_f64"][col]
        assert np.array_equal(a, b), f"{rnames[col]}: {np.max(np.abs(a - b))}"
    # RelaxIc (src/initialize.c:476-553) + InitVar
    assert np.array_equal(y0, tb["y0"])


def test_rhs_of_the_file_built_model_equals_the_table_built_one(tmp_path):
    """ODE() of the reference on the structures its front end built from the files == ODE() on the structures
    the shim fills from the generator's tables (the route every other synthetic test takes)."""
    tb = W.make_watershed(40, 30, keep_mesh=True)
    ne, nr = tb["nelem"], tb["nriver"]
    y = W.wet_state(tb, seed=11)
    forc = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    outs = []
    for from_files in (True, False):
        if from_files:
            ref = _open(tmp_path, tb, False)
        else:
            ref = reflib.RefModel(fbr=False).create_from_tables(tb)
        ref.set_forcing(forc, np.zeros(nr)); ref.set_ovlflow(np.zeros((3, ne)))
        outs.append([ref.ode(y), ref.ode(y)])
        ref.close()
    scale = np.abs(outs[1][1]).max()
    for call in range(2):
        d = np.abs(outs[0][call] - outs[1][call]).max()
        assert d <= 1e-13 * scale, f"call {call}: {d:.3e} (kmach / kmacv differ in the last place at most)"


@pytest.mark.parametrize("fbr", [False, True])
def test_dirichlet_edges_travel_as_bc_series(tmp_path, fbr):
    """bc type = index of a series of the .bc file (src/forcing.c:55-86): one constant series per Dirichlet edge.
    The reference's Initialize() must mark exactly the generator's edges, ApplyBc() must hand every edge the
    generator's head, and ODE() on the file-built model must equal ODE() on the table-built one."""
    tb = W.make_watershed(40, 30, fbr=fbr, dirichlet_edges=True, keep_mesh=True)
    ne, nr = tb["nelem"], tb["nriver"]
    ref = _open(tmp_path, tb, fbr)
    got = ref.pack_tables()
    nrow = got["elem_i32"].shape[0]
    bc_rows = list(range(W.EI_BC0, W.EI_BC2 + 1)) + (list(range(W.EI_FBRBC0, W.EI_FBRBC2 + 1)) if fbr else [])
    for r in range(nrow):
        if r in bc_rows:
            assert np.array_equal(got["elem_i32"][r] > 0, tb["elem_i32"][r] > 0) and (got["elem_i32"][r] >= 0).all()
        else:
            assert np.array_equal(got["elem_i32"][r], tb["elem_i32"][r])
    nset = int((tb["elem_i32"][bc_rows] > 0).sum())
    assert nset > 0 and sorted(got["elem_i32"][bc_rows][got["elem_i32"][bc_rows] > 0]) == list(range(1, nset + 1))
    ref.apply_forcing(0)                                    # ApplyBc at the start time
    forc, _ = ref.get_forcing()
    want = W.storm_forcing(tb, 0.0)
    for j in range(3):
        sel = tb["elem_i32"][W.EI_BC0 + j] > 0
        assert np.allclose(forc[W.F_BC0 + j, sel], want[W.F_BC0 + j, sel], rtol=ULP2, atol=0)
        if fbr:
            sel = tb["elem_i32"][W.EI_FBRBC0 + j] > 0
            assert np.allclose(forc[W.F_FBRBC0 + j, sel], want[W.F_FBRBC0 + j, sel], rtol=ULP2, atol=0)
    # the same right-hand side as the table-built model (which carries the flag 1 where the files carry an index)
    y = W.wet_state(tb, seed=3)
    f = W.storm_forcing(tb, 3 * 3600.0, ws0_surf=np.maximum(y[:ne], 0))
    ref.set_forcing(f, np.zeros(nr)); ref.set_ovlflow(np.zeros((3, ne)))
    a = ref.ode(y)
    ref.close()
    tab = reflib.RefModel(fbr=fbr).create_from_tables(tb)
    tab.set_forcing(f, np.zeros(nr)); tab.set_ovlflow(np.zeros((3, ne)))
    b = tab.ode(y)
    tab.close()
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()


def test_lai_series_is_read(tmp_path):
    """option lai_series: every element refers to LAI series 1 of the .lai file (src/read_lai.c, forcing.c:242-258)"""
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    tb = W.make_watershed(8, 6, keep_mesh=True)
    PF.write_project(tb, str(tmp_path), "synth", hours=6, lai_series=True)
    ref = reflib.RefModel(fbr=False).open_project(str(tmp_path), "synth")
    assert ref.et_dims()["nlai"] == 1 and ref.et_dims()["nmeteo"] == 1
    _, eti = ref.pack_et_tables()
    assert (eti[W.ETI_LAI_TYPE] == 1).all()
    ref.apply_forcing(0)
    meteo, lai = ref.et_get_forc()
    assert lai[0] == 3.0 and meteo[0, 1] == 285.15           # the constant series / SFCTMP of write_meteo
    ref.close()


def test_writer_refuses_what_it_cannot_express(tmp_path):
    with pytest.raises(ValueError):
        PF.write_project(W.make_watershed(8, 6), str(tmp_path))                      # no node-level mesh kept
    tb = W.make_watershed(8, 6, dirichlet_edges=True, keep_mesh=True)
    tb["elem_i32"] = tb["elem_i32"].copy()
    tb["elem_i32"][W.EI_BC1, 3] = -1                                                  # a Neumann edge
    with pytest.raises(ValueError):
        PF.write_project(tb, str(tmp_path))
