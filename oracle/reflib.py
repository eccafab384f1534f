"""ctypes front-end of oracle/_ref/libpihm*_ref.so -- TEST INFRASTRUCTURE ONLY.

The libraries hold the unmodified MM-PIHM reference (src/*.c + vendored CVODE
2.9.0) plus oracle/ref_shim.c.  Because the reference keeps its model in
process globals (src/main.c:4-16) and SetCVodeParam() has a function-static
`reset` flag (src/ode.c:343), every RefModel loads a PRIVATE COPY of the
library (fresh globals / statics), so several models can coexist.

May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# column counts of include/pihm_b200.h (checked against the shim at load)
PB_E_NCOL, PB_EI_NCOL, PB_F_NCOL, PB_R_NCOL, PB_RI_NCOL, PB_X_NCOL = 37, 9, 10, 17, 5, 18
NUM_RIVFLX = 11


class MeshStruct(C.Structure):
    """struct pihm_b200_mesh of include/pihm_b200.h"""
    _fields_ = [
        ("nelem", C.c_int32), ("nriver", C.c_int32), ("fbr", C.c_int32),
        ("surf_mode", C.c_int32), ("riv_mode", C.c_int32), ("reserved", C.c_int32),
        ("stepsize", C.c_double),
        ("elem_f64", C.c_void_p), ("elem_i32", C.c_void_p),
        ("riv_f64", C.c_void_p), ("riv_i32", C.c_void_p),
    ]


def lib_path(fbr: bool = False, cvode_omp: bool = False) -> str:
    name = "libpihm_fbr_ref.so" if fbr else (
        "libpihm_ref_cvomp.so" if cvode_omp else "libpihm_ref.so")
    return os.path.join(REF_DIR, name)


def available(fbr: bool = False) -> bool:
    return os.path.exists(lib_path(fbr))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefModel:
    """One reference model instance (private copy of the library)."""

    def __init__(self, fbr: bool = False, cvode_omp: bool = False, threads: int = 0):
        src = lib_path(fbr, cvode_omp)
        if not os.path.exists(src):
            raise FileNotFoundError(
                f"{src} missing: run `make -C oracle ref` where /root/reference exists")
        fd, self._tmp = tempfile.mkstemp(suffix=".so", prefix="pihm_ref_")
        os.close(fd)
        shutil.copyfile(src, self._tmp)
        self.lib = C.CDLL(self._tmp)
        os.unlink(self._tmp)          # mapping stays valid
        L = self.lib
        L.ref_time_ode.restype = C.c_double
        L.ref_nvec_op.restype = C.c_double
        L.ref_nvec_op.argtypes = [C.c_int, C.c_long, C.c_double, C.c_void_p,
                                  C.c_double, C.c_void_p, C.c_void_p]
        L.ref_ode.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        L.ref_set_max_step.argtypes = [C.c_double]
        L.ref_set_maxstep_ctrl.argtypes = [C.c_double] * 6
        L.ref_create_from_tables.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        self.fbr = bool(L.ref_is_fbr())
        assert self.fbr == fbr
        self.threads = L.ref_set_threads(int(threads))
        self.nelem = self.nriver = self.nsv = 0
        self._keep = None
        self.opened = False

    # -- construction ------------------------------------------------------
    def open_project(self, rundir: str, project: str):
        """ReadAlloc + Initialize of the reference on input/<project>/ under rundir."""
        cwd = os.getcwd()
        try:
            rc = self.lib.ref_open_project(rundir.encode(), project.encode(), 0)
        finally:
            os.chdir(cwd)
        if rc != 0:
            raise RuntimeError("ref_open_project failed")
        self._dims()
        return self

    def create_from_tables(self, tables: dict, reltol=1e-3, abstol=1e-4, initstep=5e-5):
        """tables: dict with nelem, nriver, fbr, surf_mode, riv_mode, stepsize,
        elem_f64, elem_i32, riv_f64, riv_i32 (numpy, C-contiguous)."""
        m = MeshStruct()
        m.nelem, m.nriver = int(tables["nelem"]), int(tables["nriver"])
        m.fbr = int(tables["fbr"])
        m.surf_mode, m.riv_mode = int(tables["surf_mode"]), int(tables["riv_mode"])
        m.stepsize = float(tables["stepsize"])
        keep = []
        for key, dt in (("elem_f64", np.float64), ("elem_i32", np.int32),
                        ("riv_f64", np.float64), ("riv_i32", np.int32)):
            a = np.ascontiguousarray(tables[key], dtype=dt)
            keep.append(a)
            setattr(m, key, a.ctypes.data)
        self._keep = keep
        rc = self.lib.ref_create_from_tables(C.byref(m), reltol, abstol, initstep)
        if rc != 0:
            raise RuntimeError(f"ref_create_from_tables failed ({rc})")
        self._dims()
        return self

    def _dims(self):
        ne, nr, nsv = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_get_dims(C.byref(ne), C.byref(nr), C.byref(nsv))
        self.nelem, self.nriver, self.nsv = ne.value, nr.value, nsv.value
        self.opened = True

    def close(self):
        if self.opened:
            self.lib.ref_close()
            self.opened = False

    # -- tables --------------------------------------------------------------
    def ctrl(self) -> dict:
        ic = (C.c_int * 7)()
        dc = (C.c_double * 10)()
        self.lib.ref_get_ctrl(ic, dc)
        keys_i = ["surf_mode", "riv_mode", "stepsize", "etstep", "starttime", "endtime", "nstep"]
        keys_d = ["abstol", "reltol", "initstep", "maxstep", "stmin", "nncfn",
                  "nnimax", "nnimin", "decr", "incr"]
        out = {k: ic[i] for i, k in enumerate(keys_i)}
        out.update({k: dc[i] for i, k in enumerate(keys_d)})
        return out

    def pack_tables(self) -> dict:
        ne, nr = self.nelem, self.nriver
        ef = np.zeros((PB_E_NCOL, ne)); ei = np.zeros((PB_EI_NCOL, ne), np.int32)
        rf = np.zeros((PB_R_NCOL, max(nr, 0))); ri = np.zeros((PB_RI_NCOL, max(nr, 0)), np.int32)
        self.lib.ref_pack_tables(_ptr(ef), _ptr(ei), _ptr(rf), _ptr(ri))
        c = self.ctrl()
        return dict(nelem=ne, nriver=nr, fbr=int(self.fbr), surf_mode=c["surf_mode"],
                    riv_mode=c["riv_mode"], stepsize=float(c["stepsize"]),
                    elem_f64=ef, elem_i32=ei, riv_f64=rf, riv_i32=ri)

    def get_forcing(self):
        f = np.zeros((PB_F_NCOL, self.nelem)); rb = np.zeros(max(self.nriver, 1))
        self.lib.ref_get_forcing(_ptr(f), _ptr(rb))
        return f, rb[:self.nriver]

    def set_forcing(self, forc, rivbc=None):
        f = np.ascontiguousarray(forc, np.float64)
        assert f.shape == (PB_F_NCOL, self.nelem)
        rb = None if rivbc is None else np.ascontiguousarray(rivbc, np.float64)
        self.lib.ref_set_forcing(_ptr(f), None if rb is None else _ptr(rb))

    def get_ovlflow(self):
        o = np.zeros((3, self.nelem)); self.lib.ref_get_ovlflow(_ptr(o)); return o

    def set_ovlflow(self, ovl):
        o = np.ascontiguousarray(ovl, np.float64); assert o.shape == (3, self.nelem)
        self.lib.ref_set_ovlflow(_ptr(o))

    def get_y(self):
        y = np.zeros(self.nsv); self.lib.ref_get_y(_ptr(y)); return y

    def set_y(self, y):
        y = np.ascontiguousarray(y, np.float64); assert y.shape == (self.nsv,)
        self.lib.ref_set_y(_ptr(y))

    def init_state(self, y):
        y = np.ascontiguousarray(y, np.float64); assert y.shape == (self.nsv,)
        self.lib.ref_init_state(_ptr(y))

    def get_ws(self):
        y = np.zeros(self.nsv); self.lib.ref_get_ws(_ptr(y)); return y

    # -- interception / snow / ET (files mode) -----------------------------------
    def et_dims(self):
        d = (C.c_int * 4)()
        if self.lib.ref_et_dims(d) != 0:
            raise RuntimeError("ET hooks need a project opened from files")
        return dict(nmeteo=d[0], nlai=d[1], nlc=d[2], etstep=d[3])

    def pack_et_tables(self):
        f = np.zeros((13, self.nelem)); ii = np.zeros((3, self.nelem), np.int32)
        self.lib.ref_pack_et_tables(_ptr(f), _ptr(ii))
        return f, ii

    def et_set_types(self, eti):
        ii = np.ascontiguousarray(eti, np.int32); assert ii.shape == (3, self.nelem)
        self.lib.ref_et_set_types(_ptr(ii))

    def et_monthly(self, t, nlc=40):
        a = np.zeros(nlc); b = np.zeros(nlc); mf = C.c_double()
        self.lib.ref_et_monthly(int(t), int(nlc), _ptr(a), _ptr(b), C.byref(mf))
        return a, b, mf.value

    def et_get_forc(self):
        d = self.et_dims()
        m = np.zeros((d["nmeteo"], 7)); l = np.zeros(max(d["nlai"], 1))
        self.lib.ref_et_get_forc(_ptr(m), _ptr(l))
        return m, l[:d["nlai"]]

    def et_get_cal(self):
        c = np.zeros(3); self.lib.ref_et_get_cal(_ptr(c)); return c

    def et_get(self):
        o = np.zeros((7, self.nelem)); self.lib.ref_et_get(_ptr(o)); return o

    def et_set_state(self, sneqv, cmc):
        a = np.ascontiguousarray(sneqv, np.float64); b = np.ascontiguousarray(cmc, np.float64)
        self.lib.ref_et_set_state(_ptr(a), _ptr(b))

    def et_run(self, t, stepsize, meteo, lai, y):
        m = np.ascontiguousarray(meteo, np.float64); l = np.ascontiguousarray(lai if len(lai) else [0.0], np.float64)
        y = np.ascontiguousarray(y, np.float64)
        self.lib.ref_et_run.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        if self.lib.ref_et_run(int(t), float(stepsize), _ptr(m), _ptr(l), _ptr(y)) != 0:
            raise RuntimeError("ref_et_run: files mode only")

    def tout(self, cstep):
        return int(self.lib.ref_tout(int(cstep)))

    # -- print accumulation: the reference's UpdPrintVar / PrintData on a field -----
    def print_add(self, src, col, upd_intvl=0, intvl=60):
        vid = self.lib.ref_print_add(int(src), int(col), int(upd_intvl), int(intvl))
        if vid < 0:
            raise RuntimeError("ref_print_add: unknown field")
        return vid

    def print_update(self, module_step=0):
        self.lib.ref_print_update(int(module_step))

    def print_data(self, vid, t, lapse, n):
        out = np.zeros(max(n, 1))
        rc = self.lib.ref_print_data(int(vid), int(t), int(lapse), _ptr(out))
        return (out[:n] if rc == 1 else None)

    def print_reset(self):
        self.lib.ref_print_reset()

    def summary(self, y):
        """Summary() of src/update.c on y (= CV_Y after SolveCVode)."""
        y = np.ascontiguousarray(y, np.float64); assert y.shape == (self.nsv,)
        self.lib.ref_summary(_ptr(y))

    def get_ws0(self):
        y = np.zeros(self.nsv); self.lib.ref_get_ws0(_ptr(y)); return y

    # -- RHS -------------------------------------------------------------------
    def ode(self, y, t=0.0):
        y = np.ascontiguousarray(y, np.float64); assert y.shape == (self.nsv,)
        dy = np.zeros(self.nsv)
        self.lib.ref_ode(float(t), _ptr(y), _ptr(dy))
        return dy

    def time_ode(self, n: int) -> float:
        return float(self.lib.ref_time_ode(int(n)))

    def get_fluxes(self):
        xf = np.zeros((PB_X_NCOL, self.nelem)); rf = np.zeros((NUM_RIVFLX, max(self.nriver, 1)))
        self.lib.ref_get_fluxes(_ptr(xf), _ptr(rf))
        return xf, rf[:, :self.nriver]

    # -- integrator ------------------------------------------------------------
    def set_cvode_param(self):
        self.lib.ref_set_cvode_param()

    def set_maxstep_ctrl(self, stmin=1.0, nncfn=0.0, nnimax=3.0, nnimin=1.0, decr=1.2, incr=1.2):
        """the AdjCVodeMaxStep parameters (ctrl_struct; the reference reads them from the .para file)"""
        self.lib.ref_set_maxstep_ctrl(stmin, nncfn, nnimax, nnimin, decr, incr)

    def set_max_step(self, h):
        self.lib.ref_set_max_step(float(h))

    def model_step(self, cstep: int, adj_max_step: bool = False, skip_forcing: bool = False) -> int:
        return int(self.lib.ref_model_step(int(cstep), int(adj_max_step), int(skip_forcing)))

    def apply_forcing(self, cstep: int):
        self.lib.ref_apply_forcing(int(cstep))

    def stats(self) -> dict:
        s = (C.c_long * 13)(); d = (C.c_double * 3)()
        self.lib.ref_get_stats(s, d)
        keys = ["nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS",
                "njtimes", "nor", "nsetups", "qlast", "qcur"]
        out = {k: int(s[i]) for i, k in enumerate(keys)}
        out.update(hlast=d[0], hcur=d[1], tcur=d[2])
        return out

    # -- serial N_Vector kernels (nvector_serial.c) ------------------------------
    def nvec_op(self, op: int, a=0.0, x=None, b=0.0, y=None, n=None):
        n = int(n if n is not None else (len(x) if x is not None else len(y)))
        z = np.zeros(n)
        xx = None if x is None else np.ascontiguousarray(x, np.float64)
        yy = None if y is None else np.ascontiguousarray(y, np.float64)
        r = self.lib.ref_nvec_op(op, n, float(a), None if xx is None else _ptr(xx),
                                 float(b), None if yy is None else _ptr(yy), _ptr(z))
        return r, z


class RefCvodeExternal:
    """The reference CVODE (private copy of libpihm_ref.so) running on an external
    N_Vector / RHS callback: N_VNew_PihmB200 + PihmB200_ODE."""

    def __init__(self, nv_y_ptr: int, rhs_fn_ptr: int, user_data_ptr: int, reltol=1e-3,
                 abstol=1e-4, initstep=5e-5, maxstep=60.0, mxsteps=600):
        self.model = RefModel(fbr=False)
        L = self.model.lib
        L.ref_ext_cvode_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_long]
        L.ref_ext_cvode_solve.argtypes = [C.c_double, C.POINTER(C.c_double)]
        rc = L.ref_ext_cvode_init(nv_y_ptr, rhs_fn_ptr, user_data_ptr, reltol, abstol, initstep,
                                  maxstep, mxsteps)
        if rc != 0:
            raise RuntimeError(f"ref_ext_cvode_init failed ({rc})")

    def attach_lsolve(self, solve_fn_ptr: int, engine_ptr: int):
        """swap CVSPGMR for a device linear solver (pihm_b200_spgmr_solve) through cv_mem->cv_lsolve"""
        L = self.model.lib
        L.ref_ext_cvode_attach_lsolve.argtypes = [C.c_void_p, C.c_void_p]
        if L.ref_ext_cvode_attach_lsolve(solve_fn_ptr, engine_ptr) != 0:
            raise RuntimeError("ref_ext_cvode_attach_lsolve failed")

    def stepper_stats(self):
        s = (C.c_long * 5)()
        self.model.lib.ref_ext_cvode_stepper_stats(s)
        return {k: int(s[i]) for i, k in enumerate(["nst", "nfe", "nni", "ncfn", "netf"])}

    def solve(self, tout: float) -> float:
        t = C.c_double()
        rc = self.model.lib.ref_ext_cvode_solve(float(tout), C.byref(t))
        if rc < 0:
            raise RuntimeError(f"reference CVode failed ({rc})")
        return t.value

    def stats(self):
        s = (C.c_long * 8)()
        self.model.lib.ref_ext_cvode_stats(s)
        keys = ["nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl", "nfeLS"]
        return {k: int(s[i]) for i, k in enumerate(keys)}

    def free(self):
        self.model.lib.ref_ext_cvode_free()
