"""ctypes binding of libpihm_b200.so -- the host-side mirror of the reference
interface for the hot path (names follow src/ode.c: ODE, SetCVodeParam,
SolveCVode, AdjCVodeMaxStep, NumStateVar; N_V* follow nvector_serial.c).

There is deliberately NO fallback: if the CUDA library is missing or no GPU
is visible, constructing a Model raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import watershed as W

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIHM_B200_LIB", os.path.join(HERE, "libpihm_b200.so"))


class MeshStruct(C.Structure):
    """struct pihm_b200_mesh (include/pihm_b200.h)"""
    _fields_ = [
        ("nelem", C.c_int32), ("nriver", C.c_int32), ("fbr", C.c_int32),
        ("surf_mode", C.c_int32), ("riv_mode", C.c_int32), ("reserved", C.c_int32),
        ("stepsize", C.c_double),
        ("elem_f64", C.c_void_p), ("elem_i32", C.c_void_p),
        ("riv_f64", C.c_void_p), ("riv_i32", C.c_void_p),
    ]


class CvodeParam(C.Structure):
    """struct pihm_b200_cvode_param"""
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double), ("initstep", C.c_double),
                ("maxstep", C.c_double), ("mxsteps", C.c_int64),
                ("stab_lim_det", C.c_int32), ("maxl", C.c_int32)]


class EtStep(C.Structure):
    """struct pihm_b200_et_step"""
    _fields_ = [("stepsize", C.c_double), ("cal_edir", C.c_double), ("cal_ec", C.c_double),
                ("cal_ett", C.c_double), ("meltf", C.c_double),
                ("nmeteo", C.c_int32), ("nlai", C.c_int32), ("nlc", C.c_int32), ("reserved", C.c_int32),
                ("meteo", C.c_void_p), ("lai", C.c_void_p), ("lai_lc", C.c_void_p), ("z0_lc", C.c_void_p)]


def make_et_step(stepsize, cal, meltf, meteo, lai, lai_lc, z0_lc):
    """-> (EtStep, keepalive).  meteo [nmeteo, 7], lai [nlai], lai_lc / z0_lc [nlc]; cal = (edir, ec, ett)"""
    meteo = np.ascontiguousarray(meteo, np.float64).reshape(-1, W.NUM_METEO_VAR)
    lai = np.ascontiguousarray(lai if lai is not None else [], np.float64).reshape(-1)
    lai_lc = np.ascontiguousarray(lai_lc, np.float64); z0_lc = np.ascontiguousarray(z0_lc, np.float64)
    assert lai_lc.shape == z0_lc.shape
    st = EtStep(stepsize=float(stepsize), cal_edir=float(cal[0]), cal_ec=float(cal[1]), cal_ett=float(cal[2]),
                meltf=float(meltf), nmeteo=meteo.shape[0], nlai=lai.shape[0], nlc=lai_lc.shape[0], reserved=0,
                meteo=meteo.ctypes.data, lai=lai.ctypes.data if lai.size else None,
                lai_lc=lai_lc.ctypes.data, z0_lc=z0_lc.ctypes.data)
    return st, (meteo, lai, lai_lc, z0_lc)


class CvodeStats(C.Structure):
    """struct pihm_b200_cvode_stats"""
    _fields_ = [(k, C.c_int64) for k in ("nst", "nfe", "nni", "ncfn", "netf", "nli", "ncfl",
                                         "nfeLS", "njtimes", "nor", "nsetups")] + \
               [("qlast", C.c_int32), ("qcur", C.c_int32),
                ("hlast", C.c_double), ("hcur", C.c_double), ("tcur", C.c_double)]


class MaxStepCtrl(C.Structure):
    """struct pihm_b200_maxstep_ctrl"""
    _fields_ = [(k, C.c_double) for k in ("maxstep", "stepsize", "stmin", "nncfn", "nnimax",
                                          "nnimin", "decr", "incr")]


# every symbol include/pihm_b200.h declares (tests/test_abi.py checks the list
# against the header and against the built library)
_SIGS = {
    "pihm_b200_last_error": (C.c_char_p, []),
    "pihm_b200_abi_version": (C.c_int, []),
    "pihm_b200_device_count": (C.c_int, []),
    "pihm_b200_create": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
    "pihm_b200_destroy": (None, [C.c_void_p]),
    "pihm_b200_partition_create": (C.c_void_p, [C.c_void_p, C.c_int]),
    "pihm_b200_partition_destroy": (None, [C.c_void_p]),
    "pihm_b200_partition_sizes": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pihm_b200_partition_fill": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 13),
    "pihm_b200_create_part": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pihm_b200_set_halo": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7),
    "pihm_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pihm_b200_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pihm_b200_num_state_var_global": (C.c_int64, [C.c_void_p]),
    "pihm_b200_comm_paths": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_comm_init_local": (C.c_int, [C.c_void_p, C.c_int]),
    "pihm_b200_halo_pack_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_set_ghosts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_num_state_var": (C.c_int64, [C.c_void_p]),
    "pihm_b200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_get_stream": (C.c_void_p, [C.c_void_p]),
    "pihm_b200_synchronize": (C.c_int, [C.c_void_p]),
    "pihm_b200_launch_count": (C.c_longlong, [C.c_void_p]),
    "pihm_b200_slow_path_count": (C.c_longlong, [C.c_void_p]),
    "pihm_b200_get_permutation": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_set_forcing": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_set_forcing_col": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pihm_b200_set_river_bc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_set_stale_ovlflow": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_ode_host": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "pihm_b200_ode": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "pihm_b200_check_nan": (C.c_int, [C.c_void_p]),
    "pihm_b200_summary": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_test_pow": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_test_div": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_set_flux_recording": (C.c_int, [C.c_void_p, C.c_int]),
    "pihm_b200_get_fluxes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_et_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_intcp_snow_et": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_et_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_et_get": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_print_add": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "pihm_b200_print_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pihm_b200_print_data": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pihm_b200_print_open": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int]),
    "pihm_b200_print_write": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "pihm_b200_print_close": (C.c_int, [C.c_void_p]),
    "pihm_b200_print_io_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_write_ic": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_set_diagnostics": (C.c_int, [C.c_void_p, C.c_int]),
    "pihm_b200_set_ws0": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_summary_mb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double]),
    "pihm_b200_get_summary": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_vec_new": (C.c_void_p, [C.c_void_p]),
    "pihm_b200_vec_free": (None, [C.c_void_p]),
    "pihm_b200_vec_length": (C.c_int64, [C.c_void_p]),
    "pihm_b200_vec_devptr": (C.c_void_p, [C.c_void_p]),
    "pihm_b200_vec_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_vec_download": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_forcing_prefetch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pihm_b200_forcing_commit": (C.c_int, [C.c_void_p]),
    "pihm_b200_vec_download_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_transfer_wait": (C.c_int, [C.c_void_p]),
    "pihm_b200_transfer_release": (C.c_int, [C.c_void_p]),
    "pihm_b200_nv_linearsum": (None, [C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_const": (None, [C.c_double, C.c_void_p]),
    "pihm_b200_nv_prod": (None, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_div": (None, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_scale": (None, [C.c_double, C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_abs": (None, [C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_inv": (None, [C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_addconst": (None, [C.c_void_p, C.c_double, C.c_void_p]),
    "pihm_b200_nv_dotprod": (C.c_double, [C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_maxnorm": (C.c_double, [C.c_void_p]),
    "pihm_b200_nv_wrmsnorm": (C.c_double, [C.c_void_p, C.c_void_p]),
    "pihm_b200_nv_min": (C.c_double, [C.c_void_p]),
    "pihm_b200_cvode_create": (C.c_void_p, [C.c_void_p]),
    "pihm_b200_cvode_destroy": (None, [C.c_void_p]),
    "pihm_b200_cvode_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "pihm_b200_cvode_set_max_step": (C.c_int, [C.c_void_p, C.c_double]),
    "pihm_b200_cvode_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "pihm_b200_cvode_get_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_spgmr_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pihm_b200_adj_cvode_max_step": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_cvode_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "pihm_b200_cvode_get_profile": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pihm_b200_cvode_get_kernel_profile": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_void_p]),
}
# include/pihm_b200_sundials.h
_SUNDIALS_SIGS = {
    "N_VNew_PihmB200": (C.c_void_p, [C.c_void_p]),
    "N_VDestroy_PihmB200": (None, [C.c_void_p]),
    "N_VPihmB200_Push": (C.c_int, [C.c_void_p]),
    "N_VPihmB200_Pull": (C.c_int, [C.c_void_p]),
    "N_VPihmB200_Device": (C.c_void_p, [C.c_void_p]),
    "PihmB200_ODE": (C.c_int, [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load_library():
    """dlopen libpihm_b200.so and type every entry point.  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in list(_SIGS.items()) + list(_SUNDIALS_SIGS.items()):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(L, rc, what):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise RuntimeError(f"{what}: {L.pihm_b200_last_error().decode()}")
    return rc


class Vec:
    """Device-resident N_Vector (internal element order)."""

    def __init__(self, model: "Model", handle=None):
        self.model = model
        self.L = model.L
        self.h = handle if handle is not None else self.L.pihm_b200_vec_new(model.h)
        if not self.h:
            raise RuntimeError("pihm_b200_vec_new: " + self.L.pihm_b200_last_error().decode())
        self.n = int(self.L.pihm_b200_vec_length(self.h))

    def upload(self, host):
        a = np.ascontiguousarray(host, np.float64)
        assert a.shape == (self.n,)
        _check(self.L, self.L.pihm_b200_vec_upload(self.h, _ptr(a)), "vec_upload")
        return self

    def download(self):
        out = np.empty(self.n)
        _check(self.L, self.L.pihm_b200_vec_download(self.h, _ptr(out)), "vec_download")
        return out

    def free(self):
        if self.h:
            self.L.pihm_b200_vec_free(self.h)
            self.h = None


class Model:
    """Device mirror of pihm->elem / pihm->river: pihm_b200_create + the RHS."""

    def __init__(self, tables: dict, device: int = 0, reorder: int = 0):
        """tables: a whole watershed, or one part from partition.partition() (it then carries
        nown_elem / nown_riv and the exchange maps; y vectors hold the owned unknowns only)"""
        self.L = L = load_library()
        if L.pihm_b200_device_count() <= device:
            raise RuntimeError("no CUDA device visible: libpihm_b200 has no CPU path")
        m = MeshStruct()
        m.nelem, m.nriver = int(tables["nelem"]), int(tables["nriver"])
        m.fbr = int(tables["fbr"])
        m.surf_mode, m.riv_mode = int(tables["surf_mode"]), int(tables["riv_mode"])
        m.stepsize = float(tables["stepsize"])
        keep = []
        for key, dt, ncol, n in (("elem_f64", np.float64, W.E_NCOL, m.nelem),
                                 ("elem_i32", np.int32, W.EI_NCOL, m.nelem),
                                 ("riv_f64", np.float64, W.R_NCOL, m.nriver),
                                 ("riv_i32", np.int32, W.RI_NCOL, m.nriver)):
            a = np.ascontiguousarray(tables[key], dtype=dt)
            assert a.shape == (ncol, n), (key, a.shape)
            keep.append(a)
            setattr(m, key, a.ctypes.data)
        self.part = tables if "nown_elem" in tables else None
        if self.part is None:
            self.h = L.pihm_b200_create(C.byref(m), device, reorder)
        else:
            self.h = L.pihm_b200_create_part(C.byref(m), device, 0, int(tables["nown_elem"]),
                                             int(tables["nown_riv"]))
        if not self.h:
            raise RuntimeError("pihm_b200_create: " + L.pihm_b200_last_error().decode())
        self.nelem, self.nriver, self.fbr = m.nelem, m.nriver, bool(m.fbr)
        self.nown_elem = int(tables.get("nown_elem", m.nelem))
        self.nown_riv = int(tables.get("nown_riv", m.nriver))
        if self.part is not None:
            t = tables
            arrs = [np.ascontiguousarray(t[k], np.int32) for k in
                    ("nbr_rank", "send_e_ptr", "send_e_idx", "recv_e_cnt", "send_r_ptr", "send_r_idx", "recv_r_cnt")]
            _check(L, L.pihm_b200_set_halo(self.h, len(t["nbr_rank"]), *[_ptr(a) for a in arrs]), "set_halo")
        self.nsv = int(L.pihm_b200_num_state_var(self.h))       # NumStateVar()

    # lifecycle ---------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.pihm_b200_transfer_release(self.h)
            self.L.pihm_b200_destroy(self.h)
            self.h = None

    # pipelined host transfers (csrc/transfer.cu); every host array must be pinned and stay alive
    def forcing_prefetch(self, cols, arrays):
        """start the upload of forcing columns `cols` (ids) from the pinned arrays `arrays` ([nelem] each)"""
        n = len(cols)
        ids = (C.c_int * n)(*[int(c) for c in cols])
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
        for a in arrays:
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == self.nelem
        _check(self.L, self.L.pihm_b200_forcing_prefetch(self.h, n, ids, ptrs), "forcing_prefetch")
        self._prefetch_keep = list(arrays)      # the copy reads them until the columns are committed

    def forcing_commit(self):
        _check(self.L, self.L.pihm_b200_forcing_commit(self.h), "forcing_commit")

    def download_async(self, v, host):
        assert host.dtype == np.float64 and host.flags["C_CONTIGUOUS"] and host.size == self.nsv
        _check(self.L, self.L.pihm_b200_vec_download_async(v.h, host.ctypes.data), "vec_download_async")

    def transfer_wait(self):
        _check(self.L, self.L.pihm_b200_transfer_wait(self.h), "transfer_wait")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # multi-GPU ---------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        _check(L, L.pihm_b200_comm_unique_id(buf), "comm_unique_id")
        return buf.raw

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        _check(self.L, self.L.pihm_b200_comm_init(self.h, rank, nranks, buf), "comm_init")

    @staticmethod
    def comm_init_local(models):
        """ranks of one process (models[r] = rank r): peer-memory halo + in-kernel all-reduce, no NCCL"""
        L = models[0].L
        arr = (C.c_void_p * len(models))(*[m.h for m in models])
        _check(L, L.pihm_b200_comm_init_local(arr, len(models)), "comm_init_local")

    def comm_paths(self) -> dict:
        out = np.zeros(2, np.int32)
        _check(self.L, self.L.pihm_b200_comm_paths(self.h, _ptr(out)), "comm_paths")
        return {"halo": "p2p" if out[0] else "nccl", "same_process": bool(out[1])}

    @property
    def nsv_global(self) -> int:
        return int(self.L.pihm_b200_num_state_var_global(self.h))

    def halo_pack_host(self, y: "Vec"):
        nse, nsr = len(self.part["send_e_idx"]), len(self.part["send_r_idx"])
        gs = 3 if self.fbr else 2
        e = np.zeros(max(nse, 1) * gs); r = np.zeros(max(nsr, 1) * 2)
        _check(self.L, self.L.pihm_b200_halo_pack_host(self.h, y.h, _ptr(e), _ptr(r)), "halo_pack_host")
        return e[:nse * gs], r[:nsr * 2]

    def set_ghosts(self, elem_rec, riv_rec):
        e = np.ascontiguousarray(elem_rec, np.float64); r = np.ascontiguousarray(riv_rec, np.float64)
        _check(self.L, self.L.pihm_b200_set_ghosts(self.h, _ptr(e) if e.size else None,
                                                   _ptr(r) if r.size else None), "set_ghosts")

    def set_stream(self, cuda_stream_ptr: int):
        _check(self.L, self.L.pihm_b200_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def synchronize(self):
        _check(self.L, self.L.pihm_b200_synchronize(self.h), "synchronize")

    def slow_path_count(self) -> int:
        return int(self.L.pihm_b200_slow_path_count(self.h))

    @property
    def launches(self) -> int:
        return int(self.L.pihm_b200_launch_count(self.h))

    def permutation(self):
        p = np.empty(self.nelem, np.int32)
        self.L.pihm_b200_get_permutation(self.h, _ptr(p))
        return p

    # per-step pushes ---------------------------------------------------------
    def set_forcing(self, forc, rivbc=None):
        f = np.ascontiguousarray(forc, np.float64)
        assert f.shape == (W.F_NCOL, self.nelem)
        _check(self.L, self.L.pihm_b200_set_forcing(self.h, _ptr(f)), "set_forcing")
        if rivbc is not None and self.nriver:
            rb = np.ascontiguousarray(rivbc, np.float64)
            _check(self.L, self.L.pihm_b200_set_river_bc(self.h, _ptr(rb)), "set_river_bc")

    def set_forcing_col(self, col: int, values):
        v = np.ascontiguousarray(values, np.float64)
        assert v.shape == (self.nelem,)
        _check(self.L, self.L.pihm_b200_set_forcing_col(self.h, col, _ptr(v)), "set_forcing_col")

    def set_stale_ovlflow(self, ovl):
        o = np.ascontiguousarray(ovl, np.float64)
        assert o.shape == (3, self.nelem)
        _check(self.L, self.L.pihm_b200_set_stale_ovlflow(self.h, _ptr(o)), "set_stale_ovlflow")

    def Summary(self, y: "Vec"):
        """ws0.surf <- y[SURF] on the device (the RHS-relevant part of Summary, src/update.c:47)"""
        _check(self.L, self.L.pihm_b200_summary(self.h, y.h), "summary")

    # Summary() + MassBalance() on the device (src/update.c:3-160) ---------------
    def set_diagnostics(self, on: bool = True):
        _check(self.L, self.L.pihm_b200_set_diagnostics(self.h, int(on)), "set_diagnostics")

    def set_ws0(self, y: "Vec"):
        """elem[i].ws0 = elem[i].ws of InitVar (src/initialize.c:598,612)"""
        _check(self.L, self.L.pihm_b200_set_ws0(self.h, y.h), "set_ws0")

    def SummaryMB(self, y: "Vec", stepsize: float):
        """void Summary(elem, river, CV_Y, stepsize) (src/update.c:3): mass-balance wf.infil /
        wf.fbr_infil from the fluxes of the last RHS call, ws0 <- y; everything stays on the device."""
        _check(self.L, self.L.pihm_b200_summary_mb(self.h, y.h, float(stepsize)), "summary_mb")

    def get_summary(self):
        """-> (subrunoff [nelem], ws0 [N]) in reference order"""
        sr = np.zeros(self.nelem); w = np.zeros(self.nsv)
        _check(self.L, self.L.pihm_b200_get_summary(self.h, _ptr(sr), _ptr(w)), "get_summary")
        return sr, w

    # ApplyMeteoForc/ApplyLai scatter + IntcpSnowEt on the device (src/forcing.c, src/is_sm_et.c) ---
    def et_create(self, et_f64, et_i32):
        f = np.ascontiguousarray(et_f64, np.float64); ii = np.ascontiguousarray(et_i32, np.int32)
        assert f.shape == (W.ET_NCOL, self.nelem) and ii.shape == (W.ETI_NCOL, self.nelem)
        _check(self.L, self.L.pihm_b200_et_create(self.h, _ptr(f), _ptr(ii)), "et_create")

    def IntcpSnowEt(self, step: "EtStep", y: "Vec"):
        """void IntcpSnowEt(t, stepsize, elem, cal) (src/is_sm_et.c:4) incl. the per-element forcing
        assignments; writes wf.pcpdrp / edir / ett into the forcing columns of the RHS on the device"""
        _check(self.L, self.L.pihm_b200_intcp_snow_et(self.h, C.byref(step), y.h), "intcp_snow_et")

    def et_set_state(self, sneqv, cmc):
        a = np.ascontiguousarray(sneqv, np.float64); b = np.ascontiguousarray(cmc, np.float64)
        assert a.shape == b.shape == (self.nelem,)
        _check(self.L, self.L.pihm_b200_et_set_state(self.h, _ptr(a), _ptr(b)), "et_set_state")

    def et_get(self):
        out = np.zeros((W.EO_NCOL, self.nelem))
        _check(self.L, self.L.pihm_b200_et_get(self.h, _ptr(out)), "et_get")
        return out

    # print accumulation on the device (UpdPrintVar / PrintData, src/print.c:171-251) -----------
    def print_add(self, src: int, column: int) -> int:
        rc = self.L.pihm_b200_print_add(self.h, int(src), int(column))
        _check(self.L, min(rc, 0), "print_add")
        return rc

    def UpdPrintVar(self, ids, y: "Vec" = None):
        a = np.ascontiguousarray(ids, np.int32)
        _check(self.L, self.L.pihm_b200_print_update(self.h, _ptr(a), len(a), y.h if y is not None else None),
               "print_update")

    def PrintData(self, vid: int, river: bool = False):
        """-> (average over the updates since the last record, number of updates)"""
        out = np.zeros(self.nriver if river else self.nelem); cnt = C.c_int32(0)
        _check(self.L, self.L.pihm_b200_print_data(self.h, int(vid), _ptr(out), C.byref(cnt)), "print_data")
        return out, cnt.value

    # output files fed from the device (InitOutputFile / PrintData / PrintInit, src/print.c:72-313) ----
    def print_open(self, vid: int, name: str, ascii: bool = True, append: bool = False):
        _check(self.L, self.L.pihm_b200_print_open(self.h, int(vid), name.encode(), int(ascii), int(append)), "print_open")

    def print_write(self, ids, t: int, timestr: str):
        """PrintData for the variables that are due at model time t: one kernel, one D2H copy, records appended"""
        a = np.ascontiguousarray(ids, np.int32)
        _check(self.L, self.L.pihm_b200_print_write(self.h, _ptr(a), len(a), int(t), timestr.encode()), "print_write")

    def print_close(self):
        self.L.pihm_b200_print_close(self.h)

    def print_io_stats(self):
        n, b = C.c_int64(0), C.c_int64(0)
        self.L.pihm_b200_print_io_stats(self.h, C.byref(n), C.byref(b))
        return n.value, b.value

    def write_ic(self, path: str, y: "Vec", cmc=None, sneqv=None):
        """PrintInit (src/print.c:253-313): the restart file of state y"""
        c = np.ascontiguousarray(cmc, np.float64) if cmc is not None else None
        s = np.ascontiguousarray(sneqv, np.float64) if sneqv is not None else None
        _check(self.L, self.L.pihm_b200_write_ic(self.h, path.encode(), y.h, _ptr(c) if c is not None else None,
                                                 _ptr(s) if s is not None else None), "write_ic")

    def set_flux_recording(self, on: bool):
        _check(self.L, self.L.pihm_b200_set_flux_recording(self.h, int(on)), "set_flux_recording")

    # RHS -----------------------------------------------------------------------
    def ODE(self, t, y, ydot=None):
        """int ODE(t, y, ydot, pihm) with host buffers (src/ode.c:3)."""
        y = np.ascontiguousarray(y, np.float64)
        assert y.shape == (self.nsv,)
        if ydot is None:
            ydot = np.empty(self.nsv)
        rc = self.L.pihm_b200_ode_host(self.h, float(t), _ptr(y), _ptr(ydot))
        _check(self.L, rc, "ode_host")
        self.nan_flag = rc
        return ydot

    def ode_dev(self, t, y: Vec, ydot: Vec):
        _check(self.L, self.L.pihm_b200_ode(self.h, float(t), y.h, ydot.h), "ode")

    def check_nan(self) -> int:
        return int(self.L.pihm_b200_check_nan(self.h))

    def get_fluxes(self, elem=True):
        xf = np.zeros((W.X_NCOL, self.nelem)) if elem else None
        rf = np.zeros((W.NUM_RIVFLX, max(self.nriver, 1)))
        _check(self.L, self.L.pihm_b200_get_fluxes(self.h, None if xf is None else _ptr(xf), _ptr(rf)),
               "get_fluxes")
        return xf, rf[:, :self.nriver]

    # vectors -------------------------------------------------------------------
    def N_VNew(self, host=None) -> Vec:
        v = Vec(self)
        if host is not None:
            v.upload(host)
        return v

    def N_VLinearSum(self, a, x, b, y, z): self.L.pihm_b200_nv_linearsum(a, x.h, b, y.h, z.h)
    def N_VConst(self, c, z): self.L.pihm_b200_nv_const(c, z.h)
    def N_VProd(self, x, y, z): self.L.pihm_b200_nv_prod(x.h, y.h, z.h)
    def N_VDiv(self, x, y, z): self.L.pihm_b200_nv_div(x.h, y.h, z.h)
    def N_VScale(self, c, x, z): self.L.pihm_b200_nv_scale(c, x.h, z.h)
    def N_VAbs(self, x, z): self.L.pihm_b200_nv_abs(x.h, z.h)
    def N_VInv(self, x, z): self.L.pihm_b200_nv_inv(x.h, z.h)
    def N_VAddConst(self, x, b, z): self.L.pihm_b200_nv_addconst(x.h, b, z.h)
    def N_VDotProd(self, x, y): return float(self.L.pihm_b200_nv_dotprod(x.h, y.h))
    def N_VMaxNorm(self, x): return float(self.L.pihm_b200_nv_maxnorm(x.h))
    def N_VWrmsNorm(self, x, w): return float(self.L.pihm_b200_nv_wrmsnorm(x.h, w.h))
    def N_VMin(self, x): return float(self.L.pihm_b200_nv_min(x.h))


class Cvode:
    """SetCVodeParam / SolveCVode / AdjCVodeMaxStep flow (src/ode.c:340-560)
    on the device-resident BDF/Newton/SPGMR integrator."""

    def __init__(self, model: Model):
        self.model = model
        self.L = model.L
        self.h = self.L.pihm_b200_cvode_create(model.h)
        if not self.h:
            raise RuntimeError("pihm_b200_cvode_create: " + self.L.pihm_b200_last_error().decode())
        self.ctrl = None

    def close(self):
        if getattr(self, "h", None):
            self.L.pihm_b200_cvode_destroy(self.h)
            self.h = None

    def SetCVodeParam(self, y: Vec, reltol=1e-3, abstol=1e-4, initstep=5e-5, stepsize=60.0,
                      stmin=1.0, nncfn=0.0, nnimax=3.0, nnimin=1.0, decr=1.2, incr=1.2, t0=0.0):
        p = CvodeParam(reltol=reltol, abstol=abstol, initstep=initstep, maxstep=stepsize,
                       mxsteps=int(stepsize * 10), stab_lim_det=1, maxl=0)
        _check(self.L, self.L.pihm_b200_cvode_init(self.h, C.byref(p), float(t0), y.h), "cvode_init")
        self.ctrl = MaxStepCtrl(maxstep=stepsize, stepsize=stepsize, stmin=stmin, nncfn=nncfn,
                                nnimax=nnimax, nnimin=nnimin, decr=decr, incr=incr)

    def SolveCVode(self, tout: float, y: Vec) -> float:
        tret = C.c_double(0.0)
        rc = self.L.pihm_b200_cvode_solve(self.h, float(tout), y.h, C.byref(tret))
        if rc < 0:
            raise RuntimeError(f"SolveCVode failed with flag {rc}: "
                               + self.L.pihm_b200_last_error().decode())
        return tret.value

    def AdjCVodeMaxStep(self):
        _check(self.L, self.L.pihm_b200_adj_cvode_max_step(self.h, C.byref(self.ctrl)), "adj_max_step")
        return self.ctrl.maxstep

    def profile(self, on=True):
        """on = 2: also CUDA events around every vector kernel (get_kernel_profile)"""
        _check(self.L, self.L.pihm_b200_cvode_profile(self.h, int(on)), "cvode_profile")

    def get_kernel_profile(self) -> dict:
        """{kernel: dict(ms, bytes, launches)} since profile(2): event-to-event time and algorithmic bytes"""
        cap = 32
        names = C.create_string_buffer(24 * cap)
        ms, by = np.zeros(cap), np.zeros(cap)
        n_l = np.zeros(cap, dtype=np.int64)
        n = self.L.pihm_b200_cvode_get_kernel_profile(self.h, cap, names, _ptr(ms), _ptr(by), _ptr(n_l))
        if n < 0:
            raise RuntimeError("cvode_get_kernel_profile failed")
        out = {}
        for k in range(n):
            nm = names.raw[24 * k:24 * (k + 1)].split(b"\0")[0].decode()
            out[nm] = dict(ms=float(ms[k]), bytes=float(by[k]), launches=int(n_l[k]))
        return out

    def get_profile(self) -> dict:
        out = np.zeros(5)
        _check(self.L, self.L.pihm_b200_cvode_get_profile(self.h, _ptr(out)), "cvode_get_profile")
        return dict(solve_ms=out[0], host_wait_ms=out[1], host_syncs=int(out[2]), rhs_ms=out[3], rhs_evals=int(out[4]))

    def stats(self) -> dict:
        st = CvodeStats()
        self.L.pihm_b200_cvode_get_stats(self.h, C.byref(st))
        return {k: getattr(st, k) for k, _ in CvodeStats._fields_}
