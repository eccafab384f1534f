"""ctypes front-end of oracle/libpihm_oracle.so (the plain-C restatement of the
reference RHS and serial N_Vector arithmetic) -- TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg only; the product never routes through it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from reflib import (MeshStruct, NUM_RIVFLX, PB_E_NCOL, PB_EI_NCOL, PB_F_NCOL,
                    PB_R_NCOL, PB_RI_NCOL, PB_X_NCOL)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpihm_oracle.so")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise FileNotFoundError(f"{LIB} missing: run `make -C oracle port`")
        L = C.CDLL(LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_void_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_num_state_var.restype = C.c_int64
        L.oracle_num_state_var.argtypes = [C.c_void_p]
        for f in ("oracle_set_forcing", "oracle_set_river_bc", "oracle_set_stale_ovlflow"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_fluxes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for f in ("oracle_set_ws0", "oracle_get_ws0"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_summary.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.oracle_intcp_snow_et.argtypes = [C.c_void_p] * 6
        L.oracle_ode.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_nv_linearsum.argtypes = [C.c_int64, C.c_double, C.c_void_p, C.c_double,
                                          C.c_void_p, C.c_void_p]
        L.oracle_nv_scale.argtypes = [C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
        for f in ("oracle_nv_dotprod", "oracle_nv_wrmsnorm"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
        for f in ("oracle_nv_maxnorm", "oracle_nv_min"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def mesh_struct(tables: dict):
    """-> (MeshStruct, keepalive list) for a tables dict (see reflib.pack_tables)."""
    m = MeshStruct()
    m.nelem, m.nriver = int(tables["nelem"]), int(tables["nriver"])
    m.fbr = int(tables["fbr"])
    m.surf_mode, m.riv_mode = int(tables["surf_mode"]), int(tables["riv_mode"])
    m.stepsize = float(tables["stepsize"])
    keep = []
    for key, dt, ncol, n in (("elem_f64", np.float64, PB_E_NCOL, m.nelem),
                             ("elem_i32", np.int32, PB_EI_NCOL, m.nelem),
                             ("riv_f64", np.float64, PB_R_NCOL, m.nriver),
                             ("riv_i32", np.int32, PB_RI_NCOL, m.nriver)):
        a = np.ascontiguousarray(tables[key], dtype=dt)
        assert a.shape == (ncol, n), (key, a.shape, (ncol, n))
        keep.append(a)
        setattr(m, key, a.ctypes.data)
    return m, keep


class OracleModel:
    def __init__(self, tables: dict):
        self.L = lib()
        m, keep = mesh_struct(tables)
        self.h = self.L.oracle_create(C.byref(m))
        self.nelem, self.nriver, self.fbr = m.nelem, m.nriver, bool(m.fbr)
        self.nsv = int(self.L.oracle_num_state_var(self.h))

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_forcing(self, forc, rivbc=None):
        f = np.ascontiguousarray(forc, np.float64)
        assert f.shape == (PB_F_NCOL, self.nelem)
        self.L.oracle_set_forcing(self.h, _ptr(f))
        if rivbc is not None and self.nriver:
            rb = np.ascontiguousarray(rivbc, np.float64)
            assert rb.shape == (self.nriver,)
            self.L.oracle_set_river_bc(self.h, _ptr(rb))

    def set_stale_ovlflow(self, ovl):
        o = np.ascontiguousarray(ovl, np.float64)
        assert o.shape == (3, self.nelem)
        self.L.oracle_set_stale_ovlflow(self.h, _ptr(o))

    def ode(self, y, t=0.0):
        y = np.ascontiguousarray(y, np.float64)
        assert y.shape == (self.nsv,)
        dy = np.empty(self.nsv)
        self.nan_flag = self.L.oracle_ode(self.h, float(t), _ptr(y), _ptr(dy))
        return dy

    def get_fluxes(self):
        xf = np.zeros((PB_X_NCOL, self.nelem))
        rf = np.zeros((NUM_RIVFLX, max(self.nriver, 1)))
        self.L.oracle_get_fluxes(self.h, _ptr(xf), _ptr(rf))
        return xf, rf[:, :self.nriver]


    # -- IntcpSnowEt + per-element forcing assignment, src/is_sm_et.c / src/forcing.c --
    def intcp_snow_et(self, step, et_f64, et_i32, y, state):
        """step: ctypes struct pihm_b200_et_step (by reference); state [EO_NCOL, nelem] carries
        sneqv / cmc in; returns the updated [EO_NCOL, nelem] table"""
        f = np.ascontiguousarray(et_f64, np.float64); ii = np.ascontiguousarray(et_i32, np.int32)
        y = np.ascontiguousarray(y, np.float64)
        out = np.array(state, np.float64, order="C")
        self.L.oracle_intcp_snow_et(self.h, C.byref(step), _ptr(f), _ptr(ii), _ptr(y), _ptr(out))
        return out

    # -- Summary() + MassBalance(), src/update.c ---------------------------------
    def set_ws0(self, y):
        y = np.ascontiguousarray(y, np.float64); assert y.shape == (self.nsv,)
        self.L.oracle_set_ws0(self.h, _ptr(y))

    def get_ws0(self):
        y = np.zeros(self.nsv); self.L.oracle_get_ws0(self.h, _ptr(y)); return y

    def summary(self, y, stepsize):
        """-> subrunoff [nelem]; get_fluxes() then shows the mass-balance infil."""
        y = np.ascontiguousarray(y, np.float64); assert y.shape == (self.nsv,)
        sr = np.zeros(self.nelem)
        self.L.oracle_summary(self.h, _ptr(y), float(stepsize), _ptr(sr))
        return sr


class PrintVarOracle:
    """numpy restatement of one varctrl_struct of the reference's print system:
    UpdPrintVar (src/print.c:171-190): buffer[j] += *var[j]; counter++
    PrintData   (src/print.c:230-246): buffer[j] / (double)counter (buffer[j] if counter == 0), then reset."""

    def __init__(self, n):
        self.buffer = np.zeros(n)
        self.counter = 0

    def update(self, values):
        self.buffer = self.buffer + np.asarray(values, np.float64)
        self.counter += 1

    def data(self):
        out = self.buffer / float(self.counter) if self.counter > 0 else self.buffer.copy()
        n = self.counter
        self.buffer = np.zeros_like(self.buffer)
        self.counter = 0
        return out, n


# serial N_Vector arithmetic -------------------------------------------------
def nv_linearsum(a, x, b, y, inplace=None):
    """z = a x + b y with nvector_serial.c's special-case rounding.
    inplace: None (fresh z), 'x' (z is x) or 'y' (z is y)."""
    L = lib()
    x = np.array(x, np.float64); y = np.array(y, np.float64)
    z = x if inplace == "x" else (y if inplace == "y" else np.empty_like(x))
    L.oracle_nv_linearsum(len(x), float(a), _ptr(x), float(b), _ptr(y), _ptr(z))
    return z


def nv_scale(c, x, inplace=False):
    L = lib()
    x = np.array(x, np.float64)
    z = x if inplace else np.empty_like(x)
    L.oracle_nv_scale(len(x), float(c), _ptr(x), _ptr(z))
    return z


def nv_dotprod(x, y):
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    return float(lib().oracle_nv_dotprod(len(x), _ptr(x), _ptr(y)))


def nv_wrmsnorm(x, w):
    x = np.ascontiguousarray(x, np.float64); w = np.ascontiguousarray(w, np.float64)
    return float(lib().oracle_nv_wrmsnorm(len(x), _ptr(x), _ptr(w)))


def nv_maxnorm(x):
    x = np.ascontiguousarray(x, np.float64)
    return float(lib().oracle_nv_maxnorm(len(x), _ptr(x)))


def nv_min(x):
    x = np.ascontiguousarray(x, np.float64)
    return float(lib().oracle_nv_min(len(x), _ptr(x)))
