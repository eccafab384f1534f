"""Mesh partitioning for the multi-GPU path: ctypes front-end of
pihm_b200_partition_* (csrc/partition.cpp).  Pure host code, no GPU needed.

Each part is again a tables dict (local numbering, owned entities first, ghosts
grouped by owner rank) plus the exchange maps of the per-RHS halo exchange."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import watershed as W
from .lib import MeshStruct, load_library


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _mesh_struct(tables):
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
    return m, keep


def state_index(ne_own, nr_own, fbr, elem_gid, riv_gid, ne_glob, nr_glob):
    """global positions (block layout of pihm_func.h:7-15) of a part's owned unknowns,
    in the part's local block layout"""
    eg = np.asarray(elem_gid[:ne_own], np.int64)
    rg = np.asarray(riv_gid[:nr_own], np.int64)
    idx = [eg, ne_glob + eg, 2 * ne_glob + eg, 3 * ne_glob + rg, 3 * ne_glob + nr_glob + rg]
    if fbr:
        idx += [3 * ne_glob + 2 * nr_glob + eg, 4 * ne_glob + 2 * nr_glob + eg]
    return np.concatenate(idx)


def partition(tables: dict, nparts: int, parts=None) -> list:
    """Split a watershed into `nparts` local meshes.  Returns one dict per part
    (or only for the part numbers listed in `parts`) with the local tables, the
    owned counts, global ids, exchange maps, and `state_idx`: where the part's
    owned unknowns sit in the global state vector."""
    L = load_library()
    L.pihm_b200_partition_create.restype = C.c_void_p
    L.pihm_b200_partition_create.argtypes = [C.c_void_p, C.c_int]
    L.pihm_b200_partition_destroy.argtypes = [C.c_void_p]
    L.pihm_b200_partition_sizes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.pihm_b200_partition_fill.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 13
    m, keep = _mesh_struct(tables)
    P = L.pihm_b200_partition_create(C.byref(m), int(nparts))
    if not P:
        raise RuntimeError("pihm_b200_partition_create: " + L.pihm_b200_last_error().decode())
    out = []
    fbr = bool(tables["fbr"])
    for p in (range(nparts) if parts is None else parts):
        sz = np.zeros(8, np.int32)
        L.pihm_b200_partition_sizes(P, p, _ptr(sz))
        nl, rl, no, ro, nn, nse, nsr = (int(v) for v in sz[:7])
        ef = np.zeros((W.E_NCOL, nl)); ei = np.zeros((W.EI_NCOL, nl), np.int32)
        rf = np.zeros((W.R_NCOL, rl)); ri = np.zeros((W.RI_NCOL, rl), np.int32)
        eg = np.zeros(nl, np.int32); rg = np.zeros(max(rl, 1), np.int32)
        nbr = np.zeros(max(nn, 1), np.int32)
        sep = np.zeros(nn + 1, np.int32); sei = np.zeros(max(nse, 1), np.int32); rec = np.zeros(max(nn, 1), np.int32)
        srp = np.zeros(nn + 1, np.int32); sri = np.zeros(max(nsr, 1), np.int32); rrc = np.zeros(max(nn, 1), np.int32)
        L.pihm_b200_partition_fill(P, p, _ptr(ef), _ptr(ei), _ptr(rf), _ptr(ri), _ptr(eg), _ptr(rg), _ptr(nbr),
                                   _ptr(sep), _ptr(sei), _ptr(rec), _ptr(srp), _ptr(sri), _ptr(rrc))
        d = dict(nelem=nl, nriver=rl, fbr=int(fbr), surf_mode=tables["surf_mode"], riv_mode=tables["riv_mode"],
                 stepsize=tables["stepsize"], elem_f64=ef, elem_i32=ei, riv_f64=rf, riv_i32=ri,
                 nown_elem=no, nown_riv=ro, elem_gid=eg, riv_gid=rg[:rl], nbr_rank=nbr[:nn],
                 send_e_ptr=sep, send_e_idx=sei[:nse], recv_e_cnt=rec[:nn],
                 send_r_ptr=srp, send_r_idx=sri[:nsr], recv_r_cnt=rrc[:nn], part=p, nparts=nparts,
                 _ne_glob=int(tables["nelem"]), _nr_glob=int(tables["nriver"]))
        d["state_idx"] = state_index(no, ro, fbr, eg, d["riv_gid"], tables["nelem"], tables["nriver"])
        # element-wise helper arrays of the global tables restricted to this part
        for key in ("xc", "yc"):
            if key in tables:
                d[key] = np.asarray(tables[key])[eg]
        if "bc_head" in tables:
            d["bc_head"] = np.asarray(tables["bc_head"])[:, eg]
        out.append(d)
    L.pihm_b200_partition_destroy(P)
    return out


def local_state(part: dict, y_global: np.ndarray, extended: bool = False) -> np.ndarray:
    """owned part of a global state vector; with extended=True the vector of the
    whole local mesh (owned + ghosts), as a plain local model would see it"""
    if not extended:
        return np.ascontiguousarray(y_global[part["state_idx"]])
    idx = state_index(part["nelem"], part["nriver"], bool(part["fbr"]), part["elem_gid"], part["riv_gid"],
                      _glob(part, "ne"), _glob(part, "nr"))
    return np.ascontiguousarray(y_global[idx])


def _glob(part, what):
    return part["_ne_glob"] if what == "ne" else part["_nr_glob"]


def owned_in_local(part: dict) -> np.ndarray:
    """positions of the owned unknowns inside the state vector of the whole local mesh (owned + ghosts)"""
    return state_index(part["nown_elem"], part["nown_riv"], bool(part["fbr"]), np.arange(part["nelem"]),
                       np.arange(part["nriver"]), part["nelem"], part["nriver"])


def ghost_records(part: dict, y_global: np.ndarray):
    """the ghost records a halo exchange would deliver, {surf, gw[, fbr_gw]} per ghost element and
    {stage, gw} per ghost river in local (= receive) order, taken from a global state vector"""
    ne, nr, fbr = _glob(part, "ne"), _glob(part, "nr"), bool(part["fbr"])
    ge = np.asarray(part["elem_gid"][part["nown_elem"]:], np.int64)
    gr = np.asarray(part["riv_gid"][part["nown_riv"]:], np.int64)
    cols = [y_global[ge], y_global[2 * ne + ge]]
    if fbr:
        cols.append(y_global[4 * ne + 2 * nr + ge])
    erec = np.ascontiguousarray(np.stack(cols, axis=1)).ravel()
    rrec = np.ascontiguousarray(np.stack([y_global[3 * ne + gr], y_global[3 * ne + nr + gr]], axis=1)).ravel()
    return erec, rrec
