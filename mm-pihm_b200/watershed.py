"""Synthetic triangulated watersheds in the column-table form of
include/pihm_b200.h.

A rectangle of nx x ny square cells (dx = 20 m) is split into 2 CCW triangles
per cell; a V-shaped valley carries a main-stem river along the grid line
y = yc from x = Lx down to the outlet at x = 0, with tributaries running down
the valley sides along vertical grid lines.  Geometry follows what the
reference computes from a .mesh/.riv pair:
  element area/centroid/edges      InitTopo   (src/init_topo.c:14-44)
  neighbour centroids / distances  InitSurfL  (src/initialize.c:365-438)
  river segment geometry, banks    InitRiver  (src/init_river.c:10-116)
  relaxed initial condition        RelaxIc    (src/initialize.c:476-553)
Soil / land-cover / river-material values are the calibrated values the
reference derives for input/example (5 soil classes, geology class 1,
rectangular channel), so the physics runs in its usual regime.

`nabr[j]` is the element across the edge OPPOSITE node j (edge j joins nodes
j+1 and j+2), which is the .mesh convention (src/init_topo.c:38-43).
"""
from __future__ import annotations

import numpy as np

# column ids -- must match include/pihm_b200.h (checked by tests/test_abi.py)
(E_AREA, E_ZMIN, E_ZMAX, E_ZBED, E_EDGE0, E_EDGE1, E_EDGE2, E_NABRDIST0,
 E_NABRDIST1, E_NABRDIST2, E_NABRX0, E_NABRX1, E_NABRX2, E_NABRY0, E_NABRY1,
 E_NABRY2, E_DEPTH, E_KSATH, E_KSATV, E_KINFV, E_DINF, E_ALPHA, E_BETA,
 E_POROSITY, E_DMAC, E_KMACH, E_KMACV, E_AREAFV, E_AREAFH, E_ROUGH, E_RZD,
 E_GDEPTH, E_GKSATH, E_GKSATV, E_GALPHA, E_GBETA, E_GPOROSITY, E_NCOL) = range(38)
(EI_NABR0, EI_NABR1, EI_NABR2, EI_BC0, EI_BC1, EI_BC2, EI_FBRBC0, EI_FBRBC1,
 EI_FBRBC2, EI_NCOL) = range(10)
(F_PCPDRP, F_EDIR, F_ETT, F_WS0SURF, F_BC0, F_BC1, F_BC2, F_FBRBC0, F_FBRBC1,
 F_FBRBC2, F_NCOL) = range(11)
(R_AREA, R_ZMIN, R_ZMAX, R_ZBED, R_NODE_ZMAX, R_DIST_LEFT, R_DIST_RIGHT,
 R_SHP_DEPTH, R_SHP_COEFF, R_SHP_LENGTH, R_SHP_WIDTH, R_ROUGH, R_CWR, R_KSATH,
 R_KSATV, R_BEDTHICK, R_POROSITY, R_NCOL) = range(18)
(RI_LEFTELE, RI_RIGHTELE, RI_DOWN, RI_BCTYPE, RI_INTRPL_ORD, RI_NCOL) = range(6)
(X_OVL0, X_OVL1, X_OVL2, X_SUB0, X_SUB1, X_SUB2, X_INFIL, X_RECHG,
 X_EDIR_SURF, X_EDIR_UNSAT, X_EDIR_GW, X_ETT_UNSAT, X_ETT_GW, X_FBR_INFIL,
 X_FBR_RECHG, X_FBRFLOW0, X_FBRFLOW1, X_FBRFLOW2, X_NCOL) = range(19)
# include/pihm_b200.h: pihm_b200_et_col / et_icol / et_out_col
(ET_ALBEDOMIN, ET_ALBEDOMAX, ET_CMCFACTR, ET_SHDFAC, ET_CFACTR, ET_RGL, ET_RSMIN, ET_RSMAX, ET_TOPT,
 ET_SMCMIN, ET_SMCWLT, ET_SMCREF, ET_ZLVL_WIND, ET_NCOL) = range(14)
(ETI_METEO_TYPE, ETI_LAI_TYPE, ETI_LC_TYPE, ETI_NCOL) = range(4)
(EO_PCPDRP, EO_EDIR, EO_ETT, EO_EC, EO_DRIP, EO_SNEQV, EO_CMC, EO_NCOL) = range(8)
NUM_METEO_VAR = 7
(PS_STATE, PS_ELEM_FLUX, PS_RIV_FLUX, PS_ET) = range(4)     # pihm_b200_print_src
NUM_RIVFLX = 11

# ksath ksatv kinfv dinf alpha beta porosity kmach kmacv areafv areafh
# (= InitSoil(example.soil x example.calib), src/init_soil.c:3-74)
SOIL_CLASSES = np.array([
    [4.901e-05, 3.724e-05, 1.516e-04, 0.1, 6.45, 1.2726, 0.30022, 0.034307, 0.01516, 0.05, 0.05],
    [6.010e-05, 3.250e-05, 9.098e-05, 0.1, 8.80, 1.3020, 0.27183, 0.042070, 0.009098, 0.05, 0.05],
    [1.1315e-04, 2.172e-05, 9.833e-05, 0.1, 6.50, 1.3209, 0.318495, 0.079205, 0.009833, 0.05, 0.05],
    [1.5235e-04, 1.3524e-05, 1.506e-05, 0.1, 5.34, 1.3230, 0.312715, 0.106645, 0.001506, 0.05, 0.05],
    [3.481e-04, 7.410e-05, 8.281e-05, 0.1, 5.82, 1.2810, 0.37638, 0.243670, 0.008281, 0.05, 0.05],
])
DMAC_TBL = 1.6                      # example: DMAC 1.0 x calib, clipped to depth (init_soil.c:57-59)
LC_CLASSES = np.array([[0.05, 0.6], [0.06, 1.0], [0.07, 1.0], [0.08, 1.0]])   # rough, rzd
GEOL = dict(ksath=6.01e-07, ksatv=3.25e-07, alpha=10.0, beta=2.1, porosity=0.03145)
RIV_SHAPE = dict(depth=0.5, order=1, coeff=2.0)      # rectangular channel
RIV_MATL = dict(rough=0.04, cwr=0.6, ksath=6.962e-05, ksatv=3.705e-05, bedthick=0.11)
DX = 20.0

SIZES = {            # name -> (nx, ny); triangles = 2 nx ny   
    "tiny": (8, 6), "small": (40, 30), "10k": (100, 50), "100k": (250, 200),
    "1M": (1000, 500), "2M": (1000, 1000), "4M": (2000, 1000), "8M": (2000, 2000),
}


def _hash2(i, j, mod):
    h = (i.astype(np.int64) * 73856093) ^ (j.astype(np.int64) * 19349663)
    return (h % mod).astype(np.int64)


def _river_eq_wid(order, depth, coeff):
    """RiverEqWid, src/init_river.c:118-145"""
    if order == 1:
        return coeff
    return 2.0 * (depth + 0.05) ** (1.0 / (order - 1.0)) / coeff ** (1.0 / (order - 1.0))


def make_watershed(nx: int, ny: int, fbr: bool = False, seed: int = 12345,
                   trib_every: int = 25, river: bool = True,
                   surf_mode: int = 2, riv_mode: int = 2, stepsize: float = 60.0,
                   dirichlet_edges: bool = False, riv_order: int = 1, keep_mesh: bool = False) -> dict:
    """Build the column tables of a synthetic watershed.

    Returns a dict: nelem, nriver, fbr, surf_mode, riv_mode, stepsize,
    elem_f64 [E_NCOL, nelem], elem_i32 [EI_NCOL, nelem], riv_f64 [R_NCOL, nriver],
    riv_i32 [RI_NCOL, nriver], y0 (RelaxIc state, block layout), plus `xc`, `yc`
    element centroids (used for partitioning / reordering).  keep_mesh=True adds `mesh`: the node-level
    description a .mesh / .att / .riv / .bedrock file set is written from (project_files.py)."""
    rng = np.random.default_rng(seed)
    nnx, nny = nx + 1, ny + 1
    ii, jj = np.meshgrid(np.arange(nnx), np.arange(nny), indexing="xy")   # [nny, nnx]
    xn = ii * DX
    yn = jj * DX
    jc = ny // 2
    ycl = jc * DX
    zmax_n = (300.0 + 0.02 * xn + 0.05 * np.abs(yn - ycl)
              + 0.5 * np.sin(2 * np.pi * xn / 400.0) * np.cos(2 * np.pi * yn / 400.0)
              + rng.uniform(-0.05, 0.05, size=xn.shape))
    d_n = (2.0 + 0.8 * np.sin(2 * np.pi * xn / 900.0) * np.cos(2 * np.pi * yn / 700.0)
           + rng.uniform(-0.2, 0.2, size=xn.shape))
    zmin_n = zmax_n - d_n
    zbed_n = zmin_n - 10.0

    def node(i, j):
        return j * nnx + i

    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ci = ci.ravel(); cj = cj.ravel()                      # cell c = cj*nx + ci
    n00, n10 = node(ci, cj), node(ci + 1, cj)
    n01, n11 = node(ci, cj + 1), node(ci + 1, cj + 1)
    ncell = nx * ny
    ne = 2 * ncell
    # element e = 2*c + k ; k=0: (n00,n10,n11)  k=1: (n00,n11,n01), both CCW
    nodes = np.empty((ne, 3), np.int64)
    nodes[0::2] = np.stack([n00, n10, n11], 1)
    nodes[1::2] = np.stack([n00, n11, n01], 1)
    xf, yf = xn.ravel(), yn.ravel()
    X = xf[nodes]; Y = yf[nodes]
    ef = np.zeros((E_NCOL, ne))
    ei = np.zeros((EI_NCOL, ne), np.int32)
    ef[E_AREA] = 0.5 * ((X[:, 1] - X[:, 0]) * (Y[:, 2] - Y[:, 0]) - (Y[:, 1] - Y[:, 0]) * (X[:, 2] - X[:, 0]))
    xc = (X[:, 0] + X[:, 1] + X[:, 2]) / 3.0
    yc = (Y[:, 0] + Y[:, 1] + Y[:, 2]) / 3.0
    for col, src in ((E_ZMIN, zmin_n), (E_ZMAX, zmax_n), (E_ZBED, zbed_n)):
        z = src.ravel()[nodes]
        ef[col] = (z[:, 0] + z[:, 1] + z[:, 2]) / 3.0
    ef[E_EDGE0] = np.sqrt((X[:, 1] - X[:, 2]) ** 2 + (Y[:, 1] - Y[:, 2]) ** 2)
    ef[E_EDGE1] = np.sqrt((X[:, 2] - X[:, 0]) ** 2 + (Y[:, 2] - Y[:, 0]) ** 2)
    ef[E_EDGE2] = np.sqrt((X[:, 0] - X[:, 1]) ** 2 + (Y[:, 0] - Y[:, 1]) ** 2)

    # neighbours (1-based, 0 = boundary)
    c = cj * nx + ci
    nabr = np.zeros((ne, 3), np.int64)
    # k=0: edge0 (n10-n11) right -> cell(i+1,j) k=1 ; edge1 diagonal -> same cell k=1 ;
    #      edge2 (n00-n10) bottom -> cell(i,j-1) k=1
    nabr[0::2, 0] = np.where(ci + 1 < nx, 2 * (c + 1) + 1 + 1, 0)
    nabr[0::2, 1] = 2 * c + 1 + 1
    nabr[0::2, 2] = np.where(cj > 0, 2 * (c - nx) + 1 + 1, 0)
    # k=1: edge0 (n11-n01) top -> cell(i,j+1) k=0 ; edge1 (n01-n00) left -> cell(i-1,j) k=0 ;
    #      edge2 diagonal -> same cell k=0
    nabr[1::2, 0] = np.where(cj + 1 < ny, 2 * (c + nx) + 1, 0)
    nabr[1::2, 1] = np.where(ci > 0, 2 * (c - 1) + 1, 0)
    nabr[1::2, 2] = 2 * c + 1

    nabr_mesh = nabr.copy() if keep_mesh else None      # as a .mesh file holds it: rivers not marked
    # ---- river network -------------------------------------------------
    segs = []        # (from_node, to_node, left_elem(1b), right_elem(1b), down(1b or code))
    if river and ny >= 2 and nx >= 2:
        # main stem along y = jc, flowing to -x: segment k covers column i = nx-1-k
        main_id = {}
        for k in range(nx):
            i = nx - 1 - k
            frm, to = node(i + 1, jc), node(i, jc)
            left = 2 * ((jc - 1) * nx + i) + 1 + 1      # cell(i,jc-1) k=1 (top edge)
            right = 2 * (jc * nx + i) + 0 + 1           # cell(i,jc)   k=0 (bottom edge)
            main_id[i] = len(segs) + 1
            segs.append([frm, to, left, right, 0])
        for k in range(nx):
            i = nx - 1 - k
            segs[k][4] = main_id[i - 1] if i > 0 else -3        # ZERO_DPTH_GRAD outlet
        # tributaries along x = i grid lines joining node (i, jc)
        if trib_every > 0:
            trib_len = max(1, min(jc - 1, ny - jc - 1))
            for i in range(trib_every, nx, trib_every):
                down_main = main_id[i - 1]
                # north side: rows j = jc+trib_len-1 ... jc, flowing to -y
                first = len(segs) + 1
                for t, j in enumerate(range(jc + trib_len - 1, jc - 1, -1)):
                    frm, to = node(i, j + 1), node(i, j)
                    left = 2 * (j * nx + i) + 1 + 1         # cell(i,j) k=1 (left edge), east side
                    right = 2 * (j * nx + i - 1) + 0 + 1    # cell(i-1,j) k=0 (right edge)
                    down = first + t + 1 if j > jc else down_main
                    segs.append([frm, to, left, right, down])
                # south side: rows j = jc-trib_len ... jc-1, flowing to +y
                first = len(segs) + 1
                for t, j in enumerate(range(jc - trib_len, jc)):
                    frm, to = node(i, j), node(i, j + 1)
                    left = 2 * (j * nx + i - 1) + 0 + 1
                    right = 2 * (j * nx + i) + 1 + 1
                    down = first + t + 1 if j < jc - 1 else down_main
                    segs.append([frm, to, left, right, down])
    nr = len(segs)
    rf = np.zeros((R_NCOL, nr))
    ri = np.zeros((RI_NCOL, nr), np.int32)
    rx = np.zeros(nr); ry = np.zeros(nr)
    if nr:
        S = np.array(segs, np.int64)
        frm, to, left, right, down = S.T
        l0, r0 = left - 1, right - 1
        # rewrite bank-element neighbours to -(river index)  (init_river.c:26-37)
        ridx = np.arange(1, nr + 1)
        for side_a, side_b in ((l0, right), (r0, left)):
            for j in range(3):
                hit = nabr[side_a, j] == side_b
                nabr[side_a[hit], j] = -ridx[hit]
        ri[RI_LEFTELE], ri[RI_RIGHTELE], ri[RI_DOWN] = left, right, down
        ri[RI_BCTYPE] = 0
        ri[RI_INTRPL_ORD] = riv_order
        rx = 0.5 * (xf[frm] + xf[to]); ry = 0.5 * (yf[frm] + yf[to])
        zmaxf = zmax_n.ravel()
        rf[R_ZMAX] = 0.5 * (zmaxf[frm] + zmaxf[to])
        rf[R_ZMIN] = rf[R_ZMAX] - (0.5 * (ef[E_ZMAX, l0] + ef[E_ZMAX, r0])
                                   - 0.5 * (ef[E_ZMIN, l0] + ef[E_ZMIN, r0]))
        rf[R_NODE_ZMAX] = zmaxf[to]
        rf[R_DIST_LEFT] = np.sqrt((rx - xc[l0]) * (rx - xc[l0]) + (ry - yc[l0]) * (ry - yc[l0]))
        rf[R_DIST_RIGHT] = np.sqrt((rx - xc[r0]) * (rx - xc[r0]) + (ry - yc[r0]) * (ry - yc[r0]))
        depth, coeff = RIV_SHAPE["depth"], RIV_SHAPE["coeff"]
        rf[R_SHP_DEPTH] = depth
        rf[R_SHP_COEFF] = coeff
        rf[R_SHP_LENGTH] = np.sqrt((xf[frm] - xf[to]) ** 2 + (yf[frm] - yf[to]) ** 2)
        rf[R_SHP_WIDTH] = _river_eq_wid(riv_order, depth, coeff)
        rf[R_ZBED] = rf[R_ZMAX] - depth
        rf[R_ROUGH] = RIV_MATL["rough"]; rf[R_CWR] = RIV_MATL["cwr"]
        rf[R_KSATH] = RIV_MATL["ksath"]; rf[R_KSATV] = RIV_MATL["ksatv"]
        rf[R_BEDTHICK] = RIV_MATL["bedthick"]
        rf[R_AREA] = rf[R_SHP_LENGTH] * _river_eq_wid(riv_order, depth, coeff)

    ei[EI_NABR0], ei[EI_NABR1], ei[EI_NABR2] = nabr[:, 0], nabr[:, 1], nabr[:, 2]

    # ---- neighbour centroids and distances (InitSurfL) ----------------------
    for j in range(3):
        a, b = (j + 1) % 3, (j + 2) % 3
        nb = nabr[:, j]
        distx = xc - 0.5 * (X[:, a] + X[:, b])
        disty = yc - 0.5 * (Y[:, a] + Y[:, b])
        bx = xc - 2.0 * distx
        by = yc - 2.0 * disty
        circ = ef[E_EDGE0] * ef[E_EDGE1] * ef[E_EDGE2] / (4.0 * ef[E_AREA])
        bd = np.sqrt(np.maximum(circ ** 2 - (ef[E_EDGE0 + j] / 2.0) ** 2, 0.0))
        en = np.clip(nb - 1, 0, ne - 1)
        rn = np.clip(-nb - 1, 0, max(nr - 1, 0))
        nbx = np.where(nb > 0, xc[en], rx[rn] if nr else 0.0)
        nby = np.where(nb > 0, yc[en], ry[rn] if nr else 0.0)
        dd = (xc - nbx) * (xc - nbx)
        dd = dd + (yc - nby) * (yc - nby)
        dd = np.sqrt(dd)
        ef[E_NABRX0 + j] = np.where(nb == 0, bx, nbx)
        ef[E_NABRY0 + j] = np.where(nb == 0, by, nby)
        ef[E_NABRDIST0 + j] = np.where(nb == 0, bd, dd)

    # ---- soil, land cover, geology ------------------------------------------
    blk = 8
    soil_cls = _hash2(ci // blk, cj // blk, len(SOIL_CLASSES))
    lc_cls = _hash2(ci // (2 * blk) + 7, cj // (2 * blk) + 3, len(LC_CLASSES))
    soil_e = np.repeat(soil_cls, 2); lc_e = np.repeat(lc_cls, 2)
    sp = SOIL_CLASSES[soil_e]
    ef[E_DEPTH] = ef[E_ZMAX] - ef[E_ZMIN]
    (ef[E_KSATH], ef[E_KSATV], ef[E_KINFV], ef[E_DINF], ef[E_ALPHA], ef[E_BETA],
     ef[E_POROSITY], ef[E_KMACH], ef[E_KMACV], ef[E_AREAFV], ef[E_AREAFH]) = sp.T
    ef[E_DMAC] = np.minimum(DMAC_TBL, ef[E_DEPTH])
    ef[E_ROUGH], ef[E_RZD] = LC_CLASSES[lc_e].T
    if fbr:
        ef[E_GDEPTH] = ef[E_ZMIN] - ef[E_ZBED]
        ef[E_GKSATH] = GEOL["ksath"]; ef[E_GKSATV] = GEOL["ksatv"]
        ef[E_GALPHA] = GEOL["alpha"]; ef[E_GBETA] = GEOL["beta"]
        ef[E_GPOROSITY] = GEOL["porosity"]
    else:
        ef[E_ZBED] = 0.0
    if nr:
        rf[R_POROSITY] = 0.5 * (ef[E_POROSITY, l0] + ef[E_POROSITY, r0])

    forc_bc = np.zeros((3, ne))
    if dirichlet_edges:
        # Dirichlet head on the x = Lx boundary edges (k=0 elements' edge 0 at i = nx-1)
        sel = np.where((nabr[:, 0] == 0) & (np.arange(ne) % 2 == 0))[0]
        ei[EI_BC0, sel] = 1
        forc_bc[0, sel] = ef[E_ZMIN, sel] + 0.6 * ef[E_DEPTH, sel]
        if fbr:
            ei[EI_FBRBC0, sel] = 1

    # ---- relaxed initial condition (RelaxIc) --------------------------------
    y0 = [np.zeros(ne), np.full(ne, 0.1), ef[E_DEPTH] - 0.1,
          np.zeros(nr), (rf[R_ZBED] - rf[R_ZMIN] - 0.1) if nr else np.zeros(0)]
    if fbr:
        fg = np.minimum(5.0, ef[E_GDEPTH])
        y0 += [0.5 * (ef[E_GDEPTH] - fg), fg]
    y0 = np.concatenate(y0)

    out = dict(nelem=ne, nriver=nr, fbr=int(fbr), surf_mode=surf_mode, riv_mode=riv_mode,
               stepsize=float(stepsize), elem_f64=ef, elem_i32=ei, riv_f64=rf, riv_i32=ri,
               y0=y0, xc=xc, yc=yc, nx=nx, ny=ny, bc_head=forc_bc)
    if keep_mesh:
        S = np.array(segs, np.int64).reshape(-1, 5)
        out["mesh"] = dict(node=nodes + 1, nabr=nabr_mesh, x=xf, y=yf, zmin=zmin_n.ravel(), zmax=zmax_n.ravel(),
                           zbed=zbed_n.ravel(), soil_type=soil_e + 1, lc_type=lc_e + 1,
                           riv_from=S[:, 0] + 1, riv_to=S[:, 1] + 1, riv_order=riv_order)
    return out


def make_named(name: str, **kw) -> dict:
    nx, ny = SIZES[name]
    return make_watershed(nx, ny, **kw)


def storm_rain(t: float) -> float:
    """rain rate (m/s) of the synthetic storm at model time t (s): a 6 h half-sine pulse, 12 mm/h at its peak, every 24 h"""
    th = (t / 3600.0) % 24.0
    return 12.0e-3 / 3600.0 * np.sin(np.pi * (th - 1.0) / 6.0) if 1.0 <= th < 7.0 else 0.0


def storm_modulation(tables: dict) -> np.ndarray:
    """spatial factor of the rain rate per element"""
    return 0.75 + 0.25 * np.sin(2 * np.pi * tables["xc"] / 5000.0)


def storm_forcing(tables: dict, t: float, ws0_surf=None) -> np.ndarray:
    """Synthetic forcing table [F_NCOL, nelem] at model time t (s):
    a 6 h rain pulse (12 mm/h peak, spatially modulated) every 24 h and
    constant evaporation / transpiration demands.  Stands in for ApplyForc +
    IntcpSnowEt (src/pihm.c:27-48), which stay host code in the reference."""
    ne = tables["nelem"]
    f = np.zeros((F_NCOL, ne))
    f[F_PCPDRP] = storm_rain(t) * storm_modulation(tables)
    f[F_EDIR] = 2.0e-8
    f[F_ETT] = 3.0e-8
    if ws0_surf is not None:
        f[F_WS0SURF] = ws0_surf
    f[F_BC0:F_BC0 + 3] = tables["bc_head"]
    if tables["fbr"]:
        f[F_FBRBC0:F_FBRBC0 + 3] = tables["bc_head"] - 8.0
    return f


def state_slices(ne: int, nr: int, fbr: bool) -> dict:
    """Block layout of y (src/include/pihm_func.h:7-15)."""
    s = dict(surf=slice(0, ne), unsat=slice(ne, 2 * ne), gw=slice(2 * ne, 3 * ne),
             stage=slice(3 * ne, 3 * ne + nr), rivgw=slice(3 * ne + nr, 3 * ne + 2 * nr))
    if fbr:
        s["fbr_unsat"] = slice(3 * ne + 2 * nr, 4 * ne + 2 * nr)
        s["fbr_gw"] = slice(4 * ne + 2 * nr, 5 * ne + 2 * nr)
    return s


def wet_state(tables: dict, seed: int = 7, ponded_frac: float = 0.3) -> np.ndarray:
    """A branch-rich test state: part of the surface ponded above/below the
    depression storage, water tables spanning depth-dinf and depth-dmac,
    rivers from dry to over-bank.  Used by parity tests and the RHS bench."""
    rng = np.random.default_rng(seed)
    ne, nr, fbr = tables["nelem"], tables["nriver"], bool(tables["fbr"])
    ef = tables["elem_f64"]
    sl = state_slices(ne, nr, fbr)
    y = tables["y0"].copy()
    u = rng.uniform(size=ne)
    surf = np.where(u < ponded_frac, rng.uniform(0, 0.02, ne),
                    np.where(u < 0.6, rng.uniform(0, 1.0e-4, ne), 0.0))
    surf[rng.uniform(size=ne) < 0.02] = -1.0e-6          # clamped negatives
    y[sl["surf"]] = surf
    depth = ef[E_DEPTH]
    gw = depth * rng.uniform(0.2, 1.02, ne)
    y[sl["gw"]] = gw
    y[sl["unsat"]] = np.maximum(depth - gw, 0.0) * rng.uniform(0.0, 1.1, ne)
    if nr:
        rfz = tables["riv_f64"]
        y[sl["stage"]] = rfz[R_SHP_DEPTH] * rng.uniform(0.0, 1.6, nr) * (rng.uniform(size=nr) > 0.1)
        y[sl["rivgw"]] = (rfz[R_ZBED] - rfz[R_ZMIN]) * rng.uniform(0.5, 1.1, nr)
    if fbr:
        gd = ef[E_GDEPTH]
        fg = gd * rng.uniform(0.1, 1.01, ne)
        y[sl["fbr_gw"]] = fg
        y[sl["fbr_unsat"]] = np.maximum(gd - fg, 0.0) * rng.uniform(0.0, 1.05, ne)
    return y
