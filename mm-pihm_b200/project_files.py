"""A synthetic watershed as an MM-PIHM project: the text input files the reference's own readers parse
(SURVEY.md appendix C), so that `ReadAlloc()` + `Initialize()` of the unchanged program -- and with them
`InitTopo`, `InitSurfL`, `InitRiver`, `InitSoil`, `InitLc`, `InitGeol`, `RelaxIc` -- build the very tables
`watershed.make_watershed()` states directly.  tests/test_project_files.py reads the files back through
oracle/_ref and compares table for table; the files also feed the unchanged `pihm` / `pihm-fbr` drivers
linked with glue/pihm_b200_glue.c (`make -C oracle drivers`) with meshes of any size.

Formats (parser in the reference -> writer here):
  .mesh     src/read_mesh.c:3-78          write_mesh       NUMELE / idx n1 n2 n3 nabr1 nabr2 nabr3 / NUMNODE / idx x y zmin zmax
  .att      src/read_att.c:3-47           write_att        idx soil geol lc meteo lai ss bc0 bc1 bc2
  .riv      src/read_river.c:3-183        write_riv        NUMRIV / segments / SHAPE / MATERIAL / BC 0 / RES 0
  .soil     src/read_soil.c:3-151         write_soil       NUMSOIL / 16 columns / DINF KMACV_RO KMACH_RO
  .geol     src/fbr/read_geol.c:3-48      write_geol       NUMGEOL / idx ksatv ksath maxsmc minsmc alpha beta
  .bedrock  src/fbr/read_bedrock.c:3-82   write_bedrock    fbr bc types / idx zbed per node / 5 print controls
  vegprmt   src/read_lc.c:3-78            write_vegprmt    NUMLC / 15 columns / TOPT_DATA CFACTR_DATA RSMAX_DATA BARE NATURAL
  .meteo    src/read_forc.c:3-79          write_meteo      METEO_TS i WIND_LVL z / 2 header lines / time + 7 values
  .bc       src/read_bc.c:3-102           write_bc         BC_TS k / 2 header lines / time + head  (one series per Dirichlet edge)
  .lai      src/read_lai.c:3-110          write_lai        LAI_TS 1 / 2 header lines / time + LAI  (option lai_series)
  .para     src/read_para.c:3-200         write_para       fixed keyword ORDER
  .calib    src/read_calib.c:3-173        write_calib      all multipliers 1 (the class tables already hold calibrated values)

Numbers are written with 17 significant digits: `sscanf("%lf")` returns the double that was written.
The class tables of watershed.py hold *calibrated* values (what InitSoil derives for input/example), so the
.calib written here is neutral and the .soil columns are chosen so that InitSoil (src/init_soil.c:25-66)
returns them: MINSMC 0 and MAXSMC = porosity, KMACV_RO = kmacv / kinfv, KMACH_RO = kmach / ksath (the two
products re-round: equal to the table within 1 ulp, everything else bit for bit).

Boundary conditions: MM-PIHM's bc type of an edge IS the (1-based) index of a time series of the .bc file, its sign
the kind (> 0 Dirichlet head, < 0 Neumann flux; src/forcing.c:55-86).  The Dirichlet edges of
`make_watershed(dirichlet_edges=True)` carry one head per edge, so every such edge gets a constant series of its own
and the .att / .bedrock columns hold the series indices where the generator's tables hold the flag 1 (the RHS reads
only the sign).  The bedrock layer's heads (storm_forcing: soil head - 8 m) follow as further series.
Without `lai_series` every element has lai type 0 = monthly table by land cover (src/forcing.c:248-257) and ReadLai
opens nothing (src/read_lai.c:13-24).  Not written: .ic (INIT_MODE 0 = RelaxIc), river boundary series, the module
files of Noah / BGC / Cycles / RT."""
from __future__ import annotations

import os

import numpy as np

from . import watershed as W

START = np.datetime64("2009-01-01T00:00")


def _g(v) -> str:
    return "%.17g" % float(v)


def write_mesh(path: str, mesh: dict) -> None:
    ne, nn = len(mesh["node"]), len(mesh["x"])
    with open(path, "w") as f:
        f.write(f"NUMELE\t{ne}\nINDEX\tNODE1\tNODE2\tNODE3\tNABR1\tNABR2\tNABR3\n")
        idx = np.arange(1, ne + 1)
        np.savetxt(f, np.column_stack([idx, mesh["node"], mesh["nabr"]]), fmt="%d", delimiter="\t")
        f.write(f"NUMNODE\t{nn}\nINDEX\tX\tY\tZMIN\tZMAX\n")
        for i in range(nn):
            f.write(f"{i + 1}\t{_g(mesh['x'][i])}\t{_g(mesh['y'][i])}\t{_g(mesh['zmin'][i])}\t{_g(mesh['zmax'][i])}\n")


def bc_series(tables: dict):
    """-> (bc index columns [3, ne] for .att, fbr bc index columns [3, ne] for .bedrock, heads of the series).
    One constant series per Dirichlet edge, soil edges first (element-major), then the bedrock edges."""
    ei, ne = tables["elem_i32"], tables["nelem"]
    if np.any(ei[W.EI_BC0:W.EI_FBRBC2 + 1] < 0) or np.any(tables["riv_i32"][W.RI_BCTYPE]):
        raise ValueError("Neumann edges and river boundary conditions are not written")
    heads = []
    soil = np.zeros((3, ne), np.int64)
    fbr = np.zeros((3, ne), np.int64)
    for dst, first, off in ((soil, W.EI_BC0, 0.0), (fbr, W.EI_FBRBC0, -8.0)):
        for i, j in zip(*np.nonzero(ei[first:first + 3].T > 0)):       # element-major
            heads.append(tables["bc_head"][j, i] + off)
            dst[j, i] = len(heads)
    return soil, fbr, heads


def write_att(path: str, tables: dict, bc_idx=None, lai_type: int = 0) -> None:
    mesh, ei = tables["mesh"], tables["elem_i32"]
    ne = tables["nelem"]
    bc = ei[W.EI_BC0:W.EI_BC2 + 1] if bc_idx is None else bc_idx
    cols = [np.arange(1, ne + 1), mesh["soil_type"], np.ones(ne, int), mesh["lc_type"], np.ones(ne, int),
            np.full(ne, lai_type, int), np.zeros(ne, int), bc[0], bc[1], bc[2]]
    with open(path, "w") as f:
        f.write("INDEX\tSOIL\tGEOL\tLC\tMETEO\tLAI\tSS\tBC0\tBC1\tBC2\n")
        np.savetxt(f, np.column_stack(cols), fmt="%d", delimiter="\t")


def write_riv(path: str, tables: dict) -> None:
    mesh, ri = tables["mesh"], tables["riv_i32"]
    nr = tables["nriver"]
    one = np.ones(nr, int)
    cols = [np.arange(1, nr + 1), mesh["riv_from"], mesh["riv_to"], ri[W.RI_DOWN], ri[W.RI_LEFTELE],
            ri[W.RI_RIGHTELE], one, one, ri[W.RI_BCTYPE], 0 * one]
    s, m = W.RIV_SHAPE, W.RIV_MATL
    with open(path, "w") as f:
        f.write(f"NUMRIV\t{nr}\nINDEX\tFROM\tTO\tDOWN\tLEFT\tRIGHT\tSHAPE\tMATL\tBC\tRES\n")
        if nr:
            np.savetxt(f, np.column_stack(cols), fmt="%d", delimiter="\t")
        f.write("SHAPE\t1\nINDEX\tDPTH\tOINT\tCWID\n")
        f.write(f"1\t{_g(s['depth'])}\t{int(mesh['riv_order'])}\t{_g(s['coeff'])}\n")
        f.write("MATERIAL\t1\nINDEX\tROUGH\tCWR\tKH\tKV\tBEDTHCK\n")
        f.write(f"1\t{_g(m['rough'])}\t{_g(m['cwr'])}\t{_g(m['ksath'])}\t{_g(m['ksatv'])}\t{_g(m['bedthick'])}\n")
        f.write("BC\t0\nRES\t0\n")


def soil_file_rows() -> tuple[list[list[float]], float, float]:
    """The .soil columns that InitSoil turns into watershed.SOIL_CLASSES with a neutral .calib.
    KMACV_RO / KMACH_RO are file-wide scalars: the class tables keep kmacv / kinfv and kmach / ksath
    constant across classes (100 and 700, the calibrated ratios of input/example)."""
    rows = []
    rv = rh = None
    for ksath, ksatv, kinfv, dinf, alpha, beta, por, kmach, kmacv, areafv, areafh in W.SOIL_CLASSES:
        rv_c, rh_c = round(kmacv / kinfv), round(kmach / ksath)
        assert rv in (None, rv_c) and rh in (None, rh_c), "KMACV_RO / KMACH_RO are one value per file"
        rv, rh = rv_c, rh_c
        #            silt  clay  om   bd   kinf   ksatv  ksath  maxsmc minsmc alpha beta machf(areafh) macvf(areafv) dmac qtz
        rows.append([35.0, 15.0, 3.5, 1.4, kinfv, ksatv, ksath, por, 0.0, alpha, beta, areafh, areafv, W.DMAC_TBL, 0.25])
    return rows, float(rv), float(rh)


def write_soil(path: str) -> None:
    rows, kmacv_ro, kmach_ro = soil_file_rows()
    with open(path, "w") as f:
        f.write(f"NUMSOIL\t{len(rows)}\n")
        f.write("INDEX\tSILT\tCLAY\tOM\tBD\tKINF\tKSATV\tKSATH\tMAXSMC\tMINSMC\tALPHA\tBETA\tMACHF\tMACVF\tDMAC\tQTZ\n")
        for i, r in enumerate(rows):
            f.write("\t".join([str(i + 1)] + [_g(v) for v in r]) + "\n")
        f.write(f"DINF\t{_g(W.SOIL_CLASSES[0][3])}\nKMACV_RO\t{_g(kmacv_ro)}\nKMACH_RO\t{_g(kmach_ro)}\n")


def write_geol(path: str) -> None:
    g = W.GEOL
    with open(path, "w") as f:
        f.write("NUMGEOL\t1\nINDEX\tKSATV\tKSATH\tMAXSMC\tMINSMC\tALPHA\tBETA\n")
        f.write(f"1\t{_g(g['ksatv'])}\t{_g(g['ksath'])}\t{_g(g['porosity'])}\t0\t{_g(g['alpha'])}\t{_g(g['beta'])}\n")


def write_bedrock(path: str, tables: dict, fbr_bc_idx=None) -> None:
    mesh, ei = tables["mesh"], tables["elem_i32"]
    ne, nn = tables["nelem"], len(mesh["x"])
    bc = ei[W.EI_FBRBC0:W.EI_FBRBC2 + 1] if fbr_bc_idx is None else fbr_bc_idx
    with open(path, "w") as f:
        f.write("INDEX\tBC0\tBC1\tBC2\n")
        np.savetxt(f, np.column_stack([np.arange(1, ne + 1), bc[0], bc[1], bc[2]]), fmt="%d", delimiter="\t")
        f.write("INDEX\tZBED\n")
        for i in range(nn):
            f.write(f"{i + 1}\t{_g(mesh['zbed'][i])}\n")
        for key in ("FBRUNSAT", "FBRGW", "FBRINFIL", "FBRRECHG", "FBRFLOW"):
            f.write(f"{key}\tHOURLY\n")


def write_vegprmt(path: str) -> None:
    """input/vegprmt.tbl with the classes of watershed.LC_CLASSES (rough, rzd); the ET columns are those of a
    mixed forest -- the RHS path reads only ROUGH and DROOT (src/init_lc.c:21,38)."""
    with open(path, "w") as f:
        f.write(f"NUMLC\t{len(W.LC_CLASSES)}\n")
        f.write("INDEX\tSHDFAC\tDROOT\tRS\tRGL\tHS\tSNUP\tLAIMIN\tLAIMAX\tEMISMIN\tEMISMAX\tALBMIN\tALBMAX\tZ0MIN\tZ0MAX\tROUGH\n")
        for i, (rough, rzd) in enumerate(W.LC_CLASSES):
            f.write(f"{i + 1}\t0.8\t{_g(rzd)}\t125\t30\t51.93\t0.08\t2.8\t5.5\t0.93\t0.97\t0.17\t0.25\t0.2\t0.5\t{_g(rough)}\n")
        f.write("TOPT_DATA\t298.0\nCFACTR_DATA\t0.5\nRSMAX_DATA\t5000.0\n")
        f.write(f"BARE\t{len(W.LC_CLASSES) + 1}\nNATURAL\t{len(W.LC_CLASSES) + 2}\n")


def write_meteo(path: str, hours: int) -> None:
    """one station, hourly records: the rain pulse of watershed.storm_forcing as a precipitation series"""
    with open(path, "w") as f:
        f.write("METEO_TS\t1\tWIND_LVL\t10.0\n")
        f.write("TIME\tPRCP\tSFCTMP\tRH\tSFCSPD\tSOLAR\tLONGWV\tPRES\n")
        f.write("TS\tkg/m2/s\tK\t%\tm/s\tW/m2\tW/m2\tPa\n")
        for h in range(hours + 1):
            th = h % 24
            rain = 12.0 / 3600.0 * np.sin(np.pi * (th - 1.0) / 6.0) if 1 <= th < 7 else 0.0    # kg/m2/s
            sun = 600.0 * max(0.0, np.sin(np.pi * (th - 6.0) / 12.0))
            t = str(START + np.timedelta64(h, "h")).replace("T", " ")
            f.write(f"{t}\t{rain:.8f}\t285.15\t70.0\t2.5\t{sun:.2f}\t300.0\t97000.0\n")


def _span(hours: int):
    """first and last record of a constant series: the run's span with a day to spare on both sides"""
    fmt = lambda t: str(t).replace("T", " ")     # noqa: E731
    return fmt(START - np.timedelta64(24, "h")), fmt(START + np.timedelta64(hours + 24, "h"))


def write_bc(path: str, heads, hours: int) -> None:
    t0, t1 = _span(hours)
    with open(path, "w") as f:
        for k, h in enumerate(heads):
            f.write(f"BC_TS\t{k + 1}\nTIME\tHEAD\nTS\tm\n{t0}\t{_g(h)}\n{t1}\t{_g(h)}\n")


def write_lai(path: str, hours: int, lai: float = 3.0) -> None:
    t0, t1 = _span(hours)
    with open(path, "w") as f:
        f.write(f"LAI_TS\t1\nTIME\tLAI\nTS\tm2/m2\n{t0}\t{_g(lai)}\n{t1}\t{_g(lai)}\n")


PRINT_KEYS = ("SURF", "UNSAT", "GW", "RIVSTG", "RIVGW", "SNOW", "CMC", "INFIL", "RECHARGE", "EC", "ETT", "EDIR",
              "RIVFLX0", "RIVFLX1", "RIVFLX2", "RIVFLX3", "RIVFLX4", "RIVFLX5", "RIVFLX6", "RIVFLX7", "RIVFLX8",
              "RIVFLX9", "RIVFLX10", "SUBFLX", "SURFFLX")


def write_para(path: str, tables: dict, hours: int, reltol=1e-3, abstol=1e-4, initstep=5e-5,
               print_interval="HOURLY") -> None:
    end = str(START + np.timedelta64(hours, "h")).replace("T", " ")
    rows = [("SIMULATION_MODE", 0), ("INIT_MODE", 0), ("ASCII_OUTPUT", 0), ("WATBAL_OUTPUT", 0), ("WRITE_IC", 0),
            ("UNSAT_MODE", 2), ("SURF_MODE", int(tables["surf_mode"])), ("RIV_MODE", int(tables["riv_mode"])),
            ("START", str(START).replace("T", " ")), ("END", end), ("MAX_SPINUP_YEAR", 1),
            ("MODEL_STEPSIZE", int(tables["stepsize"])), ("LSM_STEP", 15 * int(tables["stepsize"])),
            ("ABSTOL", _g(abstol)), ("RELTOL", _g(reltol)), ("INIT_SOLVER_STEP", _g(initstep)),
            ("NUM_NONCOV_FAIL", 0.0), ("MAX_NONLIN_ITER", 3.0), ("MIN_NONLIN_ITER", 1.0), ("DECR_FACTOR", 1.2),
            ("INCR_FACTOR", 1.2), ("MIN_MAXSTEP", 1.0)]
    rows += [(k, print_interval) for k in PRINT_KEYS] + [("IC", "MONTHLY")]
    with open(path, "w") as f:
        for k, v in rows:
            f.write(f"{k}\t{v}\n")


CALIB_HYDRO = ("KSATH", "KSATV", "KINF", "KMACSATH", "KMACSATV", "DINF", "DROOT", "DMAC", "POROSITY", "ALPHA",
               "BETA", "MACVF", "MACHF", "VEGFRAC", "ALBEDO", "ROUGH", "EC", "ETT", "EDIR", "ROUGH_RIV", "KRIVH",
               "KRIVV", "BEDTHCK", "RIV_DPTH", "RIV_WDTH")
CALIB_LSM = ("DRIP", "CMCMAX", "RS", "CZIL", "FXEXP", "CFACTR", "RGL", "HS", "REFSMC", "WLTSMC")


def write_calib(path: str) -> None:
    with open(path, "w") as f:
        for k in CALIB_HYDRO:
            f.write(f"{k}\t1.0\n")
        f.write("\nLSM_CALIBRATION\n")
        for k in CALIB_LSM:
            f.write(f"{k}\t1.0\n")
        f.write("\nBGC_CALIBRATION\nMORTALITY\t1.0\nSLA\t1.0\n")
        f.write("\nRT_CALIBRATION\nRATE\t0.0\nSSA\t1.0\nGWINFLUX\t1.0\nPRCPCONC\t1.0\nINITCONC\t1.0\nXSORPTION\t0.0\n")
        f.write("\nSCENARIO\nPRCP\t1.0\nSFCTMP\t0.0\n")


def write_project(tables: dict, rundir: str, name: str = "synth", hours: int = 24, lai_series: bool = False,
                  **para) -> str:
    """Write input/<name>/<name>.* and input/vegprmt.tbl under rundir (created if missing) from a watershed
    made with keep_mesh=True.  Returns the project directory."""
    if "mesh" not in tables:
        raise ValueError("make_watershed(..., keep_mesh=True) is needed: the files hold nodes, not centroids")
    soil_bc, fbr_bc, heads = bc_series(tables)
    d = os.path.join(rundir, "input", name)
    os.makedirs(d, exist_ok=True)
    p = lambda ext: os.path.join(d, f"{name}.{ext}")     # noqa: E731
    write_mesh(p("mesh"), tables["mesh"])
    write_att(p("att"), tables, soil_bc, lai_type=1 if lai_series else 0)
    if heads:
        write_bc(p("bc"), heads, hours)
    if lai_series:
        write_lai(p("lai"), hours)
    write_riv(p("riv"), tables)
    write_soil(p("soil"))
    write_meteo(p("meteo"), hours)
    write_para(p("para"), tables, hours, **para)
    write_calib(p("calib"))
    write_vegprmt(os.path.join(rundir, "input", "vegprmt.tbl"))
    if tables["fbr"]:
        write_geol(p("geol"))
        write_bedrock(p("bedrock"), tables, fbr_bc)
    return d
