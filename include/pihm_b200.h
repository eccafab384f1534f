/*
 * pihm_b200.h -- C ABI of the B200-native MM-PIHM hot path (libpihm_b200.so).
 *
 * Everything here is plain C: pointers, sizes and opaque handles.  No torch,
 * no C++ types.  A PIHM driver written in C links this library and keeps its
 * own elem_struct/river_struct arrays; glue/pihm_b200_glue.c is the translation
 * unit that replaces src/ode.c in the pihm / pihm-fbr drivers: it packs those
 * AoS arrays into the column tables below and defines ODE / SetCVodeParam /
 * SolveCVode / AdjCVodeMaxStep / NumStateVar on top of this ABI.
 *
 * Each entry point cites the reference interface it replaces
 * (paths relative to the MM-PIHM tree).
 */
#ifndef PIHM_B200_H
#define PIHM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIHM_B200_ABI_VERSION 1

/* ------------------------------------------------------------------------
 * Column tables.  Every table is a dense row-major array [NCOL][n]; column c
 * of entity i lives at table[c * n + i].  Values are exactly the fields the
 * reference RHS reads after Initialize() (src/initialize.c:5-205).
 * ---------------------------------------------------------------------- */

/* element, static double columns (src/include/elem_struct.h:28-111,136,240) */
enum pihm_b200_elem_col {
    PB_E_AREA = 0, PB_E_ZMIN, PB_E_ZMAX, PB_E_ZBED,
    PB_E_EDGE0, PB_E_EDGE1, PB_E_EDGE2,
    PB_E_NABRDIST0, PB_E_NABRDIST1, PB_E_NABRDIST2,
    PB_E_NABRX0, PB_E_NABRX1, PB_E_NABRX2,
    PB_E_NABRY0, PB_E_NABRY1, PB_E_NABRY2,
    PB_E_DEPTH, PB_E_KSATH, PB_E_KSATV, PB_E_KINFV, PB_E_DINF,
    PB_E_ALPHA, PB_E_BETA, PB_E_POROSITY, PB_E_DMAC, PB_E_KMACH,
    PB_E_KMACV, PB_E_AREAFV, PB_E_AREAFH,
    PB_E_ROUGH, PB_E_RZD,
    /* fractured bedrock (geol_struct), only read when fbr != 0 */
    PB_E_GDEPTH, PB_E_GKSATH, PB_E_GKSATV, PB_E_GALPHA, PB_E_GBETA,
    PB_E_GPOROSITY,
    PB_E_NCOL
};

/* element, static int32 columns (elem_struct.h:1560, attrib_struct :5-25).
 * NABRj: >0 neighbour element (1-based), <0 -(river index, 1-based), 0 boundary */
enum pihm_b200_elem_icol {
    PB_EI_NABR0 = 0, PB_EI_NABR1, PB_EI_NABR2,
    PB_EI_BC0, PB_EI_BC1, PB_EI_BC2,
    PB_EI_FBRBC0, PB_EI_FBRBC1, PB_EI_FBRBC2,
    PB_EI_NCOL
};

/* element, per-step forcing columns (wf.pcpdrp/edir/ett written by
 * IntcpSnowEt, src/is_sm_et.c:126-223; ws0.surf written by Summary,
 * src/update.c:47; bc/fbr_bc written by ApplyBc, src/forcing.c:55-86) */
enum pihm_b200_elem_forc_col {
    PB_F_PCPDRP = 0, PB_F_EDIR, PB_F_ETT, PB_F_WS0SURF,
    PB_F_BC0, PB_F_BC1, PB_F_BC2,
    PB_F_FBRBC0, PB_F_FBRBC1, PB_F_FBRBC2,
    PB_F_NCOL
};

/* river, static double columns (src/include/river_struct.h:11-60) */
enum pihm_b200_riv_col {
    PB_R_AREA = 0, PB_R_ZMIN, PB_R_ZMAX, PB_R_ZBED, PB_R_NODE_ZMAX,
    PB_R_DIST_LEFT, PB_R_DIST_RIGHT,
    PB_R_SHP_DEPTH, PB_R_SHP_COEFF, PB_R_SHP_LENGTH, PB_R_SHP_WIDTH,
    PB_R_ROUGH, PB_R_CWR, PB_R_KSATH, PB_R_KSATV, PB_R_BEDTHICK,
    PB_R_POROSITY,
    PB_R_NCOL
};

/* river, static int32 columns (river_struct.h:137-146, :41) */
enum pihm_b200_riv_icol {
    PB_RI_LEFTELE = 0, PB_RI_RIGHTELE, PB_RI_DOWN, PB_RI_BCTYPE,
    PB_RI_INTRPL_ORD,
    PB_RI_NCOL
};

/* per-element flux diagnostics returned by pihm_b200_get_elem_fluxes
 * (wflux_struct, elem_struct.h:407-477) */
enum pihm_b200_elem_flux_col {
    PB_X_OVL0 = 0, PB_X_OVL1, PB_X_OVL2,
    PB_X_SUB0, PB_X_SUB1, PB_X_SUB2,
    PB_X_INFIL, PB_X_RECHG,
    PB_X_EDIR_SURF, PB_X_EDIR_UNSAT, PB_X_EDIR_GW, PB_X_ETT_UNSAT, PB_X_ETT_GW,
    PB_X_FBR_INFIL, PB_X_FBR_RECHG, PB_X_FBRFLOW0, PB_X_FBRFLOW1, PB_X_FBRFLOW2,
    PB_X_NCOL
};

/* interception / snow / ET: static double columns read by
 * IntcpSnowEt (src/is_sm_et.c:4-225) besides those of pihm_b200_elem_col:
 * lc_struct (elem_struct.h:113-137), epc_struct, soil smc*, ps.zlvl_wind */
enum pihm_b200_et_col {
    PB_ET_ALBEDOMIN = 0, PB_ET_ALBEDOMAX, PB_ET_CMCFACTR, PB_ET_SHDFAC,
    PB_ET_CFACTR, PB_ET_RGL, PB_ET_RSMIN, PB_ET_RSMAX, PB_ET_TOPT,
    PB_ET_SMCMIN, PB_ET_SMCWLT, PB_ET_SMCREF, PB_ET_ZLVL_WIND,
    PB_ET_NCOL
};
/* attrib.meteo_type / lai_type / lc_type (elem_struct.h:5-25), 1-based;
 * lai_type 0 = monthly table by land cover (forcing.c:249-257) */
enum pihm_b200_et_icol {
    PB_ETI_METEO_TYPE = 0, PB_ETI_LAI_TYPE, PB_ETI_LC_TYPE,
    PB_ETI_NCOL
};
/* what pihm_b200_et_get returns per element: the three forcing columns of the
 * RHS, the two canopy fluxes, the two storages IntcpSnowEt integrates */
enum pihm_b200_et_out_col {
    PB_EO_PCPDRP = 0, PB_EO_EDIR, PB_EO_ETT, PB_EO_EC, PB_EO_DRIP,
    PB_EO_SNEQV, PB_EO_CMC,
    PB_EO_NCOL
};
#define PIHM_B200_NUM_METEO_VAR 7  /* src/include/pihm_const.h:48-55 */

#define PIHM_B200_NUM_RIVFLX 11   /* src/include/pihm_const.h:71,130-140 */

typedef struct pihm_b200_mesh {
    int32_t         nelem;
    int32_t         nriver;
    int32_t         fbr;        /* 0: pihm (3 states/elem); 1: pihm-fbr (5) */
    int32_t         surf_mode;  /* 1 kinematic, 2 diffusion wave (ctrl_struct) */
    int32_t         riv_mode;
    int32_t         reserved;
    double          stepsize;   /* ctrl.stepsize: dt of Infil(), hydrol.c:22 */
    const double   *elem_f64;   /* [PB_E_NCOL ][nelem]  */
    const int32_t  *elem_i32;   /* [PB_EI_NCOL][nelem]  */
    const double   *riv_f64;    /* [PB_R_NCOL ][nriver] */
    const int32_t  *riv_i32;    /* [PB_RI_NCOL][nriver] */
} pihm_b200_mesh;

typedef struct pihm_b200_ctx pihm_b200_ctx;    /* opaque device context */
typedef struct pihm_b200_vec pihm_b200_vec;    /* opaque device vector   */

/* Error handling: every int-returning call gives 0 on success, <0 on failure;
 * the message of the last failure on the calling thread is kept here. */
const char     *pihm_b200_last_error(void);
int             pihm_b200_abi_version(void);
int             pihm_b200_device_count(void);

/* ------------------------------------------------------------------------
 * Context: device mirror of pihm->elem / pihm->river.
 *   replaces: nothing in the reference (it computes in place on the AoS);
 *   call after Initialize() (src/main.c:77) with the packed tables.
 * `reorder`: 0 keep reference order on the device, 1 locality reordering
 * (patch ordering of elements; the permutation is internal -- every host
 * buffer crossing this ABI is in reference order).
 * ---------------------------------------------------------------------- */
pihm_b200_ctx  *pihm_b200_create(const pihm_b200_mesh *mesh, int device,
                                 int reorder);
void            pihm_b200_destroy(pihm_b200_ctx *ctx);

/* ------------------------------------------------------------------------
 * Multi-GPU: the mesh is partitioned, one process (rank) per GPU (SURVEY 8(e);
 * the reference itself is shared-memory only).  A partition's local
 * mesh is an ordinary pihm_b200_mesh whose first nown_elem elements and
 * nown_riv rivers are owned and whose remaining entities are ghosts, grouped
 * by owner rank.  y / ydot of such a context hold the owned unknowns only.
 * ---------------------------------------------------------------------- */
typedef struct pihm_b200_partition pihm_b200_partition;
/* host-only: split `global` into nparts (the tables must stay alive) */
pihm_b200_partition *pihm_b200_partition_create(const pihm_b200_mesh *global,
                                                int nparts);
void            pihm_b200_partition_destroy(pihm_b200_partition *p);
/* sizes[8] = {ne_local, nr_local, ne_owned, nr_owned, n_neighbours,
 *             n_send_elem, n_send_riv, 0} */
int             pihm_b200_partition_sizes(const pihm_b200_partition *p, int part,
                                          int32_t *sizes);
/* local column tables, global ids of the local entities, exchange maps */
int             pihm_b200_partition_fill(const pihm_b200_partition *p, int part,
                    double *elem_f64, int32_t *elem_i32, double *riv_f64,
                    int32_t *riv_i32, int32_t *elem_gid, int32_t *riv_gid,
                    int32_t *nbr_rank, int32_t *send_e_ptr, int32_t *send_e_idx,
                    int32_t *recv_e_cnt, int32_t *send_r_ptr,
                    int32_t *send_r_idx, int32_t *recv_r_cnt);
pihm_b200_ctx  *pihm_b200_create_part(const pihm_b200_mesh *local, int device,
                                      int reorder, int nown_elem, int nown_riv);
int             pihm_b200_set_halo(pihm_b200_ctx *ctx, int n_neighbours,
                    const int32_t *nbr_rank, const int32_t *send_e_ptr,
                    const int32_t *send_e_idx, const int32_t *recv_e_cnt,
                    const int32_t *send_r_ptr, const int32_t *send_r_idx,
                    const int32_t *recv_r_cnt);
/* NCCL communicator: rank 0 makes the 128-byte id, the launcher broadcasts it.
 * After comm_init every RHS call starts with the halo exchange of {surf, gw[, fbr_gw]} /
 * {stage, gw} records and every reduction of the integrator is all-reduced over the
 * ranks.  Both run inside the library's own kernels over NVLink peer memory (buffers
 * mapped with CUDA IPC; set_halo must have been called before comm_init); if the mapping
 * is not available on every rank, grouped ncclSend/ncclRecv and ncclAllReduce do it. */
int             pihm_b200_comm_unique_id(void *out128);
int             pihm_b200_comm_init(pihm_b200_ctx *ctx, int rank, int nranks,
                                    const void *id128);
int64_t         pihm_b200_num_state_var_global(const pihm_b200_ctx *ctx);
/* pihm_b200_comm_init fails (all ranks, the set-up is collective) when the peer-memory halo
 * exchange cannot be mapped; PIHM_B200_NO_P2P=1 selects NCCL send/recv + ncclAllReduce instead,
 * PIHM_B200_P2P_OPTIONAL=1 allows the silent fallback.  pihm_b200_comm_paths reports what runs:
 * out[0] = 1 halo over peer memory inside the first RHS kernel (0: NCCL send/recv),
 * out[1] = 1 the ranks share one process (pihm_b200_comm_init_local). */
int             pihm_b200_comm_paths(const pihm_b200_ctx *ctx, int32_t *out2);
/* Ranks that live in ONE process (ctxs[r] = rank r, one context per rank on one or several
 * devices with peer access; one host thread per rank drives its integrator): the same
 * peer-memory halo exchange and in-kernel all-reduce, wired with plain pointers -- no NCCL.
 * pihm_b200_set_halo first; create the integrators of ALL ranks before the first solve. */
int             pihm_b200_comm_init_local(pihm_b200_ctx **ctxs, int nranks);
/* transport-free access to the exchange (single-process emulation of several
 * ranks on one GPU, used by the tests): packed send records of y, and the
 * ghost records of this context */
int             pihm_b200_halo_pack_host(pihm_b200_ctx *ctx,
                    const pihm_b200_vec *y, double *elem_rec, double *riv_rec);
int             pihm_b200_set_ghosts(pihm_b200_ctx *ctx, const double *elem_rec,
                                     const double *riv_rec);
/* number of ODE unknowns, NumStateVar() (src/ode.c:313-339) */
int64_t         pihm_b200_num_state_var(const pihm_b200_ctx *ctx);
/* run all work of this context on `stream` (a cudaStream_t); NULL = default */
int             pihm_b200_set_stream(pihm_b200_ctx *ctx, void *stream);
void           *pihm_b200_get_stream(const pihm_b200_ctx *ctx);
int             pihm_b200_synchronize(pihm_b200_ctx *ctx);

/* kernels launched by this context so far (bench.py's gpu_launches) */
long long       pihm_b200_launch_count(const pihm_b200_ctx *ctx);
/* perm[internal element] = reference element (0-based), [nelem] */
int             pihm_b200_get_permutation(const pihm_b200_ctx *ctx,
                                          int32_t *perm);

/* per-step host -> device pushes.
 * forc: [PB_F_NCOL][nelem] table, reference order.  The BC columns are only
 * read where the matching bc_type != 0. */
int             pihm_b200_set_forcing(pihm_b200_ctx *ctx, const double *forc);
/* single columns of the forcing table */
int             pihm_b200_set_forcing_col(pihm_b200_ctx *ctx, int col,
                                          const double *values);
/* river bc.head|flux (river_struct.h:63-68), [nriver] */
int             pihm_b200_set_river_bc(pihm_b200_ctx *ctx, const double *bc);
/* the river-edge overland flows of the previous RHS call that Infil() still
 * sees (elem.wf.ovlflow[j] for nabr[j] < 0): [3][nelem] table,
 * only river edges are read.  Zero after create (src/initialize.c:694). */
int             pihm_b200_set_stale_ovlflow(pihm_b200_ctx *ctx,
                                            const double *ovl);

/* ------------------------------------------------------------------------
 * RHS.   replaces: int ODE(realtype t, N_Vector y, N_Vector ydot, void *pihm)
 *        (src/ode.c:3-300) and everything under Hydrol() (src/hydrol.c:3-25).
 * Block layout [SURF;UNSAT;GW;RIVSTG;RIVGW;FBRUNSAT;FBRGW] (pihm_func.h:7-15).
 * Returns 0, or 1 when a NaN was produced (reference: CheckDy -> exit).
 * ---------------------------------------------------------------------- */
/* host buffers in reference order: H2D, kernels, D2H, synchronous */
int             pihm_b200_ode_host(pihm_b200_ctx *ctx, double t,
                                   const double *y, double *ydot);
/* device vectors (internal order), asynchronous on the context stream */
int             pihm_b200_ode(pihm_b200_ctx *ctx, double t,
                              const pihm_b200_vec *y, pihm_b200_vec *ydot);
/* NaN flag of the RHS calls since the last query (device flag, D2H) */
int             pihm_b200_check_nan(pihm_b200_ctx *ctx);
/* test hook: the inlined pow of the RHS kernels against libdevice pow() */
int             pihm_b200_test_pow(int n, const double *x, const double *y,
                                   double *fast, double *ref);
/* test hook: the branch-free division of the RHS kernels against `/`;
 * ok[i] = 0 where the operands are outside its domain (kernels recompute
 * such elements with the hardware division), -1 on an internal mismatch */
int             pihm_b200_test_div(int n, const double *a, const double *b,
                                   double *fast, double *ok, double *ref);
/* diagnostic: elements recomputed on the exact (hardware `/`, pow()) path
 * since the context was created */
long long       pihm_b200_slow_path_count(pihm_b200_ctx *ctx);
/* The part of Summary() (src/update.c:19-47) that feeds back into the RHS:
 * ws0.surf = y[SURF] for the next model step's Infil() (vert_flow.c:122).
 * Runs on the device (no y round trip); call after every SolveCVode. */
int             pihm_b200_summary(pihm_b200_ctx *ctx, const pihm_b200_vec *y);
/* Summary()/print need the wf.* fields of the last RHS call.
 * Writing them costs 144 B/element per call, so it is off unless asked for. */
int             pihm_b200_set_flux_recording(pihm_b200_ctx *ctx, int on);
/* flux diagnostics of the last RHS call, reference order:
 * elem_flux [PB_X_NCOL][nelem], rivflow [11][nriver]; either may be NULL */
int             pihm_b200_get_fluxes(pihm_b200_ctx *ctx, double *elem_flux,
                                     double *rivflow);

/* ------------------------------------------------------------------------
 * Summary() + MassBalance() on the device.
 *   replaces: void Summary(elem_struct *, river_struct *, N_Vector CV_Y,
 *             double stepsize) and MassBalance() (src/update.c:3-160), called
 *             after every SolveCVode (src/pihm.c:57).
 * The reference reads the wf.* fields the LAST ODE() call left in the structs
 *  -- usually a difference-quotient call of the Krylov solver.
 * With diagnostics on, the library remembers the input vector of the last RHS
 * call (and saves a copy if that vector is about to be overwritten by one of
 * its own vector kernels); pihm_b200_summary_mb evaluates that call once more
 * with the PB_X_* columns switched on (hidden state untouched), then
 *   wf.infil / wf.fbr_infil <- mass balance (PB_X_INFIL, PB_X_FBR_INFIL are
 *   overwritten like update.c:135,150), subrunoff, ws0 <- y (all components;
 *   also the ws0.surf column Infil() reads, like pihm_b200_summary).
 * Costs one extra RHS evaluation per model step instead of 144 B/element per
 * RHS call (pihm_b200_set_flux_recording) or a D2H of y plus a host loop.
 * ---------------------------------------------------------------------- */
int             pihm_b200_set_diagnostics(pihm_b200_ctx *ctx, int on);
/* ws0 = y: InitVar (src/initialize.c:598,612); also sets the ws0.surf column */
int             pihm_b200_set_ws0(pihm_b200_ctx *ctx, const pihm_b200_vec *y);
int             pihm_b200_summary_mb(pihm_b200_ctx *ctx, const pihm_b200_vec *y,
                                     double stepsize);
/* D2H, reference order; either may be NULL.  subrunoff [nelem] (the local of
 * MassBalance, update.c:128-133,154-158; Noah's runoff2), ws0 [N] in the block
 * layout of y.  The fluxes themselves: pihm_b200_get_fluxes. */
int             pihm_b200_get_summary(pihm_b200_ctx *ctx, double *subrunoff,
                                      double *ws0);

/* ------------------------------------------------------------------------
 * Forcing scatter + interception / snow / ET on the device.
 *   replaces: the per-element loops of ApplyMeteoForc and ApplyLai
 *             (src/forcing.c:134-160, 242-258) and
 *             void IntcpSnowEt(int t, double stepsize, elem_struct *elem,
 *             const calib_struct *cal) (src/is_sm_et.c:4-225), run every
 *             ctrl.etstep (src/pihm.c:27-48).
 * The host keeps the time-series interpolation (IntrplForc, a handful of
 * stations) and the month-of-year lookups, and passes their results by type;
 * the device gathers them per element, runs the element physics and writes
 * wf.pcpdrp / wf.edir / wf.ett straight into the forcing columns the RHS
 * reads -- no [nelem] array crosses PCIe.  ws.sneqv and ws.cmc live on the
 * device (zero after et_create, like Initialize without an .ic file).
 * ---------------------------------------------------------------------- */
/* et_f64 [PB_ET_NCOL][nelem], et_i32 [PB_ETI_NCOL][nelem], reference order */
int             pihm_b200_et_create(pihm_b200_ctx *ctx, const double *et_f64,
                                    const int32_t *et_i32);
typedef struct pihm_b200_et_step {
    double          stepsize;    /* (double)ctrl.etstep (pihm.c:41)          */
    double          cal_edir, cal_ec, cal_ett;   /* calib_struct             */
    double          meltf;       /* MonthlyMf(t) (forcing.c:603)             */
    int32_t         nmeteo, nlai, nlc;
    int32_t         reserved;
    const double   *meteo;       /* [nmeteo][PIHM_B200_NUM_METEO_VAR]:
                                    forc->meteo[k].value[] after IntrplForc  */
    const double   *lai;         /* [nlai] forc->lai[k].value[0]; may be NULL */
    const double   *lai_lc;      /* [nlc] MonthlyLai(t, lc) for lc = 1..nlc  */
    const double   *z0_lc;       /* [nlc] MonthlyRl(t, lc)                   */
} pihm_b200_et_step;
/* y: the state Summary() left in elem.ws (= CV_Y after the last model step,
 * or the initial condition); only its UNSAT and GW blocks are read */
int             pihm_b200_intcp_snow_et(pihm_b200_ctx *ctx,
                                        const pihm_b200_et_step *step,
                                        const pihm_b200_vec *y);
/* ws.sneqv / ws.cmc (e.g. from an .ic file): [nelem] each, reference order */
int             pihm_b200_et_set_state(pihm_b200_ctx *ctx, const double *sneqv,
                                       const double *cmc);
/* D2H of [PB_EO_NCOL][nelem], reference order (print / restart only) */
int             pihm_b200_et_get(pihm_b200_ctx *ctx, double *out);

/* ------------------------------------------------------------------------
 * Print accumulation on the device.
 *   replaces: UpdPrintVar (src/print.c:171-190: buffer[j] += *var[j];
 *             counter++) and the averaging of PrintData (src/print.c:192-251:
 *             buffer[j] / counter, then buffer[j] = 0, counter = 0).
 * The reference keeps one pointer per element and variable into its structs
 * (map_output.c) and therefore needs every printed field on the host every
 * step; here a print variable is (source, column) and its running sum lives
 * on the device: nothing is copied until PrintNow() (src/print.c:578, host
 * time logic, unchanged) says a record is due.
 * ---------------------------------------------------------------------- */
enum pihm_b200_print_src {
    PB_PS_STATE = 0,     /* column = block of y: 0 SURF 1 UNSAT 2 GW 3 RIVSTG 4 RIVGW
                            5 FBRUNSAT 6 FBRGW (elem.ws / river.ws after Summary) */
    PB_PS_ELEM_FLUX,     /* column = PB_X_*  (needs diagnostics or flux recording)  */
    PB_PS_RIV_FLUX,      /* column = 0..10, river.wf.rivflow[k]                     */
    PB_PS_ET             /* column = PB_EO_* (needs pihm_b200_et_create)            */
};
/* -> id >= 0 of the new print variable (one varctrl_struct), or < 0 */
int             pihm_b200_print_add(pihm_b200_ctx *ctx, int src, int column);
/* UpdPrintVar for the listed variables (the caller groups them by
 * upd_intvl like print.c:180); y = CV_Y, read by PB_PS_STATE variables */
int             pihm_b200_print_update(pihm_b200_ctx *ctx, const int32_t *ids,
                                       int n, const pihm_b200_vec *y);
/* PrintData for one variable: out[j] = buffer[j] / counter (buffer[j] when
 * counter == 0), reference order, [nelem] or [nriver]; then reset.
 * *counter_out (may be NULL) = number of updates averaged. */
int             pihm_b200_print_data(pihm_b200_ctx *ctx, int id, double *out,
                                     int32_t *counter_out);

/* ------------------------------------------------------------------------
 * Output files written from the device columns (second half).
 *   replaces: InitOutputFile (src/print.c:72-151), the record writing of
 *             PrintData (src/print.c:193-251) and PrintInit (src/print.c:253-313),
 *             i.e. the consumers of map_output.c's pointer lists.
 * All variables that are due at a print time are averaged, reset and packed by
 * one kernel and cross PCIe in ONE copy; the records have the reference's bytes:
 *   <name>.dat   { double t ; double value[nelem | nriver] }   per record
 *   <name>.txt   "<timestr>"\tvalue...\n  with %lf
 *   restart .ic  per element { cmc, sneqv, surf, unsat, gw [, fbr_unsat, fbr_gw] },
 *                then per river { stage, gw }
 * ---------------------------------------------------------------------- */
/* name: path without extension (varctrl.name, map_output.c: "<outputdir><project>.<var>") */
int             pihm_b200_print_open(pihm_b200_ctx *ctx, int id, const char *name,
                                     int ascii, int append);
/* ids: the variables PrintNow() (src/print.c:578, host time logic) says are due
 * at model time t; timestr = pihm_time.str of t */
int             pihm_b200_print_write(pihm_b200_ctx *ctx, const int32_t *ids, int n,
                                      int t, const char *timestr);
int             pihm_b200_print_close(pihm_b200_ctx *ctx);
/* device -> host copies made by the writers so far, and their bytes */
int             pihm_b200_print_io_stats(pihm_b200_ctx *ctx, int64_t *copies,
                                         int64_t *bytes);
/* cmc / sneqv: [nelem] host arrays in reference order, or NULL = the device ET
 * state (pihm_b200_et_create; zero without it) */
int             pihm_b200_write_ic(pihm_b200_ctx *ctx, const char *path,
                                   const pihm_b200_vec *y, const double *cmc,
                                   const double *sneqv);

/* ------------------------------------------------------------------------
 * Device-resident N_Vector.
 *   replaces: cvode/src/nvec_ser/nvector_serial.c:421-770 (ops) and
 *             :76-419 (constructors), same arithmetic per component.
 * ---------------------------------------------------------------------- */
pihm_b200_vec  *pihm_b200_vec_new(pihm_b200_ctx *ctx);
void            pihm_b200_vec_free(pihm_b200_vec *v);
int64_t         pihm_b200_vec_length(const pihm_b200_vec *v);
void           *pihm_b200_vec_devptr(pihm_b200_vec *v);
/* host (reference order) <-> device (internal order) */
int             pihm_b200_vec_upload(pihm_b200_vec *v, const double *host);
int             pihm_b200_vec_download(const pihm_b200_vec *v, double *host);

/* ------------------------------------------------------------------------
 * Pipelined host transfers (csrc/transfer.cu): the PCIe copies of a driver
 * that pushes forcing columns and pulls the state every model step
 *  run on a copy stream of the context's
 * own and overlap the kernels of the step.  Host buffers must be PINNED.
 *   replaces: nothing in the reference (host arrays are the model state
 *   there); the synchronous forms are pihm_b200_set_forcing_col and
 *   pihm_b200_vec_download.
 * forcing_prefetch: start copying `ncol` forcing columns (ids `cols[k]`,
 *   values `values_pinned[k][nelem]`, reference order) to the device; returns
 *   at once.  One prefetch may be outstanding.
 * forcing_commit: scatter the prefetched columns into the device tables, in
 *   stream order (after every RHS evaluation issued so far).  No-op when
 *   nothing was prefetched.
 * vec_download_async: copy `v` (reference order) into `host_pinned`; the
 *   copy is ordered behind everything issued so far and overlaps what is
 *   issued afterwards.  transfer_wait returns when the last such copy has
 *   arrived.  transfer_release frees the pipeline early (pihm_b200_destroy
 *   does it as well; nothing to do for a context that never used it).
 * ---------------------------------------------------------------------- */
int             pihm_b200_forcing_prefetch(pihm_b200_ctx *ctx, int ncol,
                                           const int *cols,
                                           const double *const *values_pinned);
int             pihm_b200_forcing_commit(pihm_b200_ctx *ctx);
int             pihm_b200_vec_download_async(const pihm_b200_vec *v,
                                             double *host_pinned);
int             pihm_b200_transfer_wait(pihm_b200_ctx *ctx);
int             pihm_b200_transfer_release(pihm_b200_ctx *ctx);

void            pihm_b200_nv_linearsum(double a, const pihm_b200_vec *x,
                                       double b, const pihm_b200_vec *y,
                                       pihm_b200_vec *z);
void            pihm_b200_nv_const(double c, pihm_b200_vec *z);
void            pihm_b200_nv_prod(const pihm_b200_vec *x,
                                  const pihm_b200_vec *y, pihm_b200_vec *z);
void            pihm_b200_nv_div(const pihm_b200_vec *x,
                                 const pihm_b200_vec *y, pihm_b200_vec *z);
void            pihm_b200_nv_scale(double c, const pihm_b200_vec *x,
                                   pihm_b200_vec *z);
void            pihm_b200_nv_abs(const pihm_b200_vec *x, pihm_b200_vec *z);
void            pihm_b200_nv_inv(const pihm_b200_vec *x, pihm_b200_vec *z);
void            pihm_b200_nv_addconst(const pihm_b200_vec *x, double b,
                                      pihm_b200_vec *z);
double          pihm_b200_nv_dotprod(const pihm_b200_vec *x,
                                     const pihm_b200_vec *y);
double          pihm_b200_nv_maxnorm(const pihm_b200_vec *x);
double          pihm_b200_nv_wrmsnorm(const pihm_b200_vec *x,
                                      const pihm_b200_vec *w);
double          pihm_b200_nv_min(const pihm_b200_vec *x);

/* ------------------------------------------------------------------------
 * Integrator: variable-order BDF + Newton + scaled GMRES(5), all vector
 * work on the device.
 *   replaces: SetCVodeParam / SolveCVode / AdjCVodeMaxStep (src/ode.c:340-560)
 *   and underneath them CVodeInit/CVode (cvode/src/cvode/cvode.c:438,1074),
 *   CVSpgmr (cvode_spgmr.c:112), SpgmrSolve (sundials_spgmr.c:165),
 *   CVSpilsDQJtimes (cvode_spils.c:665).
 * ---------------------------------------------------------------------- */
typedef struct pihm_b200_cvode pihm_b200_cvode;

typedef struct pihm_b200_cvode_param {
    double          reltol;      /* ctrl.reltol   (ode.c:388) */
    double          abstol;      /* ctrl.abstol                */
    double          initstep;    /* ctrl.initstep (ode.c:402)  */
    double          maxstep;     /* ctrl.maxstep  (ode.c:414)  */
    int64_t         mxsteps;     /* ctrl.stepsize*10 (ode.c:420) */
    int32_t         stab_lim_det;/* TRUE (ode.c:408)           */
    int32_t         maxl;        /* 0 -> 5 (CVSpgmr(.., PREC_NONE, 0)) */
} pihm_b200_cvode_param;

/* counters, same meaning as CVodeGetNum* / CVSpilsGetNum* */
typedef struct pihm_b200_cvode_stats {
    int64_t         nst, nfe, nni, ncfn, netf, nli, ncfl, nfeLS, njtimes;
    int64_t         nor, nsetups;
    int32_t         qlast, qcur;
    double          hlast, hcur, tcur;
} pihm_b200_cvode_stats;

/* SetCVodeParam: CVodeCreate(CV_BDF,CV_NEWTON)+CVodeInit(t0=0, y)+tolerances+
 * CVSpgmr.  y is copied into the integrator's history at each (re)init. */
pihm_b200_cvode *pihm_b200_cvode_create(pihm_b200_ctx *ctx);
void            pihm_b200_cvode_destroy(pihm_b200_cvode *cv);
int             pihm_b200_cvode_init(pihm_b200_cvode *cv,
                                     const pihm_b200_cvode_param *p,
                                     double t0, const pihm_b200_vec *y0);
/* CVodeSetMaxStep (used by AdjCVodeMaxStep, ode.c:551) */
int             pihm_b200_cvode_set_max_step(pihm_b200_cvode *cv, double hmax);
/* SolveCVode: CVodeSetStopTime(tout); CVode(tout, CV_NORMAL).  On return y
 * holds y(tout), *tret = tout.  Returns 0/1 (CV_SUCCESS/CV_TSTOP_RETURN) or a
 * negative CVODE flag value. */
int             pihm_b200_cvode_solve(pihm_b200_cvode *cv, double tout,
                                      pihm_b200_vec *y, double *tret);
int             pihm_b200_cvode_get_stats(const pihm_b200_cvode *cv,
                                          pihm_b200_cvode_stats *st);
/* The linear solver alone: the lsolve hook for a CVODE that keeps its own BDF /
 * Newton stepper (cv_mem->cv_lsolve, cvode_impl.h:198-211).
 *   replaces: CVSpgmrSolve (cvode/src/cvode/cvode_spgmr.c:355-441) and under it
 *   SpgmrSolve, ModifiedGS, QRfact/QRsol, CVSpilsAtimes, CVSpilsDQJtimes.
 * cv: an engine from pihm_b200_cvode_create (no cvode_init needed); tn, gamma,
 * tq4 = cv_mem->cv_tn, cv_gamma, cv_tq[4]; mnewt = cv_mem->cv_mnewt; b, weight,
 * ycur, fcur as handed to lsolve.  Returns 0, >0 (recoverable) or <0; the
 * solution overwrites b.  Counters: pihm_b200_cvode_get_stats (nli, ncfl,
 * nfeLS, njtimes).  The attach code is in INTEGRATION.md 3a'. */
int             pihm_b200_spgmr_solve(pihm_b200_cvode *cv, double tn,
                                      double gamma, double tq4, int mnewt,
                                      pihm_b200_vec *b,
                                      const pihm_b200_vec *weight,
                                      const pihm_b200_vec *ycur,
                                      const pihm_b200_vec *fcur);
/* In-situ profile of the integrator (no reference counterpart; bench.py): CUDA events around each
 * RHS evaluation it issues and the host time spent at its synchronisation points.  on = 1 clears
 * the accumulators.  out5 = {ms inside cvode_solve, ms of host waiting, number of host
 * synchronisations, ms of RHS kernels event to event, RHS evaluations timed}. */
int             pihm_b200_cvode_profile(pihm_b200_cvode *cv, int on);
int             pihm_b200_cvode_get_profile(pihm_b200_cvode *cv, double *out5);
/* on = 2 in pihm_b200_cvode_profile also brackets every vector kernel of the integrator with CUDA events
 * (programmatic serialization off while it lasts).  Per kernel that ran: name (24 bytes each), summed
 * event-to-event ms, bytes read + written (vector passes x 8 N) and launches; returns the entry count. */
int             pihm_b200_cvode_get_kernel_profile(pihm_b200_cvode *cv, int cap,
                    char *names, double *ms, double *bytes, int64_t *launches);
/* AdjCVodeMaxStep (ode.c:500-560) on the integrator's own counters */
typedef struct pihm_b200_maxstep_ctrl {
    double          maxstep, stepsize, stmin, nncfn, nnimax, nnimin, decr, incr;
} pihm_b200_maxstep_ctrl;
int             pihm_b200_adj_cvode_max_step(pihm_b200_cvode *cv,
                                             pihm_b200_maxstep_ctrl *c);

#ifdef __cplusplus
}
#endif
#endif /* PIHM_B200_H */
