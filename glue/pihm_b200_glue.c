/*
 * pihm_b200_glue.c -- the translation unit that takes the place of src/ode.c in the MM-PIHM
 * drivers (pihm, pihm-fbr).  PRODUCT code: compiled against the reference's own headers
 * (src/include/pihm.h) and linked with libpihm_b200.so; main.c, pihm.c, spinup.c and every
 * other reference source are linked unchanged.
 *
 * It defines the five external symbols of src/ode.c that the rest of the driver uses
 *
 *     int  ODE(realtype t, N_Vector y, N_Vector ydot, void *pihm_data)     src/ode.c:3
 *     int  NumStateVar(void)                                               src/ode.c:313
 *     void SetCVodeParam(pihm_struct, void *cvode_mem, N_Vector CV_Y)       src/ode.c:340
 *     void SolveCVode(int starttime, int *t, int nextptr, double cputime,
 *                     void *cvode_mem, N_Vector CV_Y)                      src/ode.c:456
 *     void AdjCVodeMaxStep(void *cvode_mem, ctrl_struct *ctrl)             src/ode.c:500
 *
 * (prototypes: src/include/pihm_func.h:104,231,294,296; CheckDy / SetAbsTol of ode.c are
 * only used inside ode.c -- CheckDy is the device NaN flag here, SetAbsTol is BGC/Cycles only)
 * plus the lifecycle hooks of SURVEY 8(b):
 *
 *     PihmB200Init(pihm)          pack elem_struct / river_struct -> column tables, create the
 *                                 device context (called by the first SetCVodeParam: after
 *                                 Initialize(), src/main.c:77)
 *     PihmB200PushForcing(pihm)   wf.pcpdrp / wf.edir / wf.ett, bc.head|flux, river bc ->
 *                                 device (after ApplyBc / IntcpSnowEt, src/pihm.c:22-48)
 *     PihmB200PullState(pihm, y)  y -> NV_DATA(CV_Y) and the wf.* fields the last RHS call left
 *                                 behind -> pihm->elem[].wf / pihm->river[].wf (before Summary,
 *                                 src/pihm.c:57, and the print pointers of map_output.c)
 *     PihmB200Free()              at exit
 *
 * The hooks are called from inside SetCVodeParam / SolveCVode, so the driver needs no edit.
 *
 * Two routes, chosen by the environment variable PIHM_B200_ROUTE:
 *   2 (default)  the library's device integrator (pihm_b200_cvode_*); cvode_mem, which
 *                Initialize() created (src/initialize.c:204), stays unused
 *   1            the driver's own CVODE (cvode_mem) steps on the device-resident N_Vector
 *                (N_VNew_PihmB200) with PihmB200_ODE as its CVRhsFn and CVSpgmr untouched
 * Both give the same bits (tests/test_dropin_gpu.py).  CV_Y stays the host vector main.c made
 * with N_VNew(NumStateVar()) (src/main.c:69): it is the host mirror Summary() reads.
 *
 * There is no CPU path: every function below fails through PIHMexit when the CUDA library
 * reports an error.
 */
#include "pihm.h"
#include "pihm_b200.h"
#include "pihm_b200_sundials.h"

static struct
{
    pihm_struct     pihm;
    pihm_b200_ctx  *ctx;
    pihm_b200_cvode *cv;        /* route 2 */
    pihm_b200_vec  *y;          /* route 2: the integrator's output vector; also host-vector ODE() */
    pihm_b200_vec  *ydot;
    N_Vector        nv_y;       /* route 1: device N_Vector handed to CVODE */
    int             route;
    int             in_solve;
    double         *forc;       /* [PB_F_NCOL][nelem] staging */
    double         *rivbc;      /* [nriver] */
    double         *xf;         /* [PB_X_NCOL][nelem] */
    double         *rivflow;    /* [11][nriver] */
} G;

static void Fail(const char *what)
{
    PIHMprintf(VL_ERROR, "Error: libpihm_b200: %s: %s\n", what, pihm_b200_last_error());
    PIHMexit(EXIT_FAILURE);
}

int NumStateVar(void)
{
    /* src/ode.c:313-339 for the pihm / pihm-fbr configurations (no BGC, no Cycles) */
    int             nsv = 3 * nelem + 2 * nriver;

#if defined(_BGC_) || defined(_CYCLES_)
# error "libpihm_b200 covers the pihm and pihm-fbr drivers (SURVEY 8(a)); BGC / Cycles are out of scope"
#endif
#if defined(_FBR_)
    nsv += 2 * nelem;
#endif
    return nsv;
}

/*
 * The packer: every field of elem_struct / river_struct / ctrl_struct the RHS reads after
 * Initialize(), as dense column tables [NCOL][n].
 */
void PihmB200Init(pihm_struct pihm)
{
    pihm_b200_mesh  m;
    double         *ef, *rf, *ovl;
    int32_t        *ei, *ri;
    int             i, j;
    const char     *env;
    size_t          ne = (size_t)nelem, nr = (size_t)nriver;

    if (G.ctx != NULL)
    {
        return;
    }
    memset(&m, 0, sizeof(m));
    ef = (double *)calloc(PB_E_NCOL * ne + 1, sizeof(double));
    ei = (int32_t *)calloc(PB_EI_NCOL * ne + 1, sizeof(int32_t));
    rf = (double *)calloc(PB_R_NCOL * nr + 1, sizeof(double));
    ri = (int32_t *)calloc(PB_RI_NCOL * nr + 1, sizeof(int32_t));
    ovl = (double *)calloc(3 * ne + 1, sizeof(double));

#define EF(c) ef[(size_t)(c) * ne + i]
#define EI(c) ei[(size_t)(c) * ne + i]
    for (i = 0; i < nelem; i++)
    {
        const elem_struct *e = &pihm->elem[i];

        /* topo_struct, elem_struct.h:28-52 */
        EF(PB_E_AREA) = e->topo.area;
        EF(PB_E_ZMIN) = e->topo.zmin;
        EF(PB_E_ZMAX) = e->topo.zmax;
        for (j = 0; j < NUM_EDGE; j++)
        {
            EF(PB_E_EDGE0 + j) = e->topo.edge[j];
            EF(PB_E_NABRDIST0 + j) = e->topo.nabrdist[j];
            EF(PB_E_NABRX0 + j) = e->topo.nabr_x[j];
            EF(PB_E_NABRY0 + j) = e->topo.nabr_y[j];
            /* nabr[j] after InitRiver(): > 0 element, < 0 -(river), 0 boundary (elem_struct.h:1560) */
            EI(PB_EI_NABR0 + j) = e->nabr[j];
            EI(PB_EI_BC0 + j) = e->attrib.bc_type[j];
            /* the river-edge overland flows Infil() still sees from the previous call (H2) */
            ovl[(size_t)j * ne + i] = e->wf.ovlflow[j];
        }
        /* soil_struct, elem_struct.h:55-96 */
        EF(PB_E_DEPTH) = e->soil.depth;
        EF(PB_E_KSATH) = e->soil.ksath;
        EF(PB_E_KSATV) = e->soil.ksatv;
        EF(PB_E_KINFV) = e->soil.kinfv;
        EF(PB_E_DINF) = e->soil.dinf;
        EF(PB_E_ALPHA) = e->soil.alpha;
        EF(PB_E_BETA) = e->soil.beta;
        EF(PB_E_POROSITY) = e->soil.porosity;
        EF(PB_E_DMAC) = e->soil.dmac;
        EF(PB_E_KMACH) = e->soil.kmach;
        EF(PB_E_KMACV) = e->soil.kmacv;
        EF(PB_E_AREAFV) = e->soil.areafv;
        EF(PB_E_AREAFH) = e->soil.areafh;
        /* lc_struct.rough (:136), pstate_struct.rzd (:240) */
        EF(PB_E_ROUGH) = e->lc.rough;
        EF(PB_E_RZD) = e->ps.rzd;
#if defined(_FBR_)
        /* fractured bedrock: topo.zbed, geol_struct (:99-111), attrib.fbrbc_type */
        EF(PB_E_ZBED) = e->topo.zbed;
        EF(PB_E_GDEPTH) = e->geol.depth;
        EF(PB_E_GKSATH) = e->geol.ksath;
        EF(PB_E_GKSATV) = e->geol.ksatv;
        EF(PB_E_GALPHA) = e->geol.alpha;
        EF(PB_E_GBETA) = e->geol.beta;
        EF(PB_E_GPOROSITY) = e->geol.porosity;
        for (j = 0; j < NUM_EDGE; j++)
        {
            EI(PB_EI_FBRBC0 + j) = e->attrib.fbrbc_type[j];
        }
#endif
    }
#undef EF
#undef EI
#define RF(c) rf[(size_t)(c) * nr + i]
#define RI(c) ri[(size_t)(c) * nr + i]
    for (i = 0; i < nriver; i++)
    {
        const river_struct *r = &pihm->river[i];

        /* river_struct.h:137-168: topology, river_topo_struct, shp_struct, matl_struct */
        RI(PB_RI_LEFTELE) = r->leftele;
        RI(PB_RI_RIGHTELE) = r->rightele;
        RI(PB_RI_DOWN) = r->down;
        RI(PB_RI_BCTYPE) = r->attrib.riverbc_type;
        RI(PB_RI_INTRPL_ORD) = r->shp.intrpl_ord;
        RF(PB_R_AREA) = r->topo.area;
        RF(PB_R_ZMIN) = r->topo.zmin;
        RF(PB_R_ZMAX) = r->topo.zmax;
        RF(PB_R_ZBED) = r->topo.zbed;
        RF(PB_R_NODE_ZMAX) = r->topo.node_zmax;
        RF(PB_R_DIST_LEFT) = r->topo.dist_left;
        RF(PB_R_DIST_RIGHT) = r->topo.dist_right;
        RF(PB_R_SHP_DEPTH) = r->shp.depth;
        RF(PB_R_SHP_COEFF) = r->shp.coeff;
        RF(PB_R_SHP_LENGTH) = r->shp.length;
        RF(PB_R_SHP_WIDTH) = r->shp.width;
        RF(PB_R_ROUGH) = r->matl.rough;
        RF(PB_R_CWR) = r->matl.cwr;
        RF(PB_R_KSATH) = r->matl.ksath;
        RF(PB_R_KSATV) = r->matl.ksatv;
        RF(PB_R_BEDTHICK) = r->matl.bedthick;
        RF(PB_R_POROSITY) = r->matl.porosity;
    }
#undef RF
#undef RI

    m.nelem = nelem;
    m.nriver = nriver;
#if defined(_FBR_)
    m.fbr = 1;
#endif
    m.surf_mode = pihm->ctrl.surf_mode;
    m.riv_mode = pihm->ctrl.riv_mode;
    m.stepsize = (double)pihm->ctrl.stepsize;   /* dt of Infil(), hydrol.c:22 */
    m.elem_f64 = ef;
    m.elem_i32 = ei;
    m.riv_f64 = rf;
    m.riv_i32 = ri;

    env = getenv("PIHM_B200_DEVICE");
    G.ctx = pihm_b200_create(&m, (env != NULL) ? atoi(env) : 0, 1);
    if (G.ctx == NULL)
    {
        Fail("pihm_b200_create");
    }
    if (pihm_b200_num_state_var(G.ctx) != (int64_t)NumStateVar())
    {
        PIHMprintf(VL_ERROR, "Error: libpihm_b200 was packed with a different NumStateVar().\n");
        PIHMexit(EXIT_FAILURE);
    }
    if (pihm_b200_set_stale_ovlflow(G.ctx, ovl) != 0 || pihm_b200_set_diagnostics(G.ctx, 1) != 0)
    {
        Fail("pihm_b200_set_diagnostics");
    }
    free(ef);
    free(ei);
    free(rf);
    free(ri);
    free(ovl);

    G.pihm = pihm;
    G.forc = (double *)calloc(PB_F_NCOL * ne + 1, sizeof(double));
    G.rivbc = (double *)calloc(nr + 1, sizeof(double));
    G.xf = (double *)calloc(PB_X_NCOL * ne + 1, sizeof(double));
    G.rivflow = (double *)calloc(PIHM_B200_NUM_RIVFLX * nr + 1, sizeof(double));
    G.y = pihm_b200_vec_new(G.ctx);
    G.ydot = pihm_b200_vec_new(G.ctx);
    if (G.y == NULL || G.ydot == NULL)
    {
        Fail("pihm_b200_vec_new");
    }
    env = getenv("PIHM_B200_ROUTE");
    G.route = (env != NULL && atoi(env) == 1) ? 1 : 2;
    PIHMprintf(VL_VERBOSE, "libpihm_b200: %d elements, %d river segments on the device, route %d (%s)\n",
        nelem, nriver, G.route, (G.route == 1) ? "CVODE on the device N_Vector" : "device integrator");
}

void PihmB200Free(void)
{
    if (G.ctx == NULL)
    {
        return;
    }
    if (G.cv != NULL)
    {
        pihm_b200_cvode_destroy(G.cv);
    }
    if (G.nv_y != NULL)
    {
        N_VDestroy(G.nv_y);
    }
    pihm_b200_vec_free(G.y);
    pihm_b200_vec_free(G.ydot);
    pihm_b200_destroy(G.ctx);
    free(G.forc);
    free(G.rivbc);
    free(G.xf);
    free(G.rivflow);
    memset(&G, 0, sizeof(G));
}

/*
 * What ApplyBc / IntcpSnowEt / Summary wrote on the host since the last solve and the RHS reads
 * (SURVEY Appendix D 1-2): wf.pcpdrp, wf.edir, wf.ett (is_sm_et.c:126-223), ws0.surf
 * (update.c:47), bc.head|flux per edge (forcing.c:55-86), river bc.
 */
void PihmB200PushForcing(pihm_struct pihm)
{
    int             i, j;
    size_t          ne = (size_t)nelem;

    for (i = 0; i < nelem; i++)
    {
        const elem_struct *e = &pihm->elem[i];

        G.forc[(size_t)PB_F_PCPDRP * ne + i] = e->wf.pcpdrp;
        G.forc[(size_t)PB_F_EDIR * ne + i] = e->wf.edir;
        G.forc[(size_t)PB_F_ETT * ne + i] = e->wf.ett;
        G.forc[(size_t)PB_F_WS0SURF * ne + i] = e->ws0.surf;
        for (j = 0; j < NUM_EDGE; j++)
        {
            G.forc[(size_t)(PB_F_BC0 + j) * ne + i] = e->bc.head[j];    /* union with bc.flux[j] */
#if defined(_FBR_)
            G.forc[(size_t)(PB_F_FBRBC0 + j) * ne + i] = e->fbr_bc.head[j];
#endif
        }
    }
    for (i = 0; i < nriver; i++)
    {
        G.rivbc[i] = pihm->river[i].bc.head;                            /* union with bc.flux */
    }
    if (pihm_b200_set_forcing(G.ctx, G.forc) != 0 ||
        (nriver > 0 && pihm_b200_set_river_bc(G.ctx, G.rivbc) != 0))
    {
        Fail("pihm_b200_set_forcing");
    }
}

/* the wf.* fields an ODE() call leaves in the structs (wflux_struct, elem_struct.h:407-477;
 * river wf.rivflow[11]): from the device columns into pihm->elem / pihm->river */
static void ScatterFluxes(pihm_struct pihm)
{
    int             i, j;
    size_t          ne = (size_t)nelem, nr = (size_t)nriver;

    if (pihm_b200_get_fluxes(G.ctx, G.xf, (nriver > 0) ? G.rivflow : NULL) != 0)
    {
        Fail("pihm_b200_get_fluxes");
    }
#define XF(c) G.xf[(size_t)(c) * ne + i]
    for (i = 0; i < nelem; i++)
    {
        wflux_struct   *wf = &pihm->elem[i].wf;

        for (j = 0; j < NUM_EDGE; j++)
        {
            wf->ovlflow[j] = XF(PB_X_OVL0 + j);
            wf->subsurf[j] = XF(PB_X_SUB0 + j);
        }
        wf->infil = XF(PB_X_INFIL);
        wf->rechg = XF(PB_X_RECHG);
        wf->edir_surf = XF(PB_X_EDIR_SURF);
        wf->edir_unsat = XF(PB_X_EDIR_UNSAT);
        wf->edir_gw = XF(PB_X_EDIR_GW);
        wf->ett_unsat = XF(PB_X_ETT_UNSAT);
        wf->ett_gw = XF(PB_X_ETT_GW);
#if defined(_FBR_)
        wf->fbr_infil = XF(PB_X_FBR_INFIL);
        wf->fbr_rechg = XF(PB_X_FBR_RECHG);
        for (j = 0; j < NUM_EDGE; j++)
        {
            wf->fbrflow[j] = XF(PB_X_FBRFLOW0 + j);
        }
#endif
    }
#undef XF
    for (i = 0; i < nriver; i++)
    {
        for (j = 0; j < NUM_RIVFLX; j++)
        {
            pihm->river[i].wf.rivflow[j] = G.rivflow[(size_t)j * nr + i];
        }
    }
}

/*
 * Before Summary() (src/pihm.c:57): y(tout) into the host vector and the fluxes of the last
 * RHS call into the structs.  pihm_b200_summary_mb evaluates that call once more with the
 * flux columns on (hidden state untouched) and keeps the device copy of ws0 in step with what
 * the host Summary() is about to do (update.c:47,94).
 */
void PihmB200PullState(pihm_struct pihm, N_Vector CV_Y)
{
    pihm_b200_vec  *y = (G.route == 1) ? N_VPihmB200_Device(G.nv_y) : G.y;

    if (pihm_b200_vec_download(y, NV_DATA(CV_Y)) != 0)
    {
        Fail("pihm_b200_vec_download");
    }
    if (pihm_b200_summary_mb(G.ctx, y, (double)pihm->ctrl.stepsize) != 0)
    {
        Fail("pihm_b200_summary_mb");
    }
    ScatterFluxes(pihm);
    /* CheckDy (src/ode.c:302-311): NAN in dy ends the run */
    if (pihm_b200_check_nan(G.ctx) != 0)
    {
        PIHMprintf(VL_ERROR, "Error: NAN error in dy (device flag) at step ending %d\n",
            pihm->ctrl.tout[pihm->ctrl.cstep + 1]);
        PIHMexit(EXIT_FAILURE);
    }
}

/*
 * ODE() with the reference signature.  Device-resident vectors (N_VNew_PihmB200) go straight
 * to the kernels; host vectors (the N_Vector main.c made) are staged through the device --
 * that is the compatibility path for a caller that hands ODE to a CVODE of its own on host
 * vectors: forcing push, H2D of y, kernels, D2H of ydot, and the wf.* fields scattered into the
 * structs like the reference leaves them.
 */
int ODE(realtype t, N_Vector y, N_Vector ydot, void *pihm_data)
{
    pihm_struct     pihm = (pihm_struct)pihm_data;
    pihm_b200_vec  *dy = N_VPihmB200_Device(y);
    pihm_b200_vec  *dydot = N_VPihmB200_Device(ydot);

    if (G.ctx == NULL)
    {
        PihmB200Init(pihm);
    }
    if (dy != NULL && dydot != NULL)
    {
        return (pihm_b200_ode(G.ctx, (double)t, dy, dydot) != 0) ? -1 : 0;
    }
    if (!G.in_solve)
    {
        PihmB200PushForcing(pihm);
    }
    if (pihm_b200_vec_upload(G.y, NV_DATA(y)) != 0 ||
        pihm_b200_ode(G.ctx, (double)t, G.y, G.ydot) != 0 ||
        pihm_b200_vec_download(G.ydot, NV_DATA(ydot)) != 0)
    {
        Fail("ODE");
    }
    if (pihm_b200_check_nan(G.ctx) != 0)
    {
        PIHMprintf(VL_ERROR, "Error: NAN error in dy (device flag) at %lf\n", (double)t);
        PIHMexit(EXIT_FAILURE);
    }
    if (!G.in_solve)
    {
        /* flux columns of this very call (replayed with the columns on) */
        if (pihm_b200_summary_mb(G.ctx, G.y, (double)pihm->ctrl.stepsize) != 0)
        {
            Fail("pihm_b200_summary_mb");
        }
        ScatterFluxes(pihm);
    }
    return 0;
}

void SetCVodeParam(pihm_struct pihm, void *cvode_mem, N_Vector CV_Y)
{
    static int      reset;
    int             cv_flag;

    PihmB200Init(pihm);         /* first call: right after Initialize() / MapOutput (src/main.c:77-112) */
    G.pihm = pihm;

    pihm->ctrl.maxstep = pihm->ctrl.stepsize;

    if (G.route == 2)
    {
        pihm_b200_cvode_param p;

        if (G.cv == NULL && (G.cv = pihm_b200_cvode_create(G.ctx)) == NULL)
        {
            Fail("pihm_b200_cvode_create");
        }
        memset(&p, 0, sizeof(p));
        p.reltol = pihm->ctrl.reltol;                   /* CVodeSStolerances, ode.c:388 */
        p.abstol = pihm->ctrl.abstol;
        p.initstep = pihm->ctrl.initstep;               /* CVodeSetInitStep, ode.c:402 */
        p.maxstep = pihm->ctrl.maxstep;                 /* CVodeSetMaxStep, ode.c:414 */
        p.mxsteps = (int64_t)pihm->ctrl.stepsize * 10;  /* CVodeSetMaxNumSteps, ode.c:420 */
        p.stab_lim_det = 1;                             /* CVodeSetStabLimDet(TRUE), ode.c:408 */
        p.maxl = 0;                                     /* CVSpgmr(PREC_NONE, 0), ode.c:426 */
        /* CVodeInit / CVodeReInit(t0 = 0, CV_Y), ode.c:350-367 */
        if (pihm_b200_vec_upload(G.y, NV_DATA(CV_Y)) != 0 ||
            pihm_b200_set_ws0(G.ctx, G.y) != 0 ||
            pihm_b200_cvode_init(G.cv, &p, 0.0, G.y) != 0)
        {
            Fail("pihm_b200_cvode_init");
        }
        reset = 1;
        return;
    }

    /* route 1: the statements of src/ode.c:350-431 with the device vector and PihmB200_ODE */
    if (G.nv_y == NULL && (G.nv_y = N_VNew_PihmB200(G.ctx)) == NULL)
    {
        Fail("N_VNew_PihmB200");
    }
    memcpy(NV_DATA_S(G.nv_y), NV_DATA(CV_Y), sizeof(realtype) * NumStateVar());
    if (N_VPihmB200_Push(G.nv_y) != 0 || pihm_b200_set_ws0(G.ctx, N_VPihmB200_Device(G.nv_y)) != 0)
    {
        Fail("N_VPihmB200_Push");
    }
    if (reset)
    {
        cv_flag = CVodeReInit(cvode_mem, 0.0, G.nv_y);
    }
    else
    {
        cv_flag = CVodeInit(cvode_mem, PihmB200_ODE, 0.0, G.nv_y);
        reset = 1;
    }
    if (!CheckCVodeFlag(cv_flag))
    {
        PIHMexit(EXIT_FAILURE);
    }
    if (!CheckCVodeFlag(CVodeSStolerances(cvode_mem, (realtype)pihm->ctrl.reltol, (realtype)pihm->ctrl.abstol)) ||
        !CheckCVodeFlag(CVodeSetUserData(cvode_mem, G.ctx)) ||
        !CheckCVodeFlag(CVodeSetInitStep(cvode_mem, (realtype)pihm->ctrl.initstep)) ||
        !CheckCVodeFlag(CVodeSetStabLimDet(cvode_mem, TRUE)) ||
        !CheckCVodeFlag(CVodeSetMaxStep(cvode_mem, (realtype)pihm->ctrl.maxstep)) ||
        !CheckCVodeFlag(CVodeSetMaxNumSteps(cvode_mem, pihm->ctrl.stepsize * 10)) ||
        !CheckCVodeFlag(CVSpgmr(cvode_mem, PREC_NONE, 0)))
    {
        PIHMexit(EXIT_FAILURE);
    }
}

void SolveCVode(int starttime, int *t, int nextptr, double cputime, void *cvode_mem, N_Vector CV_Y)
{
    realtype        solvert;
    realtype        tout = (realtype)(nextptr - starttime);
    pihm_t_struct   pihm_time;
    int             cv_flag;

    if (G.ctx == NULL)
    {
        PIHMprintf(VL_ERROR, "Error: SolveCVode before SetCVodeParam.\n");
        PIHMexit(EXIT_FAILURE);
    }
    PihmB200PushForcing(G.pihm);
    G.in_solve = 1;
    if (G.route == 2)
    {
        double          tret = 0.0;

        /* CVodeSetStopTime(tout) + CVode(tout, CV_NORMAL), ode.c:466-476; negative flag -> exit */
        cv_flag = pihm_b200_cvode_solve(G.cv, (double)tout, G.y, &tret);
        if (cv_flag < 0)
        {
            PIHMprintf(VL_ERROR, "CVode error %d: %s\n", cv_flag, pihm_b200_last_error());
            PIHMexit(EXIT_FAILURE);
        }
        solvert = (realtype)tret;
    }
    else
    {
        if (!CheckCVodeFlag(CVodeSetStopTime(cvode_mem, tout)) ||
            !CheckCVodeFlag(CVode(cvode_mem, tout, G.nv_y, &solvert, CV_NORMAL)))
        {
            PIHMexit(EXIT_FAILURE);
        }
    }
    G.in_solve = 0;
    PihmB200PullState(G.pihm, CV_Y);

    *t = (int)round(solvert) + starttime;

    pihm_time = PIHMTime(*t);

    if (debug_mode)
    {
        PIHMprintf(VL_NORMAL, " Step = %s (%d)\n", pihm_time.str, *t);
    }
    else if (spinup_mode)
    {
        if (pihm_time.t % DAYINSEC == 0)
        {
            PIHMprintf(VL_NORMAL, " Step = %s\n", pihm_time.str);
        }
    }
    else if (pihm_time.t % 3600 == 0)
    {
        PIHMprintf(VL_NORMAL, " Step = %s (cputime %f)\n", pihm_time.str, cputime);
    }
}

void AdjCVodeMaxStep(void *cvode_mem, ctrl_struct *ctrl)
{
    if (G.route == 2)
    {
        /* the controller of ode.c:500-560 on the device integrator's own counters */
        pihm_b200_maxstep_ctrl c;

        c.maxstep = ctrl->maxstep;
        c.stepsize = (double)ctrl->stepsize;
        c.stmin = ctrl->stmin;
        c.nncfn = ctrl->nncfn;
        c.nnimax = ctrl->nnimax;
        c.nnimin = ctrl->nnimin;
        c.decr = ctrl->decr;
        c.incr = ctrl->incr;
        if (pihm_b200_adj_cvode_max_step(G.cv, &c) != 0)
        {
            Fail("pihm_b200_adj_cvode_max_step");
        }
        ctrl->maxstep = c.maxstep;
    }
    else
    {
        /* route 1: CVODE's counters (CVodeGetNumSteps / ...NonlinSolvConvFails / ...NonlinSolvIters) */
        static long int nst0, ncfn0, nni0;
        long int        nst, ncfn, nni;
        double          nsteps, nfails, niters;

        if (!CheckCVodeFlag(CVodeGetNumSteps(cvode_mem, &nst)) ||
            !CheckCVodeFlag(CVodeGetNumNonlinSolvConvFails(cvode_mem, &ncfn)) ||
            !CheckCVodeFlag(CVodeGetNumNonlinSolvIters(cvode_mem, &nni)))
        {
            PIHMexit(EXIT_FAILURE);
        }
        nsteps = (double)(nst - nst0);
        nfails = (double)(ncfn - ncfn0) / nsteps;
        niters = (double)(nni - nni0) / nsteps;
        if (nfails > ctrl->nncfn || niters >= ctrl->nnimax)
        {
            ctrl->maxstep /= ctrl->decr;
        }
        if (nfails == 0.0 && niters <= ctrl->nnimin)
        {
            ctrl->maxstep *= ctrl->incr;
        }
        ctrl->maxstep = (ctrl->maxstep < ctrl->stepsize) ? ctrl->maxstep : ctrl->stepsize;
        ctrl->maxstep = (ctrl->maxstep > ctrl->stmin) ? ctrl->maxstep : ctrl->stmin;
        if (!CheckCVodeFlag(CVodeSetMaxStep(cvode_mem, (realtype)ctrl->maxstep)))
        {
            PIHMexit(EXIT_FAILURE);
        }
        nst0 = nst;
        ncfn0 = ncfn;
        nni0 = nni;
    }
}
