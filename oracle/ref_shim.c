/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product.
 *
 * Compiled TOGETHER WITH the unmodified MM-PIHM sources where they lie under
 * /root/reference (see oracle/Makefile) into oracle/_ref/libpihm_ref.so
 * (-D_PIHM_) and oracle/_ref/libpihm_fbr_ref.so (-D_PIHM_ -D_FBR_).  It plays
 * the role of the reference's main.c (src/main.c:4-16 globals, :62-112 setup)
 * and exposes, over a plain C ABI for ctypes:
 *   - project loading through the reference's own ReadAlloc()/Initialize(),
 *   - construction of a pihm_struct straight from the column tables of
 *     include/pihm_b200.h (so a synthetic watershed needs no text files),
 *   - the packing stub pihm_struct -> column tables (the code a maintainer
 *     would add to the pihm driver),
 *   - the reference ODE() (src/ode.c:3), its flux fields, and the reference
 *     SetCVodeParam()/SolveCVode()/Summary() model-step sequence of
 *     src/pihm.c:3-134 minus printing.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * --impl reference legs may load these libraries.
 */
#include "pihm.h"
#include "cvode_spils.h"
#include "cvode_impl.h"      /* CVodeMem: the linear-solver hooks (cvode/src/cvode/cvode_impl.h:198-226) */
#include "pihm_b200.h"
#include "pihm_b200_sundials.h"

/* globals that src/main.c:4-16 defines (main.c itself is not compiled) */
int             verbose_mode;
int             debug_mode;
int             append_mode;
int             corr_mode;
int             spinup_mode;
int             fixed_length;
int             tecplot;
char            project[MAXSTRING];
int             nelem;
int             nriver;
#if defined(_OPENMP)
int             nthreads = 1;
#endif

typedef struct ref_handle
{
    pihm_struct     pihm;
    N_Vector        CV_Y;
    void           *cvode_mem;
    int             from_files;     /* forcing tables available */
    int             cvode_ready;
} ref_handle;

static ref_handle H;

int ref_is_fbr(void)
{
#if defined(_FBR_)
    return 1;
#else
    return 0;
#endif
}

int ref_set_threads(int n)
{
#if defined(_OPENMP)
    if (n > 0)
    {
        omp_set_num_threads(n);
    }
    nthreads = omp_get_max_threads();
    return nthreads;
#else
    (void)n;
    return 1;
#endif
}

int ref_sizeof_elem(void) { return (int)sizeof(elem_struct); }
int ref_sizeof_river(void) { return (int)sizeof(river_struct); }

/*
 * Load a project through the reference's own readers and Initialize().
 * rundir must contain input/<project>/... and input/vegprmt.tbl
 * (read_alloc.c:22-33 builds relative paths).
 */
int ref_open_project(const char *rundir, const char *proj, int verbose)
{
    if (chdir(rundir) != 0)
    {
        return -1;
    }
    memset(&H, 0, sizeof(H));
    verbose_mode = verbose ? VL_NORMAL : VL_SILENT;
    debug_mode = 0;
    strncpy(project, proj, MAXSTRING - 1);
#if defined(_OPENMP)
    nthreads = omp_get_max_threads();
#endif
    H.pihm = (pihm_struct)calloc(1, sizeof(*H.pihm));
    ReadAlloc(H.pihm);
    H.CV_Y = N_VNew(NumStateVar());
    Initialize(H.pihm, H.CV_Y, &H.cvode_mem);
    H.from_files = 1;
    return 0;
}

/*
 * Build a pihm_struct directly from the column tables (no files).  Only the
 * fields read by ODE()/Hydrol()/Summary() are filled; the rest stay zero.
 */
int ref_create_from_tables(const pihm_b200_mesh *m, double reltol,
    double abstol, double initstep)
{
    int             i, j;

#if defined(_FBR_)
    if (!m->fbr) return -2;
#else
    if (m->fbr) return -2;
#endif
    memset(&H, 0, sizeof(H));
    verbose_mode = VL_SILENT;
    nelem = m->nelem;
    nriver = m->nriver;
#if defined(_OPENMP)
    nthreads = omp_get_max_threads();
#endif
    H.pihm = (pihm_struct)calloc(1, sizeof(*H.pihm));
    H.pihm->elem = (elem_struct *)calloc(nelem, sizeof(elem_struct));
    H.pihm->river = (river_struct *)calloc(nriver > 0 ? nriver : 1,
        sizeof(river_struct));
    H.pihm->ctrl.surf_mode = m->surf_mode;
    H.pihm->ctrl.riv_mode = m->riv_mode;
    H.pihm->ctrl.stepsize = (int)m->stepsize;
    H.pihm->ctrl.etstep = (int)m->stepsize;
    H.pihm->ctrl.reltol = reltol;
    H.pihm->ctrl.abstol = abstol;
    H.pihm->ctrl.initstep = initstep;
    H.pihm->ctrl.maxstep = m->stepsize;

#define EF(c) (m->elem_f64[(size_t)(c) * nelem + i])
#define EI(c) (m->elem_i32[(size_t)(c) * nelem + i])
    for (i = 0; i < nelem; i++)
    {
        elem_struct    *e = &H.pihm->elem[i];

        e->ind = i + 1;
        e->topo.area = EF(PB_E_AREA);
        e->topo.zmin = EF(PB_E_ZMIN);
        e->topo.zmax = EF(PB_E_ZMAX);
#if defined(_FBR_)
        e->topo.zbed = EF(PB_E_ZBED);
#endif
        for (j = 0; j < NUM_EDGE; j++)
        {
            e->nabr[j] = EI(PB_EI_NABR0 + j);
            e->attrib.bc_type[j] = EI(PB_EI_BC0 + j);
#if defined(_FBR_)
            e->attrib.fbrbc_type[j] = EI(PB_EI_FBRBC0 + j);
#endif
            e->topo.edge[j] = EF(PB_E_EDGE0 + j);
            e->topo.nabrdist[j] = EF(PB_E_NABRDIST0 + j);
            e->topo.nabr_x[j] = EF(PB_E_NABRX0 + j);
            e->topo.nabr_y[j] = EF(PB_E_NABRY0 + j);
        }
        e->soil.depth = EF(PB_E_DEPTH);
        e->soil.ksath = EF(PB_E_KSATH);
        e->soil.ksatv = EF(PB_E_KSATV);
        e->soil.kinfv = EF(PB_E_KINFV);
        e->soil.dinf = EF(PB_E_DINF);
        e->soil.alpha = EF(PB_E_ALPHA);
        e->soil.beta = EF(PB_E_BETA);
        e->soil.porosity = EF(PB_E_POROSITY);
        e->soil.dmac = EF(PB_E_DMAC);
        e->soil.kmach = EF(PB_E_KMACH);
        e->soil.kmacv = EF(PB_E_KMACV);
        e->soil.areafv = EF(PB_E_AREAFV);
        e->soil.areafh = EF(PB_E_AREAFH);
        e->lc.rough = EF(PB_E_ROUGH);
        e->ps.rzd = EF(PB_E_RZD);
#if defined(_FBR_)
        e->geol.depth = EF(PB_E_GDEPTH);
        e->geol.ksath = EF(PB_E_GKSATH);
        e->geol.ksatv = EF(PB_E_GKSATV);
        e->geol.alpha = EF(PB_E_GALPHA);
        e->geol.beta = EF(PB_E_GBETA);
        e->geol.porosity = EF(PB_E_GPOROSITY);
#endif
        InitWFlux(&e->wf);
    }
#undef EF
#undef EI
#define RF(c) (m->riv_f64[(size_t)(c) * nriver + i])
#define RI(c) (m->riv_i32[(size_t)(c) * nriver + i])
    for (i = 0; i < nriver; i++)
    {
        river_struct   *r = &H.pihm->river[i];

        r->ind = i + 1;
        r->leftele = RI(PB_RI_LEFTELE);
        r->rightele = RI(PB_RI_RIGHTELE);
        r->down = RI(PB_RI_DOWN);
        r->attrib.riverbc_type = RI(PB_RI_BCTYPE);
        r->shp.intrpl_ord = RI(PB_RI_INTRPL_ORD);
        r->topo.area = RF(PB_R_AREA);
        r->topo.zmin = RF(PB_R_ZMIN);
        r->topo.zmax = RF(PB_R_ZMAX);
        r->topo.zbed = RF(PB_R_ZBED);
        r->topo.node_zmax = RF(PB_R_NODE_ZMAX);
        r->topo.dist_left = RF(PB_R_DIST_LEFT);
        r->topo.dist_right = RF(PB_R_DIST_RIGHT);
        r->shp.depth = RF(PB_R_SHP_DEPTH);
        r->shp.coeff = RF(PB_R_SHP_COEFF);
        r->shp.length = RF(PB_R_SHP_LENGTH);
        r->shp.width = RF(PB_R_SHP_WIDTH);
        r->matl.rough = RF(PB_R_ROUGH);
        r->matl.cwr = RF(PB_R_CWR);
        r->matl.ksath = RF(PB_R_KSATH);
        r->matl.ksatv = RF(PB_R_KSATV);
        r->matl.bedthick = RF(PB_R_BEDTHICK);
        r->matl.porosity = RF(PB_R_POROSITY);
        for (j = 0; j < NUM_RIVFLX; j++)
        {
            r->wf.rivflow[j] = 0.0;
        }
    }
#undef RF
#undef RI
    H.CV_Y = N_VNew(NumStateVar());
    H.cvode_mem = CVodeCreate(CV_BDF, CV_NEWTON);
    H.from_files = 0;
    return 0;
}

void ref_close(void)
{
    if (H.CV_Y) N_VDestroy(H.CV_Y);
    if (H.cvode_mem) CVodeFree(&H.cvode_mem);
    if (H.pihm)
    {
        if (!H.from_files)
        {
            free(H.pihm->elem);
            free(H.pihm->river);
        }
        /* files mode: FreeMem() (src/free_mem.c) also frees the print
         * structures that only MapOutput() sets up; the shim never calls
         * MapOutput, so the tables are simply left to the process exit */
        free(H.pihm);
    }
    memset(&H, 0, sizeof(H));
}

void ref_get_dims(int *ne, int *nr, int *nsv)
{
    *ne = nelem;
    *nr = nriver;
    *nsv = NumStateVar();
}

/* ctrl: [surf_mode, riv_mode, stepsize, etstep, starttime, endtime, nstep] */
void ref_get_ctrl(int *ictrl, double *dctrl)
{
    const ctrl_struct *c = &H.pihm->ctrl;

    ictrl[0] = c->surf_mode;
    ictrl[1] = c->riv_mode;
    ictrl[2] = c->stepsize;
    ictrl[3] = c->etstep;
    ictrl[4] = c->starttime;
    ictrl[5] = c->endtime;
    ictrl[6] = c->nstep;
    dctrl[0] = c->abstol;
    dctrl[1] = c->reltol;
    dctrl[2] = c->initstep;
    dctrl[3] = c->maxstep;
    dctrl[4] = c->stmin;
    dctrl[5] = c->nncfn;
    dctrl[6] = c->nnimax;
    dctrl[7] = c->nnimin;
    dctrl[8] = c->decr;
    dctrl[9] = c->incr;
}

/*
 * The packing stub: pihm->elem / pihm->river (AoS) -> column tables.
 * This is the code INTEGRATION.md asks the pihm driver to run once after
 * Initialize() (src/main.c:77).
 */
void ref_pack_tables(double *ef, int32_t *ei, double *rf, int32_t *ri)
{
    int             i, j;

#define EF(c) (ef[(size_t)(c) * nelem + i])
#define EI(c) (ei[(size_t)(c) * nelem + i])
    for (i = 0; i < nelem; i++)
    {
        const elem_struct *e = &H.pihm->elem[i];

        EF(PB_E_AREA) = e->topo.area;
        EF(PB_E_ZMIN) = e->topo.zmin;
        EF(PB_E_ZMAX) = e->topo.zmax;
#if defined(_FBR_)
        EF(PB_E_ZBED) = e->topo.zbed;
#else
        EF(PB_E_ZBED) = 0.0;
#endif
        for (j = 0; j < NUM_EDGE; j++)
        {
            EI(PB_EI_NABR0 + j) = e->nabr[j];
            EI(PB_EI_BC0 + j) = e->attrib.bc_type[j];
#if defined(_FBR_)
            EI(PB_EI_FBRBC0 + j) = e->attrib.fbrbc_type[j];
#else
            EI(PB_EI_FBRBC0 + j) = 0;
#endif
            EF(PB_E_EDGE0 + j) = e->topo.edge[j];
            EF(PB_E_NABRDIST0 + j) = e->topo.nabrdist[j];
            EF(PB_E_NABRX0 + j) = e->topo.nabr_x[j];
            EF(PB_E_NABRY0 + j) = e->topo.nabr_y[j];
        }
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
        EF(PB_E_ROUGH) = e->lc.rough;
        EF(PB_E_RZD) = e->ps.rzd;
#if defined(_FBR_)
        EF(PB_E_GDEPTH) = e->geol.depth;
        EF(PB_E_GKSATH) = e->geol.ksath;
        EF(PB_E_GKSATV) = e->geol.ksatv;
        EF(PB_E_GALPHA) = e->geol.alpha;
        EF(PB_E_GBETA) = e->geol.beta;
        EF(PB_E_GPOROSITY) = e->geol.porosity;
#else
        EF(PB_E_GDEPTH) = 0.0;
        EF(PB_E_GKSATH) = 0.0;
        EF(PB_E_GKSATV) = 0.0;
        EF(PB_E_GALPHA) = 0.0;
        EF(PB_E_GBETA) = 0.0;
        EF(PB_E_GPOROSITY) = 0.0;
#endif
    }
#undef EF
#undef EI
#define RF(c) (rf[(size_t)(c) * nriver + i])
#define RI(c) (ri[(size_t)(c) * nriver + i])
    for (i = 0; i < nriver; i++)
    {
        const river_struct *r = &H.pihm->river[i];

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
}

/* forcing table [PB_F_NCOL][nelem] + river bc [nriver]; get and set */
void ref_get_forcing(double *forc, double *rivbc)
{
    int             i, j;

    for (i = 0; i < nelem; i++)
    {
        const elem_struct *e = &H.pihm->elem[i];

        forc[(size_t)PB_F_PCPDRP * nelem + i] = e->wf.pcpdrp;
        forc[(size_t)PB_F_EDIR * nelem + i] = e->wf.edir;
        forc[(size_t)PB_F_ETT * nelem + i] = e->wf.ett;
        forc[(size_t)PB_F_WS0SURF * nelem + i] = e->ws0.surf;
        for (j = 0; j < NUM_EDGE; j++)
        {
            forc[(size_t)(PB_F_BC0 + j) * nelem + i] = e->bc.head[j];
#if defined(_FBR_)
            forc[(size_t)(PB_F_FBRBC0 + j) * nelem + i] = e->fbr_bc.head[j];
#else
            forc[(size_t)(PB_F_FBRBC0 + j) * nelem + i] = 0.0;
#endif
        }
    }
    for (i = 0; i < nriver; i++)
    {
        rivbc[i] = H.pihm->river[i].bc.head;
    }
}

void ref_set_forcing(const double *forc, const double *rivbc)
{
    int             i, j;

    for (i = 0; i < nelem; i++)
    {
        elem_struct    *e = &H.pihm->elem[i];

        e->wf.pcpdrp = forc[(size_t)PB_F_PCPDRP * nelem + i];
        e->wf.edir = forc[(size_t)PB_F_EDIR * nelem + i];
        e->wf.ett = forc[(size_t)PB_F_ETT * nelem + i];
        e->ws0.surf = forc[(size_t)PB_F_WS0SURF * nelem + i];
        for (j = 0; j < NUM_EDGE; j++)
        {
            e->bc.head[j] = forc[(size_t)(PB_F_BC0 + j) * nelem + i];
#if defined(_FBR_)
            e->fbr_bc.head[j] = forc[(size_t)(PB_F_FBRBC0 + j) * nelem + i];
#endif
        }
    }
    if (rivbc)
    {
        for (i = 0; i < nriver; i++)
        {
            H.pihm->river[i].bc.head = rivbc[i];
        }
    }
}

/* wf.ovlflow[3] table get/set (the H2 hidden state) */
void ref_get_ovlflow(double *ovl)
{
    int             i, j;

    for (i = 0; i < nelem; i++)
        for (j = 0; j < NUM_EDGE; j++)
            ovl[(size_t)j * nelem + i] = H.pihm->elem[i].wf.ovlflow[j];
}

void ref_set_ovlflow(const double *ovl)
{
    int             i, j;

    for (i = 0; i < nelem; i++)
        for (j = 0; j < NUM_EDGE; j++)
            H.pihm->elem[i].wf.ovlflow[j] = ovl[(size_t)j * nelem + i];
}

void ref_get_y(double *y)
{
    memcpy(y, NV_DATA(H.CV_Y), sizeof(double) * NumStateVar());
}

void ref_set_y(const double *y)
{
    memcpy(NV_DATA(H.CV_Y), y, sizeof(double) * NumStateVar());
}

/* the reference RHS, src/ode.c:3 */
int ref_ode(double t, const double *y, double *dy)
{
    N_Vector        vy, vdy;
    int             flag;

    vy = N_VNew(NumStateVar());
    vdy = N_VNew(NumStateVar());
    memcpy(NV_DATA(vy), y, sizeof(double) * NumStateVar());
    flag = ODE((realtype)t, vy, vdy, H.pihm);
    memcpy(dy, NV_DATA(vdy), sizeof(double) * NumStateVar());
    N_VDestroy(vy);
    N_VDestroy(vdy);
    return flag;
}

/* timing helper: n back-to-back reference RHS calls, returns seconds */
double ref_time_ode(int n)
{
    N_Vector        vdy = N_VNew(NumStateVar());
    double          t0, t1;
    int             k;
    struct timeval  tv;

    gettimeofday(&tv, NULL);
    t0 = tv.tv_sec + 1e-6 * tv.tv_usec;
    for (k = 0; k < n; k++)
    {
        ODE(0.0, H.CV_Y, vdy, H.pihm);
    }
    gettimeofday(&tv, NULL);
    t1 = tv.tv_sec + 1e-6 * tv.tv_usec;
    N_VDestroy(vdy);
    return t1 - t0;
}

/* flux fields of the last ODE() call */
void ref_get_fluxes(double *xf, double *rivflow)
{
    int             i, j;

#define XF(c) (xf[(size_t)(c) * nelem + i])
    if (xf)
    {
        for (i = 0; i < nelem; i++)
        {
            const wflux_struct *wf = &H.pihm->elem[i].wf;

            for (j = 0; j < NUM_EDGE; j++)
            {
                XF(PB_X_OVL0 + j) = wf->ovlflow[j];
                XF(PB_X_SUB0 + j) = wf->subsurf[j];
#if defined(_FBR_)
                XF(PB_X_FBRFLOW0 + j) = wf->fbrflow[j];
#else
                XF(PB_X_FBRFLOW0 + j) = 0.0;
#endif
            }
            XF(PB_X_INFIL) = wf->infil;
            XF(PB_X_RECHG) = wf->rechg;
            XF(PB_X_EDIR_SURF) = wf->edir_surf;
            XF(PB_X_EDIR_UNSAT) = wf->edir_unsat;
            XF(PB_X_EDIR_GW) = wf->edir_gw;
            XF(PB_X_ETT_UNSAT) = wf->ett_unsat;
            XF(PB_X_ETT_GW) = wf->ett_gw;
#if defined(_FBR_)
            XF(PB_X_FBR_INFIL) = wf->fbr_infil;
            XF(PB_X_FBR_RECHG) = wf->fbr_rechg;
#else
            XF(PB_X_FBR_INFIL) = 0.0;
            XF(PB_X_FBR_RECHG) = 0.0;
#endif
        }
    }
#undef XF
    if (rivflow)
    {
        for (i = 0; i < nriver; i++)
            for (j = 0; j < NUM_RIVFLX; j++)
                rivflow[(size_t)j * nriver + i] =
                    H.pihm->river[i].wf.rivflow[j];
    }
}

/*
 * Integrator side: reference SetCVodeParam (src/ode.c:340).  SetCVodeParam
 * keeps a static `reset` flag, so it may be called once per process for a
 * fresh cvode_mem and afterwards re-inits (CVodeReInit) -- same as a spin-up.
 */
void ref_set_cvode_param(void)
{
    SetCVodeParam(H.pihm, H.cvode_mem, H.CV_Y);
    H.cvode_ready = 1;
}

/* the AdjCVodeMaxStep controller parameters of ctrl_struct (read_para.c); tables mode leaves them 0 */
void ref_set_maxstep_ctrl(double stmin, double nncfn, double nnimax,
    double nnimin, double decr, double incr)
{
    H.pihm->ctrl.stmin = stmin;
    H.pihm->ctrl.nncfn = nncfn;
    H.pihm->ctrl.nnimax = nnimax;
    H.pihm->ctrl.nnimin = nnimin;
    H.pihm->ctrl.decr = decr;
    H.pihm->ctrl.incr = incr;
}

void ref_set_max_step(double hmax)
{
    H.pihm->ctrl.maxstep = hmax;
    CVodeSetMaxStep(H.cvode_mem, (realtype)hmax);
}

/*
 * One model step of src/pihm.c:3-134 without the print calls.
 * from_files: ApplyBc/ApplyForc/IntcpSnowEt run as in the reference.
 * tables: forcing was injected with ref_set_forcing().
 * skip_forcing: the caller already ran ref_apply_forcing(cstep).
 * Returns the model time after the step (ctime seconds).
 */
int ref_model_step(int cstep, int adj_max_step, int skip_forcing)
{
    pihm_struct     pihm = H.pihm;
    int             t;

    if (H.from_files)
    {
        pihm->ctrl.cstep = cstep;
        t = pihm->ctrl.tout[cstep];
        if (!skip_forcing)
        {
            ApplyBc(&pihm->forc, pihm->elem, pihm->river, t);
            if ((t - pihm->ctrl.starttime) % pihm->ctrl.etstep == 0)
            {
                ApplyForc(&pihm->forc, pihm->elem, t);
                IntcpSnowEt(t, (double)pihm->ctrl.etstep, pihm->elem,
                    &pihm->cal);
            }
        }
        SolveCVode(pihm->ctrl.starttime, &t, pihm->ctrl.tout[cstep + 1], 0.0,
            H.cvode_mem, H.CV_Y);
    }
    else
    {
        /* starttime = 0, tout[k] = k * stepsize */
        t = cstep * pihm->ctrl.stepsize;
        SolveCVode(0, &t, (cstep + 1) * pihm->ctrl.stepsize, 0.0,
            H.cvode_mem, H.CV_Y);
    }
    Summary(pihm->elem, pihm->river, H.CV_Y, (double)pihm->ctrl.stepsize);
    if (adj_max_step)
    {
        AdjCVodeMaxStep(H.cvode_mem, &pihm->ctrl);
    }
    return t;
}

/* only the forcing part of a model step (files mode): lets a test pull the
 * reference's pcpdrp/edir/ett for time tout[cstep] and push them to the GPU */
void ref_apply_forcing(int cstep)
{
    pihm_struct     pihm = H.pihm;
    int             t;

    if (!H.from_files) return;
    pihm->ctrl.cstep = cstep;
    t = pihm->ctrl.tout[cstep];
    ApplyBc(&pihm->forc, pihm->elem, pihm->river, t);
    if ((t - pihm->ctrl.starttime) % pihm->ctrl.etstep == 0)
    {
        ApplyForc(&pihm->forc, pihm->elem, t);
        IntcpSnowEt(t, (double)pihm->ctrl.etstep, pihm->elem, &pihm->cal);
    }
}

/* counters: nst nfe nni ncfn netf nli ncfl nfeLS njtimes nor nsetups qlast qcur */
void ref_get_stats(long int *s, double *d)
{
    int             q;

    CVodeGetNumSteps(H.cvode_mem, &s[0]);
    CVodeGetNumRhsEvals(H.cvode_mem, &s[1]);
    CVodeGetNumNonlinSolvIters(H.cvode_mem, &s[2]);
    CVodeGetNumNonlinSolvConvFails(H.cvode_mem, &s[3]);
    CVodeGetNumErrTestFails(H.cvode_mem, &s[4]);
    CVSpilsGetNumLinIters(H.cvode_mem, &s[5]);
    CVSpilsGetNumConvFails(H.cvode_mem, &s[6]);
    CVSpilsGetNumRhsEvals(H.cvode_mem, &s[7]);
    CVSpilsGetNumJtimesEvals(H.cvode_mem, &s[8]);
    CVodeGetNumStabLimOrderReds(H.cvode_mem, &s[9]);
    CVodeGetNumLinSolvSetups(H.cvode_mem, &s[10]);
    CVodeGetLastOrder(H.cvode_mem, &q);
    s[11] = q;
    CVodeGetCurrentOrder(H.cvode_mem, &q);
    s[12] = q;
    CVodeGetLastStep(H.cvode_mem, &d[0]);
    CVodeGetCurrentStep(H.cvode_mem, &d[1]);
    CVodeGetCurrentTime(H.cvode_mem, &d[2]);
}

/* Summary() (src/update.c:3-98) on caller data: y plays CV_Y after SolveCVode,
 * the wf.* fields are those the last ODE() call left behind.  Afterwards
 * ref_get_fluxes() shows the mass-balance wf.infil / wf.fbr_infil. */
void ref_summary(const double *y)
{
    ref_set_y(y);
    Summary(H.pihm->elem, H.pihm->river, H.CV_Y,
        (double)H.pihm->ctrl.stepsize);
}

/* ws0 of elements and rivers in the block layout of y */
void ref_get_ws0(double *y)
{
    int             i;

    for (i = 0; i < nelem; i++)
    {
        y[SURF(i)] = H.pihm->elem[i].ws0.surf;
        y[UNSAT(i)] = H.pihm->elem[i].ws0.unsat;
        y[GW(i)] = H.pihm->elem[i].ws0.gw;
#if defined(_FBR_)
        y[FBRUNSAT(i)] = H.pihm->elem[i].ws0.fbr_unsat;
        y[FBRGW(i)] = H.pihm->elem[i].ws0.fbr_gw;
#endif
    }
    for (i = 0; i < nriver; i++)
    {
        y[RIVSTG(i)] = H.pihm->river[i].ws0.stage;
        y[RIVGW(i)] = H.pihm->river[i].ws0.gw;
    }
}

/*
 * Interception / snow / ET, files mode only (the land-cover
 * and vegetation fields come from Initialize()).
 */
/* dims[4] = {nmeteo, nlai, 40 (rows of the monthly tables, forcing.c:359), etstep} */
int ref_et_dims(int *dims)
{
    if (!H.from_files) return -1;
    dims[0] = H.pihm->forc.nmeteo;
    dims[1] = H.pihm->forc.nlai;
    dims[2] = 40;
    dims[3] = H.pihm->ctrl.etstep;
    return 0;
}

/* the static columns IntcpSnowEt reads: [PB_ET_NCOL][nelem], [PB_ETI_NCOL][nelem] */
void ref_pack_et_tables(double *etf, int32_t *eti)
{
    int             i;

#define TF(c) (etf[(size_t)(c) * nelem + i])
#define TI(c) (eti[(size_t)(c) * nelem + i])
    for (i = 0; i < nelem; i++)
    {
        const elem_struct *e = &H.pihm->elem[i];

        TF(PB_ET_ALBEDOMIN) = e->lc.albedomin;
        TF(PB_ET_ALBEDOMAX) = e->lc.albedomax;
        TF(PB_ET_CMCFACTR) = e->lc.cmcfactr;
        TF(PB_ET_SHDFAC) = e->lc.shdfac;
        TF(PB_ET_CFACTR) = e->lc.cfactr;
        TF(PB_ET_RGL) = e->epc.rgl;
        TF(PB_ET_RSMIN) = e->epc.rsmin;
        TF(PB_ET_RSMAX) = e->epc.rsmax;
        TF(PB_ET_TOPT) = e->epc.topt;
        TF(PB_ET_SMCMIN) = e->soil.smcmin;
        TF(PB_ET_SMCWLT) = e->soil.smcwlt;
        TF(PB_ET_SMCREF) = e->soil.smcref;
        TF(PB_ET_ZLVL_WIND) = e->ps.zlvl_wind;
        TI(PB_ETI_METEO_TYPE) = e->attrib.meteo_type;
        TI(PB_ETI_LAI_TYPE) = e->attrib.lai_type;
        TI(PB_ETI_LC_TYPE) = e->attrib.lc_type;
    }
#undef TF
#undef TI
}

/* overwrite attrib.lai_type / lc_type (test variety: input/example uses one LAI series and one
 * land-cover class, so the monthly-table branches of forcing.c:249-257 / is_sm_et.c:58-65,92 would
 * stay unvisited); eti as in ref_pack_et_tables, meteo_type is left alone */
void ref_et_set_types(const int32_t *eti)
{
    int             i;

    for (i = 0; i < nelem; i++)
    {
        H.pihm->elem[i].attrib.lai_type = eti[(size_t)PB_ETI_LAI_TYPE * nelem + i];
        H.pihm->elem[i].attrib.lc_type = eti[(size_t)PB_ETI_LC_TYPE * nelem + i];
    }
}

/* month-of-year lookups of the reference at model time t (forcing.c:351-618) */
void ref_et_monthly(int t, int nlc, double *lai_lc, double *z0_lc,
    double *meltf)
{
    int             k;

    for (k = 0; k < nlc; k++)
    {
        lai_lc[k] = MonthlyLai(t, k + 1);
        z0_lc[k] = MonthlyRl(t, k + 1);
    }
    *meltf = MonthlyMf(t);
}

/* forc->meteo[k].value[] / forc->lai[k].value[0] as the last ApplyForc left them */
void ref_et_get_forc(double *meteo, double *lai)
{
    int             k, j;

    for (k = 0; k < H.pihm->forc.nmeteo; k++)
        for (j = 0; j < NUM_METEO_VAR; j++)
            meteo[k * NUM_METEO_VAR + j] = H.pihm->forc.meteo[k].value[j];
    for (k = 0; k < H.pihm->forc.nlai; k++)
        lai[k] = H.pihm->forc.lai[k].value[0];
}

void ref_et_get_cal(double *cal3)
{
    cal3[0] = H.pihm->cal.edir;
    cal3[1] = H.pihm->cal.ec;
    cal3[2] = H.pihm->cal.ett;
}

/* out[PB_EO_NCOL][nelem] */
void ref_et_get(double *out)
{
    int             i;

    for (i = 0; i < nelem; i++)
    {
        const elem_struct *e = &H.pihm->elem[i];

        out[(size_t)PB_EO_PCPDRP * nelem + i] = e->wf.pcpdrp;
        out[(size_t)PB_EO_EDIR * nelem + i] = e->wf.edir;
        out[(size_t)PB_EO_ETT * nelem + i] = e->wf.ett;
        out[(size_t)PB_EO_EC * nelem + i] = e->wf.ec;
        out[(size_t)PB_EO_DRIP * nelem + i] = e->wf.drip;
        out[(size_t)PB_EO_SNEQV * nelem + i] = e->ws.sneqv;
        out[(size_t)PB_EO_CMC * nelem + i] = e->ws.cmc;
    }
}

void ref_et_set_state(const double *sneqv, const double *cmc)
{
    int             i;

    for (i = 0; i < nelem; i++)
    {
        H.pihm->elem[i].ws.sneqv = sneqv[i];
        H.pihm->elem[i].ws.cmc = cmc[i];
    }
}

/*
 * The reference's IntcpSnowEt (src/is_sm_et.c:4) on caller data: station
 * values meteo[nmeteo][NUM_METEO_VAR] and lai[nlai] take the place of
 * IntrplForc's results, the per-element assignments of ApplyMeteoForc /
 * ApplyLai (forcing.c:134-160, 242-258) are made as there, y supplies
 * elem.ws.unsat / ws.gw as Summary() would have left them.
 */
int ref_et_run(int t, double stepsize, const double *meteo, const double *lai,
    const double *y)
{
    int             i, k, j;
    pihm_struct     pihm = H.pihm;

    if (!H.from_files) return -1;
    for (k = 0; k < pihm->forc.nmeteo; k++)
        for (j = 0; j < NUM_METEO_VAR; j++)
            pihm->forc.meteo[k].value[j] = meteo[k * NUM_METEO_VAR + j];
    for (k = 0; k < pihm->forc.nlai; k++)
        pihm->forc.lai[k].value[0] = lai[k];
    for (i = 0; i < nelem; i++)
    {
        elem_struct    *e = &pihm->elem[i];
        int             ind = e->attrib.meteo_type - 1;

        e->wf.prcp = pihm->forc.meteo[ind].value[PRCP_TS] / 1000.0;
        e->es.sfctmp = pihm->forc.meteo[ind].value[SFCTMP_TS];
        e->ps.rh = pihm->forc.meteo[ind].value[RH_TS];
        e->ps.sfcspd = pihm->forc.meteo[ind].value[SFCSPD_TS];
        e->ef.soldn = pihm->forc.meteo[ind].value[SOLAR_TS];
        e->ef.soldn = (e->ef.soldn > 0.0) ? e->ef.soldn : 0.0;
        if (e->attrib.lai_type > 0)
        {
            e->ps.proj_lai = pihm->forc.lai[e->attrib.lai_type - 1].value[0];
        }
        else
        {
            e->ps.proj_lai = MonthlyLai(t, e->attrib.lc_type);
        }
        e->ws.unsat = y[UNSAT(i)];
        e->ws.gw = y[GW(i)];
    }
    IntcpSnowEt(t, stepsize, pihm->elem, &pihm->cal);
    return 0;
}

/* model time of step k (ctrl.tout[k]) */
int ref_tout(int cstep)
{
    return H.from_files ? H.pihm->ctrl.tout[cstep] : 0;
}

/*
 * Print accumulation: a varctrl_struct over one field, fed to
 * the reference's own UpdPrintVar() and PrintData() (src/print.c:171-251); the
 * record PrintData writes goes to a temporary file and is read back.
 * src / col as in enum pihm_b200_print_src.
 */
#define MAX_PV 64
static varctrl_struct PV[MAX_PV];
static int      npv;

static const double *pv_field(int src, int col, int j)
{
    const elem_struct *e = &H.pihm->elem[(src == PB_PS_RIV_FLUX ||
        (src == PB_PS_STATE && (col == 3 || col == 4))) ? 0 : j];
    const river_struct *r = &H.pihm->river[(nriver > 0 && j < nriver) ? j : 0];

    switch (src)
    {
        case PB_PS_STATE:
            switch (col)
            {
                case 0: return &e->ws.surf;
                case 1: return &e->ws.unsat;
                case 2: return &e->ws.gw;
                case 3: return &r->ws.stage;
                case 4: return &r->ws.gw;
#if defined(_FBR_)
                case 5: return &e->ws.fbr_unsat;
                case 6: return &e->ws.fbr_gw;
#endif
            }
            return NULL;
        case PB_PS_ELEM_FLUX:
            if (col >= PB_X_OVL0 && col <= PB_X_OVL2) return &e->wf.ovlflow[col - PB_X_OVL0];
            if (col >= PB_X_SUB0 && col <= PB_X_SUB2) return &e->wf.subsurf[col - PB_X_SUB0];
            switch (col)
            {
                case PB_X_INFIL: return &e->wf.infil;
                case PB_X_RECHG: return &e->wf.rechg;
                case PB_X_EDIR_SURF: return &e->wf.edir_surf;
                case PB_X_EDIR_UNSAT: return &e->wf.edir_unsat;
                case PB_X_EDIR_GW: return &e->wf.edir_gw;
                case PB_X_ETT_UNSAT: return &e->wf.ett_unsat;
                case PB_X_ETT_GW: return &e->wf.ett_gw;
#if defined(_FBR_)
                case PB_X_FBR_INFIL: return &e->wf.fbr_infil;
                case PB_X_FBR_RECHG: return &e->wf.fbr_rechg;
#endif
            }
#if defined(_FBR_)
            if (col >= PB_X_FBRFLOW0 && col <= PB_X_FBRFLOW2) return &e->wf.fbrflow[col - PB_X_FBRFLOW0];
#endif
            return NULL;
        case PB_PS_RIV_FLUX:
            return (col >= 0 && col < NUM_RIVFLX) ? &r->wf.rivflow[col] : NULL;
        case PB_PS_ET:
            switch (col)
            {
                case PB_EO_PCPDRP: return &e->wf.pcpdrp;
                case PB_EO_EDIR: return &e->wf.edir;
                case PB_EO_ETT: return &e->wf.ett;
                case PB_EO_EC: return &e->wf.ec;
                case PB_EO_DRIP: return &e->wf.drip;
                case PB_EO_SNEQV: return &e->ws.sneqv;
                case PB_EO_CMC: return &e->ws.cmc;
            }
            return NULL;
    }
    return NULL;
}

int ref_print_add(int src, int col, int upd_intvl, int intvl)
{
    varctrl_struct *v;
    int             j, n;
    int             river = (src == PB_PS_RIV_FLUX) ||
        (src == PB_PS_STATE && (col == 3 || col == 4));

    if (npv >= MAX_PV || pv_field(src, col, 0) == NULL) return -1;
    v = &PV[npv];
    memset(v, 0, sizeof(*v));
    n = river ? nriver : nelem;
    v->nvar = n;
    v->intvl = intvl;
    v->upd_intvl = upd_intvl;
    v->var = (const double **)malloc(sizeof(double *) * (n > 0 ? n : 1));
    v->buffer = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    for (j = 0; j < n; j++) v->var[j] = pv_field(src, col, j);
    v->datfile = tmpfile();
    return npv++;
}

/* UpdPrintVar(varctrl, nprint, module_step) over all variables made so far */
void ref_print_update(int module_step)
{
    UpdPrintVar(PV, npv, module_step);
}

/* PrintData(varctrl, nprint, t, lapse, ascii = 0) for all variables; then the
 * last record of variable id is read back: returns 1 and fills out[nvar] when
 * a record was written by this call, 0 when PrintNow() said no */
int ref_print_data(int id, int t, int lapse, double *out)
{
    varctrl_struct *v = &PV[id];
    long            before, after;
    double          tt;

    before = ftell(v->datfile);
    PrintData(v, 1, t, lapse, 0);
    after = ftell(v->datfile);
    if (after == before) return 0;
    fseek(v->datfile, before, SEEK_SET);
    if (fread(&tt, sizeof(double), 1, v->datfile) != 1) return -1;
    if (fread(out, sizeof(double), v->nvar, v->datfile) != (size_t)v->nvar) return -1;
    fseek(v->datfile, 0, SEEK_END);
    return 1;
}

void ref_print_reset(void)
{
    int             i;

    for (i = 0; i < npv; i++)
    {
        free((void *)PV[i].var);
        free(PV[i].buffer);
        if (PV[i].datfile) fclose(PV[i].datfile);
    }
    npv = 0;
}

/* element/river water states after Summary() (ws), for trajectory checks */
void ref_get_ws(double *y)
{
    int             i;

    for (i = 0; i < nelem; i++)
    {
        y[SURF(i)] = H.pihm->elem[i].ws.surf;
        y[UNSAT(i)] = H.pihm->elem[i].ws.unsat;
        y[GW(i)] = H.pihm->elem[i].ws.gw;
#if defined(_FBR_)
        y[FBRUNSAT(i)] = H.pihm->elem[i].ws.fbr_unsat;
        y[FBRGW(i)] = H.pihm->elem[i].ws.fbr_gw;
#endif
    }
    for (i = 0; i < nriver; i++)
    {
        y[RIVSTG(i)] = H.pihm->river[i].ws.stage;
        y[RIVGW(i)] = H.pihm->river[i].ws.gw;
    }
}

/* initial condition of a tables-mode handle: y -> CV_Y and ws/ws0 (InitVar,
 * src/initialize.c:569-617) */
void ref_init_state(const double *y)
{
    int             i;

    ref_set_y(y);
    for (i = 0; i < nelem; i++)
    {
        elem_struct    *e = &H.pihm->elem[i];

        e->ws.surf = y[SURF(i)];
        e->ws.unsat = y[UNSAT(i)];
        e->ws.gw = y[GW(i)];
#if defined(_FBR_)
        e->ws.fbr_unsat = y[FBRUNSAT(i)];
        e->ws.fbr_gw = y[FBRGW(i)];
#endif
        e->ws0 = e->ws;
    }
    for (i = 0; i < nriver; i++)
    {
        H.pihm->river[i].ws.stage = y[RIVSTG(i)];
        H.pihm->river[i].ws.gw = y[RIVGW(i)];
        H.pihm->river[i].ws0 = H.pihm->river[i].ws;
    }
}

/*
 * Reference N_Vector kernels for the vector-op parity tests: call the serial
 * implementation (cvode/src/nvec_ser/nvector_serial.c) on caller buffers.
 * op: 0 linearsum 1 const 2 prod 3 div 4 scale 5 abs 6 inv 7 addconst
 *     8 dot 9 maxnorm 10 wrmsnorm 11 min
 */
double ref_nvec_op(int op, long int n, double a, const double *x, double b,
    const double *y, double *z)
{
    N_Vector        vx, vy, vz;
    double          r = 0.0;

    vx = N_VMake_Serial(n, (realtype *)x);
    vy = N_VMake_Serial(n, (realtype *)y);
    vz = N_VMake_Serial(n, z);
    switch (op)
    {
        case 0: N_VLinearSum(a, vx, b, vy, vz); break;
        case 1: N_VConst(a, vz); break;
        case 2: N_VProd(vx, vy, vz); break;
        case 3: N_VDiv(vx, vy, vz); break;
        case 4: N_VScale(a, vx, vz); break;
        case 5: N_VAbs(vx, vz); break;
        case 6: N_VInv(vx, vz); break;
        case 7: N_VAddConst(vx, b, vz); break;
        case 8: r = N_VDotProd(vx, vy); break;
        case 9: r = N_VMaxNorm(vx); break;
        case 10: r = N_VWrmsNorm(vx, vy); break;
        case 11: r = N_VMin(vx); break;
        default: r = -1.0;
    }
    N_VDestroy(vx);
    N_VDestroy(vy);
    N_VDestroy(vz);
    return r;
}

/*
 * The reference CVODE (this library) driven with an EXTERNAL N_Vector and RHS
 * callback -- the drop-in of INTEGRATION.md: same call sequence as
 * SetCVodeParam()/SolveCVode() (src/ode.c:340-498), with N_VNew_PihmB200's
 * vector and PihmB200_ODE in place of N_VNew_Serial / ODE.
 */
static void    *X_mem;
static N_Vector X_y;

int ref_ext_cvode_init(void *nv_y, void *rhs_fn, void *user_data, double reltol,
    double abstol, double initstep, double maxstep, long int mxsteps)
{
    int             flag;

    X_y = (N_Vector)nv_y;
    X_mem = CVodeCreate(CV_BDF, CV_NEWTON);
    if (X_mem == NULL) return -1;
    flag = CVodeInit(X_mem, (CVRhsFn)rhs_fn, 0.0, X_y);
    if (flag < 0) return flag;
    flag = CVodeSStolerances(X_mem, (realtype)reltol, (realtype)abstol);
    if (flag < 0) return flag;
    CVodeSetUserData(X_mem, user_data);
    CVodeSetInitStep(X_mem, (realtype)initstep);
    CVodeSetStabLimDet(X_mem, TRUE);
    CVodeSetMaxStep(X_mem, (realtype)maxstep);
    CVodeSetMaxNumSteps(X_mem, mxsteps);
    return CVSpgmr(X_mem, PREC_NONE, 0);
}

/*
 * Linear-solver plug-in: attach a device SPGMR to the reference
 * CVODE the way CVSpgmr() attaches its own (cvode_spgmr.c:135-138,174,220).
 * This is the binding a maintainer adds (INTEGRATION.md 3a'); `solve` is
 * pihm_b200_spgmr_solve, `engine` a pihm_b200_cvode from pihm_b200_cvode_create.
 */
typedef int     (*b200_lsolve_fn)(void *engine, double tn, double gamma,
    double tq4, int mnewt, void *b, void *weight, void *ycur, void *fcur);
static b200_lsolve_fn X_lsolve;
static void    *X_engine;

#define DEVVEC(v) ((void *)((N_VectorContent_PihmB200)((v)->content))->dev)

static int x_linit(CVodeMem cv_mem)
{
    (void)cv_mem;
    return 0;
}

static int x_lsolve(CVodeMem cv_mem, N_Vector b, N_Vector weight, N_Vector ycur,
    N_Vector fcur)
{
    return X_lsolve(X_engine, cv_mem->cv_tn, cv_mem->cv_gamma,
        cv_mem->cv_tq[4], cv_mem->cv_mnewt, DEVVEC(b), DEVVEC(weight),
        DEVVEC(ycur), DEVVEC(fcur));
}

static int x_lfree(CVodeMem cv_mem)
{
    cv_mem->cv_lmem = NULL;
    return 0;
}

int ref_ext_cvode_attach_lsolve(void *solve, void *engine)
{
    CVodeMem        cv_mem = (CVodeMem)X_mem;

    if (cv_mem == NULL || solve == NULL) return -1;
    if (cv_mem->cv_lfree != NULL) cv_mem->cv_lfree(cv_mem);   /* drop CVSPGMR */
    X_lsolve = (b200_lsolve_fn)solve;
    X_engine = engine;
    cv_mem->cv_linit = x_linit;
    cv_mem->cv_lsetup = NULL;
    cv_mem->cv_lsolve = x_lsolve;
    cv_mem->cv_lfree = x_lfree;
    cv_mem->cv_lmem = engine;
    cv_mem->cv_setupNonNull = FALSE;
    return 0;
}

/* nst nfe nni ncfn netf of the stepper only (the plug-in keeps its own nli, ncfl, nfeLS) */
void ref_ext_cvode_stepper_stats(long int *s)
{
    CVodeGetNumSteps(X_mem, &s[0]);
    CVodeGetNumRhsEvals(X_mem, &s[1]);
    CVodeGetNumNonlinSolvIters(X_mem, &s[2]);
    CVodeGetNumNonlinSolvConvFails(X_mem, &s[3]);
    CVodeGetNumErrTestFails(X_mem, &s[4]);
}

int ref_ext_cvode_solve(double tout, double *tret)
{
    realtype        solvert;
    int             flag;

    flag = CVodeSetStopTime(X_mem, (realtype)tout);
    if (flag < 0) return flag;
    flag = CVode(X_mem, (realtype)tout, X_y, &solvert, CV_NORMAL);
    *tret = solvert;
    return flag;
}

void ref_ext_cvode_stats(long int *s)
{
    CVodeGetNumSteps(X_mem, &s[0]);
    CVodeGetNumRhsEvals(X_mem, &s[1]);
    CVodeGetNumNonlinSolvIters(X_mem, &s[2]);
    CVodeGetNumNonlinSolvConvFails(X_mem, &s[3]);
    CVodeGetNumErrTestFails(X_mem, &s[4]);
    CVSpilsGetNumLinIters(X_mem, &s[5]);
    CVSpilsGetNumConvFails(X_mem, &s[6]);
    CVSpilsGetNumRhsEvals(X_mem, &s[7]);
}

void ref_ext_cvode_free(void)
{
    if (X_mem) CVodeFree(&X_mem);
    X_mem = NULL;
}
