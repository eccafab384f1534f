/*
 * pihm_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product.
 *
 * Plain-C, CPU restatement of the MM-PIHM RHS (ODE() and everything under
 * Hydrol()) on the column tables of include/pihm_b200.h, plus the serial
 * N_Vector arithmetic.  Every function cites the reference lines it follows.
 * Operation order is kept identical to the reference so that, compiled without
 * FMA contraction, it reproduces the reference bit for bit; that claim is
 * pinned by tests/test_oracle_vs_reference.py (against oracle/_ref, i.e. the
 * real reference compiled from /root/reference) and by the golden vectors in
 * tests/golden/ (generated from the reference by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  The product (libpihm_b200.so) never links or calls it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "pihm_b200.h"

/* src/include/pihm_const.h:7,77-82,158-159,143-146 */
#define GRAV        9.80665
#define PSIMIN      -70.0
#define DEPRSTG     1.0E-4
#define GRADMIN     5.0E-8
#define SATMIN      0.1
#define RIVDPTHMIN  0.05
#define RIVGRADMIN  0.05
#define KINEMATIC   1
#define DIFF_WAVE   2
#define RECTANGLE   1
#define TRIANGLE    2
#define QUADRATIC   3
#define CUBIC       4
#define BC_DIRICHLET      -1
#define BC_NEUMANN        -2
#define BC_ZERO_DPTH_GRAD -3
#define BC_CRIT_DPTH      -4
/* rivflow slots, pihm_const.h:130-140 */
enum { UP_C2C = 0, DOWN_C2C, LEFT_S2C, RIGHT_S2C, LEFT_A2C, RIGHT_A2C,
       CHANL_LKG, LEFT_A2A, RIGHT_A2A, DOWN_A2A, UP_A2A };

typedef struct oracle_model
{
    int             ne, nr, fbr, surf_mode, riv_mode;
    double          dt;
    double         *ef;     /* [PB_E_NCOL][ne]  */
    int32_t        *ei;     /* [PB_EI_NCOL][ne] */
    double         *rf;     /* [PB_R_NCOL][nr]  */
    int32_t        *ri;     /* [PB_RI_NCOL][nr] */
    double         *forc;   /* [PB_F_NCOL][ne]  */
    double         *rivbc;  /* [nr] */
    /* working state = the ws/wf fields of the reference structs */
    double         *surf, *unsat, *gw, *fbr_unsat, *fbr_gw, *surfh;
    double         *stage, *rgw;
    double         *xf;     /* [PB_X_NCOL][ne]; OVL*, SUB* persist across calls */
    double         *rivflow;/* [11][nr] */
    double         *dhbydx, *dhbydy;
    double         *ws0;    /* [nsv] elem/river ws0 (block layout of y)  */
    double         *subrunoff; /* [ne] MassBalance's local, update.c:131 */
} oracle_model;

#define E(c, i)   (om->ef[(size_t)(c) * om->ne + (i)])
#define EI(c, i)  (om->ei[(size_t)(c) * om->ne + (i)])
#define R(c, i)   (om->rf[(size_t)(c) * om->nr + (i)])
#define RI(c, i)  (om->ri[(size_t)(c) * om->nr + (i)])
#define F(c, i)   (om->forc[(size_t)(c) * om->ne + (i)])
#define X(c, i)   (om->xf[(size_t)(c) * om->ne + (i)])
#define RFLX(k, i) (om->rivflow[(size_t)(k) * om->nr + (i)])

static void *dup_mem(const void *p, size_t n)
{
    void *q = malloc(n ? n : 1);
    if (p && n) memcpy(q, p, n);
    return q;
}

oracle_model *oracle_create(const pihm_b200_mesh *m)
{
    oracle_model   *om = (oracle_model *)calloc(1, sizeof(*om));
    size_t          ne = (size_t)m->nelem, nr = (size_t)m->nriver;

    om->ne = m->nelem; om->nr = m->nriver; om->fbr = m->fbr;
    om->surf_mode = m->surf_mode; om->riv_mode = m->riv_mode;
    om->dt = m->stepsize;
    om->ef = (double *)dup_mem(m->elem_f64, sizeof(double) * PB_E_NCOL * ne);
    om->ei = (int32_t *)dup_mem(m->elem_i32, sizeof(int32_t) * PB_EI_NCOL * ne);
    om->rf = (double *)dup_mem(m->riv_f64, sizeof(double) * PB_R_NCOL * nr);
    om->ri = (int32_t *)dup_mem(m->riv_i32, sizeof(int32_t) * PB_RI_NCOL * nr);
    om->forc = (double *)calloc(PB_F_NCOL * ne + 1, sizeof(double));
    om->rivbc = (double *)calloc(nr + 1, sizeof(double));
    om->surf = (double *)calloc(ne + 1, sizeof(double));
    om->unsat = (double *)calloc(ne + 1, sizeof(double));
    om->gw = (double *)calloc(ne + 1, sizeof(double));
    om->fbr_unsat = (double *)calloc(ne + 1, sizeof(double));
    om->fbr_gw = (double *)calloc(ne + 1, sizeof(double));
    om->surfh = (double *)calloc(ne + 1, sizeof(double));
    om->stage = (double *)calloc(nr + 1, sizeof(double));
    om->rgw = (double *)calloc(nr + 1, sizeof(double));
    om->xf = (double *)calloc(PB_X_NCOL * ne + 1, sizeof(double));
    om->rivflow = (double *)calloc(PIHM_B200_NUM_RIVFLX * nr + 1, sizeof(double));
    om->dhbydx = (double *)calloc(ne + 1, sizeof(double));
    om->dhbydy = (double *)calloc(ne + 1, sizeof(double));
    om->ws0 = (double *)calloc(5 * ne + 2 * nr + 1, sizeof(double));
    om->subrunoff = (double *)calloc(ne + 1, sizeof(double));
    return om;
}

void oracle_destroy(oracle_model *om)
{
    if (!om) return;
    free(om->ef); free(om->ei); free(om->rf); free(om->ri); free(om->forc);
    free(om->rivbc); free(om->surf); free(om->unsat); free(om->gw);
    free(om->fbr_unsat); free(om->fbr_gw); free(om->surfh); free(om->stage);
    free(om->rgw); free(om->xf); free(om->rivflow); free(om->dhbydx);
    free(om->dhbydy); free(om->ws0); free(om->subrunoff); free(om);
}

int64_t oracle_num_state_var(const oracle_model *om)   /* ode.c:313-339 */
{
    return (int64_t)3 * om->ne + 2 * om->nr + (om->fbr ? 2 * (int64_t)om->ne : 0);
}

void oracle_set_forcing(oracle_model *om, const double *forc)
{
    memcpy(om->forc, forc, sizeof(double) * PB_F_NCOL * (size_t)om->ne);
}

void oracle_set_river_bc(oracle_model *om, const double *bc)
{
    memcpy(om->rivbc, bc, sizeof(double) * (size_t)om->nr);
}

/* hidden state H2: wf.ovlflow[] of river edges survives from the last call */
void oracle_set_stale_ovlflow(oracle_model *om, const double *ovl)
{
    memcpy(&X(PB_X_OVL0, 0), ovl, sizeof(double) * 3 * (size_t)om->ne);
}

void oracle_get_fluxes(const oracle_model *om, double *xf, double *rivflow)
{
    if (xf) memcpy(xf, om->xf, sizeof(double) * PB_X_NCOL * (size_t)om->ne);
    if (rivflow) memcpy(rivflow, om->rivflow,
        sizeof(double) * PIHM_B200_NUM_RIVFLX * (size_t)om->nr);
}

/* ---- Summary() + MassBalance(), src/update.c:3-160 ---------------------
 * ws0 is kept as a vector in the block layout of y (pihm_func.h:7-15). */
void oracle_set_ws0(oracle_model *om, const double *y)  /* initialize.c:598,612 */
{
    memcpy(om->ws0, y, sizeof(double) * (size_t)oracle_num_state_var(om));
}

void oracle_get_ws0(const oracle_model *om, double *y)
{
    memcpy(y, om->ws0, sizeof(double) * (size_t)oracle_num_state_var(om));
}

/* y = CV_Y after SolveCVode; the wf.* fields are those of the last oracle_ode
 * call.  Overwrites X(INFIL) (and X(FBR_INFIL)) like update.c:135,150-152 and
 * keeps the local `subrunoff` (update.c:128-133,154-158) for inspection. */
void oracle_summary(oracle_model *om, const double *y, double stepsize,
    double *subrunoff_out)
{
    const size_t    ne = (size_t)om->ne, nr = (size_t)om->nr;
    const size_t    o_unsat = ne, o_gw = 2 * ne, o_fu = 3 * ne + 2 * nr,
                    o_fg = 4 * ne + 2 * nr;
    size_t          i;
    int             j;

    for (i = 0; i < ne; i++)
    {
        /* update.c:19-25: ws = y (not clamped) */
        const double    unsat = y[o_unsat + i], gw = y[o_gw + i];
        const double    unsat0 = om->ws0[o_unsat + i], gw0 = om->ws0[o_gw + i];
        const double    depth = E(PB_E_DEPTH, i), area = E(PB_E_AREA, i);
        double          soilw0, soilw1, subrunoff, infil;

        /* MassBalance, update.c:103-160 */
        soilw0 = gw0 + unsat0;
        soilw0 = (soilw0 > depth) ? depth : soilw0;
        soilw0 = (soilw0 < 0.0) ? 0.0 : soilw0;

        soilw1 = gw + unsat;
        soilw1 = (soilw1 > depth) ? depth : soilw1;
        soilw1 = (soilw1 < 0.0) ? 0.0 : soilw1;

        subrunoff = 0.0;
        for (j = 0; j < 3; j++) subrunoff += X(PB_X_SUB0 + j, i) / area;

        infil = (soilw1 - soilw0) * E(PB_E_POROSITY, i) / stepsize + subrunoff +
            X(PB_X_EDIR_UNSAT, i) + X(PB_X_EDIR_GW, i) + X(PB_X_ETT_UNSAT, i) +
            X(PB_X_ETT_GW, i);

        if (om->fbr)
        {
            const double    fbrw0 = om->ws0[o_fg + i] + om->ws0[o_fu + i];
            const double    fbrw1 = y[o_fg + i] + y[o_fu + i];
            double          fbrrunoff = 0.0, fbr_infil;

            for (j = 0; j < 3; j++) fbrrunoff += X(PB_X_FBRFLOW0 + j, i) / area;
            fbr_infil = (fbrw1 - fbrw0) * E(PB_E_GPOROSITY, i) / stepsize +
                fbrrunoff;
            X(PB_X_FBR_INFIL, i) = fbr_infil;
            infil += fbr_infil;
        }

        if (infil < 0.0)
        {
            subrunoff -= infil;
            infil = 0.0;
        }
        X(PB_X_INFIL, i) = infil;
        om->subrunoff[i] = subrunoff;
        /* update.c:47: ws0 = ws; ws0.surf is what Infil() reads next step */
        F(PB_F_WS0SURF, i) = y[i];
    }
    /* update.c:47,94 */
    memcpy(om->ws0, y, sizeof(double) * (size_t)oracle_num_state_var(om));
    if (subrunoff_out) memcpy(subrunoff_out, om->subrunoff, sizeof(double) * ne);
}

/* ---- ApplyMeteoForc/ApplyLai per-element part + IntcpSnowEt ------------
 * forcing.c:134-160, 242-258; is_sm_et.c:4-225.  etf [PB_ET_NCOL][ne], eti
 * [PB_ETI_NCOL][ne]; out [PB_EO_NCOL][ne] carries ws.sneqv / ws.cmc in and
 * everything out; y supplies elem.ws.unsat / ws.gw (Summary, update.c:19-21).
 * Also writes wf.pcpdrp/edir/ett into the forcing table of the RHS. */
void oracle_intcp_snow_et(oracle_model *om, const pihm_b200_et_step *st,
    const double *etf, const int32_t *eti, const double *y, double *out)
{
    const double    CP = 1004.0, LVH2O = 2.501e6, SIGMA = 5.67e-8, RD = 287.04,
                    RV = 461.5;                      /* pihm_const.h:8-14 */
    const double    TSNOW = -3.0, TRAIN = 1.0, T0 = 0.0;
    const size_t    ne = (size_t)om->ne;
    const double    stepsize = st->stepsize;
    size_t          i;

#define TF(c) (etf[(size_t)(c) * ne + i])
#define TI(c) (eti[(size_t)(c) * ne + i])
#define O(c)  (out[(size_t)(c) * ne + i])
    for (i = 0; i < ne; i++)
    {
        const double   *mv = st->meteo +
            (size_t)(TI(PB_ETI_METEO_TYPE) - 1) * PIHM_B200_NUM_METEO_VAR;
        double          prcp = mv[0] / 1000.0, sfctmp = mv[1], rh = mv[2];
        double          wind = mv[3], soldn = (mv[4] > 0.0) ? mv[4] : 0.0;
        int             lc = TI(PB_ETI_LC_TYPE) - 1;
        double          lai = (TI(PB_ETI_LAI_TYPE) > 0) ?
            st->lai[TI(PB_ETI_LAI_TYPE) - 1] : st->lai_lc[lc];
        double          shdfac = TF(PB_ET_SHDFAC), cfactr = TF(PB_ET_CFACTR);
        double          depth = E(PB_E_DEPTH, i), rzd = E(PB_E_RZD, i);
        double          unsat = y[ne + i], gw = y[2 * ne + i];
        double          sneqv = O(PB_EO_SNEQV), cmc = O(PB_EO_CMC);
        double          albedo, radnet, vp, pres, qv, qvsat, frac_snow,
                        snow_rate, melt_rate, intcp_max, z0, zlvl, ra, gamma,
                        delta, etp, satn, betas, edir, ec, ett, drip, isval;

        albedo = 0.5 * (TF(PB_ET_ALBEDOMIN) + TF(PB_ET_ALBEDOMAX));
        radnet = soldn * (1.0 - albedo);
        sfctmp = sfctmp - 273.15;
        rh = rh / 100.0;
        vp = 611.2 * exp(17.67 * sfctmp / (sfctmp + 243.5)) * rh;
        pres = 101.325 * 1.0e3 *
            pow((293.0 - 0.0065 * E(PB_E_ZMAX, i)) / 293.0, 5.26);
        qv = 0.622 * vp / pres;
        qvsat = 0.622 * (vp / rh) / pres;

        frac_snow = (sfctmp < TSNOW) ? 1.0 :
            ((sfctmp > TRAIN) ? 0.0 : (TRAIN - sfctmp) / (TRAIN - TSNOW));
        snow_rate = frac_snow * prcp;
        sneqv += snow_rate * stepsize;
        melt_rate = (sfctmp > T0) ? (sfctmp - T0) * st->meltf : 0.0;
        if (sneqv > melt_rate * stepsize)
        {
            sneqv -= melt_rate * stepsize;
        }
        else
        {
            melt_rate = sneqv / stepsize;
            sneqv = 0.0;
        }

        intcp_max = TF(PB_ET_CMCFACTR) * lai * shdfac;
        z0 = st->z0_lc[lc];
        zlvl = TF(PB_ET_ZLVL_WIND);
        ra = log(zlvl / z0) * log(10.0 * zlvl / z0) / (wind * 0.16);
        gamma = 4.0 * 0.7 * SIGMA * RD / CP * pow(sfctmp + 273.15, 4) /
            (pres / ra) + 1.0;
        delta = LVH2O * LVH2O * 0.622 / RV / CP / pow(sfctmp + 273.15, 2) *
            qvsat;
        etp = (radnet * delta + gamma * (1.2 * LVH2O * (qvsat - qv) / ra)) /
            (1000.0 * LVH2O * (delta + gamma));

        if (depth - gw < rzd)
        {
            satn = 1.0;
        }
        else
        {
            double          ratio = unsat / (depth - gw);

            satn = (ratio > 1.0) ? 1.0 : ((ratio < 0.0) ? 0.0 :
                0.5 * (1.0 - cos(3.14 * ratio)));
        }
        betas = (satn * E(PB_E_POROSITY, i) + TF(PB_ET_SMCMIN) -
            TF(PB_ET_SMCWLT)) / (TF(PB_ET_SMCREF) - TF(PB_ET_SMCWLT));
        betas = (betas < 0.0001) ? 0.0001 : ((betas > 1.0) ? 1.0 : betas);
        edir = (1.0 - shdfac) * pow(betas, 2) * etp;
        edir *= st->cal_edir;
        edir = (edir < 0.0) ? 0.0 : edir;

        if (lai > 0.0)
        {
            double          cmc_c = (cmc < 0.0) ? 0.0 :
                ((cmc > intcp_max) ? intcp_max : cmc);
            double          cfrac = (cmc < 0.0) ? 0.0 :
                ((cmc > intcp_max) ? intcp_max : cmc) / intcp_max;
            double          rsmin = TF(PB_ET_RSMIN), rsmax = TF(PB_ET_RSMAX);
            double          fr, alphar, etas, gammas, rs, pc;

            ec = shdfac * pow(cmc_c / intcp_max, cfactr) * etp;
            ec *= st->cal_ec;
            ec = (ec < 0.0) ? 0.0 : ec;
            fr = 1.1 * radnet / (TF(PB_ET_RGL) * lai);
            fr = (fr < 0.0) ? 0.0 : fr;
            alphar = (1.0 + fr) / (fr + (rsmin / rsmax));
            alphar = (alphar > 10000.0) ? 10000.0 : alphar;
            etas = 1.0 - 0.0016 * (pow((TF(PB_ET_TOPT) - 273.15 - sfctmp), 2));
            etas = (etas < 0.0001) ? 0.0001 : etas;
            gammas = 1.0 / (1.0 + 0.00025 * (vp / rh - vp));
            gammas = (gammas < 0.01) ? 0.01 : gammas;
            rs = rsmin * alphar / (betas * lai * etas * gammas);
            rs = (rs > rsmax) ? rsmax : rs;
            pc = (1.0 + delta / gamma) / (1.0 + rs / ra + delta / gamma);
            ett = shdfac * pc * (1.0 - pow(cfrac, cfactr)) * etp;
            ett *= st->cal_ett;
            ett = (ett < 0.0) ? 0.0 : ett;
            ett = ((gw < (depth - rzd)) && unsat <= 0.0) ? 0.0 : ett;
            drip = (cmc <= 0.0) ? 0.0 :
                6.52E-7 * intcp_max * exp(3.89 * cmc / intcp_max);
        }
        else
        {
            ett = 0.0; ec = 0.0; drip = 0.0;
        }

        if (drip < 0.0) drip = 0.0;
        if (drip * stepsize > cmc) drip = cmc / stepsize;
        isval = cmc + (1.0 - frac_snow) * prcp * shdfac * stepsize -
            ec * stepsize - drip * stepsize;
        if (isval > intcp_max)
        {
            cmc = intcp_max;
            drip += (isval - intcp_max) / stepsize;
        }
        else if (isval < 0.0)
        {
            cmc = 0.0;
            if (ec + drip > 0.0)
            {
                ec = ec / (ec + drip) *
                    (cmc + (1.0 - frac_snow) * prcp * shdfac * stepsize);
                drip = drip / (ec + drip) *
                    (cmc + (1.0 - frac_snow) * prcp * shdfac * stepsize);
            }
        }
        else
        {
            cmc = isval;
        }
        O(PB_EO_PCPDRP) = (1.0 - shdfac) * (1.0 - frac_snow) * prcp + drip +
            melt_rate;
        O(PB_EO_EDIR) = edir; O(PB_EO_ETT) = ett; O(PB_EO_EC) = ec;
        O(PB_EO_DRIP) = drip; O(PB_EO_SNEQV) = sneqv; O(PB_EO_CMC) = cmc;
        F(PB_F_PCPDRP, i) = O(PB_EO_PCPDRP);
        F(PB_F_EDIR, i) = edir;
        F(PB_F_ETT, i) = ett;
    }
#undef TF
#undef TI
#undef O
}

/* ---- small physics helpers ------------------------------------------- */

static double surf_h(double surfeqv)                 /* hydrol.c:92-126 */
{
    if (surfeqv < 0.0) return 0.0;
    if (surfeqv <= 0.5 * DEPRSTG) return sqrt(2.0 * DEPRSTG * surfeqv);
    return DEPRSTG + (surfeqv - 0.5 * DEPRSTG);
}

static double avg_hsurf(double diff, double hsurf, double hnabr) /* lat_flow.c:175-203 */
{
    if (diff > 0.0)
        return (hsurf > DEPRSTG) ? 1.0 * (hsurf - DEPRSTG) : 0.0;
    return (hnabr > DEPRSTG) ? 1.0 * (hnabr - DEPRSTG) : 0.0;
}

static double avg_h(double diff, double hsub, double hnabr)      /* lat_flow.c:205-225 */
{
    double a = 0.0;
    if (diff > 0.0) { if (hsub > 0.0) a = hsub; }
    else { if (hnabr > 0.0) a = hnabr; }
    return a;
}

static double dh_by_dl(const double *l1, const double *l2, const double *h) /* lat_flow.c:227-234 */
{
    return -1.0 *
        (l1[2] * (h[1] - h[0]) + l1[1] * (h[0] - h[2]) + l1[0] * (h[2] - h[1])) /
        (l2[2] * (l1[1] - l1[0]) + l2[1] * (l1[0] - l1[2]) +
        l2[0] * (l1[2] - l1[1]));
}

/* EffKh of element i at groundwater level gw, lat_flow.c:236-265 */
static double eff_kh(const oracle_model *om, int i, double gw)
{
    double depth = E(PB_E_DEPTH, i), dmac = E(PB_E_DMAC, i);
    double ksath = E(PB_E_KSATH, i);
    double k1, k2, d1, d2;

    gw = (gw > 0.0) ? gw : 0.0;
    if (gw > depth - dmac)
    {
        k1 = E(PB_E_KMACH, i) * E(PB_E_AREAFV, i) +
            ksath * (1.0 - E(PB_E_AREAFV, i));
        k2 = ksath;
        if (gw > depth) { d1 = dmac; d2 = depth - dmac; }
        else { d1 = gw - (depth - dmac); d2 = depth - dmac; }
        return (k1 * d1 + k2 * d2) / (d1 + d2);
    }
    return ksath;
}

static double overland_flow(double avgh, double grad, double sf, double crossa,
    double rough)                                    /* lat_flow.c:267-271 */
{
    return crossa * pow(avgh, 0.6666667) * grad / (sqrt(sf) * rough);
}

static double kr_func(double beta, double satn)      /* soil.c:3-8 */
{
    return sqrt(satn) *
        (1.0 - pow(1.0 - pow(satn, beta / (beta - 1.0)), (beta - 1.0) / beta)) *
        (1.0 - pow(1.0 - pow(satn, beta / (beta - 1.0)), (beta - 1.0) / beta));
}

static double psi_func(double satn, double alpha, double beta) /* vert_flow.c:272-278 */
{
    satn = (satn < SATMIN) ? SATMIN : satn;
    return -pow(pow(1.0 / satn, beta / (beta - 1.0)) - 1.0, 1.0 / beta) / alpha;
}

/* vert_flow.c:211-270 */
static double eff_kinf(const oracle_model *om, int i, double dh_by_dz,
    double ksatfunc, double elemsatn, double applrate, double surfh)
{
    double kinfv = E(PB_E_KINFV, i), kmacv = E(PB_E_KMACV, i);
    double areafh = E(PB_E_AREAFH, i);
    double keff, kmax;

    if (areafh == 0.0)
        keff = kinfv * ksatfunc;
    else if (surfh > DEPRSTG)
        keff = kinfv * (1.0 - areafh) * ksatfunc + kmacv * areafh;
    else if (applrate <= dh_by_dz * kinfv * ksatfunc)
        keff = kinfv * ksatfunc;
    else
    {
        kmax = dh_by_dz * (kmacv * areafh + kinfv * (1.0 - areafh) * ksatfunc);
        if (applrate < kmax)
            keff = kinfv * (1.0 - areafh) * ksatfunc +
                kmacv * areafh * kr_func(2.0, elemsatn);
        else
            keff = kinfv * (1.0 - areafh) * ksatfunc + kmacv * areafh;
    }
    return keff;
}

/* vert_flow.c:33-130 (non-Noah) */
static double infil_func(const oracle_model *om, int i)
{
    double depth = E(PB_E_DEPTH, i), dinf = E(PB_E_DINF, i);
    double area = E(PB_E_AREA, i), zmax = E(PB_E_ZMAX, i), zmin = E(PB_E_ZMIN, i);
    double surfh = om->surfh[i], gw = om->gw[i], unsat = om->unsat[i];
    double applrate, wetfrac, dh_by_dz, satn, satkfunc, infil, infil_max, kinf;
    double deficit, psi_u, h_u, ws0surf;
    int    j;

    if (unsat + gw > depth) return 0.0;

    applrate = 0.0;
    for (j = 0; j < 3; j++) applrate += -X(PB_X_OVL0 + j, i) / area;
    applrate = (applrate > 0.0) ? applrate : 0.0;
    applrate += F(PB_F_PCPDRP, i);

    wetfrac = surfh / DEPRSTG;
    wetfrac = (wetfrac > 0.0) ? wetfrac : 0.0;
    wetfrac = (wetfrac < 1.0) ? wetfrac : 1.0;

    if (gw > depth - dinf)
    {
        dh_by_dz = (surfh + zmax - (gw + zmin)) / (0.5 * (surfh + dinf));
        dh_by_dz = (surfh <= 0.0 && dh_by_dz > 0.0) ? 0.0 : dh_by_dz;
        satn = 1.0;
        satkfunc = kr_func(E(PB_E_BETA, i), satn);
        kinf = eff_kinf(om, i, dh_by_dz, satkfunc, satn, applrate, surfh);
        infil = kinf * dh_by_dz;
    }
    else
    {
        deficit = depth - gw;
        satn = unsat / deficit;
        satn = (satn > 1.0) ? 1.0 : satn;
        satn = (satn < SATMIN) ? SATMIN : satn;
        psi_u = psi_func(satn, E(PB_E_ALPHA, i), E(PB_E_BETA, i));
        psi_u = (psi_u > PSIMIN) ? psi_u : PSIMIN;
        h_u = psi_u + zmax - 0.5 * dinf;
        dh_by_dz = (surfh + zmax - h_u) / (0.5 * (surfh + dinf));
        dh_by_dz = (surfh <= 0.0 && dh_by_dz > 0.0) ? 0.0 : dh_by_dz;
        satkfunc = kr_func(E(PB_E_BETA, i), satn);
        kinf = eff_kinf(om, i, dh_by_dz, satkfunc, satn, applrate, surfh);
        infil = kinf * dh_by_dz;
        infil = (infil > 0.0) ? infil : 0.0;
    }

    ws0surf = F(PB_F_WS0SURF, i);
    infil_max = applrate + ((ws0surf > 0.0) ? ws0surf / om->dt : 0.0);
    infil = (infil > infil_max) ? infil_max : infil;
    infil *= wetfrac;
    return infil;
}

/* vert_flow.c:172-209 with _ARITH_ (pihm_func.h:4) */
static double avg_kv(const oracle_model *om, int i, double deficit, double gw,
    double satkfunc)
{
    double ksatv = E(PB_E_KSATV, i), dmac = E(PB_E_DMAC, i);
    double areafh = E(PB_E_AREAFH, i), kmacv = E(PB_E_KMACV, i);
    double k1, k2, k3, d1, d2, d3;

    if (deficit > dmac)
    {
        k1 = satkfunc * ksatv; d1 = dmac;
        k2 = satkfunc * ksatv; d2 = deficit - dmac;
        k3 = ksatv; d3 = gw;
    }
    else
    {
        k1 = satkfunc * ksatv; d1 = deficit;
        k2 = (areafh > 0.0) ? kmacv * areafh + ksatv * (1.0 - areafh) : ksatv;
        d2 = dmac - deficit;
        k3 = ksatv; d3 = gw - (dmac - deficit);
    }
    return (k1 * d1 + k2 * d2 + k3 * d3) / (d1 + d2 + d3);
}

/* vert_flow.c:132-170 */
static double recharge_func(const oracle_model *om, int i, double infil)
{
    double depth = E(PB_E_DEPTH, i), dinf = E(PB_E_DINF, i);
    double gw = om->gw[i], unsat = om->unsat[i];
    double deficit, satn, satkfunc, psi_u, dh_by_dz, kavg, rechg;

    if (gw > depth - dinf) return infil;
    deficit = depth - gw;
    satn = unsat / deficit;
    satn = (satn > 1.0) ? 1.0 : satn;
    satn = (satn < SATMIN) ? SATMIN : satn;
    satkfunc = kr_func(E(PB_E_BETA, i), satn);
    psi_u = psi_func(satn, E(PB_E_ALPHA, i), E(PB_E_BETA, i));
    dh_by_dz = (0.5 * deficit + psi_u) / (0.5 * (deficit + gw));
    kavg = avg_kv(om, i, deficit, gw, satkfunc);
    rechg = kavg * dh_by_dz;
    rechg = (rechg > 0.0 && unsat <= 0.0) ? 0.0 : rechg;
    rechg = (rechg < 0.0 && gw <= 0.0) ? 0.0 : rechg;
    return rechg;
}

/* vert_flow.c:284-330 */
static double fbr_infil_func(const oracle_model *om, int i)
{
    double gdepth = E(PB_E_GDEPTH, i), zmin = E(PB_E_ZMIN, i);
    double fu = om->fbr_unsat[i], fg = om->fbr_gw[i], gw = om->gw[i];
    double deficit, satn, psi_u, h_u, satkfunc, dh_by_dz, kavg;

    if (fg >= gdepth) return -E(PB_E_KSATV, i);
    if (fu + fg > gdepth || gw <= 0.0) return 0.0;
    deficit = gdepth - fg;
    satn = fu / deficit;
    satn = (satn > 1.0) ? 1.0 : satn;
    satn = (satn < SATMIN) ? SATMIN : satn;
    psi_u = psi_func(satn, E(PB_E_GALPHA, i), E(PB_E_GBETA, i));
    psi_u = (psi_u > PSIMIN) ? psi_u : PSIMIN;
    h_u = psi_u + zmin - 0.5 * deficit;
    satkfunc = kr_func(E(PB_E_GBETA, i), satn);
    dh_by_dz = (zmin + gw - h_u) / (0.5 * (gw + deficit));
    kavg = (gw + deficit) /
        (gw / E(PB_E_KSATV, i) + deficit / (E(PB_E_GKSATV, i) * satkfunc));
    return kavg * dh_by_dz;
}

/* vert_flow.c:332-373 */
static double fbr_recharge_func(const oracle_model *om, int i, double fbr_infil)
{
    double gdepth = E(PB_E_GDEPTH, i), gksatv = E(PB_E_GKSATV, i);
    double fu = om->fbr_unsat[i], fg = om->fbr_gw[i];
    double deficit, satn, psi_u, satkfunc, dh_by_dz, kavg, rechg;

    if (fg >= gdepth) return fbr_infil;
    deficit = gdepth - fg;
    satn = fu / deficit;
    satn = (satn > 1.0) ? 1.0 : satn;
    satn = (satn < SATMIN) ? SATMIN : satn;
    psi_u = psi_func(satn, E(PB_E_GALPHA, i), E(PB_E_GBETA, i));
    psi_u = (psi_u > PSIMIN) ? psi_u : PSIMIN;
    satkfunc = kr_func(E(PB_E_GBETA, i), satn);
    dh_by_dz = (0.5 * deficit + psi_u) / (0.5 * (deficit + fg));
    kavg = (fu * gksatv * satkfunc + fg * gksatv) / (fu + fg);
    rechg = kavg * dh_by_dz;
    rechg = (rechg > 0.0 && fu <= 0.0) ? 0.0 : rechg;
    rechg = (rechg < 0.0 && fg <= 0.0) ? 0.0 : rechg;
    return rechg;
}

/* ---- river helpers ------------------------------------------------------ */

static double riv_area(int order, double depth, double coeff)  /* river_flow.c:518-546 */
{
    depth = (depth > 0.0) ? depth : 0.0;
    switch (order)
    {
        case RECTANGLE: return depth * coeff;
        case TRIANGLE:  return depth * depth / coeff;
        case QUADRATIC: return 4.0 * depth * sqrt(depth) / (3.0 * sqrt(coeff));
        case CUBIC:
            return 3.0 * pow(depth, 4.0 / 3.0) / (2.0 * pow(coeff, 1.0 / 3.0));
    }
    return 0.0;
}

static double riv_perim(int order, double depth, double coeff) /* river_flow.c:548-581 */
{
    depth = (depth > 0.0) ? depth : 0.0;
    switch (order)
    {
        case RECTANGLE: return 2.0 * depth + coeff;
        case TRIANGLE:  return 2.0 * depth * sqrt(1.0 + coeff * coeff) / coeff;
        case QUADRATIC:
            return sqrt(depth * (1.0 + 4.0 * coeff * depth) / coeff) +
                log(2.0 * sqrt(coeff * depth) +
                sqrt(1.0 + 4.0 * coeff * depth)) / (2.0 * coeff);
        case CUBIC:
            return 2.0 * ((pow(depth * (1.0 + 9.0 * pow(coeff, 2.0 / 3.0) *
                depth), 0.5) / 3.0) +
                (log(3.0 * pow(coeff, 1.0 / 3.0) * sqrt(depth) +
                pow(1.0 + 9.0 * pow(coeff, 2.0 / 3.0) * depth, 0.5)) /
                (9.0 * pow(coeff, 1.0 / 3.0))));
    }
    return 0.0;
}

/* river_flow.c:183-253; e = bank element (0-based), r = river (0-based) */
static double ovl_elem_to_river(const oracle_model *om, int e, int r)
{
    double zbank, flux, elem_h, rivseg_h;
    double rzmax = R(PB_R_ZMAX, r), ezmax = E(PB_E_ZMAX, e);
    double cwr = R(PB_R_CWR, r), len = R(PB_R_SHP_LENGTH, r);

    zbank = (rzmax > ezmax) ? rzmax : ezmax;
    elem_h = ezmax + om->surfh[e];
    rivseg_h = R(PB_R_ZBED, r) + om->stage[r];

    if (rivseg_h > elem_h)
    {
        if (elem_h > zbank)
            flux = cwr * 2.0 * sqrt(2.0 * GRAV) * len * sqrt(rivseg_h - elem_h) *
                (rivseg_h - zbank) / 3.0;
        else if (zbank < rivseg_h)
            flux = cwr * 2.0 * sqrt(2.0 * GRAV) * len * sqrt(rivseg_h - zbank) *
                (rivseg_h - zbank) / 3.0;
        else
            flux = 0.0;
    }
    else if (om->surfh[e] > DEPRSTG)
    {
        if (rivseg_h > zbank)
            flux = -cwr * 2.0 * sqrt(2.0 * GRAV) * len * sqrt(elem_h - rivseg_h) *
                (elem_h - zbank) / 3.0;
        else if (zbank < elem_h)
            flux = -cwr * 2.0 * sqrt(2.0 * GRAV) * len * sqrt(elem_h - zbank) *
                (elem_h - zbank) / 3.0;
        else
            flux = 0.0;
    }
    else
        flux = 0.0;
    return flux;
}

/* river_flow.c:427-458 */
static double chan_elem_to_river(const oracle_model *om, int e, double effk,
    int r, double distance)
{
    double diff_h, avgh, grad_h, avg_ksat;
    double zbed = R(PB_R_ZBED, r), ezmin = E(PB_E_ZMIN, e), egw = om->gw[e];

    diff_h = (om->stage[r] + zbed) - (egw + ezmin);
    if (ezmin > zbed) avgh = egw;
    else if (ezmin + egw > zbed) avgh = ezmin + egw - zbed;
    else avgh = 0.0;
    avgh = avg_h(diff_h, om->stage[r], avgh);
    grad_h = diff_h / distance;
    avg_ksat = 0.5 * (effk + R(PB_R_KSATH, r));
    return R(PB_R_SHP_LENGTH, r) * avg_ksat * grad_h * avgh;
}

/* river_flow.c:460-496 with _ARITH_ */
static double sub_elem_to_river(const oracle_model *om, int e, double effk,
    int r, double effk_riv, double distance)
{
    double diff_h, avgh, avg_ksat, grad_h;
    double zbed = R(PB_R_ZBED, r), ezmin = E(PB_E_ZMIN, e), egw = om->gw[e];

    diff_h = (om->rgw[r] + R(PB_R_ZMIN, r)) - (egw + ezmin);
    if (ezmin > zbed) avgh = 0.0;
    else if (ezmin + egw > zbed) avgh = zbed - ezmin;
    else avgh = egw;
    avgh = avg_h(diff_h, om->rgw[r], avgh);
    avg_ksat = 0.5 * (effk + effk_riv);
    grad_h = diff_h / distance;
    return R(PB_R_SHP_LENGTH, r) * avg_ksat * grad_h * avgh;
}

/* river_flow.c:255-298 */
static double chan_river_to_river(const oracle_model *om, int r, int d)
{
    int    ord = RI(PB_RI_INTRPL_ORD, r), ordd = RI(PB_RI_INTRPL_ORD, d);
    double total_h, perim, total_h_down, perim_down, avg_perim, avg_rough;
    double distance, diff_h, grad_h, avg_sf, crossa, crossa_down, avg_crossa, avgh;

    total_h = om->stage[r] + R(PB_R_ZBED, r);
    perim = riv_perim(ord, om->stage[r], R(PB_R_SHP_COEFF, r));
    total_h_down = om->stage[d] + R(PB_R_ZBED, d);
    perim_down = riv_perim(ordd, om->stage[d], R(PB_R_SHP_COEFF, d));
    avg_perim = (perim + perim_down) / 2.0;
    avg_rough = (R(PB_R_ROUGH, r) + R(PB_R_ROUGH, d)) / 2.0;
    distance = 0.5 * (R(PB_R_SHP_LENGTH, r) + R(PB_R_SHP_LENGTH, d));
    diff_h = (om->riv_mode == KINEMATIC) ?
        (R(PB_R_ZBED, r) - R(PB_R_ZBED, d)) : (total_h - total_h_down);
    grad_h = diff_h / distance;
    avg_sf = (grad_h > 0.0) ? grad_h : RIVGRADMIN;
    crossa = riv_area(ord, om->stage[r], R(PB_R_SHP_COEFF, r));
    crossa_down = riv_area(ordd, om->stage[d], R(PB_R_SHP_COEFF, d));
    avg_crossa = 0.5 * (crossa + crossa_down);
    avgh = (avg_perim == 0.0) ? 0.0 : (avg_crossa / avg_perim);
    return overland_flow(avgh, grad_h, avg_sf, crossa, avg_rough);
}

/* river_flow.c:300-329 with _ARITH_ */
static double sub_river_to_river(const oracle_model *om, int r, double effk,
    int d, double effk_nabr)
{
    double total_h, total_h_down, avg_wid, diff_h, avgh, distance, grad_h, avg_ksat;

    total_h = om->rgw[r] + R(PB_R_ZMIN, r);
    total_h_down = om->rgw[d] + R(PB_R_ZMIN, d);
    avg_wid = (R(PB_R_SHP_WIDTH, r) + R(PB_R_SHP_WIDTH, d)) / 2.0;
    diff_h = total_h - total_h_down;
    avgh = avg_h(diff_h, om->rgw[r], om->rgw[d]);
    distance = 0.5 * (R(PB_R_SHP_LENGTH, r) + R(PB_R_SHP_LENGTH, d));
    grad_h = diff_h / distance;
    avg_ksat = 0.5 * (effk + effk_nabr);
    return avg_ksat * grad_h * avgh * avg_wid;
}

/* river_flow.c:331-388 */
static double outlet_flux(const oracle_model *om, int r, int down)
{
    int    ord = RI(PB_RI_INTRPL_ORD, r);
    double stage = om->stage[r], coeff = R(PB_R_SHP_COEFF, r);
    double total_h, total_h_down, distance, grad_h, avgh, avg_perim, crossa;
    double discharge = 0.0, bchead = om->rivbc[r];

    switch (down)
    {
        case BC_DIRICHLET:
            total_h = stage + R(PB_R_ZBED, r);
            total_h_down = bchead;
            distance = 0.5 * R(PB_R_SHP_LENGTH, r);
            grad_h = (total_h - total_h_down) / distance;
            avg_perim = riv_perim(ord, stage, coeff);
            crossa = riv_area(ord, stage, coeff);
            avgh = (avg_perim == 0.0) ? 0.0 : (crossa / avg_perim);
            discharge = overland_flow(avgh, grad_h, grad_h, crossa,
                R(PB_R_ROUGH, r));
            break;
        case BC_NEUMANN:
            discharge = -bchead;
            break;
        case BC_ZERO_DPTH_GRAD:
            distance = 0.5 * R(PB_R_SHP_LENGTH, r);
            grad_h = (R(PB_R_ZBED, r) -
                (R(PB_R_NODE_ZMAX, r) - R(PB_R_SHP_DEPTH, r))) / distance;
            avg_perim = riv_perim(ord, stage, coeff);
            crossa = riv_area(ord, stage, coeff);
            discharge = sqrt(grad_h) * crossa * ((avg_perim > 0.0) ?
                pow(crossa / avg_perim, 2.0 / 3.0) : 0.0) / R(PB_R_ROUGH, r);
            break;
        case BC_CRIT_DPTH:
            crossa = riv_area(ord, stage, coeff);
            discharge = crossa * sqrt(GRAV * stage);
            break;
        default:
            discharge = NAN;    /* reference: PIHMexit */
    }
    return discharge;
}

/* river_flow.c:390-425 */
static double bound_flux_river(const oracle_model *om, int r, int bctype)
{
    int    ord = RI(PB_RI_INTRPL_ORD, r);
    double stage = om->stage[r], coeff = R(PB_R_SHP_COEFF, r);
    double total_h, total_h_down, distance, grad_h, avgh, avg_perim, crossa;
    double flux = 0.0;

    if (bctype > 0)
    {
        total_h = stage + R(PB_R_ZBED, r);
        total_h_down = om->rivbc[r];
        distance = 0.5 * R(PB_R_SHP_LENGTH, r);
        grad_h = (total_h - total_h_down) / distance;
        avg_perim = riv_perim(ord, stage, coeff);
        crossa = riv_area(ord, stage, coeff);
        avgh = (avg_perim == 0.0) ? 0.0 : (crossa / avg_perim);
        flux = overland_flow(avgh, grad_h, grad_h, crossa, R(PB_R_ROUGH, r));
    }
    else if (bctype < 0)
        flux = -om->rivbc[r];
    return flux;
}

/* river_flow.c:498-516 */
static double chan_leak(const oracle_model *om, int r)
{
    double diff_h, grad_h;

    if (R(PB_R_ZBED, r) - (om->rgw[r] + R(PB_R_ZMIN, r)) > 0.0)
        diff_h = om->stage[r];
    else
        diff_h = om->stage[r] + R(PB_R_ZBED, r) - (om->rgw[r] + R(PB_R_ZMIN, r));
    grad_h = diff_h / R(PB_R_BEDTHICK, r);
    return R(PB_R_KSATV, r) * R(PB_R_SHP_WIDTH, r) * R(PB_R_SHP_LENGTH, r) * grad_h;
}

/* ---- the RHS ------------------------------------------------------------ */

/* lat_flow.c:118-173 */
static void frict_slope(oracle_model *om)
{
    int i, j;

    if (om->surf_mode != DIFF_WAVE) return;
    for (i = 0; i < om->ne; i++)
    {
        double h[3], nx[3], ny[3];

        for (j = 0; j < 3; j++)
        {
            int nb = EI(PB_EI_NABR0 + j, i);

            nx[j] = E(PB_E_NABRX0 + j, i);
            ny[j] = E(PB_E_NABRY0 + j, i);
            if (nb > 0)
                h[j] = E(PB_E_ZMAX, nb - 1) + om->surfh[nb - 1];
            else if (nb < 0)
            {
                int r = -nb - 1;
                h[j] = (om->stage[r] > R(PB_R_SHP_DEPTH, r)) ?
                    R(PB_R_ZBED, r) + om->stage[r] : R(PB_R_ZMAX, r);
            }
            else if (EI(PB_EI_BC0 + j, i) == 0)
                h[j] = E(PB_E_ZMAX, i) + om->surfh[i];
            else
                h[j] = F(PB_F_BC0 + j, i);
        }
        om->dhbydx[i] = dh_by_dl(ny, nx, h);
        om->dhbydy[i] = dh_by_dl(nx, ny, h);
    }
}

/* lat_flow.c:3-116 */
static void lateral_flow(oracle_model *om)
{
    int i, j, k;

    frict_slope(om);
    for (i = 0; i < om->ne; i++)
    {
        for (j = 0; j < 3; j++)
        {
            int nb = EI(PB_EI_NABR0 + j, i);

            if (nb > 0)
            {
                int    n = nb - 1;
                double diff_h, avgh, grad_h, effk, effk_n, avg_ksat, avg_sf;
                double avg_rough, crossa;

                /* SubFlowElemToElem, lat_flow.c:273-298 */
                diff_h = (om->gw[i] + E(PB_E_ZMIN, i)) - (om->gw[n] + E(PB_E_ZMIN, n));
                avgh = avg_h(diff_h, om->gw[i], om->gw[n]);
                grad_h = diff_h / E(PB_E_NABRDIST0 + j, i);
                effk = eff_kh(om, i, om->gw[i]);
                effk_n = eff_kh(om, n, om->gw[n]);
                avg_ksat = 0.5 * (effk + effk_n);
                X(PB_X_SUB0 + j, i) = avg_ksat * grad_h * avgh * E(PB_E_EDGE0 + j, i);

                /* lat_flow.c:33-38, OvlFlowElemToElem :300-327 */
                avg_sf = 0.5 *
                    (sqrt(om->dhbydx[i] * om->dhbydx[i] + om->dhbydy[i] * om->dhbydy[i]) +
                     sqrt(om->dhbydx[n] * om->dhbydx[n] + om->dhbydy[n] * om->dhbydy[n]));
                diff_h = (om->surf_mode == KINEMATIC) ?
                    E(PB_E_ZMAX, i) - E(PB_E_ZMAX, n) :
                    (om->surfh[i] + E(PB_E_ZMAX, i)) - (om->surfh[n] + E(PB_E_ZMAX, n));
                avgh = avg_hsurf(diff_h, om->surfh[i], om->surfh[n]);
                grad_h = diff_h / E(PB_E_NABRDIST0 + j, i);
                if (om->surf_mode == KINEMATIC)
                    avg_sf = (grad_h > 0.0) ? grad_h : GRADMIN;
                else
                    avg_sf = (avg_sf > GRADMIN) ? avg_sf : GRADMIN;
                avg_rough = 0.5 * (E(PB_E_ROUGH, i) + E(PB_E_ROUGH, n));
                crossa = avgh * E(PB_E_EDGE0 + j, i);
                X(PB_X_OVL0 + j, i) = overland_flow(avgh, grad_h, avg_sf, crossa, avg_rough);
            }
            else if (nb < 0)
            {
                /* river edge: left to river_flow (and stale until then) */
            }
            else
            {
                /* BoundFluxElem, lat_flow.c:329-371 */
                int bc = EI(PB_EI_BC0 + j, i);

                if (bc == 0)
                {
                    X(PB_X_OVL0 + j, i) = 0.0;
                    X(PB_X_SUB0 + j, i) = 0.0;
                }
                else if (bc > 0)
                {
                    double head = F(PB_F_BC0 + j, i);
                    double diff_h, avgh, effk, grad_h;

                    X(PB_X_OVL0 + j, i) = 0.0;
                    diff_h = om->gw[i] + E(PB_E_ZMIN, i) - head;
                    avgh = avg_h(diff_h, om->gw[i], head - E(PB_E_ZMIN, i));
                    effk = eff_kh(om, i, om->gw[i]);
                    grad_h = diff_h / E(PB_E_NABRDIST0 + j, i);
                    X(PB_X_SUB0 + j, i) = effk * grad_h * avgh * E(PB_E_EDGE0 + j, i);
                }
                else
                {
                    X(PB_X_OVL0 + j, i) = 0.0;
                    X(PB_X_SUB0 + j, i) = -F(PB_F_BC0 + j, i);
                }
            }
        }
    }

    if (!om->fbr) return;
    /* lat_flow.c:56-115 */
    for (i = 0; i < om->ne; i++)
    {
        for (j = 0; j < 3; j++)
        {
            int nb = EI(PB_EI_NABR0 + j, i);

            if (nb == 0)
            {
                /* FbrBoundFluxElem, lat_flow.c:392-424 */
                int    bc = EI(PB_EI_FBRBC0 + j, i);
                double flux;

                if (bc == 0) flux = 0.0;
                else if (bc > 0)
                {
                    double head = F(PB_F_FBRBC0 + j, i);
                    double diff_h = om->fbr_gw[i] + E(PB_E_ZBED, i) - head;
                    double avgh = avg_h(diff_h, om->fbr_gw[i], head - E(PB_E_ZBED, i));
                    double effk = E(PB_E_GKSATH, i);
                    double grad_h = diff_h / E(PB_E_NABRDIST0 + j, i);
                    flux = effk * grad_h * avgh * E(PB_E_EDGE0 + j, i);
                }
                else flux = -F(PB_F_FBRBC0 + j, i);
                X(PB_X_FBRFLOW0 + j, i) = flux;
            }
            else
            {
                int    n;
                double dist = 0.0, diff_h, avgh, grad_h, avg_ksat;

                if (nb > 0)
                {
                    n = nb - 1;
                    dist = E(PB_E_NABRDIST0 + j, i);
                }
                else
                {
                    int r = -nb - 1;
                    n = (RI(PB_RI_LEFTELE, r) == i + 1) ?
                        RI(PB_RI_RIGHTELE, r) - 1 : RI(PB_RI_LEFTELE, r) - 1;
                    for (k = 0; k < 3; k++)
                    {
                        if (EI(PB_EI_NABR0 + k, n) == nb)
                        {
                            dist = E(PB_E_NABRDIST0 + j, i) + E(PB_E_NABRDIST0 + k, n);
                            break;
                        }
                    }
                }
                /* FbrFlowElemToElem, lat_flow.c:374-390 */
                diff_h = (om->fbr_gw[i] + E(PB_E_ZBED, i)) -
                    (om->fbr_gw[n] + E(PB_E_ZBED, n));
                avgh = avg_h(diff_h, om->fbr_gw[i], om->fbr_gw[n]);
                grad_h = diff_h / dist;
                avg_ksat = 0.5 * (E(PB_E_GKSATH, i) + E(PB_E_GKSATH, n));
                X(PB_X_FBRFLOW0 + j, i) = avg_ksat * grad_h * avgh * E(PB_E_EDGE0 + j, i);
            }
        }
    }
}

/* river_flow.c:3-109 */
static void river_flow(oracle_model *om)
{
    int i, j;

    for (i = 0; i < om->nr; i++)
    {
        int    down = RI(PB_RI_DOWN, i);
        int    l = RI(PB_RI_LEFTELE, i) - 1, r = RI(PB_RI_RIGHTELE, i) - 1;
        double effk, effk_nabr, effk_left, effk_right;

        if (down > 0)
        {
            int d = down - 1;
            int dl = RI(PB_RI_LEFTELE, d) - 1, dr = RI(PB_RI_RIGHTELE, d) - 1;

            if (RI(PB_RI_BCTYPE, i) != 0)
                RFLX(UP_C2C, i) += bound_flux_river(om, i, RI(PB_RI_BCTYPE, i));
            RFLX(DOWN_C2C, i) = chan_river_to_river(om, i, d);
            effk = 0.5 * (eff_kh(om, l, om->gw[l]) + eff_kh(om, r, om->gw[r]));
            effk_nabr = 0.5 * (eff_kh(om, dl, om->gw[dl]) + eff_kh(om, dr, om->gw[dr]));
            RFLX(DOWN_A2A, i) = sub_river_to_river(om, i, effk, d, effk_nabr);
        }
        else
        {
            RFLX(DOWN_C2C, i) = outlet_flux(om, i, down);
            RFLX(DOWN_A2A, i) = 0.0;
        }

        /* RiverToElem, river_flow.c:111-181 (leftele/rightele > 0 always here) */
        RFLX(LEFT_S2C, i) = ovl_elem_to_river(om, l, i);
        RFLX(RIGHT_S2C, i) = ovl_elem_to_river(om, r, i);
        effk_left = eff_kh(om, l, om->gw[l]);
        effk_right = eff_kh(om, r, om->gw[r]);
        RFLX(LEFT_A2C, i) = chan_elem_to_river(om, l, effk_left, i, R(PB_R_DIST_LEFT, i));
        RFLX(RIGHT_A2C, i) = chan_elem_to_river(om, r, effk_right, i, R(PB_R_DIST_RIGHT, i));
        RFLX(LEFT_A2A, i) = sub_elem_to_river(om, l, effk_left, i,
            0.5 * (effk_left + effk_right), R(PB_R_DIST_LEFT, i));
        RFLX(RIGHT_A2A, i) = sub_elem_to_river(om, r, effk_right, i,
            0.5 * (effk_left + effk_right), R(PB_R_DIST_RIGHT, i));
        for (j = 0; j < 3; j++)
        {
            if (EI(PB_EI_NABR0 + j, l) == -(i + 1))
            {
                X(PB_X_OVL0 + j, l) = -RFLX(LEFT_S2C, i);
                X(PB_X_SUB0 + j, l) = -(RFLX(LEFT_A2C, i) + RFLX(LEFT_A2A, i));
                break;
            }
        }
        for (j = 0; j < 3; j++)
        {
            if (EI(PB_EI_NABR0 + j, r) == -(i + 1))
            {
                X(PB_X_OVL0 + j, r) = -RFLX(RIGHT_S2C, i);
                X(PB_X_SUB0 + j, r) = -(RFLX(RIGHT_A2C, i) + RFLX(RIGHT_A2A, i));
                break;
            }
        }
        RFLX(CHANL_LKG, i) = chan_leak(om, i);
    }

    /* serial upstream accumulation, river_flow.c:94-108 */
    for (i = 0; i < om->nr; i++)
    {
        int down = RI(PB_RI_DOWN, i);

        if (down > 0)
        {
            RFLX(UP_C2C, down - 1) -= RFLX(DOWN_C2C, i);
            RFLX(UP_A2A, down - 1) -= RFLX(DOWN_A2A, i);
        }
    }
}

/* ODE(), src/ode.c:3-300; returns 0, or 1 if any dy is NaN (CheckDy) */
int oracle_ode(oracle_model *om, double t, const double *y, double *dy)
{
    const int ne = om->ne, nr = om->nr;
    const double *ysurf = y, *yunsat = y + ne, *ygw = y + 2 * (size_t)ne;
    const double *ystg = y + 3 * (size_t)ne, *yrgw = y + 3 * (size_t)ne + nr;
    const double *yfu = y + 3 * (size_t)ne + 2 * (size_t)nr;
    const double *yfg = y + 4 * (size_t)ne + 2 * (size_t)nr;
    double *dsurf = dy, *dunsat = dy + ne, *dgw = dy + 2 * (size_t)ne;
    double *dstg = dy + 3 * (size_t)ne, *drgw = dy + 3 * (size_t)ne + nr;
    double *dfu = dy + 3 * (size_t)ne + 2 * (size_t)nr;
    double *dfg = dy + 4 * (size_t)ne + 2 * (size_t)nr;
    int i, j, nan_flag = 0;

    (void)t;
    for (i = 0; i < oracle_num_state_var(om); i++) dy[i] = 0.0;

    for (i = 0; i < ne; i++)                       /* ode.c:25-49 */
    {
        om->surf[i] = (ysurf[i] >= 0.0) ? ysurf[i] : 0.0;
        om->unsat[i] = (yunsat[i] >= 0.0) ? yunsat[i] : 0.0;
        om->gw[i] = (ygw[i] >= 0.0) ? ygw[i] : 0.0;
        if (om->fbr)
        {
            om->fbr_unsat[i] = (yfu[i] >= 0.0) ? yfu[i] : 0.0;
            om->fbr_gw[i] = (yfg[i] >= 0.0) ? yfg[i] : 0.0;
        }
    }
    for (i = 0; i < nr; i++)                       /* ode.c:59-74 */
    {
        om->stage[i] = (ystg[i] >= 0.0) ? ystg[i] : 0.0;
        om->rgw[i] = (yrgw[i] >= 0.0) ? yrgw[i] : 0.0;
        RFLX(UP_C2C, i) = 0.0;
        RFLX(UP_A2A, i) = 0.0;
    }

    /* Hydrol, hydrol.c:3-25 */
    for (i = 0; i < ne; i++) om->surfh[i] = surf_h(om->surf[i]);
    for (i = 0; i < ne; i++)                       /* EtExtract, hydrol.c:51-87 */
    {
        double edir = F(PB_F_EDIR, i), ett = F(PB_F_ETT, i);
        double depth = E(PB_E_DEPTH, i);

        if (om->surfh[i] >= DEPRSTG)
        {
            X(PB_X_EDIR_SURF, i) = edir; X(PB_X_EDIR_UNSAT, i) = 0.0; X(PB_X_EDIR_GW, i) = 0.0;
        }
        else if (om->gw[i] > depth - E(PB_E_DINF, i))
        {
            X(PB_X_EDIR_SURF, i) = 0.0; X(PB_X_EDIR_UNSAT, i) = 0.0; X(PB_X_EDIR_GW, i) = edir;
        }
        else
        {
            X(PB_X_EDIR_SURF, i) = 0.0; X(PB_X_EDIR_UNSAT, i) = edir; X(PB_X_EDIR_GW, i) = 0.0;
        }
        if (om->gw[i] > depth - E(PB_E_RZD, i))
        {
            X(PB_X_ETT_UNSAT, i) = 0.0; X(PB_X_ETT_GW, i) = ett;
        }
        else
        {
            X(PB_X_ETT_UNSAT, i) = ett; X(PB_X_ETT_GW, i) = 0.0;
        }
    }
    lateral_flow(om);
    for (i = 0; i < ne; i++)                       /* VerticalFlow, vert_flow.c:3-31 */
    {
        X(PB_X_INFIL, i) = infil_func(om, i);
        X(PB_X_RECHG, i) = recharge_func(om, i, X(PB_X_INFIL, i));
        if (om->fbr)
        {
            X(PB_X_FBR_INFIL, i) = fbr_infil_func(om, i);
            X(PB_X_FBR_RECHG, i) = fbr_recharge_func(om, i, X(PB_X_FBR_INFIL, i));
        }
    }
    river_flow(om);

    for (i = 0; i < ne; i++)                       /* ode.c:108-206 */
    {
        double area = E(PB_E_AREA, i);

        dsurf[i] += F(PB_F_PCPDRP, i) - X(PB_X_INFIL, i) - X(PB_X_EDIR_SURF, i);
        dunsat[i] += X(PB_X_INFIL, i) - X(PB_X_RECHG, i) - X(PB_X_EDIR_UNSAT, i) -
            X(PB_X_ETT_UNSAT, i);
        dgw[i] += X(PB_X_RECHG, i) - X(PB_X_EDIR_GW, i) - X(PB_X_ETT_GW, i);
        if (om->fbr)
        {
            dgw[i] -= X(PB_X_FBR_INFIL, i);
            dfu[i] += X(PB_X_FBR_INFIL, i) - X(PB_X_FBR_RECHG, i);
            dfg[i] += X(PB_X_FBR_RECHG, i);
        }
        for (j = 0; j < 3; j++)
        {
            dsurf[i] -= X(PB_X_OVL0 + j, i) / area;
            dgw[i] -= X(PB_X_SUB0 + j, i) / area;
            if (om->fbr) dfg[i] -= X(PB_X_FBRFLOW0 + j, i) / area;
        }
        dunsat[i] /= E(PB_E_POROSITY, i);
        dgw[i] /= E(PB_E_POROSITY, i);
        if (om->fbr)
        {
            dfu[i] /= E(PB_E_GPOROSITY, i);
            dfg[i] /= E(PB_E_GPOROSITY, i);
            if (isnan(dfu[i]) || isnan(dfg[i])) nan_flag = 1;
        }
        if (isnan(dsurf[i]) || isnan(dunsat[i]) || isnan(dgw[i])) nan_flag = 1;
    }
    for (i = 0; i < nr; i++)                       /* ode.c:228-297 */
    {
        double area = R(PB_R_AREA, i);

        for (j = 0; j <= 6; j++) dstg[i] -= RFLX(j, i) / area;
        drgw[i] += -RFLX(LEFT_A2A, i) - RFLX(RIGHT_A2A, i) - RFLX(DOWN_A2A, i) -
            RFLX(UP_A2A, i) + RFLX(CHANL_LKG, i);
        drgw[i] /= R(PB_R_POROSITY, i) * area;
        if (isnan(dstg[i]) || isnan(drgw[i])) nan_flag = 1;
    }
    return nan_flag;
}

/* ---- serial N_Vector arithmetic (cvode/src/nvec_ser/nvector_serial.c) ---- */

/* N_VLinearSum with its special cases, nvector_serial.c:421-506:
 * the a==1/b==1/-1 and a==b / a==-b branches round differently from a*x+b*y */
void oracle_nv_linearsum(int64_t n, double a, const double *x, double b,
    const double *y, double *z)
{
    int64_t i;
    double c;
    const double *v1, *v2;

    if (b == 1.0 && z == y) { for (i = 0; i < n; i++) z[i] += a * x[i]; return; }   /* Vaxpy :1007 */
    if (a == 1.0 && z == x) { for (i = 0; i < n; i++) z[i] += b * y[i]; return; }
    if (a == 1.0 && b == 1.0) { for (i = 0; i < n; i++) z[i] = x[i] + y[i]; return; }
    if ((a == 1.0 && b == -1.0) || (a == -1.0 && b == 1.0))
    {
        v1 = (a == 1.0) ? y : x; v2 = (a == 1.0) ? x : y;
        for (i = 0; i < n; i++) z[i] = v2[i] - v1[i];
        return;
    }
    if (a == 1.0 || b == 1.0)
    {
        c = (a == 1.0) ? b : a; v1 = (a == 1.0) ? y : x; v2 = (a == 1.0) ? x : y;
        for (i = 0; i < n; i++) z[i] = c * v1[i] + v2[i];
        return;
    }
    if (a == -1.0 || b == -1.0)
    {
        c = (a == -1.0) ? b : a; v1 = (a == -1.0) ? y : x; v2 = (a == -1.0) ? x : y;
        for (i = 0; i < n; i++) z[i] = c * v1[i] - v2[i];
        return;
    }
    if (a == b) { for (i = 0; i < n; i++) z[i] = a * (x[i] + y[i]); return; }
    if (a == -b) { for (i = 0; i < n; i++) z[i] = a * (x[i] - y[i]); return; }
    for (i = 0; i < n; i++) z[i] = a * x[i] + b * y[i];
}

/* N_VScale special cases, nvector_serial.c:558-584 */
void oracle_nv_scale(int64_t n, double c, const double *x, double *z)
{
    int64_t i;

    if (z == x) { for (i = 0; i < n; i++) z[i] *= c; return; }
    if (c == 1.0) { for (i = 0; i < n; i++) z[i] = x[i]; return; }
    if (c == -1.0) { for (i = 0; i < n; i++) z[i] = -x[i]; return; }
    for (i = 0; i < n; i++) z[i] = c * x[i];
}

double oracle_nv_dotprod(int64_t n, const double *x, const double *y) /* :637-651 */
{
    int64_t i; double s = 0.0;
    for (i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}

double oracle_nv_wrmsnorm(int64_t n, const double *x, const double *w) /* :669-686 */
{
    int64_t i; double s = 0.0, p;
    for (i = 0; i < n; i++) { p = x[i] * w[i]; s += p * p; }
    return sqrt(s / n);
}

double oracle_nv_maxnorm(int64_t n, const double *x)                   /* :653-667 */
{
    int64_t i; double m = 0.0;
    for (i = 0; i < n; i++) if (fabs(x[i]) > m) m = fabs(x[i]);
    return m;
}

double oracle_nv_min(int64_t n, const double *x)                       /* :709-725 */
{
    int64_t i; double m = x[0];
    for (i = 1; i < n; i++) if (x[i] < m) m = x[i];
    return m;
}
