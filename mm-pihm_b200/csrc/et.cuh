// Forcing scatter + interception / snow / evapotranspiration on the device
// (SURVEY 8(f) f2).  Written from the formulas of the reference:
//   ApplyMeteoForc / ApplyLai per-element loops  src/forcing.c:134-160, 242-258
//   IntcpSnowEt                                  src/is_sm_et.c:4-225
// One thread per owned element, run once per ctrl.etstep (every 15 model steps in the
// example project), so it is written for parity, not for speed: the expression order of
// the reference is kept, FMA contraction is off for the whole library, and exp / log /
// pow / cos are the libdevice ones (<= 2 ulp from glibc's, like pow in the RHS).
#pragma once
#include "common.cuh"

namespace pb {

struct EtStepDev {
    double stepsize, cal_edir, cal_ec, cal_ett, meltf;
    int nmeteo, nlai, nlc;
    const double *meteo;   // [nmeteo][PIHM_B200_NUM_METEO_VAR]
    const double *lai;     // [nlai]
    const double *lai_lc;  // [nlc]
    const double *z0_lc;   // [nlc]
};

// tile slot / dictionary entry of an element (same layout macros as rhs.cuh)
#define TSC(slot, i) (m.es[((size_t)((i) >> 5) * TS_NCOL + (slot)) * PB_TILE + ((i) & 31)])
#ifndef CLE
#define CLE(c, i) (m.cls[(size_t)m.cid[i] * CC_STRIDE + (c)])
#endif
#define ETC(c, i) (etf[(size_t)(c) * m.nes + (i)])
#define ETO(c, i) (out[(size_t)(c) * m.nes + (i)])

static __global__ void __launch_bounds__(128)
k_intcp_snow_et(const DevMesh m, const EtStepDev st, const double *__restrict__ etf,
                const int *__restrict__ eti, const double *__restrict__ y, double *__restrict__ out,
                double *__restrict__ ft)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nown) return;
    // pihm_const.h:8-14
    const double CP = 1004.0, LVH2O = 2.501e6, SIGMA = 5.67e-8, RD = 287.04, RV = 461.5;
    const double TSNOW = -3.0, TRAIN = 1.0, T0 = 0.0;          // is_sm_et.c:8-10
    const double stepsize = st.stepsize;

    // ApplyMeteoForc, forcing.c:134-160 (the non-Noah assignments)
    const int ind = eti[(size_t)PB_ETI_METEO_TYPE * m.nes + i] - 1;
    const double *mv = st.meteo + (size_t)ind * PIHM_B200_NUM_METEO_VAR;
    const double prcp = mv[0] / 1000.0;
    const double sfctmp_k = mv[1];
    const double rh_pct = mv[2];
    const double sfcspd = mv[3];
    double soldn = mv[4];
    soldn = (soldn > 0.0) ? soldn : 0.0;
    // ApplyLai, forcing.c:242-258 and is_sm_et.c:58-65 (same value either way)
    const int lai_type = eti[(size_t)PB_ETI_LAI_TYPE * m.nes + i];
    const int lc = eti[(size_t)PB_ETI_LC_TYPE * m.nes + i] - 1;
    const double lai = (lai_type > 0) ? st.lai[lai_type - 1] : st.lai_lc[lc];

    const double shdfac = ETC(PB_ET_SHDFAC, i), cfactr = ETC(PB_ET_CFACTR, i);
    const double depth = TSC(TS_DEPTH, i), rzd = CLE(CC_RZD, i), porosity = CLE(CC_POROSITY, i);
    const double ws_unsat = y[m.o_unsat + i], ws_gw = y[m.o_gw + i];     // elem.ws after Summary (update.c:19-21)
    double sneqv = ETO(PB_EO_SNEQV, i), cmc = ETO(PB_EO_CMC, i);

    // is_sm_et.c:43-55
    const double albedo = 0.5 * (ETC(PB_ET_ALBEDOMIN, i) + ETC(PB_ET_ALBEDOMAX, i));
    const double radnet = soldn * (1.0 - albedo);
    const double sfctmp = sfctmp_k - 273.15;
    const double wind = sfcspd;
    const double rh = rh_pct / 100.0;
    const double vp = 611.2 * exp(17.67 * sfctmp / (sfctmp + 243.5)) * rh;
    const double pres = 101.325 * 1.0e3 * pow((293.0 - 0.0065 * TSC(TS_ZMAX, i)) / 293.0, 5.26);
    const double qv = 0.622 * vp / pres;
    const double qvsat = 0.622 * (vp / rh) / pres;
    const double meltf = st.meltf;

    // snow accumulation and melt, is_sm_et.c:69-87
    const double frac_snow = (sfctmp < TSNOW) ? 1.0 : ((sfctmp > TRAIN) ? 0.0 : (TRAIN - sfctmp) / (TRAIN - TSNOW));
    const double snow_rate = frac_snow * prcp;
    sneqv += snow_rate * stepsize;
    double melt_rate = (sfctmp > T0) ? (sfctmp - T0) * meltf : 0.0;
    if (sneqv > melt_rate * stepsize) {
        sneqv -= melt_rate * stepsize;
    } else {
        melt_rate = sneqv / stepsize;
        sneqv = 0.0;
    }

    // is_sm_et.c:89-106
    const double intcp_max = ETC(PB_ET_CMCFACTR, i) * lai * shdfac;
    const double z0 = st.z0_lc[lc];
    const double zlvl = ETC(PB_ET_ZLVL_WIND, i);
    const double ra = log(zlvl / z0) * log(10.0 * zlvl / z0) / (wind * 0.16);
    const double gamma = 4.0 * 0.7 * SIGMA * RD / CP * pow(sfctmp + 273.15, 4.0) / (pres / ra) + 1.0;
    const double delta = LVH2O * LVH2O * 0.622 / RV / CP / pow(sfctmp + 273.15, 2.0) * qvsat;
    const double etp = (radnet * delta + gamma * (1.2 * LVH2O * (qvsat - qv) / ra)) /
        (1000.0 * LVH2O * (delta + gamma));

    // is_sm_et.c:108-131
    double satn;
    if (depth - ws_gw < rzd) {
        satn = 1.0;
    } else {
        const double ratio = ws_unsat / (depth - ws_gw);
        satn = (ratio > 1.0) ? 1.0 : ((ratio < 0.0) ? 0.0 : 0.5 * (1.0 - cos(3.14 * ratio)));
    }
    const double smcwlt = ETC(PB_ET_SMCWLT, i);
    double betas = (satn * porosity + ETC(PB_ET_SMCMIN, i) - smcwlt) / (ETC(PB_ET_SMCREF, i) - smcwlt);
    betas = (betas < 0.0001) ? 0.0001 : ((betas > 1.0) ? 1.0 : betas);
    double edir = (1.0 - shdfac) * pow(betas, 2.0) * etp;
    edir *= st.cal_edir;
    edir = (edir < 0.0) ? 0.0 : edir;

    // is_sm_et.c:133-184
    double ec, ett, drip;
    if (lai > 0.0) {
        const double cmc_c = (cmc < 0.0) ? 0.0 : ((cmc > intcp_max) ? intcp_max : cmc);
        ec = shdfac * pow(cmc_c / intcp_max, cfactr) * etp;
        ec *= st.cal_ec;
        ec = (ec < 0.0) ? 0.0 : ec;

        const double rsmin = ETC(PB_ET_RSMIN, i), rsmax = ETC(PB_ET_RSMAX, i);
        double fr = 1.1 * radnet / (ETC(PB_ET_RGL, i) * lai);
        fr = (fr < 0.0) ? 0.0 : fr;
        double alphar = (1.0 + fr) / (fr + (rsmin / rsmax));
        alphar = (alphar > 10000.0) ? 10000.0 : alphar;
        double etas = 1.0 - 0.0016 * (pow((ETC(PB_ET_TOPT, i) - 273.15 - sfctmp), 2.0));
        etas = (etas < 0.0001) ? 0.0001 : etas;
        double gammas = 1.0 / (1.0 + 0.00025 * (vp / rh - vp));
        gammas = (gammas < 0.01) ? 0.01 : gammas;
        double rs = rsmin * alphar / (betas * lai * etas * gammas);
        rs = (rs > rsmax) ? rsmax : rs;
        const double pc = (1.0 + delta / gamma) / (1.0 + rs / ra + delta / gamma);

        // (is_sm_et.c:163-167 as parsed: the quotient sits inside the else branch of cmc < 0)
        const double cfrac = (cmc < 0.0) ? 0.0 : ((cmc > intcp_max) ? intcp_max : cmc) / intcp_max;
        ett = shdfac * pc * (1.0 - pow(cfrac, cfactr)) * etp;
        ett *= st.cal_ett;
        ett = (ett < 0.0) ? 0.0 : ett;
        ett = ((ws_gw < (depth - rzd)) && ws_unsat <= 0.0) ? 0.0 : ett;

        drip = (cmc <= 0.0) ? 0.0 : 6.52E-7 * intcp_max * exp(3.89 * cmc / intcp_max);
    } else {
        ett = 0.0;
        ec = 0.0;
        drip = 0.0;
    }

    // is_sm_et.c:186-222
    if (drip < 0.0) drip = 0.0;
    if (drip * stepsize > cmc) drip = cmc / stepsize;
    const double isval = cmc + (1.0 - frac_snow) * prcp * shdfac * stepsize - ec * stepsize - drip * stepsize;
    if (isval > intcp_max) {
        cmc = intcp_max;
        drip += (isval - intcp_max) / stepsize;
    } else if (isval < 0.0) {
        cmc = 0.0;
        if (ec + drip > 0.0) {
            ec = ec / (ec + drip) * (cmc + (1.0 - frac_snow) * prcp * shdfac * stepsize);
            drip = drip / (ec + drip) * (cmc + (1.0 - frac_snow) * prcp * shdfac * stepsize);   // the updated ec, as written
        }
    } else {
        cmc = isval;
    }
    const double pcpdrp = (1.0 - shdfac) * (1.0 - frac_snow) * prcp + drip + melt_rate;

    ETO(PB_EO_PCPDRP, i) = pcpdrp;
    ETO(PB_EO_EDIR, i) = edir;
    ETO(PB_EO_ETT, i) = ett;
    ETO(PB_EO_EC, i) = ec;
    ETO(PB_EO_DRIP, i) = drip;
    ETO(PB_EO_SNEQV, i) = sneqv;
    ETO(PB_EO_CMC, i) = cmc;
    // the three forcing columns k_main reads (tile layout, common.cuh FTC)
    double *f = ft + (size_t)(i >> 5) * 4 * PB_TILE + (i & 31);
    f[PB_F_PCPDRP * PB_TILE] = pcpdrp;
    f[PB_F_EDIR * PB_TILE] = edir;
    f[PB_F_ETT * PB_TILE] = ett;
}

#undef TSC
#undef ETC
#undef ETO
}  // namespace pb
