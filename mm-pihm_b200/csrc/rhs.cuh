// rhs.cuh -- the MM-PIHM right-hand side as two sm_100a FP64 kernels.
//
//   k_pre  : one thread per element  -> surfh, EffKh, friction slope |grad h|
//            one thread per river    -> all river fluxes except the upstream sums
//   k_main : one thread per element  -> lateral + vertical fluxes, dy
//            one thread per river    -> ordered upstream accumulation, dy
//
// The split is forced by the data flow of the reference: the
// overland flux of edge (i,n) needs |grad h| of BOTH elements, and |grad h| of
// n needs the surface heads of n's neighbours; river->element fluxes are
// computed per river and consumed per element.
//
// Arithmetic follows the reference expression by expression (file:line cited
// at each function) and this translation unit is compiled with -fmad=false, so
// + - * / sqrt round exactly like the reference's x86-64 build; only pow/log
// (libdevice vs glibc) can differ, by <= 2 ulp.
#pragma once
#include "common.cuh"
#include "rhs_layout.cuh"
#include "fastpow.cuh"
#include "fdiv.cuh"

namespace pb {

// global-memory access to a tile slot of an arbitrary element (river kernels, fbr gathers)
#define TSC(slot, i) (m.es[((size_t)((i) >> 5) * TS_NCOL + (slot)) * PB_TILE + ((i) & 31)])
// dictionary entry c of an arbitrary element (river kernels, fbr gathers)
#define CLE(c, i) (m.cls[(size_t)m.cid[i] * CC_STRIDE + (c)])
#define FTC(c, i) (m.ft[((size_t)((i) >> 5) * 4 + (c)) * PB_TILE + ((i) & 31)])

// ---- TMA bulk copy + mbarrier (PTX; no CUTLASS dependency) ---------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// more bytes for the current phase, without an arrival (the static part of a stage requested
// in the prologue, before the dependency wait; the arrival comes with the dynamic part)
__device__ __forceinline__ void mbar_expect_tx_only(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// pull a contiguous range into L2 (no destination): the rows other tiles are about to gather from
__device__ __forceinline__ void tma_prefetch_l2(const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase)
{
    asm volatile("{\n .reg .pred P1;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 " @P1 bra DONE;\n bra LAB_WAIT;\n DONE:\n }" ::"r"(bar), "r"(phase) : "memory");
}

// dictionary fetches pinned at their point of use (asm volatile): the rows are L1-resident, so
// the ~40 cycles are cheaper than holding a dozen doubles in registers from the top of the kernel
__device__ __forceinline__ double2 ldg2_here(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Scheduling fence without instructions: the value x may not be consumed before `dep` has been
// computed.  Warps issue in order, so the first consumer of a gathered value stalls the warp
// for the rest of the gather's latency even when independent arithmetic follows in the
// instruction stream; tying the gathered registers to the result of the element's own-state
// chain (van Genuchten pows / the overland power) keeps that chain in front of them.
#ifndef PB_NO_ORDER
#define PB_AFTER(x, dep) asm volatile("" : "+d"(x) : "d"(dep))
#else
#define PB_AFTER(x, dep)
#endif

// Neighbour gathers pinned at their point of issue (asm volatile keeps its order relative to the
// other volatile statements, e.g. the dictionary fetches that start the own-state chains): all
// of a tile's gathers leave at the top of the tile instead of being sunk towards their consumers.
#ifndef PB_NO_ORDER
__device__ __forceinline__ double ld_here(const double *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double4 ld4_here(const double4 *p)
{
    double4 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2+16];" : "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
#else
__device__ __forceinline__ double ld_here(const double *p) { return *p; }
__device__ __forceinline__ double4 ld4_here(const double4 *p) { return *p; }
#endif
// Per-lane asynchronous gathers (LDGSTS): global -> this warp's gather buffer in shared memory,
// no registers held while they are in flight; completion per thread (cp.async.wait_group).
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool GH> __device__ __forceinline__ const double *p_surf(const DevMesh &m, const double *y, int i)
{
    return (!GH || i < m.nown) ? y + i : m.gel + (size_t)(i - m.nown) * m.gs;
}
template <bool GH> __device__ __forceinline__ const double *p_gw(const DevMesh &m, const double *y, int i)
{
    return (!GH || i < m.nown) ? y + m.o_gw + i : m.gel + (size_t)(i - m.nown) * m.gs + 1;
}
template <bool GH> __device__ __forceinline__ const double *p_fg(const DevMesh &m, const double *y, int i)
{
    return (!GH || i < m.nown) ? y + m.o_fg + i : m.gel + (size_t)(i - m.nown) * m.gs + 2;
}

#define RFC(c, r) (m.rf[(size_t)(c) * m.nrs + (r)])
#define RIC(c, r) (m.ri[(size_t)(c) * m.nrs + (r)])
#define FOC(c, i) (m.forc[(size_t)(c) * m.nes + (i)])   /* bc columns only (rare path) */
#define RFLX(k, r) (m.rivflow[(size_t)(k) * m.nrs + (r)])
#define XFC(c, i) (m.xflux[(size_t)(c) * m.nes + (i)])

__device__ __forceinline__ double max0(double v) { return (v >= 0.0) ? v : 0.0; }

// State access in a partitioned run: local entities [0, nown) / [0, rown) are
// the rank's own unknowns inside y; the rest are ghosts whose values arrived
// with the halo exchange as records {surf, gw[, fbr_gw]} / {stage, gw}.
// Single-GPU: nown == ne, rown == nr, the ghost branch is never taken.
template <bool GH> __device__ __forceinline__ double y_surf(const DevMesh &m, const double *__restrict__ y, int i)
{
    return (!GH || i < m.nown) ? y[i] : m.gel[(size_t)(i - m.nown) * m.gs];
}
template <bool GH> __device__ __forceinline__ double y_gw(const DevMesh &m, const double *__restrict__ y, int i)
{
    return (!GH || i < m.nown) ? y[m.o_gw + i] : m.gel[(size_t)(i - m.nown) * m.gs + 1];
}
template <bool GH> __device__ __forceinline__ double y_fg(const DevMesh &m, const double *__restrict__ y, int i)
{
    return (!GH || i < m.nown) ? y[m.o_fg + i] : m.gel[(size_t)(i - m.nown) * m.gs + 2];
}
template <bool GH> __device__ __forceinline__ double y_stage(const DevMesh &m, const double *__restrict__ y, int r)
{
    return (!GH || r < m.rown) ? y[m.o_stg + r] : m.gri[(size_t)(r - m.rown) * 2];
}
template <bool GH> __device__ __forceinline__ double y_rgw(const DevMesh &m, const double *__restrict__ y, int r)
{
    return (!GH || r < m.rown) ? y[m.o_rgw + r] : m.gri[(size_t)(r - m.rown) * 2 + 1];
}

// a / b for a finite b > 0.  IEEE gives 0/b = 0 with the sign of a, i.e. a
// itself; taking that shortcut in a branch keeps zero numerators (dry edges,
// dry surface) out of the FP64 division's denormal/zero slow path, which a
// whole warp otherwise pays for (ncu: 21 % of k_main's instructions).
// a, or 1.0 where a == 0.  Written as opaque PTX: in C++ the optimiser sees
// that the quotient of the substituted lanes is discarded and divides the raw
// zero again (checked in the SASS).
__device__ __forceinline__ double nonzero_or_one(double a)
{
    double r;
    asm("{ .reg .pred p; setp.neu.f64 p, %1, 0d0000000000000000; selp.f64 %0, %1, 0d3FF0000000000000, p; }"
        : "=d"(r) : "d"(a));
    return r;
}
__device__ __forceinline__ double div_pos(double a, double b)
{
    // branch-free on purpose: the compiler if-converts a guard anyway, and then
    // the zero lanes would still send the warp into __cuda_sm20_div_rn_f64_full.
    // Zero lanes divide 1.0 instead (fast path) and the result is discarded.
    const bool nz = (a != 0.0);
    const double q = nonzero_or_one(a) / b;
    return nz ? q : a;
}
// general divisor (either sign, finite, non-zero): 0/b = a*b in sign and value
__device__ __forceinline__ double div_any(double a, double b)
{
    const bool plain = (a != 0.0) || !(fabs(b) > 0.0) || isinf(b);
    const double q = (plain ? a : nonzero_or_one(a)) / b;
    return plain ? q : a * b;
}

// SurfH, src/hydrol.c:92-126
__device__ __forceinline__ double surf_h(double surfeqv)
{
    if (surfeqv < 0.0) return 0.0;
    if (surfeqv <= 0.5 * PB_DEPRSTG) return sqrt(2.0 * PB_DEPRSTG * surfeqv);
    return PB_DEPRSTG + (surfeqv - 0.5 * PB_DEPRSTG);
}

// AvgHsurf, src/lat_flow.c:175-203
__device__ __forceinline__ double avg_hsurf(double diff, double hsurf, double hnabr)
{
    if (diff > 0.0) return (hsurf > PB_DEPRSTG) ? 1.0 * (hsurf - PB_DEPRSTG) : 0.0;
    return (hnabr > PB_DEPRSTG) ? 1.0 * (hnabr - PB_DEPRSTG) : 0.0;
}

// AvgH, src/lat_flow.c:205-225
__device__ __forceinline__ double avg_h(double diff, double hsub, double hnabr)
{
    double a = 0.0;
    if (diff > 0.0) { if (hsub > 0.0) a = hsub; }
    else { if (hnabr > 0.0) a = hnabr; }
    return a;
}

// AvgH for heads that are already clamped at zero (max0): the inner tests are no-ops then
__device__ __forceinline__ double avg_h_nn(double diff, double hsub, double hnabr)
{
#ifdef PB_AVGH_GENERIC
    return avg_h(diff, hsub, hnabr);
#else
    return (diff > 0.0) ? hsub : hnabr;
#endif
}

// DhByDl, src/lat_flow.c:227-234
__device__ __forceinline__ double dh_by_dl(const double *l1, const double *l2, const double *h)
{
    return div_any(-1.0 *
        (l1[2] * (h[1] - h[0]) + l1[1] * (h[0] - h[2]) + l1[0] * (h[2] - h[1])),
        (l2[2] * (l1[1] - l1[0]) + l2[1] * (l1[0] - l1[2]) + l2[0] * (l1[2] - l1[1])));
}

// EffKh, src/lat_flow.c:236-265
__device__ __forceinline__ double eff_kh(double depth, double dmac, double kmach,
                                         double areafv, double ksath, double gw)
{
    gw = (gw > 0.0) ? gw : 0.0;
    if (gw > depth - dmac) {
        double k1 = kmach * areafv + ksath * (1.0 - areafv);
        double k2 = ksath;
        double d1, d2;
        if (gw > depth) { d1 = dmac; d2 = depth - dmac; }
        else { d1 = gw - (depth - dmac); d2 = depth - dmac; }
        return (k1 * d1 + k2 * d2) / (d1 + d2);
    }
    return ksath;
}

__device__ __forceinline__ double eff_kh_elem(const DevMesh &m, int e, double gw)
{
    const double *row = m.cls + (size_t)m.cid[e] * CC_STRIDE;
    return eff_kh(TSC(TS_DEPTH, e), TSC(TS_DMAC, e), row[CC_KMACH], row[CC_AREAFV], row[CC_KSATH], gw);
}

// OverLandFlow, src/lat_flow.c:267-271.  pow(0, 0.6666667) == 0 exactly, so the
// libdevice call is skipped for dry edges (bitwise the same product).
__device__ __forceinline__ double overland_flow(double avgh, double grad, double sf,
                                                double crossa, double rough)
{
    double p = (avgh == 0.0) ? 0.0 : pow_pos(avgh, 0.6666667);
    return div_pos(crossa * p * grad, sqrt(sf) * rough);
}

// KrFunc, src/soil.c:3-8.  The reference spells the squared factor out twice;
// (sqrt(s) * a) * a is the same left-to-right product.
__device__ __forceinline__ double kr_func(double beta, double satn)
{
    double a = 1.0 - pow_pos(1.0 - pow_pos(satn, beta / (beta - 1.0)), (beta - 1.0) / beta);
    return sqrt(satn) * a * a;
}

// KrFunc(beta, satn) and Psi(satn, alpha, beta) (vert_flow.c:272-278) of the same
// element: the four pow() form two independent pairs, pow(s, m) / pow(1/s, m) and
// pow(1 - A, (b-1)/b) / pow(C - 1, 1/b), each evaluated as one interleaved
// straight-line block (pow_pos2).  Values are those of kr_func() / psi_func().
// m1 = beta/(beta-1), m2 = (beta-1)/beta, m3 = 1/beta come from the class dictionary.
__device__ __forceinline__ void vg_kr_psi(double satn, double alpha, double m1, double m2, double m3,
                                          double &kr, double &psi)
{
    const double sp = (satn < PB_SATMIN) ? PB_SATMIN : satn;
    double A, C, B, D;
    pow_pos2(satn, m1, 1.0 / sp, m1, A, C);
    pow_pos2(1.0 - A, m2, C - 1.0, m3, B, D);
    const double a = 1.0 - B;
    kr = sqrt(satn) * a * a;
    psi = -D / alpha;
}

// KrFunc(BETA_CRACK = 2.0, s), src/vert_flow.c:258: both exponents are exact
// (2.0 and 0.5), where a correctly rounded pow equals s*s and sqrt().
__device__ __forceinline__ double kr_func_crack(double satn)
{
    double a = 1.0 - sqrt(1.0 - satn * satn);
    return sqrt(satn) * a * a;
}

// Psi, src/vert_flow.c:272-278
__device__ __forceinline__ double psi_func(double satn, double alpha, double beta)
{
    satn = (satn < PB_SATMIN) ? PB_SATMIN : satn;
    return -pow_pos(pow_pos(1.0 / satn, beta / (beta - 1.0)) - 1.0, 1.0 / beta) / alpha;
}

// EffKinf, src/vert_flow.c:211-270
__device__ __forceinline__ double eff_kinf(double kinfv, double kmacv, double areafh,
                                           double dh_by_dz, double ksatfunc, double elemsatn,
                                           double applrate, double surfh)
{
    if (areafh == 0.0) return kinfv * ksatfunc;
    if (surfh > PB_DEPRSTG) return kinfv * (1.0 - areafh) * ksatfunc + kmacv * areafh;
    if (applrate <= dh_by_dz * kinfv * ksatfunc) return kinfv * ksatfunc;
    double kmax = dh_by_dz * (kmacv * areafh + kinfv * (1.0 - areafh) * ksatfunc);
    if (applrate < kmax)
        return kinfv * (1.0 - areafh) * ksatfunc + kmacv * areafh * kr_func_crack(elemsatn);
    return kinfv * (1.0 - areafh) * ksatfunc + kmacv * areafh;
}

// RiverCroSectArea, src/river_flow.c:518-546
__device__ __forceinline__ double riv_area(int order, double depth, double coeff)
{
    depth = (depth > 0.0) ? depth : 0.0;
    switch (order) {
        case 1: return depth * coeff;
        case 2: return depth * depth / coeff;
        case 3: return 4.0 * depth * sqrt(depth) / (3.0 * sqrt(coeff));
        case 4: return 3.0 * pow(depth, 4.0 / 3.0) / (2.0 * pow(coeff, 1.0 / 3.0));
    }
    return 0.0;
}

// RiverPerim, src/river_flow.c:548-581
__device__ __forceinline__ double riv_perim(int order, double depth, double coeff)
{
    depth = (depth > 0.0) ? depth : 0.0;
    switch (order) {
        case 1: return 2.0 * depth + coeff;
        case 2: return 2.0 * depth * sqrt(1.0 + coeff * coeff) / coeff;
        case 3:
            return sqrt(depth * (1.0 + 4.0 * coeff * depth) / coeff) +
                log(2.0 * sqrt(coeff * depth) + sqrt(1.0 + 4.0 * coeff * depth)) / (2.0 * coeff);
        case 4:
            return 2.0 * ((pow(depth * (1.0 + 9.0 * pow(coeff, 2.0 / 3.0) * depth), 0.5) / 3.0) +
                (log(3.0 * pow(coeff, 1.0 / 3.0) * sqrt(depth) +
                     pow(1.0 + 9.0 * pow(coeff, 2.0 / 3.0) * depth, 0.5)) /
                 (9.0 * pow(coeff, 1.0 / 3.0))));
    }
    return 0.0;
}

// river-side view of one bank element
struct Bank {
    double surfh, gw, zmax, zmin, effk;
};

template <bool GH> __device__ __forceinline__ Bank load_bank(const DevMesh &m, const double *__restrict__ y, int e)
{
    Bank b;
    b.surfh = surf_h(max0(y_surf<GH>(m, y, e)));
    b.gw = max0(y_gw<GH>(m, y, e));
    b.zmax = TSC(TS_ZMAX, e);
    b.zmin = TSC(TS_ZMIN, e);
    b.effk = eff_kh_elem(m, e, b.gw);
    return b;
}

// OvlFlowElemToRiver, src/river_flow.c:183-253
__device__ __forceinline__ double ovl_elem_to_river(const Bank &b, double rzmax, double zbed,
                                                    double stage, double cwr, double len)
{
    double zbank = (rzmax > b.zmax) ? rzmax : b.zmax;
    double elem_h = b.zmax + b.surfh;
    double rivseg_h = zbed + stage;
    double flux;
    if (rivseg_h > elem_h) {
        if (elem_h > zbank)
            flux = cwr * 2.0 * sqrt(2.0 * PB_GRAV) * len * sqrt(rivseg_h - elem_h) * (rivseg_h - zbank) / 3.0;
        else if (zbank < rivseg_h)
            flux = cwr * 2.0 * sqrt(2.0 * PB_GRAV) * len * sqrt(rivseg_h - zbank) * (rivseg_h - zbank) / 3.0;
        else
            flux = 0.0;
    } else if (b.surfh > PB_DEPRSTG) {
        if (rivseg_h > zbank)
            flux = -cwr * 2.0 * sqrt(2.0 * PB_GRAV) * len * sqrt(elem_h - rivseg_h) * (elem_h - zbank) / 3.0;
        else if (zbank < elem_h)
            flux = -cwr * 2.0 * sqrt(2.0 * PB_GRAV) * len * sqrt(elem_h - zbank) * (elem_h - zbank) / 3.0;
        else
            flux = 0.0;
    } else {
        flux = 0.0;
    }
    return flux;
}

// ChanFlowElemToRiver, src/river_flow.c:427-458
__device__ __forceinline__ double chan_elem_to_river(const Bank &b, double zbed, double stage,
                                                     double rksath, double len, double distance)
{
    double diff_h = (stage + zbed) - (b.gw + b.zmin);
    double avgh;
    if (b.zmin > zbed) avgh = b.gw;
    else if (b.zmin + b.gw > zbed) avgh = b.zmin + b.gw - zbed;
    else avgh = 0.0;
    avgh = avg_h(diff_h, stage, avgh);
    double grad_h = div_pos(diff_h, distance);
    double avg_ksat = 0.5 * (b.effk + rksath);
    return len * avg_ksat * grad_h * avgh;
}

// SubFlowElemToRiver, src/river_flow.c:460-496 (_ARITH_)
__device__ __forceinline__ double sub_elem_to_river(const Bank &b, double zbed, double rzmin,
                                                    double rgw, double effk_riv, double len,
                                                    double distance)
{
    double diff_h = (rgw + rzmin) - (b.gw + b.zmin);
    double avgh;
    if (b.zmin > zbed) avgh = 0.0;
    else if (b.zmin + b.gw > zbed) avgh = zbed - b.zmin;
    else avgh = b.gw;
    avgh = avg_h(diff_h, rgw, avgh);
    double avg_ksat = 0.5 * (b.effk + effk_riv);
    double grad_h = div_pos(diff_h, distance);
    return len * avg_ksat * grad_h * avgh;
}

// ---------------------------------------------------------------------------
// river part of k_pre: RiverFlow() parallel loop, src/river_flow.c:10-86
// ---------------------------------------------------------------------------
template <bool GH> __device__ __forceinline__ void river_fluxes(const DevMesh &m, const double *__restrict__ y, int r)
{
    const double stage = max0(y_stage<GH>(m, y, r));
    const double rgw = max0(y_rgw<GH>(m, y, r));
    const int down = RIC(PB_RI_DOWN, r);
    const int ord = RIC(PB_RI_INTRPL_ORD, r);
    const double zbed = RFC(PB_R_ZBED, r), rzmin = RFC(PB_R_ZMIN, r), rzmax = RFC(PB_R_ZMAX, r);
    const double len = RFC(PB_R_SHP_LENGTH, r), coeff = RFC(PB_R_SHP_COEFF, r);
    const double rough = RFC(PB_R_ROUGH, r);
    const Bank L = load_bank<GH>(m, y, RIC(PB_RI_LEFTELE, r));
    const Bank R = load_bank<GH>(m, y, RIC(PB_RI_RIGHTELE, r));
    double rf_up = 0.0, rf_down, rf_down_a2a;

    // the previous call's bank overland fluxes stay visible to Infil (H2)
    // (a re-evaluation of the last call for its fluxes sees what that call saw)
    if (!m.replay) {
        m.s2c_stale[r] = RFLX(RF_LEFT_S2C, r);
        m.s2c_stale[m.nrs + r] = RFLX(RF_RIGHT_S2C, r);
    }

    if (down > 0) {
        const int d = down - 1;
        const int bct = RIC(PB_RI_BCTYPE, r);
        if (bct != 0) {   // BoundFluxRiver, river_flow.c:390-425
            double flux = 0.0;
            if (bct > 0) {
                double total_h = stage + zbed;
                double total_h_down = m.rivbc[r];
                double distance = 0.5 * len;
                double grad_h = (total_h - total_h_down) / distance;
                double avg_perim = riv_perim(ord, stage, coeff);
                double crossa = riv_area(ord, stage, coeff);
                double avgh = (avg_perim == 0.0) ? 0.0 : (crossa / avg_perim);
                flux = overland_flow(avgh, grad_h, grad_h, crossa, rough);
            } else {
                flux = -m.rivbc[r];
            }
            rf_up += flux;
        }
        // ChanFlowRiverToRiver, river_flow.c:255-298
        const double stage_d = max0(y_stage<GH>(m, y, d));
        const double rgw_d = max0(y_rgw<GH>(m, y, d));
        const int ord_d = RIC(PB_RI_INTRPL_ORD, d);
        const double len_d = RFC(PB_R_SHP_LENGTH, d), coeff_d = RFC(PB_R_SHP_COEFF, d);
        {
            double total_h = stage + zbed;
            double perim = riv_perim(ord, stage, coeff);
            double total_h_down = stage_d + RFC(PB_R_ZBED, d);
            double perim_down = riv_perim(ord_d, stage_d, coeff_d);
            double avg_perim = (perim + perim_down) / 2.0;
            double avg_rough = (rough + RFC(PB_R_ROUGH, d)) / 2.0;
            double distance = 0.5 * (len + len_d);
            double diff_h = (m.riv_mode == PB_KINEMATIC) ? (zbed - RFC(PB_R_ZBED, d))
                                                         : (total_h - total_h_down);
            double grad_h = div_pos(diff_h, distance);
            double avg_sf = (grad_h > 0.0) ? grad_h : PB_RIVGRADMIN;
            double crossa = riv_area(ord, stage, coeff);
            double crossa_down = riv_area(ord_d, stage_d, coeff_d);
            double avg_crossa = 0.5 * (crossa + crossa_down);
            double avgh = (avg_perim == 0.0) ? 0.0 : (avg_crossa / avg_perim);
            rf_down = overland_flow(avgh, grad_h, avg_sf, crossa, avg_rough);
        }
        // SubFlowRiverToRiver, river_flow.c:300-329 (_ARITH_)
        {
            const Bank DL = load_bank<GH>(m, y, RIC(PB_RI_LEFTELE, d));
            const Bank DR = load_bank<GH>(m, y, RIC(PB_RI_RIGHTELE, d));
            double effk = 0.5 * (L.effk + R.effk);
            double effk_nabr = 0.5 * (DL.effk + DR.effk);
            double total_h = rgw + rzmin;
            double total_h_down = rgw_d + RFC(PB_R_ZMIN, d);
            double avg_wid = (RFC(PB_R_SHP_WIDTH, r) + RFC(PB_R_SHP_WIDTH, d)) / 2.0;
            double diff_h = total_h - total_h_down;
            double avgh = avg_h(diff_h, rgw, rgw_d);
            double distance = 0.5 * (len + len_d);
            double grad_h = div_pos(diff_h, distance);
            double avg_ksat = 0.5 * (effk + effk_nabr);
            rf_down_a2a = avg_ksat * grad_h * avgh * avg_wid;
        }
    } else {
        // OutletFlux, river_flow.c:331-388
        double discharge;
        switch (down) {
            case -1: {
                double total_h = stage + zbed;
                double total_h_down = m.rivbc[r];
                double distance = 0.5 * len;
                double grad_h = (total_h - total_h_down) / distance;
                double avg_perim = riv_perim(ord, stage, coeff);
                double crossa = riv_area(ord, stage, coeff);
                double avgh = (avg_perim == 0.0) ? 0.0 : (crossa / avg_perim);
                discharge = overland_flow(avgh, grad_h, grad_h, crossa, rough);
                break;
            }
            case -2:
                discharge = -m.rivbc[r];
                break;
            case -3: {
                double distance = 0.5 * len;
                double grad_h = (zbed - (RFC(PB_R_NODE_ZMAX, r) - RFC(PB_R_SHP_DEPTH, r))) / distance;
                double avg_perim = riv_perim(ord, stage, coeff);
                double crossa = riv_area(ord, stage, coeff);
                discharge = sqrt(grad_h) * crossa *
                    ((avg_perim > 0.0) ? pow(crossa / avg_perim, 2.0 / 3.0) : 0.0) / rough;
                break;
            }
            case -4: {
                double crossa = riv_area(ord, stage, coeff);
                discharge = crossa * sqrt(PB_GRAV * stage);
                break;
            }
            default:      // reference exits; rejected at pihm_b200_create
                discharge = 0.0;
        }
        rf_down = discharge;
        rf_down_a2a = 0.0;
    }

    // RiverToElem, river_flow.c:111-181
    const double cwr = RFC(PB_R_CWR, r), rksath = RFC(PB_R_KSATH, r);
    const double dl = RFC(PB_R_DIST_LEFT, r), dr = RFC(PB_R_DIST_RIGHT, r);
    RFLX(RF_UP_C2C, r) = rf_up;
    RFLX(RF_DOWN_C2C, r) = rf_down;
    RFLX(RF_LEFT_S2C, r) = ovl_elem_to_river(L, rzmax, zbed, stage, cwr, len);
    RFLX(RF_RIGHT_S2C, r) = ovl_elem_to_river(R, rzmax, zbed, stage, cwr, len);
    RFLX(RF_LEFT_A2C, r) = chan_elem_to_river(L, zbed, stage, rksath, len, dl);
    RFLX(RF_RIGHT_A2C, r) = chan_elem_to_river(R, zbed, stage, rksath, len, dr);
    const double effk_riv = 0.5 * (L.effk + R.effk);
    RFLX(RF_LEFT_A2A, r) = sub_elem_to_river(L, zbed, rzmin, rgw, effk_riv, len, dl);
    RFLX(RF_RIGHT_A2A, r) = sub_elem_to_river(R, zbed, rzmin, rgw, effk_riv, len, dr);
    RFLX(RF_DOWN_A2A, r) = rf_down_a2a;
    RFLX(RF_UP_A2A, r) = 0.0;
    // ChanLeak, river_flow.c:498-516
    {
        double diff_h;
        if (zbed - (rgw + rzmin) > 0.0) diff_h = stage;
        else diff_h = stage + zbed - (rgw + rzmin);
        double grad_h = div_pos(diff_h, RFC(PB_R_BEDTHICK, r));
        RFLX(RF_CHANL_LKG, r) = RFC(PB_R_KSATV, r) * RFC(PB_R_SHP_WIDTH, r) * len * grad_h;
    }
}

// ---------------------------------------------------------------------------
// element part of k_pre: SurfH (hydrol.c:10-14), EffKh of the own column, and
// FrictSlope (lat_flow.c:118-173) reduced to |grad h| = sqrt(dhbydx^2+dhbydy^2),
// the only form LateralFlow uses it in (lat_flow.c:33-36).  The results go into
// the 32-byte "dynamic neighbour record" {surfh, effkh, |grad h|, (surfh-D)^(2/3)} that
// k_main gathers with one sector per neighbour.
// ---------------------------------------------------------------------------
// SurfH / EffKh / DhByDl through Arith<FAST>: with FAST the whole of elem_pre is one
// straight-line block (selects instead of branches around sqrt and `/`)
template <bool FAST>
__device__ __forceinline__ double surf_h_a(Arith<FAST> &A, double surfeqv)
{
    if (!FAST) return surf_h(surfeqv);
    // the root is needed only for 0 < surfeqv <= DEPRSTG / 2 (sqrt(0) = 0): when no lane of the warp
    // is in that range -- ponded or dry ground -- the whole warp skips it (same bits either way)
    double lo = 0.0;
#ifdef PB_SQRT_VOTE       // (measured: with mixed states the vote costs more than the skipped roots give)
    if (__any_sync(__activemask(), surfeqv > 0.0 && surfeqv <= 0.5 * PB_DEPRSTG))
#endif
        lo = A.sqrtp(2.0 * PB_DEPRSTG * ((surfeqv > 0.0) ? surfeqv : 0.0));
    const double hi = PB_DEPRSTG + (surfeqv - 0.5 * PB_DEPRSTG);
    return (surfeqv < 0.0) ? 0.0 : ((surfeqv <= 0.5 * PB_DEPRSTG) ? lo : hi);
}
template <bool FAST>
__device__ __forceinline__ double eff_kh_a(Arith<FAST> &A, double depth, double dmac, double kmach,
                                           double areafv, double ksath, double gw)
{
    if (!FAST) return eff_kh(depth, dmac, kmach, areafv, ksath, gw);
    gw = (gw > 0.0) ? gw : 0.0;
    const bool mac = gw > depth - dmac;
    const double k1 = kmach * areafv + ksath * (1.0 - areafv);
    const double d2 = depth - dmac;
    const double d1 = (gw > depth) ? dmac : gw - (depth - dmac);
    const double q = A.quor(mac ? k1 * d1 + ksath * d2 : 1.0, mac ? d1 + d2 : 1.0);
    return mac ? q : ksath;
}
template <bool FAST>
__device__ __forceinline__ double dh_by_dl_a(Arith<FAST> &A, const double *l1, const double *l2, const double *h)
{
    if (!FAST) return dh_by_dl(l1, l2, h);
    return A.quor(-1.0 *
        (l1[2] * (h[1] - h[0]) + l1[1] * (h[0] - h[2]) + l1[0] * (h[2] - h[1])),
        (l2[2] * (l1[1] - l1[0]) + l2[1] * (l1[0] - l1[2]) + l2[0] * (l1[2] - l1[1])));
}

// KrFunc(beta, satn) and Psi(satn, alpha, beta) (vert_flow.c:272-278) of the same
// element.  The four pow() form two independent pairs, pow(s, m) / pow(1/s, m) and
// pow(1 - A, (b-1)/b) / pow(C - 1, 1/b); with Arith<true> all of it is one
// straight-line block and the pairs interleave.  Values are those of kr_func() /
// psi_func().
template <bool FAST>
__device__ __forceinline__ void vg_kr_psi_a(Arith<FAST> &A, double satn, double alpha, double r_alpha, double m1,
                                            double m2, double m3, double &kr, double &psi)
{
    if (!FAST) { vg_kr_psi(satn, alpha, m1, m2, m3, kr, psi); return; }
    const double sp = (satn < PB_SATMIN) ? PB_SATMIN : satn;
#if PB_RELAX & 4
    // With A = s^m1 the reference's C = (1/s)^m1 is 1/A, so C - 1 = (1 - A) / A and
    //   B = (1 - A)^m2 = exp(m2 L1),  D = (C - 1)^m3 = exp(m3 (L1 - m1 Ls)),  L1 = log(1 - A), Ls = log s:
    // two logarithms and three exponentials instead of three and four, no 1/s.  (The callers
    // clamp satn to [SATMIN, 1], so sp == satn; satn = 1 gives A = 1, B = D = 0 exactly.)
    {
        const LogDD Ls = A.logp(sp);
        const double Av = A.expy(Ls.H, Ls.Lo, m1);
        const double oma = 1.0 - Av;
        const bool nz = oma > 0.0;
        const LogDD L1 = A.logp(nz ? oma : 1.0);
        const double Bv = A.expy(L1.H, L1.Lo, m2);
        // m1 * Ls as a double-double, then L1 - m1 Ls (two-sum of the heads)
        const double P = __dmul_rn(Ls.H, m1);
        const double Pl = __fma_rn(Ls.Lo, m1, __fma_rn(Ls.H, m1, -P));
        const double Hs = __dsub_rn(L1.H, P);
        const double bb = __dsub_rn(Hs, L1.H);
        const double He = __dsub_rn(__dsub_rn(L1.H, __dsub_rn(Hs, bb)), __dadd_rn(P, bb));
        const double Dv = A.expy(Hs, __dadd_rn(He, __dsub_rn(L1.Lo, Pl)), m3);
        const double a = 1.0 - (nz ? Bv : 0.0);
        kr = A.sqrtr(satn) * a * a;
        psi = A.divs(nz ? -Dv : -0.0, alpha, r_alpha);
        return;
    }
#endif
#ifndef PB_POW_NO_SHARE
    // pow(s, m1) and pow(1/s, m1): one logarithm.  q = fl(1/s) satisfies s q = 1 - e with
    // e = fma(-s, q, 1) exact, so log q = -log s + log(1 - e) = -log s - e (e^2 < 2^-106): the
    // double-double logarithm of q at the accuracy the pow algorithm works with.  (The callers
    // clamp satn to [SATMIN, 1], so sp == satn.)
    const LogDD Ls = A.logp(sp);
    const double Av = A.expy(Ls.H, Ls.Lo, m1);
    const double q = A.divs(1.0, sp);
    const double e = __fma_rn(-sp, q, 1.0);
    const double Cv = A.expy(-Ls.H, -Ls.Lo - e, m1);
#else
    const double Av = A.powp(satn, m1);
    const double Cv = A.powp(A.divs(1.0, sp), m1);
#endif
    const double Bv = A.powp(1.0 - Av, m2);
    const double Dv = A.powp(Cv - 1.0, m3);
    const double a = 1.0 - Bv;
    kr = A.sqrtr(satn) * a * a;
    psi = A.divs(-Dv, alpha, r_alpha);
}

// returns false (nothing written) when FAST arithmetic left its domain
template <bool FAST, bool GH>
__device__ __forceinline__ bool elem_pre(const DevMesh &m, const double *__restrict__ y, int i,
                                         const double *st, unsigned bar, unsigned phase, bool ys)
{
    // ys: the tile's own surf / gw columns arrived with the stage (behind the static slab)
    // st: this lane's column 0 of the stage's tile slab (slots TS_PRE0..TS_PRE1)
#define EC(c) st[((c) - TS_PRE0) * PB_TILE]
    Arith<FAST> A;
    mbar_wait(bar, phase);      // the tile slab has landed (it was requested STAGES tiles ago)
    // neighbour codes first, then every gather unconditionally (non-element
    // edges gather the element itself) so the loads are in flight together
    int code[3], nn[3], cid;
    {
        const int *nbs = reinterpret_cast<const int *>(st - (i & 31) + (TS_NB0 - TS_PRE0) * PB_TILE);
#pragma unroll
        for (int j = 0; j < 3; j++) {
            code[j] = nbs[j * PB_TILE + (i & 31)];
            nn[j] = (code[j] >= 0) ? code[j] : i;
        }
        cid = nbs[3 * PB_TILE + (i & 31)];
    }
#ifdef PB_FAKE_GATHER
    nn[0] = nn[1] = nn[2] = i;
#endif
    const double2 *crow = reinterpret_cast<const double2 *>(m.cls + (size_t)cid * CC_STRIDE);
    const double2 c_mach = __ldg(crow + CC_KMACH / 2);      // {kmach, areafv}
    const double c_ksath = __ldg(m.cls + (size_t)cid * CC_STRIDE + CC_KSATH);
    double ysn[3], zmaxn[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        ysn[j] = ld_here(p_surf<GH>(m, y, nn[j]));
#ifndef PB_PRE_ZMAX_SNB
        // neighbour's zmax out of its tile's column of the static table: that tile's slab is being
        // fetched by some CTA at about the same time, so the line is in L2 (and the static
        // neighbour records are not read by this kernel at all)
        zmaxn[j] = ld_here(&TSC(TS_ZMAX, nn[j]));
#else
        zmaxn[j] = ld_here(&m.snb[nn[j]].y);
#endif
    }
    const double *sy = st + (TS_PRE1 - TS_PRE0) * PB_TILE;
    const double surfh = surf_h_a<FAST>(A, max0(ys ? sy[0] : y_surf<GH>(m, y, i)));
    const double gw = max0(ys ? sy[PB_TILE] : y_gw<GH>(m, y, i));
    const double effkh = eff_kh_a<FAST>(A, EC(TS_DEPTH), EC(TS_DMAC), c_mach.x, c_mach.y, c_ksath, gw);
    // pow(avg_h, 0.6666667) of OverLandFlow (lat_flow.c:270): AvgHsurf (lat_flow.c:175-203)
    // returns the depth above DEPRSTG of the UPWIND element, so the power is a
    // per-element quantity -- evaluated once here instead of once per edge side.  It needs
    // the element's own state only: placed ahead of the friction slope, its FP64 chain runs
    // while the neighbour gathers are in flight.
    const double hd = (surfh > PB_DEPRSTG) ? 1.0 * (surfh - PB_DEPRSTG) : 0.0;
    const double p23 = A.pow23(hd);
#pragma unroll
    for (int j = 0; j < 3; j++) { PB_AFTER(ysn[j], p23); PB_AFTER(zmaxn[j], p23); }
    double sf = 0.0;
    if (m.surf_mode == PB_DIFF_WAVE) {
        const double zmax = EC(TS_ZMAX);
        double h[3], nx[3], ny[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            nx[j] = EC(TS_NABRX0 + j);
            ny[j] = EC(TS_NABRY0 + j);
            h[j] = zmaxn[j] + surf_h_a<FAST>(A, max0(ysn[j]));
        }
        if ((code[0] | code[1] | code[2]) < 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (code[j] == PB_NB_BOUNDARY) {
                    if (m.bct[(size_t)j * m.nes + i] == 0) h[j] = zmax + surfh;
                    else h[j] = FOC(PB_F_BC0 + j, i);
                } else if (code[j] < 0) {
                    const int r = (-code[j] - 2) >> 2;
                    const double stage = max0(y_stage<GH>(m, y, r));
                    h[j] = (stage > RFC(PB_R_SHP_DEPTH, r)) ? RFC(PB_R_ZBED, r) + stage : RFC(PB_R_ZMAX, r);
                }
            }
        }
        const double dx = dh_by_dl_a<FAST>(A, ny, nx, h);
        const double dy = dh_by_dl_a<FAST>(A, nx, ny, h);
        sf = A.sqrtr(dx * dx + dy * dy);
    }
    if (FAST && !A.ok) return false;
    m.dnb[i] = make_double4(surfh, effkh, sf, p23);
    return true;
#undef EC
}

template <bool GH>
__device__ __noinline__ void elem_pre_exact(const DevMesh *gm, const double *__restrict__ y, int i,
                                            const double *st, unsigned bar, unsigned phase, bool ys)
{
    const DevMesh &m = *gm;
    elem_pre<false, GH>(m, y, i, st, bar, phase, ys);
    if (m.slow_count) atomicAdd(m.slow_count, 1ULL);
}

// per-warp gather buffer of k_main: per edge [4][32] x 16 B + [32] x 8 B (+ [32] x 8 B fbr)
#ifndef PB_GATHER_ASYNC
#define PB_GATHER_ASYNC 0
#endif
template <bool FBR> struct MainGather {
    static constexpr int EDGE = 2048 + 256 + (FBR ? 256 : 0);
    static constexpr int WARP = PB_GATHER_ASYNC ? 3 * EDGE + 256 : 0;    // + one scratch double per lane
};

// ---------------------------------------------------------------------------
// element part of k_main
// ---------------------------------------------------------------------------
// FAST: every division / pow goes through Arith<true> (fdiv.cuh) and the
// function returns false, without having written anything, when one of them
// left the fast-path domain; the caller then runs the FAST = false version.
template <bool FBR, bool FAST, bool GH>
__device__ __forceinline__ bool elem_main(const DevMesh &m, const double *__restrict__ y,
                                          double *__restrict__ dy, int i, const double *st,
                                          const double *f, unsigned bar, unsigned phase, int ys,
                                          unsigned char *gb)
{
    // gb: this warp's gather buffer (FAST): per edge [4][32] 16-byte chunks {dnb, snb} of the
    // neighbour + [32] gw (+ [32] fbr_gw), filled by per-lane cp.async at the top of the tile
    // st / f: this lane's column 0 of the stage's slabs (static slots TS_MAIN0.., hot
    // forcing columns), filled by the bulk copies issued STAGES tiles ago in k_main; behind
    // them the tile's own state columns (ys bit 0: gw [, fbr_gw], bit 1: unsat [, fbr_unsat]
    // arrived with the stage; otherwise they are read from y) and its own dynamic records
#define EC(c) st[((c) - TS_MAIN0) * PB_TILE]
#define DISTC(j) m.dist_cold[(size_t)(j) * m.nes + i]
    Arith<FAST> A;
    mbar_wait(bar, phase);      // static + forcing slabs have landed
    // ---- loads: neighbour codes, then all gathers (unconditional) ---------------
    int code[3], nn[3], cid;
    {
        const int *nbs = reinterpret_cast<const int *>(st - (i & 31) + (TS_NB0 - TS_MAIN0) * PB_TILE);
#pragma unroll
        for (int j = 0; j < 3; j++) {
            code[j] = nbs[j * PB_TILE + (i & 31)];
            nn[j] = (code[j] >= 0) ? code[j] : i;
        }
        cid = nbs[3 * PB_TILE + (i & 31)];
    }
    // the element's class row (soil / land cover / geology parameters and derived exponents)
    const double2 *crow = reinterpret_cast<const double2 *>(m.cls + (size_t)cid * CC_STRIDE);
    const double2 c_por = __ldg(crow + CC_POROSITY / 2);    // {porosity, rough}
    const double c_rzd = __ldg(m.cls + (size_t)cid * CC_STRIDE + CC_RZD);
    double4 dn[3], sn[3];     // {surfh, effkh, sf, p23} and {zmin, zmax, rough, zbed} of the neighbours
    double gwn[3];
    double fgn[3] = {0.0, 0.0, 0.0};
    int cidn[3] = {0, 0, 0};
    constexpr int GBE = MainGather<FBR>::EDGE;      // bytes per edge in the gather buffer
    constexpr bool GA = FAST && (PB_GATHER_ASYNC != 0);
#ifdef PB_FAKE_GATHER      // timing experiment only (wrong results): every gather hits the element itself
    nn[0] = nn[1] = nn[2] = i;
#endif
    if (GA) {
        // neighbour gathers straight into shared memory; they are collected below, after the
        // element's own-state chains (van Genuchten), which need none of them
        const unsigned g0 = smem_u32(gb) + (i & 31) * 16;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const char *pd = reinterpret_cast<const char *>(m.dnb + nn[j]);
            const char *ps = reinterpret_cast<const char *>(m.snb + nn[j]);
            const unsigned g = g0 + j * GBE;
            cp_async16(g, pd);
            cp_async16(g + 512, pd + 16);
            cp_async16(g + 1024, ps);
            cp_async16(g + 1536, ps + 16);
            cp_async8(smem_u32(gb) + j * GBE + 2048 + (i & 31) * 8, p_gw<GH>(m, y, nn[j]));
            if (FBR) {
                cp_async8(smem_u32(gb) + j * GBE + 2304 + (i & 31) * 8, p_fg<GH>(m, y, nn[j]));
                cidn[j] = m.cid[nn[j]];
            }
        }
        cp_async_commit();
    } else {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            dn[j] = m.dnb[nn[j]];
            sn[j] = m.snb[nn[j]];
            gwn[j] = max0(y_gw<GH>(m, y, nn[j]));
            if (FBR && FAST) {
                fgn[j] = max0(y_fg<GH>(m, y, nn[j]));
                cidn[j] = m.cid[nn[j]];
            }
        }
    }
    const double *sy = f + 4 * PB_TILE;
    const double4 own = PB_GATHER_ASYNC ? m.dnb[i]
                                        : *reinterpret_cast<const double4 *>(sy - (i & 31) + (FBR ? 4 : 2) * PB_TILE + 4 * (i & 31));
    // ode.c:25-49
    const double unsat = max0((ys & 2) ? sy[1 * PB_TILE] : y[m.o_unsat + i]);
    const double gw = max0((ys & 1) ? sy[0] : y[m.o_gw + i]);
    const double surfh = own.x;
    const double effkh = own.y;
    const double area = EC(TS_AREA);
    const double zmin = EC(TS_ZMIN), zmax = EC(TS_ZMAX);
    const double depth = EC(TS_DEPTH), dinf = EC(TS_DINF);
    const double rough = c_por.y;
    const double pcpdrp = f[0 * PB_TILE];
    // (FAST: the refined reciprocals of the static divisors come with the tile; a NaN marks one outside
    // the fast division's domain and sends the element to the exact path)
    const double r_area = FAST ? EC(TS_RAREA) : 0.0;

    // EtExtract (non-Noah), hydrol.c:51-87
    double edir_surf = 0.0, edir_unsat = 0.0, edir_gw = 0.0, ett_unsat = 0.0, ett_gw = 0.0;
    {
        const double edir = f[1 * PB_TILE], ett = f[2 * PB_TILE];
        if (surfh >= PB_DEPRSTG) edir_surf = edir;
        else if (gw > depth - dinf) edir_gw = edir;
        else edir_unsat = edir;
        if (gw > depth - c_rzd) ett_gw = ett;
        else ett_unsat = ett;
    }

    // van Genuchten pair of the unsaturated zone (FAST): it depends on the element's own state
    // only, so its long dependent FP64 chain is placed ahead of the lateral block and runs while
    // the neighbour gathers are still in flight.  Branch-free: a saturated lane evaluates it for
    // satn = 1 and its results are selected away below.
    const bool sat = gw > depth - dinf;
    double satn = 1.0, satkfunc = 1.0, psi_u = 0.0, deficit = 0.0;
#if !defined(PB_VG_LATE)
    if (FAST) {
        deficit = depth - gw;
        double sv = A.divs(sat ? 1.0 : unsat, sat ? 1.0 : deficit);
        sv = (sv > 1.0) ? 1.0 : sv;
        sv = (sv < PB_SATMIN) ? PB_SATMIN : sv;
        satn = sv;
        const double2 c_vg0 = ldg2_here(crow + CC_ALPHA / 2);   // {alpha, m1}
        const double2 c_vg1 = ldg2_here(crow + CC_M2 / 2);      // {m2, m3}
        const double r_alpha = ldg2_here(crow + CC_KSATH / 2).y;
        double kr;
        vg_kr_psi_a<FAST>(A, satn, c_vg0.x, r_alpha, c_vg0.y, c_vg1.x, c_vg1.y, kr, psi_u);
        satkfunc = sat ? 1.0 : kr;
    }
#endif
    // ... and the same for the bedrock layer (vert_flow.c:284-373), whose pair of pows is
    // independent of the soil's: the two chains interleave
    double fu = 0.0, fg = 0.0, g_deficit = 0.0, g_kr = 1.0, g_psi = 0.0;
    bool full = true;
    if (FBR && FAST) {
        fu = max0((ys & 2) ? sy[3 * PB_TILE] : y[m.o_fu + i]);
        fg = max0((ys & 1) ? sy[2 * PB_TILE] : y[m.o_fg + i]);
        const double gdepth = EC(TS_GDEPTH);
        full = fg >= gdepth;
        g_deficit = gdepth - fg;
        double sv = A.divs(full ? 1.0 : fu, full ? 1.0 : g_deficit);
        sv = (sv > 1.0) ? 1.0 : sv;
        sv = (sv < PB_SATMIN) ? PB_SATMIN : sv;
        const double2 c_g0 = ldg2_here(crow + CC_GALPHA / 2);   // {galpha, gm1}
        const double2 c_g1 = ldg2_here(crow + CC_GM2 / 2);      // {gm2, gm3}
        const double r_galpha = ldg2_here(crow + CC_RPOR / 2).y;
        double kr, ps;
        vg_kr_psi_a<FAST>(A, sv, c_g0.x, r_galpha, c_g0.y, c_g1.x, c_g1.y, kr, ps);
        g_kr = full ? 1.0 : kr;
        g_psi = (ps > PB_PSIMIN) ? ps : PB_PSIMIN;
    }

    if (GA) {
        // collect the gathers (each lane reads what it copied itself); the wait takes the result
        // of the own-state chains as an operand so that it cannot be scheduled ahead of them
        {
            // (a volatile shared store of that result right in front of the wait: ptxas keeps
            // memory operations in order around the wait, and the store needs the chains' result)
            const double dep = FBR ? satkfunc + g_kr + psi_u : satkfunc + psi_u;
            asm volatile("st.volatile.shared.f64 [%0], %1;\n\tcp.async.wait_group 0;"
                         ::"r"(smem_u32(gb) + 3 * GBE + (i & 31) * 8), "d"(dep) : "memory");
        }
        const unsigned char *g0 = gb + (i & 31) * 16;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const unsigned char *g = g0 + j * GBE;
            const double2 a = *reinterpret_cast<const double2 *>(g), b = *reinterpret_cast<const double2 *>(g + 512);
            const double2 c = *reinterpret_cast<const double2 *>(g + 1024), d = *reinterpret_cast<const double2 *>(g + 1536);
            dn[j] = make_double4(a.x, a.y, b.x, b.y);
            sn[j] = make_double4(c.x, c.y, d.x, d.y);
            gwn[j] = max0(*reinterpret_cast<const double *>(gb + j * GBE + 2048 + (i & 31) * 8));
            if (FBR) fgn[j] = max0(*reinterpret_cast<const double *>(gb + j * GBE + 2304 + (i & 31) * 8));
        }
    }

    // LateralFlow, lat_flow.c:17-51.  The element-to-element formulas run for all
    // three edges as one straight-line block (an edge without an element behind it
    // sees the element itself: every difference is zero); boundary and river edges
    // are patched afterwards.
    double ovl[3], sub[3], ovl_infil[3], fbrflow[3] = {0.0, 0.0, 0.0};
    const double sf_i = own.z;
    {
        double num[3], den[3], p[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double edge = EC(TS_EDGE0 + j);
            const double r_dist = FAST ? EC(TS_RDIST0 + j) : 0.0;
            // the distance itself: exact path only (and the fast one of a build without PB_RELAX & 1)
            const double dist = (!FAST || !(PB_RELAX & 1)) ? ((code[j] >= 0) ? DISTC(j) : 1.0) : 1.0;
            if (FAST) A.ok = A.ok && (r_dist == r_dist);
            if (FBR && FAST) {
                // FbrFlowElemToElem, lat_flow.c:374-390 (an edge without an element behind it
                // sees the element itself: zero; boundary and river edges are patched below)
                const double gksath = __ldg(m.cls + (size_t)cid * CC_STRIDE + CC_GKSATH);
                const double gk_n = __ldg(m.cls + (size_t)cidn[j] * CC_STRIDE + CC_GKSATH);
                const double dfh = (fg + EC(TS_ZBED)) - (fgn[j] + sn[j].w);
                fbrflow[j] = 0.5 * (gksath + gk_n) * A.divr(dfh, dist, r_dist) * avg_h_nn(dfh, fg, fgn[j]) * edge;
            }
            const double gw_n = gwn[j], surfh_n = dn[j].x;
            const double zmin_n = sn[j].x, zmax_n = sn[j].y;
            // SubFlowElemToElem, lat_flow.c:273-298
            double diff_h = (gw + zmin) - (gw_n + zmin_n);
            double avgh = avg_h_nn(diff_h, gw, gw_n);
            double grad_h = A.divr(diff_h, dist, r_dist);
            double avg_ksat = 0.5 * (effkh + dn[j].y);
            sub[j] = avg_ksat * grad_h * avgh * edge;
            // OvlFlowElemToElem, lat_flow.c:300-327 with avg_sf of :33-36
            double avg_sf;
            diff_h = (m.surf_mode == PB_KINEMATIC) ? zmax - zmax_n
                                                   : (surfh + zmax) - (surfh_n + zmax_n);
            avgh = avg_hsurf(diff_h, surfh, surfh_n);
            grad_h = A.divr(diff_h, dist, r_dist);
            if (m.surf_mode == PB_KINEMATIC) {
                avg_sf = (grad_h > 0.0) ? grad_h : PB_GRADMIN;
            } else {
                avg_sf = 0.5 * (sf_i + dn[j].z);
                avg_sf = (avg_sf > PB_GRADMIN) ? avg_sf : PB_GRADMIN;
            }
            const double avg_rough = 0.5 * (rough + sn[j].z);
            p[j] = (diff_h > 0.0) ? own.w : dn[j].w;   // pow(avgh, 0.6666667) of the upwind element (k_pre)
#if PB_RELAX & 64
            den[j] = FAST ? A.rsqrt_times_rcp(avg_sf, avg_rough) : A.sqrtr(avg_sf) * avg_rough;   // FAST: 1 / den
#else
            den[j] = A.sqrtr(avg_sf) * avg_rough;
#endif
            num[j] = avgh * edge;                     // crossa; crossa * p * grad left to right (lat_flow.c:270)
            ovl_infil[j] = grad_h;                    // parked: grad_h of the overland flux
        }
        // OverLandFlow, lat_flow.c:267-271
#pragma unroll
        for (int j = 0; j < 3; j++) {
#if PB_RELAX & 64
            ovl[j] = FAST ? num[j] * p[j] * ovl_infil[j] * den[j] : A.divr(num[j] * p[j] * ovl_infil[j], den[j]);
#else
            ovl[j] = A.divr(num[j] * p[j] * ovl_infil[j], den[j]);
#endif
            ovl_infil[j] = ovl[j];
        }
    }
    if ((code[0] | code[1] | code[2]) < 0) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (code[j] == PB_NB_BOUNDARY) {
                // BoundFluxElem, lat_flow.c:329-371
                const int bc = m.bct[(size_t)j * m.nes + i];
                ovl[j] = 0.0;
                if (bc == 0) {
                    sub[j] = 0.0;
                } else if (bc > 0) {
                    const double head = FOC(PB_F_BC0 + j, i);
                    double diff_h = gw + zmin - head;
                    double avgh = avg_h(diff_h, gw, head - zmin);
                    double grad_h = div_pos(diff_h, DISTC(j));
                    sub[j] = effkh * grad_h * avgh * EC(TS_EDGE0 + j);
                } else {
                    sub[j] = -FOC(PB_F_BC0 + j, i);
                }
                ovl_infil[j] = ovl[j];
            } else if (code[j] < 0) {
                // river edge: RiverToElem write-back, river_flow.c:159-180
                const int c = -code[j] - 2, r = c >> 2, side = c & 3;
                if (side < 2) {
                    ovl[j] = -RFLX(RF_LEFT_S2C + side, r);
                    sub[j] = -(RFLX(RF_LEFT_A2C + side, r) + RFLX(RF_LEFT_A2A + side, r));
                    ovl_infil[j] = -m.s2c_stale[(size_t)side * m.nrs + r];
                } else {
                    ovl[j] = sub[j] = ovl_infil[j] = 0.0;
                }
            }
        }
    }

    // VerticalFlow: Infil (vert_flow.c:33-130) and Recharge (:132-170).  Both
    // use the same satn / KrFunc / Psi in the unsaturated branch.
    double infil, rechg;
#if !defined(PB_VG_LATE)
    if (FAST) {
        // straight-line form of the block below: both branches of Infil share the denominator
        // of dh_by_dz, divisors of lanes that do not use a quotient are replaced by 1.0
        const double2 c_kinf = ldg2_here(crow + CC_KINFV / 2);      // {kinfv, kmacv}
        const double2 c_afh = ldg2_here(crow + CC_AREAFH / 2);      // {areafh, ksatv}
        const double kinfv = c_kinf.x, kmacv = c_kinf.y;
        const double areafh = c_afh.x;
        double applrate = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) applrate += A.divs(-ovl_infil[j], area, r_area);
        applrate = (applrate > 0.0) ? applrate : 0.0;
        applrate += pcpdrp;
        double wetfrac = A.divr(surfh, PB_DEPRSTG, m.r_deprstg);
        wetfrac = (wetfrac > 0.0) ? wetfrac : 0.0;
        wetfrac = (wetfrac < 1.0) ? wetfrac : 1.0;
        const double psi_c = (psi_u > PB_PSIMIN) ? psi_u : PB_PSIMIN;
        const double h_u = psi_c + zmax - 0.5 * dinf;
        double dh_by_dz = A.quor(surfh + zmax - (sat ? (gw + zmin) : h_u), 0.5 * (surfh + dinf));
        dh_by_dz = (surfh <= 0.0 && dh_by_dz > 0.0) ? 0.0 : dh_by_dz;
        const double kinf = eff_kinf(kinfv, kmacv, areafh, dh_by_dz, satkfunc, satn, applrate, surfh);
        double v = kinf * dh_by_dz;
        v = (sat || v > 0.0) ? v : 0.0;          // the unsaturated branch clamps at zero (vert_flow.c:112)
        const double ws0surf = f[3 * PB_TILE];
        const double infil_max = applrate + ((ws0surf > 0.0) ? A.divr(ws0surf, m.dt, m.r_dt) : 0.0);
        v = (v > infil_max) ? infil_max : v;
        v *= wetfrac;
        infil = (unsat + gw > depth) ? 0.0 : v;
        // Recharge; AvgKv (_ARITH_), vert_flow.c:172-209
        const double ksatv = c_afh.y, dmac = EC(TS_DMAC);
        const bool deep = deficit > dmac;
        const double k1 = satkfunc * ksatv;
        const double kmid = (areafh > 0.0) ? kmacv * areafh + ksatv * (1.0 - areafh) : ksatv;
        const double d1 = deep ? dmac : deficit;
        const double k2 = deep ? k1 : kmid;
        const double d2 = deep ? deficit - dmac : dmac - deficit;
        const double d3 = deep ? gw : gw - (dmac - deficit);
        const double kavg = A.quor(k1 * d1 + k2 * d2 + ksatv * d3, sat ? 1.0 : d1 + d2 + d3);
        const double dh2 = A.quor(0.5 * deficit + psi_u, sat ? 1.0 : 0.5 * (deficit + gw));
        double rc = kavg * dh2;
        rc = (rc > 0.0 && unsat <= 0.0) ? 0.0 : rc;
        rc = (rc < 0.0 && gw <= 0.0) ? 0.0 : rc;
        rechg = sat ? infil : rc;
    } else
#endif
    {
        const double2 c_kinf = ldg2_here(crow + CC_KINFV / 2);      // {kinfv, kmacv}
        const double2 c_afh = ldg2_here(crow + CC_AREAFH / 2);      // {areafh, ksatv}
        const double kinfv = c_kinf.x, kmacv = c_kinf.y;
        const double areafh = c_afh.x;
        if (!sat) {
            deficit = depth - gw;
            satn = A.divs(unsat, deficit);
            satn = (satn > 1.0) ? 1.0 : satn;
            satn = (satn < PB_SATMIN) ? PB_SATMIN : satn;
            const double2 c_vg0 = ldg2_here(crow + CC_ALPHA / 2);   // {alpha, m1}
            const double2 c_vg1 = ldg2_here(crow + CC_M2 / 2);      // {m2, m3}
            const double r_alpha = ldg2_here(crow + CC_KSATH / 2).y;
            vg_kr_psi_a<FAST>(A, satn, c_vg0.x, r_alpha, c_vg0.y, c_vg1.x, c_vg1.y, satkfunc, psi_u);
        }
        if (unsat + gw > depth) {
            infil = 0.0;
        } else {
            double applrate = 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++) applrate += A.divs(-ovl_infil[j], area, r_area);
            applrate = (applrate > 0.0) ? applrate : 0.0;
            applrate += pcpdrp;
            double wetfrac = A.divr(surfh, PB_DEPRSTG, m.r_deprstg);
            wetfrac = (wetfrac > 0.0) ? wetfrac : 0.0;
            wetfrac = (wetfrac < 1.0) ? wetfrac : 1.0;
            double dh_by_dz;
            if (sat) {
                // KrFunc(beta, 1.0) == 1.0 exactly (pow(1,x) = 1, pow(0,x>0) = 0)
                dh_by_dz = A.quor(surfh + zmax - (gw + zmin), 0.5 * (surfh + dinf));
                dh_by_dz = (surfh <= 0.0 && dh_by_dz > 0.0) ? 0.0 : dh_by_dz;
                double kinf = eff_kinf(kinfv, kmacv, areafh, dh_by_dz, 1.0, 1.0, applrate, surfh);
                infil = kinf * dh_by_dz;
            } else {
                double psi_c = (psi_u > PB_PSIMIN) ? psi_u : PB_PSIMIN;
                double h_u = psi_c + zmax - 0.5 * dinf;
                dh_by_dz = A.quor(surfh + zmax - h_u, 0.5 * (surfh + dinf));
                dh_by_dz = (surfh <= 0.0 && dh_by_dz > 0.0) ? 0.0 : dh_by_dz;
                double kinf = eff_kinf(kinfv, kmacv, areafh, dh_by_dz, satkfunc, satn, applrate, surfh);
                infil = kinf * dh_by_dz;
                infil = (infil > 0.0) ? infil : 0.0;
            }
            const double ws0surf = f[3 * PB_TILE];
            double infil_max = applrate + ((ws0surf > 0.0) ? A.divr(ws0surf, m.dt, m.r_dt) : 0.0);
            infil = (infil > infil_max) ? infil_max : infil;
            infil *= wetfrac;
        }
        if (sat) {
            rechg = infil;
        } else {
            // AvgKv (_ARITH_), vert_flow.c:172-209
            const double ksatv = c_afh.y, dmac = EC(TS_DMAC);
            double k1, k2, k3, d1, d2, d3;
            if (deficit > dmac) {
                k1 = satkfunc * ksatv; d1 = dmac;
                k2 = satkfunc * ksatv; d2 = deficit - dmac;
                k3 = ksatv; d3 = gw;
            } else {
                k1 = satkfunc * ksatv; d1 = deficit;
                k2 = (areafh > 0.0) ? kmacv * areafh + ksatv * (1.0 - areafh) : ksatv;
                d2 = dmac - deficit;
                k3 = ksatv; d3 = gw - (dmac - deficit);
            }
            double kavg = A.quor(k1 * d1 + k2 * d2 + k3 * d3, d1 + d2 + d3);
            double dh_by_dz = A.quor(0.5 * deficit + psi_u, 0.5 * (deficit + gw));
            rechg = kavg * dh_by_dz;
            rechg = (rechg > 0.0 && unsat <= 0.0) ? 0.0 : rechg;
            rechg = (rechg < 0.0 && gw <= 0.0) ? 0.0 : rechg;
        }
    }

    // assemble, ode.c:108-150
    double dsurf = 0.0, dunsat = 0.0, dgw = 0.0;
    dsurf += pcpdrp - infil - edir_surf;
    dunsat += infil - rechg - edir_unsat - ett_unsat;
    dgw += rechg - edir_gw - ett_gw;

    double fbr_infil = 0.0, fbr_rechg = 0.0, dfu = 0.0, dfg = 0.0;
    if (FBR && FAST) {
        // straight-line form of the block below (FbrInfil, FbrRecharge): divisors of lanes that
        // do not use a quotient are replaced by 1.0
        const double2 c_gk = ldg2_here(crow + CC_GKSATV / 2);   // {gksatv, gksath}
        const double ksatv_s = ldg2_here(crow + CC_AREAFH / 2).y;   // soil ksatv
        const double gksatv = c_gk.x, gksath = c_gk.y;
        const double zbed = EC(TS_ZBED);
        const double gdepth = EC(TS_GDEPTH);
        const bool none = (fu + fg > gdepth) || (gw <= 0.0);     // vert_flow.c:298-301
        const bool calc = !full && !none;
        const double h_u = g_psi + zmin - 0.5 * g_deficit;
        const double dh1 = A.quor(zmin + gw - h_u, calc ? 0.5 * (gw + g_deficit) : 1.0);
        const double q1 = A.quor(gw, ksatv_s), q2 = A.quor(g_deficit, calc ? gksatv * g_kr : 1.0);
        const double kavg1 = A.quor(gw + g_deficit, calc ? q1 + q2 : 1.0);
        fbr_infil = full ? -ksatv_s : (none ? 0.0 : kavg1 * dh1);
        const double dh2 = A.quor(0.5 * g_deficit + g_psi, full ? 1.0 : 0.5 * (g_deficit + fg));
        const double kavg2 = A.quor(fu * gksatv * g_kr + fg * gksatv, full ? 1.0 : fu + fg);
        double rc = kavg2 * dh2;
        rc = (rc > 0.0 && fu <= 0.0) ? 0.0 : rc;
        rc = (rc < 0.0 && fg <= 0.0) ? 0.0 : rc;
        fbr_rechg = full ? fbr_infil : rc;
        if ((code[0] | code[1] | code[2]) < 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (code[j] == PB_NB_BOUNDARY) {
                    // FbrBoundFluxElem, lat_flow.c:392-424
                    const int bc = m.fbct[(size_t)j * m.nes + i];
                    if (bc == 0) {
                        fbrflow[j] = 0.0;
                    } else if (bc > 0) {
                        const double head = FOC(PB_F_FBRBC0 + j, i);
                        double diff_h = fg + zbed - head;
                        double avgh = avg_h(diff_h, fg, head - zbed);
                        double grad_h = div_pos(diff_h, DISTC(j));
                        fbrflow[j] = gksath * grad_h * avgh * EC(TS_EDGE0 + j);
                    } else {
                        fbrflow[j] = -FOC(PB_F_FBRBC0 + j, i);
                    }
                } else if (code[j] < 0) {
                    // neighbour across the river, lat_flow.c:85-100
                    const int r = (-code[j] - 2) >> 2;
                    const int l = RIC(PB_RI_LEFTELE, r);
                    const int n = (l == i) ? RIC(PB_RI_RIGHTELE, r) : l;
                    const double fg_n = max0(y_fg<GH>(m, y, n));
                    double diff_h = (fg + zbed) - (fg_n + m.snb[n].w);
                    double avgh = avg_h(diff_h, fg, fg_n);
                    double grad_h = div_any(diff_h, m.fbr_dist[r]);
                    double avg_ksat = 0.5 * (gksath + CLE(CC_GKSATH, n));
                    fbrflow[j] = avg_ksat * grad_h * avgh * EC(TS_EDGE0 + j);
                }
            }
        }
        dgw -= fbr_infil;
        dfu += fbr_infil - fbr_rechg;
        dfg += fbr_rechg;
    } else if (FBR) {
        const double fu = max0((ys & 2) ? sy[3 * PB_TILE] : y[m.o_fu + i]);
        const double fg = max0((ys & 1) ? sy[2 * PB_TILE] : y[m.o_fg + i]);
        const double2 c_g0 = ldg2_here(crow + CC_GALPHA / 2);   // {galpha, gm1}
        const double2 c_g1 = ldg2_here(crow + CC_GM2 / 2);      // {gm2, gm3}
        const double2 c_gk = ldg2_here(crow + CC_GKSATV / 2);   // {gksatv, gksath}
        const double ksatv_s = ldg2_here(crow + CC_AREAFH / 2).y;   // soil ksatv
        const double2 c_rg = ldg2_here(crow + CC_RPOR / 2);         // {1/porosity, 1/galpha}
        const double gdepth = EC(TS_GDEPTH), gksatv = c_gk.x;
        const double zbed = EC(TS_ZBED), gksath = c_gk.y;
        const bool full = fg >= gdepth;
        double deficit = 0.0, satkfunc = 1.0, psi_c = 0.0;
        if (!full) {
            deficit = gdepth - fg;
            double satn = A.divs(fu, deficit);
            satn = (satn > 1.0) ? 1.0 : satn;
            satn = (satn < PB_SATMIN) ? PB_SATMIN : satn;
            vg_kr_psi_a<FAST>(A, satn, c_g0.x, c_rg.y, c_g0.y, c_g1.x, c_g1.y, satkfunc, psi_c);
            psi_c = (psi_c > PB_PSIMIN) ? psi_c : PB_PSIMIN;
        }
        // FbrInfil, vert_flow.c:284-330
        if (full) {
            fbr_infil = -ksatv_s;
        } else if (fu + fg > gdepth || gw <= 0.0) {
            fbr_infil = 0.0;
        } else {
            double h_u = psi_c + zmin - 0.5 * deficit;
            double dh_by_dz = A.quor(zmin + gw - h_u, 0.5 * (gw + deficit));
            double kavg = A.quor(gw + deficit, A.quor(gw, ksatv_s) + A.quor(deficit, gksatv * satkfunc));
            fbr_infil = kavg * dh_by_dz;
        }
        // FbrRecharge, vert_flow.c:332-373
        if (full) {
            fbr_rechg = fbr_infil;
        } else {
            double dh_by_dz = A.quor(0.5 * deficit + psi_c, 0.5 * (deficit + fg));
            double kavg = A.quor(fu * gksatv * satkfunc + fg * gksatv, fu + fg);
            fbr_rechg = kavg * dh_by_dz;
            fbr_rechg = (fbr_rechg > 0.0 && fu <= 0.0) ? 0.0 : fbr_rechg;
            fbr_rechg = (fbr_rechg < 0.0 && fg <= 0.0) ? 0.0 : fbr_rechg;
        }
        // lateral bedrock flow, lat_flow.c:56-115
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (code[j] == PB_NB_BOUNDARY) {
                // FbrBoundFluxElem, lat_flow.c:392-424
                const int bc = m.fbct[(size_t)j * m.nes + i];
                if (bc == 0) {
                    fbrflow[j] = 0.0;
                } else if (bc > 0) {
                    const double head = FOC(PB_F_FBRBC0 + j, i);
                    double diff_h = fg + zbed - head;
                    double avgh = avg_h(diff_h, fg, head - zbed);
                    double grad_h = div_pos(diff_h, DISTC(j));
                    fbrflow[j] = gksath * grad_h * avgh * EC(TS_EDGE0 + j);
                } else {
                    fbrflow[j] = -FOC(PB_F_FBRBC0 + j, i);
                }
            } else {
                int n;
                double dist;
                if (code[j] >= 0) {
                    n = code[j];
                    dist = DISTC(j);
                } else {
                    // neighbour across the river, lat_flow.c:85-100
                    const int r = (-code[j] - 2) >> 2;
                    const int l = RIC(PB_RI_LEFTELE, r);
                    n = (l == i) ? RIC(PB_RI_RIGHTELE, r) : l;
                    dist = m.fbr_dist[r];
                }
                // FbrFlowElemToElem, lat_flow.c:374-390
                const double fg_n = max0(y_fg<GH>(m, y, n));
                double diff_h = (fg + zbed) - (fg_n + m.snb[n].w);
                double avgh = avg_h(diff_h, fg, fg_n);
                double grad_h = A.divr(diff_h, dist);
                double avg_ksat = 0.5 * (gksath + CLE(CC_GKSATH, n));
                fbrflow[j] = avg_ksat * grad_h * avgh * EC(TS_EDGE0 + j);
            }
        }
        dgw -= fbr_infil;
        dfu += fbr_infil - fbr_rechg;
        dfg += fbr_rechg;
    }

#pragma unroll
    for (int j = 0; j < 3; j++) {
        dsurf -= A.divr(ovl[j], area, r_area);
        dgw -= A.divr(sub[j], area, r_area);
        if (FBR) dfg -= A.divr(fbrflow[j], area, r_area);
    }
    const double porosity = c_por.x;
    const double r_por = ldg2_here(crow + CC_RPOR / 2).x;
    dunsat = A.divr(dunsat, porosity, r_por);
    dgw = A.divr(dgw, porosity, r_por);
    if (FBR) {
        const double gporosity = __ldg(m.cls + (size_t)cid * CC_STRIDE + CC_GPOROSITY);
        const double r_gpor = ldg2_here(crow + CC_RGPOR / 2).x;
        dfu = A.divr(dfu, gporosity, r_gpor);
        dfg = A.divr(dfg, gporosity, r_gpor);
    }
    if (FAST && !A.ok) return false;        // nothing written yet: the caller recomputes this element

    dy[i] = dsurf;
    dy[m.o_unsat + i] = dunsat;
    dy[m.o_gw + i] = dgw;
    bool bad = isnan(dsurf) || isnan(dunsat) || isnan(dgw);
    if (FBR) {
        dy[m.o_fu + i] = dfu;
        dy[m.o_fg + i] = dfg;
        bad = bad || isnan(dfu) || isnan(dfg);
    }
    if (bad) atomicOr(m.nan_flag, 1);   // CheckDy, ode.c:302-311

    if (m.record) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            XFC(PB_X_OVL0 + j, i) = ovl[j];
            XFC(PB_X_SUB0 + j, i) = sub[j];
            XFC(PB_X_FBRFLOW0 + j, i) = fbrflow[j];
        }
        XFC(PB_X_INFIL, i) = infil;
        XFC(PB_X_RECHG, i) = rechg;
        XFC(PB_X_EDIR_SURF, i) = edir_surf;
        XFC(PB_X_EDIR_UNSAT, i) = edir_unsat;
        XFC(PB_X_EDIR_GW, i) = edir_gw;
        XFC(PB_X_ETT_UNSAT, i) = ett_unsat;
        XFC(PB_X_ETT_GW, i) = ett_gw;
        XFC(PB_X_FBR_INFIL, i) = fbr_infil;
        XFC(PB_X_FBR_RECHG, i) = fbr_rechg;
    }
    return true;
#undef EC
#undef DISTC
}

// the rare elements whose arithmetic left the fast-path domain: plain `/` and pow()
template <bool FBR, bool GH>
__device__ __noinline__ void elem_main_exact(const DevMesh *gm, const double *__restrict__ y,
                                             double *__restrict__ dy, int i, const double *st,
                                             const double *f, unsigned bar, unsigned phase, int ys)
{
    const DevMesh &m = *gm;
    elem_main<FBR, false, GH>(m, y, dy, i, st, f, bar, phase, ys, nullptr);
    if (m.slow_count) atomicAdd(m.slow_count, 1ULL);
}

// ---------------------------------------------------------------------------
// river part of k_main: the serial accumulation of river_flow.c:94-108 as an
// ordered gather over the upstream list, then ode.c:228-252
__device__ __forceinline__ void river_main(const DevMesh &m, double *__restrict__ dy, int r)
{
    double up_c2c = RFLX(RF_UP_C2C, r);
    double up_a2a = 0.0;
    for (int k = m.up_ptr[r]; k < m.up_ptr[r + 1]; k++) {
        const int u = m.up_idx[k];
        up_c2c -= RFLX(RF_DOWN_C2C, u);
        up_a2a -= RFLX(RF_DOWN_A2A, u);
    }
    RFLX(RF_UP_C2C, r) = up_c2c;
    RFLX(RF_UP_A2A, r) = up_a2a;
    const double area = RFC(PB_R_AREA, r);
    double dstg = 0.0;
    dstg -= div_pos(up_c2c, area);
#pragma unroll
    for (int j = 1; j <= 6; j++) dstg -= div_pos(RFLX(j, r), area);
    double drgw = 0.0;
    drgw += -RFLX(RF_LEFT_A2A, r) - RFLX(RF_RIGHT_A2A, r) - RFLX(RF_DOWN_A2A, r) - up_a2a +
        RFLX(RF_CHANL_LKG, r);
    drgw /= RFC(PB_R_POROSITY, r) * area;
    dy[m.o_stg + r] = dstg;
    dy[m.o_rgw + r] = drgw;
    if (isnan(dstg) || isnan(drgw)) atomicOr(m.nan_flag, 1);
}

// ---------------------------------------------------------------------------
// kernels: persistent CTAs, one ring of TMA-filled stages per CTA (see the top of
// this file).  Tickets q = 0, 1, .. of CTA b map to tile (q / 8 * gridDim.x + b) * 8 + q % 8,
// i.e. the 8 warps of a CTA work on 8 consecutive tiles (two locality patches) at a time, so
// most neighbour gathers hit lines another lane of the same CTA has brought into L1.  The
// river tiles (32 segments per warp; few, slow) come first, padded to a whole group, then
// the element tiles.
// ---------------------------------------------------------------------------
#ifndef PB_RHS_WARPS
#define PB_RHS_WARPS 8     // k_main: 8 warps x 2 CTAs -> 128 registers, no spills: 113.8 us against 116.6 us with
#endif                     // 10 warps at 96 registers (7 / 6 / 5 warps: 121 / 122 / 133 us; 11 at 80: 121 us)
#define PB_RHS_THREADS (PB_RHS_WARPS * 32)
#ifndef PB_PRE_WARPS
#define PB_PRE_WARPS 10    // k_pre: 3 x 10 (3 x 8, 2 x 12, 2 x 10, 4 x 6: 1.5-5 us slower)
#endif
#define PB_PRE_THREADS (PB_PRE_WARPS * 32)
#ifndef PB_PRE_HALO_FRAC
#define PB_PRE_HALO_FRAC 35    // per cent of the interior tiles k_pre takes before the tiles that wait for the halo
#endif
#ifndef PB_PRE_STAGES
#define PB_PRE_STAGES 12
#endif
#ifndef PB_MAIN_STAGES
#define PB_MAIN_STAGES 12
#endif
#ifndef PB_MAIN_STAGES_FBR
#define PB_MAIN_STAGES_FBR 12
#endif
#ifndef PB_PRE_MINB
#define PB_PRE_MINB 3      // CTAs of 10 warps per SM: 3 -> 64 registers
#endif
#ifndef PB_MAIN_MINB
#define PB_MAIN_MINB 2      // 2 CTAs of 8 warps -> 128 registers, no spills
#endif
#ifndef PB_MAIN_MINB_FBR
#define PB_MAIN_MINB_FBR 2
#endif
#ifndef PB_MAIN_WARPS_FBR
#define PB_MAIN_WARPS_FBR 10   // fbr: 10 warps x 2 CTAs at 96 registers (8 x 2 at 128: 134.6 -> 130.3 us at 1M, r02q)
#endif

#ifndef PB_RING_GROUP
#define PB_RING_GROUP 1     // consecutive tiles a CTA takes at a time (1: tile t goes to CTA t mod grid)
#endif
template <int STAGES, int STAGE_BYTES> struct Ring {
    unsigned char *base;
    unsigned bar0;           // shared-space address of the first mbarrier
    volatile int *seq;       // per stage: tiles finished on it (= index of the phase it is in)
    int *ticket;
    __device__ __forceinline__ Ring(unsigned char *smem)
        : base(smem), bar0(smem_u32(smem + (size_t)STAGES * STAGE_BYTES)),
          seq(reinterpret_cast<volatile int *>(smem + (size_t)STAGES * STAGE_BYTES + 8 * STAGES)),
          ticket(reinterpret_cast<int *>(smem + (size_t)STAGES * STAGE_BYTES + 12 * STAGES)) {}
    __device__ __forceinline__ void init()
    {
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; s++) { mbar_init(bar0 + 8 * s, 1); seq[s] = 0; }
            *ticket = 0;
        }
        __syncthreads();
    }
    // A parity wait cannot tell phase n+1 from phase n-1, so before waiting for the copy of
    // its tile (phase n of the stage) a warp makes sure the stage has left phase n-1, i.e.
    // the previous tile on it is finished and the stage has been re-armed or is about to be.
    __device__ __forceinline__ void acquire(int s, int n) const
    {
        while (seq[s] != n) { }
    }
    __device__ __forceinline__ void release(int s, int n) { seq[s] = n + 1; }
    __device__ __forceinline__ int take(int lane)
    {
        int q = 0;
        if (lane == 0) q = atomicAdd(ticket, 1);
        return __shfl_sync(0xffffffffu, q, 0);
    }
    __device__ __forceinline__ unsigned bar(int s) const { return bar0 + 8 * s; }
    __device__ __forceinline__ unsigned char *stage(int s) const { return base + (size_t)s * STAGE_BYTES; }
    // behind the control words: up to 8 copy descriptors of a stage (lane c issues copy c)
    __host__ __device__ static constexpr int desc_off() { return (STAGES * STAGE_BYTES + 12 * STAGES + 16 + 31) / 32 * 32; }
    __host__ __device__ static constexpr int smem_bytes() { return desc_off() + 8 * 32; }
};
// One bulk copy of a stage: src = base + tile * stride (bytes), dst = stage + off; kind 1 = L2 prefetch only
struct __align__(16) StageCopy { unsigned long long base; unsigned stride, off, bytes, kind, pad0, pad1; };

template <bool FBR> struct MainCfg {
    static constexpr int NC = (FBR ? TS_FBR1 : TS_MAIN1) - TS_MAIN0;
    // static slab | hot forcing columns | own state {gw, unsat[, fbr_gw, fbr_unsat]} | own dynamic record
    static constexpr int SBS = NC * PB_TILE * 8, SBF = 4 * PB_TILE * 8;
    // (with asynchronous gathers the own dynamic record is fetched with them: shared memory is short)
    static constexpr int NY = FBR ? 4 : 2, SBY = NY * PB_TILE * 8, SBD = PB_GATHER_ASYNC ? 0 : PB_TILE * 32;
    static constexpr int SB = SBS + SBF + SBY + SBD;
    static constexpr int STAGES = FBR ? PB_MAIN_STAGES_FBR : PB_MAIN_STAGES;
    static constexpr int MINB = FBR ? PB_MAIN_MINB_FBR : PB_MAIN_MINB;
    static constexpr int WARPS = FBR ? PB_MAIN_WARPS_FBR : PB_RHS_WARPS;
    static constexpr int THREADS = WARPS * 32;
    typedef Ring<STAGES, SB> ring_t;
    static constexpr int RING_BYTES = (ring_t::smem_bytes() + 127) / 128 * 128;
    static constexpr int SMEM = RING_BYTES + WARPS * MainGather<FBR>::WARP;
};
struct PreCfg {
    static constexpr int NC = TS_PRE1 - TS_PRE0;
    static constexpr int SBS = NC * PB_TILE * 8, SBY = 2 * PB_TILE * 8;     // static slab | own {surf, gw}
    static constexpr int SB = SBS + SBY;
    typedef Ring<PB_PRE_STAGES, SB> ring_t;
};

// Timing experiment (-DPB_HALO_TIMING, tools/mgpu_rhs_probe.py): %globaltimer stamps of the halo
// protocol per RHS evaluation, [seq & 255][slot]; slots: 0 first CTA past the dependency wait,
// 1 arrival flags raised, 2 / 3 first / last warp reaches the flag wait, 4 last warp leaves it,
// 5 k_pre ends, 6 k_main's first CTA past its dependency wait, 7 k_main ends
#ifdef PB_HALO_TIMING
__device__ unsigned long long g_ht[256][8];
__device__ int g_ht_seq;
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define HT_MIN(sq, k) atomicMin(&g_ht[(sq) & 255][k], gtime())
#define HT_MAX(sq, k) atomicMax(&g_ht[(sq) & 255][k], gtime())
#else
#define HT_MIN(sq, k) ((void)0)
#define HT_MAX(sq, k) ((void)0)
#endif

// Halo exchange over peer memory, folded into k_pre (partitioned run).  At its start the grid
// stores this rank's boundary states straight into the neighbours' ghost buffers (NVLink
// stores) as records in their ghost order -- elements {surf, gw[, fbr_gw]}, rivers {stage, gw}
// -- and the last CTA to finish raises this rank's arrival flag in each of them.  The CTAs then
// work on the INTERIOR element tiles (the partitioner lists the owned elements as [interior |
// boundary]: an interior tile reads no ghost) and only before the first river / boundary / ghost
// tile does a warp wait for the flags of this rank's neighbours.  Two parity copies of the ghost
// buffers alternate per RHS: a neighbour can start the exchange of RHS n+2 only after this rank
// has sent n+1, i.e. after it has finished reading the copy of RHS n.
struct HaloPut {
    int nse, nsr;                     // records to send: elements, rivers
    const int *send_e, *send_r;       // local owned indices, grouped by neighbour (hp.e_ptr / r_ptr)
    HaloPeers hp;
    int par;                          // parity copy of this exchange
    double seq;                       // its sequence number (the flag value)
    unsigned int *counter;            // CTAs that have finished their stores
    int nput;                         // CTAs that put (the first nput of the grid), one record per thread
};

// Executed by the first h.nput CTAs of k_pre before they turn to their work items: one record per thread,
// one system-scope fence per CTA, by the thread that counts the CTA in after the CTA barrier (the barrier
// orders the other threads' stores before it).  Measured on 2 B200 (PB_HALO_TIMING, 3.4 k records): flags up
// 11 us after the kernel starts (two sc fences and the NVLink round trip in between), seen by the
// neighbour ~10 us later.  Variants tried and dropped: every CTA of the grid putting and fencing (same
// 11 us, but all of them start late); a single communication CTA that does nothing else (27 rounds of
// dependent latencies: 25-29 us); a few dedicated CTAs with four records per thread (same time as this).
__device__ __forceinline__ void halo_put(const DevMesh &m, const double *__restrict__ y, const HaloPut &h)
{
    const HaloPeers &hp = h.hp;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < h.nse + h.nsr; k += h.nput * blockDim.x) {
        if (k < h.nse) {
            int kk = 0;
            while (kk + 1 < hp.nn && k >= hp.e_ptr[kk + 1]) kk++;
            const int i = h.send_e[k];
            double *dst = hp.base[kk] + h.par * hp.pstride[kk] + hp.gel_off[kk] + (size_t)(k - hp.e_ptr[kk]) * m.gs;
            dst[0] = y[i];
            dst[1] = y[m.o_gw + i];
            if (m.gs == 3) dst[2] = y[m.o_fg + i];
        } else {
            const int kr = k - h.nse;
            int kk = 0;
            while (kk + 1 < hp.nn && kr >= hp.r_ptr[kk + 1]) kk++;
            const int r = h.send_r[kr];
            double *dst = hp.base[kk] + h.par * hp.pstride[kk] + hp.gri_off[kk] + (size_t)(kr - hp.r_ptr[kk]) * 2;
            dst[0] = y[m.o_stg + r];
            dst[1] = y[m.o_rgw + r];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_sys();
        if (atomicAdd(h.counter, 1u) == (unsigned)h.nput - 1) {
            *h.counter = 0u;
            fence_sys();
            for (int kk = 0; kk < hp.nn; kk++) {
                volatile double *f = hp.base[kk] + hp.flag_off[kk] + h.par * PB_MAX_RANKS_H + hp.myrank;
                *f = h.seq;
            }
            HT_MAX((int)h.seq, 1);
        }
    }
}

template <bool GH>     // GH: the local mesh has ghost entities (partitioned run)
__global__ void __launch_bounds__(PB_PRE_THREADS, PB_PRE_MINB)
k_pre(const DevMesh m, const double *__restrict__ y, int ntile_e, int ntile_r, int ntile_int, const HaloWait hw,
      const HaloPut hput, int ys)
{
    constexpr int SBS = PreCfg::SBS;
    // programmatic dependent launch: k_main's CTAs may become resident as this kernel's CTAs
    // retire and run their prologue (ring set-up, first slab requests); they wait for the
    // completion of this grid (griddepcontrol.wait) before they touch what it writes
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ __align__(128) unsigned char smem[];
    PreCfg::ring_t ring(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, b = blockIdx.x;
    // Work items of the grid, taken by ticket (ticket q of CTA b = item q * G + b):
    //   [ clean river tiles | interior A | other river tiles | boundary + ghost element tiles | interior B ]
    // "clean" / "interior": the tile reads no ghost state (pihm_b200_create_part classifies them; the
    // partitioner lists the owned elements as [interior | boundary], the river tiles are taken through
    // m.riv_tile_order, clean ones first).  With a peer-memory halo the first two sections run while the
    // neighbours' records are on their way (interior A is the first PB_PRE_HALO_FRAC per cent of the
    // interior tiles); every warp waits for the flags before its first item behind them.  The tiles that
    // need the records come next and NOT last: a river tile is a long serial computation -- at the end
    // of the kernel the few that wait for the halo were a 16 us tail (PB_HALO_TIMING: 128 -> 113 us per RHS
    // on 2 B200) -- so they run under the cover of interior B.  Without a peer-memory halo there is nothing
    // to wait for and the order is simply [ river tiles | element tiles ], slow tiles first.
    // (Tried and dropped: only the warps that take a halo-dependent item wait, the clean river tiles behind
    // them -- 1 us faster, but the same-process rank-group tests hung intermittently with it.)
    const bool halo = GH && hw.nn > 0;
    const int R = ntile_r;
    const int Rc = halo ? min(m.ntile_rc, R) : R;
    const int E_int = halo ? min(ntile_int, ntile_e) : 0;
    const int E_a = (int)((long long)E_int * PB_PRE_HALO_FRAC / 100);
    const int E_nog = GH ? min(ntile_int, ntile_e) : ntile_e;      // leading tiles served by the ghost-free code
    // section boundaries: [0,C0) river | [C0,C1) elements | [C1,C2) river | [C2,C3) elements | [C3,C4) elements
    const long long C0 = Rc, C1 = C0 + E_a, C2 = C1 + (R - Rc), C3 = C2 + (ntile_e - E_int), C4 = C3 + (E_int - E_a);
    auto first_q = [&](long long B) { return (B > b) ? (int)((B - b + G - 1) / G) : 0; };   // first ticket with item >= B
    const int qa0 = first_q(C0), qa1 = first_q(C1), qb0 = first_q(C2);
    const int na = qa1 - qa0;                                       // element tickets of interior A
    auto q_of_k = [&](int k) { return (k < na) ? k + qa0 : k - na + qb0; };     // k-th element ticket of this CTA
    auto k_of_q = [&](int q) { return (q < qa1) ? q - qa0 : q - qb0 + na; };
    auto tile_of_g = [&](long long g) -> long long {   // element tile of an element item (ntile_e: past the end)
        return (g < C1) ? g - C0 : (g < C3) ? g - C2 + E_int : (g < C4) ? g - C3 + E_a : (long long)ntile_e;
    };
    auto tile_of_q = [&](int q) { return tile_of_g((long long)q * G + b); };
    ring.init();
    // request the tile of the k-th element ticket (no-op past the end): the static slab does not
    // depend on the previous kernel (prologue), the own surf / gw columns of a tile of owned
    // elements do (after the dependency wait; the arrival travels with this part)
    const int ntile_own = ys ? m.nown / PB_TILE : 0;        // tiles whose 32 elements all live in y
    auto request = [&](int k, bool stat, bool dyn) {
        const long long tile = tile_of_q(q_of_k(k));
        if (tile < ntile_e) {
            const int s = k % PB_PRE_STAGES;
            if (stat) {
                mbar_expect_tx_only(ring.bar(s), SBS);
                tma_bulk_g2s(smem_u32(ring.stage(s)), m.es + ((size_t)tile * TS_NCOL + TS_PRE0) * PB_TILE, SBS, ring.bar(s));
            }
            if (dyn) {
                if (tile < ntile_own) {
                    mbar_expect_tx(ring.bar(s), 2 * PB_TILE * 8);
                    tma_bulk_g2s(smem_u32(ring.stage(s) + SBS), y + tile * PB_TILE, PB_TILE * 8, ring.bar(s));
                    tma_bulk_g2s(smem_u32(ring.stage(s) + SBS + PB_TILE * 8), y + m.o_gw + tile * PB_TILE, PB_TILE * 8, ring.bar(s));
                } else {
                    mbar_expect_tx(ring.bar(s), 0);
                }
            }
        }
    };
    if (lane == 0)
        for (int k = warp; k < PB_PRE_STAGES; k += PB_PRE_WARPS) request(k, true, false);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // y comes from the previous kernel of the stream
#ifdef PB_HALO_TIMING
    if (GH && threadIdx.x == 0) { HT_MIN((int)hput.seq, 0); if (b == 0) g_ht_seq = (int)hput.seq; }
#endif
    if (GH && hput.hp.nn > 0 && b < hput.nput) halo_put(m, y, hput);
    if (lane == 0)
        for (int k = warp; k < PB_PRE_STAGES; k += PB_PRE_WARPS) request(k, false, true);
    const long long g_end = C4;
    bool halo_seen = !halo;
    for (;;) {
        const int q = ring.take(lane);
        const long long g = (long long)q * G + b;
        if (g >= g_end) break;
        if (GH && !halo_seen && g >= C1) {
            // the neighbours' halo records of this RHS have arrived (one lane per neighbour)
            if (lane == 0) { HT_MIN((int)hw.seq, 2); HT_MAX((int)hw.seq, 3); }
            if (lane < hw.nn) {
                long long spins = 0;
                while (hw.flags[hw.rank[lane]] != hw.seq)
                    if (++spins > (1LL << 31)) { atomicOr(m.nan_flag, 2); break; }   // lost neighbour: flag, do not hang
                fence_sys();
            }
            __syncwarp();
            if (lane == 0) HT_MAX((int)hw.seq, 4);
            halo_seen = true;
        }
        if (g < C0 || (g >= C1 && g < C2)) {
            const int t = (int)((g < C0) ? g : g - C1 + Rc);
            const int r = (GH ? m.riv_tile_order[t] : t) * PB_TILE + lane;
            if (r < m.nr) {
                if (GH && t >= m.ntile_rc) river_fluxes<GH>(m, y, r);
                else river_fluxes<false>(m, y, r);          // a clean tile reads owned state only
            }
            continue;
        }
        const long long tile = tile_of_g(g);
        const int k = k_of_q(q), s = k % PB_PRE_STAGES, n = k / PB_PRE_STAGES;
        const unsigned phase = (unsigned)n & 1u;
        ring.acquire(s, n);
        const int i = (int)tile * PB_TILE + lane;
        const double *st = reinterpret_cast<const double *>(ring.stage(s)) + lane;
        if (i < m.ne) {
            const bool tys = tile < ntile_own;
            // (a tile of the interior reads no ghost: the variant without the owned / ghost distinction)
            const bool done = (GH && tile >= E_nog) ? elem_pre<true, GH>(m, y, i, st, ring.bar(s), phase, tys)
                                                    : elem_pre<true, false>(m, y, i, st, ring.bar(s), phase, tys);
            if (!done) elem_pre_exact<GH>(m.self, y, i, st, ring.bar(s), phase, tys);
        } else {
            mbar_wait(ring.bar(s), phase);
        }
        __syncwarp();           // every lane is done with the stage: hand it to the tile STAGES tickets ahead
        if (lane == 0) ring.release(s, n);
        const long long tn = tile_of_q(q_of_k(k + PB_PRE_STAGES));
        if (tn < ntile_e) {
            // lane 0 the static slab, lanes 1 and 2 the own surf / gw columns of a tile of owned elements
            const bool own = tn < ntile_own;
            if (lane == 0) mbar_expect_tx(ring.bar(s), SBS + (own ? 2 * PB_TILE * 8 : 0));
            __syncwarp();
            if (lane < (own ? 3 : 1)) {
                const double *src = (lane == 0) ? m.es + ((size_t)tn * TS_NCOL + TS_PRE0) * PB_TILE
                                                : y + (lane == 1 ? 0 : m.o_gw) + (size_t)tn * PB_TILE;
                const unsigned off = (lane == 0) ? 0 : SBS + (lane - 1) * PB_TILE * 8;
                tma_bulk_g2s(smem_u32(ring.stage(s)) + off, src, (lane == 0) ? SBS : PB_TILE * 8, ring.bar(s));
            }
        }
    }
#ifdef PB_HALO_TIMING
    if (GH && lane == 0) HT_MAX((int)hput.seq, 5);
#endif
}

template <bool FBR, bool GH>
__global__ void __launch_bounds__(MainCfg<FBR>::THREADS, MainCfg<FBR>::MINB)
k_main(const DevMesh m, const double *__restrict__ y, double *__restrict__ dy, int ntile_e, int ntile_r, int ys)
{
    constexpr int SBS = MainCfg<FBR>::SBS, SBF = MainCfg<FBR>::SBF, SBY = MainCfg<FBR>::SBY, SBD = MainCfg<FBR>::SBD;
    constexpr int STAGES = MainCfg<FBR>::STAGES;
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ __align__(128) unsigned char smem[];
    typename MainCfg<FBR>::ring_t ring(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, b = blockIdx.x;
    const int gr = (ntile_r + PB_RING_GROUP - 1) / PB_RING_GROUP;
    const int qe0 = (gr > b) ? (gr - b + G - 1) / G * PB_RING_GROUP : 0;
    ring.init();
    auto item = [&](int q) {
        return ((long long)(q / PB_RING_GROUP) * G + b) * PB_RING_GROUP + q % PB_RING_GROUP;
    };
    // element tiles are walked from the last to the first: k_pre ran in ascending order, so
    // its most recent tiles (columns shared by both kernels, neighbour records, y) are still in L2
    // A stage = the tile's static slab and forcing columns (they do not depend on the kernels
    // before this one: requested in the prologue) + the tile's own state columns and dynamic
    // records (written by the predecessors: requested after the dependency wait).  The arrival
    // travels with the dynamic part; the static part only announces its bytes.
    const unsigned dyn_bytes = ((ys & 1) ? (FBR ? 2 : 1) * PB_TILE * 8 : 0) + ((ys & 2) ? (FBR ? 2 : 1) * PB_TILE * 8 : 0) + SBD;
    auto request = [&](int q, bool stat, bool dyn) {
        const long long k = item(q) - (long long)gr * PB_RING_GROUP;
        if (k < ntile_e) {
            const size_t tile = (size_t)(ntile_e - 1 - k);
            const int s = (q - qe0) % STAGES;
            unsigned char *sp = ring.stage(s);
            if (stat) {
                mbar_expect_tx_only(ring.bar(s), SBS + SBF);
                tma_bulk_g2s(smem_u32(sp), m.es + (tile * TS_NCOL + TS_MAIN0) * PB_TILE, SBS, ring.bar(s));
                tma_bulk_g2s(smem_u32(sp + SBS), m.ft + tile * 4 * PB_TILE, SBF, ring.bar(s));
            }
            if (dyn) {
                sp += SBS + SBF;
                mbar_expect_tx(ring.bar(s), dyn_bytes);
                if (ys & 1) tma_bulk_g2s(smem_u32(sp), y + m.o_gw + tile * PB_TILE, PB_TILE * 8, ring.bar(s));
                if (ys & 2) tma_bulk_g2s(smem_u32(sp + PB_TILE * 8), y + m.o_unsat + tile * PB_TILE, PB_TILE * 8, ring.bar(s));
                if (FBR) {
                    if (ys & 1) tma_bulk_g2s(smem_u32(sp + 2 * PB_TILE * 8), y + m.o_fg + tile * PB_TILE, PB_TILE * 8, ring.bar(s));
                    if (ys & 2) tma_bulk_g2s(smem_u32(sp + 3 * PB_TILE * 8), y + m.o_fu + tile * PB_TILE, PB_TILE * 8, ring.bar(s));
                }
                if (SBD) tma_bulk_g2s(smem_u32(sp + SBY), m.dnb + tile * PB_TILE, SBD, ring.bar(s));
#ifndef PB_NO_SNB_PREFETCH
                // the static neighbour records of this tile's elements are gathered by the tiles
                // around it, which are in flight on other SMs now: have them in L2
                tma_prefetch_l2(m.snb + tile * PB_TILE, PB_TILE * 32);
#endif
            }
        }
    };
    // In the loop below a stage is re-armed by the whole warp: lane c issues copy c of the table
    // (one address computation per lane instead of one copy after the other on lane 0)
    StageCopy *desc = reinterpret_cast<StageCopy *>(smem + MainCfg<FBR>::ring_t::desc_off());
    int ncopy = 0;
    unsigned total_bytes = SBS + SBF + dyn_bytes;
    {
        auto add = [&](const void *base, unsigned stride, unsigned off, unsigned bytes, unsigned kind) {
            if (threadIdx.x == 0) {
                StageCopy c;
                c.base = (unsigned long long)base; c.stride = stride; c.off = off;
                c.bytes = bytes; c.kind = kind; c.pad0 = c.pad1 = 0;
                desc[ncopy] = c;
            }
            ncopy++;
        };
        add(m.es + (size_t)TS_MAIN0 * PB_TILE, TS_NCOL * PB_TILE * 8, 0, SBS, 0);
        add(m.ft, 4 * PB_TILE * 8, SBS, SBF, 0);
        if (ys & 1) add(y + m.o_gw, PB_TILE * 8, SBS + SBF, PB_TILE * 8, 0);
        if (ys & 2) add(y + m.o_unsat, PB_TILE * 8, SBS + SBF + PB_TILE * 8, PB_TILE * 8, 0);
        if (FBR && (ys & 1)) add(y + m.o_fg, PB_TILE * 8, SBS + SBF + 2 * PB_TILE * 8, PB_TILE * 8, 0);
        if (FBR && (ys & 2)) add(y + m.o_fu, PB_TILE * 8, SBS + SBF + 3 * PB_TILE * 8, PB_TILE * 8, 0);
        if (SBD) add(m.dnb, PB_TILE * 32, SBS + SBF + SBY, SBD, 0);
#ifndef PB_NO_SNB_PREFETCH
        add(m.snb, PB_TILE * 32, 0, PB_TILE * 32, 1);
#endif
    }
    if (lane == 0)
        for (int k = warp; k < STAGES; k += MainCfg<FBR>::WARPS) request(qe0 + k, true, false);
    // the static slabs above do not depend on k_pre; everything below does (no-op when the
    // kernel was not launched with the programmatic-serialization attribute)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef PB_HALO_TIMING
    const int ht_seq = g_ht_seq;
    if (GH && threadIdx.x == 0) HT_MIN(ht_seq, 6);
#endif
    if (lane == 0)
        for (int k = warp; k < STAGES; k += MainCfg<FBR>::WARPS) request(qe0 + k, false, true);
    __syncthreads();        // the descriptor table is visible to every warp
    const long long r_end = (long long)gr * PB_RING_GROUP, e_end = r_end + ntile_e;
    for (;;) {
        const int q = ring.take(lane);
        const long long g = item(q);
        if (g >= e_end) break;
        if (g < r_end) {
            const int r = (int)g * PB_TILE + lane;
            if (r < m.rown) river_main(m, dy, r);
            continue;
        }
        const int k = q - qe0, s = k % STAGES, n = k / STAGES;
        const unsigned phase = (unsigned)n & 1u;
        ring.acquire(s, n);
        const int i = (ntile_e - 1 - (int)(g - r_end)) * PB_TILE + lane;
        const double *st = reinterpret_cast<const double *>(ring.stage(s)) + lane;
        const double *f = reinterpret_cast<const double *>(ring.stage(s) + SBS) + lane;
        if (i < m.nown) {
            if (!elem_main<FBR, true, GH>(m, y, dy, i, st, f, ring.bar(s), phase, ys,
                                      smem + MainCfg<FBR>::RING_BYTES + warp * MainGather<FBR>::WARP))
                elem_main_exact<FBR, GH>(m.self, y, dy, i, st, f, ring.bar(s), phase, ys);
        } else {
            mbar_wait(ring.bar(s), phase);
        }
        __syncwarp();           // every lane is done with the stage: hand it to the tile STAGES tickets ahead
        if (lane == 0) ring.release(s, n);
        const long long kk = item(q + STAGES) - r_end;
        if (kk < ntile_e) {
            const unsigned tile = (unsigned)(ntile_e - 1 - (int)kk);
            if (lane == 0) mbar_expect_tx(ring.bar(s), total_bytes);
            __syncwarp();
            if (lane < ncopy) {
                const StageCopy d = desc[lane];
                const void *src = reinterpret_cast<const void *>(d.base + (unsigned long long)tile * d.stride);
                if (d.kind == 0) tma_bulk_g2s(smem_u32(ring.stage(s)) + d.off, src, d.bytes, ring.bar(s));
                else tma_prefetch_l2(src, d.bytes);
            }
        }
    }
#ifdef PB_HALO_TIMING
    if (GH && lane == 0) HT_MAX(ht_seq, 7);
#endif
}

// One-off: the tiles' RAREA / RDIST* slots.  The packer left the neighbour distances in RDIST*: they move to
// the cold table, and the slot gets the refined reciprocal (of 1.0 on an edge without an element behind it:
// LateralFlow's straight-line block sees the element itself there), NaN outside Arith<true>::rcp's domain.
static __global__ void k_tile_rcp(double *es, double *dist_cold, int nes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nes) return;
    double *tile = es + (size_t)(i >> 5) * TS_NCOL * PB_TILE + (i & 31);
    const int *nbs = reinterpret_cast<const int *>(es + ((size_t)(i >> 5) * TS_NCOL + TS_NB0) * PB_TILE);
    tile[TS_RAREA * PB_TILE] = rcp_or_nan(tile[TS_AREA * PB_TILE]);
    for (int j = 0; j < 3; j++) {
        const double d = tile[(TS_RDIST0 + j) * PB_TILE];
        dist_cold[(size_t)j * nes + i] = d;
        tile[(TS_RDIST0 + j) * PB_TILE] = rcp_or_nan((nbs[j * PB_TILE + (i & 31)] >= 0) ? d : 1.0);
    }
}

// One-off: the refined reciprocals of the dictionary's divisors and of the model constants
// (same instruction sequence as Arith<true>::rcp, so div(a, b, r) stays bitwise `a / b`)
static __global__ void k_class_rcp(double *cls, int ncls, double dt, double *consts)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncls) {
        double *row = cls + (size_t)c * CC_STRIDE;
        row[CC_RALPHA] = rcp_or_nan(row[CC_ALPHA]);
        row[CC_RPOR] = rcp_or_nan(row[CC_POROSITY]);
        row[CC_RGALPHA] = rcp_or_nan(row[CC_GALPHA]);
        row[CC_RGPOR] = rcp_or_nan(row[CC_GPOROSITY]);
    }
    if (c == 0) {
        consts[0] = rcp_or_nan(PB_DEPRSTG);
        consts[1] = rcp_or_nan(dt);
    }
}

// Halo pack: the owned values other ranks need, as records in their ghost order
//   elements {surf, gw[, fbr_gw]} (gs doubles), rivers {stage, gw}
static __global__ void __launch_bounds__(256)
k_halo_pack(const DevMesh m, const double *__restrict__ y, int nse, const int *__restrict__ send_e,
            double *__restrict__ buf_e, int nsr, const int *__restrict__ send_r, double *__restrict__ buf_r)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nse) {
        const int i = send_e[k];
        buf_e[(size_t)k * m.gs] = y[i];
        buf_e[(size_t)k * m.gs + 1] = y[m.o_gw + i];
        if (m.gs == 3) buf_e[(size_t)k * m.gs + 2] = y[m.o_fg + i];
    } else if (k < nse + nsr) {
        const int r = send_r[k - nse];
        buf_r[(size_t)(k - nse) * 2] = y[m.o_stg + r];
        buf_r[(size_t)(k - nse) * 2 + 1] = y[m.o_rgw + r];
    }
}

// test hook: pow_pos() against libdevice pow() (tests/test_fastpow_gpu.py)
static __global__ void k_test_pow(int n, const double *__restrict__ x, const double *__restrict__ y,
                                  double *__restrict__ fast, double *__restrict__ ref)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a, b;
    pow_pos2(x[i], y[i], x[n - 1 - i], y[n - 1 - i], a, b);
    fast[i] = a;
    if (b != pow_pos(x[n - 1 - i], y[n - 1 - i]) && !(b != b)) fast[i] = -1.0;   // pair form == single form
    ref[i] = pow(x[i], y[i]);
}

// test hook: Arith<true> division against the hardware `/` (tests/test_fastpow_gpu.py);
// ok[i] = 0 where the operands left the branch-free domain (the kernels then recompute)
static __global__ void k_test_div(int n, const double *__restrict__ a, const double *__restrict__ b,
                                  double *__restrict__ fast, double *__restrict__ ok, double *__restrict__ ref)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Arith<true> A;
    const double q = A.quo(a[i], b[i]);
    const bool ok1 = A.ok;
    // shared-reciprocal form for positive divisors must agree with the single form
    const double bp = fabs(b[i]);
    const double r = A.rcp(bp);
    const double q2 = A.div(a[i], bp, r), q3 = A.div(a[n - 1 - i], bp, r);
    fast[i] = q;
    ok[i] = ok1 ? 1.0 : 0.0;
    ref[i] = a[i] / b[i];
    const double a2 = a[n - 1 - i];
    const bool dom = (fabs(a[i]) >= 1e-291 || a[i] == 0.0) && (fabs(a2) >= 1e-291 || a2 == 0.0);   // |a| >= 2^-969
    if (A.ok && dom && (q2 != a[i] / bp || q3 != a2 / bp) && q2 == q2 && q3 == q3) ok[i] = -1.0;
    Arith<true> S;
    const double sq = S.sqrtp(bp);
    if (S.ok && sq != sqrt(bp)) ok[i] = -2.0;
    if (!S.ok && bp >= 1e-290 && bp < 1e300) ok[i] = -3.0;
}

#undef TSC
#undef RFC
#undef RIC
#undef FOC
#undef RFLX
#undef XFC

}  // namespace pb
