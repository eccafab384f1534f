// summary.cuh -- Summary() + MassBalance() on the device.  Compiled with
// -fmad=false (pihm_b200.cu): equal inputs give the reference's bits.
#pragma once
#include "common.cuh"
#include "rhs_layout.cuh"

namespace pb {

#define TSC(slot, i) (m.es[((size_t)((i) >> 5) * TS_NCOL + (slot)) * PB_TILE + ((i) & 31)])
#define CLE(c, i) (m.cls[(size_t)m.cid[i] * CC_STRIDE + (c)])
#define XFC(c, i) (m.xflux[(size_t)(c) * m.nes + (i)])

// Summary() + MassBalance() (src/update.c:3-160): ws = y (not clamped), the mass-balance
// infiltration from the change of soil storage and the wf.* fields of the last RHS call
// (PB_X_* columns), subrunoff, fbr: wf.fbr_infil.  One thread per owned element; plain
// IEEE operators (-fmad=false), so equal inputs give the reference's bits.  ws0 itself is
// replaced by y afterwards (a copy on the same stream).
static __global__ void __launch_bounds__(256)
k_summary_mb(const DevMesh m, const double *__restrict__ y, const double *__restrict__ ws0,
             double *__restrict__ subrunoff_out, double stepsize)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nown) return;
    const double depth = TSC(TS_DEPTH, i), area = TSC(TS_AREA, i);
    // update.c:122-128
    double soilw0 = ws0[m.o_gw + i] + ws0[m.o_unsat + i];
    soilw0 = (soilw0 > depth) ? depth : soilw0;
    soilw0 = (soilw0 < 0.0) ? 0.0 : soilw0;
    double soilw1 = y[m.o_gw + i] + y[m.o_unsat + i];
    soilw1 = (soilw1 > depth) ? depth : soilw1;
    soilw1 = (soilw1 < 0.0) ? 0.0 : soilw1;
    // update.c:130-136
    double subrunoff = 0.0;
#pragma unroll
    for (int j = 0; j < 3; j++) subrunoff += XFC(PB_X_SUB0 + j, i) / area;
    double infil = (soilw1 - soilw0) * CLE(CC_POROSITY, i) / stepsize + subrunoff + XFC(PB_X_EDIR_UNSAT, i) +
        XFC(PB_X_EDIR_GW, i) + XFC(PB_X_ETT_UNSAT, i) + XFC(PB_X_ETT_GW, i);
    if (m.fbr) {
        // update.c:138-152
        const double fbrw0 = ws0[m.o_fg + i] + ws0[m.o_fu + i];
        const double fbrw1 = y[m.o_fg + i] + y[m.o_fu + i];
        double fbrrunoff = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) fbrrunoff += XFC(PB_X_FBRFLOW0 + j, i) / area;
        const double fbr_infil = (fbrw1 - fbrw0) * CLE(CC_GPOROSITY, i) / stepsize + fbrrunoff;
        XFC(PB_X_FBR_INFIL, i) = fbr_infil;
        infil += fbr_infil;
    }
    // update.c:154-158
    if (infil < 0.0) {
        subrunoff -= infil;
        infil = 0.0;
    }
    XFC(PB_X_INFIL, i) = infil;
    subrunoff_out[i] = subrunoff;
}

#undef TSC
#undef CLE
#undef XFC

}  // namespace pb
