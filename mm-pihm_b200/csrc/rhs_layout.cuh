// rhs_layout.cuh -- device data layout of the element tables (host packer + kernels).
#pragma once
#include "common.cuh"

namespace pb {

// Static element columns are stored warp-tiled ("AoSoA-32"): tile t = i >> 5
// holds all columns of its 32 elements contiguously, 256 B per column, in the
// slot order below; the three neighbour codes of the tile (int32 [3][32]) ride
// along as two pseudo-columns.  One warp works on one tile at a time.  Both RHS
// kernels are persistent: a CTA owns a ring of PB_*_STAGES shared-memory stages,
// each filled by ONE bulk async copy (cp.async.bulk, the TMA engine; SASS UBLKCP)
// of the contiguous slot range the kernel needs (+ one for the forcing columns),
// completion signalled on the stage's mbarrier.  Warps take tiles by ticket; the
// warp that finishes a tile re-arms its stage with the tile PB_*_STAGES tickets
// ahead, so (STAGES - WARPS) tile loads per CTA are always in flight while the
// warps compute -- the loads no longer sit at the head of every warp's
// dependency chain (ncu r01: both kernels were latency-, not bandwidth-bound).
#define PB_TILE 32
#ifndef PB_PATCH
#define PB_PATCH 128       // elements per locality patch of the internal ordering (reorder.h)
#endif
// Soil / land-cover / geology parameters are CLASS data in MM-PIHM (soil, lc and geol tables
// indexed by type; ReadSoil/ReadLc/ReadGeol, calibrated by global multipliers): the create
// call deduplicates their rows into a dictionary `cls[ncls][CC_STRIDE]` (a few dozen rows, L1
// resident) and the tiles carry a 4-byte class id per element instead of 12 (fbr: 17) doubles.
// The dictionary also holds the van Genuchten exponents derived from beta, computed once on
// the host with the same IEEE divisions.  dmac / dinf stay per element (clipped to the soil
// depth, init_soil.c:57-59).  A table without repeated rows still works (one class per element).
enum {   // k_pre reads [TS_PRE0, TS_PRE1), k_main [TS_MAIN0, TS_MAIN1 / TS_FBR1)
    TS_NABRX0 = 0, TS_NABRX1, TS_NABRX2, TS_NABRY0, TS_NABRY1, TS_NABRY2,
    TS_ZMAX, TS_DEPTH, TS_DMAC, TS_NB0, TS_NB1,     // NB0/NB1: int32 [4][32] = 3 neighbour codes + class id
    // RAREA / RDIST*: refined reciprocals of the area and of the three neighbour distances (1.0 on an edge
    // without an element behind it), filled on the device at create time (k_tile_rcp); the distances
    // themselves are needed on boundary edges and by the exact path only: cold table DevMesh::dist_cold
    TS_AREA, TS_RAREA, TS_DINF, TS_ZMIN, TS_EDGE0, TS_EDGE1, TS_EDGE2, TS_RDIST0, TS_RDIST1, TS_RDIST2,
    TS_ZBED, TS_GDEPTH,
    TS_NCOL,
    TS_PRE0 = TS_NABRX0, TS_PRE1 = TS_AREA, TS_MAIN0 = TS_ZMAX, TS_MAIN1 = TS_ZBED, TS_FBR1 = TS_NCOL
};
enum {   // dictionary row; pairs are fetched as double2
    CC_ALPHA = 0, CC_M1, CC_M2, CC_M3,              // m1 = beta/(beta-1), m2 = (beta-1)/beta, m3 = 1/beta
    CC_KINFV, CC_KMACV, CC_AREAFH, CC_KSATV,
    CC_POROSITY, CC_ROUGH, CC_RZD, CC_BETA,
    CC_KMACH, CC_AREAFV, CC_KSATH, CC_RALPHA,       // R*: refined reciprocals (k_class_rcp), shared by divisions
    CC_GALPHA, CC_GM1, CC_GM2, CC_GM3,
    CC_GKSATV, CC_GKSATH, CC_GPOROSITY, CC_GBETA,
    CC_RPOR, CC_RGALPHA, CC_RGPOR, CC_PAD,
    CC_STRIDE
};
// ABI column (include/pihm_b200.h) -> tile slot; used by the host packer
__host__ __device__ inline int tile_slot_of(int abi_col)
{
    switch (abi_col) {
        case PB_E_AREA: return TS_AREA;  case PB_E_ZMIN: return TS_ZMIN;  case PB_E_ZMAX: return TS_ZMAX;
        case PB_E_ZBED: return TS_ZBED;
        case PB_E_EDGE0: return TS_EDGE0; case PB_E_EDGE1: return TS_EDGE1; case PB_E_EDGE2: return TS_EDGE2;
        case PB_E_NABRDIST0: return TS_RDIST0; case PB_E_NABRDIST1: return TS_RDIST1;    // the packer stores the
        case PB_E_NABRDIST2: return TS_RDIST2;      // distance; k_tile_rcp moves it to dist_cold and leaves 1 / d
        case PB_E_NABRX0: return TS_NABRX0; case PB_E_NABRX1: return TS_NABRX1; case PB_E_NABRX2: return TS_NABRX2;
        case PB_E_NABRY0: return TS_NABRY0; case PB_E_NABRY1: return TS_NABRY1; case PB_E_NABRY2: return TS_NABRY2;
        case PB_E_DEPTH: return TS_DEPTH; case PB_E_DINF: return TS_DINF; case PB_E_DMAC: return TS_DMAC;
        case PB_E_GDEPTH: return TS_GDEPTH;
    }
    return -1;
}
// ABI column -> dictionary slot (-1: per-element column)
__host__ __device__ inline int class_slot_of(int abi_col)
{
    switch (abi_col) {
        case PB_E_ALPHA: return CC_ALPHA; case PB_E_BETA: return CC_BETA; case PB_E_KINFV: return CC_KINFV;
        case PB_E_KMACV: return CC_KMACV; case PB_E_AREAFH: return CC_AREAFH; case PB_E_KSATV: return CC_KSATV;
        case PB_E_POROSITY: return CC_POROSITY; case PB_E_ROUGH: return CC_ROUGH; case PB_E_RZD: return CC_RZD;
        case PB_E_KMACH: return CC_KMACH; case PB_E_AREAFV: return CC_AREAFV; case PB_E_KSATH: return CC_KSATH;
        case PB_E_GALPHA: return CC_GALPHA; case PB_E_GBETA: return CC_GBETA; case PB_E_GKSATV: return CC_GKSATV;
        case PB_E_GKSATH: return CC_GKSATH; case PB_E_GPOROSITY: return CC_GPOROSITY;
    }
    return -1;
}
}  // namespace pb
