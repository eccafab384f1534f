// common.cuh -- shared declarations of libpihm_b200 (host + device).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include "pihm_b200.h"

namespace pb {

// ---- error plumbing --------------------------------------------------------
void set_error(const std::string &msg);

#define PB_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            pb::set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
            return -1;                                                         \
        }                                                                      \
    } while (0)

#define PB_CUDA_PTR(call)                                                      \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            pb::set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
            return nullptr;                                                    \
        }                                                                      \
    } while (0)

// Every state-sized device vector is allocated with PB_VEC_PAD extra doubles: the RHS kernels
// fetch the own-state columns of a 32-element tile with bulk copies of a whole tile, which for
// the last (partial) tile of a block reach past the end of the vector (read only, never used).
#define PB_VEC_PAD 64

// ---- physical constants (src/include/pihm_const.h:7,77-82) -----------------
#define PB_GRAV 9.80665
#define PB_PSIMIN (-70.0)
#define PB_DEPRSTG 1.0E-4
#define PB_GRADMIN 5.0E-8
#define PB_SATMIN 0.1
#define PB_RIVGRADMIN 0.05
#define PB_KINEMATIC 1
#define PB_DIFF_WAVE 2

// rivflow slots (pihm_const.h:130-140)
enum { RF_UP_C2C = 0, RF_DOWN_C2C, RF_LEFT_S2C, RF_RIGHT_S2C, RF_LEFT_A2C,
       RF_RIGHT_A2C, RF_CHANL_LKG, RF_LEFT_A2A, RF_RIGHT_A2A, RF_DOWN_A2A,
       RF_UP_A2A };

// Neighbour code of an element edge on the device:
//   code >= 0  : neighbour element (internal index)
//   code == -1 : domain boundary
//   code <= -2 : river edge; c = -code-2; river = c >> 2; side = c & 3
//                (0 = this element is the river's left bank on this edge,
//                 1 = right bank, 2 = edge never written by RiverToElem)
#define PB_NB_BOUNDARY (-1)
__host__ __device__ inline int nb_river_code(int river, int side) { return -((river << 2 | side) + 2); }

// Device view of one model: everything the RHS kernels read.  Passed by value.
struct DevMesh {
    int ne, nr;          // local elements / river segments (owned + ghosts)
    int nown, rown;      // owned ones come first; single GPU: nown == ne, rown == nr
    int gs;              // doubles per ghost element record (2, or 3 with fbr)
    const double *gel;   // [ne - nown][gs] ghost element records {surf, gw[, fbr_gw]}
    const double *gri;   // [nr - rown][2]  ghost river records {stage, gw}
    int nes, nrs;        // column strides (padded)
    int fbr, surf_mode, riv_mode;
    int record;          // write the PB_X_* flux columns
    int replay;          // re-evaluation of the last call for its fluxes: keep s2c_stale as it is
    double dt;
    double r_deprstg, r_dt;   // refined reciprocals of DEPRSTG and dt (k_class_rcp)
    // offsets of the state blocks inside y / ydot (pihm_func.h:7-15)
    long long o_unsat, o_gw, o_stg, o_rgw, o_fu, o_fg;
    const double *es;    // [ntile][PB_E_NCOL][32] static element columns, warp-tiled
    const double *ft;    // [ntile][4][32]     hot forcing columns (pcpdrp, edir, ett, ws0.surf), warp-tiled
    const double4 *snb;  // [nes] static neighbour record  {zmin, zmax, rough, zbed}
    double4 *dnb;        // [nes] dynamic neighbour record {surfh, EffKh, |grad h|, (surfh-D)^(2/3)}, written by k_pre
    const double *cls;   // [nclass][CC_STRIDE] dictionary of the soil / land-cover / geology parameter rows
    const int *cid;      // [nes] class id of each element (also inside the tile slab)
    const int *bct;      // [3][nes]           bc_type
    const int *fbct;     // [3][nes]           fbrbc_type
    const double *forc;  // [PB_F_NCOL][nes]   flat forcing table; the kernels read only its bc columns
    const double *rf;    // [PB_R_NCOL][nrs]   static river columns
    const int *ri;       // [PB_RI_NCOL][nrs]  LEFT/RIGHT = internal element idx
    const double *rivbc; // [nrs]
    const double *fbr_dist;  // [nrs] nabrdist(left bank) + nabrdist(right bank)
    const int *riv_tile_order;   // [ceil(nr / 32)] river tiles, those that read no ghost state first (partitions)
    int ntile_rc;                //   ... and how many of them there are
    const double *dist_cold; // [3][nes] nabrdist (boundary edges / exact path; the tiles carry 1 / nabrdist)
    const int *up_ptr;   // [nr+1] CSR of upstream segments, ascending index
    const int *up_idx;
    double *rivflow;     // [11][nrs]
    double *s2c_stale;   // [2][nrs]  rivflow[LEFT/RIGHT_S2C] of the previous call
    double *xflux;       // [PB_X_NCOL][nes] (record != 0)
    int *nan_flag;
    unsigned long long *slow_count;   // elements recomputed by elem_main_exact (diagnostic)
    const DevMesh *self;              // device copy of this struct (for non-inlined rare paths)
};

}  // namespace pb

namespace pb {
// System-scope fence between a result and the flag / ticket that announces it (peer GPUs over NVLink, the
// host over a mapped page).  __threadfence_system() is membar.sys = fence.sc.sys (MEMBAR.SC.SYS), measured at
// ~6.5 us per executing warp inside the busy RHS kernel (tools/mgpu_rhs_probe.py, PB_HALO_TIMING).
// PB_SYS_FENCE=1 builds the release / acquire form (fence.acq_rel.sys, MEMBAR.ALL.SYS) instead: no measurable
// difference on one GPU (8.34 vs 8.44 ms per step), and the same-process rank-group tests
// (tests/test_localgroup_gpu.py) intermittently lost a neighbour's flag with it -- so the sc fences stay.
#ifndef PB_SYS_FENCE
#define PB_SYS_FENCE 0
#endif
__device__ __forceinline__ void fence_sys()
{
#if PB_SYS_FENCE == 1
    asm volatile("fence.acq_rel.sys;" ::: "memory");
#elif PB_SYS_FENCE == 2       // timing experiment only: no fence at all
    asm volatile("" ::: "memory");
#else
    __threadfence_system();
#endif
}
// ... and the device-scope one between a block's partial result and the counter that elects the last block
// (the threadFenceReduction pattern): release / acquire as well (MEMBAR.ALL.GPU instead of MEMBAR.SC.GPU)
__device__ __forceinline__ void fence_gpu()
{
#if PB_SYS_FENCE == 0
    __threadfence();
#else
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
}  // namespace pb

namespace pb {
// Peer-memory halo exchange (partitioned run): what k_halo_put needs to store this rank's
// boundary states straight into its neighbours' ghost buffers, and what k_pre waits for.
#define PB_MAX_NBR 8
struct HaloPeers {
    double *base[PB_MAX_NBR];        // neighbour k's halo buffer, mapped into this process
    long long pstride[PB_MAX_NBR];   // doubles per parity copy of that buffer
    long long gel_off[PB_MAX_NBR];   // where this rank's element records start in it
    long long gri_off[PB_MAX_NBR];   // ... river records
    long long flag_off[PB_MAX_NBR];  // arrival flags [2][PB_MAX_RANKS_H]
    int e_ptr[PB_MAX_NBR + 1], r_ptr[PB_MAX_NBR + 1];   // send segments per neighbour
    int nn, myrank;
};
struct HaloWait {
    const volatile double *flags;    // this rank's arrival flags of the current parity
    int nn;
    int rank[PB_MAX_NBR];
    double seq;
};
#define PB_MAX_RANKS_H 8
}  // namespace pb

struct pihm_b200_ctx;
struct pihm_b200_cvode;
namespace pb {
// ranks that live in one process (pihm_b200_comm_init_local): exchange buffers as plain pointers
struct LocalGroup {
    int n = 0;
    pihm_b200_ctx *ctx[8] = {};
    double *xbuf[8] = {};            // integrator exchange buffers (cvode_b200.cu), by rank
    double **peer_tab[8] = {};       // each integrator's device table of those
};
// comm.cu: NCCL (dlopen'ed) halo exchange and scalar all-reduce
int comm_halo_exchange(pihm_b200_ctx *ctx);
int comm_allreduce(pihm_b200_ctx *ctx, double *dev_ptr, int count, int op /*0 sum, 1 min, 2 max*/);
void comm_destroy(pihm_b200_ctx *ctx);
int comm_share_buffer(pihm_b200_ctx *ctx, void *local, void **peers /*[nranks]*/);
void comm_unshare_buffer(pihm_b200_ctx *ctx, void **peers);
int comm_setup_halo_p2p(pihm_b200_ctx *ctx);
int comm_setup_halo_local(pihm_b200_ctx **ctxs, int n);
}  // namespace pb

namespace pb {
// rhs_kernels.cu: the RHS kernels live in their own translation unit (own -fmad setting)
int rhs_configure(pihm_b200_ctx *ctx, int sms);
int rhs_class_rcp(pihm_b200_ctx *ctx);
int rhs_tile_rcp(pihm_b200_ctx *ctx);
int rhs_halo_pack(pihm_b200_ctx *ctx, const double *y);
int rhs_launch(pihm_b200_ctx *ctx, const double *y, double *dy, bool replay);
int rhs_halo_times(unsigned long long *out, int reset);
}  // namespace pb

namespace pb {
// one varctrl_struct of the reference's print system (pihm_struct.h:193-216) on the device
struct PrintVar {
    int src = 0, col = 0;      // pihm_b200_print_src, column within it
    int len = 0;               // owned elements or river segments
    int is_river = 0;
    double *acc = nullptr;     // running sum (varctrl.buffer), internal order
    int counter = 0;           // varctrl.counter
};
}  // namespace pb

// opaque handle types of the C ABI
struct pihm_b200_ctx {
    pb::DevMesh dm{};
    int device = 0;
    int reorder = 0;
    cudaStream_t stream = nullptr;         // owned stream
    cudaStream_t user_stream = nullptr;    // caller's stream (pihm_b200_set_stream)
    int use_user_stream = 0;
    cudaStream_t s() const { return use_user_stream ? user_stream : stream; }
    int64_t nsv = 0;
    int rhs_launches = 0;              // kernels launched per RHS call
    int pre_grid = 0, main_grid = 0;   // persistent RHS grids (SMs x resident CTAs)
    int pre_smem = 0, main_smem = 0;   // dynamic shared memory of the stage rings
    cudaAccessPolicyWindow l2_window{}; // persisting-L2 window over the neighbour records
    int l2_on = 0;
    int pdl = 1;                       // launch k_main with programmatic stream serialization
    int ystage = 1;                    // own-state columns of a tile travel with the TMA stage
    // host copies kept for permutation / validation
    std::vector<int> perm;             // internal element -> reference element
    std::vector<int> iperm;            // reference element -> internal element
    std::vector<int> riv_left_edge, riv_right_edge;   // edge slot of each bank
    // device allocations
    double *d_dist_cold = nullptr;
    int *d_riv_tile_order = nullptr;
    double *d_es = nullptr, *d_ft = nullptr, *d_forc = nullptr, *d_rf = nullptr, *d_rivbc = nullptr;
    double4 *d_snb = nullptr, *d_dnb = nullptr;
    double *d_cls = nullptr;
    int *d_cid = nullptr, nclass = 0;
    double *d_fbr_dist = nullptr;
    int *d_bct = nullptr, *d_fbct = nullptr, *d_ri = nullptr;
    int *d_up_ptr = nullptr, *d_up_idx = nullptr;
    double *d_rivflow = nullptr, *d_stale = nullptr, *d_xflux = nullptr;
    int *d_nan = nullptr;
    unsigned long long *d_slow = nullptr;
    pb::DevMesh *d_dm = nullptr;
    int *d_perm = nullptr, *d_iperm = nullptr;   // device copies (state gather)
    // multi-GPU: ghost buffers, send lists, communicator (comm.cu)
    double *d_gel = nullptr, *d_gri = nullptr, *d_send_e = nullptr, *d_send_r = nullptr;
    int *d_send_e_idx = nullptr, *d_send_r_idx = nullptr;
    std::vector<int> nbr_rank, send_e_ptr, recv_e_cnt, send_r_ptr, recv_r_cnt;
    int nse = 0, nsr = 0;
    // peer-memory halo exchange (comm.cu: comm_setup_halo_p2p)
    int halo_p2p = 0;
    int ntile_int = 0;                 // leading element tiles that read no ghost in k_pre (interior of the partition)
    double *d_hx = nullptr;            // [2][gel | gri] + flags, mapped into the neighbours
    long long hx_stride = 0;           // doubles per parity copy
    void *hx_peer[8] = {};             // all ranks' buffers mapped here (PB_MAX_RANKS_H)
    pb::HaloPeers hpeers{};
    pb::DevMesh *d_dm_par[2] = {nullptr, nullptr};   // device copies of dm with the parity's ghost pointers
    unsigned int *d_hcount = nullptr;
    long long halo_seq = 0;
    void *comm = nullptr;              // ncclComm_t
    std::shared_ptr<pb::LocalGroup> lgroup;   // same-process rank group (no NCCL)
    int rank = 0, nranks = 1;
    long long nsv_global = 0;
    // staging for host <-> device vectors
    double *d_stage = nullptr;         // nsv (reference order)
    pihm_b200_vec *y_tmp = nullptr, *yd_tmp = nullptr;
    // reduction scratch
    double *d_red = nullptr;           // partial sums
    double *h_red = nullptr;           // pinned scalars
    int red_blocks = 0;
    long long launches = 0;            // total kernel launches (for gpu_launches)
    // Summary()/MassBalance() on the device (pihm_b200_set_diagnostics): the wf.* fields of the
    // LAST RHS call are produced on demand by evaluating that call once more with record = 1.
    int diag = 0;
    const double *last_in = nullptr;   // input vector of the last RHS call (device pointer)
    int flux_fresh = 0;                // d_xflux holds the fluxes of the last RHS call
    double *d_last_snap = nullptr;     // copy of *last_in, taken just before that vector is overwritten
    double *d_rec_dy = nullptr;        // ydot scratch of the re-evaluation
    double *d_ws0 = nullptr;           // [nsv] ws0 of elements and rivers, block layout of y
    double *d_subrunoff = nullptr;     // [nes] MassBalance's subrunoff (update.c:128-133,154-158)
    pb::DevMesh *d_dm_rec = nullptr;   // device copy of the record/replay view (rare exact paths)
    // interception / snow / ET on the device (pihm_b200_et_create)
    double *d_etf = nullptr;           // [PB_ET_NCOL][nes] static columns, internal order
    int *d_eti = nullptr;              // [PB_ETI_NCOL][nes]
    double *d_eto = nullptr;           // [PB_EO_NCOL][nes] outputs + the two storages
    double *d_et_tab = nullptr;        // per-step tables by type: meteo | lai | lai_lc | z0_lc
    double *h_et_tab = nullptr;        // pinned staging of the same
    cudaEvent_t et_ev = nullptr;       // the last upload from h_et_tab has finished
    size_t et_tab_cap = 0;             // doubles
    int et_max_meteo = 0, et_max_lai = 0, et_max_lc = 0;   // largest type index used by an element
    std::vector<pb::PrintVar> pvars;   // pihm_b200_print_add
};

namespace pb {
inline void snapshot_last_in(pihm_b200_ctx *ctx)
{
    cudaMemcpyAsync(ctx->d_last_snap, ctx->last_in, sizeof(double) * ctx->nsv, cudaMemcpyDeviceToDevice, ctx->s());
    ctx->last_in = ctx->d_last_snap;
}
// Call BEFORE launching anything that overwrites device vector p: if p is the input of the last
// RHS call and the diagnostics are on, its contents are saved first (stream order).
inline void note_write(pihm_b200_ctx *ctx, const double *p)
{
    if (ctx->diag && p == ctx->last_in) snapshot_last_in(ctx);
}
// ... and before freeing it: without diagnostics there is nowhere to keep the contents, so the
// library simply forgets the input (a later pihm_b200_summary_mb then asks for an RHS call first)
inline void note_free(pihm_b200_ctx *ctx, const double *p)
{
    if (p != ctx->last_in) return;
    if (ctx->diag) snapshot_last_in(ctx);
    else ctx->last_in = nullptr;
}
}  // namespace pb

struct pihm_b200_vec {
    pihm_b200_ctx *ctx = nullptr;
    double *d = nullptr;
    int64_t n = 0;
    bool owns = true;
};
