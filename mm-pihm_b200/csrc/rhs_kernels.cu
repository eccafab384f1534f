// rhs_kernels.cu -- the translation unit of the two RHS kernels (rhs.cuh), their launcher and
// their configuration.  Compiled apart from the rest of the library because the element
// physics may be built with FMA contraction (csrc/Makefile: RHS_FMAD), while the N_Vector /
// integrator / Summary / ET kernels keep -fmad=false (they reproduce the reference's bits).
#include <algorithm>
#include <string>
#include "common.cuh"
#include "rhs.cuh"

namespace pb {

// persistent RHS kernels: opt in to the ring's dynamic shared memory, size the grids to what
// is resident at once (SMs x CTAs per SM)
int rhs_configure(pihm_b200_ctx *ctx, int sms)
{
    const DevMesh &dm = ctx->dm;
    const int pre_smem = PreCfg::ring_t::smem_bytes();
    const int main_smem = dm.fbr ? MainCfg<true>::SMEM : MainCfg<false>::SMEM;
    int bpre = 0, bmain = 0;
    // GH = ghost entities present (a partition): the variant without them skips the owned /
    // ghost distinction in every state access
    const bool gh = (dm.nown != dm.ne) || (dm.rown != dm.nr);
    const void *kp = gh ? (const void *)k_pre<true> : (const void *)k_pre<false>;
    const void *km = dm.fbr ? (gh ? (const void *)k_main<true, true> : (const void *)k_main<true, false>)
                            : (gh ? (const void *)k_main<false, true> : (const void *)k_main<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, pre_smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, main_smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpre, kp, PB_PRE_THREADS, pre_smem);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bmain, km, dm.fbr ? MainCfg<true>::THREADS : MainCfg<false>::THREADS,
                                                          main_smem);
    if (e != cudaSuccess || bpre < 1 || bmain < 1) {
        set_error(std::string("pihm_b200_create: RHS kernel configuration failed: ") + cudaGetErrorString(e));
        return -1;
    }
    if (const char *ov = std::getenv("PIHM_B200_PRE_CTAS")) bpre = std::max(1, std::min(bpre, std::atoi(ov)));
    if (const char *ov = std::getenv("PIHM_B200_MAIN_CTAS")) bmain = std::max(1, std::min(bmain, std::atoi(ov)));
    ctx->pre_grid = sms * bpre;
    ctx->main_grid = sms * bmain;
    ctx->pre_smem = pre_smem;
    ctx->main_smem = main_smem;
    ctx->ystage = std::getenv("PIHM_B200_NO_YSTAGE") ? 0 : 1;
    return 0;
}

// reciprocals of the dictionary's divisors and of DEPRSTG / dt, computed on the device
int rhs_class_rcp(pihm_b200_ctx *ctx)
{
    double *d_c = nullptr, h_c[2] = {0.0, 0.0};
    bool good = cudaMalloc((void **)&d_c, sizeof(h_c)) == cudaSuccess;
    if (good) {
        k_class_rcp<<<(ctx->nclass + 127) / 128, 128>>>(ctx->d_cls, ctx->nclass, ctx->dm.dt, d_c);
        good = cudaMemcpy(h_c, d_c, sizeof(h_c), cudaMemcpyDeviceToHost) == cudaSuccess;
        cudaFree(d_c);
    }
    if (!good) return -1;
    ctx->dm.r_deprstg = h_c[0];
    ctx->dm.r_dt = h_c[1];
    return 0;
}

// RAREA / RDIST* slots of the tiles and the cold distance table (k_tile_rcp)
int rhs_tile_rcp(pihm_b200_ctx *ctx)
{
    const int nes = ctx->dm.nes;
    k_tile_rcp<<<(nes + 255) / 256, 256>>>(ctx->d_es, ctx->d_dist_cold, nes);
    return (cudaDeviceSynchronize() == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : -1;
}

// the packed send records of a state vector (NCCL transport / single-process emulation)
int rhs_halo_pack(pihm_b200_ctx *ctx, const double *y)
{
    const int n = ctx->nse + ctx->nsr;
    if (n > 0) {
        k_halo_pack<<<(n + 255) / 256, 256, 0, ctx->s()>>>(ctx->dm, y, ctx->nse, ctx->d_send_e_idx, ctx->d_send_e,
                                                           ctx->nsr, ctx->d_send_r_idx, ctx->d_send_r);
        ctx->launches++;
    }
    return 0;
}

// timing experiment (build with RHS_EXTRA=-DPB_HALO_TIMING): read / reset the %globaltimer stamps
int rhs_halo_times(unsigned long long *out, int reset)
{
#ifdef PB_HALO_TIMING
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (out && cudaMemcpyFromSymbol(out, g_ht, sizeof(unsigned long long) * 256 * 8) != cudaSuccess) return -1;
    if (reset) {
        std::vector<unsigned long long> h(256 * 8);
        for (int i = 0; i < 256; i++)
            for (int k = 0; k < 8; k++) h[i * 8 + k] = (k == 0 || k == 2 || k == 6) ? ~0ULL : 0ULL;
        if (cudaMemcpyToSymbol(g_ht, h.data(), sizeof(unsigned long long) * 256 * 8) != cudaSuccess) return -1;
    }
    return 0;
#else
    (void)out; (void)reset;
    return -2;
#endif
}

// ---------------------------------------------------------------------------
// RHS
// ---------------------------------------------------------------------------
// replay: evaluate the LAST call once more (same input, same ghosts, same stale river-edge
// flows) with the flux columns switched on -- no halo exchange, hidden state left as it is.
int rhs_launch(pihm_b200_ctx *ctx, const double *y, double *dy, bool replay)
{
    DevMesh dm = ctx->dm;
    HaloWait hw{};
    HaloPut hput{};
    if (replay) {
        if (ctx->nranks > 1 && ctx->halo_p2p) {     // the ghost records of the last exchange
            const int par = (int)(ctx->halo_seq & 1);
            dm.gel = ctx->d_hx + par * ctx->hx_stride;
            dm.gri = dm.gel + (size_t)dm.gs * (dm.ne - dm.nown);
        }
        dm.record = 1;
        dm.replay = 1;
        dm.xflux = ctx->d_xflux;
        dm.self = ctx->d_dm_rec;
        PB_CUDA(cudaMemcpyAsync(ctx->d_dm_rec, &dm, sizeof(DevMesh), cudaMemcpyHostToDevice, ctx->s()));
    } else if (ctx->nranks > 1 && ctx->halo_p2p) {
        // halo exchange over peer memory, inside k_pre: stores into the neighbours' ghost buffers +
        // arrival flags at the start of the kernel, the wait before its first non-interior tile
        const long long seq = ++ctx->halo_seq;
        const int par = (int)(seq & 1);
        hput.nse = ctx->nse; hput.nsr = ctx->nsr;
        hput.send_e = ctx->d_send_e_idx; hput.send_r = ctx->d_send_r_idx;
        hput.hp = ctx->hpeers;
        hput.par = par;
        hput.seq = (double)seq;
        hput.counter = ctx->d_hcount;
        hput.nput = std::max(1, (ctx->nse + ctx->nsr + PB_PRE_THREADS - 1) / PB_PRE_THREADS);     // one record per thread
        dm.gel = ctx->d_hx + par * ctx->hx_stride;
        dm.gri = dm.gel + (size_t)dm.gs * (dm.ne - dm.nown);
        dm.self = ctx->d_dm_par[par];
        hw.flags = ctx->d_hx + 2 * ctx->hx_stride + par * PB_MAX_RANKS_H;
        hw.nn = ctx->hpeers.nn;
        for (int k = 0; k < hw.nn; k++) hw.rank[k] = ctx->nbr_rank[k];
        hw.seq = (double)seq;
        // timing experiments only (wrong results): bit 0 drops the flag wait, bit 1 the put
        static const int dbg = std::getenv("PIHM_B200_HALO_DEBUG") ? std::atoi(std::getenv("PIHM_B200_HALO_DEBUG")) : 0;
        if (dbg & 1) hw.nn = 0;
        if (dbg & 2) hput.hp.nn = 0;
    } else if (ctx->nranks > 1) {
        // one-ring(+) halo exchange of neighbour and river states before the RHS
        const int n = ctx->nse + ctx->nsr;
        if (n > 0) {
            k_halo_pack<<<(n + 255) / 256, 256, 0, ctx->s()>>>(dm, y, ctx->nse, ctx->d_send_e_idx, ctx->d_send_e,
                                                               ctx->nsr, ctx->d_send_r_idx, ctx->d_send_r);
            ctx->launches++;
        }
        if (pb::comm_halo_exchange(ctx) != 0) return -1;
    }
    // work items = 32-entity tiles; k_pre: owned + ghost, k_main: owned only
    const int te = (dm.ne + PB_TILE - 1) / PB_TILE, tr = (dm.nr + PB_TILE - 1) / PB_TILE;
    const int te_own = (dm.nown + PB_TILE - 1) / PB_TILE, tr_own = (dm.rown + PB_TILE - 1) / PB_TILE;
    const auto groups = [](int t) { return (t + PB_RING_GROUP - 1) / PB_RING_GROUP; };
    const int gpre = std::max(1, std::min(ctx->pre_grid, groups(te) + groups(tr)));
    if (hput.nput > gpre) hput.nput = gpre;
    const int gmain = std::max(1, std::min(ctx->main_grid, groups(te_own) + groups(tr_own)));
    // own-state columns of a tile by bulk copy: needs 16-byte alignment of each block inside y
    int ys = 0;
    if (ctx->ystage && ((uintptr_t)y & 15) == 0) ys = 1 | ((dm.nown % 2 == 0) ? 2 : 0);
    const bool gh = (dm.nown != dm.ne) || (dm.rown != dm.nr);
    {
        // k_pre too is launched with programmatic stream serialization: its prologue (ring set-up, the
        // first static-slab requests) depends on nothing the predecessor writes and overlaps its tail;
        // everything else comes after griddepcontrol.wait
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gpre);
        cfg.blockDim = dim3(PB_PRE_THREADS);
        cfg.dynamicSmemBytes = ctx->pre_smem;
        cfg.stream = ctx->s();
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = (ctx->pdl && !std::getenv("PIHM_B200_PRE_NO_PDL")) ? 1 : 0;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const int ys1 = ys & 1;
        const cudaError_t e = gh ? cudaLaunchKernelEx(&cfg, k_pre<true>, dm, y, te, tr, ctx->ntile_int, hw, hput, ys1)
                                 : cudaLaunchKernelEx(&cfg, k_pre<false>, dm, y, te, tr, te, hw, hput, ys1);
        if (e != cudaSuccess) { set_error(std::string("k_pre launch: ") + cudaGetErrorString(e)); return -1; }
    }
    {
        // k_main right behind k_pre with programmatic stream serialization (PDL): its launch
        // latency and prologue overlap k_pre's tail
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gmain);
        cfg.blockDim = dim3(dm.fbr ? MainCfg<true>::THREADS : MainCfg<false>::THREADS);
        cfg.dynamicSmemBytes = ctx->main_smem;
        cfg.stream = ctx->s();
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = ctx->pdl ? 1 : 0;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const cudaError_t e =
            dm.fbr ? (gh ? cudaLaunchKernelEx(&cfg, k_main<true, true>, dm, y, dy, te_own, tr_own, ys)
                         : cudaLaunchKernelEx(&cfg, k_main<true, false>, dm, y, dy, te_own, tr_own, ys))
                   : (gh ? cudaLaunchKernelEx(&cfg, k_main<false, true>, dm, y, dy, te_own, tr_own, ys)
                         : cudaLaunchKernelEx(&cfg, k_main<false, false>, dm, y, dy, te_own, tr_own, ys));
        if (e != cudaSuccess) { set_error(std::string("k_main launch: ") + cudaGetErrorString(e)); return -1; }
    }
    ctx->launches += 2;
    if (!replay) ctx->last_in = y;
    ctx->flux_fresh = dm.record;
    return 0;
}


}  // namespace pb

using namespace pb;

extern "C" {

// test hook (not part of the reference interface): pow_pos vs libdevice pow on n pairs
int pihm_b200_test_pow(int n, const double *x, const double *y, double *fast, double *ref)
{
    double *d = nullptr;
    PB_CUDA(cudaMalloc((void **)&d, sizeof(double) * 4 * (size_t)n));
    PB_CUDA(cudaMemcpy(d, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(d + n, y, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_test_pow<<<(n + 255) / 256, 256>>>(n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n);
    PB_CUDA(cudaMemcpy(fast, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    PB_CUDA(cudaMemcpy(ref, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

// test hook: the branch-free division of the RHS kernels vs the hardware division
int pihm_b200_test_div(int n, const double *a, const double *b, double *fast, double *ok, double *ref)
{
    double *d = nullptr;
    PB_CUDA(cudaMalloc((void **)&d, sizeof(double) * 5 * (size_t)n));
    PB_CUDA(cudaMemcpy(d, a, sizeof(double) * n, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(d + n, b, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_test_div<<<(n + 255) / 256, 256>>>(n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n);
    PB_CUDA(cudaMemcpy(fast, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    PB_CUDA(cudaMemcpy(ok, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    PB_CUDA(cudaMemcpy(ref, d + 4 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}


}  // extern "C"
