// comm.cu -- NCCL plumbing of the partitioned run: one process per GPU, the
// per-RHS halo exchange (grouped ncclSend/ncclRecv between neighbour ranks) and
// the one-double all-reduce behind every dot product / norm.
// NCCL is dlopen'ed (libnccl.so.2 -- the copy torch already loaded when the
// process uses torch.distributed), so the library itself has no link-time
// dependency and loads on machines without NCCL.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#include "common.cuh"

namespace {

typedef void *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclFloat64 = 8 };                 // ncclDataType_t
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nccl &nccl()
{
    static Nccl n;
    if (n.h) return n;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) { n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.h) break; }
    if (!n.h) return n;
#define PB_SYM(f) *(void **)(&n.f) = dlsym(n.h, "nccl" #f)
    PB_SYM(GetUniqueId); PB_SYM(CommInitRank); PB_SYM(CommDestroy); PB_SYM(AllReduce); PB_SYM(AllGather);
    PB_SYM(Send); PB_SYM(Recv); PB_SYM(GroupStart); PB_SYM(GroupEnd); PB_SYM(GetErrorString);
#undef PB_SYM
    n.ok = n.GetUniqueId && n.CommInitRank && n.AllReduce && n.Send && n.Recv && n.GroupStart && n.GroupEnd;
    return n;
}

int check(int rc, const char *what)
{
    if (rc == 0) return 0;
    Nccl &n = nccl();
    pb::set_error(std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
    return -1;
}

}  // namespace

namespace pb {

int comm_halo_exchange(pihm_b200_ctx *ctx)
{
    Nccl &n = nccl();
    if (!ctx->comm) { set_error("halo exchange without a communicator"); return -1; }
    const int gs = ctx->dm.gs;
    size_t roff_e = 0, roff_r = 0;
    if (check(n.GroupStart(), "ncclGroupStart")) return -1;
    for (size_t k = 0; k < ctx->nbr_rank.size(); k++) {
        const int peer = ctx->nbr_rank[k];
        const size_t se = (size_t)(ctx->send_e_ptr[k + 1] - ctx->send_e_ptr[k]);
        const size_t sr = (size_t)(ctx->send_r_ptr[k + 1] - ctx->send_r_ptr[k]);
        if (se) n.Send(ctx->d_send_e + (size_t)ctx->send_e_ptr[k] * gs, se * gs, ncclFloat64, peer, ctx->comm, ctx->s());
        if (sr) n.Send(ctx->d_send_r + (size_t)ctx->send_r_ptr[k] * 2, sr * 2, ncclFloat64, peer, ctx->comm, ctx->s());
        if (ctx->recv_e_cnt[k]) n.Recv(ctx->d_gel + roff_e * gs, (size_t)ctx->recv_e_cnt[k] * gs, ncclFloat64, peer, ctx->comm, ctx->s());
        if (ctx->recv_r_cnt[k]) n.Recv(ctx->d_gri + roff_r * 2, (size_t)ctx->recv_r_cnt[k] * 2, ncclFloat64, peer, ctx->comm, ctx->s());
        roff_e += ctx->recv_e_cnt[k];
        roff_r += ctx->recv_r_cnt[k];
    }
    return check(n.GroupEnd(), "ncclGroupEnd (halo)");
}

int comm_allreduce(pihm_b200_ctx *ctx, double *dev_ptr, int count, int op)
{
    if (ctx->nranks <= 1) return 0;
    if (!ctx->comm) { set_error("all-reduce without a communicator (same-process rank group: no NCCL)"); return -1; }
    Nccl &n = nccl();
    const int nop = (op == 0) ? ncclSum : (op == 1 ? ncclMin : ncclMax);
    return check(n.AllReduce(dev_ptr, dev_ptr, (size_t)count, ncclFloat64, nop, ctx->comm, ctx->s()), "ncclAllReduce");
}

// Map one small device buffer of every rank into every other rank (CUDA IPC over NVLink peer
// access): peers[r] = this process's pointer to rank r's buffer, peers[rank] = local.  The
// integrator's reduction kernels use it to all-reduce their scalars themselves (cvode_kernels.cuh).
// Collective; returns 0 only if it worked on ALL ranks (the handles travel by ncclAllGather).
int comm_share_buffer(pihm_b200_ctx *ctx, void *local, void **peers)
{
    Nccl &n = nccl();
    if (ctx->nranks <= 1 || !ctx->comm || !n.AllGather) return -1;
    const int R = ctx->nranks;
    // A local failure must not leave the collective sequence: every rank goes through the
    // all-gather and the agreement all-reduce, and a rank that failed votes "no".  The staging
    // area is the context's reduction scratch (red_blocks + 64 doubles >= (R + 1) handles).
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    int good = (local && cudaIpcGetMemHandle(&mine, local) == cudaSuccess) ? 1 : 0;
    static_assert(sizeof(cudaIpcMemHandle_t) * (PB_MAX_RANKS_H + 1) <= sizeof(double) * 64 + 8 * 64, "staging area");
    unsigned char *d_h = reinterpret_cast<unsigned char *>(ctx->d_red);
    std::vector<cudaIpcMemHandle_t> all(R);
    cudaMemcpyAsync(d_h + sizeof(mine) * R, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->s());
    int rc = n.AllGather(d_h + sizeof(mine) * R, d_h, sizeof(mine), /*ncclInt8*/ 0, ctx->comm, ctx->s());
    cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * R, cudaMemcpyDeviceToHost, ctx->s());
    if (cudaStreamSynchronize(ctx->s()) != cudaSuccess || rc != 0) good = 0;
    for (int r = 0; r < R; r++) peers[r] = nullptr;
    for (int r = 0; r < R && good; r++) {
        if (r == ctx->rank) { peers[r] = local; continue; }
        if (cudaIpcOpenMemHandle(&peers[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { peers[r] = nullptr; good = 0; }
    }
    cudaGetLastError();
    // agree: every rank must have every mapping, otherwise nobody uses them
    double v = (double)good, *d = ctx->d_red;
    cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->s());
    n.AllReduce(d, d, 1, ncclFloat64, ncclMin, ctx->comm, ctx->s());
    cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->s());
    cudaStreamSynchronize(ctx->s());
    if (v > 0.5) return 0;
    comm_unshare_buffer(ctx, peers);        // somebody failed: close what this rank had opened
    return -1;
}

void comm_unshare_buffer(pihm_b200_ctx *ctx, void **peers)
{
    for (int r = 0; r < ctx->nranks; r++)
        if (r != ctx->rank && peers[r]) { cudaIpcCloseMemHandle(peers[r]); peers[r] = nullptr; }
}

// Peer-memory halo exchange: every rank allocates [2 parities][ghost element records | ghost
// river records] + arrival flags, maps it into all ranks, and learns from each neighbour where
// its own records go in the neighbour's buffer (one grouped ncclSend/ncclRecv of 3 doubles).
// Collective over all ranks; a rank with a local failure stays in the sequence and votes "no",
// and then nobody uses the mappings.
static long long halo_stride(const pihm_b200_ctx *ctx)
{
    const long long ng = ctx->dm.ne - ctx->dm.nown, nrg = ctx->dm.nr - ctx->dm.rown;
    return ((ctx->dm.gs * ng + 2 * nrg + 3) / 4) * 4 + 4;
}
// local allocations: the buffer itself, the CTA counter of the put, the two parity views of the mesh
static int halo_alloc(pihm_b200_ctx *ctx)
{
    const long long stride = halo_stride(ctx);
    const size_t bytes = sizeof(double) * (size_t)(2 * stride + 2 * PB_MAX_RANKS_H);
    int good = 1;
    if (cudaMalloc((void **)&ctx->d_hx, bytes) != cudaSuccess) { ctx->d_hx = nullptr; good = 0; }
    else cudaMemset(ctx->d_hx, 0, bytes);
    if (!ctx->d_hcount) {
        if (cudaMalloc((void **)&ctx->d_hcount, sizeof(unsigned int)) == cudaSuccess) cudaMemset(ctx->d_hcount, 0, sizeof(unsigned int));
        else { ctx->d_hcount = nullptr; good = 0; }
    }
    for (int p = 0; p < 2; p++)
        if (!ctx->d_dm_par[p] && cudaMalloc((void **)&ctx->d_dm_par[p], sizeof(DevMesh)) != cudaSuccess) { ctx->d_dm_par[p] = nullptr; good = 0; }
    cudaGetLastError();
    return good;
}
// where neighbour k's records land in THIS rank's buffer: {stride, element offset, river offset}
static void halo_my_layout(const pihm_b200_ctx *ctx, std::vector<double> &mine)
{
    const int nn = (int)ctx->nbr_rank.size(), gs = ctx->dm.gs;
    const long long ng = ctx->dm.ne - ctx->dm.nown, stride = halo_stride(ctx);
    mine.assign(3 * (size_t)std::max(nn, 1), 0.0);
    long long roff_e = 0, roff_r = 0;
    for (int k = 0; k < nn; k++) {
        mine[3 * k] = (double)stride;
        mine[3 * k + 1] = (double)(roff_e * gs);
        mine[3 * k + 2] = (double)(gs * ng + roff_r * 2);
        roff_e += ctx->recv_e_cnt[k];
        roff_r += ctx->recv_r_cnt[k];
    }
}
// everything agreed: theirs[3k..] = where this rank's records land in neighbour k's buffer
static void halo_finish(pihm_b200_ctx *ctx, const std::vector<double> &theirs)
{
    const int nn = (int)ctx->nbr_rank.size(), gs = ctx->dm.gs;
    const long long ng = ctx->dm.ne - ctx->dm.nown, stride = halo_stride(ctx);
    HaloPeers &hp = ctx->hpeers;
    hp.nn = nn;
    hp.myrank = ctx->rank;
    for (int k = 0; k < nn; k++) {
        hp.base[k] = static_cast<double *>(ctx->hx_peer[ctx->nbr_rank[k]]);
        hp.pstride[k] = (long long)theirs[3 * k];
        hp.gel_off[k] = (long long)theirs[3 * k + 1];
        hp.gri_off[k] = (long long)theirs[3 * k + 2];
        hp.flag_off[k] = 2 * hp.pstride[k];
    }
    for (int k = 0; k <= nn; k++) { hp.e_ptr[k] = ctx->send_e_ptr[k]; hp.r_ptr[k] = ctx->send_r_ptr[k]; }
    ctx->hx_stride = stride;
    // device copies of the mesh view whose ghost pointers select a parity (the rare exact paths read them)
    for (int p = 0; p < 2; p++) {
        DevMesh dm = ctx->dm;
        dm.gel = ctx->d_hx + p * stride;
        dm.gri = ctx->d_hx + p * stride + gs * ng;
        dm.self = ctx->d_dm_par[p];
        cudaMemcpy(ctx->d_dm_par[p], &dm, sizeof(DevMesh), cudaMemcpyHostToDevice);
    }
    ctx->halo_p2p = 1;
}
static void halo_release(pihm_b200_ctx *ctx)
{
    if (ctx->d_hx) { cudaFree(ctx->d_hx); ctx->d_hx = nullptr; }
    for (int p = 0; p < 2; p++) if (ctx->d_dm_par[p]) { cudaFree(ctx->d_dm_par[p]); ctx->d_dm_par[p] = nullptr; }
}

int comm_setup_halo_p2p(pihm_b200_ctx *ctx)
{
    Nccl &n = nccl();
    if (ctx->nranks <= 1 || ctx->nranks > PB_MAX_RANKS_H || !ctx->comm) return -1;
    const int nn = (int)ctx->nbr_rank.size();
    int good = (nn <= PB_MAX_NBR) ? 1 : 0;
    if (!halo_alloc(ctx)) good = 0;
    const int shared = (comm_share_buffer(ctx, ctx->d_hx, ctx->hx_peer) == 0);
    if (!shared) good = 0;
    // tell every neighbour where its records land here; staging in the reduction scratch (no
    // allocation that could fail between the collectives)
    std::vector<double> mine, theirs(3 * (size_t)std::max(nn, 1), 0.0);
    halo_my_layout(ctx, mine);
    double *d_x = ctx->d_red;
    const bool fits = 6 * (size_t)std::max(nn, 1) <= 64 + (size_t)ctx->red_blocks;
    if (!fits) good = 0;
    if (fits) cudaMemcpyAsync(d_x, mine.data(), sizeof(double) * 3 * nn, cudaMemcpyHostToDevice, ctx->s());
    n.GroupStart();
    for (int k = 0; k < nn && fits; k++) {
        n.Send(d_x + 3 * k, 3, ncclFloat64, ctx->nbr_rank[k], ctx->comm, ctx->s());
        n.Recv(d_x + 3 * nn + 3 * k, 3, ncclFloat64, ctx->nbr_rank[k], ctx->comm, ctx->s());
    }
    if (n.GroupEnd() != 0) good = 0;
    if (fits) cudaMemcpyAsync(theirs.data(), d_x + 3 * nn, sizeof(double) * 3 * nn, cudaMemcpyDeviceToHost, ctx->s());
    if (cudaStreamSynchronize(ctx->s()) != cudaSuccess) good = 0;
    // agree
    double v = (double)good;
    cudaMemcpyAsync(d_x, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->s());
    n.AllReduce(d_x, d_x, 1, ncclFloat64, ncclMin, ctx->comm, ctx->s());
    cudaMemcpyAsync(&v, d_x, sizeof(double), cudaMemcpyDeviceToHost, ctx->s());
    cudaStreamSynchronize(ctx->s());
    cudaGetLastError();
    if (v < 0.5) {      // somebody failed: nobody uses peer memory; undo the local part
        if (shared) comm_unshare_buffer(ctx, ctx->hx_peer);
        halo_release(ctx);
        return -1;
    }
    halo_finish(ctx, theirs);
    return 0;
}

// The same wiring for ranks that live in ONE process (one context per rank, on one or several
// devices): buffers are exchanged as plain pointers, no NCCL, no IPC.  Used by the single-GPU
// tests of the peer-memory halo exchange / in-kernel all-reduce, and by a threaded driver.
int comm_setup_halo_local(pihm_b200_ctx **ctxs, int n)
{
    for (int r = 0; r < n; r++) {
        cudaSetDevice(ctxs[r]->device);
        if ((int)ctxs[r]->nbr_rank.size() > PB_MAX_NBR || !halo_alloc(ctxs[r])) {
            for (int q = 0; q <= r; q++) halo_release(ctxs[q]);
            set_error("comm_init_local: halo buffers");
            return -1;
        }
    }
    for (int r = 0; r < n; r++) {
        pihm_b200_ctx *ctx = ctxs[r];
        for (int q = 0; q < n; q++) ctx->hx_peer[q] = ctxs[q]->d_hx;
        const int nn = (int)ctx->nbr_rank.size();
        std::vector<double> theirs(3 * (size_t)std::max(nn, 1), 0.0), lay;
        for (int k = 0; k < nn; k++) {
            const pihm_b200_ctx *o = ctxs[ctx->nbr_rank[k]];
            halo_my_layout(o, lay);
            int ko = -1;
            for (size_t j = 0; j < o->nbr_rank.size(); j++) if (o->nbr_rank[j] == r) ko = (int)j;
            if (ko < 0) { set_error("comm_init_local: neighbour relation is not symmetric"); return -1; }
            for (int c = 0; c < 3; c++) theirs[3 * k + c] = lay[3 * ko + c];
        }
        cudaSetDevice(ctx->device);
        halo_finish(ctx, theirs);
    }
    return 0;
}

void comm_destroy(pihm_b200_ctx *ctx)
{
    if (ctx->halo_p2p) { if (!ctx->lgroup) comm_unshare_buffer(ctx, ctx->hx_peer); ctx->halo_p2p = 0; }
    if (ctx->d_hx) { cudaFree(ctx->d_hx); ctx->d_hx = nullptr; }
    if (ctx->d_hcount) { cudaFree(ctx->d_hcount); ctx->d_hcount = nullptr; }
    for (int p = 0; p < 2; p++) if (ctx->d_dm_par[p]) { cudaFree(ctx->d_dm_par[p]); ctx->d_dm_par[p] = nullptr; }
    if (ctx->comm) {
        Nccl &n = nccl();
        if (n.CommDestroy) n.CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
}

}  // namespace pb

extern "C" {

// rank 0 creates the id; the launcher (torch.distributed broadcast) ships it to the others
int pihm_b200_comm_unique_id(void *out128)
{
    Nccl &n = nccl();
    if (!n.ok) { pb::set_error("libnccl.so.2 not found"); return -1; }
    ncclUniqueId id;
    if (check(n.GetUniqueId(&id), "ncclGetUniqueId")) return -1;
    std::memcpy(out128, &id, sizeof(id));
    return 0;
}

int pihm_b200_comm_init(pihm_b200_ctx *ctx, int rank, int nranks, const void *id128)
{
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) { pb::set_error("comm_init: bad argument"); return -1; }
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (nranks == 1) return 0;
    Nccl &n = nccl();
    if (!n.ok) { pb::set_error("libnccl.so.2 not found"); return -1; }
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    cudaSetDevice(ctx->device);
    if (check(n.CommInitRank(&ctx->comm, nranks, id, rank), "ncclCommInitRank")) return -1;
    // global number of unknowns (N of the WRMS norms, sqrt(N) of the SPGMR tolerance)
    double *d = ctx->d_red;
    double v = (double)ctx->nsv;
    PB_CUDA(cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->s()));
    if (pb::comm_allreduce(ctx, d, 1, 0)) return -1;
    PB_CUDA(cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->s()));
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    ctx->nsv_global = (long long)(v + 0.5);
    // neighbours' ghost buffers over NVLink peer memory.  PIHM_B200_NO_P2P=1 asks for the NCCL
    // send/recv path instead; without it a failed mapping is an error (all ranks see it: the
    // set-up is collective), unless PIHM_B200_P2P_OPTIONAL=1 allows the silent fallback.
    const bool no_p2p = std::getenv("PIHM_B200_NO_P2P") && std::atoi(std::getenv("PIHM_B200_NO_P2P"));
    if (!no_p2p && pb::comm_setup_halo_p2p(ctx) != 0) {
        const bool optional = std::getenv("PIHM_B200_P2P_OPTIONAL") && std::atoi(std::getenv("PIHM_B200_P2P_OPTIONAL"));
        if (!optional) {
            pb::set_error("comm_init: the peer-memory halo exchange could not be set up (CUDA IPC / peer access); "
                          "set PIHM_B200_NO_P2P=1 to run over NCCL send/recv");
            return -1;
        }
    }
    return 0;
}

// Ranks of ONE process (ctxs[r] = rank r; any mix of devices with peer access): peer-memory halo
// exchange and in-kernel all-reduce wired with plain pointers, no NCCL.  Create the integrators
// (pihm_b200_cvode_create) of ALL ranks before the first solve: each one registers its exchange
// buffer with the group.
int pihm_b200_comm_init_local(pihm_b200_ctx **ctxs, int nranks)
{
    if (!ctxs || nranks < 1 || nranks > PB_MAX_RANKS_H) { pb::set_error("comm_init_local: bad argument"); return -1; }
    auto g = std::make_shared<pb::LocalGroup>();
    g->n = nranks;
    long long nsv = 0;
    for (int r = 0; r < nranks; r++) {
        if (!ctxs[r] || ctxs[r]->comm || ctxs[r]->lgroup) { pb::set_error("comm_init_local: context already in a group"); return -1; }
        nsv += ctxs[r]->nsv;
        g->ctx[r] = ctxs[r];
    }
    for (int r = 0; r < nranks; r++)
        for (int q = 0; q < nranks; q++)
            if (ctxs[r]->device != ctxs[q]->device) {
                cudaSetDevice(ctxs[r]->device);
                cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                cudaGetLastError();         // already enabled is fine
            }
    for (int r = 0; r < nranks; r++) {
        ctxs[r]->rank = r;
        ctxs[r]->nranks = nranks;
        ctxs[r]->nsv_global = nsv;
        ctxs[r]->lgroup = g;
        // Ranks that SHARE a device wait for each other inside kernels (halo flags in k_pre, tickets
        // in the reduction kernels), so a waiting kernel must leave room for its peers' kernels: the
        // k_pre grids of ALL ranks together leave 16 SMs without any of their CTAs (an SM that hosts a
        // waiting k_pre CTA keeps that kernel's shared-memory configuration: a peer's kernel that wants
        // another one -- the state permutation of a download, a vector kernel -- cannot join it there,
        // and with every SM taken the peer never gets to the launch its neighbour waits for: the tests
        // hung about one time in four), and no kernel is launched ahead of its predecessor's end (no
        // PDL: a dependent grid parked on the SMs could keep a peer's kernel off them).  Host calls
        // that synchronise the device (cudaMalloc / cudaFree) must not be issued while a peer's RHS is
        // in flight: allocate before the first evaluation.
        int share = 0, sms = 1;
        for (int q = 0; q < nranks; q++) share += (ctxs[q]->device == ctxs[r]->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctxs[r]->device);
        if (share > 1) {
            ctxs[r]->pdl = 0;
            ctxs[r]->pre_grid = std::max(1, std::min(ctxs[r]->pre_grid, std::max(1, (sms - 16) / share)));
            ctxs[r]->main_grid = std::max(1, ctxs[r]->main_grid / share);
        }
    }
    if (nranks == 1) return 0;
    return pb::comm_setup_halo_local(ctxs, nranks);
}

// which transport the partitioned run uses: out[0] halo (1 peer memory inside k_pre, 0 NCCL send/recv)
int pihm_b200_comm_paths(const pihm_b200_ctx *ctx, int32_t *out)
{
    if (!ctx || !out) return -1;
    out[0] = ctx->halo_p2p;
    out[1] = ctx->lgroup ? 1 : 0;
    return 0;
}

int64_t pihm_b200_num_state_var_global(const pihm_b200_ctx *ctx) { return ctx ? ctx->nsv_global : 0; }

}  // extern "C"
