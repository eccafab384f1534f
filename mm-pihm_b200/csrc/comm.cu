// comm.cu -- NCCL plumbing of the partitioned run: one process per GPU, the
// per-RHS halo exchange (grouped ncclSend/ncclRecv between neighbour ranks) and
// the one-double all-reduce behind every dot product / norm.
// NCCL is dlopen'ed (libnccl.so.2 -- the copy torch already loaded when the
// process uses torch.distributed), so the library itself has no link-time
// dependency and loads on machines without NCCL.
#include <dlfcn.h>
#include <cstring>
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
    cudaIpcMemHandle_t mine;
    int good = (cudaIpcGetMemHandle(&mine, local) == cudaSuccess) ? 1 : 0;
    unsigned char *d_h = nullptr;
    std::vector<cudaIpcMemHandle_t> all(R);
    if (cudaMalloc((void **)&d_h, sizeof(mine) * (R + 1)) != cudaSuccess) return -1;
    cudaMemcpyAsync(d_h + sizeof(mine) * R, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->s());
    int rc = n.AllGather(d_h + sizeof(mine) * R, d_h, sizeof(mine), /*ncclInt8*/ 0, ctx->comm, ctx->s());
    cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * R, cudaMemcpyDeviceToHost, ctx->s());
    if (cudaStreamSynchronize(ctx->s()) != cudaSuccess || rc != 0) good = 0;
    for (int r = 0; r < R && good; r++) {
        if (r == ctx->rank) { peers[r] = local; continue; }
        if (cudaIpcOpenMemHandle(&peers[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) good = 0;
    }
    cudaGetLastError();
    // agree: every rank must have every mapping, otherwise nobody uses them
    double v = (double)good, *d = reinterpret_cast<double *>(d_h);
    cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->s());
    n.AllReduce(d, d, 1, ncclFloat64, ncclMin, ctx->comm, ctx->s());
    cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->s());
    cudaStreamSynchronize(ctx->s());
    cudaFree(d_h);
    return (v > 0.5) ? 0 : -1;
}

void comm_unshare_buffer(pihm_b200_ctx *ctx, void **peers)
{
    for (int r = 0; r < ctx->nranks; r++)
        if (r != ctx->rank && peers[r]) { cudaIpcCloseMemHandle(peers[r]); peers[r] = nullptr; }
}

void comm_destroy(pihm_b200_ctx *ctx)
{
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
    return 0;
}

int64_t pihm_b200_num_state_var_global(const pihm_b200_ctx *ctx) { return ctx ? ctx->nsv_global : 0; }

}  // extern "C"
