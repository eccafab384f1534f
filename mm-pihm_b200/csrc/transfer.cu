// transfer.cu -- pipelined host transfers around the model step (C ABI: pihm_b200_forcing_prefetch /
// pihm_b200_forcing_commit / pihm_b200_vec_download_async / pihm_b200_transfer_wait / pihm_b200_transfer_release).
//
// The synchronous calls (pihm_b200_set_forcing_col, pihm_b200_vec_download) put their PCIe copies on the
// compute stream: 24 MB in and 24 MB out per model step at 1M triangles stand between two steps.  Here the
// copies run on a copy stream of the context's own, ordered against the compute stream by events:
//   * pull:  the state is brought into reference order by k_permute_state ON the compute stream (it reads the
//     vector), the device-to-host copy of that staging buffer overlaps whatever is launched next;
//   * push:  the forcing columns of the NEXT step travel to a staging table while this step computes;
//     pihm_b200_forcing_commit scatters them into the tile layout in stream order, i.e. after the last RHS
//     evaluation (and the Summary replay) that reads the previous values.
// Host buffers must be pinned (cudaHostAlloc / torch pin_memory): a pageable buffer makes the copy synchronous.
// The state of a context's pipeline lives here, keyed by the context (the context structure of common.cuh stays
// as it is); pihm_b200_destroy releases it.
#include <mutex>
#include <string>
#include <unordered_map>
#include "common.cuh"
#include "nvec.cuh"

using namespace pb;

namespace {

struct Pipe {
    cudaStream_t copy = nullptr;
    cudaEvent_t ready = nullptr;       // compute stream: the permuted state sits in d_pull
    cudaEvent_t pulled = nullptr;      // copy stream: the host buffer of the last pull is complete
    cudaEvent_t fetched = nullptr;     // copy stream: the prefetched forcing columns are on the device
    cudaEvent_t scattered = nullptr;   // compute stream: the last commit has read d_pref
    double *d_pull = nullptr;          // [nsv], reference order
    double *d_pref = nullptr;          // [PB_F_NCOL][ne], reference order
    int cols[PB_F_NCOL] = {};
    int ncol = 0;
    bool pull_pending = false, pref_pending = false, scatter_recorded = false;
};

std::mutex g_mu;
std::unordered_map<pihm_b200_ctx *, Pipe> g_pipes;

void destroy(Pipe &p)
{
    if (p.copy) cudaStreamSynchronize(p.copy);
    for (cudaEvent_t e : {p.ready, p.pulled, p.fetched, p.scattered}) if (e) cudaEventDestroy(e);
    if (p.d_pull) cudaFree(p.d_pull);
    if (p.d_pref) cudaFree(p.d_pref);
    if (p.copy) cudaStreamDestroy(p.copy);
    p = Pipe();
}

// the context's pipeline, created on first use; nullptr (error set) when a CUDA call fails
Pipe *pipe_of(pihm_b200_ctx *ctx)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_pipes.find(ctx);
    if (it != g_pipes.end()) return &it->second;
    Pipe p;
    cudaError_t e = cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking);
    for (cudaEvent_t *ev : {&p.ready, &p.pulled, &p.fetched, &p.scattered})
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p.d_pull, sizeof(double) * std::max<int64_t>(ctx->nsv, 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p.d_pref, sizeof(double) * PB_F_NCOL * std::max(ctx->dm.ne, 1));
    if (e != cudaSuccess) {
        set_error(std::string("transfer pipeline: ") + cudaGetErrorString(e));
        destroy(p);
        return nullptr;
    }
    return &(g_pipes[ctx] = p);
}

// ... without creating it (nullptr: the context has no pipeline)
Pipe *find_pipe(pihm_b200_ctx *ctx)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_pipes.find(ctx);
    return (it == g_pipes.end()) ? nullptr : &it->second;
}

int vec_grid(const pihm_b200_ctx *ctx, long long n)
{
    const long long b = (n + PB_VEC_THREADS - 1) / PB_VEC_THREADS;
    return (int)std::max<long long>(1, std::min<long long>(b, ctx->red_blocks));
}

}  // namespace

extern "C" {

int pihm_b200_vec_download_async(const pihm_b200_vec *v, double *host_pinned)
{
    if (!v || !host_pinned) { set_error("vec_download_async: bad argument"); return -1; }
    pihm_b200_ctx *ctx = v->ctx;
    Pipe *p = pipe_of(ctx);
    if (!p) return -1;
    // the staging buffer is free once the previous pull has left it
    if (p->pull_pending) PB_CUDA(cudaStreamWaitEvent(ctx->s(), p->pulled, 0));
    if (ctx->reorder) {
        k_permute_state<<<vec_grid(ctx, v->n), PB_VEC_THREADS, 0, ctx->s()>>>(
            ctx->dm.ne, ctx->dm.nr, ctx->dm.fbr, ctx->d_perm, v->d, p->d_pull, 0);
        ctx->launches++;
    } else {
        PB_CUDA(cudaMemcpyAsync(p->d_pull, v->d, sizeof(double) * v->n, cudaMemcpyDeviceToDevice, ctx->s()));
    }
    PB_CUDA(cudaEventRecord(p->ready, ctx->s()));
    PB_CUDA(cudaStreamWaitEvent(p->copy, p->ready, 0));
    PB_CUDA(cudaMemcpyAsync(host_pinned, p->d_pull, sizeof(double) * v->n, cudaMemcpyDeviceToHost, p->copy));
    PB_CUDA(cudaEventRecord(p->pulled, p->copy));
    p->pull_pending = true;
    return 0;
}

int pihm_b200_transfer_wait(pihm_b200_ctx *ctx)
{
    if (!ctx) { set_error("transfer_wait: bad argument"); return -1; }
    Pipe *p = find_pipe(ctx);
    if (p && p->pull_pending) PB_CUDA(cudaEventSynchronize(p->pulled));
    return 0;
}

int pihm_b200_forcing_prefetch(pihm_b200_ctx *ctx, int ncol, const int *cols, const double *const *values_pinned)
{
    if (!ctx || ncol < 1 || ncol > PB_F_NCOL || !cols || !values_pinned) { set_error("forcing_prefetch: bad argument"); return -1; }
    for (int k = 0; k < ncol; k++)
        if (cols[k] < 0 || cols[k] >= PB_F_NCOL || !values_pinned[k]) { set_error("forcing_prefetch: bad column"); return -1; }
    Pipe *p = pipe_of(ctx);
    if (!p) return -1;
    if (p->pref_pending) { set_error("forcing_prefetch: the previous prefetch has not been committed"); return -1; }
    // the staging table is free once the scatter kernels of the last commit have read it
    if (p->scatter_recorded) PB_CUDA(cudaStreamWaitEvent(p->copy, p->scattered, 0));
    const size_t ne = (size_t)ctx->dm.ne;
    for (int k = 0; k < ncol; k++) {
        p->cols[k] = cols[k];
        PB_CUDA(cudaMemcpyAsync(p->d_pref + (size_t)k * ne, values_pinned[k], sizeof(double) * ne, cudaMemcpyHostToDevice, p->copy));
    }
    PB_CUDA(cudaEventRecord(p->fetched, p->copy));
    p->ncol = ncol;
    p->pref_pending = true;
    return 0;
}

int pihm_b200_forcing_commit(pihm_b200_ctx *ctx)
{
    if (!ctx) { set_error("forcing_commit: bad argument"); return -1; }
    Pipe *p = find_pipe(ctx);
    if (!p || !p->pref_pending) return 0;
    const int ne = ctx->dm.ne;
    PB_CUDA(cudaStreamWaitEvent(ctx->s(), p->fetched, 0));
    for (int k = 0; k < p->ncol; k++) {
        k_scatter_forcing<<<(ne + 255) / 256, 256, 0, ctx->s()>>>(ne, ctx->dm.nes, p->cols[k], ctx->d_perm,
                                                                  p->d_pref + (size_t)k * ne, ctx->d_ft, ctx->d_forc);
        ctx->launches++;
    }
    PB_CUDA(cudaEventRecord(p->scattered, ctx->s()));
    p->scatter_recorded = true;
    p->pref_pending = false;
    PB_CUDA(cudaGetLastError());
    return 0;
}

// frees the copy stream and staging buffers of the context's pipeline (pihm_b200_destroy calls it; a context that
// never used the calls above has nothing to release)
int pihm_b200_transfer_release(pihm_b200_ctx *ctx)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_pipes.find(ctx);
    if (it == g_pipes.end()) return 0;
    if (ctx) cudaStreamSynchronize(ctx->s());
    destroy(it->second);
    g_pipes.erase(it);
    return 0;
}

}  // extern "C"
