// cvode_kernels.cuh -- fused vector kernels of the device-resident BDF/Newton/
// SPGMR integrator.
//
// Each kernel replaces a run of N_V* calls of CVODE 2.9.0 / SPGMR and performs,
// per component, exactly the floating-point operations of that run in the same
// order (the special-case forms of N_VLinearSum_Serial included; file:line
// cited per kernel).  Scalars produced by one kernel and consumed by the next
// (WRMS norms, Gram-Schmidt coefficients) stay in device memory (`sc[]`) so the
// host is not in the loop; the host reads them from a mapped pinned mirror only
// where CVODE's control flow branches on them.
//
// Reductions: per-thread grid-stride partials -> warp shuffles -> block partials
// in `part`; the last block to finish sums the partials in index order and
// stores the RAW sum (or min) to sc[slot] and its host mirror.  Consumers apply
// sqrt( /N) themselves, so a multi-GPU build can all-reduce the raw value in
// between.  Compiled with -fmad=false.
#pragma once
#include "common.cuh"
#include "nvec.cuh"

namespace pb {

// device / host scalar slots
enum {
    SC_EWT_MIN = 0,   // min(reltol*|y|+abstol)
    SC_EWT_NRM,       // sum (zn0*ewt)^2
    SC_BSUM,          // sum (b*ewt)^2
    SC_VNRM,          // sum (vtemp*ewt)^2  (DQ increment)
    SC_VK2,           // dot(V[k],V[k]) before Gram-Schmidt
    SC_H0,            // SC_H0+i = dot(V[i],V[k]), i = 0..5
    SC_NEW2 = SC_H0 + 6,   // dot(V[k],V[k]) after Gram-Schmidt
    SC_DEL,           // sum (b*ewt)^2 of the Newton correction
    SC_ACNRM,         // sum (acor*ewt)^2
    SC_ETA_M1,        // sum (zn[q]*ewt)^2
    SC_ETA_P1,        // sum ((acor - cquot zn[qmax])*ewt)^2
    SC_STAB1,         // sum (zn[q]*ewt)^2
    SC_STAB2,         // sum (zn[q-1]*ewt)^2
    SC_TMP,
    SC_SEQ,           // ticket of the last finished reduction kernel (host spin-wait)
    SC_GATE,          // device copy of that ticket for k_mgs_chain's CTAs (RedBuf::gate)
    SC_COUNT = 32
};

#define PB_MAX_RANKS 8
struct RedBuf {
    double *part;            // [SC_COUNT][max_blocks]
    unsigned int *counter;
    double *sc;              // device scalars
    volatile double *hsc;    // mapped host mirror: [SC_COUNT] pairs {value, ticket}
    int max_blocks;
    double seq;              // ticket of this reduction kernel (travels with each result)
    // partitioned run: the exchange buffer of every rank, mapped into this one (comm_share_buffer);
    // null = single GPU, or the scalars are all-reduced by NCCL after the kernel
    double *const *peer;
    int nranks, rank;
    // device-side gate: when set, the finishing thread also publishes the ticket here, for the
    // other CTAs of a cooperative kernel that runs several reductions in a row (k_mgs_chain)
    double *gate;
};
// exchange buffer: [parity of the ticket][slot][rank]{value, ticket}
#define PB_XB_DOUBLES (2 * SC_COUNT * PB_MAX_RANKS * 2)
__device__ __forceinline__ int xb_index(int par, int slot, int r) { return ((par * SC_COUNT + slot) * PB_MAX_RANKS + r) * 2; }

// All-reduce of the finishing thread's raw sums over the ranks, inside the reduction kernel: the thread stores
// {value, ticket} straight into every rank's exchange buffer over NVLink, waits for the pair of every rank in
// its own buffer and combines the values in rank order, so all ranks hold the same bits.  Consecutive tickets
// alternate between two copies of the buffer: a rank can be at most one reduction ahead of the slowest one,
// so a value is never overwritten before it has been read.
// {value, ticket} travel as ONE naturally aligned 16-byte store / load: the pair arrives (or not) as a whole, so
// it validates itself and the exchange needs no fence between a value and its ticket (PB_RED_PAIR = 0: the
// protocol of round 1 -- values, system-scope fence, tickets, wait, fence -- whose two fences cost ~5 us each)
#ifndef PB_RED_PAIR
#define PB_RED_PAIR 1
#endif
__device__ __forceinline__ void st_pair(double *p, double v, double t)
{
    asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(p)), "d"(v), "d"(t) : "memory");
}
__device__ __forceinline__ void ld_pair(const double *p, double &v, double &t)
{
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(t) : "l"(__cvta_generic_to_global(p)) : "memory");
}

template <bool MIN_A>
__device__ __forceinline__ void red_exchange(const RedBuf &rb, double &a, int slotA, double &b, int slotB)
{
    const int par = (int)((long long)rb.seq & 1);
#if PB_RED_PAIR
    for (int p = 0; p < rb.nranks; p++) {
        double *x = rb.peer[p];
        st_pair(x + xb_index(par, slotA, rb.rank), a, rb.seq);
        if (slotB >= 0) st_pair(x + xb_index(par, slotB, rb.rank), b, rb.seq);
    }
    const double *mine = rb.peer[rb.rank];
    bool lost = false;
    double sa = MIN_A ? __longlong_as_double(0x7ff0000000000000LL) : 0.0, sb = 0.0;
    for (int r = 0; r < rb.nranks; r++) {       // rank order: every rank adds the same values in the same order
        double va = 0.0, vb = 0.0, tk = 0.0;
        long long spins = 0;
        for (;;) {
            ld_pair(mine + xb_index(par, slotA, r), va, tk);
            if (tk == rb.seq) break;
            if (++spins > (1LL << 31)) { lost = true; break; }   // a lost rank must not hang the GPU forever
        }
        if (slotB >= 0 && !lost) {
            for (;;) {
                ld_pair(mine + xb_index(par, slotB, r), vb, tk);
                if (tk == rb.seq) break;
                if (++spins > (1LL << 31)) { lost = true; break; }
            }
        }
        sa = MIN_A ? ((va < sa) ? va : sa) : sa + va;
        sb += vb;
    }
#else
    for (int p = 0; p < rb.nranks; p++) {
        volatile double *x = rb.peer[p];
        x[xb_index(par, slotA, rb.rank)] = a;
        if (slotB >= 0) x[xb_index(par, slotB, rb.rank)] = b;
    }
    fence_sys();
    for (int p = 0; p < rb.nranks; p++) {
        volatile double *x = rb.peer[p];
        x[xb_index(par, slotA, rb.rank) + 1] = rb.seq;
    }
    const volatile double *mine = rb.peer[rb.rank];
    bool lost = false;
    for (int r = 0; r < rb.nranks; r++) {
        long long spins = 0;
        while (mine[xb_index(par, slotA, r) + 1] != rb.seq)
            if (++spins > (1LL << 31)) { lost = true; break; }   // a lost rank must not hang the GPU forever
    }
    fence_sys();
    double sa = MIN_A ? __longlong_as_double(0x7ff0000000000000LL) : 0.0, sb = 0.0;
    for (int r = 0; r < rb.nranks; r++) {
        const double va = mine[xb_index(par, slotA, r)];
        sa = MIN_A ? ((va < sa) ? va : sa) : sa + va;
        if (slotB >= 0) sb += mine[xb_index(par, slotB, r)];
    }
#endif
    // a rank that never arrived: poison the result so that the integrator stops (NaN norm)
    a = lost ? __longlong_as_double(0x7ff8000000000000LL) : sa;
    b = lost ? __longlong_as_double(0x7ff8000000000000LL) : sb;
}

template <bool IS_MIN>
__device__ __forceinline__ double red_block(double v)
{
    __shared__ double sh[PB_VEC_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_down_sync(0xffffffffu, v, o);
        v = IS_MIN ? ((w < v) ? w : v) : v + w;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();             // protect sh[] across successive calls
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = (lane < PB_VEC_THREADS / 32) ? sh[lane] : (IS_MIN ? __longlong_as_double(0x7ff0000000000000LL) : 0.0);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const double w = __shfl_down_sync(0xffffffffu, v, o);
            v = IS_MIN ? ((w < v) ? w : v) : v + w;
        }
    }
    return v;
}

// finish up to 2 reductions of one kernel.  slotA: sum (or min if MIN_A); slotB: sum, -1 = none
template <bool MIN_A>
__device__ __forceinline__ void red_finish(const RedBuf &rb, double a, int slotA, double b, int slotB)
{
    a = red_block<MIN_A>(a);
    if (slotB >= 0) b = red_block<false>(b);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        rb.part[(size_t)slotA * rb.max_blocks + blockIdx.x] = a;
        if (slotB >= 0) rb.part[(size_t)slotB * rb.max_blocks + blockIdx.x] = b;
        fence_gpu();
        last = (atomicAdd(rb.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    fence_gpu();
    double sa = MIN_A ? __longlong_as_double(0x7ff0000000000000LL) : 0.0, sb = 0.0;
    const volatile double *pa = rb.part + (size_t)slotA * rb.max_blocks;
    const volatile double *pb_ = rb.part + (size_t)(slotB >= 0 ? slotB : slotA) * rb.max_blocks;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += PB_VEC_THREADS) {
        const double va = pa[k];
        sa = MIN_A ? ((va < sa) ? va : sa) : sa + va;
        if (slotB >= 0) sb += pb_[k];
    }
    sa = red_block<MIN_A>(sa);
    if (slotB >= 0) sb = red_block<false>(sb);
    if (threadIdx.x == 0) {
        if (rb.peer) red_exchange<MIN_A>(rb, sa, slotA, sb, slotB);
        // to the host: {value, ticket} of each slot as ONE 16-byte store into the mapped page -- the pair
        // validates itself (the host polls the ticket word of every slot it is about to read), so no
        // system-scope fence sits between the results and their announcement, i.e. in the kernel's tail
        double *hp = const_cast<double *>(rb.hsc);
        rb.sc[slotA] = sa;
        st_pair(hp + 2 * slotA, sa, rb.seq);
        if (slotB >= 0) { rb.sc[slotB] = sb; st_pair(hp + 2 * slotB, sb, rb.seq); }
        *rb.counter = 0u;
        if (rb.gate) { fence_gpu(); *reinterpret_cast<volatile double *>(rb.gate) = rb.seq; }
    }
}

// the same for up to NV sums of one kernel (slot[k] < 0: not wanted; value 0 may be a minimum): per value the
// block tree, the partials, the last block's sum in index order -- each value goes through exactly the
// operations red_finish would apply to it, so a sum does not depend on which kernel produced it
template <int NV, bool MIN_0>
__device__ __forceinline__ void red_finish_n(const RedBuf &rb, double (&v)[NV], const int (&slot)[NV])
{
#pragma unroll
    for (int k = 0; k < NV; k++)
        if (slot[k] >= 0) v[k] = (MIN_0 && k == 0) ? red_block<true>(v[k]) : red_block<false>(v[k]);
    __shared__ bool last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++)
            if (slot[k] >= 0) rb.part[(size_t)slot[k] * rb.max_blocks + blockIdx.x] = v[k];
        fence_gpu();
        last = (atomicAdd(rb.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    fence_gpu();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        if (slot[k] < 0) continue;
        const bool is_min = MIN_0 && k == 0;
        double sa = is_min ? __longlong_as_double(0x7ff0000000000000LL) : 0.0;
        const volatile double *pa = rb.part + (size_t)slot[k] * rb.max_blocks;
        for (int j = threadIdx.x; j < (int)gridDim.x; j += PB_VEC_THREADS) {
            const double va = pa[j];
            sa = is_min ? ((va < sa) ? va : sa) : sa + va;
        }
        v[k] = is_min ? red_block<true>(sa) : red_block<false>(sa);
    }
    if (threadIdx.x == 0) {
        if (rb.peer) {
            // one exchange per value (they share the kernel's ticket; the buffer is indexed by slot)
#pragma unroll
            for (int k = 0; k < NV; k++) {
                if (slot[k] < 0) continue;
                double dummy = 0.0;
                if (MIN_0 && k == 0) red_exchange<true>(rb, v[k], slot[k], dummy, -1);
                else red_exchange<false>(rb, v[k], slot[k], dummy, -1);
            }
        }
        double *hp = const_cast<double *>(rb.hsc);
#pragma unroll
        for (int k = 0; k < NV; k++)
            if (slot[k] >= 0) { rb.sc[slot[k]] = v[k]; st_pair(hp + 2 * slot[k], v[k], rb.seq); }
        *rb.counter = 0u;
        if (rb.gate) { fence_gpu(); *reinterpret_cast<volatile double *>(rb.gate) = rb.seq; }
    }
}

#define PB_GRID_STRIDE(i, n)                                                  \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x,      \
                   stride_ = (long long)gridDim.x * blockDim.x;               \
         i < (n); i += stride_)

// The same walk, PB_VB grid-stride steps at a time: the kernels below first load the
// inputs of all PB_VB steps (independent loads, PB_VB x the bytes in flight per thread)
// and then compute / store / accumulate them in step order.  In the plain loop the
// in-place store of step u keeps the compiler from hoisting the loads of step u+1 above
// it (ncu r01: 4.0-4.4 TB/s for the in-place kernels against 6.0 TB/s for k_krylov_b).
// A thread visits the same elements in the same order, so every sum is bitwise unchanged.
#ifndef PB_VB
#define PB_VB 4
#endif
#ifndef PB_VB_MINB
#define PB_VB_MINB 4     // resident CTAs the batched kernels are compiled for (<= 64 registers)
#endif
#define PB_GRID_STRIDE_BATCH(i0, n)                                           \
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x,     \
                   stride_ = (long long)gridDim.x * blockDim.x;               \
         i0 < (n); i0 += PB_VB * stride_)
#define PB_BATCH(u, i, i0, n)                                                 \
    _Pragma("unroll") for (int u = 0; u < PB_VB; u++)                         \
        for (long long i = i0 + u * stride_, once_ = 1; once_ && i < (n); once_ = 0)

// cvEwtSetSS (cvode.c:4081-4090: Abs, Scale, AddConst, Min, Inv) fused with the
// "too much accuracy" norm N_VWrmsNorm(zn[0], ewt) (cvode.c:1376)
static __global__ void __launch_bounds__(PB_VEC_THREADS, PB_VB_MINB)
k_ewt(long long n, double reltol, double abstol, const double *__restrict__ y,
      double *__restrict__ ewt, RedBuf rb)
{
    pdl_enter();
    double mn = __longlong_as_double(0x7ff0000000000000LL), s = 0.0;
    PB_GRID_STRIDE_BATCH(i0, n) {
        double yv[PB_VB];
        PB_BATCH(u, i, i0, n) yv[u] = y[i];
        PB_BATCH(u, i, i0, n) {
        const double yi = yv[u];
        double t = fabs(yi);
        t = reltol * t;
        t = t + abstol;
        mn = (t < mn) ? t : mn;
        const double w = 1.0 / t;
        ewt[i] = w;
        const double p = yi * w;
        s += p * p;
        }
    }
    red_finish<true>(rb, mn, SC_EWT_MIN, s, SC_EWT_NRM);
}

struct ZnPtrs { double *z[6]; };
struct Coef6 { double c[6]; };

// cvPredict (cvode.c:2285-2287) / cvRestore (:2887-2889): the q(q+1)/2 in-place
// N_VLinearSum(1, zn[j-1], +-1, zn[j], zn[j-1]) per component, in registers
// RESC: the cvRescale that precedes the prediction (cvode.c:2257-2260, zn[j] *= eta^j with the factors
// from the host) is applied to the loaded values first -- the same products the separate k_rescale
// launch stored and this kernel read back
template <int SIGN, bool RESC>
__global__ void __launch_bounds__(PB_VEC_THREADS, RESC ? 3 : PB_VB_MINB)
k_predict(long long n, int q, ZnPtrs zn, Coef6 f)
{
    pdl_enter();
    // two grid-stride steps at a time: all loads of both before the first in-place store
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x,
                   stride_ = (long long)gridDim.x * blockDim.x; i0 < n; i0 += 2 * stride_) {
        double z[2][6];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
#pragma unroll
                for (int j = 0; j < 6; j++) if (j <= q) z[u][j] = zn.z[j][i];
                if (SIGN > 0 && RESC) {
#pragma unroll
                    for (int j = 1; j < 6; j++) if (j <= q) z[u][j] = f.c[j] * z[u][j];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
                for (int k = 1; k <= q; k++)
#pragma unroll
                    for (int j = 5; j >= 1; j--)
                        if (j <= q && j >= k) z[u][j - 1] = (SIGN > 0) ? z[u][j - 1] + z[u][j] : z[u][j - 1] - z[u][j];
#pragma unroll
                for (int j = 0; j < 5; j++) if (j < q) zn.z[j][i] = z[u][j];
                if (SIGN > 0 && RESC) zn.z[q][i] = z[u][q];      // rescaled, otherwise untouched by the prediction
            }
        }
    }
}

// cvRescale (cvode.c:2257-2260): zn[j] *= eta^j, j = 1..q (factors from the host)
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_rescale(long long n, int q, ZnPtrs zn, Coef6 f)
{
    pdl_enter();
    PB_GRID_STRIDE(i, n) {
#pragma unroll
        for (int j = 1; j < 6; j++) if (j <= q) zn.z[j][i] = f.c[j] * zn.z[j][i];
    }
}

// cvNlsNewton/cvNewtonIteration residual (cvode.c:2700-2701, 2744-2745) fused
// with CVSpgmrSolve's norm test (cvode_spgmr.c:370) and SpgmrSolve's initial
// residual scaling (sundials_spgmr.c:209-233):
//   first : acor = 0 ; y = zn0
//   tempv = rl1*zn1 + acor ; b = gamma*ftemp - tempv
//   V0 = ewt * b ;  S = sum (b*ewt)^2   [bnorm = sqrt(S/N), beta = sqrt(S)]
template <bool FIRST>
__global__ void __launch_bounds__(PB_VEC_THREADS, PB_VB_MINB)
k_newton_res(long long n, double rl1, double gamma, const double *__restrict__ zn0,
             const double *__restrict__ zn1, const double *__restrict__ ftemp,
             const double *__restrict__ ewt, double *__restrict__ acor, double *__restrict__ y,
             double *__restrict__ b, double *__restrict__ V0, RedBuf rb)
{
    pdl_enter();
    double s = 0.0;
    PB_GRID_STRIDE_BATCH(i0, n) {
        double z0[PB_VB], z1[PB_VB], a0[PB_VB], ft[PB_VB], ew[PB_VB];
        PB_BATCH(u, i, i0, n) {
            if (FIRST) z0[u] = zn0[i]; else a0[u] = acor[i];
            z1[u] = zn1[i]; ft[u] = ftemp[i]; ew[u] = ewt[i];
        }
        PB_BATCH(u, i, i0, n) {
            double ac;
            if (FIRST) { ac = 0.0; acor[i] = 0.0; y[i] = z0[u]; }
            else ac = a0[u];
            double t = rl1 * z1[u] + ac;
            t = gamma * ft[u] - t;
            b[i] = t;
            const double p = ew[u] * t;
            V0[i] = p;
            s += p * p;
        }
    }
    red_finish<false>(rb, s, SC_BSUM, 0.0, -1);
}

// Head of CVSpgmrSolve when it is used as the lsolve hook of an external CVODE
// (pihm_b200_spgmr_solve): the part of k_newton_res that belongs to the linear solver,
//   V0 = weight * b ;  S = sum (b*weight)^2   [bnorm = sqrt(S/N), beta = sqrt(S)]
// (cvode_spgmr.c:370, sundials_spgmr.c:209-233 with x0 = 0, no preconditioner)
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_lsolve_head(long long n, const double *__restrict__ b, const double *__restrict__ w,
              double *__restrict__ V0, RedBuf rb)
{
    pdl_enter();
    double s = 0.0;
    PB_GRID_STRIDE(i, n) {
        const double p = w[i] * b[i];
        V0[i] = p;
        s += p * p;
    }
    red_finish<false>(rb, s, SC_BSUM, 0.0, -1);
}

// Krylov step, part a (sundials_spgmr.c:264 or :341, then :278):
//   V[l] = c * V[l]   (normalisation, c = 1/r_norm or 1/Hes[l][l-1])
//   vtemp = V[l] / ewt ;  SC_VNRM = sum (vtemp*ewt)^2   (cvode_spils.c:679)
static __global__ void __launch_bounds__(PB_VEC_THREADS, PB_VB_MINB)
k_krylov_a(long long n, double c, double *__restrict__ Vl, const double *__restrict__ ewt,
           double *__restrict__ vtemp, RedBuf rb)
{
    pdl_enter();
    double s = 0.0;
    PB_GRID_STRIDE_BATCH(i0, n) {
        double vl[PB_VB], ew[PB_VB];
        PB_BATCH(u, i, i0, n) { vl[u] = Vl[i]; ew[u] = ewt[i]; }
        PB_BATCH(u, i, i0, n) {
            const double v = c * vl[u];
            Vl[i] = v;
            const double w = ew[u];
            const double t = v / w;
            vtemp[i] = t;
            const double p = t * w;
            s += p * p;
        }
    }
    red_finish<false>(rb, s, SC_VNRM, 0.0, -1);
}

// part b (cvode_spils.c:679-684): sig = 1/||vtemp||_wrms ; work = sig*vtemp + y
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_krylov_b(long long n, double n_global, const double *__restrict__ sc,
           const double *__restrict__ vtemp, const double *__restrict__ y,
           double *__restrict__ work)
{
    pdl_enter();
    const double sig = 1.0 / sqrt(sc[SC_VNRM] / n_global);
    PB_GRID_STRIDE(i, n) work[i] = sig * vtemp[i] + y[i];
}

// part c (cvode_spils.c:697-699 VScaleDiff, :619 VLin1, sundials_spgmr.c:305-311)
//   Jv = siginv*(Jv - fy) ; z = (-gamma)*Jv + vtemp ; V[l+1] = ewt * z
//   SC_VK2 = dot(V[l+1],V[l+1]) ; SC_H0 = dot(V[0],V[l+1])   (ModifiedGS :50,:57)
static __global__ void __launch_bounds__(PB_VEC_THREADS, PB_VB_MINB)
k_krylov_c(long long n, double n_global, double gamma, const double *__restrict__ sc,
           const double *__restrict__ vtemp, const double *__restrict__ fy,
           const double *__restrict__ ewt, const double *__restrict__ V0,
           double *__restrict__ Vk, RedBuf rb)
{
    pdl_enter();
    const double sig = 1.0 / sqrt(sc[SC_VNRM] / n_global);
    const double siginv = 1.0 / sig;
    const double mg = -gamma;
    double s = 0.0, h = 0.0;
    PB_GRID_STRIDE_BATCH(i0, n) {
        double a[PB_VB], f[PB_VB], vt[PB_VB], ew[PB_VB], v0[PB_VB];
        PB_BATCH(u, i, i0, n) { a[u] = Vk[i]; f[u] = fy[i]; vt[u] = vtemp[i]; ew[u] = ewt[i]; v0[u] = V0[i]; }
        PB_BATCH(u, i, i0, n) {
            double jv = siginv * (a[u] - f[u]);
            double z = mg * jv + vt[u];
            z = ew[u] * z;
            Vk[i] = z;
            s += z * z;
            h += v0[u] * z;
        }
    }
    red_finish<false>(rb, s, SC_VK2, h, SC_H0);
}

// Modified Gram-Schmidt step (sundials_iterative.c:56-63), chained on device:
//   V[k] += (-h_prev) * V[prev]          (Vaxpy form of N_VLinearSum)
//   next dot: SC_H0+inext = dot(V[inext], V[k])   or, when Vnext == V[k] itself,
//   SC_NEW2 = dot(V[k], V[k])
static __global__ void __launch_bounds__(PB_VEC_THREADS, PB_VB_MINB)
k_mgs_step(long long n, const double *__restrict__ sc, int slot_prev,
           const double *__restrict__ Vprev, const double *Vnext, double *Vk,
           int slot_next, RedBuf rb)
{
    pdl_enter();
    const double mh = -sc[slot_prev];
    const bool self = (Vnext == Vk);
    double s = 0.0;
    PB_GRID_STRIDE_BATCH(i0, n) {
        double a[PB_VB], p[PB_VB], x[PB_VB];
        PB_BATCH(u, i, i0, n) { a[u] = Vk[i]; p[u] = Vprev[i]; x[u] = self ? 0.0 : Vnext[i]; }
        PB_BATCH(u, i, i0, n) {
            const double v = a[u] + mh * p[u];
            Vk[i] = v;
            s += (self ? v : x[u]) * v;
        }
    }
    red_finish<false>(rb, s, slot_next, 0.0, -1);
}

struct KryPtrs { const double *v[5]; };

// The whole modified Gram-Schmidt chain of one Krylov iteration (nsteps = l + 1 k_mgs_step
// launches) in ONE cooperative launch: every thread keeps its elements of V[k] in shared memory
// across the steps, so a step reads V[i] (for the update) and V[i+1] (for the next dot product)
// and nothing else -- 1-2 vector passes instead of 4.  Same grid and block shape as k_mgs_step and
// the same per-thread order of operations and of the summation, hence the same bits.  Between the
// steps the CTAs wait at a device-side gate for the ticket of the reduction that produces the next
// coefficient (red_finish publishes it; in a partitioned run after the exchange over peer memory).
// Requires all CTAs to be resident (cooperative launch) and slice <= the shared memory given.
static __global__ void __launch_bounds__(PB_VEC_THREADS, 4)
k_mgs_chain(long long n, int nsteps, KryPtrs V, double *Vk, RedBuf rb, int per_thread)
{
    extern __shared__ double wsh[];
    const long long stride_ = (long long)gridDim.x * blockDim.x;
    const long long i_first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    {
        int k = 0;
        for (long long i = i_first; i < n; i += stride_, k++) wsh[k * PB_VEC_THREADS + threadIdx.x] = Vk[i];
    }
    const double seq0 = rb.seq;
    const volatile double *gate = rb.gate;
    const volatile double *sc = rb.sc;
    for (int st = 0; st < nsteps; st++) {
        if (st > 0) {
            if (threadIdx.x == 0) {
                const double want = seq0 + (double)(st - 1);
                while (*gate != want) { }
                fence_gpu();
            }
            __syncthreads();
        }
        const double mh = -sc[SC_H0 + st];
        const bool self = (st == nsteps - 1);
        const double *__restrict__ Vprev = V.v[st];
        const double *__restrict__ Vnext = self ? nullptr : V.v[st + 1];
        double s = 0.0;
        // PB_VB grid-stride steps at a time: all global loads of a batch first (see PB_GRID_STRIDE_BATCH)
        int k0 = 0;
        for (long long i0 = i_first; i0 < n; i0 += PB_VB * stride_, k0 += PB_VB) {
            double p[PB_VB], x[PB_VB];
#pragma unroll
            for (int u = 0; u < PB_VB; u++) {
                const long long i = i0 + u * stride_;
                if (i < n) { p[u] = Vprev[i]; x[u] = self ? 0.0 : Vnext[i]; }
            }
#pragma unroll
            for (int u = 0; u < PB_VB; u++) {
                const long long i = i0 + u * stride_;
                if (i < n) {
                    double *wp = wsh + (k0 + u) * PB_VEC_THREADS + threadIdx.x;
                    const double v = *wp + mh * p[u];
                    *wp = v;
                    s += (self ? v : x[u]) * v;
                }
            }
        }
        RedBuf r = rb;
        r.seq = seq0 + (double)st;
        red_finish<false>(r, s, self ? SC_NEW2 : SC_H0 + st + 1, 0.0, -1);
    }
    {
        int k = 0;
        for (long long i = i_first; i < n; i += stride_, k++) Vk[i] = wsh[k * PB_VEC_THREADS + threadIdx.x];
    }
    (void)per_thread;
}

// End of SpgmrSolve + CVSpgmrSolve + Newton update, fused:
//   xcor = sum_k yg[k]*V[k]  (Vaxpy chain from 0, sundials_spgmr.c:348-357)
//   xcor = xcor / ewt ; x = 0 + xcor ; b = x         (:366-378, cvode_spgmr.c:389)
//   del^2 sum = sum (b*ewt)^2 ; acor += b ; y = zn0 + acor   (cvode.c:2762-2764)
//   and sum (acor*ewt)^2 of the new acor: acnrm of cvode.c:2779 if the iteration converges (saves k_wsq + a sync)
static __global__ void __launch_bounds__(PB_VEC_THREADS, PB_VB_MINB)
k_spgmr_final(long long n, int krydim, KryPtrs V, Coef6 yg, const double *__restrict__ ewt,
              const double *__restrict__ zn0, double *__restrict__ acor, double *__restrict__ y,
              RedBuf rb)
{
    pdl_enter();
    double s = 0.0, s_ac = 0.0;
    constexpr int VB2 = 2;      // 8 input streams: two steps at a time
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x,
                   stride_ = (long long)gridDim.x * blockDim.x; i0 < n; i0 += VB2 * stride_) {
        double v[VB2][5], ew[VB2], a0[VB2], z0[VB2];
#pragma unroll
        for (int u = 0; u < VB2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
#pragma unroll
                for (int k = 0; k < 5; k++) if (k < krydim) v[u][k] = V.v[k][i];
                ew[u] = ewt[i]; a0[u] = acor[i]; z0[u] = zn0[i];
            }
        }
#pragma unroll
        for (int u = 0; u < VB2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
                double xc = 0.0;
#pragma unroll
                for (int k = 0; k < 5; k++) if (k < krydim) xc = xc + yg.c[k] * v[u][k];
                const double w = ew[u];
                xc = xc / w;
                const double b = 0.0 + xc;
                const double p = b * w;
                s += p * p;
                const double ac = a0[u] + b;
                acor[i] = ac;
                y[i] = z0[u] + ac;
                const double pa = ac * w;       // N_VWrmsNorm(acor, ewt) of cvode.c:2779, in case this iteration converges
                s_ac += pa * pa;
            }
        }
    }
    red_finish<false>(rb, s, SC_DEL, s_ac, SC_ACNRM);
}

// Tail of the lsolve hook: the solver part of k_spgmr_final,
//   xcor = sum_k yg[k]*V[k] ; xcor /= weight ; x = 0 + xcor ; b = x
// (sundials_spgmr.c:348-378, cvode_spgmr.c:389); krydim == 0: b = 0 (x stayed 0)
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_lsolve_tail(long long n, int krydim, KryPtrs V, Coef6 yg, const double *__restrict__ w, double *__restrict__ b)
{
    pdl_enter();
    PB_GRID_STRIDE(i, n) {
        double xc = 0.0;
#pragma unroll
        for (int k = 0; k < 5; k++) if (k < krydim) xc = xc + yg.c[k] * V.v[k][i];
        if (krydim > 0) xc = xc / w[i];
        b[i] = 0.0 + xc;
    }
}

// Newton update when the linear solve returned b itself (early outs of
// CVSpgmrSolve, cvode_spgmr.c:371-374 and sundials_spgmr.c:238-239):
//   ZERO_B: b = 0 first.   del^2 = sum (b*ewt)^2 ; acor += b ; y = zn0 + acor
template <bool ZERO_B>
__global__ void __launch_bounds__(PB_VEC_THREADS)
k_newton_update(long long n, const double *__restrict__ bvec, const double *__restrict__ ewt,
                const double *__restrict__ zn0, double *__restrict__ acor, double *__restrict__ y,
                RedBuf rb)
{
    pdl_enter();
    double s = 0.0, s_ac = 0.0;
    PB_GRID_STRIDE(i, n) {
        const double b = ZERO_B ? 0.0 : bvec[i];
        const double w = ewt[i];
        const double p = b * w;
        s += p * p;
        const double ac = acor[i] + b;
        acor[i] = ac;
        y[i] = zn0[i] + ac;
        const double pa = ac * w;           // N_VWrmsNorm(acor, ewt), cvode.c:2779
        s_ac += pa * pa;
    }
    red_finish<false>(rb, s, SC_DEL, s_ac, SC_ACNRM);
}

// sum (x*w)^2 into one slot; optionally a second vector into a second slot
// (N_VWrmsNorm, nvector_serial.c:669-686; pairs: cvBDFStab cvode.c:3267-3268)
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_wsq(long long n, const double *__restrict__ x1, const double *x2, const double *__restrict__ w,
      int slot1, int slot2, RedBuf rb)
{
    pdl_enter();
    double s1 = 0.0, s2 = 0.0;
    PB_GRID_STRIDE(i, n) {
        const double wi = w[i];
        const double p = x1[i] * wi;
        s1 += p * p;
        if (slot2 >= 0) { const double r = x2[i] * wi; s2 += r * r; }
    }
    red_finish<false>(rb, s1, slot1, s2, slot2);
}

// cvCompleteStep (cvode.c:3010-3016): zn[j] += l[j]*acor (Vaxpy), j = 0..q,
// and the optional save zn[qmax] = acor
static __global__ void __launch_bounds__(PB_VEC_THREADS, PB_VB_MINB)
k_complete(long long n, int q, ZnPtrs zn, Coef6 l, const double *__restrict__ acor, double *save)
{
    pdl_enter();
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x,
                   stride_ = (long long)gridDim.x * blockDim.x; i0 < n; i0 += 2 * stride_) {
        double z[2][6], a[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
                a[u] = acor[i];
#pragma unroll
                for (int j = 0; j < 6; j++) if (j <= q) z[u][j] = zn.z[j][i];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
#pragma unroll
                for (int j = 0; j < 6; j++) if (j <= q) zn.z[j][i] = z[u][j] + l.c[j] * a[u];
                if (save) save[i] = a[u];
            }
        }
    }
}

// cvCompleteStep fused with everything the rest of the step reads from the vectors it has just written:
//   * the NEXT step's error weights from the new zn[0] (cvEwtSetSS, cvode.c:4081-4090) into a second buffer
//     (this step's k_eta / stability norms still need the old ones; the host swaps the pointers at the top
//     of the next step) with min(reltol |y| + abstol) and the tolsf norm sum (zn0 * ewt_next)^2 (cvode.c:1376),
//   * sum (zn[q] * ewt)^2 (cvComputeEtaqm1 cvode.c:3097 and cvBDFStab :3267), sum (zn[q-1] * ewt)^2 (:3268),
//   * sum ((acor - cquot zn[qmax]) * ewt)^2 (cvComputeEtaqp1 :3118-3119),
// each with the operations and the per-thread summation order of k_ewt / k_wsq / k_eta, whose launches (and
// the host synchronisations behind two of them) it replaces.
struct CompleteNorms {
    double reltol, abstol, cquot;
    double *ewt_next;
    const double *ewt, *znmax;
    int do_zq, do_zqm1, do_p1;
};
static __global__ void __launch_bounds__(PB_VEC_THREADS, 3)
k_complete_norms(long long n, int q, ZnPtrs zn, Coef6 l, const double *__restrict__ acor, double *save,
                 CompleteNorms cn, RedBuf rb)
{
    pdl_enter();
    double acc[5] = {__longlong_as_double(0x7ff0000000000000LL), 0.0, 0.0, 0.0, 0.0};
    const double mc = -cn.cquot;
    const bool any_w = cn.do_zq || cn.do_zqm1 || cn.do_p1;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x,
                   stride_ = (long long)gridDim.x * blockDim.x; i0 < n; i0 += 2 * stride_) {
        double z[2][6], a[2], w[2] = {0.0, 0.0}, zm[2] = {0.0, 0.0};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
                a[u] = acor[i];
#pragma unroll
                for (int j = 0; j < 6; j++) if (j <= q) z[u][j] = zn.z[j][i];
                if (any_w) w[u] = cn.ewt[i];
                if (cn.do_p1) zm[u] = cn.znmax[i];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long i = i0 + u * stride_;
            if (i < n) {
                double zq = 0.0, zqm1 = 0.0;
#pragma unroll
                for (int j = 0; j < 6; j++)
                    if (j <= q) {
                        const double v = z[u][j] + l.c[j] * a[u];
                        zn.z[j][i] = v;
                        z[u][j] = v;
                        if (j == q) zq = v;
                        if (j == q - 1) zqm1 = v;
                    }
                if (save) save[i] = a[u];
                {
                    const double yi = z[u][0];
                    double t = fabs(yi);
                    t = cn.reltol * t;
                    t = t + cn.abstol;
                    acc[0] = (t < acc[0]) ? t : acc[0];
                    const double wn = 1.0 / t;
                    cn.ewt_next[i] = wn;
                    const double p = yi * wn;
                    acc[1] += p * p;
                }
                if (cn.do_zq) { const double p = zq * w[u]; acc[2] += p * p; }
                if (cn.do_zqm1) { const double r = zqm1 * w[u]; acc[3] += r * r; }
                if (cn.do_p1) { const double t = mc * zm[u] + a[u]; const double p = t * w[u]; acc[4] += p * p; }
            }
        }
    }
    const int slot[5] = {SC_EWT_MIN, SC_EWT_NRM, cn.do_zq ? SC_STAB1 : -1, cn.do_zqm1 ? SC_STAB2 : -1,
                         cn.do_p1 ? SC_ETA_P1 : -1};
    red_finish_n<5, true>(rb, acc, slot);
}

// cvComputeEtaqm1 / cvComputeEtaqp1 norms (cvode.c:3097, :3118-3119):
//   SC_ETA_M1 = sum (znq*ewt)^2 ; tempv = (-cquot)*znmax + acor ; SC_ETA_P1 = sum (tempv*ewt)^2
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_eta(long long n, int do_m1, int do_p1, double cquot, const double *znq, const double *znmax,
      const double *__restrict__ acor, const double *__restrict__ ewt, RedBuf rb)
{
    pdl_enter();
    double s1 = 0.0, s2 = 0.0;
    const double mc = -cquot;
    PB_GRID_STRIDE(i, n) {
        const double w = ewt[i];
        if (do_m1) { const double p = znq[i] * w; s1 += p * p; }
        if (do_p1) { const double t = mc * znmax[i] + acor[i]; const double p = t * w; s2 += p * p; }
    }
    red_finish<false>(rb, s1, SC_ETA_M1, s2, SC_ETA_P1);
}

// CVodeGetDky with k = 0 (cvode.c:1545-1556): Horner form, VLin1 per term
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_dky(long long n, int q, double s, ZnPtrs zn, double *__restrict__ dky)
{
    pdl_enter();
    PB_GRID_STRIDE(i, n) {
        double d = 0.0;
#pragma unroll
        for (int j = 5; j >= 0; j--)
            if (j <= q) d = (j == q) ? zn.z[j][i] : s * d + zn.z[j][i];
        dky[i] = d;
    }
}

}  // namespace pb
