// nvec.cuh -- device-resident N_Vector kernels (streaming + reductions).
//
// Component arithmetic is that of cvode/src/nvec_ser/nvector_serial.c:421-770.
// Every special case of N_VLinearSum_Serial (Vaxpy/VSum/VDiff/VLin1/VLin2) is
// bitwise equal to (a*x)+(b*y) evaluated without FMA, except VScaleSum
// a*(x+y) and VScaleDiff a*(x-y) (:469-480), which get their own mode.  This
// translation unit is compiled with -fmad=false.
//
// Reductions are deterministic: fixed per-thread strides, shuffle tree inside
// a warp, fixed-order sum of the block partials by the last block to finish.
#pragma once
#include "common.cuh"

namespace pb {

// programmatic dependent launch: let the next kernel of the stream start launching, then wait
// until everything the previous one wrote is visible (both are no-ops for a plain launch)
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}


#define PB_VEC_THREADS 256

enum { LS_GENERAL = 0, LS_SCALESUM = 1, LS_SCALEDIFF = 2 };

template <int MODE>
__global__ void __launch_bounds__(PB_VEC_THREADS)
k_linearsum(long long n, double a, const double *__restrict__ x, double b,
            const double *__restrict__ y, double *__restrict__ z)
{
    pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        if (MODE == LS_GENERAL) z[i] = (a * x[i]) + (b * y[i]);
        else if (MODE == LS_SCALESUM) z[i] = a * (x[i] + y[i]);
        else z[i] = a * (x[i] - y[i]);
    }
}

// in-place variants must not carry __restrict__ on aliased pointers
template <int MODE>
__global__ void __launch_bounds__(PB_VEC_THREADS)
k_linearsum_alias(long long n, double a, const double *x, double b, const double *y, double *z)
{
    pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double xv = x[i], yv = y[i];
        if (MODE == LS_GENERAL) z[i] = (a * xv) + (b * yv);
        else if (MODE == LS_SCALESUM) z[i] = a * (xv + yv);
        else z[i] = a * (xv - yv);
    }
}

enum { EW_CONST = 0, EW_SCALE, EW_ABS, EW_INV, EW_ADDCONST, EW_PROD, EW_DIV, EW_COPY };

template <int OP>
__global__ void __launch_bounds__(PB_VEC_THREADS)
k_elementwise(long long n, double c, const double *x, const double *y, double *z)
{
    pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        if (OP == EW_CONST) z[i] = c;
        else if (OP == EW_SCALE) z[i] = c * x[i];
        else if (OP == EW_ABS) z[i] = fabs(x[i]);
        else if (OP == EW_INV) z[i] = 1.0 / x[i];
        else if (OP == EW_ADDCONST) z[i] = x[i] + c;
        else if (OP == EW_PROD) z[i] = x[i] * y[i];
        else if (OP == EW_DIV) z[i] = x[i] / y[i];
        else z[i] = x[i];
    }
}

// ---- reductions ------------------------------------------------------------
enum { RD_DOT = 0, RD_WSQ, RD_MAXABS, RD_MIN };

template <int OP> __device__ __forceinline__ double rd_identity()
{
    return (OP == RD_MIN) ? __longlong_as_double(0x7ff0000000000000LL) : 0.0;
}
template <int OP> __device__ __forceinline__ double rd_combine(double a, double b)
{
    if (OP == RD_MAXABS) return (b > a) ? b : a;
    if (OP == RD_MIN) return (b < a) ? b : a;
    return a + b;
}
template <int OP> __device__ __forceinline__ double rd_term(double x, double y)
{
    if (OP == RD_DOT) return x * y;
    if (OP == RD_WSQ) { double p = x * y; return p * p; }
    if (OP == RD_MAXABS) return fabs(x);
    return x;
}

template <int OP>
__device__ __forceinline__ double block_reduce(double v)
{
    __shared__ double sh[PB_VEC_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = rd_combine<OP>(v, __shfl_down_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < PB_VEC_THREADS / 32) ? sh[lane] : rd_identity<OP>();
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v = rd_combine<OP>(v, __shfl_down_sync(0xffffffffu, v, o));
    }
    return v;   // valid in thread 0
}

// One launch per reduction: block partials -> `part`, the last block to arrive
// sums them in index order and writes the scalar to out_dev (device) and
// out_host (mapped pinned host memory).  `counter` must be 0 on entry and is
// reset on exit.  POST: 0 raw, 1 sqrt(v / n)  (N_VWrmsNorm, :669-686).
template <int OP, int POST>
__global__ void __launch_bounds__(PB_VEC_THREADS)
k_reduce(long long n, const double *__restrict__ x, const double *__restrict__ y,
         double *part, unsigned int *counter, double *out_dev, volatile double *out_host)
{
    pdl_enter();
    double v = rd_identity<OP>();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double xv = x[i];
        const double yv = (OP == RD_DOT || OP == RD_WSQ) ? y[i] : 0.0;
        v = rd_combine<OP>(v, rd_term<OP>(xv, yv));
    }
    v = block_reduce<OP>(v);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = v;
        fence_gpu();
        const unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    fence_gpu();
    double s = rd_identity<OP>();
    for (int b = threadIdx.x; b < (int)gridDim.x; b += PB_VEC_THREADS)
        s = rd_combine<OP>(s, ((volatile double *)part)[b]);
    s = block_reduce<OP>(s);
    if (threadIdx.x == 0) {
        if (POST == 1) s = sqrt(s / (double)n);
        *out_dev = s;
        if (out_host) *out_host = s;
        *counter = 0u;
    }
}

// reference order <-> internal order of the state blocks
static __global__ void __launch_bounds__(PB_VEC_THREADS)
k_permute_state(int ne, int nr, int fbr, const int *__restrict__ perm,
                const double *__restrict__ src, double *__restrict__ dst, int to_internal)
{
    pdl_enter();
    const long long n = (long long)(fbr ? 5 : 3) * ne + 2LL * nr;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        long long blk_off, k;
        const long long e3 = 3LL * ne, r2 = 2LL * nr;
        if (i < e3) { blk_off = (i / ne) * ne; k = i - blk_off; }
        else if (i < e3 + r2) { dst[i] = src[i]; continue; }
        else { blk_off = e3 + r2 + ((i - e3 - r2) / ne) * ne; k = i - blk_off; }
        // perm[internal] = reference
        if (to_internal) dst[i] = src[blk_off + perm[k]];
        else dst[blk_off + perm[k]] = src[i];
    }
}

// one forcing column, reference order -> internal order; hot columns (pcpdrp, edir, ett,
// ws0.surf) go to the warp-tiled table [tile][4][32], bc columns to the flat table
static __global__ void __launch_bounds__(256)
k_scatter_forcing(int ne, int nes, int col, const int *__restrict__ perm, const double *__restrict__ src,
                  double *__restrict__ ft, double *__restrict__ forc)
{
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    const double v = src[perm[i]];
    if (col <= PB_F_WS0SURF) ft[((size_t)(i >> 5) * 4 + col) * 32 + (i & 31)] = v;
    else forc[(size_t)col * nes + i] = v;
}

}  // namespace pb
