// cvode_b200.cu -- variable-order BDF + Newton + scaled GMRES(maxl) with all
// vector work on the device.
//
// The algorithm is the one the reference reaches through SetCVodeParam()/SolveCVode()
// (src/ode.c:340-498): CVODE 2.9.0's BDF/Newton stepper (cvode/src/cvode/cvode.c), its
// CVSPGMR glue (cvode_spgmr.c, cvode_spils.c) and the generic SPGMR solver
// (cvode/src/sundials/sundials_spgmr.c, sundials_iterative.c), restricted to what MM-PIHM
// configures: CV_BDF, CV_NEWTON, scalar tolerances, PREC_NONE, modified Gram-Schmidt, no
// restarts, difference-quotient J*v, tstop mode, stability-limit detection on.
//
// What is new here is the device side: the N_V* calls are replaced by the fused kernels of
// cvode_kernels.cuh, scalars stay on the device, the host only synchronises where CVODE
// branches on a norm.  The HOST control flow is a restatement of CVODE's: every branch,
// counter and heuristic constant follows the reference so that step sequences agree bit for
// bit (tests/test_dropin_gpu.py), and the scalar routines cvSLdet (stability-limit detection,
// cvode.c:3335-3602), cvSetBDF / cvSetTqBDF (cvode.c:2400-2531) and cvAdjustOrder /
// cvIncreaseBDF / cvDecreaseBDF (cvode.c:2106-2241) follow cvode.c statement by statement.
// That code is derived from SUNDIALS:
//
//   Copyright (c) 2002-2016, Lawrence Livermore National Security.
//   Produced at the Lawrence Livermore National Laboratory.
//   Written by A.C. Hindmarsh, D.R. Reynolds, R. Serban, C.S. Woodward, S.D. Cohen,
//   A.G. Taylor, S. Peles, L.E. Banks, and D. Shumaker.  UCRL-CODE-155951 (CVODE).
//   All rights reserved.  BSD 3-clause license: see THIRD_PARTY_NOTICES.md at the
//   repository root for the full text, which applies to those portions.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <utility>
#include "common.cuh"
#include "cvode_kernels.cuh"

using namespace pb;

namespace {

// cvode.c:60-330 constants
constexpr double FUZZ_FACTOR = 100.0, CORTES = 0.1, THRESH = 1.5, ETAMX1 = 10000.0;
constexpr double ETAMX2 = 10.0, ETAMX3 = 10.0, ETAMXF = 0.2, ETAMIN = 0.1, ETACF = 0.25;
constexpr double ADDON = 0.000001, BIAS1 = 6.0, BIAS2 = 6.0, BIAS3 = 10.0, ONEPSM = 1.000001;
constexpr double CRDOWN = 0.3, RDIV = 2.0, TINY = 1.0e-10, HUNDRED = 100.0;
constexpr int SMALL_NST = 10, MXNCF = 10, MXNEF = 7, MXNEF1 = 3, SMALL_NEF = 2, LONG_WAIT = 10;
constexpr int NLS_MAXCOR = 3, BDF_Q_MAX = 5, MXHNIL_DEFAULT = 10;
constexpr double CVSPILS_EPLIN = 0.05;
constexpr int CVSPILS_MAXL = 5;
// return flags (cvode/include/cvode/cvode.h:91-131)
constexpr int CV_SUCCESS = 0, CV_TSTOP_RETURN = 1, CV_TOO_MUCH_WORK = -1, CV_TOO_MUCH_ACC = -2;
constexpr int CV_ERR_FAILURE = -3, CV_CONV_FAILURE = -4, CV_LSOLVE_FAIL = -7, CV_RHSFUNC_FAIL = -8;
constexpr int CV_ILL_INPUT = -22, CV_BAD_T = -25;
// internal flags (cvode.c:104-114)
constexpr int DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, CONV_FAIL = 4, TRY_AGAIN = 5;
constexpr int FIRST_CALL = 6, PREV_CONV_FAIL = 7, PREV_ERR_FAIL = 8;

// cvode/src/sundials/sundials_math.c: SUNRpowerI is a multiplication loop,
// SUNRpowerR / SUNRsqrt clamp non-positive arguments to 0
inline double rpowerI(double base, int exponent)
{
    double prod = 1.0;
    const int expt = std::abs(exponent);
    for (int i = 1; i <= expt; i++) prod *= base;
    if (exponent < 0) prod = 1.0 / prod;
    return prod;
}
inline double rpowerR(double base, double exponent) { return (base <= 0.0) ? 0.0 : std::pow(base, exponent); }
inline double rsqrt_s(double x) { return (x <= 0.0) ? 0.0 : std::sqrt(x); }

}  // namespace

// Every integrator kernel is launched with programmatic stream serialization (PDL): it may
// become resident while its predecessor drains; all of them start with pdl_enter()
// (cvode_kernels.cuh: launch_dependents + wait), so nothing is read or written early.
static thread_local cudaError_t g_launch_error = cudaSuccess;
template <typename... ExpTypes, typename... ActTypes>
static inline void launch_pdl(cudaStream_t st, int pdl, int grid, int block, void (*kernel)(ExpTypes...),
                              ActTypes &&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<ActTypes>(args)...);
    if (e != cudaSuccess && g_launch_error == cudaSuccess) g_launch_error = e;     // first failure wins; solve() reports it
}

struct pihm_b200_cvode {
    pihm_b200_ctx *ctx = nullptr;
    long long N = 0;
    double n_global = 0.0;
    // vectors (cvAllocVectors cvode.c:1640-1698; SpgmrMalloc sundials_spgmr.c:45-157)
    double *zn[6] = {}, *ewt = nullptr, *acor = nullptr, *tempv = nullptr, *ftemp = nullptr;
    double *V[6] = {}, *vtemp = nullptr, *ytemp = nullptr;
    double *y = nullptr;                 // the caller's vector doubles as Newton iterate (cvode.c:1105)
    pihm_b200_vec wrap_a, wrap_b;        // non-owning views handed to pihm_b200_ode
    RedBuf rb{};
    double *d_part = nullptr, *d_sc = nullptr, *h_sc = nullptr;
    unsigned int *d_counter = nullptr;
    int blocks = 1;
    // integrator state (struct CVodeMemRec, cvode_impl.h:62-300)
    double reltol = 0, abstol = 0, uround = DBL_EPSILON;
    double tn = 0, h = 0, hu = 0, hprime = 0, hscale = 0, eta = 0, etamax = 0, next_h = 0, h0u = 0;
    double etaq = 0, etaqm1 = 0, etaqp1 = 0, hin = 0, hmin = 0, hmax_inv = 0, tstop = 0, tolsf = 1;
    double gamma = 0, gammap = 0, gamrat = 1, rl1 = 0, crate = 1, acnrm = 0, nlscoef = CORTES;
    double saved_tq5 = 0, tretlast = 0;
    double l[13] = {}, tq[6] = {}, tau[14] = {}, ssdat[6][4] = {};
    int q = 1, qprime = 1, qu = 0, L = 2, qwait = 2, qmax = BDF_Q_MAX, next_q = 0, nscon = 0;
    int mnewt = 0, maxcor = NLS_MAXCOR, maxnef = MXNEF, maxncf = MXNCF, indx_acor = 0;
    bool tstopset = false, sldeton = true, initialised = false, ewt_pending = false;
    bool resc_pending = false;           // cvRescale factors waiting for the next k_predict
    Coef6 resc_f{};
    // cvCompleteStep fused with the next step's error weights and the norms of cvPrepareNextStep / cvBDFStab
    // (k_complete_norms; PIHM_B200_FOLD=0: the separate k_ewt / k_eta / k_wsq launches)
    bool fold = true, ewt_ready = false, fold_m1 = false, fold_p1 = false, fold_stab = false;
    double *ewt_next = nullptr;
    long long nst = 0, nfe = 0, ncfn = 0, netf = 0, nni = 0, nsetups = 0, nhnil = 0, nor = 0;
    long long mxstep = 500, mxhnil = MXHNIL_DEFAULT;
    // CVSPGMR state
    int maxl = CVSPILS_MAXL;
    double eplifac = CVSPILS_EPLIN, sqrtN = 0, deltar = 0, delta = 0;
    long long nli = 0, ncfl = 0, nfes = 0, njtimes = 0;
    double Hes[6][5] = {}, givens[10] = {}, yg[6] = {};
    int krydim_last = 0;                 // dimension of the last Krylov space
    // the Gram-Schmidt chain of a Krylov iteration as one cooperative launch (k_mgs_chain)
    bool chain_ok = false;
    int chain_per_thread = 0;
    size_t chain_smem = 0;
    // AdjCVodeMaxStep statics (ode.c:506-508)
    long long nst0 = 0, ncfn0 = 0, nni0 = 0;

    cudaStream_t s() const { return ctx->s(); }
    void count(int k = 1) { ctx->launches += k; }
    // Multi-GPU: kernels leave rank-local partial sums in d_sc; red() all-reduces
    // the slots in stream order and sync() then fetches the global values.
    double *h_sc_map = nullptr;          // mapped page the kernels write: [SC_COUNT] pairs {value, ticket}
    // p2p: the fused reduction kernels all-reduce their scalars themselves over peer memory
    // (red_exchange, cvode_kernels.cuh) and write the GLOBAL values to the mapped mirror, like
    // on one GPU; otherwise NCCL does it after the kernel.
    bool p2p = false;
    double *d_xbuf = nullptr;            // this rank's exchange buffer
    double **d_peer = nullptr;           // device array of the ranks' exchange buffers
    void *h_peer[PB_MAX_RANKS] = {};
    void sync()
    {
        if (ctx->nranks > 1 && !p2p)
            cudaMemcpyAsync(h_sc, d_sc, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, ctx->s());
        note_status(cudaStreamSynchronize(ctx->s()));
        if (!(ctx->nranks > 1 && !p2p)) sync_spin_();      // the stream has drained: every pair is there
        else npend = 0;
    }
    // a failed launch or a device fault: remember the first one; the step loop stops at its next check
    cudaError_t dev_error = cudaSuccess;
    void note_status(cudaError_t e) { if (e != cudaSuccess && dev_error == cudaSuccess) dev_error = e; }
    bool failed()
    {
        if (g_launch_error != cudaSuccess) { note_status(g_launch_error); g_launch_error = cudaSuccess; }
        return dev_error != cudaSuccess;
    }
    void red(int slot, int n = 1, int op = 0) { if (ctx->nranks > 1 && !p2p) comm_allreduce(ctx, d_sc + slot, n, op); }
    // a plain k_reduce launch (N_Vector kernel, rank-local result): always NCCL
    void red_nccl(int slot)
    {
        if (ctx->nranks <= 1) return;
        comm_allreduce(ctx, d_sc + slot, 1, 0);
        if (p2p) cudaMemcpyAsync(h_sc + slot, d_sc + slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->s());
    }
    // Every reduction kernel gets a ticket; the kernel's finishing thread stores it to the
    // mapped host mirror after its results.  sync_spin() waits for the newest ticket by polling
    // that word (~2 us) instead of cudaStreamSynchronize (~15 us); use it only when the last
    // kernel launched was a reduction kernel.
    long long seq_ctr = 0;
    // R(slots): the next ticket; the slots the kernel will produce are noted so that the next host
    // synchronisation waits for exactly those pairs
    int pend_slot[2 * SC_COUNT] = {}, npend = 0;
    double pend_seq[2 * SC_COUNT] = {};
    void expect(int slot, double seq)
    {
        for (int k = 0; k < npend; k++) if (pend_slot[k] == slot) { pend_seq[k] = seq; return; }
        pend_slot[npend] = slot; pend_seq[npend] = seq; npend++;
    }
    RedBuf &R(int slotA, int slotB = -1)
    {
        rb.seq = (double)(++seq_ctr);
        expect(slotA, rb.seq);
        if (slotB >= 0) expect(slotB, rb.seq);
        return rb;
    }
    // PIHM_B200_PROFILE=1: host wait time per sync and in-situ RHS event times, printed at destroy
    bool prof = false;
    long long prof_nsync = 0, prof_wait_ns = 0, prof_solve_ns = 0;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    static long long now_ns()
    {
        return std::chrono::duration_cast<std::chrono::nanoseconds>(
                   std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    void sync_spin()
    {
        if (prof) {
            const long long t0 = now_ns();
            sync_spin_();
            prof_wait_ns += now_ns() - t0;
            prof_nsync++;
        } else
            sync_spin_();
    }
    void sync_spin_()
    {
        if (ctx->nranks > 1 && !p2p) { sync(); return; }
        // every pair the kernels since the last synchronisation produce: wait for its ticket, take its value
        for (int k = 0; k < npend; k++) {
            volatile double *pr = h_sc_map + 2 * pend_slot[k];
            const double want = pend_seq[k];
            for (long long it = 0; pr[1] != want; it++) {
                if ((it & 0xfff) == 0xfff && cudaStreamQuery(ctx->s()) != cudaErrorNotReady) {
                    // finished (or failed) without the ticket: stop spinning; a ticket that is still missing
                    // after the stream has drained means the scalar is stale -> fatal, not a branch on garbage
                    note_status(cudaStreamSynchronize(ctx->s()));
                    std::atomic_thread_fence(std::memory_order_acquire);
                    if (pr[1] != want) note_status(cudaErrorUnknown);
                    break;
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            h_sc[pend_slot[k]] = pr[0];
        }
        npend = 0;
    }
    // N_VWrmsNorm = SUNRsqrt(sum / N)  (nvector_serial.c:685)
    double wrms(int slot) const { const double v = h_sc[slot] / n_global; return (v <= 0.0) ? 0.0 : std::sqrt(v); }

    // ---- per-kernel in-situ profile (pihm_b200_cvode_profile(cv, 2); bench.py's vector_roofline) ----
    // Every vector kernel launch is bracketed by a pair of CUDA events (PDL off so the pair sees the
    // kernel alone) and booked under its kernel id together with the bytes it reads and writes
    // (vector passes x 8 N).  Off by default: one branch per launch.
    enum KId { K_EWT = 0, K_PREDICT, K_RESCALE, K_NEWTON_RES, K_KRYLOV_A, K_KRYLOV_B, K_KRYLOV_C, K_MGS_STEP,
               K_MGS_CHAIN, K_SPGMR_FINAL, K_NEWTON_UPDATE, K_WSQ, K_COMPLETE, K_ETA, K_DKY, K_SCALE, K_COPY,
               K_AXPY, K_DOT, K_NKERN };
    struct KAcc { double ms = 0.0, bytes = 0.0; long long n = 0; } kacc[K_NKERN];
    struct KPend { int id; double bytes; cudaEvent_t a, b; };
    std::vector<KPend> kpend;
    std::vector<cudaEvent_t> kfree;
    bool kprof = false;
    cudaEvent_t kev()
    {
        if (kfree.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; }
        cudaEvent_t e = kfree.back(); kfree.pop_back(); return e;
    }
    void kflush()
    {
        cudaStreamSynchronize(s());
        for (const KPend &k : kpend) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, k.a, k.b) == cudaSuccess) { kacc[k.id].ms += ms; kacc[k.id].bytes += k.bytes; kacc[k.id].n++; }
            kfree.push_back(k.a); kfree.push_back(k.b);
        }
        kpend.clear();
    }
    void kbegin(int id, double passes)
    {
        if (kpend.size() >= 2048) kflush();
        KPend k{id, passes * 8.0 * (double)N, kev(), kev()};
        cudaEventRecord(k.a, s());
        kpend.push_back(k);
    }
    void kend() { cudaEventRecord(kpend.back().b, s()); }
    // launch of a vector kernel: kernel id + vector passes (N doubles read or written) for the profile
    template <typename... ExpTypes, typename... ActTypes>
    void LK(int id, double passes, void (*kernel)(ExpTypes...), ActTypes &&... args)
    {
        if (!kprof) {
            launch_pdl(s(), ctx->pdl, blocks, PB_VEC_THREADS, kernel, std::forward<ActTypes>(args)...);
            return;
        }
        kbegin(id, passes);
        launch_pdl(s(), 0, blocks, PB_VEC_THREADS, kernel, std::forward<ActTypes>(args)...);
        kend();
    }

    // ---- thin launch helpers -------------------------------------------------
    void rhs(const double *yin, double *ydot)
    {
        wrap_a.d = const_cast<double *>(yin);
        wrap_b.d = ydot;
        if (prof && prof_used + 2 <= prof_ev.size()) {
            cudaEventRecord(prof_ev[prof_used], s());
            if (pihm_b200_ode(ctx, tn, &wrap_a, &wrap_b) != 0) note_status(cudaErrorLaunchFailure);
            cudaEventRecord(prof_ev[prof_used + 1], s());
            prof_used += 2;
            return;
        }
        if (pihm_b200_ode(ctx, tn, &wrap_a, &wrap_b) != 0) note_status(cudaErrorLaunchFailure);
    }
    // diagnostics (pihm_b200_set_diagnostics): say so BEFORE a kernel overwrites a vector that
    // can be the input of the last RHS call (zn[0], y, ytemp), see common.cuh note_write
    void clobber(const double *p) { pb::note_write(ctx, p); }
    ZnPtrs znp() const { ZnPtrs p; for (int j = 0; j < 6; j++) p.z[j] = zn[j]; return p; }
    void scale_inplace(double c, double *v)
    {
        LK(K_SCALE, 2, k_elementwise<EW_SCALE>, N, c, v, nullptr, v);
        count();
    }
    void copy(const double *src, double *dst)
    {
        clobber(dst);
        LK(K_COPY, 2, k_elementwise<EW_COPY>, N, 0.0, src, nullptr, dst);
        count();
    }
    void axpy(double a, const double *x, double *yv)     // Vaxpy: y += a*x
    {
        LK(K_AXPY, 3, k_linearsum_alias<LS_GENERAL>, N, a, x, 1.0, yv, yv);
        count();
    }
    void launch_ewt()                                    // efun + tolsf norm (cvode.c:1349,1376)
    {
        LK(K_EWT, 2, k_ewt, N, reltol, abstol, zn[0], ewt, R(SC_EWT_MIN, SC_EWT_NRM));
        count();
        red(SC_EWT_MIN, 1, 1);
        red(SC_EWT_NRM);
        ewt_pending = true;
    }

    // ---- CVODE pieces ----------------------------------------------------------
    void cvRescale();
    void cvPredict();
    void cvRestore(double saved_t);
    void cvAdjustOrder(int deltaq);
    void cvAdjustParams();
    void cvSetBDF();
    void cvSet();
    int cvNlsNewton(int nflag);
    int cvNewtonIteration();
    int cvHandleNFlag(int *nflagPtr, double saved_t, int *ncfPtr);
    int cvDoErrorTest(int *nflagPtr, double saved_t, int *nefPtr, double *dsmPtr);
    void cvCompleteStep();
    void cvPrepareNextStep(double dsm);
    void cvSetEta();
    void cvChooseEta();
    void cvBDFStab();
    int cvSLdet();
    int cvStep();
    int spgmrSolve(bool *b_is_zero);
    int getDky(double t, double *dky);
    int solve(double tout, double *yout, double *tret);
};

// cvode.c:2251-2265
void pihm_b200_cvode::cvRescale()
{
    Coef6 f{};
    double factor = eta;
    for (int j = 1; j <= q; j++) { f.c[j] = factor; factor *= eta; }
    // (folding this into the k_predict that always follows was built and measured: the fused variant spills
    // and takes 56 us against 19 + 19 us for the two launches -- PB_FUSE_RESCALE)
#ifdef PB_FUSE_RESCALE
    resc_f = f;
    resc_pending = true;
#else
    LK(K_RESCALE, 2.0 * q, k_rescale, N, q, znp(), f);
    count();
#endif
    h = hscale * eta;
    next_h = h;
    hscale = h;
    nscon = 0;
}

// cvode.c:2277-2288
void pihm_b200_cvode::cvPredict()
{
    tn += h;
    if (tstopset) {
        if ((tn - tstop) * h > 0.0) tn = tstop;
    }
    clobber(zn[0]);
    if (resc_pending) LK(K_PREDICT, 2.0 * q + 2, k_predict<1, true>, N, q, znp(), resc_f);
    else LK(K_PREDICT, 2.0 * q + 1, k_predict<1, false>, N, q, znp(), Coef6{});
    resc_pending = false;
    count();
}

// cvode.c:2882-2890
void pihm_b200_cvode::cvRestore(double saved_t)
{
    tn = saved_t;
    clobber(zn[0]);
    LK(K_PREDICT, 2.0 * q + 1, k_predict<-1, false>, N, q, znp(), Coef6{});
    count();
}

// cvAdjustOrder / cvIncreaseBDF / cvDecreaseBDF, cvode.c:2105-2241
void pihm_b200_cvode::cvAdjustOrder(int deltaq)
{
    if ((q == 2) && (deltaq != 1)) return;
    if (deltaq == 1) {
        double alpha0, alpha1, prod, xi, xiold, hsum, A1;
        for (int i = 0; i <= qmax; i++) l[i] = 0.0;
        l[2] = alpha1 = prod = xiold = 1.0;
        alpha0 = -1.0;
        hsum = hscale;
        if (q > 1) {
            for (int j = 1; j < q; j++) {
                hsum += tau[j + 1];
                xi = hsum / hscale;
                prod *= xi;
                alpha0 -= 1.0 / (j + 1);
                alpha1 += 1.0 / xi;
                for (int i = j + 2; i >= 2; i--) l[i] = l[i] * xiold + l[i - 1];
                xiold = xi;
            }
        }
        A1 = (-alpha0 - alpha1) / prod;
        // zn[L] = A1 * zn[indx_acor]
        LK(K_SCALE, 2, k_elementwise<EW_SCALE>, N, A1, zn[indx_acor], nullptr, zn[L]);
        count();
        for (int j = 2; j <= q; j++) axpy(l[j], zn[L], zn[j]);
    } else if (deltaq == -1) {
        double hsum, xi;
        for (int i = 0; i <= qmax; i++) l[i] = 0.0;
        l[2] = 1.0;
        hsum = 0.0;
        for (int j = 1; j <= q - 2; j++) {
            hsum += tau[j];
            xi = hsum / hscale;
            for (int i = j + 2; i >= 2; i--) l[i] = l[i] * xi + l[i - 1];
        }
        for (int j = 2; j < q; j++) axpy(-l[j], zn[q], zn[j]);
    }
}

// cvode.c:2078-2087
void pihm_b200_cvode::cvAdjustParams()
{
    if (qprime != q) {
        cvAdjustOrder(qprime - q);
        q = qprime;
        L = q + 1;
        qwait = L;
    }
    cvRescale();
}

// cvSetBDF + cvSetTqBDF, cvode.c:2467-2531
void pihm_b200_cvode::cvSetBDF()
{
    double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
    l[0] = l[1] = xi_inv = xistar_inv = 1.0;
    for (int i = 2; i <= q; i++) l[i] = 0.0;
    alpha0 = alpha0_hat = -1.0;
    hsum = h;
    if (q > 1) {
        for (int j = 2; j < q; j++) {
            hsum += tau[j - 1];
            xi_inv = h / hsum;
            alpha0 -= 1.0 / j;
            for (int i = j; i >= 1; i--) l[i] += l[i - 1] * xi_inv;
        }
        alpha0 -= 1.0 / q;
        xistar_inv = -l[1] - alpha0;
        hsum += tau[q - 1];
        xi_inv = h / hsum;
        alpha0_hat = -l[1] - xi_inv;
        for (int i = q; i >= 1; i--) l[i] += l[i - 1] * xistar_inv;
    }
    // cvSetTqBDF
    double A1, A2, A3, A4, A5, A6, C, Cpinv, Cppinv;
    A1 = 1.0 - alpha0_hat + alpha0;
    A2 = 1.0 + q * A1;
    tq[2] = std::fabs(A1 / (alpha0 * A2));
    tq[5] = std::fabs(A2 * xistar_inv / (l[q] * xi_inv));
    if (qwait == 1) {
        if (q > 1) {
            C = xistar_inv / l[q];
            A3 = alpha0 + 1.0 / q;
            A4 = alpha0_hat + xi_inv;
            Cpinv = (1.0 - A4 + A3) / A3;
            tq[1] = std::fabs(C * Cpinv);
        } else {
            tq[1] = 1.0;
        }
        hsum += tau[q];
        xi_inv = h / hsum;
        A5 = alpha0 - (1.0 / (q + 1));
        A6 = alpha0_hat - xi_inv;
        Cppinv = (1.0 - A6 + A5) / A2;
        tq[3] = std::fabs(Cppinv / (xi_inv * (q + 2) * A5));
    }
    tq[4] = nlscoef / tq[2];
}

// cvode.c:2308-2322
void pihm_b200_cvode::cvSet()
{
    cvSetBDF();
    rl1 = 1.0 / l[1];
    gamma = h * rl1;
    if (nst == 0) gammap = gamma;
    gamrat = (nst > 0) ? gamma / gammap : 1.0;
}

// CVSpgmrSolve (cvode_spgmr.c:355-441) after the bnorm test, i.e. SpgmrSolve
// (sundials_spgmr.c:165-433) with x0 = 0, s1 = s2 = ewt, PREC_NONE, MODIFIED_GS,
// max_restarts = 0.  On entry V[0] = ewt*b and h_sc[SC_BSUM] = sum (ewt*b)^2.
// Returns the lsolve value (0 ok, >0 recoverable, <0 fatal); *b_is_zero says the
// solution is x = 0; otherwise yg[0..krydim) / V[] define it (see k_spgmr_final).
int pihm_b200_cvode::spgmrSolve(bool *b_is_zero)
{
    *b_is_zero = false;
    krydim_last = 0;
    const double beta = rsqrt_s(h_sc[SC_BSUM]);
    double r_norm = beta, rho = beta;
    int nli_inc = 0, retval;
    bool converged = false;

    if (r_norm <= delta) {               // sundials_spgmr.c:238-239: x stays 0
        *b_is_zero = true;
        retval = 0;                      // SPGMR_SUCCESS
    } else {
        for (int i = 0; i <= maxl; i++)
            for (int j = 0; j < maxl; j++) Hes[i][j] = 0.0;
        double rotation_product = 1.0;
        double cnorm = 1.0 / r_norm;     // N_VScale(ONE/r_norm, V[0], V[0])
        int krydim = 0;
        for (int lk = 0; lk < maxl; lk++) {
            nli_inc++;
            const int l_plus_1 = lk + 1;
            krydim = l_plus_1;
            // A-tilde V[l]: right scaling, DQ J*v, I - gamma J, left scaling, first MGS dot
            LK(K_KRYLOV_A, 4, k_krylov_a, N, cnorm, V[lk], ewt, vtemp, R(SC_VNRM));
            red(SC_VNRM);
            clobber(ytemp);
            LK(K_KRYLOV_B, 3, k_krylov_b, N, n_global, d_sc, vtemp, y, ytemp);
            count(2);
            rhs(ytemp, V[l_plus_1]);     // Jv = f(tn, y + sig*v)   (cvode_spils.c:687)
            nfes++;
            njtimes++;
            LK(K_KRYLOV_C, 6, k_krylov_c, N, n_global, gamma, d_sc, vtemp, ftemp, ewt,
                                                            V[0], V[l_plus_1], R(SC_VK2, SC_H0));
            count();
            red(SC_VK2);
            red(SC_H0);
            // ModifiedGS (sundials_iterative.c:44-92), i0 = 0 since k <= p
            if (chain_ok) {
                KryPtrs kp{};
                for (int i = 0; i < l_plus_1; i++) kp.v[i] = V[i];
                RedBuf r = rb;
                r.seq = (double)(seq_ctr + 1);           // one ticket per step
                r.gate = d_sc + SC_GATE;
                for (int st = 0; st < l_plus_1; st++)
                    expect((st == l_plus_1 - 1) ? SC_NEW2 : SC_H0 + st + 1, (double)(seq_ctr + 1 + st));
                seq_ctr += l_plus_1;
                long long n_ = N;
                int nsteps = l_plus_1, per = chain_per_thread;
                double *vk = V[l_plus_1];
                void *args[] = {&n_, &nsteps, &kp, &vk, &r, &per};
                if (kprof) kbegin(K_MGS_CHAIN, 2.0 + 2.0 * l_plus_1 - 1.0);
                if (cudaLaunchCooperativeKernel((void *)k_mgs_chain, dim3(blocks), dim3(PB_VEC_THREADS), args,
                                                chain_smem, s()) != cudaSuccess) {
                    set_error(std::string("k_mgs_chain launch: ") + cudaGetErrorString(cudaGetLastError()));
                    return -1;
                }
                if (kprof) kend();
                count();
            } else
            for (int i = 0; i < l_plus_1; i++) {
                const double *vnext = (i + 1 < l_plus_1) ? V[i + 1] : V[l_plus_1];
                const int slot_next = (i + 1 < l_plus_1) ? SC_H0 + i + 1 : SC_NEW2;
                LK(K_MGS_STEP, (vnext == V[l_plus_1]) ? 3 : 4, k_mgs_step, N, d_sc, SC_H0 + i, V[i], vnext,
                                                                V[l_plus_1], slot_next, R(slot_next));
                count();
                red(slot_next);
            }
            sync_spin();
            const double vk_norm = rsqrt_s(h_sc[SC_VK2]);
            for (int i = 0; i < l_plus_1; i++) Hes[i][lk] = h_sc[SC_H0 + i];
            double new_vk_norm = rsqrt_s(h_sc[SC_NEW2]);
            double temp = 1000.0 * vk_norm;
            if ((temp + new_vk_norm) == temp) {
                // re-orthogonalisation branch (sundials_iterative.c:73-88); rare
                double new_norm_2 = 0.0;
                for (int i = 0; i < l_plus_1; i++) {
                    LK(K_DOT, 2, k_reduce<RD_DOT, 0>, 
                        N, V[i], V[l_plus_1], d_part, d_counter, d_sc + SC_TMP, h_sc_map + 2 * SC_TMP);
                    count();
                    red_nccl(SC_TMP);
                    sync();
                    if (!(ctx->nranks > 1)) h_sc[SC_TMP] = h_sc_map[2 * SC_TMP];      // (partitioned: red_nccl brought it)
                    const double new_product = h_sc[SC_TMP];
                    temp = 1000.0 * Hes[i][lk];
                    if ((temp + new_product) == temp) continue;
                    Hes[i][lk] += new_product;
                    axpy(-new_product, V[i], V[l_plus_1]);
                    new_norm_2 += new_product * new_product;
                }
                if (new_norm_2 != 0.0) {
                    const double new_product = new_vk_norm * new_vk_norm - new_norm_2;
                    new_vk_norm = (new_product > 0.0) ? rsqrt_s(new_product) : 0.0;
                }
            }
            Hes[l_plus_1][lk] = new_vk_norm;
            // QRfact(krydim, Hes, givens, job = l)   (sundials_iterative.c:160-248)
            {
                const int n = krydim;
                double c, sg, temp1, temp2, temp3;
                int code = 0;
                if (lk == 0) {
                    for (int k = 0; k < n; k++) {
                        for (int j = 0; j < k - 1; j++) {
                            const int i = 2 * j;
                            temp1 = Hes[j][k]; temp2 = Hes[j + 1][k];
                            c = givens[i]; sg = givens[i + 1];
                            Hes[j][k] = c * temp1 - sg * temp2;
                            Hes[j + 1][k] = sg * temp1 + c * temp2;
                        }
                        temp1 = Hes[k][k]; temp2 = Hes[k + 1][k];
                        if (temp2 == 0.0) { c = 1.0; sg = 0.0; }
                        else if (std::fabs(temp2) >= std::fabs(temp1)) {
                            temp3 = temp1 / temp2;
                            sg = -1.0 / std::sqrt(1.0 + temp3 * temp3);
                            c = -sg * temp3;
                        } else {
                            temp3 = temp2 / temp1;
                            c = 1.0 / std::sqrt(1.0 + temp3 * temp3);
                            sg = -c * temp3;
                        }
                        givens[2 * k] = c; givens[2 * k + 1] = sg;
                        if ((Hes[k][k] = c * temp1 - sg * temp2) == 0.0) code = k + 1;
                    }
                } else {
                    const int nm1 = n - 1;
                    for (int k = 0; k < nm1; k++) {
                        const int i = 2 * k;
                        temp1 = Hes[k][nm1]; temp2 = Hes[k + 1][nm1];
                        c = givens[i]; sg = givens[i + 1];
                        Hes[k][nm1] = c * temp1 - sg * temp2;
                        Hes[k + 1][nm1] = sg * temp1 + c * temp2;
                    }
                    temp1 = Hes[nm1][nm1]; temp2 = Hes[n][nm1];
                    if (temp2 == 0.0) { c = 1.0; sg = 0.0; }
                    else if (std::fabs(temp2) >= std::fabs(temp1)) {
                        temp3 = temp1 / temp2;
                        sg = -1.0 / std::sqrt(1.0 + temp3 * temp3);
                        c = -sg * temp3;
                    } else {
                        temp3 = temp2 / temp1;
                        c = 1.0 / std::sqrt(1.0 + temp3 * temp3);
                        sg = -c * temp3;
                    }
                    givens[2 * nm1] = c; givens[2 * nm1 + 1] = sg;
                    if ((Hes[nm1][nm1] = c * temp1 - sg * temp2) == 0.0) code = n;
                }
                if (code != 0) { nli += nli_inc; ncfl++; return 1; }   // SPGMR_QRFACT_FAIL -> recoverable
            }
            rotation_product *= givens[2 * lk + 1];
            rho = std::fabs(rotation_product * r_norm);
            if (rho <= delta) { converged = true; break; }
            cnorm = 1.0 / Hes[l_plus_1][lk];   // normalisation of V[l+1], applied by the next k_krylov_a
        }
        // QRsol (sundials_iterative.c:258-293)
        yg[0] = r_norm;
        for (int i = 1; i <= krydim; i++) yg[i] = 0.0;
        {
            int code = 0;
            for (int k = 0; k < krydim; k++) {
                const double c = givens[2 * k], sg = givens[2 * k + 1];
                const double temp1 = yg[k], temp2 = yg[k + 1];
                yg[k] = c * temp1 - sg * temp2;
                yg[k + 1] = sg * temp1 + c * temp2;
            }
            for (int k = krydim - 1; k >= 0; k--) {
                if (Hes[k][k] == 0.0) { code = k + 1; break; }
                yg[k] /= Hes[k][k];
                for (int i = 0; i < k; i++) yg[i] -= yg[k] * Hes[i][k];
            }
            if (code != 0) { nli += nli_inc; ncfl++; return -1; }      // SPGMR_QRSOL_FAIL
        }
        krydim_last = krydim;
        if (converged) retval = 0;               // SPGMR_SUCCESS
        else if (rho < beta) retval = 1;         // SPGMR_RES_REDUCED
        else retval = 2;                         // SPGMR_CONV_FAIL (x stays 0)
    }
    nli += nli_inc;
    if (retval != 0) ncfl++;
    if (retval == 0) return 0;
    if (retval == 1) return (mnewt == 0) ? 0 : 1;
    return 1;
}

// cvode.c:2729-2800 with lsolve = CVSpgmrSolve inlined
int pihm_b200_cvode::cvNewtonIteration()
{
    int m = 0;
    double del = 0.0, delp = 0.0, dcon;
    mnewt = 0;
    for (;;) {
        clobber(y);
        if (m == 0)
            LK(K_NEWTON_RES, 8, k_newton_res<true>, N, rl1, gamma, zn[0], zn[1], ftemp, ewt,
                                                                   acor, y, tempv, V[0], R(SC_BSUM));
        else
            LK(K_NEWTON_RES, 6, k_newton_res<false>, N, rl1, gamma, zn[0], zn[1], ftemp, ewt,
                                                                    acor, y, tempv, V[0], R(SC_BSUM));
        count();
        red(SC_BSUM);
        sync_spin();
        if (ewt_pending) {               // deferred checks of cvode.c:1349-1388
            ewt_pending = false;
            if (h_sc[SC_EWT_MIN] <= 0.0) return CV_ILL_INPUT;
            if (uround * wrms(SC_EWT_NRM) > 1.0) return CV_TOO_MUCH_ACC;
        }
        // CVSpgmrSolve (cvode_spgmr.c:355-441)
        deltar = eplifac * tq[4];
        const double bnorm = wrms(SC_BSUM);
        int retval;
        clobber(y);                      // k_newton_update / k_spgmr_final: y = zn[0] + acor
        if (bnorm <= deltar) {
            // x = b (first iteration) or x = 0
            if (mnewt > 0)
                LK(K_NEWTON_UPDATE, 5, k_newton_update<true>, N, tempv, ewt, zn[0], acor, y, R(SC_DEL, SC_ACNRM));
            else
                LK(K_NEWTON_UPDATE, 6, k_newton_update<false>, N, tempv, ewt, zn[0], acor, y, R(SC_DEL, SC_ACNRM));
            count();
            retval = 0;
        } else {
            delta = deltar * sqrtN;
            bool zero;
            retval = spgmrSolve(&zero);
            if (retval == 0) {
                if (zero || krydim_last == 0) {
                    LK(K_NEWTON_UPDATE, 5, k_newton_update<true>, N, tempv, ewt, zn[0], acor, y, R(SC_DEL, SC_ACNRM));
                } else {
                    KryPtrs kp{};
                    Coef6 c{};
                    for (int k = 0; k < krydim_last; k++) { kp.v[k] = V[k]; c.c[k] = yg[k]; }
                    LK(K_SPGMR_FINAL, krydim_last + 5.0, k_spgmr_final, N, krydim_last, kp, c, ewt, zn[0], acor, y, R(SC_DEL, SC_ACNRM));
                }
                count();
            }
        }
        nni++;
        if (retval < 0) return CV_LSOLVE_FAIL;
        if (retval > 0) return CONV_FAIL;      // setupNonNull == FALSE (cvode_spgmr.c:262)

        red(SC_DEL, 2);                  // SC_DEL, SC_ACNRM
        sync_spin();
        del = wrms(SC_DEL);
        if (m > 0) crate = std::max(CRDOWN * crate, del / delp);
        dcon = del * std::min(1.0, crate) / tq[4];
        if (dcon <= 1.0) {
            if (m == 0) {
                acnrm = del;
            } else {
                acnrm = wrms(SC_ACNRM);  // N_VWrmsNorm(acor, ewt): summed by the update kernel that wrote acor
            }
            return CV_SUCCESS;
        }
        mnewt = ++m;
        if ((m == maxcor) || ((m >= 2) && (del > RDIV * delp))) return CONV_FAIL;
        delp = del;
        rhs(y, ftemp);
        nfe++;
    }
}

// cvode.c:2651-2714 (setupNonNull == FALSE: no lsetup, crate = 1)
int pihm_b200_cvode::cvNlsNewton(int nflag)
{
    (void)nflag;
    crate = 1.0;
    rhs(zn[0], ftemp);
    nfe++;
    return cvNewtonIteration();
}

// cvode.c:2834-2871
int pihm_b200_cvode::cvHandleNFlag(int *nflagPtr, double saved_t, int *ncfPtr)
{
    const int nflag = *nflagPtr;
    if (nflag == CV_SUCCESS) return DO_ERROR_TEST;
    ncfn++;
    cvRestore(saved_t);
    if (nflag == CV_LSOLVE_FAIL) return CV_LSOLVE_FAIL;
    if (nflag == CV_RHSFUNC_FAIL) return CV_RHSFUNC_FAIL;
    if (nflag == CV_ILL_INPUT || nflag == CV_TOO_MUCH_ACC) return nflag;
    (*ncfPtr)++;
    etamax = 1.0;
    if ((std::fabs(h) <= hmin * ONEPSM) || (*ncfPtr == maxncf)) {
        if (nflag == CONV_FAIL) return CV_CONV_FAILURE;
    }
    eta = std::max(ETACF, hmin / std::fabs(h));
    *nflagPtr = PREV_CONV_FAIL;
    cvRescale();
    return PREDICT_AGAIN;
}

// cvode.c:2915-2976
int pihm_b200_cvode::cvDoErrorTest(int *nflagPtr, double saved_t, int *nefPtr, double *dsmPtr)
{
    const double dsm = acnrm * tq[2];
    *dsmPtr = dsm;
    if (dsm <= 1.0) return CV_SUCCESS;
    (*nefPtr)++;
    netf++;
    *nflagPtr = PREV_ERR_FAIL;
    cvRestore(saved_t);
    if ((std::fabs(h) <= hmin * ONEPSM) || (*nefPtr == maxnef)) return CV_ERR_FAILURE;
    etamax = 1.0;
    if (*nefPtr <= MXNEF1) {
        eta = 1.0 / (rpowerR(BIAS2 * dsm, 1.0 / L) + ADDON);
        eta = std::max(ETAMIN, std::max(eta, hmin / std::fabs(h)));
        if (*nefPtr >= SMALL_NEF) eta = std::min(eta, ETAMXF);
        cvRescale();
        return TRY_AGAIN;
    }
    if (q > 1) {
        eta = std::max(ETAMIN, hmin / std::fabs(h));
        cvAdjustOrder(-1);
        L = q;
        q--;
        qwait = L;
        cvRescale();
        return TRY_AGAIN;
    }
    eta = std::max(ETAMIN, hmin / std::fabs(h));
    h *= eta;
    next_h = h;
    hscale = h;
    qwait = LONG_WAIT;
    nscon = 0;
    rhs(zn[0], tempv);
    nfe++;
    LK(K_SCALE, 2, k_elementwise<EW_SCALE>, N, h, tempv, nullptr, zn[1]);
    count();
    return TRY_AGAIN;
}

// cvode.c:2996-3018
void pihm_b200_cvode::cvCompleteStep()
{
    nst++;
    nscon++;
    hu = h;
    qu = q;
    for (int i = q; i >= 2; i--) tau[i] = tau[i - 1];
    if ((q == 1) && (nst > 1)) tau[2] = tau[1];
    tau[1] = h;
    Coef6 lc{};
    for (int j = 0; j <= q; j++) lc.c[j] = l[j];
    qwait--;
    double *save = nullptr;
    if ((qwait == 1) && (q != qmax)) {
        save = zn[qmax];
        saved_tq5 = tq[5];
        indx_acor = qmax;
    }
    clobber(zn[0]);
    if (!fold) {
        LK(K_COMPLETE, 2.0 * (q + 1) + 1 + (save ? 1 : 0), k_complete, N, q, znp(), lc, acor, save);
        count();
        return;
    }
    // what cvPrepareNextStep (cvode.c:3029-3059) and cvBDFStab (:3249-3297) will ask for, decided here with the
    // conditions they apply (etamax, qwait after its decrement above, q, saved_tq5, sldeton)
    const bool need_eta = (etamax != 1.0) && (qwait == 0);
    fold_m1 = need_eta && (q > 1);
    fold_p1 = need_eta && (q != qmax) && (saved_tq5 != 0.0);
    fold_stab = sldeton && (q >= 3);
    CompleteNorms cn{};
    cn.reltol = reltol; cn.abstol = abstol;
    cn.cquot = fold_p1 ? (tq[5] / saved_tq5) * rpowerI(h / tau[2], L) : 0.0;
    cn.ewt_next = ewt_next; cn.ewt = ewt; cn.znmax = zn[qmax];
    cn.do_zq = (fold_m1 || fold_stab) ? 1 : 0;
    cn.do_zqm1 = fold_stab ? 1 : 0;
    cn.do_p1 = fold_p1 ? 1 : 0;
    RedBuf &r = R(SC_EWT_MIN, SC_EWT_NRM);
    if (cn.do_zq) expect(SC_STAB1, r.seq);
    if (cn.do_zqm1) expect(SC_STAB2, r.seq);
    if (cn.do_p1) expect(SC_ETA_P1, r.seq);
    LK(K_COMPLETE, 2.0 * (q + 1) + 2 + (save ? 1 : 0) + ((cn.do_zq || cn.do_p1) ? 1 : 0) + (cn.do_p1 ? 1 : 0),
       k_complete_norms, N, q, znp(), lc, acor, save, cn, r);
    count();
    red(SC_EWT_MIN, 1, 1);
    red(SC_EWT_NRM);
    if (cn.do_zq) red(SC_STAB1, cn.do_zqm1 ? 2 : 1);
    if (cn.do_p1) red(SC_ETA_P1);
    ewt_ready = true;
}

// cvode.c:3067-3082
void pihm_b200_cvode::cvSetEta()
{
    if (eta < THRESH) {
        eta = 1.0;
        hprime = h;
    } else {
        eta = std::min(eta, etamax);
        eta /= std::max(1.0, std::fabs(h) * hmax_inv * eta);
        hprime = h * eta;
        if (qprime < q) nscon = 0;
    }
}

// cvode.c:3138-3176
void pihm_b200_cvode::cvChooseEta()
{
    const double etam = std::max(etaqm1, std::max(etaq, etaqp1));
    if (etam < THRESH) {
        eta = 1.0;
        qprime = q;
        return;
    }
    if (etam == etaq) {
        eta = etaq;
        qprime = q;
    } else if (etam == etaqm1) {
        eta = etaqm1;
        qprime = q - 1;
    } else {
        eta = etaqp1;
        qprime = q + 1;
        copy(acor, zn[qmax]);
    }
}

// cvode.c:3029-3059 with cvComputeEtaqm1/qp1 (:3090-3124) fused into one kernel
void pihm_b200_cvode::cvPrepareNextStep(double dsm)
{
    if (etamax == 1.0) {
        qwait = std::max(qwait, 2);
        qprime = q;
        hprime = h;
        eta = 1.0;
        return;
    }
    etaq = 1.0 / (rpowerR(BIAS2 * dsm, 1.0 / L) + ADDON);
    if (qwait != 0) {
        eta = etaq;
        qprime = q;
        cvSetEta();
        return;
    }
    qwait = 2;
    const bool do_m1 = (q > 1);
    const bool do_p1 = (q != qmax) && (saved_tq5 != 0.0);
    double cquot = 0.0;
    if (do_p1) cquot = (tq[5] / saved_tq5) * rpowerI(h / tau[2], L);
    etaqm1 = 0.0;
    etaqp1 = 0.0;
    if (fold && (do_m1 || do_p1)) {
        // produced by k_complete_norms (cvCompleteStep decided with the same conditions)
        sync_spin();
        h_sc[SC_ETA_M1] = h_sc[SC_STAB1];
        if (do_m1) {
            const double ddn = wrms(SC_ETA_M1) * tq[1];
            etaqm1 = 1.0 / (rpowerR(BIAS1 * ddn, 1.0 / q) + ADDON);
        }
        if (do_p1) {
            const double dup = wrms(SC_ETA_P1) * tq[3];
            etaqp1 = 1.0 / (rpowerR(BIAS3 * dup, 1.0 / (L + 1)) + ADDON);
        }
    } else if (do_m1 || do_p1) {
        LK(K_ETA, 1.0 + (do_m1 ? 1 : 0) + (do_p1 ? 2 : 0), k_eta, N, do_m1, do_p1, cquot, zn[q], zn[qmax], acor, ewt, R(SC_ETA_M1, SC_ETA_P1));
        count();
        red(SC_ETA_M1, 2);
        sync_spin();
        if (do_m1) {
            const double ddn = wrms(SC_ETA_M1) * tq[1];
            etaqm1 = 1.0 / (rpowerR(BIAS1 * ddn, 1.0 / q) + ADDON);
        }
        if (do_p1) {
            const double dup = wrms(SC_ETA_P1) * tq[3];
            etaqp1 = 1.0 / (rpowerR(BIAS3 * dup, 1.0 / (L + 1)) + ADDON);
        }
    }
    cvChooseEta();
    cvSetEta();
}

// cvode.c:3249-3297
void pihm_b200_cvode::cvBDFStab()
{
    if (q >= 3) {
        for (int k = 1; k <= 3; k++)
            for (int i = 5; i >= 2; i--) ssdat[i][k] = ssdat[i - 1][k];
        int factorial = 1;
        for (int i = 1; i <= q - 1; i++) factorial *= i;
        if (!fold) {
            LK(K_WSQ, 3, k_wsq, N, zn[q], zn[q - 1], ewt, SC_STAB1, SC_STAB2, R(SC_STAB1, SC_STAB2));
            count();
            red(SC_STAB1, 2);
        }
        sync_spin();                     // (fold: the two sums come from k_complete_norms)
        const double sq = factorial * q * (q + 1) * acnrm / std::max(tq[5], TINY);
        const double sqm1 = factorial * q * wrms(SC_STAB1);
        const double sqm2 = factorial * wrms(SC_STAB2);
        ssdat[1][1] = sqm2 * sqm2;
        ssdat[1][2] = sqm1 * sqm1;
        ssdat[1][3] = sq * sq;
    }
    if (qprime >= q) {
        if ((q >= 3) && (nscon >= q + 5)) {
            const int ldflag = cvSLdet();
            if (ldflag > 3) {
                qprime = q - 1;
                eta = etaqm1;
                eta = std::min(eta, etamax);
                eta = eta / std::max(1.0, std::fabs(h) * hmax_inv * eta);
                hprime = h * eta;
                nor = nor + 1;
            }
        }
    } else {
        nscon = 0;
    }
}

// cvode.c:3336-3604 (STALD); scalar code on ssdat[][]
int pihm_b200_cvode::cvSLdet()
{
    int kmin = 0, kflag = 0;
    double rat[5][4], rav[4], qkr[4], sigsq[4], smax[4], ssmax[4];
    double drr[4], rrc[4], sqmx[4], qjk[4][4], vrat[5], qc[6][4], qco[6][4];
    double rr, smink, smaxk, sumrat, sumrsq, vmin, vmax, drrmax, adrr;
    double tem, sqmax, saqk, qp, sv, sqmaxk, saqj, sqmin = 0.0;
    double rsa, rsb, rsc, rsd, rd1a, rd1b, rd1c, rd2a, rd2b, rd3a, cest1, corr1;
    double ratp, ratm, qfac1, qfac2, bb, rrb;
    const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;

    rr = 0.0;
    for (int k = 1; k <= 3; k++) {
        smink = ssdat[1][k];
        smaxk = 0.0;
        for (int i = 1; i <= 5; i++) {
            smink = std::min(smink, ssdat[i][k]);
            smaxk = std::max(smaxk, ssdat[i][k]);
        }
        if (smink < TINY * smaxk) return -1;
        smax[k] = smaxk;
        ssmax[k] = smaxk * smaxk;
        sumrat = 0.0;
        sumrsq = 0.0;
        for (int i = 1; i <= 4; i++) {
            rat[i][k] = ssdat[i][k] / ssdat[i + 1][k];
            sumrat = sumrat + rat[i][k];
            sumrsq = sumrsq + rat[i][k] * rat[i][k];
        }
        rav[k] = 0.25 * sumrat;
        vrat[k] = std::fabs(0.25 * sumrsq - rav[k] * rav[k]);
        qc[5][k] = ssdat[1][k] * ssdat[3][k] - ssdat[2][k] * ssdat[2][k];
        qc[4][k] = ssdat[2][k] * ssdat[3][k] - ssdat[1][k] * ssdat[4][k];
        qc[3][k] = 0.0;
        qc[2][k] = ssdat[2][k] * ssdat[5][k] - ssdat[3][k] * ssdat[4][k];
        qc[1][k] = ssdat[4][k] * ssdat[4][k] - ssdat[3][k] * ssdat[5][k];
        for (int i = 1; i <= 5; i++) qco[i][k] = qc[i][k];
    }
    vmin = std::min(vrat[1], std::min(vrat[2], vrat[3]));
    vmax = std::max(vrat[1], std::max(vrat[2], vrat[3]));
    if (vmin < vrrtol * vrrtol) {
        if (vmax > vrrt2 * vrrt2) return -2;
        rr = (rav[1] + rav[2] + rav[3]) / 3.0;
        drrmax = 0.0;
        for (int k = 1; k <= 3; k++) {
            adrr = std::fabs(rav[k] - rr);
            drrmax = std::max(drrmax, adrr);
        }
        if (drrmax > vrrt2) kflag = -3;
        kflag = 1;
    } else {
        if (std::fabs(qco[1][1]) < TINY * ssmax[1]) return -4;
        tem = qco[1][2] / qco[1][1];
        for (int i = 2; i <= 5; i++) qco[i][2] = qco[i][2] - tem * qco[i][1];
        qco[1][2] = 0.0;
        tem = qco[1][3] / qco[1][1];
        for (int i = 2; i <= 5; i++) qco[i][3] = qco[i][3] - tem * qco[i][1];
        qco[1][3] = 0.0;
        if (std::fabs(qco[2][2]) < TINY * ssmax[2]) return -4;
        tem = qco[2][3] / qco[2][2];
        for (int i = 3; i <= 5; i++) qco[i][3] = qco[i][3] - tem * qco[i][2];
        if (std::fabs(qco[4][3]) < TINY * ssmax[3]) return -4;
        rr = -qco[5][3] / qco[4][3];
        if (rr < TINY || rr > HUNDRED) return -5;
        for (int k = 1; k <= 3; k++)
            qkr[k] = qc[5][k] + rr * (qc[4][k] + rr * rr * (qc[2][k] + rr * qc[1][k]));
        sqmax = 0.0;
        for (int k = 1; k <= 3; k++) {
            saqk = std::fabs(qkr[k]) / ssmax[k];
            if (saqk > sqmax) sqmax = saqk;
        }
        if (sqmax < sqtol) {
            kflag = 2;
        } else {
            for (int it = 1; it <= 3; it++) {
                for (int k = 1; k <= 3; k++) {
                    qp = qc[4][k] + rr * rr * (3.0 * qc[2][k] + rr * 4.0 * qc[1][k]);
                    drr[k] = 0.0;
                    if (std::fabs(qp) > TINY * ssmax[k]) drr[k] = -qkr[k] / qp;
                    rrc[k] = rr + drr[k];
                }
                for (int k = 1; k <= 3; k++) {
                    sv = rrc[k];
                    sqmaxk = 0.0;
                    for (int j = 1; j <= 3; j++) {
                        qjk[j][k] = qc[5][j] + sv * (qc[4][j] + sv * sv * (qc[2][j] + sv * qc[1][j]));
                        saqj = std::fabs(qjk[j][k]) / ssmax[j];
                        if (saqj > sqmaxk) sqmaxk = saqj;
                    }
                    sqmx[k] = sqmaxk;
                }
                sqmin = sqmx[1] + 1.0;
                for (int k = 1; k <= 3; k++) {
                    if (sqmx[k] < sqmin) {
                        kmin = k;
                        sqmin = sqmx[k];
                    }
                }
                rr = rrc[kmin];
                if (sqmin < sqtol) {
                    kflag = 3;
                    break;
                } else {
                    for (int j = 1; j <= 3; j++) qkr[j] = qjk[j][kmin];
                }
            }
            if (sqmin > sqtol) return -6;
        }
    }
    for (int k = 1; k <= 3; k++) {
        rsa = ssdat[1][k];
        rsb = ssdat[2][k] * rr;
        rsc = ssdat[3][k] * rr * rr;
        rsd = ssdat[4][k] * rr * rr * rr;
        rd1a = rsa - rsb;
        rd1b = rsb - rsc;
        rd1c = rsc - rsd;
        rd2a = rd1a - rd1b;
        rd2b = rd1b - rd1c;
        rd3a = rd2a - rd2b;
        if (std::fabs(rd1b) < TINY * smax[k]) return -7;
        cest1 = -rd3a / rd1b;
        if (cest1 < TINY || cest1 > 4.0) return -7;
        corr1 = (rd2b / cest1) / (rr * rr);
        sigsq[k] = ssdat[3][k] + corr1;
    }
    if (sigsq[2] < TINY) return -8;
    ratp = sigsq[3] / sigsq[2];
    ratm = sigsq[1] / sigsq[2];
    qfac1 = 0.25 * (q * q - 1.0);
    qfac2 = 2.0 / (q - 1.0);
    bb = ratp * ratm - 1.0 - qfac1 * ratp;
    tem = 1.0 - qfac2 * bb;
    if (std::fabs(tem) < TINY) return -8;
    rrb = 1.0 / tem;
    if (std::fabs(rrb - rr) > rrtol) return -9;
    if (rr > rrcut) {
        if (kflag == 1) kflag = 4;
        if (kflag == 2) kflag = 5;
        if (kflag == 3) kflag = 6;
    }
    return kflag;
}

// cvode.c:2005-2066
int pihm_b200_cvode::cvStep()
{
    const double saved_t = tn;
    double dsm = 0.0;
    int ncf = 0, nef = 0, nflag = FIRST_CALL, kflag, eflag;

    if ((nst > 0) && (hprime != h)) cvAdjustParams();
    for (;;) {
        cvPredict();
        cvSet();
        nflag = cvNlsNewton(nflag);
        kflag = cvHandleNFlag(&nflag, saved_t, &ncf);
        if (kflag == PREDICT_AGAIN) continue;
        if (kflag != DO_ERROR_TEST) return kflag;
        eflag = cvDoErrorTest(&nflag, saved_t, &nef, &dsm);
        if (eflag == TRY_AGAIN) continue;
        if (eflag != CV_SUCCESS) return eflag;
        break;
    }
    cvCompleteStep();
    cvPrepareNextStep(dsm);
    if (sldeton) cvBDFStab();
    etamax = (nst <= SMALL_NST) ? ETAMX2 : ETAMX3;
    // cvode.c:2063 rescales acor to the estimated local error vector; nothing in
    // the PIHM flow reads it (acor is zeroed at the next Newton start), so the
    // two vector passes are not spent.
    return CV_SUCCESS;
}

// CVodeGetDky, k = 0 (cvode.c:1509-1562)
int pihm_b200_cvode::getDky(double t, double *dky)
{
    double tfuzz = FUZZ_FACTOR * uround * (std::fabs(tn) + std::fabs(hu));
    if (hu < 0.0) tfuzz = -tfuzz;
    const double tp = tn - hu - tfuzz;
    const double tn1 = tn + tfuzz;
    if ((t - tp) * (t - tn1) > 0.0) return CV_BAD_T;
    const double sv = (t - tn) / h;
    clobber(dky);
    LK(K_DKY, q + 2.0, k_dky, N, q, sv, znp(), dky);
    count();
    return CV_SUCCESS;
}

// CVode(..., CV_NORMAL) with tstop set (cvode.c:1074-1489)
int pihm_b200_cvode::solve(double tout, double *yout, double *tret)
{
    int istate = CV_SUCCESS;
    y = yout;
    // CVodeSetStopTime (cvode_io.c:356-383)
    if (nst > 0 && (tout - tn) * h < 0.0) { set_error("SolveCVode: tstop behind current t"); return CV_ILL_INPUT; }
    tstop = tout;
    tstopset = true;

    if (nst == 0) {
        tretlast = *tret = tn;
        // cvInitialSetup: ewt, CVSpgmrInit counters (cvode.c:1730-1770, cvode_spgmr.c:245-275)
        launch_ewt();
        nli = ncfl = nfes = njtimes = 0;
        rhs(zn[0], zn[1]);
        nfe++;
        sync();
        ewt_pending = false;
        if (h_sc[SC_EWT_MIN] <= 0.0) { set_error("initial ewt has a non-positive component"); return CV_ILL_INPUT; }
        if ((tstop - tn) * (tout - tn) <= 0.0) { set_error("SolveCVode: tstop not beyond t0"); return CV_ILL_INPUT; }
        h = hin;
        if ((h != 0.0) && ((tout - tn) * h < 0.0)) return CV_ILL_INPUT;
        if (h == 0.0) { set_error("INIT_SOLVER_STEP must be non-zero (cvHin is not on the PIHM path)"); return CV_ILL_INPUT; }
        const double rh = std::fabs(h) * hmax_inv;
        if (rh > 1.0) h /= rh;
        if (std::fabs(h) < hmin) h *= hmin / std::fabs(h);
        if ((tn + h - tstop) * h > 0.0) h = (tstop - tn) * (1.0 - 4.0 * uround);
        hscale = h;
        h0u = h;
        hprime = h;
        scale_inplace(h, zn[1]);
    }

    if (nst > 0) {
        double troundoff = FUZZ_FACTOR * uround * (std::fabs(tn) + std::fabs(h));
        if ((tn - tout) * h >= 0.0) {
            tretlast = *tret = tout;
            if (getDky(tout, yout) != CV_SUCCESS) return CV_ILL_INPUT;
            sync();
            return CV_SUCCESS;
        }
        if (std::fabs(tn - tstop) <= troundoff) {
            if (getDky(tstop, yout) != CV_SUCCESS) return CV_ILL_INPUT;
            tretlast = *tret = tstop;
            tstopset = false;
            sync();
            return CV_TSTOP_RETURN;
        }
        if ((tn + hprime - tstop) * h > 0.0) {
            hprime = (tstop - tn) * (1.0 - 4.0 * uround);
            eta = hprime / h;
        }
    }

    long long nstloc = 0;
    for (;;) {
        next_h = h;
        next_q = q;
        if (nst > 0) {                   // efun; checked at the first sync inside the step
            if (fold && ewt_ready) { std::swap(ewt, ewt_next); ewt_pending = true; }    // written by k_complete_norms
            else launch_ewt();
            ewt_ready = false;
        }
        if ((mxstep > 0) && (nstloc >= mxstep)) {
            set_error("mxstep steps taken before reaching tout");
            istate = CV_TOO_MUCH_WORK;
            tretlast = *tret = tn;
            copy(zn[0], yout);
            break;
        }
        if (nst == 0) {
            tolsf = uround * wrms(SC_EWT_NRM);
            if (tolsf > 1.0) { istate = CV_TOO_MUCH_ACC; tretlast = *tret = tn; copy(zn[0], yout); break; }
            tolsf = 1.0;
        }
        if (tn + h == tn) nhnil++;

        int kflag = cvStep();
        if (failed()) {
            set_error(std::string("device error inside the integrator: ") + cudaGetErrorString(dev_error));
            dev_error = cudaSuccess;
            tretlast = *tret = tn;
            return CV_RHSFUNC_FAIL;
        }
        if (kflag != CV_SUCCESS) {
            istate = kflag;
            set_error("cvStep failed with flag " + std::to_string(kflag) + " at t = " + std::to_string(tn) +
                      ", h = " + std::to_string(h));
            tretlast = *tret = tn;
            copy(zn[0], yout);
            break;
        }
        nstloc++;

        if ((tn - tout) * h >= 0.0) {
            istate = CV_SUCCESS;
            tretlast = *tret = tout;
            getDky(tout, yout);
            next_q = qprime;
            next_h = hprime;
            break;
        }
        {
            const double troundoff = FUZZ_FACTOR * uround * (std::fabs(tn) + std::fabs(h));
            if (std::fabs(tn - tstop) <= troundoff) {
                getDky(tstop, yout);
                tretlast = *tret = tstop;
                tstopset = false;
                istate = CV_TSTOP_RETURN;
                break;
            }
            if ((tn + hprime - tstop) * h > 0.0) {
                hprime = (tstop - tn) * (1.0 - 4.0 * uround);
                eta = hprime / h;
            }
        }
    }
    sync();
    return istate;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

pihm_b200_cvode *pihm_b200_cvode_create(pihm_b200_ctx *ctx)
{
    if (!ctx) { set_error("cvode_create: null context"); return nullptr; }
    pihm_b200_cvode *cv = new pihm_b200_cvode();
    cv->ctx = ctx;
    cv->N = ctx->nsv;
    cv->n_global = (double)ctx->nsv_global;      // N of the WRMS norms: all ranks' unknowns
    cv->blocks = (int)std::max<long long>(1, std::min<long long>((cv->N + PB_VEC_THREADS - 1) / PB_VEC_THREADS,
                                                                  ctx->red_blocks));
    const size_t bytes = sizeof(double) * ((size_t)std::max<long long>(cv->N, 1) + PB_VEC_PAD);
    double **all[] = {&cv->zn[0], &cv->zn[1], &cv->zn[2], &cv->zn[3], &cv->zn[4], &cv->zn[5], &cv->ewt,
                      &cv->acor, &cv->tempv, &cv->ftemp, &cv->V[0], &cv->V[1], &cv->V[2], &cv->V[3],
                      &cv->V[4], &cv->V[5], &cv->vtemp, &cv->ytemp, &cv->ewt_next};
    bool ok = true;
    for (double **p : all) {
        if (cudaMalloc((void **)p, bytes) != cudaSuccess) { ok = false; break; }
        cudaMemsetAsync(*p, 0, bytes, ctx->s());
    }
    ok = ok && cudaMalloc((void **)&cv->d_part, sizeof(double) * SC_COUNT * ctx->red_blocks) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&cv->d_sc, sizeof(double) * SC_COUNT) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&cv->d_counter, sizeof(unsigned int) * 4) == cudaSuccess;
    ok = ok && cudaHostAlloc((void **)&cv->h_sc_map, sizeof(double) * 2 * SC_COUNT, cudaHostAllocMapped) == cudaSuccess;
    ok = ok && cudaHostAlloc((void **)&cv->h_sc, sizeof(double) * SC_COUNT, cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        set_error(std::string("cvode_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        pihm_b200_cvode_destroy(cv);
        return nullptr;
    }
    cudaMemsetAsync(cv->d_sc, 0, sizeof(double) * SC_COUNT, ctx->s());
    cudaMemsetAsync(cv->d_counter, 0, sizeof(unsigned int) * 4, ctx->s());
    std::memset(cv->h_sc, 0, sizeof(double) * SC_COUNT);
    std::memset(cv->h_sc_map, 0, sizeof(double) * 2 * SC_COUNT);
    cv->rb.part = cv->d_part;
    cv->rb.counter = cv->d_counter;
    cv->rb.sc = cv->d_sc;
    cv->rb.hsc = cv->h_sc_map;
    if (const char *e = getenv("PIHM_B200_PROFILE")) {
        cv->prof = atoi(e) != 0;
        if (cv->prof) {
            cv->prof_ev.resize(16384);
            for (cudaEvent_t &ev : cv->prof_ev) cudaEventCreate(&ev);
        }
    }
    if (const char *e = getenv("PIHM_B200_FOLD")) cv->fold = atoi(e) != 0;
    cv->rb.peer = nullptr;
    cv->rb.nranks = ctx->nranks;
    cv->rb.rank = ctx->rank;
    if (ctx->nranks > 1 && ctx->nranks <= PB_MAX_RANKS && !(getenv("PIHM_B200_NO_P2P") && atoi(getenv("PIHM_B200_NO_P2P")))) {
        // exchange buffers of all ranks mapped into each other (collective; all or nobody)
        bool got = cudaMalloc((void **)&cv->d_xbuf, sizeof(double) * PB_XB_DOUBLES) == cudaSuccess &&
                   cudaMalloc((void **)&cv->d_peer, sizeof(double *) * PB_MAX_RANKS) == cudaSuccess;
        if (got) {
            cudaMemsetAsync(cv->d_xbuf, 0, sizeof(double) * PB_XB_DOUBLES, ctx->s());
            cudaStreamSynchronize(ctx->s());
        }
        if (ctx->lgroup) {
            // ranks of one process: register with the group and refresh every registered table
            // (complete once the integrators of all ranks exist -- before the first solve)
            pb::LocalGroup &g = *ctx->lgroup;
            if (got) {
                g.xbuf[ctx->rank] = cv->d_xbuf;
                g.peer_tab[ctx->rank] = cv->d_peer;
                for (int r = 0; r < g.n; r++)
                    if (g.peer_tab[r]) cudaMemcpy(g.peer_tab[r], g.xbuf, sizeof(double *) * PB_MAX_RANKS, cudaMemcpyHostToDevice);
                cv->rb.peer = cv->d_peer;
                cv->p2p = true;
            }
        } else if (comm_share_buffer(ctx, got ? (void *)cv->d_xbuf : nullptr, cv->h_peer) == 0 && got) {
            cudaMemcpy(cv->d_peer, cv->h_peer, sizeof(double *) * PB_MAX_RANKS, cudaMemcpyHostToDevice);
            cv->rb.peer = cv->d_peer;
            cv->p2p = true;
        }
        if (!cv->p2p && (ctx->lgroup || ctx->halo_p2p)) {
            // the halo runs over peer memory, so the scalars must too (or the caller asked for NCCL for both)
            set_error("cvode_create: the exchange buffers of the in-kernel all-reduce could not be shared");
            pihm_b200_cvode_destroy(cv);
            return nullptr;
        }
    }
    // (partitioned run without peer memory: the kernels' rank-local values go to the mapped pairs and are
    // ignored; the host reads h_sc, which sync() fills from d_sc after the NCCL all-reduces)
    cv->rb.max_blocks = ctx->red_blocks;
    {
        // k_mgs_chain: every CTA must be resident and a thread's elements of one vector must fit
        // its share of the shared memory; reductions must complete inside the kernel (single GPU,
        // or the peer-memory exchange).  Opt-in (PIHM_B200_MGS_CHAIN=1): bit-identical to the k_mgs_step
        // launches and measured neutral at 1M triangles on a B200 (6.44 vs 6.43 ms per model step --
        // the vectors of the chain already sit in the 126 MB L2 between the launches).
        const char *e = getenv("PIHM_B200_MGS_CHAIN");
        const bool want = e && atoi(e) != 0;
        const long long per = (cv->N + (long long)cv->blocks * PB_VEC_THREADS - 1) / ((long long)cv->blocks * PB_VEC_THREADS);
        const size_t smem = sizeof(double) * PB_VEC_THREADS * (size_t)std::max<long long>(per, 1);
        int dev_smem = 0, coop = 0, nsm = 0, occ = 0;
        cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
        if (want && coop && cv->N > 0 && smem <= (size_t)dev_smem && (ctx->nranks == 1 || cv->p2p) &&
            cudaFuncSetAttribute(k_mgs_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mgs_chain, PB_VEC_THREADS, smem) == cudaSuccess &&
            (long long)occ * nsm >= cv->blocks) {
            cv->chain_ok = true;
            cv->chain_per_thread = (int)per;
            cv->chain_smem = smem;
        }
        cudaGetLastError();
    }
    cv->wrap_a.ctx = cv->wrap_b.ctx = ctx;
    cv->wrap_a.n = cv->wrap_b.n = cv->N;
    cv->wrap_a.owns = cv->wrap_b.owns = false;
    return cv;
}

void pihm_b200_cvode_destroy(pihm_b200_cvode *cv)
{
    if (!cv) return;
    pb::note_free(cv->ctx, cv->zn[0]);  // the vectors go away: keep the last RHS input if it is one of them
    pb::note_free(cv->ctx, cv->ytemp);
    cudaStreamSynchronize(cv->ctx->s());
    if (cv->prof && getenv("PIHM_B200_PROFILE")) {
        double rhs_ms = 0.0;
        for (size_t i = 0; i + 1 < cv->prof_used; i += 2) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, cv->prof_ev[i], cv->prof_ev[i + 1]);
            rhs_ms += ms;
        }
        fprintf(stderr, "[pihm_b200 profile] solve %.3f ms  host-wait %.3f ms in %lld syncs  "
                        "rhs(in situ) %.3f ms over %zu evals (%.1f us each)  launches %lld\n",
                cv->prof_solve_ns * 1e-6, cv->prof_wait_ns * 1e-6, cv->prof_nsync, rhs_ms, cv->prof_used / 2,
                cv->prof_used ? rhs_ms * 2e3 / cv->prof_used : 0.0, (long long)cv->ctx->launches);
    }
    for (cudaEvent_t e : cv->prof_ev) cudaEventDestroy(e);
    cv->prof_ev.clear();
    for (const auto &k : cv->kpend) { cudaEventDestroy(k.a); cudaEventDestroy(k.b); }
    for (cudaEvent_t e : cv->kfree) cudaEventDestroy(e);
    double *all[] = {cv->zn[0], cv->zn[1], cv->zn[2], cv->zn[3], cv->zn[4], cv->zn[5], cv->ewt, cv->acor,
                     cv->tempv, cv->ftemp, cv->V[0], cv->V[1], cv->V[2], cv->V[3], cv->V[4], cv->V[5],
                     cv->vtemp, cv->ytemp, cv->ewt_next, cv->d_part, cv->d_sc};
    for (double *p : all) if (p) cudaFree(p);
    if (cv->p2p && !cv->ctx->lgroup) comm_unshare_buffer(cv->ctx, cv->h_peer);
    if (cv->ctx->lgroup && cv->ctx->lgroup->xbuf[cv->ctx->rank] == cv->d_xbuf) {
        cv->ctx->lgroup->xbuf[cv->ctx->rank] = nullptr;
        cv->ctx->lgroup->peer_tab[cv->ctx->rank] = nullptr;
    }
    if (cv->d_xbuf) cudaFree(cv->d_xbuf);
    if (cv->d_peer) cudaFree(cv->d_peer);
    if (cv->d_counter) cudaFree(cv->d_counter);
    if (cv->h_sc_map && cv->h_sc_map != cv->h_sc) cudaFreeHost(cv->h_sc_map);
    if (cv->h_sc) cudaFreeHost(cv->h_sc);
    delete cv;
}

// CVodeInit / CVodeReInit + CVodeSStolerances + CVodeSet* + CVSpgmr of
// SetCVodeParam (src/ode.c:340-431)
int pihm_b200_cvode_init(pihm_b200_cvode *cv, const pihm_b200_cvode_param *p, double t0,
                         const pihm_b200_vec *y0)
{
    if (!cv || !p || !y0 || y0->n != cv->N) { set_error("cvode_init: bad argument"); return -1; }
    if (p->reltol < 0.0 || p->abstol < 0.0 || p->maxstep < 0.0) { set_error("cvode_init: negative tolerance / hmax"); return -1; }
    cv->tn = t0;
    cv->q = 1; cv->L = 2; cv->qwait = cv->L; cv->etamax = ETAMX1;
    cv->qu = 0; cv->hu = 0.0; cv->tolsf = 1.0;
    cv->copy(y0->d, cv->zn[0]);
    cv->nst = cv->nfe = cv->ncfn = cv->netf = cv->nni = cv->nsetups = cv->nhnil = 0;
    cv->nscon = 0; cv->h0u = 0.0; cv->next_h = 0.0; cv->next_q = 0; cv->nor = 0;
    cv->qprime = 1; cv->h = cv->hprime = cv->hscale = cv->eta = 0.0;
    cv->saved_tq5 = 0.0; cv->indx_acor = 0; cv->crate = 1.0; cv->gammap = 0.0;
    std::memset(cv->ssdat, 0, sizeof(cv->ssdat));
    std::memset(cv->tau, 0, sizeof(cv->tau));
    std::memset(cv->l, 0, sizeof(cv->l));
    std::memset(cv->tq, 0, sizeof(cv->tq));
    cv->reltol = p->reltol; cv->abstol = p->abstol;
    cv->hin = p->initstep;
    cv->sldeton = p->stab_lim_det != 0;
    cv->hmax_inv = (p->maxstep == 0.0) ? 0.0 : 1.0 / p->maxstep;
    cv->mxstep = (p->mxsteps == 0) ? 500 : p->mxsteps;      // cvode_io.c: 0 -> default, <0 -> no limit
    cv->maxl = (p->maxl <= 0) ? CVSPILS_MAXL : std::min(p->maxl, 5);
    cv->sqrtN = std::sqrt(cv->n_global);    // sqrt(dot(1,1)), cvode_spgmr.c:195-196
    cv->nli = cv->ncfl = cv->nfes = cv->njtimes = 0;
    cv->tstopset = false;
    cv->ewt_pending = false;
    cv->ewt_ready = false;
    cv->resc_pending = false;
    cv->initialised = true;
    // ODE() hidden state restarts like a fresh model only when the caller says so;
    // SetCVodeParam itself does not touch elem.wf (src/ode.c:340-431)
    return 0;
}

// ---------------------------------------------------------------------------
// The linear solver alone, as the lsolve hook of a CVODE that keeps its own stepper
// (SURVEY 8(b) "Linear-solver plug-in"): CVSpgmrSolve (cvode_spgmr.c:355-441) with SpgmrSolve,
// CVSpilsAtimes and CVSpilsDQJtimes underneath -- the same fused kernels the integrator uses.
// b, weight, ycur, fcur are the caller's vectors (cv_mem->cv_tempv, cv_ewt, cv_y, cv_ftemp).
// Returns 0 / >0 recoverable / <0 fatal like cv_lsolve; the solution overwrites b.
// ---------------------------------------------------------------------------
int pihm_b200_spgmr_solve(pihm_b200_cvode *cv, double tn, double gamma, double tq4, int mnewt, pihm_b200_vec *b,
                          const pihm_b200_vec *weight, const pihm_b200_vec *ycur, const pihm_b200_vec *fcur)
{
    if (!cv || !b || !weight || !ycur || !fcur || b->n != cv->N || weight->n != cv->N || ycur->n != cv->N ||
        fcur->n != cv->N) {
        set_error("spgmr_solve: bad argument");
        return -1;
    }
    cudaSetDevice(cv->ctx->device);
    // the integrator's own ewt / y / ftemp step aside for the caller's vectors
    double *ewt0 = cv->ewt, *y0 = cv->y, *ft0 = cv->ftemp;
    cv->ewt = weight->d; cv->y = ycur->d; cv->ftemp = fcur->d;
    cv->tn = tn; cv->gamma = gamma; cv->mnewt = mnewt;
    if (cv->maxl <= 0 || cv->maxl > 5) cv->maxl = CVSPILS_MAXL;
    cv->sqrtN = std::sqrt(cv->n_global);            // cvode_spgmr.c:195-196
    int ret;
    {
        cv->clobber(cv->V[0]);
        launch_pdl(cv->s(), cv->ctx->pdl, cv->blocks, PB_VEC_THREADS, k_lsolve_head, cv->N, (const double *)b->d,
                   (const double *)weight->d, cv->V[0], cv->R(SC_BSUM));
        cv->count();
        cv->red(SC_BSUM);
        cv->sync_spin();
        cv->deltar = cv->eplifac * tq4;
        const double bnorm = cv->wrms(SC_BSUM);
        KryPtrs kp{};
        Coef6 c{};
        int krydim = 0;
        bool write_b = false;
        if (bnorm <= cv->deltar) {
            ret = 0;                                 // x = b, or x = 0 after the first Newton iteration
            write_b = mnewt > 0;
        } else {
            cv->delta = cv->deltar * cv->sqrtN;
            bool zero = false;
            ret = cv->spgmrSolve(&zero);
            if (ret == 0) {
                write_b = true;
                if (!zero) {
                    krydim = cv->krydim_last;
                    for (int k = 0; k < krydim; k++) { kp.v[k] = cv->V[k]; c.c[k] = cv->yg[k]; }
                }
            }
        }
        if (write_b) {
            cv->clobber(b->d);
            launch_pdl(cv->s(), cv->ctx->pdl, cv->blocks, PB_VEC_THREADS, k_lsolve_tail, cv->N, krydim, kp, c,
                       (const double *)weight->d, b->d);
            cv->count();
        }
    }
    cv->ewt = ewt0; cv->y = y0; cv->ftemp = ft0;
    if (cudaGetLastError() != cudaSuccess) { set_error("spgmr_solve: kernel launch failed"); return -1; }
    return ret;
}

int pihm_b200_cvode_set_max_step(pihm_b200_cvode *cv, double hmax)
{
    if (!cv || hmax < 0.0) { set_error("set_max_step: bad argument"); return -1; }
    cv->hmax_inv = (hmax == 0.0) ? 0.0 : 1.0 / hmax;
    return 0;
}

int pihm_b200_cvode_solve(pihm_b200_cvode *cv, double tout, pihm_b200_vec *y, double *tret)
{
    if (!cv || !cv->initialised || !y || y->n != cv->N || !tret) { set_error("cvode_solve: bad argument"); return CV_ILL_INPUT; }
    const long long t_solve0 = cv->prof ? pihm_b200_cvode::now_ns() : 0;
    const int flag = cv->solve(tout, y->d, tret);
    if (cv->prof) { cudaStreamSynchronize(cv->ctx->s()); cv->prof_solve_ns += pihm_b200_cvode::now_ns() - t_solve0; }
    int nan_any = (flag >= 0) ? pihm_b200_check_nan(cv->ctx) : 0;
    if (cv->ctx->nranks > 1) {      // every rank must take the same exit
        double v = (double)nan_any;
        cudaMemcpyAsync(cv->d_sc + SC_TMP, &v, sizeof(double), cudaMemcpyHostToDevice, cv->ctx->s());
        comm_allreduce(cv->ctx, cv->d_sc + SC_TMP, 1, 2);
        cudaMemcpyAsync(&v, cv->d_sc + SC_TMP, sizeof(double), cudaMemcpyDeviceToHost, cv->ctx->s());
        cudaStreamSynchronize(cv->ctx->s());
        nan_any = v > 0.0;
    }
    if (flag >= 0 && nan_any) {
        set_error("NAN error in dy (CheckDy, src/ode.c:302-311)");
        return CV_RHSFUNC_FAIL;
    }
    return flag;
}

int pihm_b200_cvode_get_stats(const pihm_b200_cvode *cv, pihm_b200_cvode_stats *st)
{
    if (!cv || !st) return -1;
    st->nst = cv->nst; st->nfe = cv->nfe; st->nni = cv->nni; st->ncfn = cv->ncfn; st->netf = cv->netf;
    st->nli = cv->nli; st->ncfl = cv->ncfl; st->nfeLS = cv->nfes; st->njtimes = cv->njtimes;
    st->nor = cv->nor; st->nsetups = cv->nsetups;
    st->qlast = cv->qu; st->qcur = cv->next_q;
    st->hlast = cv->hu; st->hcur = cv->next_h; st->tcur = cv->tn;
    return 0;
}

// In-situ profile (bench.py): CUDA events around every RHS evaluation the integrator issues and the
// host time spent waiting at its synchronisation points.  Switching it on clears the accumulators.
int pihm_b200_cvode_profile(pihm_b200_cvode *cv, int on)
{
    if (!cv) return -1;
    cudaStreamSynchronize(cv->ctx->s());
    if (on && cv->prof_ev.empty()) {
        cv->prof_ev.resize(16384);
        for (cudaEvent_t &ev : cv->prof_ev)
            if (cudaEventCreate(&ev) != cudaSuccess) { set_error("cvode_profile: cudaEventCreate"); return -1; }
    }
    cv->prof = on != 0;
    if (on) { cv->prof_used = 0; cv->prof_nsync = 0; cv->prof_wait_ns = 0; cv->prof_solve_ns = 0; }
    // on == 2: also a pair of events around every vector kernel (pihm_b200_cvode_get_kernel_profile)
    if (cv->kprof) cv->kflush();
    cv->kprof = on == 2;
    if (cv->kprof) for (auto &k : cv->kacc) k = pihm_b200_cvode::KAcc();
    return 0;
}

// Per-kernel in-situ figures since pihm_b200_cvode_profile(cv, 2): for each vector kernel of the integrator
// that ran, its name (24 bytes, NUL-terminated), the summed event-to-event time, the bytes it read and
// wrote (vector passes x 8 N, the algorithmic traffic) and the number of launches.  Returns the number
// of entries written (<= cap).
int pihm_b200_cvode_get_kernel_profile(pihm_b200_cvode *cv, int cap, char *names, double *ms, double *bytes,
                                       int64_t *launches)
{
    if (!cv || !names || !ms || !bytes || !launches) return -1;
    static const char *const nm[pihm_b200_cvode::K_NKERN] = {
        "k_ewt", "k_predict", "k_rescale", "k_newton_res", "k_krylov_a", "k_krylov_b", "k_krylov_c", "k_mgs_step",
        "k_mgs_chain", "k_spgmr_final", "k_newton_update", "k_wsq", "k_complete", "k_eta", "k_dky",
        "k_elementwise<scale>", "k_elementwise<copy>", "k_linearsum<axpy>", "k_reduce<dot>"};
    cv->kflush();
    int n = 0;
    for (int k = 0; k < pihm_b200_cvode::K_NKERN && n < cap; k++) {
        if (cv->kacc[k].n == 0) continue;
        std::snprintf(names + 24 * n, 24, "%s", nm[k]);
        ms[n] = cv->kacc[k].ms; bytes[n] = cv->kacc[k].bytes; launches[n] = cv->kacc[k].n;
        n++;
    }
    return n;
}

// out[0] ms inside pihm_b200_cvode_solve, out[1] ms of host waiting in out[2] synchronisations,
// out[3] ms of RHS kernels (k_pre + k_main, event to event) over out[4] timed evaluations
int pihm_b200_cvode_get_profile(pihm_b200_cvode *cv, double *out5)
{
    if (!cv || !out5) return -1;
    cudaStreamSynchronize(cv->ctx->s());
    double rhs_ms = 0.0;
    for (size_t i = 0; i + 1 < cv->prof_used; i += 2) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, cv->prof_ev[i], cv->prof_ev[i + 1]);
        rhs_ms += ms;
    }
    out5[0] = cv->prof_solve_ns * 1e-6;
    out5[1] = cv->prof_wait_ns * 1e-6;
    out5[2] = (double)cv->prof_nsync;
    out5[3] = rhs_ms;
    out5[4] = (double)(cv->prof_used / 2);
    return 0;
}

// AdjCVodeMaxStep (src/ode.c:500-560)
int pihm_b200_adj_cvode_max_step(pihm_b200_cvode *cv, pihm_b200_maxstep_ctrl *c)
{
    if (!cv || !c) return -1;
    const double nsteps = (double)(cv->nst - cv->nst0);
    const double nfails = (double)(cv->ncfn - cv->ncfn0) / nsteps;
    const double niters = (double)(cv->nni - cv->nni0) / nsteps;
    if (nfails > c->nncfn || niters >= c->nnimax) c->maxstep /= c->decr;
    if (nfails == 0.0 && niters <= c->nnimin) c->maxstep *= c->incr;
    c->maxstep = (c->maxstep < c->stepsize) ? c->maxstep : c->stepsize;
    c->maxstep = (c->maxstep > c->stmin) ? c->maxstep : c->stmin;
    pihm_b200_cvode_set_max_step(cv, c->maxstep);
    cv->nst0 = cv->nst; cv->ncfn0 = cv->ncfn; cv->nni0 = cv->nni;
    return 0;
}

}  // extern "C"
