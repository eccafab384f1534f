// pihm_b200.cu -- context, RHS entry points and N_Vector ops of libpihm_b200.so
// (C ABI declared in include/pihm_b200.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include "common.cuh"
#include "nvec.cuh"
#include "reorder.h"
#include "rhs_layout.cuh"
#include "summary.cuh"
#include "et.cuh"

namespace pb {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <typename T>
static int upload(T **dst, const std::vector<T> &src)
{
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    PB_CUDA(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) PB_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static int vec_blocks(const pihm_b200_ctx *ctx, long long n)
{
    long long b = (n + PB_VEC_THREADS - 1) / PB_VEC_THREADS;
    const long long cap = (long long)ctx->red_blocks;     // SMs x 4 resident CTAs
    return (int)std::max<long long>(1, std::min(b, cap));
}
// N_VDotProd / N_VMaxNorm / N_VWrmsNorm / N_VMin of the stand-alone N_Vector (nvector_b200.cu).
// On a partitioned context the vectors hold the owned unknowns only: the raw sum / max / min is
// all-reduced over the ranks (NCCL) and the WRMS norm divides by the GLOBAL length, so every
// rank of an external CVODE sees the same bits and takes the same branches.
template <int OP, int POST>
static double reduce_sync(pihm_b200_ctx *ctx, long long n, const double *x, const double *y)
{
    const int g = vec_blocks(ctx, n);
    unsigned int *counter = (unsigned int *)(ctx->d_nan + 2);
    double *part = ctx->d_red + 64;
    if (ctx->nranks <= 1) {
        k_reduce<OP, POST><<<g, PB_VEC_THREADS, 0, ctx->s()>>>(n, x, y, part, counter, ctx->d_red, ctx->h_red);
        ctx->launches++;
        cudaStreamSynchronize(ctx->s());
        return ctx->h_red[0];
    }
    k_reduce<OP, 0><<<g, PB_VEC_THREADS, 0, ctx->s()>>>(n, x, y, part, counter, ctx->d_red, nullptr);
    ctx->launches++;
    const int op = (OP == RD_MIN) ? 1 : (OP == RD_MAXABS ? 2 : 0);
    if (pb::comm_allreduce(ctx, ctx->d_red, 1, op) != 0) return std::nan("");
    double v = 0.0;
    cudaMemcpyAsync(&v, ctx->d_red, sizeof(double), cudaMemcpyDeviceToHost, ctx->s());
    cudaStreamSynchronize(ctx->s());
    if (POST == 1) v = std::sqrt(v / (double)ctx->nsv_global);
    return v;
}

}  // namespace pb

using namespace pb;

extern "C" {

const char *pihm_b200_last_error(void) { return g_err.c_str(); }
int pihm_b200_abi_version(void) { return PIHM_B200_ABI_VERSION; }

int pihm_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
pihm_b200_ctx *pihm_b200_create(const pihm_b200_mesh *mesh, int device, int reorder)
{
    if (!mesh) { set_error("pihm_b200_create: bad mesh descriptor"); return nullptr; }
    return pihm_b200_create_part(mesh, device, reorder, mesh->nelem, mesh->nriver);
}

// One partition of a larger mesh (mm-pihm_b200/csrc/partition.cpp): the first
// nown_elem elements / nown_riv rivers of the local tables are owned, the rest
// are ghosts that receive their state through the halo exchange.
pihm_b200_ctx *pihm_b200_create_part(const pihm_b200_mesh *mesh, int device, int reorder, int nown_elem,
                                     int nown_riv)
{
    if (!mesh || mesh->nelem <= 0 || mesh->nriver < 0 || !mesh->elem_f64 || !mesh->elem_i32) {
        set_error("pihm_b200_create: bad mesh descriptor");
        return nullptr;
    }
    if (pihm_b200_device_count() <= device) {
        set_error("pihm_b200_create: no CUDA device (this library has no CPU path)");
        return nullptr;
    }
    const int ne = mesh->nelem, nr = mesh->nriver;
    if (nown_elem < 1 || nown_elem > ne || nown_riv < 0 || nown_riv > nr) {
        set_error("pihm_b200_create_part: owned counts out of range");
        return nullptr;
    }
    const bool partitioned = (nown_elem != ne) || (nown_riv != nr);
    if (partitioned && reorder) {
        set_error("pihm_b200_create_part: a partition is already in locality order (reorder must be 0)");
        return nullptr;
    }
    auto EF = [&](int c, int e) { return mesh->elem_f64[(size_t)c * ne + e]; };
    auto EI = [&](int c, int e) { return mesh->elem_i32[(size_t)c * ne + e]; };
    auto RF = [&](int c, int r) { return mesh->riv_f64[(size_t)c * nr + r]; };
    auto RI = [&](int c, int r) { return mesh->riv_i32[(size_t)c * nr + r]; };

    // ---- validation (what the reference would trip over at run time) -------
    for (int r = 0; r < nr; r++) {
        const int l = RI(PB_RI_LEFTELE, r), rt = RI(PB_RI_RIGHTELE, r), d = RI(PB_RI_DOWN, r);
        const int ord = RI(PB_RI_INTRPL_ORD, r);
        if (l < 1 || l > ne || rt < 1 || rt > ne) {
            set_error("river " + std::to_string(r + 1) + ": bank element out of range");
            return nullptr;
        }
        if (d == 0 || d > nr || d < -4) {   // river_flow.c:381-385 exits on unknown outlet code
            set_error("river " + std::to_string(r + 1) + ": bad downstream code " + std::to_string(d));
            return nullptr;
        }
        if (ord < 1 || ord > 4) {           // river_flow.c:539-543
            set_error("river " + std::to_string(r + 1) + ": river order not defined");
            return nullptr;
        }
    }
    for (int e = 0; e < ne; e++)
        for (int j = 0; j < 3; j++) {
            const int nb = EI(PB_EI_NABR0 + j, e);
            if (nb > ne || nb < -nr) {
                set_error("element " + std::to_string(e + 1) + ": neighbour out of range");
                return nullptr;
            }
        }

    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return nullptr; }

    pihm_b200_ctx *ctx = new pihm_b200_ctx();
    ctx->device = device;
    ctx->reorder = reorder;
    DevMesh &dm = ctx->dm;
    dm.ne = ne; dm.nr = nr;
    dm.nown = nown_elem; dm.rown = nown_riv;
    dm.gs = mesh->fbr ? 3 : 2;
    dm.nes = round_up(ne, 32); dm.nrs = round_up(std::max(nr, 1), 32);
    dm.fbr = mesh->fbr ? 1 : 0;
    dm.surf_mode = mesh->surf_mode; dm.riv_mode = mesh->riv_mode;
    dm.dt = mesh->stepsize;
    // y / ydot hold the OWNED unknowns only, in the block layout of pihm_func.h:7-15
    {
        const long long no = nown_elem, ro = nown_riv;
        dm.o_unsat = no; dm.o_gw = 2 * no; dm.o_stg = 3 * no; dm.o_rgw = 3 * no + ro;
        dm.o_fu = 3 * no + 2 * ro; dm.o_fg = 4 * no + 2 * ro;
        ctx->nsv = 3 * no + 2 * ro + (dm.fbr ? 2 * no : 0);
        ctx->nsv_global = ctx->nsv;
    }

    // ---- internal element order -------------------------------------------
    ctx->perm.resize(ne);
    if (reorder) {
        std::vector<int> nabr((size_t)3 * ne);
        for (int j = 0; j < 3; j++)
            for (int e = 0; e < ne; e++) nabr[(size_t)j * ne + e] = EI(PB_EI_NABR0 + j, e);
        std::vector<int> lr((size_t)2 * nr);
        for (int r = 0; r < nr; r++) { lr[r] = RI(PB_RI_LEFTELE, r); lr[nr + r] = RI(PB_RI_RIGHTELE, r); }
        patch_order(ne, nr, nabr.data(), lr.data(), PB_PATCH, ctx->perm.data());
    } else {
        std::iota(ctx->perm.begin(), ctx->perm.end(), 0);
    }
    ctx->iperm.resize(ne);
    for (int i = 0; i < ne; i++) ctx->iperm[ctx->perm[i]] = i;
    const std::vector<int> &perm = ctx->perm, &iperm = ctx->iperm;

    // ---- river bank edge slots (first match, river_flow.c:159-180) ---------
    ctx->riv_left_edge.assign(nr, -1);
    ctx->riv_right_edge.assign(nr, -1);
    for (int r = 0; r < nr; r++) {
        const int l = RI(PB_RI_LEFTELE, r) - 1, rt = RI(PB_RI_RIGHTELE, r) - 1;
        for (int j = 0; j < 3; j++)
            if (EI(PB_EI_NABR0 + j, l) == -(r + 1)) { ctx->riv_left_edge[r] = j; break; }
        for (int j = 0; j < 3; j++)
            if (EI(PB_EI_NABR0 + j, rt) == -(r + 1)) { ctx->riv_right_edge[r] = j; break; }
    }

    // ---- element columns in internal order ----------------------------------
    const int nes = dm.nes, nrs = dm.nrs;
    // warp-tiled static table [ntile][PB_E_NCOL][32] (padding lanes repeat a benign 1.0)
    const int ntile = nes / 32;
    std::vector<double> es((size_t)ntile * TS_NCOL * 32, 1.0);
    for (int i = 0; i < ne; i++)
        for (int c = 0; c < PB_E_NCOL; c++) {
            const int slot = tile_slot_of(c);
            if (slot >= 0) es[((size_t)(i >> 5) * TS_NCOL + slot) * 32 + (i & 31)] = EF(c, perm[i]);
        }
    // ---- class dictionary of the soil / land-cover / geology columns (rhs.cuh: CC_*) ----------
    std::vector<int> cid(nes, 0);
    std::vector<double> cls;
    {
        std::vector<int> ccols;
        for (int c = 0; c < PB_E_NCOL; c++) {
            const int slot = class_slot_of(c);
            if (slot >= 0 && (dm.fbr || slot < CC_GALPHA)) ccols.push_back(c);
        }
        std::unordered_map<std::string, int> seen;      // key = the row's bit patterns
        std::string key(ccols.size() * sizeof(double), '\0');
        for (int i = 0; i < ne; i++) {
            for (size_t k = 0; k < ccols.size(); k++) {
                const double v = EF(ccols[k], perm[i]);
                std::memcpy(&key[k * sizeof(double)], &v, sizeof(double));
            }
            auto it = seen.find(key);
            if (it == seen.end()) {
                const int id = (int)seen.size();
                it = seen.emplace(key, id).first;
                cls.resize((size_t)(id + 1) * CC_STRIDE, 0.0);
                double *row = &cls[(size_t)id * CC_STRIDE];
                for (size_t k = 0; k < ccols.size(); k++) row[class_slot_of(ccols[k])] = EF(ccols[k], perm[i]);
                // van Genuchten exponents, the divisions of soil.c:5-6 / vert_flow.c:276 done once
                row[CC_M1] = row[CC_BETA] / (row[CC_BETA] - 1.0);
                row[CC_M2] = (row[CC_BETA] - 1.0) / row[CC_BETA];
                row[CC_M3] = 1.0 / row[CC_BETA];
                if (dm.fbr) {
                    row[CC_GM1] = row[CC_GBETA] / (row[CC_GBETA] - 1.0);
                    row[CC_GM2] = (row[CC_GBETA] - 1.0) / row[CC_GBETA];
                    row[CC_GM3] = 1.0 / row[CC_GBETA];
                }
            }
            cid[i] = it->second;
        }
        ctx->nclass = (int)seen.size();
    }
    std::vector<double4> snb(nes, make_double4(0.0, 0.0, 1.0, 0.0));
    for (int i = 0; i < ne; i++)
        snb[i] = make_double4(EF(PB_E_ZMIN, perm[i]), EF(PB_E_ZMAX, perm[i]), EF(PB_E_ROUGH, perm[i]),
                              EF(PB_E_ZBED, perm[i]));
    std::vector<int> nb((size_t)3 * nes, PB_NB_BOUNDARY), bct((size_t)3 * nes, 0), fbct((size_t)3 * nes, 0);
    for (int i = 0; i < ne; i++) {
        const int e = perm[i];
        for (int j = 0; j < 3; j++) {
            const int n = EI(PB_EI_NABR0 + j, e);
            int code;
            if (n > 0) code = iperm[n - 1];
            else if (n == 0) code = PB_NB_BOUNDARY;
            else {
                const int r = -n - 1;
                int side = 2;
                if (RI(PB_RI_LEFTELE, r) - 1 == e && ctx->riv_left_edge[r] == j) side = 0;
                if (RI(PB_RI_RIGHTELE, r) - 1 == e && ctx->riv_right_edge[r] == j) side = 1;
                code = nb_river_code(r, side);
            }
            nb[(size_t)j * nes + i] = code;
            bct[(size_t)j * nes + i] = EI(PB_EI_BC0 + j, e);
            fbct[(size_t)j * nes + i] = EI(PB_EI_FBRBC0 + j, e);
        }
    }
    // leading tiles whose elements read no ghost in k_pre (all neighbours and adjacent rivers owned):
    // the partitioner lists the owned elements as [interior | boundary] (partition.cpp)
    {
        int first_dep = nown_elem;
        for (int i = 0; i < nown_elem && first_dep == nown_elem; i++)
            for (int j = 0; j < 3; j++) {
                const int code = nb[(size_t)j * nes + i];
                if (code >= nown_elem) first_dep = i;
                else if (code <= -2 && ((-code - 2) >> 2) >= nown_riv) first_dep = i;
            }
        ctx->ntile_int = first_dep / 32;
    }
    // leading element tiles above: the strict rule for pihm-fbr also wants both banks of an adjacent river owned
    // (FbrFlow across the river, lat_flow.c:85-100) -- k_main does not use the split, k_pre reads the stage only.
    // the neighbour codes and the class id ride in the tile slab as int32 [4][32] (pseudo-columns TS_NB0/1)
    for (int t = 0; t < ntile; t++) {
        int *dst = reinterpret_cast<int *>(&es[((size_t)t * TS_NCOL + TS_NB0) * 32]);
        for (int j = 0; j < 3; j++)
            for (int l = 0; l < 32; l++) dst[j * 32 + l] = nb[(size_t)j * nes + (size_t)t * 32 + l];
        for (int l = 0; l < 32; l++) dst[96 + l] = cid[(size_t)t * 32 + l];
    }
    // ---- river columns --------------------------------------------------------
    std::vector<double> rf((size_t)PB_R_NCOL * nrs, 0.0), fbr_dist(nrs, 0.0);
    std::vector<int> ri((size_t)PB_RI_NCOL * nrs, 0);
    std::vector<int> up_ptr(nr + 1, 0), up_idx;
    for (int r = 0; r < nr; r++) {
        for (int c = 0; c < PB_R_NCOL; c++) rf[(size_t)c * nrs + r] = RF(c, r);
        for (int c = 0; c < PB_RI_NCOL; c++) ri[(size_t)c * nrs + r] = RI(c, r);
        ri[(size_t)PB_RI_LEFTELE * nrs + r] = iperm[RI(PB_RI_LEFTELE, r) - 1];
        ri[(size_t)PB_RI_RIGHTELE * nrs + r] = iperm[RI(PB_RI_RIGHTELE, r) - 1];
        const int jl = ctx->riv_left_edge[r], jr = ctx->riv_right_edge[r];
        if (jl >= 0 && jr >= 0)
            fbr_dist[r] = EF(PB_E_NABRDIST0 + jl, RI(PB_RI_LEFTELE, r) - 1) +
                EF(PB_E_NABRDIST0 + jr, RI(PB_RI_RIGHTELE, r) - 1);
        else if (dm.fbr && r < nown_riv) {   // lat_flow.c:102-107 "Error finding distance between elements"
            set_error("river " + std::to_string(r + 1) + ": bank elements do not list the river as neighbour");
            delete ctx;
            return nullptr;
        }
        const int d = RI(PB_RI_DOWN, r);
        if (d > 0) up_ptr[d]++;          // count into slot d (= (d-1)+1)
    }
    for (int r = 0; r < nr; r++) up_ptr[r + 1] += up_ptr[r];
    up_idx.resize(up_ptr[nr]);
    {
        std::vector<int> fill(up_ptr.begin(), up_ptr.end() - 1);
        for (int r = 0; r < nr; r++) {   // ascending r == order of river_flow.c:94-108
            const int d = RI(PB_RI_DOWN, r);
            if (d > 0) up_idx[fill[d - 1]++] = r;
        }
    }

    // river tiles (32 segments) in the order k_pre takes them: first those that read owned state only --
    // the segment, both banks, the downstream segment and its banks (river_flow.c:10-86) -- then the rest
    std::vector<int> riv_tile_order;
    int ntile_rc = 0;
    {
        const int ntr = (nr + 31) / 32;
        std::vector<char> ghosty(ntr, 0);
        auto ghost_elem = [&](int e_ref) { return iperm[e_ref - 1] >= nown_elem; };   // 1-based reference index
        for (int r = 0; r < nr; r++) {
            bool g = r >= nown_riv || ghost_elem(RI(PB_RI_LEFTELE, r)) || ghost_elem(RI(PB_RI_RIGHTELE, r));
            const int d = RI(PB_RI_DOWN, r);
            if (d > 0)
                g = g || (d - 1) >= nown_riv || ghost_elem(RI(PB_RI_LEFTELE, d - 1)) || ghost_elem(RI(PB_RI_RIGHTELE, d - 1));
            if (g) ghosty[r >> 5] = 1;
        }
        for (int t = 0; t < ntr; t++) if (!ghosty[t]) riv_tile_order.push_back(t);
        ntile_rc = (int)riv_tile_order.size();
        for (int t = 0; t < ntr; t++) if (ghosty[t]) riv_tile_order.push_back(t);
        if (riv_tile_order.empty()) riv_tile_order.push_back(0);
    }

    int rc = 0;
    rc |= upload(&ctx->d_riv_tile_order, riv_tile_order);
    rc |= upload(&ctx->d_es, es);
    rc |= upload(&ctx->d_cls, cls);
    rc |= upload(&ctx->d_cid, cid);
    rc |= upload(&ctx->d_bct, bct);
    rc |= upload(&ctx->d_fbct, fbct);
    rc |= upload(&ctx->d_rf, rf);
    rc |= upload(&ctx->d_ri, ri);
    rc |= upload(&ctx->d_fbr_dist, fbr_dist);
    rc |= upload(&ctx->d_up_ptr, up_ptr);
    rc |= upload(&ctx->d_up_idx, up_idx);
    rc |= upload(&ctx->d_perm, ctx->perm);
    rc |= upload(&ctx->d_iperm, ctx->iperm);
    auto zalloc = [&](void **p, size_t bytes) {
        if (cudaMalloc(p, bytes) != cudaSuccess) { rc = -1; return; }
        cudaMemset(*p, 0, bytes);
    };
    zalloc((void **)&ctx->d_forc, sizeof(double) * PB_F_NCOL * nes);
    zalloc((void **)&ctx->d_rivbc, sizeof(double) * nrs);
    zalloc((void **)&ctx->d_ft, sizeof(double) * 4 * nes);
    // neighbour records [dnb | snb] in ONE allocation: the L2 access-policy window below covers both
    zalloc((void **)&ctx->d_dnb, sizeof(double4) * nes * 2);
    if (rc == 0) {
        ctx->d_snb = ctx->d_dnb + nes;
        if (cudaMemcpy(ctx->d_snb, snb.data(), sizeof(double4) * nes, cudaMemcpyHostToDevice) != cudaSuccess) rc = -1;
    }
    zalloc((void **)&ctx->d_dist_cold, sizeof(double) * 3 * nes);
    zalloc((void **)&ctx->d_rivflow, sizeof(double) * PIHM_B200_NUM_RIVFLX * nrs);
    zalloc((void **)&ctx->d_stale, sizeof(double) * 2 * nrs);
    zalloc((void **)&ctx->d_nan, sizeof(int) * 4);
    zalloc((void **)&ctx->d_slow, sizeof(unsigned long long));
    zalloc((void **)&ctx->d_dm, sizeof(DevMesh));
    zalloc((void **)&ctx->d_stage, sizeof(double) * std::max<long long>(ctx->nsv, ne));
    zalloc((void **)&ctx->d_gel, sizeof(double) * dm.gs * (size_t)std::max(ne - nown_elem, 1));
    zalloc((void **)&ctx->d_gri, sizeof(double) * 2 * (size_t)std::max(nr - nown_riv, 1));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    ctx->red_blocks = sms * 4;     // one resident wave of the batched vector kernels (<= 64 registers, 4 CTAs / SM)
    if (const char *ov = std::getenv("PIHM_B200_VEC_CTAS"))      // tuning knob: CTAs per SM of the vector kernels
        ctx->red_blocks = sms * std::max(1, std::min(16, std::atoi(ov)));
    zalloc((void **)&ctx->d_red, sizeof(double) * (ctx->red_blocks + 64));
    if (cudaHostAlloc((void **)&ctx->h_red, sizeof(double) * 64, cudaHostAllocMapped) != cudaSuccess) rc = -1;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) rc = -1;
    if (rc != 0) {
        set_error(std::string("pihm_b200_create: device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        pihm_b200_destroy(ctx);
        return nullptr;
    }
    dm.es = ctx->d_es; dm.ft = ctx->d_ft; dm.snb = ctx->d_snb; dm.dnb = ctx->d_dnb; dm.cls = ctx->d_cls; dm.cid = ctx->d_cid; dm.bct = ctx->d_bct; dm.fbct = ctx->d_fbct;
    dm.forc = ctx->d_forc; dm.rf = ctx->d_rf; dm.ri = ctx->d_ri; dm.rivbc = ctx->d_rivbc;
    dm.dist_cold = ctx->d_dist_cold;
    dm.riv_tile_order = ctx->d_riv_tile_order; dm.ntile_rc = ntile_rc;
    dm.fbr_dist = ctx->d_fbr_dist; dm.up_ptr = ctx->d_up_ptr; dm.up_idx = ctx->d_up_idx;
    dm.rivflow = ctx->d_rivflow; dm.s2c_stale = ctx->d_stale;
    if (pb::rhs_tile_rcp(ctx) != 0 ||       // reciprocals of the element areas and neighbour distances
        pb::rhs_class_rcp(ctx) != 0) {      // reciprocals of the dictionary's divisors and of DEPRSTG / dt
        set_error("pihm_b200_create: class dictionary set-up failed");
        pihm_b200_destroy(ctx);
        return nullptr;
    }
    dm.gel = ctx->d_gel; dm.gri = ctx->d_gri;
    dm.xflux = nullptr; dm.record = 0;
    dm.nan_flag = ctx->d_nan;
    dm.slow_count = ctx->d_slow;
    dm.self = ctx->d_dm;
    if (cudaMemcpy(ctx->d_dm, &dm, sizeof(DevMesh), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("pihm_b200_create: DevMesh upload failed");
        pihm_b200_destroy(ctx);
        return nullptr;
    }
    // persistent RHS kernels: opt in to the ring's dynamic shared memory, size the grids to
    // what is resident at once (SMs x CTAs per SM)
    {
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, ctx->device);
        if (pb::rhs_configure(ctx, prop.multiProcessorCount) != 0) {
            pihm_b200_destroy(ctx);
            return nullptr;
        }
        if (const char *ov = std::getenv("PIHM_B200_PDL")) ctx->pdl = std::atoi(ov);
        // B200's 126 MB L2 as a scratchpad: the 64 B / element of neighbour records (written by
        // k_pre, gathered three times per element by k_main, 64 MB at 1M triangles) are marked
        // persisting, everything else that passes through the RHS stream keeps normal priority --
        // the gathers then hit L2 instead of HBM and the records never travel to DRAM and back.
        const char *pe = std::getenv("PIHM_B200_L2_PERSIST");
        // Measured (r01, 1M triangles): RHS 127.6 -> 125.2 us, but the integrator's vector kernels lose
        // the 61 MB set aside (model step 6.98 -> 7.34 ms), so it is opt-in: PIHM_B200_L2_PERSIST=1.
        if (pe && std::atoi(pe) != 0 && prop.persistingL2CacheMaxSize > 0) {
            const size_t rec = sizeof(double4) * (size_t)nes * 2;
            const size_t win = std::min(rec, (size_t)prop.accessPolicyMaxWindowSize);
            const size_t keep = std::min(win, (size_t)prop.persistingL2CacheMaxSize);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, keep) == cudaSuccess) {
                ctx->l2_window.base_ptr = ctx->d_dnb;
                ctx->l2_window.num_bytes = win;
                ctx->l2_window.hitRatio = (float)std::min(1.0, (double)keep / (double)win);
                ctx->l2_window.hitProp = cudaAccessPropertyPersisting;
                ctx->l2_window.missProp = cudaAccessPropertyStreaming;
                ctx->l2_on = 1;
                cudaStreamAttrValue av;
                av.accessPolicyWindow = ctx->l2_window;
                cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            }
            if (std::getenv("PIHM_B200_VERBOSE"))
                fprintf(stderr, "pihm_b200: L2 %d MB, persisting max %d MB, window max %d MB -> records %zu MB, keep %zu MB\n",
                        prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20,
                        rec >> 20, keep >> 20);
            cudaGetLastError();
        }
    }
    ctx->y_tmp = pihm_b200_vec_new(ctx);
    ctx->yd_tmp = pihm_b200_vec_new(ctx);
    if (!ctx->y_tmp || !ctx->yd_tmp) { pihm_b200_destroy(ctx); return nullptr; }
    return ctx;
}

void pihm_b200_destroy(pihm_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    pihm_b200_transfer_release(ctx);                   // copy stream / staging of the pipelined transfers, if used (transfer.cu)
    if (ctx->l2_on) cudaCtxResetPersistingL2Cache();   // hand the set-aside lines back
    pihm_b200_vec_free(ctx->y_tmp);
    pihm_b200_vec_free(ctx->yd_tmp);
    void *dev[] = {ctx->d_riv_tile_order, ctx->d_dist_cold, ctx->d_es, ctx->d_ft, ctx->d_dnb, ctx->d_cls, ctx->d_cid, ctx->d_forc, ctx->d_rf, ctx->d_rivbc, ctx->d_fbr_dist,
                   ctx->d_bct, ctx->d_fbct, ctx->d_ri, ctx->d_up_ptr, ctx->d_up_idx,
                   ctx->d_rivflow, ctx->d_stale, ctx->d_xflux, ctx->d_nan, ctx->d_perm,
                   ctx->d_iperm, ctx->d_stage, ctx->d_red, ctx->d_gel, ctx->d_gri, ctx->d_send_e, ctx->d_send_r,
                   ctx->d_send_e_idx, ctx->d_send_r_idx, ctx->d_slow, ctx->d_dm,
                   ctx->d_last_snap, ctx->d_rec_dy, ctx->d_ws0, ctx->d_subrunoff, ctx->d_dm_rec,
                   ctx->d_etf, ctx->d_eti, ctx->d_eto, ctx->d_et_tab};
    pb::comm_destroy(ctx);
    for (void *p : dev) if (p) cudaFree(p);
    pihm_b200_print_close(ctx);        // files and staging buffers of the output writers (output.cu)
    for (pb::PrintVar &v : ctx->pvars) if (v.acc) cudaFree(v.acc);
    if (ctx->h_et_tab) cudaFreeHost(ctx->h_et_tab);
    if (ctx->et_ev) cudaEventDestroy(ctx->et_ev);
    if (ctx->h_red) cudaFreeHost(ctx->h_red);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int64_t pihm_b200_num_state_var(const pihm_b200_ctx *ctx) { return ctx ? ctx->nsv : 0; }

int pihm_b200_set_stream(pihm_b200_ctx *ctx, void *stream)
{
    if (!ctx) return -1;
    cudaStreamSynchronize(ctx->stream);
    // the context keeps its own stream object alive but launches on the given one
    ctx->user_stream = (cudaStream_t)stream;
    ctx->use_user_stream = 1;
    if (ctx->l2_on) {
        cudaStreamAttrValue av;
        av.accessPolicyWindow = ctx->l2_window;
        cudaStreamSetAttribute(ctx->user_stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    return 0;
}

void *pihm_b200_get_stream(const pihm_b200_ctx *ctx) { return ctx ? (void *)ctx->s() : nullptr; }

int pihm_b200_synchronize(pihm_b200_ctx *ctx)
{
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    return 0;
}

long long pihm_b200_launch_count(const pihm_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

long long pihm_b200_slow_path_count(pihm_b200_ctx *ctx)
{
    if (!ctx) return -1;
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, ctx->d_slow, sizeof(v), cudaMemcpyDeviceToHost, ctx->s()) != cudaSuccess ||
        cudaStreamSynchronize(ctx->s()) != cudaSuccess) { set_error("slow_path_count: copy failed"); return -1; }
    return (long long)v;
}

int pihm_b200_get_permutation(const pihm_b200_ctx *ctx, int32_t *perm)
{
    if (!ctx || !perm) return -1;
    std::copy(ctx->perm.begin(), ctx->perm.end(), perm);
    return 0;
}

// ---- per-step pushes ----------------------------------------------------------
int pihm_b200_set_forcing_col(pihm_b200_ctx *ctx, int col, const double *values)
{
    if (!ctx || col < 0 || col >= PB_F_NCOL || !values) { set_error("set_forcing_col: bad argument"); return -1; }
    const int ne = ctx->dm.ne;
    // straight from the caller's buffer (fast when it is pinned) into device staging, then one
    // kernel applies the internal element order and the tile layout
    PB_CUDA(cudaMemcpyAsync(ctx->d_stage, values, sizeof(double) * ne, cudaMemcpyHostToDevice, ctx->s()));
    k_scatter_forcing<<<(ne + 255) / 256, 256, 0, ctx->s()>>>(ne, ctx->dm.nes, col, ctx->d_perm, ctx->d_stage,
                                                              ctx->d_ft, ctx->d_forc);
    ctx->launches++;
    PB_CUDA(cudaStreamSynchronize(ctx->s()));      // the caller may reuse `values`
    return 0;
}

int pihm_b200_set_forcing(pihm_b200_ctx *ctx, const double *forc)
{
    if (!ctx || !forc) { set_error("set_forcing: bad argument"); return -1; }
    for (int c = 0; c < PB_F_NCOL; c++) {
        // bc columns are only read where a bc_type is set; skip the copy when unused
        if (pihm_b200_set_forcing_col(ctx, c, forc + (size_t)c * ctx->dm.ne) != 0) return -1;
    }
    return 0;
}

int pihm_b200_set_river_bc(pihm_b200_ctx *ctx, const double *bc)
{
    if (!ctx || (!bc && ctx->dm.nr)) { set_error("set_river_bc: bad argument"); return -1; }
    if (ctx->dm.nr == 0) return 0;
    PB_CUDA(cudaMemcpyAsync(ctx->d_rivbc, bc, sizeof(double) * ctx->dm.nr, cudaMemcpyHostToDevice, ctx->s()));
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    return 0;
}

int pihm_b200_set_stale_ovlflow(pihm_b200_ctx *ctx, const double *ovl)
{
    if (!ctx || !ovl) { set_error("set_stale_ovlflow: bad argument"); return -1; }
    const int ne = ctx->dm.ne, nr = ctx->dm.nr, nrs = ctx->dm.nrs;
    if (nr == 0) return 0;
    // elem.wf.ovlflow[j] of a bank edge is -rivflow[LEFT|RIGHT_SURF2CHANL]
    // (river_flow.c:163,175); the next RHS call moves these rows to "stale".
    std::vector<double> rows((size_t)2 * nrs, 0.0);
    std::vector<int> h_ri((size_t)PB_RI_NCOL * nrs);
    PB_CUDA(cudaMemcpy(h_ri.data(), ctx->d_ri, sizeof(int) * h_ri.size(), cudaMemcpyDeviceToHost));
    for (int r = 0; r < nr; r++) {
        const int l = ctx->perm[h_ri[(size_t)PB_RI_LEFTELE * nrs + r]];
        const int rt = ctx->perm[h_ri[(size_t)PB_RI_RIGHTELE * nrs + r]];
        if (ctx->riv_left_edge[r] >= 0) rows[r] = -ovl[(size_t)ctx->riv_left_edge[r] * ne + l];
        if (ctx->riv_right_edge[r] >= 0) rows[nrs + r] = -ovl[(size_t)ctx->riv_right_edge[r] * ne + rt];
    }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    PB_CUDA(cudaMemcpy(ctx->d_rivflow + (size_t)RF_LEFT_S2C * nrs, rows.data(), sizeof(double) * 2 * nrs,
                       cudaMemcpyHostToDevice));
    return 0;
}

// Summary()'s effect on the hot path (src/update.c:19-47): ws0.surf = y[SURF],
// which Infil() reads during the next model step (vert_flow.c:122).  Done on
// the device: one D2D copy of the SURF block into the forcing table.
int pihm_b200_summary(pihm_b200_ctx *ctx, const pihm_b200_vec *y)
{
    if (!ctx || !y || y->n != ctx->nsv) { set_error("pihm_b200_summary: bad argument"); return -1; }
    const int ne = ctx->dm.nown, full = ne / 32, rem = ne % 32;
    double *dst = ctx->d_ft + (size_t)PB_F_WS0SURF * 32;
    if (full)
        PB_CUDA(cudaMemcpy2DAsync(dst, 4 * 32 * sizeof(double), y->d, 32 * sizeof(double), 32 * sizeof(double),
                                  full, cudaMemcpyDeviceToDevice, ctx->s()));
    if (rem)
        PB_CUDA(cudaMemcpyAsync(dst + (size_t)full * 4 * 32, y->d + (size_t)full * 32, rem * sizeof(double),
                                cudaMemcpyDeviceToDevice, ctx->s()));
    return 0;
}

// Ghost records written from the host: lets ONE process emulate all ranks of a
// partitioned run on one GPU (tests), and is what a non-NCCL transport would call.
int pihm_b200_set_ghosts(pihm_b200_ctx *ctx, const double *elem_rec, const double *riv_rec)
{
    if (!ctx) return -1;
    const int ng = ctx->dm.ne - ctx->dm.nown, nrg = ctx->dm.nr - ctx->dm.rown;
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    if (ng > 0 && elem_rec)
        PB_CUDA(cudaMemcpy(ctx->d_gel, elem_rec, sizeof(double) * ctx->dm.gs * ng, cudaMemcpyHostToDevice));
    if (nrg > 0 && riv_rec)
        PB_CUDA(cudaMemcpy(ctx->d_gri, riv_rec, sizeof(double) * 2 * nrg, cudaMemcpyHostToDevice));
    return 0;
}

// Exchange maps of the halo exchange (from pihm_b200_partition_fill)
int pihm_b200_set_halo(pihm_b200_ctx *ctx, int nn, const int32_t *nbr_rank, const int32_t *send_e_ptr,
                       const int32_t *send_e_idx, const int32_t *recv_e_cnt, const int32_t *send_r_ptr,
                       const int32_t *send_r_idx, const int32_t *recv_r_cnt)
{
    if (!ctx || nn < 0) { set_error("set_halo: bad argument"); return -1; }
    ctx->nbr_rank.assign(nbr_rank, nbr_rank + nn);
    ctx->send_e_ptr.assign(send_e_ptr, send_e_ptr + nn + 1);
    ctx->send_r_ptr.assign(send_r_ptr, send_r_ptr + nn + 1);
    ctx->recv_e_cnt.assign(recv_e_cnt, recv_e_cnt + nn);
    ctx->recv_r_cnt.assign(recv_r_cnt, recv_r_cnt + nn);
    ctx->nse = nn ? send_e_ptr[nn] : 0;
    ctx->nsr = nn ? send_r_ptr[nn] : 0;
    long long re = 0, rr = 0;
    for (int k = 0; k < nn; k++) { re += recv_e_cnt[k]; rr += recv_r_cnt[k]; }
    if (re != ctx->dm.ne - ctx->dm.nown || rr != ctx->dm.nr - ctx->dm.rown) {
        set_error("set_halo: receive counts do not match the ghost counts of the local mesh");
        return -1;
    }
    for (int k = 0; k < ctx->nse; k++)
        if (send_e_idx[k] < 0 || send_e_idx[k] >= ctx->dm.nown) { set_error("set_halo: send index not owned"); return -1; }
    for (int k = 0; k < ctx->nsr; k++)
        if (send_r_idx[k] < 0 || send_r_idx[k] >= ctx->dm.rown) { set_error("set_halo: river send index not owned"); return -1; }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    for (void *p : {(void *)ctx->d_send_e, (void *)ctx->d_send_r, (void *)ctx->d_send_e_idx, (void *)ctx->d_send_r_idx})
        if (p) cudaFree(p);
    PB_CUDA(cudaMalloc((void **)&ctx->d_send_e, sizeof(double) * ctx->dm.gs * (size_t)std::max(ctx->nse, 1)));
    PB_CUDA(cudaMalloc((void **)&ctx->d_send_r, sizeof(double) * 2 * (size_t)std::max(ctx->nsr, 1)));
    PB_CUDA(cudaMalloc((void **)&ctx->d_send_e_idx, sizeof(int) * (size_t)std::max(ctx->nse, 1)));
    PB_CUDA(cudaMalloc((void **)&ctx->d_send_r_idx, sizeof(int) * (size_t)std::max(ctx->nsr, 1)));
    if (ctx->nse) PB_CUDA(cudaMemcpy(ctx->d_send_e_idx, send_e_idx, sizeof(int) * ctx->nse, cudaMemcpyHostToDevice));
    if (ctx->nsr) PB_CUDA(cudaMemcpy(ctx->d_send_r_idx, send_r_idx, sizeof(int) * ctx->nsr, cudaMemcpyHostToDevice));
    return 0;
}

// the packed send records of a state vector, on the host (single-process emulation / tests)
int pihm_b200_halo_pack_host(pihm_b200_ctx *ctx, const pihm_b200_vec *y, double *elem_rec, double *riv_rec)
{
    if (!ctx || !y) return -1;
    const int n = ctx->nse + ctx->nsr;
    if (n > 0) {
        pb::rhs_halo_pack(ctx, y->d);
    }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    if (ctx->nse && elem_rec)
        PB_CUDA(cudaMemcpy(elem_rec, ctx->d_send_e, sizeof(double) * ctx->dm.gs * ctx->nse, cudaMemcpyDeviceToHost));
    if (ctx->nsr && riv_rec)
        PB_CUDA(cudaMemcpy(riv_rec, ctx->d_send_r, sizeof(double) * 2 * ctx->nsr, cudaMemcpyDeviceToHost));
    return 0;
}

int pihm_b200_set_flux_recording(pihm_b200_ctx *ctx, int on)
{
    if (!ctx) return -1;
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    if (on && !ctx->d_xflux) {
        PB_CUDA(cudaMalloc((void **)&ctx->d_xflux, sizeof(double) * PB_X_NCOL * ctx->dm.nes));
        PB_CUDA(cudaMemset(ctx->d_xflux, 0, sizeof(double) * PB_X_NCOL * ctx->dm.nes));
    }
    ctx->dm.xflux = ctx->d_xflux;
    ctx->dm.record = on ? 1 : 0;
    PB_CUDA(cudaMemcpy(ctx->d_dm, &ctx->dm, sizeof(DevMesh), cudaMemcpyHostToDevice));
    for (int p = 0; p < 2; p++) {       // the parity views of a peer-memory halo exchange
        if (!ctx->d_dm_par[p]) continue;
        DevMesh dm = ctx->dm;
        dm.gel = ctx->d_hx + p * ctx->hx_stride;
        dm.gri = dm.gel + (size_t)dm.gs * (dm.ne - dm.nown);
        dm.self = ctx->d_dm_par[p];
        PB_CUDA(cudaMemcpy(ctx->d_dm_par[p], &dm, sizeof(DevMesh), cudaMemcpyHostToDevice));
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Summary() + MassBalance() on the device (src/update.c:3-160)
// ---------------------------------------------------------------------------
int pihm_b200_set_diagnostics(pihm_b200_ctx *ctx, int on)
{
    if (!ctx) return -1;
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    if (on && !ctx->d_ws0) {
        const size_t nb = sizeof(double) * (std::max<int64_t>(ctx->nsv, 1) + PB_VEC_PAD);
        if (!ctx->d_xflux) {
            PB_CUDA(cudaMalloc((void **)&ctx->d_xflux, sizeof(double) * PB_X_NCOL * ctx->dm.nes));
            PB_CUDA(cudaMemset(ctx->d_xflux, 0, sizeof(double) * PB_X_NCOL * ctx->dm.nes));
        }
        PB_CUDA(cudaMalloc((void **)&ctx->d_last_snap, nb));
        PB_CUDA(cudaMalloc((void **)&ctx->d_rec_dy, nb));
        PB_CUDA(cudaMalloc((void **)&ctx->d_ws0, nb));
        PB_CUDA(cudaMemset(ctx->d_ws0, 0, nb));
        PB_CUDA(cudaMalloc((void **)&ctx->d_subrunoff, sizeof(double) * ctx->dm.nes));
        PB_CUDA(cudaMemset(ctx->d_subrunoff, 0, sizeof(double) * ctx->dm.nes));
        PB_CUDA(cudaMalloc((void **)&ctx->d_dm_rec, sizeof(DevMesh)));
    }
    if (on && !ctx->dm.xflux) {
        // the column pointer travels with every launch; record stays as it is (0: no writes)
        ctx->dm.xflux = ctx->d_xflux;
        PB_CUDA(cudaMemcpy(ctx->d_dm, &ctx->dm, sizeof(DevMesh), cudaMemcpyHostToDevice));
    }
    if (!on && ctx->last_in == ctx->d_last_snap) ctx->last_in = nullptr;
    ctx->diag = on ? 1 : 0;
    return 0;
}

int pihm_b200_set_ws0(pihm_b200_ctx *ctx, const pihm_b200_vec *y)
{
    if (!ctx || !y || y->n != ctx->nsv) { set_error("pihm_b200_set_ws0: bad argument"); return -1; }
    if (!ctx->d_ws0) { set_error("pihm_b200_set_ws0: call pihm_b200_set_diagnostics(ctx, 1) first"); return -1; }
    PB_CUDA(cudaMemcpyAsync(ctx->d_ws0, y->d, sizeof(double) * ctx->nsv, cudaMemcpyDeviceToDevice, ctx->s()));
    return pihm_b200_summary(ctx, y);      // ws0.surf of Infil()
}

int pihm_b200_summary_mb(pihm_b200_ctx *ctx, const pihm_b200_vec *y, double stepsize)
{
    if (!ctx || !y || y->n != ctx->nsv || !(stepsize > 0.0)) { set_error("pihm_b200_summary_mb: bad argument"); return -1; }
    if (!ctx->diag) { set_error("pihm_b200_summary_mb: call pihm_b200_set_diagnostics(ctx, 1) first"); return -1; }
    if (!ctx->flux_fresh) {
        // wf.* of the last ODE() call: evaluate that call again with the columns on
        if (!ctx->last_in) { set_error("pihm_b200_summary_mb: no RHS call to take the fluxes from"); return -1; }
        if (pb::rhs_launch(ctx, ctx->last_in, ctx->d_rec_dy, true) != 0) return -1;
        ctx->flux_fresh = 1;
    }
    const int ne = ctx->dm.nown;
    if (ne > 0) {
        k_summary_mb<<<(ne + 255) / 256, 256, 0, ctx->s()>>>(ctx->dm, y->d, ctx->d_ws0, ctx->d_subrunoff, stepsize);
        ctx->launches++;
    }
    // update.c:47,94: ws0 = ws
    PB_CUDA(cudaMemcpyAsync(ctx->d_ws0, y->d, sizeof(double) * ctx->nsv, cudaMemcpyDeviceToDevice, ctx->s()));
    PB_CUDA(cudaGetLastError());
    return pihm_b200_summary(ctx, y);
}

int pihm_b200_get_summary(pihm_b200_ctx *ctx, double *subrunoff, double *ws0)
{
    if (!ctx || !ctx->d_ws0) { set_error("pihm_b200_get_summary: diagnostics are off"); return -1; }
    if (subrunoff) {
        const int ne = ctx->dm.nown, nes = ctx->dm.nes;
        std::vector<double> h((size_t)nes);
        PB_CUDA(cudaStreamSynchronize(ctx->s()));
        PB_CUDA(cudaMemcpy(h.data(), ctx->d_subrunoff, sizeof(double) * nes, cudaMemcpyDeviceToHost));
        for (int i = 0; i < ne; i++) subrunoff[ctx->perm[i]] = h[i];
    }
    if (ws0) {
        pihm_b200_vec w;
        w.ctx = ctx; w.d = ctx->d_ws0; w.n = ctx->nsv; w.owns = false;
        if (pihm_b200_vec_download(&w, ws0) != 0) return -1;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Forcing scatter + IntcpSnowEt on the device (src/forcing.c:134-160,242-258; src/is_sm_et.c:4-225)
// ---------------------------------------------------------------------------
int pihm_b200_et_create(pihm_b200_ctx *ctx, const double *et_f64, const int32_t *et_i32)
{
    if (!ctx || !et_f64 || !et_i32) { set_error("et_create: bad argument"); return -1; }
    const int ne = ctx->dm.ne, nes = ctx->dm.nes;
    std::vector<double> f((size_t)PB_ET_NCOL * nes, 0.0);
    std::vector<int> ii((size_t)PB_ETI_NCOL * nes, 1);
    ctx->et_max_meteo = ctx->et_max_lai = ctx->et_max_lc = 0;
    for (int i = 0; i < ne; i++) {
        const int r = ctx->perm[i];
        for (int c = 0; c < PB_ET_NCOL; c++) f[(size_t)c * nes + i] = et_f64[(size_t)c * ne + r];
        for (int c = 0; c < PB_ETI_NCOL; c++) ii[(size_t)c * nes + i] = et_i32[(size_t)c * ne + r];
        const int mt = et_i32[(size_t)PB_ETI_METEO_TYPE * ne + r], lt = et_i32[(size_t)PB_ETI_LAI_TYPE * ne + r],
                  lc = et_i32[(size_t)PB_ETI_LC_TYPE * ne + r];
        if (mt < 1 || lt < 0 || lc < 1) { set_error("et_create: meteo_type / lc_type must be >= 1, lai_type >= 0"); return -1; }
        ctx->et_max_meteo = std::max(ctx->et_max_meteo, mt);
        ctx->et_max_lai = std::max(ctx->et_max_lai, lt);
        ctx->et_max_lc = std::max(ctx->et_max_lc, lc);
    }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    if (!ctx->d_etf) {
        PB_CUDA(cudaMalloc((void **)&ctx->d_etf, sizeof(double) * f.size()));
        PB_CUDA(cudaMalloc((void **)&ctx->d_eti, sizeof(int) * ii.size()));
        PB_CUDA(cudaMalloc((void **)&ctx->d_eto, sizeof(double) * PB_EO_NCOL * nes));
    }
    PB_CUDA(cudaMemcpy(ctx->d_etf, f.data(), sizeof(double) * f.size(), cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(ctx->d_eti, ii.data(), sizeof(int) * ii.size(), cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemset(ctx->d_eto, 0, sizeof(double) * PB_EO_NCOL * nes));
    return 0;
}

int pihm_b200_intcp_snow_et(pihm_b200_ctx *ctx, const pihm_b200_et_step *st, const pihm_b200_vec *y)
{
    if (!ctx || !st || !y || y->n != ctx->nsv) { set_error("intcp_snow_et: bad argument"); return -1; }
    if (!ctx->d_etf) { set_error("intcp_snow_et: call pihm_b200_et_create first"); return -1; }
    if (!(st->stepsize > 0.0) || !st->meteo || !st->lai_lc || !st->z0_lc || st->nmeteo < ctx->et_max_meteo ||
        st->nlc < ctx->et_max_lc || st->nlai < ctx->et_max_lai || (ctx->et_max_lai > 0 && !st->lai)) {
        set_error("intcp_snow_et: a type table is missing or shorter than the largest type index in use");
        return -1;
    }
    // the by-type tables of this step: a few hundred bytes through a pinned staging buffer
    const size_t nm = (size_t)st->nmeteo * PIHM_B200_NUM_METEO_VAR, nl = (size_t)std::max(st->nlai, 0), nc = (size_t)st->nlc;
    const size_t ntab = nm + nl + 2 * nc;
    if (ntab > ctx->et_tab_cap) {
        PB_CUDA(cudaStreamSynchronize(ctx->s()));
        if (ctx->d_et_tab) cudaFree(ctx->d_et_tab);
        if (ctx->h_et_tab) cudaFreeHost(ctx->h_et_tab);
        PB_CUDA(cudaMalloc((void **)&ctx->d_et_tab, sizeof(double) * ntab));
        PB_CUDA(cudaHostAlloc((void **)&ctx->h_et_tab, sizeof(double) * ntab, cudaHostAllocDefault));
        if (!ctx->et_ev) PB_CUDA(cudaEventCreateWithFlags(&ctx->et_ev, cudaEventDisableTiming));
        ctx->et_tab_cap = ntab;
    } else {
        PB_CUDA(cudaEventSynchronize(ctx->et_ev));      // the previous call's upload (long finished)
    }
    double *tab = ctx->h_et_tab;
    std::copy(st->meteo, st->meteo + nm, tab);
    if (nl) std::copy(st->lai, st->lai + nl, tab + nm);
    std::copy(st->lai_lc, st->lai_lc + nc, tab + nm + nl);
    std::copy(st->z0_lc, st->z0_lc + nc, tab + nm + nl + nc);
    PB_CUDA(cudaMemcpyAsync(ctx->d_et_tab, tab, sizeof(double) * ntab, cudaMemcpyHostToDevice, ctx->s()));
    PB_CUDA(cudaEventRecord(ctx->et_ev, ctx->s()));
    EtStepDev d{};
    d.stepsize = st->stepsize; d.cal_edir = st->cal_edir; d.cal_ec = st->cal_ec; d.cal_ett = st->cal_ett;
    d.meltf = st->meltf;
    d.nmeteo = st->nmeteo; d.nlai = st->nlai; d.nlc = st->nlc;
    d.meteo = ctx->d_et_tab; d.lai = ctx->d_et_tab + nm; d.lai_lc = ctx->d_et_tab + nm + nl;
    d.z0_lc = ctx->d_et_tab + nm + nl + nc;
    const int ne = ctx->dm.nown;
    if (ne > 0) {
        k_intcp_snow_et<<<(ne + 127) / 128, 128, 0, ctx->s()>>>(ctx->dm, d, ctx->d_etf, ctx->d_eti, y->d, ctx->d_eto,
                                                                  ctx->d_ft);
        ctx->launches++;
    }
    PB_CUDA(cudaGetLastError());
    return 0;
}

int pihm_b200_et_set_state(pihm_b200_ctx *ctx, const double *sneqv, const double *cmc)
{
    if (!ctx || !ctx->d_eto) { set_error("et_set_state: call pihm_b200_et_create first"); return -1; }
    const int ne = ctx->dm.ne, nes = ctx->dm.nes;
    std::vector<double> h((size_t)nes, 0.0);
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    const double *src[2] = {sneqv, cmc};
    const int col[2] = {PB_EO_SNEQV, PB_EO_CMC};
    for (int k = 0; k < 2; k++) {
        if (!src[k]) continue;
        for (int i = 0; i < ne; i++) h[i] = src[k][ctx->perm[i]];
        PB_CUDA(cudaMemcpy(ctx->d_eto + (size_t)col[k] * nes, h.data(), sizeof(double) * nes, cudaMemcpyHostToDevice));
    }
    return 0;
}

int pihm_b200_et_get(pihm_b200_ctx *ctx, double *out)
{
    if (!ctx || !ctx->d_eto || !out) { set_error("et_get: call pihm_b200_et_create first"); return -1; }
    const int ne = ctx->dm.ne, nes = ctx->dm.nes;
    std::vector<double> h((size_t)PB_EO_NCOL * nes);
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    PB_CUDA(cudaMemcpy(h.data(), ctx->d_eto, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    for (int c = 0; c < PB_EO_NCOL; c++)
        for (int i = 0; i < ne; i++) out[(size_t)c * ne + ctx->perm[i]] = h[(size_t)c * nes + i];
    return 0;
}

// ---------------------------------------------------------------------------
// Print accumulation (UpdPrintVar / PrintData averaging, src/print.c:171-251)
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
k_print_accum(int n, const double *__restrict__ src, double *__restrict__ acc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = acc[i] + src[i];        // buffer[j] += *var[j]
}

// where a print variable reads from (internal order, contiguous), nullptr if unavailable
static const double *print_source(pihm_b200_ctx *ctx, const PrintVar &v, const pihm_b200_vec *y)
{
    const DevMesh &dm = ctx->dm;
    switch (v.src) {
    case PB_PS_STATE: {
        if (!y) return nullptr;
        const long long off[7] = {0, dm.o_unsat, dm.o_gw, dm.o_stg, dm.o_rgw, dm.o_fu, dm.o_fg};
        return y->d + off[v.col];
    }
    case PB_PS_ELEM_FLUX: return ctx->d_xflux ? ctx->d_xflux + (size_t)v.col * dm.nes : nullptr;
    case PB_PS_RIV_FLUX: return ctx->d_rivflow + (size_t)v.col * dm.nrs;
    case PB_PS_ET: return ctx->d_eto ? ctx->d_eto + (size_t)v.col * dm.nes : nullptr;
    }
    return nullptr;
}

int pihm_b200_print_add(pihm_b200_ctx *ctx, int src, int column)
{
    if (!ctx) return -1;
    const DevMesh &dm = ctx->dm;
    PrintVar v;
    v.src = src; v.col = column;
    bool ok = false;
    switch (src) {
    case PB_PS_STATE:
        ok = column >= 0 && column < (dm.fbr ? 7 : 5);
        v.is_river = (column == 3 || column == 4);
        break;
    case PB_PS_ELEM_FLUX: ok = column >= 0 && column < PB_X_NCOL; break;
    case PB_PS_RIV_FLUX: ok = column >= 0 && column < PIHM_B200_NUM_RIVFLX; v.is_river = 1; break;
    case PB_PS_ET: ok = column >= 0 && column < PB_EO_NCOL; break;
    }
    if (!ok) { set_error("print_add: unknown source / column"); return -1; }
    v.len = v.is_river ? dm.rown : dm.nown;
    PB_CUDA(cudaMalloc((void **)&v.acc, sizeof(double) * std::max(v.len, 1)));
    PB_CUDA(cudaMemsetAsync(v.acc, 0, sizeof(double) * std::max(v.len, 1), ctx->s()));
    ctx->pvars.push_back(v);
    return (int)ctx->pvars.size() - 1;
}

int pihm_b200_print_update(pihm_b200_ctx *ctx, const int32_t *ids, int n, const pihm_b200_vec *y)
{
    if (!ctx || (n > 0 && !ids) || (y && y->n != ctx->nsv)) { set_error("print_update: bad argument"); return -1; }
    for (int k = 0; k < n; k++) {
        if (ids[k] < 0 || ids[k] >= (int)ctx->pvars.size()) { set_error("print_update: unknown variable id"); return -1; }
        PrintVar &v = ctx->pvars[ids[k]];
        const double *src = print_source(ctx, v, y);
        if (!src) { set_error("print_update: the variable's source is not available (y / diagnostics / et_create)"); return -1; }
        if (v.len > 0) {
            k_print_accum<<<(v.len + 255) / 256, 256, 0, ctx->s()>>>(v.len, src, v.acc);
            ctx->launches++;
        }
        v.counter++;
    }
    PB_CUDA(cudaGetLastError());
    return 0;
}

int pihm_b200_print_data(pihm_b200_ctx *ctx, int id, double *out, int32_t *counter_out)
{
    if (!ctx || id < 0 || id >= (int)ctx->pvars.size() || !out) { set_error("print_data: bad argument"); return -1; }
    PrintVar &v = ctx->pvars[id];
    std::vector<double> h((size_t)std::max(v.len, 1));
    PB_CUDA(cudaMemcpyAsync(h.data(), v.acc, sizeof(double) * v.len, cudaMemcpyDeviceToHost, ctx->s()));
    PB_CUDA(cudaMemsetAsync(v.acc, 0, sizeof(double) * std::max(v.len, 1), ctx->s()));
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    const double cnt = (double)v.counter;
    for (int i = 0; i < v.len; i++) {
        const double val = (v.counter > 0) ? h[i] / cnt : h[i];        // print.c:234-241
        out[v.is_river ? i : ctx->perm[i]] = val;
    }
    if (counter_out) *counter_out = v.counter;
    v.counter = 0;
    return 0;
}

int pihm_b200_ode(pihm_b200_ctx *ctx, double t, const pihm_b200_vec *y, pihm_b200_vec *ydot)
{
    (void)t;    // unused by the pihm / pihm-fbr physics (ode.c:3)
    if (!ctx || !y || !ydot || y->n != ctx->nsv || ydot->n != ctx->nsv) {
        set_error("pihm_b200_ode: bad argument");
        return -1;
    }
    return pb::rhs_launch(ctx, y->d, ydot->d, false);
}

// timing experiment of the halo protocol (library built with RHS_EXTRA=-DPB_HALO_TIMING; -2 otherwise):
// out[256][8] %globaltimer stamps per RHS evaluation (rhs.cuh), then reset
int pihm_b200_debug_halo_times(pihm_b200_ctx *ctx, unsigned long long *out, int reset)
{
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    return pb::rhs_halo_times(out, reset);
}

int pihm_b200_check_nan(pihm_b200_ctx *ctx)
{
    int flag = 0;
    PB_CUDA(cudaMemcpyAsync(&flag, ctx->d_nan, sizeof(int), cudaMemcpyDeviceToHost, ctx->s()));
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    if (flag) PB_CUDA(cudaMemsetAsync(ctx->d_nan, 0, sizeof(int), ctx->s()));
    return flag ? 1 : 0;
}

int pihm_b200_ode_host(pihm_b200_ctx *ctx, double t, const double *y, double *ydot)
{
    if (!ctx || !y || !ydot) { set_error("pihm_b200_ode_host: bad argument"); return -1; }
    if (pihm_b200_vec_upload(ctx->y_tmp, y) != 0) return -1;
    if (pihm_b200_ode(ctx, t, ctx->y_tmp, ctx->yd_tmp) != 0) return -1;
    if (pihm_b200_vec_download(ctx->yd_tmp, ydot) != 0) return -1;
    PB_CUDA(cudaGetLastError());
    return pihm_b200_check_nan(ctx);
}

int pihm_b200_get_fluxes(pihm_b200_ctx *ctx, double *elem_flux, double *rivflow)
{
    if (!ctx) return -1;
    const int ne = ctx->dm.ne, nes = ctx->dm.nes, nr = ctx->dm.nr, nrs = ctx->dm.nrs;
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    if (elem_flux) {
        if (!ctx->d_xflux || !(ctx->dm.record || ctx->diag)) {
            set_error("get_fluxes: enable pihm_b200_set_flux_recording or pihm_b200_set_diagnostics first");
            return -1;
        }
        std::vector<double> h((size_t)PB_X_NCOL * nes);
        PB_CUDA(cudaMemcpy(h.data(), ctx->d_xflux, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        for (int c = 0; c < PB_X_NCOL; c++)
            for (int i = 0; i < ne; i++) elem_flux[(size_t)c * ne + ctx->perm[i]] = h[(size_t)c * nes + i];
    }
    if (rivflow && nr) {
        std::vector<double> h((size_t)PIHM_B200_NUM_RIVFLX * nrs);
        PB_CUDA(cudaMemcpy(h.data(), ctx->d_rivflow, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        for (int k = 0; k < PIHM_B200_NUM_RIVFLX; k++)
            for (int r = 0; r < nr; r++) rivflow[(size_t)k * nr + r] = h[(size_t)k * nrs + r];
    }
    return 0;
}

// ---------------------------------------------------------------------------
// device vectors
// ---------------------------------------------------------------------------
pihm_b200_vec *pihm_b200_vec_new(pihm_b200_ctx *ctx)
{
    if (!ctx) return nullptr;
    pihm_b200_vec *v = new pihm_b200_vec();
    v->ctx = ctx;
    v->n = ctx->nsv;
    if (cudaMalloc((void **)&v->d, sizeof(double) * (std::max<int64_t>(v->n, 1) + PB_VEC_PAD)) != cudaSuccess) {
        set_error("pihm_b200_vec_new: cudaMalloc failed");
        delete v;
        return nullptr;
    }
    cudaMemsetAsync(v->d, 0, sizeof(double) * (v->n + PB_VEC_PAD), ctx->s());
    return v;
}

void pihm_b200_vec_free(pihm_b200_vec *v)
{
    if (!v) return;
    if (v->owns && v->d) {
        pb::note_free(v->ctx, v->d);
        cudaStreamSynchronize(v->ctx->s());
        cudaFree(v->d);
    }
    delete v;
}

int64_t pihm_b200_vec_length(const pihm_b200_vec *v) { return v ? v->n : 0; }
void *pihm_b200_vec_devptr(pihm_b200_vec *v) { return v ? v->d : nullptr; }

int pihm_b200_vec_upload(pihm_b200_vec *v, const double *host)
{
    pihm_b200_ctx *ctx = v->ctx;
    pb::note_write(ctx, v->d);
    if (!ctx->reorder) {
        PB_CUDA(cudaMemcpyAsync(v->d, host, sizeof(double) * v->n, cudaMemcpyHostToDevice, ctx->s()));
    } else {
        PB_CUDA(cudaMemcpyAsync(ctx->d_stage, host, sizeof(double) * v->n, cudaMemcpyHostToDevice, ctx->s()));
        k_permute_state<<<vec_blocks(ctx, v->n), PB_VEC_THREADS, 0, ctx->s()>>>(
            ctx->dm.ne, ctx->dm.nr, ctx->dm.fbr, ctx->d_perm, ctx->d_stage, v->d, 1);
        ctx->launches++;
    }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));      // the caller may reuse `host`
    return 0;
}

int pihm_b200_vec_download(const pihm_b200_vec *v, double *host)
{
    pihm_b200_ctx *ctx = v->ctx;
    if (!ctx->reorder) {
        PB_CUDA(cudaMemcpyAsync(host, v->d, sizeof(double) * v->n, cudaMemcpyDeviceToHost, ctx->s()));
    } else {
        k_permute_state<<<vec_blocks(ctx, v->n), PB_VEC_THREADS, 0, ctx->s()>>>(
            ctx->dm.ne, ctx->dm.nr, ctx->dm.fbr, ctx->d_perm, v->d, ctx->d_stage, 0);
        ctx->launches++;
        PB_CUDA(cudaMemcpyAsync(host, ctx->d_stage, sizeof(double) * v->n, cudaMemcpyDeviceToHost, ctx->s()));
    }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    return 0;
}

// ---- streaming ops (nvector_serial.c:421-635) ---------------------------------
void pihm_b200_nv_linearsum(double a, const pihm_b200_vec *x, double b, const pihm_b200_vec *y,
                            pihm_b200_vec *z)
{
    pihm_b200_ctx *ctx = z->ctx;
    const long long n = z->n;
    const int g = vec_blocks(ctx, n);
    // VScaleSum / VScaleDiff are the only cases that round differently from
    // (a*x)+(b*y); they are reached only when no +-1 coefficient is involved
    // (nvector_serial.c:435-480).
    const bool unit = (a == 1.0 || a == -1.0 || b == 1.0 || b == -1.0);
    const int mode = (!unit && a == b) ? LS_SCALESUM : ((!unit && a == -b) ? LS_SCALEDIFF : LS_GENERAL);
    const bool alias = (z->d == x->d) || (z->d == y->d);
    cudaStream_t s = ctx->s();
    pb::note_write(ctx, z->d);
    if (alias) {
        if (mode == LS_GENERAL) k_linearsum_alias<LS_GENERAL><<<g, PB_VEC_THREADS, 0, s>>>(n, a, x->d, b, y->d, z->d);
        else if (mode == LS_SCALESUM) k_linearsum_alias<LS_SCALESUM><<<g, PB_VEC_THREADS, 0, s>>>(n, a, x->d, b, y->d, z->d);
        else k_linearsum_alias<LS_SCALEDIFF><<<g, PB_VEC_THREADS, 0, s>>>(n, a, x->d, b, y->d, z->d);
    } else {
        if (mode == LS_GENERAL) k_linearsum<LS_GENERAL><<<g, PB_VEC_THREADS, 0, s>>>(n, a, x->d, b, y->d, z->d);
        else if (mode == LS_SCALESUM) k_linearsum<LS_SCALESUM><<<g, PB_VEC_THREADS, 0, s>>>(n, a, x->d, b, y->d, z->d);
        else k_linearsum<LS_SCALEDIFF><<<g, PB_VEC_THREADS, 0, s>>>(n, a, x->d, b, y->d, z->d);
    }
    ctx->launches++;
}

#define PB_EW(OP, c, xp, yp)                                                                     \
    do {                                                                                         \
        pihm_b200_ctx *ctx = z->ctx;                                                             \
        pb::note_write(ctx, z->d);                                                               \
        k_elementwise<OP><<<vec_blocks(ctx, z->n), PB_VEC_THREADS, 0, ctx->s()>>>(z->n, c, xp, yp, z->d); \
        ctx->launches++;                                                                         \
    } while (0)

void pihm_b200_nv_const(double c, pihm_b200_vec *z) { PB_EW(EW_CONST, c, nullptr, nullptr); }
void pihm_b200_nv_prod(const pihm_b200_vec *x, const pihm_b200_vec *y, pihm_b200_vec *z) { PB_EW(EW_PROD, 0.0, x->d, y->d); }
void pihm_b200_nv_div(const pihm_b200_vec *x, const pihm_b200_vec *y, pihm_b200_vec *z) { PB_EW(EW_DIV, 0.0, x->d, y->d); }
void pihm_b200_nv_scale(double c, const pihm_b200_vec *x, pihm_b200_vec *z)
{
    // c == 1 copies, c == -1 negates: both equal c * x bitwise (nvector_serial.c:558-584)
    if (c == 1.0) {
        if (z->d != x->d) PB_EW(EW_COPY, 0.0, x->d, nullptr);
        return;
    }
    PB_EW(EW_SCALE, c, x->d, nullptr);
}
void pihm_b200_nv_abs(const pihm_b200_vec *x, pihm_b200_vec *z) { PB_EW(EW_ABS, 0.0, x->d, nullptr); }
void pihm_b200_nv_inv(const pihm_b200_vec *x, pihm_b200_vec *z) { PB_EW(EW_INV, 0.0, x->d, nullptr); }
void pihm_b200_nv_addconst(const pihm_b200_vec *x, double b, pihm_b200_vec *z) { PB_EW(EW_ADDCONST, b, x->d, nullptr); }

// ---- reductions (nvector_serial.c:637-725) -------------------------------------
double pihm_b200_nv_dotprod(const pihm_b200_vec *x, const pihm_b200_vec *y)
{
    return reduce_sync<RD_DOT, 0>(x->ctx, x->n, x->d, y->d);
}
double pihm_b200_nv_maxnorm(const pihm_b200_vec *x)
{
    return reduce_sync<RD_MAXABS, 0>(x->ctx, x->n, x->d, nullptr);
}
double pihm_b200_nv_wrmsnorm(const pihm_b200_vec *x, const pihm_b200_vec *w)
{
    return reduce_sync<RD_WSQ, 1>(x->ctx, x->n, x->d, w->d);
}
double pihm_b200_nv_min(const pihm_b200_vec *x)
{
    return reduce_sync<RD_MIN, 0>(x->ctx, x->n, x->d, nullptr);
}

}  // extern "C"
