// output.cu -- the output path fed from the device columns (second half):
// binary .dat / ASCII .txt print records and the restart .ic file, written by the library from
// the running sums it keeps in HBM (pihm_b200_print_add / print_update, pihm_b200.cu).
//
// Reference: InitOutputFile (src/print.c:72-151) opens "<name>.dat" and "<name>.txt" per print
// variable; PrintData (src/print.c:193-251) writes, for every variable PrintNow() says is due,
// one record { double time ; double value[nvar] } with value = buffer / counter (buffer itself
// when counter == 0) and resets the sums; PrintInit (src/print.c:253-313) writes the restart
// file { cmc, sneqv, surf, unsat, gw [, fbr_unsat, fbr_gw] } per element, { stage, gw } per river.
// The reference walks one pointer per element and variable into its host structs
// (map_output.c:33-262), so every printed field has to be on the host every step.  Here all the
// variables that are due at a print time are averaged, reset and packed by ONE kernel into one
// staging buffer and cross PCIe in ONE copy; the host un-permutes and fwrite()s the records.
// Same bytes as the reference's files for the same values (IEEE division on both sides).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "common.cuh"

using namespace pb;

namespace {

struct PackItem { const double *acc_c; double *acc; long long off; int len; int counter; };

// out[off + i] = (counter > 0) ? acc[i] / counter : acc[i];  acc[i] = 0      (print.c:232-243)
__global__ void __launch_bounds__(256)
k_print_pack(int nitems, const PackItem *__restrict__ items, double *__restrict__ out)
{
    for (int v = blockIdx.y; v < nitems; v += gridDim.y) {
        const PackItem it = items[v];
        const double cnt = (double)it.counter;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < it.len; i += gridDim.x * blockDim.x) {
            const double b = it.acc[i];
            out[it.off + i] = (it.counter > 0) ? b / cnt : b;
            it.acc[i] = 0.0;
        }
    }
}

// restart record of one element / river from the state vector (internal order) and the two
// storages IntcpSnowEt integrates; written in internal order, un-permuted on the host
__global__ void __launch_bounds__(256)
k_ic_pack(const DevMesh m, const double *__restrict__ y, const double *__restrict__ eto, int nes_eto,
          double *__restrict__ out)
{
    const int ne = m.nown, nr = m.rown, per = m.fbr ? 7 : 5;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ne) {
        double *o = out + (size_t)i * per;
        o[0] = eto ? eto[(size_t)PB_EO_CMC * nes_eto + i] : 0.0;
        o[1] = eto ? eto[(size_t)PB_EO_SNEQV * nes_eto + i] : 0.0;
        o[2] = y[i];
        o[3] = y[m.o_unsat + i];
        o[4] = y[m.o_gw + i];
        if (m.fbr) { o[5] = y[m.o_fu + i]; o[6] = y[m.o_fg + i]; }
    } else if (i < ne + nr) {
        const int r = i - ne;
        double *o = out + (size_t)ne * per + (size_t)r * 2;
        o[0] = y[m.o_stg + r];
        o[1] = y[m.o_rgw + r];
    }
}

struct OutFile { FILE *dat = nullptr, *txt = nullptr; };
struct OutState {
    std::vector<OutFile> files;        // by print variable id
    double *h_stage = nullptr;         // pinned
    double *d_stage = nullptr;
    PackItem *d_items = nullptr;
    size_t cap = 0, item_cap = 0;
    long long d2h_copies = 0, d2h_bytes = 0;
};

// one OutState per context, kept outside the ctx struct (host-only bookkeeping of this TU)
std::vector<std::pair<pihm_b200_ctx *, OutState *>> g_out;

OutState *state_of(pihm_b200_ctx *ctx, bool create)
{
    for (auto &p : g_out) if (p.first == ctx) return p.second;
    if (!create) return nullptr;
    g_out.emplace_back(ctx, new OutState());
    return g_out.back().second;
}

int ensure_stage(OutState *st, size_t doubles, size_t items)
{
    if (doubles > st->cap) {
        if (st->h_stage) cudaFreeHost(st->h_stage);
        if (st->d_stage) cudaFree(st->d_stage);
        st->h_stage = nullptr; st->d_stage = nullptr; st->cap = 0;
        PB_CUDA(cudaHostAlloc((void **)&st->h_stage, sizeof(double) * doubles, cudaHostAllocDefault));
        PB_CUDA(cudaMalloc((void **)&st->d_stage, sizeof(double) * doubles));
        st->cap = doubles;
    }
    if (items > st->item_cap) {
        if (st->d_items) cudaFree(st->d_items);
        st->d_items = nullptr; st->item_cap = 0;
        PB_CUDA(cudaMalloc((void **)&st->d_items, sizeof(PackItem) * items));
        st->item_cap = items;
    }
    return 0;
}

}  // namespace

extern "C" {

// InitOutputFile for one print variable (src/print.c:116-129): "<name>.dat" and, if ascii,
// "<name>.txt"; mode "w" or (append_mode, print.c:82-89) "a"
int pihm_b200_print_open(pihm_b200_ctx *ctx, int id, const char *name, int ascii, int append)
{
    if (!ctx || !name || id < 0 || id >= (int)ctx->pvars.size()) { set_error("print_open: bad argument"); return -1; }
    OutState *st = state_of(ctx, true);
    if (st->files.size() < ctx->pvars.size()) st->files.resize(ctx->pvars.size());
    OutFile &f = st->files[id];
    if (f.dat) { fclose(f.dat); f.dat = nullptr; }
    if (f.txt) { fclose(f.txt); f.txt = nullptr; }
    const std::string base(name);
    f.dat = fopen((base + ".dat").c_str(), append ? "a" : "w");
    if (!f.dat) { set_error("print_open: cannot open " + base + ".dat"); return -1; }
    if (ascii) {
        f.txt = fopen((base + ".txt").c_str(), append ? "a" : "w");
        if (!f.txt) { set_error("print_open: cannot open " + base + ".txt"); return -1; }
    }
    return 0;
}

// PrintData (src/print.c:193-251) for the listed variables -- the ones PrintNow() says are due at
// model time t (the caller keeps that host time logic, print.c:578-624): average, reset, ONE copy
// to the host, then per variable one record  { (double)t ; values in reference order }  to its .dat
// file and  "<timestr>"\tvalue...\n  (%lf) to its .txt file.  Variables without an open file are
// averaged and reset like the others (their record is dropped).
int pihm_b200_print_write(pihm_b200_ctx *ctx, const int32_t *ids, int n, int t, const char *timestr)
{
    if (!ctx || (n > 0 && !ids)) { set_error("print_write: bad argument"); return -1; }
    if (n <= 0) return 0;
    OutState *st = state_of(ctx, true);
    if (st->files.size() < ctx->pvars.size()) st->files.resize(ctx->pvars.size());
    std::vector<PackItem> items((size_t)n);
    long long total = 0;
    int maxlen = 1;
    for (int k = 0; k < n; k++) {
        if (ids[k] < 0 || ids[k] >= (int)ctx->pvars.size()) { set_error("print_write: unknown variable id"); return -1; }
        PrintVar &v = ctx->pvars[ids[k]];
        items[k] = PackItem{v.acc, v.acc, total, v.len, v.counter};
        total += v.len;
        maxlen = std::max(maxlen, v.len);
    }
    if (ensure_stage(st, (size_t)std::max<long long>(total, 1), (size_t)n) != 0) return -1;
    PB_CUDA(cudaMemcpyAsync(st->d_items, items.data(), sizeof(PackItem) * n, cudaMemcpyHostToDevice, ctx->s()));
    if (total > 0) {
        const dim3 grid((unsigned)std::min(256, (maxlen + 255) / 256), (unsigned)std::min(n, 64));
        k_print_pack<<<grid, 256, 0, ctx->s()>>>(n, st->d_items, st->d_stage);
        ctx->launches++;
        PB_CUDA(cudaMemcpyAsync(st->h_stage, st->d_stage, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->s()));
        st->d2h_copies++;
        st->d2h_bytes += (long long)sizeof(double) * total;
    }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    PB_CUDA(cudaGetLastError());
    std::vector<double> rec;
    const double outtime = (double)t;
    for (int k = 0; k < n; k++) {
        PrintVar &v = ctx->pvars[ids[k]];
        v.counter = 0;
        const OutFile &f = st->files[ids[k]];
        if (!f.dat && !f.txt) continue;
        const double *h = st->h_stage + items[k].off;
        rec.resize((size_t)v.len);
        if (v.is_river) std::memcpy(rec.data(), h, sizeof(double) * v.len);
        else for (int i = 0; i < v.len; i++) rec[ctx->perm[i]] = h[i];
        if (f.txt) {
            fprintf(f.txt, "\"%s\"", timestr ? timestr : "");
            for (int i = 0; i < v.len; i++) fprintf(f.txt, "\t%lf", rec[i]);
            fprintf(f.txt, "\n");
            fflush(f.txt);
        }
        if (f.dat) {
            fwrite(&outtime, sizeof(double), 1, f.dat);
            fwrite(rec.data(), sizeof(double), (size_t)v.len, f.dat);
            fflush(f.dat);
        }
    }
    return 0;
}

// device -> host copies the writers have made so far (tests: one per print time) and their bytes
int pihm_b200_print_io_stats(pihm_b200_ctx *ctx, int64_t *copies, int64_t *bytes)
{
    OutState *st = ctx ? state_of(ctx, false) : nullptr;
    if (copies) *copies = st ? st->d2h_copies : 0;
    if (bytes) *bytes = st ? st->d2h_bytes : 0;
    return 0;
}

int pihm_b200_print_close(pihm_b200_ctx *ctx)
{
    for (size_t k = 0; k < g_out.size(); k++) {
        if (g_out[k].first != ctx) continue;
        OutState *st = g_out[k].second;
        for (OutFile &f : st->files) {
            if (f.dat) fclose(f.dat);
            if (f.txt) fclose(f.txt);
        }
        if (st->h_stage) cudaFreeHost(st->h_stage);
        if (st->d_stage) cudaFree(st->d_stage);
        if (st->d_items) cudaFree(st->d_items);
        delete st;
        g_out.erase(g_out.begin() + k);
        break;
    }
    return 0;
}

// PrintInit (src/print.c:253-313): the restart file of state y.  ws.cmc / ws.sneqv come from the
// device ET state (pihm_b200_et_create) unless the caller passes them (reference order, [nelem]
// each; e.g. the unchanged host IntcpSnowEt of route 1 / 2); without either they are zero, like
// Initialize() without an .ic file leaves them.
int pihm_b200_write_ic(pihm_b200_ctx *ctx, const char *path, const pihm_b200_vec *y, const double *cmc,
                       const double *sneqv)
{
    if (!ctx || !path || !y || y->n != ctx->nsv) { set_error("write_ic: bad argument"); return -1; }
    const DevMesh &dm = ctx->dm;
    const int ne = dm.nown, nr = dm.rown, per = dm.fbr ? 7 : 5;
    const size_t total = (size_t)ne * per + (size_t)nr * 2;
    OutState *st = state_of(ctx, true);
    if (ensure_stage(st, std::max<size_t>(total, 1), 1) != 0) return -1;
    if (ne + nr > 0) {
        k_ic_pack<<<(ne + nr + 255) / 256, 256, 0, ctx->s()>>>(dm, y->d, ctx->d_eto, dm.nes, st->d_stage);
        ctx->launches++;
        PB_CUDA(cudaMemcpyAsync(st->h_stage, st->d_stage, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->s()));
        st->d2h_copies++;
        st->d2h_bytes += (long long)sizeof(double) * total;
    }
    PB_CUDA(cudaStreamSynchronize(ctx->s()));
    PB_CUDA(cudaGetLastError());
    std::vector<double> rec(total);
    for (int i = 0; i < ne; i++) {
        const int e = ctx->perm[i];
        double *o = rec.data() + (size_t)e * per;
        std::memcpy(o, st->h_stage + (size_t)i * per, sizeof(double) * per);
        if (cmc) o[0] = cmc[e];
        if (sneqv) o[1] = sneqv[e];
    }
    if (nr > 0) std::memcpy(rec.data() + (size_t)ne * per, st->h_stage + (size_t)ne * per, sizeof(double) * 2 * nr);
    FILE *fp = fopen(path, "wb");
    if (!fp) { set_error(std::string("write_ic: cannot open ") + path); return -1; }
    const size_t w = fwrite(rec.data(), sizeof(double), total, fp);
    fflush(fp);
    fclose(fp);
    if (w != total) { set_error(std::string("write_ic: short write to ") + path); return -1; }
    return 0;
}

}  // extern "C"
