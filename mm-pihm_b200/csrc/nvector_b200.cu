// nvector_b200.cu -- SUNDIALS 2.7 custom N_Vector over a device vector, and the
// CVRhsFn wrapper: what an unmodified CVODE needs to run MM-PIHM on the GPU
// (include/pihm_b200_sundials.h).  Pattern of constructors / destructor follows
// N_VNewEmpty_Serial / N_VClone_Serial / N_VDestroy_Serial
// (cvode/src/nvec_ser/nvector_serial.c:76-130, 337-399).
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "pihm_b200_sundials.h"

namespace {

inline N_VectorContent_PihmB200 C(N_Vector v) { return (N_VectorContent_PihmB200)v->content; }
inline pihm_b200_vec *D(N_Vector v) { return C(v)->dev; }

N_Vector_ID nv_id(N_Vector) { return SUNDIALS_NVEC_CUSTOM; }
N_Vector nv_clone(N_Vector w);
void nv_destroy(N_Vector v) { N_VDestroy_PihmB200(v); }
void nv_space(N_Vector v, long int *lrw, long int *liw) { *lrw = C(v)->length; *liw = 1; }
void nv_linearsum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) { pihm_b200_nv_linearsum(a, D(x), b, D(y), D(z)); }
void nv_const(realtype c, N_Vector z) { pihm_b200_nv_const(c, D(z)); }
void nv_prod(N_Vector x, N_Vector y, N_Vector z) { pihm_b200_nv_prod(D(x), D(y), D(z)); }
void nv_div(N_Vector x, N_Vector y, N_Vector z) { pihm_b200_nv_div(D(x), D(y), D(z)); }
void nv_scale(realtype c, N_Vector x, N_Vector z) { pihm_b200_nv_scale(c, D(x), D(z)); }
void nv_abs(N_Vector x, N_Vector z) { pihm_b200_nv_abs(D(x), D(z)); }
void nv_inv(N_Vector x, N_Vector z) { pihm_b200_nv_inv(D(x), D(z)); }
void nv_addconst(N_Vector x, realtype b, N_Vector z) { pihm_b200_nv_addconst(D(x), b, D(z)); }
realtype nv_dot(N_Vector x, N_Vector y) { return pihm_b200_nv_dotprod(D(x), D(y)); }
realtype nv_maxnorm(N_Vector x) { return pihm_b200_nv_maxnorm(D(x)); }
realtype nv_wrms(N_Vector x, N_Vector w) { return pihm_b200_nv_wrmsnorm(D(x), D(w)); }
realtype nv_min(N_Vector x) { return pihm_b200_nv_min(D(x)); }

N_Vector make(pihm_b200_ctx *ctx)
{
    N_Vector v = (N_Vector)std::calloc(1, sizeof(*v));
    N_Vector_Ops ops = (N_Vector_Ops)std::calloc(1, sizeof(*ops));   // unused slots stay NULL
    N_VectorContent_PihmB200 c = (N_VectorContent_PihmB200)std::calloc(1, sizeof(*c));
    if (!v || !ops || !c) { std::free(v); std::free(ops); std::free(c); return nullptr; }
    ops->nvgetvectorid = nv_id;
    ops->nvclone = nv_clone;
    ops->nvdestroy = nv_destroy;
    ops->nvspace = nv_space;
    ops->nvlinearsum = nv_linearsum;
    ops->nvconst = nv_const;
    ops->nvprod = nv_prod;
    ops->nvdiv = nv_div;
    ops->nvscale = nv_scale;
    ops->nvabs = nv_abs;
    ops->nvinv = nv_inv;
    ops->nvaddconst = nv_addconst;
    ops->nvdotprod = nv_dot;
    ops->nvmaxnorm = nv_maxnorm;
    ops->nvwrmsnorm = nv_wrms;
    ops->nvmin = nv_min;
    c->length = (long int)pihm_b200_num_state_var(ctx);
    c->own_data = 1;
    c->ctx = ctx;
    c->dev = pihm_b200_vec_new(ctx);
    if (!c->dev || cudaHostAlloc((void **)&c->data, sizeof(double) * (size_t)(c->length > 0 ? c->length : 1),
                                 cudaHostAllocDefault) != cudaSuccess) {
        if (c->dev) pihm_b200_vec_free(c->dev);
        std::free(v); std::free(ops); std::free(c);
        pb::set_error("N_VNew_PihmB200: allocation failed");
        return nullptr;
    }
    std::memset(c->data, 0, sizeof(double) * (size_t)c->length);
    v->content = c;
    v->ops = ops;
    return v;
}

N_Vector nv_clone(N_Vector w) { return make(C(w)->ctx); }

}  // namespace

extern "C" {

N_Vector N_VNew_PihmB200(pihm_b200_ctx *ctx)
{
    if (!ctx) { pb::set_error("N_VNew_PihmB200: null context"); return nullptr; }
    return make(ctx);
}

void N_VDestroy_PihmB200(N_Vector v)
{
    if (!v) return;
    N_VectorContent_PihmB200 c = C(v);
    if (c) {
        if (c->dev) pihm_b200_vec_free(c->dev);
        if (c->data) cudaFreeHost(c->data);
        std::free(c);
    }
    std::free(v->ops);
    std::free(v);
}

int N_VPihmB200_Push(N_Vector v) { return pihm_b200_vec_upload(D(v), C(v)->data); }
int N_VPihmB200_Pull(N_Vector v) { return pihm_b200_vec_download(D(v), C(v)->data); }
pihm_b200_vec *N_VPihmB200_Device(N_Vector v) { return v ? D(v) : nullptr; }

int PihmB200_ODE(realtype t, N_Vector y, N_Vector ydot, void *user_data)
{
    pihm_b200_ctx *ctx = (pihm_b200_ctx *)user_data;
    // ODE() itself always returns 0 (src/ode.c:299) and exits on a NaN (CheckDy, :302-311).  Here a failed
    // launch is an unrecoverable RHS error for CVODE (< 0, cvode.c:2683-2684); NaNs raise the context's flag
    // on the device -- the caller asks pihm_b200_check_nan(ctx) after CVode() returns (the flag costs a
    // device -> host copy, which does not belong between two kernels of the integrator).
    return (pihm_b200_ode(ctx, t, D(y), D(ydot)) != 0) ? -1 : 0;
}

}  // extern "C"
