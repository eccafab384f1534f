/*
 * pihm_b200_sundials.h -- the SUNDIALS-facing side of the drop-in boundary.
 *
 * MM-PIHM drives CVODE 2.9.0 with a serial N_Vector (N_VNew_Serial,
 * src/include/pihm_func.h:76-79) and the RHS callback ODE() (src/ode.c:3).
 * This header exports the two objects an UNMODIFIED CVODE needs to run on the
 * GPU instead:
 *
 *   N_VNew_PihmB200()  a custom N_Vector whose data lives on the device.  Its
 *                      content struct starts with the fields of
 *                      _N_VectorContent_Serial (nvector_serial.h:74-78), so the
 *                      driver's NV_DATA_S / NV_Ith_S macros keep working on the
 *                      pinned host mirror; N_VPihmB200_Push/Pull move the mirror.
 *                      The ops table fills the 15 slots CVODE+SPGMR call
 *                      (cvCheckNvector cvode.c:1608-1626, cvode_spgmr.c:127)
 *                      plus clone/destroy/space/getvectorid; the rest are NULL.
 *   PihmB200_ODE()     a CVRhsFn (cvode.h:159-160) whose user_data is the
 *                      pihm_b200_ctx instead of pihm_struct.
 *
 * The struct declarations below are layout-compatible restatements of
 * sundials_nvector.h:64-125 (SUNDIALS 2.7.0 NVECTOR); when the real SUNDIALS
 * header was included first they are skipped.
 */
#ifndef PIHM_B200_SUNDIALS_H
#define PIHM_B200_SUNDIALS_H

#include "pihm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef _NVECTOR_H   /* guard macro of sundials/sundials_nvector.h */
typedef double realtype;
typedef int booleantype;
typedef enum {
    SUNDIALS_NVEC_SERIAL, SUNDIALS_NVEC_PARALLEL, SUNDIALS_NVEC_OPENMP, SUNDIALS_NVEC_PTHREADS,
    SUNDIALS_NVEC_PARHYP, SUNDIALS_NVEC_PETSC, SUNDIALS_NVEC_CUSTOM
} N_Vector_ID;
typedef struct _generic_N_Vector_Ops *N_Vector_Ops;
typedef struct _generic_N_Vector *N_Vector;
struct _generic_N_Vector_Ops {
    N_Vector_ID (*nvgetvectorid)(N_Vector);
    N_Vector    (*nvclone)(N_Vector);
    N_Vector    (*nvcloneempty)(N_Vector);
    void        (*nvdestroy)(N_Vector);
    void        (*nvspace)(N_Vector, long int *, long int *);
    realtype   *(*nvgetarraypointer)(N_Vector);
    void        (*nvsetarraypointer)(realtype *, N_Vector);
    void        (*nvlinearsum)(realtype, N_Vector, realtype, N_Vector, N_Vector);
    void        (*nvconst)(realtype, N_Vector);
    void        (*nvprod)(N_Vector, N_Vector, N_Vector);
    void        (*nvdiv)(N_Vector, N_Vector, N_Vector);
    void        (*nvscale)(realtype, N_Vector, N_Vector);
    void        (*nvabs)(N_Vector, N_Vector);
    void        (*nvinv)(N_Vector, N_Vector);
    void        (*nvaddconst)(N_Vector, realtype, N_Vector);
    realtype    (*nvdotprod)(N_Vector, N_Vector);
    realtype    (*nvmaxnorm)(N_Vector);
    realtype    (*nvwrmsnorm)(N_Vector, N_Vector);
    realtype    (*nvwrmsnormmask)(N_Vector, N_Vector, N_Vector);
    realtype    (*nvmin)(N_Vector);
    realtype    (*nvwl2norm)(N_Vector, N_Vector);
    realtype    (*nvl1norm)(N_Vector);
    void        (*nvcompare)(realtype, N_Vector, N_Vector);
    booleantype (*nvinvtest)(N_Vector, N_Vector);
    booleantype (*nvconstrmask)(N_Vector, N_Vector, N_Vector);
    realtype    (*nvminquotient)(N_Vector, N_Vector);
};
struct _generic_N_Vector {
    void           *content;
    struct _generic_N_Vector_Ops *ops;
};
#endif /* _NVECTOR_H */

/* content of a PihmB200 vector: Serial-compatible prefix, then private fields */
typedef struct _N_VectorContent_PihmB200 {
    long int        length;      /* NumStateVar()                                  */
    booleantype     own_data;    /* always TRUE                                    */
    realtype       *data;        /* pinned host mirror, reference order            */
    pihm_b200_vec  *dev;         /* device vector (internal element order)         */
    pihm_b200_ctx  *ctx;
} *N_VectorContent_PihmB200;

/* replaces N_VNew(NumStateVar()) at src/main.c:69 */
N_Vector        N_VNew_PihmB200(pihm_b200_ctx *ctx);
void            N_VDestroy_PihmB200(N_Vector v);
/* host mirror -> device (after InitVar wrote the initial condition,
 * src/initialize.c:569-617) and device -> host mirror (before Summary reads y,
 * src/update.c:9) */
int             N_VPihmB200_Push(N_Vector v);
int             N_VPihmB200_Pull(N_Vector v);
pihm_b200_vec  *N_VPihmB200_Device(N_Vector v);

/* replaces ODE() as the CVRhsFn handed to CVodeInit (src/ode.c:363);
 * user_data = pihm_b200_ctx* (set with CVodeSetUserData, src/ode.c:396).
 * Always returns 0 like the reference; NaNs raise the context's flag
 * (pihm_b200_check_nan) instead of exiting inside the callback. */
int             PihmB200_ODE(realtype t, N_Vector y, N_Vector ydot, void *user_data);

#ifdef __cplusplus
}
#endif
#endif /* PIHM_B200_SUNDIALS_H */
