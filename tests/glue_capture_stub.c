/* TEST INFRASTRUCTURE: an LD_PRELOAD interposer for tests/test_glue_packer.py.  It replaces ONE entry point of
 * the C ABI, pihm_b200_create(), inside the unchanged pihm / pihm-fbr programs linked with glue/pihm_b200_glue.c:
 * the tables the glue packed from pihm_struct are written to $PIHM_B200_CAPTURE and the program ends.  It
 * computes nothing and is never loaded by the product. */
#include <stdio.h>
#include <stdlib.h>
#include "pihm_b200.h"

pihm_b200_ctx *pihm_b200_create(const pihm_b200_mesh *m, int device, int reorder)
{
    const char *path = getenv("PIHM_B200_CAPTURE");
    FILE       *f = fopen(path ? path : "glue_capture.bin", "wb");
    int32_t     head[8];
    size_t      ne = (size_t)m->nelem, nr = (size_t)m->nriver;

    if (f == NULL)
    {
        exit(3);
    }
    head[0] = m->nelem; head[1] = m->nriver; head[2] = m->fbr; head[3] = m->surf_mode;
    head[4] = m->riv_mode; head[5] = device; head[6] = reorder; head[7] = 0;
    fwrite(head, sizeof(int32_t), 8, f);
    fwrite(&m->stepsize, sizeof(double), 1, f);
    fwrite(m->elem_f64, sizeof(double), PB_E_NCOL * ne, f);
    fwrite(m->elem_i32, sizeof(int32_t), PB_EI_NCOL * ne, f);
    fwrite(m->riv_f64, sizeof(double), PB_R_NCOL * nr, f);
    fwrite(m->riv_i32, sizeof(int32_t), PB_RI_NCOL * nr, f);
    fclose(f);
    exit(0);
}
