// cvode_b200.cu -- device-resident BDF/Newton/SPGMR integrator (placeholder,
// replaced by the full implementation in the next milestone).
#include "common.cuh"
extern "C" {
pihm_b200_cvode *pihm_b200_cvode_create(pihm_b200_ctx *) { pb::set_error("integrator not built yet"); return nullptr; }
void pihm_b200_cvode_destroy(pihm_b200_cvode *) {}
int pihm_b200_cvode_init(pihm_b200_cvode *, const pihm_b200_cvode_param *, double, const pihm_b200_vec *) { return -1; }
int pihm_b200_cvode_set_max_step(pihm_b200_cvode *, double) { return -1; }
int pihm_b200_cvode_solve(pihm_b200_cvode *, double, pihm_b200_vec *, double *) { return -1; }
int pihm_b200_cvode_get_stats(const pihm_b200_cvode *, pihm_b200_cvode_stats *) { return -1; }
int pihm_b200_adj_cvode_max_step(pihm_b200_cvode *, pihm_b200_maxstep_ctrl *) { return -1; }
}
