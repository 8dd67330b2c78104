// fmm3.cu -- placeholder while the kd-tree FMM kernels are being brought up.
#include "common.cuh"
namespace nbco {
int fmm3_kd_launch(nbco_ctx *, float *, float *, int64_t, const float *, bool)
{ set_error("fmm3_kd: not built yet"); return NBCO_ERR_INVALID; }
void fmm3_destroy(nbco_ctx *) {}
}
extern "C" {
int nbco_fmm_get_info(nbco_ctx *, nbco_fmm_info *) { return NBCO_ERR_INVALID; }
int nbco_fmm_get_tree(nbco_ctx *, float *, float *, float *, float *, float *, int32_t *, int32_t *, int32_t *, int32_t *) { return NBCO_ERR_INVALID; }
int nbco_fmm_get_lists(nbco_ctx *, int32_t *, int64_t, int32_t *, int64_t) { return NBCO_ERR_INVALID; }
int nbco_fmm_get_phase_ms(nbco_ctx *, const char **, float *, int) { return 0; }
}
