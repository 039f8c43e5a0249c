// Tensor-core (tcgen05 / TMEM / TMA) scoring path -- placeholder gate until the kernel lands.
#include "ocm_common.cuh"

extern "C" int ocm_simca_score_tc_supported(int64_t, int32_t, int64_t, int32_t, int32_t, const void*) { return 0; }

extern "C" int ocm_simca_score_tc_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                      const float*, const ocm_score_params*, float*, float*, uint8_t*, void*) {
    ocm::set_error("ocm_simca_score_tc_f32: shape not supported by the tensor-core path");
    return OCM_ELIMIT;
}
