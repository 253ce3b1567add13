// tcgen05 3xTF32 dense layer (placeholder until the tensor-core path lands): reports "unsupported" so that
// api.cu falls back to the FP32 SIMT GEMM.
#include "dpe_internal.cuh"

namespace dpe {

int launch_gemm_tc(dpe_model *m, const GemmArgs &g, cudaStream_t s) {
    (void)m; (void)g; (void)s;
    return DPE_ERR_UNSUPPORTED;
}

}  // namespace dpe
