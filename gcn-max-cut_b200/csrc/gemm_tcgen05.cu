// gemm_tcgen05.cu -- tensor-core path of the dense feature transforms (placeholder until the
// tcgen05/TMA kernel lands: reports GMC_ERR_UNSUPPORTED, never silently falls back).
#include "common.cuh"

namespace gmc {

size_t tc_workspace_bytes(int, int64_t, int64_t, int64_t, int) { return 0; }

int tc_gemm(int, const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int,
            void*, size_t, cudaStream_t) {
    set_error("gmc_gemm: tcgen05 TF32 path not built in this library version");
    return GMC_ERR_UNSUPPORTED;
}

}  // namespace gmc
