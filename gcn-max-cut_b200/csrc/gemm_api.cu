// gemm_api.cu -- C-ABI dispatch of the dense feature transforms (include/gcnmaxcut.h, section b).
#include "common.cuh"

namespace gmc {
size_t simt_workspace_bytes(int64_t M, int64_t N, int64_t K);
int simt_gemm(int op, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
              int64_t ldb, int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t s);
size_t tc_workspace_bytes(int op, int64_t M, int64_t N, int64_t K, int precision);
int tc_gemm(int op, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
            int64_t ldb, int64_t ldc, int accumulate, int precision, void* workspace, size_t workspace_bytes,
            cudaStream_t s);

size_t tc_bf16_workspace_bytes(int op, int64_t M, int64_t N, int64_t K);
int tc_gemm_bf16(int op, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                 int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t s, int c_bf16,
                 const float* bias, int relu, const float* proj_w, float* proj_out, int64_t ldp, int proj_k);

static int gemm_dispatch(int op, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                         int64_t lda, int64_t ldb, int64_t ldc, int accumulate, int precision, void* workspace,
                         size_t workspace_bytes, void* stream) {
    GMC_REQUIRE(A && B && C, "gmc_gemm: null pointer");
    GMC_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gmc_gemm: negative dimension");
    const int64_t a_min = (op == 2) ? M : K, b_min = (op == 1) ? K : N;
    GMC_REQUIRE(lda >= a_min && ldb >= b_min && ldc >= N, "gmc_gemm: leading dimension too small (op %d)", op);
    cudaStream_t s = as_stream(stream);
    if (precision == GMC_GEMM_FP32)
        return simt_gemm(op, A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
    if (precision == GMC_GEMM_TF32 || precision == GMC_GEMM_TF32X3) {
        // Skinny problems (the 3-class layer: a dimension below one 16-byte row) cannot be described to TMA; they take
        // the exact CUDA-core kernel -- never a lower precision than asked for.  Anything else that TMA cannot address
        // (unaligned base / leading dimension) is still rejected by the tensor-core path with an error.
        if (N < 4 || M < 4 || K < 4)
            return simt_gemm(op, A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
        return tc_gemm(op, A, B, C, M, N, K, lda, ldb, ldc, accumulate, precision, workspace, workspace_bytes, s);
    }
    set_error("gmc_gemm: unknown precision %d", precision);
    return GMC_ERR_INVALID_ARG;
}
}  // namespace gmc

extern "C" {

size_t gmc_gemm_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K, int32_t precision) {
    const size_t simt = gmc::simt_workspace_bytes(M, N, K);
    if (precision == GMC_GEMM_FP32) return simt;
    const size_t tc = gmc::tc_workspace_bytes(op, M, N, K, precision);
    return tc > simt ? tc : simt;            // skinny shapes are routed to the CUDA-core kernel (gemm_dispatch)
}

int gmc_gemm_nn(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                int64_t ldc, int32_t accumulate, int32_t precision, void* workspace, size_t workspace_bytes,
                void* stream) {
    return gmc::gemm_dispatch(0, A, B, C, M, N, K, lda, ldb, ldc, accumulate, precision, workspace, workspace_bytes, stream);
}
int gmc_gemm_nt(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                int64_t ldc, int32_t accumulate, int32_t precision, void* workspace, size_t workspace_bytes,
                void* stream) {
    return gmc::gemm_dispatch(1, A, B, C, M, N, K, lda, ldb, ldc, accumulate, precision, workspace, workspace_bytes, stream);
}
int gmc_gemm_tn(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                int64_t ldc, int32_t accumulate, int32_t precision, void* workspace, size_t workspace_bytes,
                void* stream) {
    return gmc::gemm_dispatch(2, A, B, C, M, N, K, lda, ldb, ldc, accumulate, precision, workspace, workspace_bytes, stream);
}

size_t gmc_gemm_bf16_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K) {
    return gmc::tc_bf16_workspace_bytes(op, M, N, K);
}

// C[M,N] (+)= op(A) op(B) with bf16 operands in device memory (op: 0 nn, 1 nt, 2 tn as gmc_gemm_*), fp32 out.
int gmc_gemm_bf16(int32_t op, const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                  int64_t ldb, int64_t ldc, int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(A && B && C, "gmc_gemm_bf16: null pointer");
    GMC_REQUIRE(op >= 0 && op <= 2 && M >= 0 && N >= 0 && K >= 0, "gmc_gemm_bf16: bad op or negative dimension");
    const int64_t a_min = (op == 2) ? M : K, b_min = (op == 1) ? K : N;
    GMC_REQUIRE(lda >= a_min && ldb >= b_min && ldc >= N, "gmc_gemm_bf16: leading dimension too small (op %d)", op);
    return tc_gemm_bf16(op, A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, as_stream(stream), 0, nullptr, 0, nullptr, nullptr, 0, 0);
}

// C[M,N] = op(A) op(B), bf16 operands AND a bf16 result: the fp32 accumulators are rounded to nearest even in the
// epilogue and leave through TMA stores, so the activation is written once at half the bytes (T1 = X W1 of
// TrainingNeural.py:80 when the layer-1 activations are kept in bf16).  Optional fp32 bias[N] (N % 4 == 0) and ReLU
// applied to the accumulators before the rounding: with pre-aggregated features the GEMM IS the whole first layer,
// H1 = relu((A_hat X) W1 + b1) (TrainingNeural.py:80-81).  Optional fused skinny projection (proj_w != NULL): P = bf16(C)
// projW with projW [N rounded up to 64][4] fp32, zero padded (the second GraphConv layer's th.matmul, :83, on the rounded
// activations the backward pass will read); P [M, ldp] is zeroed here and receives n_proj <= 4 columns; needs N <= 512
// (at most two n-tiles per row, so the atomic partial sums commute: reproducible).  No split-K, no accumulate.
int gmc_gemm_bf16_bf16out(int32_t op, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                          int64_t ldb, int64_t ldc, const float* bias, int32_t relu, const float* proj_w, float* proj_out,
                          int64_t ldp, int32_t n_proj, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(A && B && C, "gmc_gemm_bf16_bf16out: null pointer");
    GMC_REQUIRE(op >= 0 && op <= 2 && M >= 0 && N >= 0 && K >= 0, "gmc_gemm_bf16_bf16out: bad op or negative dimension");
    const int64_t a_min = (op == 2) ? M : K, b_min = (op == 1) ? K : N;
    GMC_REQUIRE(lda >= a_min && ldb >= b_min && ldc >= N, "gmc_gemm_bf16_bf16out: leading dimension too small (op %d)", op);
    GMC_REQUIRE(ldc % 8 == 0 && aligned16(C), "gmc_gemm_bf16_bf16out: C needs a 16-byte aligned base and ldc %% 8 == 0");
    return tc_gemm_bf16(op, A, B, C, M, N, K, lda, ldb, ldc, 0, nullptr, 0, as_stream(stream), 1, bias, relu, proj_w, proj_out,
                        ldp, n_proj);
}

}  // extern "C"
