// split.cu -- fp32-grade feature transforms at bf16 tensor-core rates (engine precision 'bf16x3').
//
// The reference computes GraphConv layer 1 in fp32 (TORCH_DTYPE, TrainingNeural.py:33-34, :80).  For the graphs it
// trains on, the pre-aggregated features A_hat X factor into a per-row scale and a matrix of small integers:
//   (A_hat X)[v, :] = s_v * sum_{u ~ v} X[u, :],   s_v = deg(v)^-1/2 deg(u)^-1/2  (one value per row when the neighbours
//   of v share a degree -- every regular graph),   X = the zero-padded 0/1 adjacency rows (graphExtender.py:106-111)
// so sum_u X[u, :] counts 2-step paths: integers <= max degree, EXACT in bf16.  The only inexact bf16 operand of
//   H1 = relu(s . (XI W1) + b1)        and        dW1 = XI^T (s . dH1pre)
// is then the fp32 one (W1, resp. s . dH1pre).  It is split into bf16 parts hi + lo (+ lo2) (8 mantissa bits each; three
// parts carry all 24 bits of fp32), the tcgen05 GEMM multiplies the exact integer tile with all parts in one MMA per
// k-step and adds the partial accumulators in its epilogue (gemm_tcgen05.cu, NS > 1).
//
//   gmc_f32_split_bf16     fp32 matrix -> n_split stacked bf16 parts
//   gmc_row_scale_f32      s_v from the per-edge coefficients, with a uniformity check
//   gmc_gemm_bf16_split    the GEMM (nn / tn), optional fp32 epilogue act(s_m * acc + bias_n)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gmc {

size_t tc_bf16_split_workspace_bytes(int op, int64_t M, int64_t N, int64_t K, int n_split, int with_projection);
int tc_gemm_bf16_split(int op, const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc, int n_split, int64_t b_split_rows, const float* row_scale,
                       const float* bias, int relu, const float* proj_w, float* proj_out, int64_t ldp, int proj_k,
                       int accumulate, int lo_shift, void* workspace, size_t workspace_bytes, cudaStream_t s);

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// one thread per 8 output columns (16 bytes per part); rows >= n_rows and columns >= n_cols are written as zeros
template <int NS>
__global__ void __launch_bounds__(256)
f32_split_bf16_kernel(const float* __restrict__ src, int64_t lds, __nv_bfloat16* __restrict__ dst, int64_t ldd,
                      int64_t n_rows, int n_cols, int64_t split_rows, int units_per_row) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = i / units_per_row;
    const int c0 = (int)(i - r * units_per_row) * 8;
    if (r >= split_rows) return;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (r < n_rows && c0 + j < n_cols) ? __ldg(src + r * lds + c0 + j) : 0.f;
#pragma unroll
    for (int sp = 0; sp < NS; ++sp) {
        __nv_bfloat16 part[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            part[j] = __float2bfloat16_rn(x[j]);
            x[j] -= __bfloat162float(part[j]);                     // exact: the residual has at most 16 significant bits
        }
        *reinterpret_cast<uint4*>(dst + ((int64_t)sp * split_rows + r) * ldd + c0) = *reinterpret_cast<const uint4*>(part);
    }
}

// fp16 parts: 11 significant bits each, so TWO parts carry 22 of fp32's 24 bits.  fp16's exponent range is narrow: the
// residual of a weight of magnitude 1e-2 is ~1e-5, a subnormal -- every part after the first is therefore stored
// multiplied by 2^(shift * part index) and the GEMM epilogue divides it out again.
template <int NS>
__global__ void __launch_bounds__(256)
f32_split_f16_kernel(const float* __restrict__ src, int64_t lds, __half* __restrict__ dst, int64_t ldd, int64_t n_rows,
                     int n_cols, int64_t split_rows, int units_per_row, float up) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = i / units_per_row;
    const int c0 = (int)(i - r * units_per_row) * 8;
    if (r >= split_rows) return;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (r < n_rows && c0 + j < n_cols) ? __ldg(src + r * lds + c0 + j) : 0.f;
#pragma unroll
    for (int sp = 0; sp < NS; ++sp) {
        __half part[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            part[j] = __float2half_rn(x[j]);
            x[j] = (x[j] - __half2float(part[j])) * up;            // exact residual, scaled by a power of two (exact)
        }
        *reinterpret_cast<uint4*>(dst + ((int64_t)sp * split_rows + r) * ldd + c0) = *reinterpret_cast<const uint4*>(part);
    }
}

__global__ void __launch_bounds__(256)
row_scale_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ coef, int64_t n_rows,
                 float* __restrict__ row_scale, int32_t* __restrict__ nonuniform) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
    float s = 0.f;
    bool bad = false;
    if (e1 > e0) {
        s = __ldg(coef + e0);
        for (int e = e0 + 1; e < e1; ++e) bad |= __ldg(coef + e) != s;
    }
    row_scale[v] = s;
    if (bad && nonuniform) atomicAdd(nonuniform, 1);
}

}  // namespace gmc

extern "C" {

int gmc_f32_split_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t n_rows, int32_t n_cols,
                       int32_t n_split, int64_t split_rows, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(src && dst, "gmc_f32_split_bf16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && lds >= n_cols && split_rows >= n_rows, "gmc_f32_split_bf16: bad sizes");
    GMC_REQUIRE(n_split >= 1 && n_split <= 3, "gmc_f32_split_bf16: n_split must be 1, 2 or 3");
    const int units = (n_cols + 7) / 8;
    GMC_REQUIRE(ldd % 8 == 0 && ldd >= (int64_t)units * 8 && aligned16(dst),
                "gmc_f32_split_bf16: dst needs a 16-byte aligned base and ldd %% 8 == 0 covering n_cols rounded up to 8");
    if (split_rows == 0) return GMC_OK;
    const int64_t total = split_rows * units;
    const unsigned grid = (unsigned)ceil_div<int64_t>(total, 256);
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
    cudaStream_t s = as_stream(stream);
    switch (n_split) {
        case 1: GMC_CUDA(launch_pdl(f32_split_bf16_kernel<1>, grid, 256, 0, s, src, lds, d, ldd, n_rows, n_cols, split_rows, units)); break;
        case 2: GMC_CUDA(launch_pdl(f32_split_bf16_kernel<2>, grid, 256, 0, s, src, lds, d, ldd, n_rows, n_cols, split_rows, units)); break;
        default: GMC_CUDA(launch_pdl(f32_split_bf16_kernel<3>, grid, 256, 0, s, src, lds, d, ldd, n_rows, n_cols, split_rows, units)); break;
    }
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_f32_split_f16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t n_rows, int32_t n_cols,
                      int32_t n_split, int64_t split_rows, int32_t lo_shift, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(src && dst, "gmc_f32_split_f16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && lds >= n_cols && split_rows >= n_rows, "gmc_f32_split_f16: bad sizes");
    GMC_REQUIRE(n_split >= 1 && n_split <= 3 && lo_shift >= 1 && lo_shift <= 24,
                "gmc_f32_split_f16: n_split must be 1..3 and lo_shift 1..24");
    const int units = (n_cols + 7) / 8;
    GMC_REQUIRE(ldd % 8 == 0 && ldd >= (int64_t)units * 8 && aligned16(dst),
                "gmc_f32_split_f16: dst needs a 16-byte aligned base and ldd %% 8 == 0 covering n_cols rounded up to 8");
    if (split_rows == 0) return GMC_OK;
    const int64_t total = split_rows * units;
    const unsigned grid = (unsigned)ceil_div<int64_t>(total, 256);
    __half* d = reinterpret_cast<__half*>(dst);
    const float up = (float)(1u << lo_shift);
    cudaStream_t s = as_stream(stream);
    switch (n_split) {
        case 1: GMC_CUDA(launch_pdl(f32_split_f16_kernel<1>, grid, 256, 0, s, src, lds, d, ldd, n_rows, n_cols, split_rows, units, up)); break;
        case 2: GMC_CUDA(launch_pdl(f32_split_f16_kernel<2>, grid, 256, 0, s, src, lds, d, ldd, n_rows, n_cols, split_rows, units, up)); break;
        default: GMC_CUDA(launch_pdl(f32_split_f16_kernel<3>, grid, 256, 0, s, src, lds, d, ldd, n_rows, n_cols, split_rows, units, up)); break;
    }
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_row_scale_f32(const int32_t* rowptr, const float* coef, int64_t n_rows, float* row_scale,
                      int32_t* nonuniform_count, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && coef && row_scale, "gmc_row_scale_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0, "gmc_row_scale_f32: negative row count");
    if (n_rows == 0) return GMC_OK;
    row_scale_kernel<<<(unsigned)ceil_div<int64_t>(n_rows, 256), 256, 0, as_stream(stream)>>>(rowptr, coef, n_rows, row_scale,
                                                                                              nonuniform_count);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

size_t gmc_gemm_bf16_split_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K, int32_t n_split,
                                           int32_t with_projection) {
    return gmc::tc_bf16_split_workspace_bytes(op, M, N, K, n_split, with_projection);
}

int gmc_gemm_bf16_split(int32_t op, const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                        int64_t ldb, int64_t ldc, int32_t n_split, int64_t b_split_rows, const float* row_scale,
                        const float* bias, int32_t relu, const float* proj_w, float* proj_out, int64_t ldp, int32_t n_proj,
                        int32_t accumulate, int32_t lo_shift, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(A && B && C, "gmc_gemm_bf16_split: null pointer");
    GMC_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gmc_gemm_bf16_split: negative dimension");
    const int64_t a_min = (op == 2) ? M : K;
    GMC_REQUIRE(lda >= a_min && ldb >= N && ldc >= N, "gmc_gemm_bf16_split: leading dimension too small (op %d)", op);
    return tc_gemm_bf16_split(op, A, B, C, M, N, K, lda, ldb, ldc, n_split, b_split_rows, row_scale, bias, relu, proj_w,
                              proj_out, ldp, n_proj, accumulate, lo_shift, workspace, workspace_bytes, as_stream(stream));
}

}  // extern "C"
