// graph_ops.cu -- per-batch graph preparation: degree norms, A_hat edge coefficients,
// dense padded adjacency rows (device-side graphExtender).  HBM-bound integer/float work.
#include <cuda_bf16.h>

#include "common.cuh"

namespace gmc {

__global__ void degree_norm_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows,
                                   float* __restrict__ norm, int32_t* __restrict__ zero_count) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int is_zero = 0;
    if (v < n_rows) {
        int deg = rowptr[v + 1] - rowptr[v];
        is_zero = (deg == 0);
        // DGL: degs.clamp(min=1) then pow(-0.5)
        norm[v] = 1.0f / sqrtf((float)(deg < 1 ? 1 : deg));
    }
    if (zero_count) {
        unsigned m = __ballot_sync(0xffffffffu, is_zero);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(zero_count, __popc(m));
    }
}

// Rows of the reference's graphs have 6-8 edges: a warp per row left 3/4 of the lanes idle and cost 4 M warps per
// launch at config 3 (0.62 ms for edge_coef, 1.2 ms for the feature scatter inside the end-to-end step).  Eight
// lanes own a row (four rows per warp) and stride over its edges.
constexpr int kRowLanes = 8;

// graph of a row: the batches the reference builds have equal-sized graphs, so row / size(graph 0) is verified with two
// loads before falling back to the binary search
__device__ __forceinline__ int find_graph_guess(const int32_t* __restrict__ graph_ptr, int n_graphs, int64_t row) {
    const int n0 = __ldg(graph_ptr + 1) - __ldg(graph_ptr);
    if (n0 > 0) {
        const int64_t g = row / n0;
        if (g < n_graphs && (int64_t)__ldg(graph_ptr + g) <= row && row < (int64_t)__ldg(graph_ptr + g + 1)) return (int)g;
    }
    return find_graph(graph_ptr, n_graphs, row);
}

__global__ void edge_coef_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 const float* __restrict__ vals, const float* __restrict__ ns,
                                 const float* __restrict__ nd, int64_t n_rows, float* __restrict__ coef) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    float d = nd ? nd[row] : 1.0f;
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        float w = vals ? vals[e] : 1.0f;
        float s = ns ? __ldg(ns + colidx[e]) : 1.0f;
        coef[e] = (w * s) * d;
    }
}

__global__ void densify_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                               const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr,
                               int n_graphs, int64_t n_rows, int n_cols, float* __restrict__ X, int64_t ldx,
                               int* __restrict__ bad) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int g = find_graph_guess(graph_ptr, n_graphs, row);
    int base = graph_ptr[g];
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        int local = colidx[e] - base;
        if (local < 0 || local >= n_cols) { if (bad) *bad = 1; continue; }
        X[row * ldx + local] = vals ? vals[e] : 1.0f;
    }
}

__global__ void densify_bf16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, int n_graphs,
                                    int64_t n_rows, int n_cols, __nv_bfloat16* __restrict__ X, int64_t ldx) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int g = find_graph_guess(graph_ptr, n_graphs, row);
    int base = graph_ptr[g];
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        int local = colidx[e] - base;
        if (local < 0 || local >= n_cols) continue;
        X[row * ldx + local] = __float2bfloat16_rn(vals ? vals[e] : 1.0f);
    }
}

// value != 0: X[row, local col] = bf16(vals ? vals[e] : 1) at every edge; value == 0: zeros at the same positions
__global__ void scatter_bf16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, int n_graphs,
                                    int64_t n_rows, int n_cols, __nv_bfloat16* __restrict__ X, int64_t ldx, int clear) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int g = find_graph_guess(graph_ptr, n_graphs, row);
    int base = graph_ptr[g];
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        int local = colidx[e] - base;
        if (local < 0 || local >= n_cols) continue;
        X[row * ldx + local] = __float2bfloat16_rn(clear ? 0.0f : (vals ? vals[e] : 1.0f));
    }
}

// 8 elements per thread: two 128-bit loads, one 128-bit store
__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ src, int64_t lds, __nv_bfloat16* __restrict__ dst, int64_t ldd, int64_t n_rows,
                   int n_cols) {
    const int c8 = (n_cols + 7) / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = i / c8;
    const int c = (int)(i - r * c8) * 8;
    if (r >= n_rows) return;
    const float* s = src + r * lds + c;
    __nv_bfloat16* d = dst + r * ldd + c;
    if (c + 8 <= n_cols && ((lds | ldd) % 8 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
        (reinterpret_cast<uintptr_t>(dst) % 16 == 0)) {
        const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
        __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                               __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
        *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(o);
    } else {
        for (int k = 0; k < 8 && c + k < n_cols; ++k) d[k] = __float2bfloat16_rn(s[k]);
    }
}

}  // namespace gmc

extern "C" {

int gmc_degree_norm_f32(const int32_t* rowptr, int64_t n_rows, float* norm, int32_t* zero_degree_count,
                        void* stream) {
    GMC_REQUIRE(rowptr && norm && n_rows >= 0, "gmc_degree_norm_f32: null pointer or negative size");
    cudaStream_t s = gmc::as_stream(stream);
    if (zero_degree_count) GMC_CUDA(cudaMemsetAsync(zero_degree_count, 0, sizeof(int32_t), s));
    if (n_rows == 0) return GMC_OK;
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows, threads);
    gmc::degree_norm_kernel<<<(unsigned)blocks, threads, 0, s>>>(rowptr, n_rows, norm, zero_degree_count);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_edge_coef_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, const float* norm_src,
                      const float* norm_dst, int64_t n_rows, float* coef, void* stream) {
    GMC_REQUIRE(rowptr && colidx && coef && n_rows >= 0, "gmc_edge_coef_f32: null pointer or negative size");
    if (n_rows == 0) return GMC_OK;
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::edge_coef_kernel<<<(unsigned)blocks, threads, 0, gmc::as_stream(stream)>>>(rowptr, colidx, vals, norm_src,
                                                                                    norm_dst, n_rows, coef);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_csr_densify_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* graph_ptr,
                        int32_t n_graphs, int64_t n_rows, int32_t n_cols, float* X, int64_t ldx, void* stream) {
    GMC_REQUIRE(rowptr && colidx && graph_ptr && X, "gmc_csr_densify_f32: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_csr_densify_f32: bad sizes");
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = gmc::as_stream(stream);
    // rows are stored back to back: one flat memset (pad columns included) runs at write bandwidth, the pitched 2-D
    // form does not
    GMC_CUDA(cudaMemsetAsync(X, 0, (size_t)ldx * sizeof(float) * (size_t)(n_rows - 1) + (size_t)n_cols * sizeof(float), s));
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::densify_kernel<<<(unsigned)blocks, threads, 0, s>>>(rowptr, colidx, vals, graph_ptr, n_graphs, n_rows,
                                                             n_cols, X, ldx, nullptr);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// bf16 twin of gmc_csr_densify_f32 (0/1 adjacency entries are exact in bf16): the features of the bf16 GEMM path
int gmc_csr_densify_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* graph_ptr,
                         int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X, int64_t ldx, void* stream) {
    GMC_REQUIRE(rowptr && colidx && graph_ptr && X, "gmc_csr_densify_bf16: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_csr_densify_bf16: bad sizes");
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = gmc::as_stream(stream);
    GMC_CUDA(cudaMemsetAsync(X, 0, (size_t)ldx * 2 * (size_t)(n_rows - 1) + (size_t)n_cols * 2, s));
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::densify_bf16_kernel<<<(unsigned)blocks, threads, 0, s>>>(rowptr, colidx, vals, graph_ptr, n_graphs, n_rows, n_cols,
                                                                  reinterpret_cast<__nv_bfloat16*>(X), ldx);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// Incremental form of gmc_csr_densify_bf16 for a feature buffer that is reused step after step: clear = 0 writes the
// adjacency entries of the batch into an X that is zero everywhere (no memset), clear = 1 writes zeros at the same
// positions, restoring the all-zero state.  Touches nnz 32-byte sectors instead of the whole n_rows x ldx matrix
// (config 3: 0.9 GB instead of 8.4 GB) -- graphExtender.py:106-111 for a stream of graphs.
int gmc_csr_scatter_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* graph_ptr,
                         int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X, int64_t ldx, int32_t clear,
                         void* stream) {
    GMC_REQUIRE(rowptr && colidx && graph_ptr && X, "gmc_csr_scatter_bf16: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_csr_scatter_bf16: bad sizes");
    if (n_rows == 0) return GMC_OK;
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::scatter_bf16_kernel<<<(unsigned)blocks, threads, 0, gmc::as_stream(stream)>>>(
        rowptr, colidx, vals, graph_ptr, n_graphs, n_rows, n_cols, reinterpret_cast<__nv_bfloat16*>(X), ldx, clear);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// dst[r, c] = bf16(src[r, c]), round to nearest even; leading dimensions in elements
int gmc_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t n_rows, int32_t n_cols, void* stream) {
    GMC_REQUIRE(src && dst, "gmc_f32_to_bf16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols >= 0 && lds >= n_cols && ldd >= n_cols, "gmc_f32_to_bf16: bad sizes");
    if (n_rows == 0 || n_cols == 0) return GMC_OK;
    const int64_t total = n_rows * ((n_cols + 7) / 8);
    gmc::f32_to_bf16_kernel<<<(unsigned)gmc::ceil_div<int64_t>(total, 256), 256, 0, gmc::as_stream(stream)>>>(
        src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, n_rows, n_cols);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_copy2d_f32(float* dst, int64_t lddst, const float* src, int64_t ldsrc, int64_t n_rows, int32_t n_cols,
                   void* stream) {
    GMC_REQUIRE(dst && src, "gmc_copy2d_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols >= 0 && lddst >= n_cols && ldsrc >= n_cols, "gmc_copy2d_f32: bad sizes");
    if (n_rows == 0 || n_cols == 0) return GMC_OK;
    GMC_CUDA(cudaMemcpy2DAsync(dst, (size_t)lddst * sizeof(float), src, (size_t)ldsrc * sizeof(float),
                               (size_t)n_cols * sizeof(float), (size_t)n_rows, cudaMemcpyDeviceToDevice,
                               gmc::as_stream(stream)));
    return GMC_OK;
}

}  // extern "C"
