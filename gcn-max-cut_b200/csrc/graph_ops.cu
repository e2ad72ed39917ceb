// graph_ops.cu -- per-batch graph preparation: degree norms, A_hat edge coefficients,
// dense padded adjacency rows (device-side graphExtender).  HBM-bound integer/float work.
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

// one warp per row: lanes stride over the row's edges
__global__ void edge_coef_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 const float* __restrict__ vals, const float* __restrict__ ns,
                                 const float* __restrict__ nd, int64_t n_rows, float* __restrict__ coef) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    float d = nd ? nd[row] : 1.0f;
    for (int e = e0 + lane; e < e1; e += 32) {
        float w = vals ? vals[e] : 1.0f;
        float s = ns ? __ldg(ns + colidx[e]) : 1.0f;
        coef[e] = (w * s) * d;
    }
}

__global__ void densify_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                               const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr,
                               int n_graphs, int64_t n_rows, int n_cols, float* __restrict__ X, int64_t ldx,
                               int* __restrict__ bad) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int g = find_graph(graph_ptr, n_graphs, row);
    int base = graph_ptr[g];
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    for (int e = e0 + lane; e < e1; e += 32) {
        int local = colidx[e] - base;
        if (local < 0 || local >= n_cols) { if (bad) *bad = 1; continue; }
        X[row * ldx + local] = vals ? vals[e] : 1.0f;
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
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * 32, threads);
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
    GMC_CUDA(cudaMemset2DAsync(X, (size_t)ldx * sizeof(float), 0, (size_t)n_cols * sizeof(float), (size_t)n_rows, s));
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * 32, threads);
    gmc::densify_kernel<<<(unsigned)blocks, threads, 0, s>>>(rowptr, colidx, vals, graph_ptr, n_graphs, n_rows,
                                                             n_cols, X, ldx, nullptr);
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
