// spmm_fused.cu -- first GraphConv layer's aggregation fused with the skinny second-layer projection:
//
//     H[v,:] = relu( sum_e coef_e * T1[col_e,:] + b1 )            (written, needed by the backward pass)
//     T2[v,k] = sum_j H[v,j] * W2[j,k]                            (k < n_out <= 8)
//
// One warp owns a full output row (n_cols <= 512), so the projection is a register-resident epilogue:
// the second pass over H that gmc_skinny_fwd_f32 would make (8.2 GB at config 3) disappears.
// Replaces dgl update_all + bias (TrainingNeural.py:80), F.relu (:81) and th.matmul of conv2 (:83).
#include "common.cuh"

namespace gmc {

// Row schedule.  A CTA stages W^T once and then walks row blocks of 8 (one row per warp) grid-stride, so at any time
// the whole grid works on ONE contiguous window of gridDim x 8 rows (~5 graphs at config 3 = 10 MB of source rows,
// L2-resident).  The earlier blocked schedule (8 consecutive rows per warp, one CTA per 64 rows) kept ~47 graphs of
// rows in flight, overflowed L2 and re-read 40 % of the source rows from DRAM (5.64 MB/graph against 4.04 algorithmic).
template <int NV, int NOUT>
static int fused_ctas_per_sm(size_t smem);

template <int NV, int NOUT>
__global__ void __launch_bounds__(256)
spmm_fused_skinny_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                         const float* __restrict__ vals, const float* __restrict__ ns, const float* __restrict__ nd,
                         const float4* __restrict__ X, float4* __restrict__ Y, int64_t n_rows, int c4, int64_t ldx4,
                         int64_t ldy4, const float4* __restrict__ bias, int relu, const float* __restrict__ W,
                         float* __restrict__ T, int64_t ldt) {
    // Measured (bench_tf32_v5..v7): staging W^T in shared memory with 8 rows per warp costs 6.2 ms at config 3,
    // reading W through L1 with one row per warp 7.1 ms, the unfused pair of kernels 7.45 ms.
    extern __shared__ float4 Ws4[];                       // Ws[k][c4] : W^T, float4 over j
    {
        float* Ws = reinterpret_cast<float*>(Ws4);
        const int n_cols = c4 * 4;
        for (int i = threadIdx.x; i < n_cols * NOUT; i += blockDim.x) {
            const int k = i / n_cols, j = i - k * n_cols;
            Ws[i] = __ldg(W + (int64_t)j * NOUT + k);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps = blockDim.x >> 5;

    for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < n_rows; row += (int64_t)gridDim.x * warps) {
        float4 acc[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

        const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
        for (int eb = e0; eb < e1; eb += 32) {
            int my_c = 0;
            float my_a = 0.f;
            if (eb + lane < e1) {
                my_c = __ldg(colidx + eb + lane);
                my_a = vals ? __ldg(vals + eb + lane) : 1.0f;
                if (ns) my_a *= __ldg(ns + my_c);
            }
            const int cnt = min(32, e1 - eb);
            int j = 0;
            for (; j + 1 < cnt; j += 2) {
                const int ca = __shfl_sync(0xffffffffu, my_c, j), cb = __shfl_sync(0xffffffffu, my_c, j + 1);
                const float aa = __shfl_sync(0xffffffffu, my_a, j), ab = __shfl_sync(0xffffffffu, my_a, j + 1);
                const float4* xa = X + (int64_t)ca * ldx4;
                const float4* xb = X + (int64_t)cb * ldx4;
                float4 va[NV], vb[NV];
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const int col = lane + 32 * q;
                    if (col < c4) { va[q] = __ldg(xa + col); vb[q] = __ldg(xb + col); }
                    else { va[q] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q] = va[q]; }
                }
#pragma unroll
                for (int q = 0; q < NV; ++q) { fma4(acc[q], aa, va[q]); fma4(acc[q], ab, vb[q]); }
            }
            if (j < cnt) {
                const int ca = __shfl_sync(0xffffffffu, my_c, j);
                const float aa = __shfl_sync(0xffffffffu, my_a, j);
                const float4* xa = X + (int64_t)ca * ldx4;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const int col = lane + 32 * q;
                    if (col < c4) fma4(acc[q], aa, __ldg(xa + col));
                }
            }
        }

        const float d = nd ? __ldg(nd + row) : 1.0f;
        float proj[NOUT];
#pragma unroll
        for (int k = 0; k < NOUT; ++k) proj[k] = 0.f;
        float4* yr = Y + row * ldy4;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int col = lane + 32 * q;
            if (col < c4) {
                float4 r = acc[q];
                r.x *= d; r.y *= d; r.z *= d; r.w *= d;
                if (bias) { const float4 b = __ldg(bias + col); r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w; }
                if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                yr[col] = r;
#pragma unroll
                for (int k = 0; k < NOUT; ++k) {
                    const float4 w = Ws4[k * c4 + col];
                    proj[k] = fmaf(r.x, w.x, proj[k]); proj[k] = fmaf(r.y, w.y, proj[k]);
                    proj[k] = fmaf(r.z, w.z, proj[k]); proj[k] = fmaf(r.w, w.w, proj[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NOUT; ++k) proj[k] = warp_sum(proj[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NOUT; ++k) T[row * ldt + k] = proj[k];
        }
    }
}

// resident CTAs per SM of one instantiation (registers and the W^T staging decide): queried once
template <int NV, int NOUT>
static int fused_ctas_per_sm(size_t smem) {
    static int cached = 0;
    if (cached == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmm_fused_skinny_kernel<NV, NOUT>, 256, smem) != cudaSuccess || n < 1) {
            cudaGetLastError();
            n = 4;
        }
        cached = n;
    }
    return cached;
}

}  // namespace gmc

extern "C" int gmc_spmm_fused_skinny_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                                         const float* norm_src, const float* norm_dst, const float* X, float* Y,
                                         int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy, const float* bias,
                                         int32_t relu, const float* W, int32_t n_out, float* T, int64_t ldt,
                                         void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && colidx && X && Y && W && T, "gmc_spmm_fused_skinny_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols && ldt >= n_out,
                "gmc_spmm_fused_skinny_f32: bad sizes");
    GMC_REQUIRE(n_out >= 1 && n_out <= kMaxClasses, "gmc_spmm_fused_skinny_f32: n_out must be 1..8");
    GMC_REQUIRE(X != Y, "gmc_spmm_fused_skinny_f32: in-place SpMM is not supported");
    const bool ok = (n_cols % 4 == 0) && n_cols <= 512 && n_cols >= 16 && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                    aligned16(X) && aligned16(Y) && (!bias || aligned16(bias)) && aligned16(W);
    if (!ok) {
        set_error("gmc_spmm_fused_skinny_f32: needs 16 <= n_cols <= 512, n_cols %% 4 == 0 and 16-byte aligned rows; "
                  "use gmc_spmm_symnorm_f32 + gmc_skinny_fwd_f32");
        return GMC_ERR_UNSUPPORTED;
    }
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = as_stream(stream);
    const int c4 = n_cols / 4;
    const int warps = 8;
    const int64_t blocks = ceil_div<int64_t>(n_rows, warps);
    const size_t smem = (size_t)n_cols * n_out * sizeof(float);
    const float4* X4 = reinterpret_cast<const float4*>(X);
    float4* Y4 = reinterpret_cast<float4*>(Y);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#define GMC_LAUNCH(NV, K)                                                                                           \
    {                                                                                                               \
        const int64_t resident = (int64_t)sm_count() * fused_ctas_per_sm<NV, K>(smem);                              \
        const unsigned grid = (unsigned)(blocks < resident ? blocks : resident);                                    \
        spmm_fused_skinny_kernel<NV, K><<<grid, warps * 32, smem, s>>>(rowptr, colidx, vals, norm_src, norm_dst,    \
                                                                       X4, Y4, n_rows, c4, ldx / 4, ldy / 4, b4,   \
                                                                       relu, W, T, ldt);                            \
    }
#define GMC_NV(K)                                      \
    if (c4 <= 32) GMC_LAUNCH(1, K)                     \
    else if (c4 <= 64) GMC_LAUNCH(2, K)                \
    else GMC_LAUNCH(4, K)
    switch (n_out) {
        case 1: GMC_NV(1); break;
        case 2: GMC_NV(2); break;
        case 3: GMC_NV(3); break;
        case 4: GMC_NV(4); break;
        case 5: GMC_NV(5); break;
        case 6: GMC_NV(6); break;
        case 7: GMC_NV(7); break;
        default: GMC_NV(8); break;
    }
#undef GMC_NV
#undef GMC_LAUNCH
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}
