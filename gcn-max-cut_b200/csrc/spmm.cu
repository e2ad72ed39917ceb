// spmm.cu -- (a) symmetric-normalised CSR SpMM, forward == backward (A_hat symmetric).
//
//   Y[v,:] = act( nd[v] * sum_{e in row v} w_e * ns[col_e] * X[col_e,:] + bias )
//
// Replaces dgl update_all(copy_u,sum) + both degree norms + bias inside GraphConv.forward
// (reference python/Training/TrainingNeural.py:80, :83) and F.relu (:81).
//
// HBM-bound gather.  Block-diagonal batches of small graphs keep a graph's source rows
// (n*C*4 B = 2 MB at n=1000, C=500) L2-resident, so the algorithmic traffic is one read of X,
// one write of Y and the CSR:  bytes = 8*N*C + 4*nnz + 4*(N+1)   (SURVEY.md 8(d)).
//
// Kernels
//   spmm_warp_row_v4 : one warp per row, lanes own float4 column slots (128-bit loads), the
//                      row's (col, coef) pairs are loaded once by the lanes and broadcast with
//                      shuffles; edges are processed two at a time for memory-level parallelism.
//   spmm_thread_row  : tiny C (the 3-class layer): one thread per row, scalar loads.
//   spmm_warp_row_s  : generic fallback (C % 4 != 0 or unaligned rows).
#include "common.cuh"

namespace gmc {

template <int NV>
__global__ void __launch_bounds__(256)
spmm_warp_row_v4(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                 const float* __restrict__ vals, const float* __restrict__ ns, const float* __restrict__ nd,
                 const float4* __restrict__ X, float4* __restrict__ Y, int64_t n_rows, int c4, int64_t ldx4,
                 int64_t ldy4, const float4* __restrict__ bias, int relu) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int cbase = blockIdx.y * (NV * 32);            // column-chunk (float4 units) of this CTA row

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
        for (; j + 1 < cnt; j += 2) {                     // two neighbours in flight
            const int ca = __shfl_sync(0xffffffffu, my_c, j), cb = __shfl_sync(0xffffffffu, my_c, j + 1);
            const float aa = __shfl_sync(0xffffffffu, my_a, j), ab = __shfl_sync(0xffffffffu, my_a, j + 1);
            const float4* xa = X + (int64_t)ca * ldx4 + cbase;
            const float4* xb = X + (int64_t)cb * ldx4 + cbase;
            float4 va[NV], vb[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int col = lane + 32 * q;
                if (cbase + col < c4) { va[q] = __ldg(xa + col); vb[q] = __ldg(xb + col); }
                else { va[q] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q] = va[q]; }
            }
#pragma unroll
            for (int q = 0; q < NV; ++q) { fma4(acc[q], aa, va[q]); fma4(acc[q], ab, vb[q]); }
        }
        if (j < cnt) {
            const int ca = __shfl_sync(0xffffffffu, my_c, j);
            const float aa = __shfl_sync(0xffffffffu, my_a, j);
            const float4* xa = X + (int64_t)ca * ldx4 + cbase;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int col = lane + 32 * q;
                if (cbase + col < c4) fma4(acc[q], aa, __ldg(xa + col));
            }
        }
    }

    const float d = nd ? __ldg(nd + row) : 1.0f;
    float4* yr = Y + row * ldy4 + cbase;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int col = lane + 32 * q;
        if (cbase + col < c4) {
            float4 r = acc[q];
            r.x *= d; r.y *= d; r.z *= d; r.w *= d;
            if (bias) { const float4 b = __ldg(bias + cbase + col); r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w; }
            if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            yr[col] = r;
        }
    }
}

template <int C>
__global__ void __launch_bounds__(256)
spmm_thread_row(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const float* __restrict__ vals, const float* __restrict__ ns, const float* __restrict__ nd,
                const float* __restrict__ X, float* __restrict__ Y, int64_t n_rows, int64_t ldx, int64_t ldy,
                const float* __restrict__ bias, int relu) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    float acc[C];
#pragma unroll
    for (int k = 0; k < C; ++k) acc[k] = 0.f;
    const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
    for (int e = e0; e < e1; ++e) {
        const int c = __ldg(colidx + e);
        float a = vals ? __ldg(vals + e) : 1.0f;
        if (ns) a *= __ldg(ns + c);
        const float* xr = X + (int64_t)c * ldx;
#pragma unroll
        for (int k = 0; k < C; ++k) acc[k] = fmaf(a, __ldg(xr + k), acc[k]);
    }
    const float d = nd ? __ldg(nd + row) : 1.0f;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        float r = acc[k] * d;
        if (bias) r += __ldg(bias + k);
        if (relu) r = fmaxf(r, 0.f);
        Y[row * ldy + k] = r;
    }
}

__global__ void __launch_bounds__(256)
spmm_warp_row_s(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const float* __restrict__ vals, const float* __restrict__ ns, const float* __restrict__ nd,
                const float* __restrict__ X, float* __restrict__ Y, int64_t n_rows, int n_cols, int64_t ldx,
                int64_t ldy, const float* __restrict__ bias, int relu) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
    const float d = nd ? __ldg(nd + row) : 1.0f;
    for (int cb = 0; cb < n_cols; cb += 32) {
        const int col = cb + lane;
        float acc = 0.f;
        for (int e = e0; e < e1; ++e) {
            const int c = __ldg(colidx + e);
            float a = vals ? __ldg(vals + e) : 1.0f;
            if (ns) a *= __ldg(ns + c);
            if (col < n_cols) acc = fmaf(a, __ldg(X + (int64_t)c * ldx + col), acc);
        }
        if (col < n_cols) {
            float r = acc * d;
            if (bias) r += __ldg(bias + col);
            if (relu) r = fmaxf(r, 0.f);
            Y[row * ldy + col] = r;
        }
    }
}

}  // namespace gmc

extern "C" int gmc_spmm_symnorm_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                                    const float* norm_src, const float* norm_dst, const float* X, float* Y,
                                    int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy, const float* bias,
                                    int32_t relu, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && colidx && X && Y, "gmc_spmm_symnorm_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols, "gmc_spmm_symnorm_f32: bad sizes");
    GMC_REQUIRE(X != Y, "gmc_spmm_symnorm_f32: in-place SpMM is not supported");
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = as_stream(stream);
    const bool vec_ok = (n_cols % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) &&
                        (!bias || aligned16(bias)) && n_cols >= 16;
    if (vec_ok) {
        const int c4 = n_cols / 4;
        const int warps = 8;
        dim3 block(warps * 32);
        const float4* X4 = reinterpret_cast<const float4*>(X);
        float4* Y4 = reinterpret_cast<float4*>(Y);
        const float4* b4 = reinterpret_cast<const float4*>(bias);
#define GMC_SPMM_LAUNCH(NV)                                                                              \
    {                                                                                                    \
        dim3 grid((unsigned)ceil_div<int64_t>(n_rows, warps), (unsigned)ceil_div(c4, (NV) * 32));        \
        spmm_warp_row_v4<NV><<<grid, block, 0, s>>>(rowptr, colidx, vals, norm_src, norm_dst, X4, Y4,    \
                                                    n_rows, c4, ldx / 4, ldy / 4, b4, relu);             \
    }
        if (c4 <= 32) GMC_SPMM_LAUNCH(1)
        else if (c4 <= 64) GMC_SPMM_LAUNCH(2)
        else GMC_SPMM_LAUNCH(4)
#undef GMC_SPMM_LAUNCH
    } else if (n_cols <= 4) {
        const int threads = 256;
        const unsigned blocks = (unsigned)ceil_div<int64_t>(n_rows, threads);
        switch (n_cols) {
            case 1: spmm_thread_row<1><<<blocks, threads, 0, s>>>(rowptr, colidx, vals, norm_src, norm_dst, X, Y, n_rows, ldx, ldy, bias, relu); break;
            case 2: spmm_thread_row<2><<<blocks, threads, 0, s>>>(rowptr, colidx, vals, norm_src, norm_dst, X, Y, n_rows, ldx, ldy, bias, relu); break;
            case 3: spmm_thread_row<3><<<blocks, threads, 0, s>>>(rowptr, colidx, vals, norm_src, norm_dst, X, Y, n_rows, ldx, ldy, bias, relu); break;
            default: spmm_thread_row<4><<<blocks, threads, 0, s>>>(rowptr, colidx, vals, norm_src, norm_dst, X, Y, n_rows, ldx, ldy, bias, relu); break;
        }
    } else {
        const int warps = 8;
        spmm_warp_row_s<<<(unsigned)ceil_div<int64_t>(n_rows, warps), warps * 32, 0, s>>>(
            rowptr, colidx, vals, norm_src, norm_dst, X, Y, n_rows, n_cols, ldx, ldy, bias, relu);
    }
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}
