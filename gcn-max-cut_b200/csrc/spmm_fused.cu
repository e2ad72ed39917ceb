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

// ---- bf16 activations ---------------------------------------------------------------------------------------
// Same row schedule with the source matrix stored in bf16 (T1 written by the bf16-output GEMM epilogue): a neighbour
// row is n_cols * 2 bytes, so the d gathers per output row move half the L2->SM bytes -- the roof of the fp32 kernel
// (profiles/r01_spmm_slab_notes.md).  Accumulation, bias, ReLU and the projection stay fp32; H is written as fp32 or,
// with YB16, rounded to bf16 (read back by gmc_skinny_bwd_bf16).  One lane owns 8 consecutive columns (one 16-byte
// load) per chunk; four neighbour rows are in flight per lane.
__device__ __forceinline__ void fma8_bf16(float (&acc)[8], float a, const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fmaf(a, __uint_as_float(w[i] << 16), acc[2 * i]);
        acc[2 * i + 1] = fmaf(a, __uint_as_float(w[i] & 0xffff0000u), acc[2 * i + 1]);
    }
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    uint32_t w;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    return w;
}

template <int NV, int NOUT, bool YB16>
__global__ void __launch_bounds__(256, 3)
spmm_fused_skinny_b16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                             const float* __restrict__ vals, const float* __restrict__ ns, const float* __restrict__ nd,
                             const uint4* __restrict__ X, void* __restrict__ Yv, int64_t n_rows, int n_cols, int c8,
                             int64_t ldx8, int64_t ldy, const float* __restrict__ bias, int relu,
                             const float* __restrict__ W, float* __restrict__ T, int64_t ldt) {
    extern __shared__ float4 Ws4[];                       // Ws[k][c8 * 8] : W^T, zero beyond n_cols
    const int wcols = c8 * 8;
    {
        float* Ws = reinterpret_cast<float*>(Ws4);
        for (int i = threadIdx.x; i < wcols * NOUT; i += blockDim.x) {
            const int k = i / wcols, j = i - k * wcols;
            Ws[i] = j < n_cols ? __ldg(W + (int64_t)j * NOUT + k) : 0.f;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps = blockDim.x >> 5;

    // Software pipeline over the warp's rows: a row costs three dependent L2 round trips (rowptr -> colidx/vals ->
    // neighbour rows); with them in sequence the gathers -- the only loads that move real bytes -- were in flight a
    // third of the time (4.97 ms at config 3, 5.8 TB/s of L2->SM traffic, far below the L2 roof).  Row r's neighbour
    // list is therefore fetched while row r-1 is gathered, and row r+1's extent one step earlier still.
    const int64_t stride = (int64_t)gridDim.x * warps;
    int64_t row = (int64_t)blockIdx.x * warps + warp;
    int e0 = 0, e1 = 0, e0n = 0, e1n = 0, my_c = 0;
    float my_a = 0.f;
    auto load_list = [&](int b0, int b1, int& c, float& a) {
        c = 0; a = 0.f;
        if (b0 + lane < b1) {
            c = __ldg(colidx + b0 + lane);
            a = vals ? __ldg(vals + b0 + lane) : 1.0f;
            if (ns) a *= __ldg(ns + c);
        }
    };
    if (row < n_rows) {
        e0 = __ldg(rowptr + row); e1 = __ldg(rowptr + row + 1);
        load_list(e0, e1, my_c, my_a);
        if (row + stride < n_rows) { e0n = __ldg(rowptr + row + stride); e1n = __ldg(rowptr + row + stride + 1); }
    }
    for (; row < n_rows; row += stride) {
        float acc[NV][8];
#pragma unroll
        for (int q = 0; q < NV; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;

        // next row's list and the extent of the row after it: issued before this row's gathers
        int nx_c = 0, e0nn = 0, e1nn = 0;
        float nx_a = 0.f;
        if (row + stride < n_rows) load_list(e0n, e1n, nx_c, nx_a);
        if (row + 2 * stride < n_rows) { e0nn = __ldg(rowptr + row + 2 * stride); e1nn = __ldg(rowptr + row + 2 * stride + 1); }

        for (int eb = e0; eb < e1; eb += 32) {
            if (eb > e0) load_list(eb, e1, my_c, my_a);   // rows with more than 32 neighbours: later chunks on demand
            const int cnt = min(32, e1 - eb);
            int j = 0;
            for (; j + 3 < cnt; j += 4) {
                uint4 v[4][NV];
                float a[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int c = __shfl_sync(0xffffffffu, my_c, j + g);
                    a[g] = __shfl_sync(0xffffffffu, my_a, j + g);
                    const uint4* x = X + (int64_t)c * ldx8;
#pragma unroll
                    for (int q = 0; q < NV; ++q) {
                        const int col = lane + 32 * q;
                        v[g][q] = col < c8 ? __ldg(x + col) : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int q = 0; q < NV; ++q) fma8_bf16(acc[q], a[g], v[g][q]);
            }
            if (j < cnt) {                                // 1..3 remaining neighbours, all in flight together
                uint4 v[3][NV];
                float a[3];
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const bool on = j + g < cnt;          // warp-uniform
                    const int c = __shfl_sync(0xffffffffu, my_c, on ? j + g : j);
                    a[g] = on ? __shfl_sync(0xffffffffu, my_a, on ? j + g : j) : 0.f;
                    const uint4* x = X + (int64_t)c * ldx8;
#pragma unroll
                    for (int q = 0; q < NV; ++q) {
                        const int col = lane + 32 * q;
                        v[g][q] = (on && col < c8) ? __ldg(x + col) : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int q = 0; q < NV; ++q) fma8_bf16(acc[q], a[g], v[g][q]);
            }
        }
        e0 = e0n; e1 = e1n; e0n = e0nn; e1n = e1nn; my_c = nx_c; my_a = nx_a;

        const float d = nd ? __ldg(nd + row) : 1.0f;
        float proj[NOUT];
#pragma unroll
        for (int k = 0; k < NOUT; ++k) proj[k] = 0.f;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int col = lane + 32 * q;                // 8-column unit
            if (col < c8) {
                const bool hi_ok = col * 8 + 4 < n_cols;  // n_cols % 4 == 0: a unit is whole or holds 4 valid columns
                float r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = acc[q][i] * d;
                if (bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + col * 2);
                    r[0] += b0.x; r[1] += b0.y; r[2] += b0.z; r[3] += b0.w;
                    if (hi_ok) {
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + col * 2 + 1);
                        r[4] += b1.x; r[5] += b1.y; r[6] += b1.z; r[7] += b1.w;
                    }
                }
                if (relu) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) r[i] = fmaxf(r[i], 0.f);
                }
                if (!hi_ok) { r[4] = 0.f; r[5] = 0.f; r[6] = 0.f; r[7] = 0.f; }   // pad columns: exact zeros, never garbage
                if (YB16) {
                    // the projection consumes what the backward pass will read: the rounded activations
                    uint4 o = make_uint4(pack2_bf16(r[0], r[1]), pack2_bf16(r[2], r[3]), pack2_bf16(r[4], r[5]), pack2_bf16(r[6], r[7]));
                    uint4* yr = reinterpret_cast<uint4*>(Yv) + row * (ldy / 8);
                    if (col * 8 + 8 <= ldy) yr[col] = o;
                    else *reinterpret_cast<uint2*>(yr + col) = make_uint2(o.x, o.y);
                    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) { r[2 * i] = __uint_as_float(w[i] << 16); r[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
                } else {
                    float4* yr = reinterpret_cast<float4*>(Yv) + row * (ldy / 4);
                    yr[col * 2] = make_float4(r[0], r[1], r[2], r[3]);
                    if (hi_ok) yr[col * 2 + 1] = make_float4(r[4], r[5], r[6], r[7]);
                }
#pragma unroll
                for (int k = 0; k < NOUT; ++k) {
                    const float4 w0 = Ws4[k * (c8 * 2) + col * 2], w1 = Ws4[k * (c8 * 2) + col * 2 + 1];
                    float pk = proj[k];
                    pk = fmaf(r[0], w0.x, pk); pk = fmaf(r[1], w0.y, pk); pk = fmaf(r[2], w0.z, pk); pk = fmaf(r[3], w0.w, pk);
                    pk = fmaf(r[4], w1.x, pk); pk = fmaf(r[5], w1.y, pk); pk = fmaf(r[6], w1.z, pk); pk = fmaf(r[7], w1.w, pk);
                    proj[k] = pk;
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

template <int NV, int NOUT, bool YB16>
static int fused_b16_ctas_per_sm(size_t smem) {
    static int cached = 0;
    if (cached == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmm_fused_skinny_b16_kernel<NV, NOUT, YB16>, 256, smem) != cudaSuccess || n < 1) {
            cudaGetLastError();
            n = 4;
        }
        cached = n;
    }
    return cached;
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

// The same fused layer with X stored in bf16 (ldx in elements, multiple of 8, covering n_cols rounded up to 8) and Y
// either fp32 (y_bf16 = 0, ldy % 4 == 0) or bf16 (y_bf16 = 1, ldy % 8 == 0).  Accumulation, bias, ReLU and the
// projection are fp32; with a bf16 Y the projection uses the rounded activations (what the backward pass reads).
// Pad columns of a bf16 Y up to the next multiple of 8 are written as zeros.
extern "C" int gmc_spmm_fused_skinny_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                                          const float* norm_src, const float* norm_dst, const void* X, void* Y,
                                          int32_t y_bf16, int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                                          const float* bias, int32_t relu, const float* W, int32_t n_out, float* T,
                                          int64_t ldt, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && colidx && X && Y && W && T, "gmc_spmm_fused_skinny_bf16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols && ldt >= n_out,
                "gmc_spmm_fused_skinny_bf16: bad sizes");
    GMC_REQUIRE(n_out >= 1 && n_out <= kMaxClasses, "gmc_spmm_fused_skinny_bf16: n_out must be 1..8");
    GMC_REQUIRE(X != Y, "gmc_spmm_fused_skinny_bf16: in-place SpMM is not supported");
    const int c8 = (n_cols + 7) / 8;
    const bool ok = (n_cols % 4 == 0) && n_cols <= 512 && n_cols >= 16 && (ldx % 8 == 0) && ldx >= (int64_t)c8 * 8 &&
                    (ldy % (y_bf16 ? 8 : 4) == 0) && aligned16(X) && aligned16(Y) && (!bias || aligned16(bias));
    if (!ok) {
        set_error("gmc_spmm_fused_skinny_bf16: needs 16 <= n_cols <= 512, n_cols %% 4 == 0, ldx %% 8 == 0 covering n_cols "
                  "rounded up to 8, and 16-byte aligned rows");
        return GMC_ERR_UNSUPPORTED;
    }
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = as_stream(stream);
    const int warps = 8;
    const int64_t blocks = ceil_div<int64_t>(n_rows, warps);
    const size_t smem = (size_t)c8 * 8 * n_out * sizeof(float);
    const uint4* X8 = reinterpret_cast<const uint4*>(X);
#define GMC_LAUNCH(NV, K, YB)                                                                                        \
    {                                                                                                               \
        const int64_t resident = (int64_t)sm_count() * fused_b16_ctas_per_sm<NV, K, YB>(smem);                      \
        const unsigned grid = (unsigned)(blocks < resident ? blocks : resident);                                    \
        spmm_fused_skinny_b16_kernel<NV, K, YB><<<grid, warps * 32, smem, s>>>(rowptr, colidx, vals, norm_src,      \
                                                                               norm_dst, X8, Y, n_rows, n_cols, c8, \
                                                                               ldx / 8, ldy, bias, relu, W, T, ldt); \
    }
#define GMC_NV(K)                                                          \
    if (y_bf16) { if (c8 <= 32) GMC_LAUNCH(1, K, true) else GMC_LAUNCH(2, K, true) } \
    else { if (c8 <= 32) GMC_LAUNCH(1, K, false) else GMC_LAUNCH(2, K, false) }
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
