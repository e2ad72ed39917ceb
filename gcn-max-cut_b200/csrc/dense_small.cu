// dense_small.cu -- the skinny (H -> n_classes) second layer and bias-gradient column sums.
//
// With n_out = 3 these are HBM-bound passes over H[N, n_in]; they run on CUDA cores by
// design (padding K=3 to an MMA tile would waste >97% of the tensor pipe).
//
//   gmc_skinny_fwd_f32 : T = H * W                          (th.matmul inside conv2, TrainingNeural.py:83)
//   gmc_skinny_bwd_f32 : dHpre = relu'(H) .* (dT * W^T),  dW = H^T dT,  dbias1 = colsum(dHpre)
//                        in ONE pass over H (autograd of TrainingNeural.py:81-83)
//   gmc_colsum_f32     : bias gradients
// All reductions over nodes are two-stage with a fixed order -> bitwise reproducible.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma_util.cuh"

namespace gmc {

constexpr int kSkinnyMaxSmemFloats = 12288;   // 48 KB of W^T

template <int NOUT>
__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const float* __restrict__ H, int64_t ldh, const float* __restrict__ W, float* __restrict__ T,
                  int64_t ldt, int64_t n_rows, int n_in, int vec) {
    extern __shared__ float Ws[];                         // Ws[k][j], row stride n_in_pad
    const int n_in_pad = (n_in + 3) & ~3;
    for (int i = threadIdx.x; i < n_in_pad * NOUT; i += blockDim.x) {
        const int k = i / n_in_pad, j = i % n_in_pad;
        Ws[i] = (j < n_in) ? W[(int64_t)j * NOUT + k] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows;
         row += (int64_t)gridDim.x * warps_per_block) {
        float acc[NOUT];
#pragma unroll
        for (int k = 0; k < NOUT; ++k) acc[k] = 0.f;
        const float* h = H + row * ldh;
        if (vec) {
            for (int j = lane * 4; j < n_in; j += 128) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(h + j));
#pragma unroll
                for (int k = 0; k < NOUT; ++k) {
                    const float4 w = *reinterpret_cast<const float4*>(&Ws[k * n_in_pad + j]);
                    acc[k] = fmaf(x.x, w.x, acc[k]); acc[k] = fmaf(x.y, w.y, acc[k]);
                    acc[k] = fmaf(x.z, w.z, acc[k]); acc[k] = fmaf(x.w, w.w, acc[k]);
                }
            }
        } else {
            for (int j = lane; j < n_in; j += 32) {
                const float x = __ldg(h + j);
#pragma unroll
                for (int k = 0; k < NOUT; ++k) acc[k] = fmaf(x, Ws[k * n_in_pad + j], acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < NOUT; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NOUT; ++k) T[row * ldt + k] = acc[k];
        }
    }
}

// bf16 H (8 columns per 16-byte unit): a warp takes four rows per step -- eight 16-byte loads in flight per lane --
// keeps its 16 columns of W in registers, and folds the 4 x NOUT partial sums with warp_multi_sum (one shuffle per
// value instead of five).  T = H W of TrainingNeural.py:83 when H1 is stored in bf16 (forward via the slab SpMM).
template <int NU, int NOUT>
__global__ void __launch_bounds__(256)
skinny_fwd_b16_kernel(const uint4* __restrict__ H, int64_t ldh8, const float* __restrict__ W, float* __restrict__ T,
                      int64_t ldt, int64_t n_rows, int n_in, int c8) {
    constexpr int R = 4, V = R * NOUT;
    constexpr int P = V <= 4 ? 4 : (V <= 8 ? 8 : (V <= 16 ? 16 : 32));
    constexpr int SH = P == 4 ? 3 : (P == 8 ? 2 : (P == 16 ? 1 : 0));
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t total_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float w[NU][8][NOUT];
#pragma unroll
    for (int q = 0; q < NU; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int col = (lane + 32 * q) * 8 + i;
#pragma unroll
            for (int k = 0; k < NOUT; ++k) w[q][i][k] = col < n_in ? __ldg(W + (int64_t)col * NOUT + k) : 0.f;
        }
    for (int64_t r0 = gwarp * R; r0 < n_rows; r0 += total_warps * R) {
        uint4 h[R][NU];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < NU; ++q) {
                const int u = lane + 32 * q;
                h[r][q] = (r0 + r < n_rows && u < c8) ? __ldg(H + (r0 + r) * ldh8 + u) : make_uint4(0u, 0u, 0u, 0u);
                if (u * 8 + 4 >= n_in) { h[r][q].z = 0u; h[r][q].w = 0u; }     // pad columns never reach the sums
            }
        float v[P];
#pragma unroll
        for (int i = 0; i < P; ++i) v[i] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < NU; ++q) {
                const uint32_t ww[4] = {h[r][q].x, h[r][q].y, h[r][q].z, h[r][q].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float lo = __uint_as_float(ww[i] << 16), hi = __uint_as_float(ww[i] & 0xffff0000u);
#pragma unroll
                    for (int k = 0; k < NOUT; ++k) {
                        v[r * NOUT + k] = fmaf(lo, w[q][2 * i][k], v[r * NOUT + k]);
                        v[r * NOUT + k] = fmaf(hi, w[q][2 * i + 1][k], v[r * NOUT + k]);
                    }
                }
            }
        const float tot = warp_multi_sum<P>(v, lane);
        const int idx = lane >> SH;
        if ((lane & ((1 << SH) - 1)) == 0 && idx < V) {
            const int r = idx / NOUT, k = idx - r * NOUT;
            if (r0 + r < n_rows) T[(r0 + r) * ldt + k] = tot;
        }
    }
}

// ws layout: [n_ctas][n_in][NOUT + 1]   (last slot = dbias partial)
template <int NOUT>
__global__ void __launch_bounds__(128)
skinny_bwd_kernel(const float* __restrict__ dT, int64_t lddt, const float* __restrict__ W, const float* __restrict__ H,
                  int64_t ldh, float* __restrict__ dH, int64_t lddh, int64_t n_rows, int n_in, float* __restrict__ ws) {
    pdl_prologue();
    const int64_t rows_per = ceil_div<int64_t>(n_rows, gridDim.x);
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = min(n_rows, r0 + rows_per);
    float* my_ws = ws + (int64_t)blockIdx.x * n_in * (NOUT + 1);
    for (int j0 = threadIdx.x * 4; j0 < n_in; j0 += blockDim.x * 4) {
        float w[4][NOUT], dw[4][NOUT], db[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            db[i] = 0.f;
#pragma unroll
            for (int k = 0; k < NOUT; ++k) {
                w[i][k] = (j0 + i < n_in) ? __ldg(W + (int64_t)(j0 + i) * NOUT + k) : 0.f;
                dw[i][k] = 0.f;
            }
        }
        auto row = [&](const float (&t)[NOUT], const float4 h4, int64_t v) {
            const float h[4] = {h4.x, h4.y, h4.z, h4.w};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < NOUT; ++k) { s = fmaf(t[k], w[i][k], s); dw[i][k] = fmaf(h[i], t[k], dw[i][k]); }
                o[i] = h[i] > 0.f ? s : 0.f;
                db[i] += o[i];
            }
            *reinterpret_cast<float4*>(dH + v * lddh + j0) = make_float4(o[0], o[1], o[2], o[3]);
        };
        // two rows per iteration, both loads issued before the first use (more bytes in flight per thread)
        int64_t v = r0;
        for (; v + 2 <= r1; v += 2) {
            float t0[NOUT], t1[NOUT];
            const float4 ha = __ldg(reinterpret_cast<const float4*>(H + v * ldh + j0));
            const float4 hb = __ldg(reinterpret_cast<const float4*>(H + (v + 1) * ldh + j0));
#pragma unroll
            for (int k = 0; k < NOUT; ++k) { t0[k] = __ldg(dT + v * lddt + k); t1[k] = __ldg(dT + (v + 1) * lddt + k); }
            row(t0, ha, v);
            row(t1, hb, v + 1);
        }
        for (; v < r1; ++v) {
            float t[NOUT];
#pragma unroll
            for (int k = 0; k < NOUT; ++k) t[k] = __ldg(dT + v * lddt + k);
            row(t, __ldg(reinterpret_cast<const float4*>(H + v * ldh + j0)), v);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j0 + i < n_in) {
#pragma unroll
                for (int k = 0; k < NOUT; ++k) my_ws[(int64_t)(j0 + i) * (NOUT + 1) + k] = dw[i][k];
                my_ws[(int64_t)(j0 + i) * (NOUT + 1) + NOUT] = db[i];
            }
        }
    }
}

// bf16 activations: H (the forward's rounded ReLU output) and dHpre are bf16 matrices, so the pass moves 4 bytes per
// (node, hidden unit) instead of 16.  A thread still owns 4 columns (one 8-byte load and store per row); four rows are
// in flight per thread to keep the bytes in flight of the fp32 kernel.  dW, dbias and the products are fp32; dbias sums
// the unrounded values.
template <int NOUT>
__global__ void __launch_bounds__(128)
skinny_bwd_b16_kernel(const float* __restrict__ dT, int64_t lddt, const float* __restrict__ W, const uint2* __restrict__ H,
                      int64_t ldh4, uint2* __restrict__ dH, int64_t lddh4, int64_t n_rows, int n_in, float* __restrict__ ws) {
    pdl_prologue();
    const int64_t rows_per = ceil_div<int64_t>(n_rows, gridDim.x);
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = min(n_rows, r0 + rows_per);
    float* my_ws = ws + (int64_t)blockIdx.x * n_in * (NOUT + 1);
    for (int j0 = threadIdx.x * 4; j0 < n_in; j0 += blockDim.x * 4) {
        const int j4 = j0 >> 2;
        float w[4][NOUT], dw[4][NOUT], db[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            db[i] = 0.f;
#pragma unroll
            for (int k = 0; k < NOUT; ++k) {
                w[i][k] = (j0 + i < n_in) ? __ldg(W + (int64_t)(j0 + i) * NOUT + k) : 0.f;
                dw[i][k] = 0.f;
            }
        }
        auto row = [&](const float (&t)[NOUT], const uint2 h2, int64_t v) {
            const float h[4] = {__uint_as_float(h2.x << 16), __uint_as_float(h2.x & 0xffff0000u),
                                __uint_as_float(h2.y << 16), __uint_as_float(h2.y & 0xffff0000u)};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < NOUT; ++k) { s = fmaf(t[k], w[i][k], s); dw[i][k] = fmaf(h[i], t[k], dw[i][k]); }
                o[i] = h[i] > 0.f ? s : 0.f;
                db[i] += o[i];
            }
            uint2 pk;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(o[1]), "f"(o[0]));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(o[3]), "f"(o[2]));
            dH[v * lddh4 + j4] = pk;
        };
        int64_t v = r0;
        for (; v + 4 <= r1; v += 4) {
            float t[4][NOUT];
            uint2 h[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) h[r] = __ldg(H + (v + r) * ldh4 + j4);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < NOUT; ++k) t[r][k] = __ldg(dT + (v + r) * lddt + k);
#pragma unroll
            for (int r = 0; r < 4; ++r) row(t[r], h[r], v + r);
        }
        for (; v < r1; ++v) {
            float t[NOUT];
#pragma unroll
            for (int k = 0; k < NOUT; ++k) t[k] = __ldg(dT + v * lddt + k);
            row(t, __ldg(H + v * ldh4 + j4), v);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j0 + i < n_in) {
#pragma unroll
                for (int k = 0; k < NOUT; ++k) my_ws[(int64_t)(j0 + i) * (NOUT + 1) + k] = dw[i][k];
                my_ws[(int64_t)(j0 + i) * (NOUT + 1) + NOUT] = db[i];
            }
        }
    }
}

// fp32-grade form for the split-operand weight-gradient GEMM (split.cu): H stays fp32, and dHpre leaves as NS stacked
// bf16 parts (hi, lo [, lo2]) of row_scale[v] * dHpre[v, :] -- part s at rows [s * split_rows, s * split_rows + n_rows)
// of dH -- the B operand of dW1 = XI^T (s . dH1pre).  dW and dbias are sums of the UNSCALED fp32 values.
// two fp32 -> one word of two 16-bit floats (round to nearest even; fp16 saturates instead of overflowing to inf), and
// what that word holds as fp32 again -- the exact residual o - parts drives the next part
template <bool F16>
__device__ __forceinline__ uint32_t pack_part(float lo, float hi) {
    uint32_t w;
    if (F16) {
        lo = fminf(fmaxf(lo, -65504.f), 65504.f); hi = fminf(fmaxf(hi, -65504.f), 65504.f);
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    } else {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    }
    return w;
}
template <bool F16>
__device__ __forceinline__ void unpack_part(uint32_t w, float& lo, float& hi) {
    if (F16) {
        const __half2 h = *reinterpret_cast<const __half2*>(&w);
        lo = __low2float(h); hi = __high2float(h);
    } else {
        lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
    }
}

template <int NOUT, int NS, bool F16 = false>
__global__ void __launch_bounds__(128)
skinny_bwd_split_kernel(const float* __restrict__ dT, int64_t lddt, const float* __restrict__ W, const float* __restrict__ H,
                        int64_t ldh, const float* __restrict__ row_scale, uint2* __restrict__ dH, int64_t lddh4,
                        int64_t split_rows, int64_t n_rows, int n_in, float* __restrict__ ws, float lo_up) {
    pdl_prologue();
    const int64_t rows_per = ceil_div<int64_t>(n_rows, gridDim.x);
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = min(n_rows, r0 + rows_per);
    float* my_ws = ws + (int64_t)blockIdx.x * n_in * (NOUT + 1);
    for (int j0 = threadIdx.x * 4; j0 < n_in; j0 += blockDim.x * 4) {
        const int j4 = j0 >> 2;
        float w[4][NOUT], dw[4][NOUT], db[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            db[i] = 0.f;
#pragma unroll
            for (int k = 0; k < NOUT; ++k) {
                w[i][k] = (j0 + i < n_in) ? __ldg(W + (int64_t)(j0 + i) * NOUT + k) : 0.f;
                dw[i][k] = 0.f;
            }
        }
        auto row = [&](const float (&t)[NOUT], const float4 h4, float rs, int64_t v) {
            const float h[4] = {h4.x, h4.y, h4.z, h4.w};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < NOUT; ++k) { s = fmaf(t[k], w[i][k], s); dw[i][k] = fmaf(h[i], t[k], dw[i][k]); }
                s = h[i] > 0.f ? s : 0.f;
                db[i] += s;
                o[i] = s * rs;
            }
#pragma unroll
            for (int sp = 0; sp < NS; ++sp) {
                uint2 pk;
                pk.x = pack_part<F16>(o[0], o[1]);
                pk.y = pack_part<F16>(o[2], o[3]);
                dH[((int64_t)sp * split_rows + v) * lddh4 + j4] = pk;
                if (sp + 1 < NS) {                                 // exact residuals (scaled up for fp16 parts): the next part
                    float p0, p1, p2, p3;
                    unpack_part<F16>(pk.x, p0, p1); unpack_part<F16>(pk.y, p2, p3);
                    o[0] = (o[0] - p0) * lo_up; o[1] = (o[1] - p1) * lo_up;
                    o[2] = (o[2] - p2) * lo_up; o[3] = (o[3] - p3) * lo_up;
                }
            }
        };
        int64_t v = r0;
        for (; v + 4 <= r1; v += 4) {
            float t[4][NOUT], rs[4];
            float4 h[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) h[r] = ldg_stream4(reinterpret_cast<const float4*>(H + (v + r) * ldh + j0));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                rs[r] = row_scale ? __ldg(row_scale + v + r) : 1.f;
#pragma unroll
                for (int k = 0; k < NOUT; ++k) t[r][k] = __ldg(dT + (v + r) * lddt + k);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) row(t[r], h[r], rs[r], v + r);
        }
        for (; v < r1; ++v) {
            float t[NOUT];
#pragma unroll
            for (int k = 0; k < NOUT; ++k) t[k] = __ldg(dT + v * lddt + k);
            row(t, __ldg(reinterpret_cast<const float4*>(H + v * ldh + j0)), row_scale ? __ldg(row_scale + v) : 1.f, v);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j0 + i < n_in) {
#pragma unroll
                for (int k = 0; k < NOUT; ++k) my_ws[(int64_t)(j0 + i) * (NOUT + 1) + k] = dw[i][k];
                my_ws[(int64_t)(j0 + i) * (NOUT + 1) + NOUT] = db[i];
            }
        }
    }
}

// ---- TMA-streamed forms of the two bf16 skinny kernels -----------------------------------------------------------
// The register-staged kernels above are latency-bound (ncu, profiles/r01c_*: 8-10 warps stalled on the long scoreboard
// per issue, 0.50-0.60 of the HBM roofline): the bytes in flight are capped by the registers that hold them.  Here the
// H rows arrive through cp.async.bulk.tensor boxes (256 columns x TR rows, out-of-range rows / pad columns zero-filled)
// in a 4-stage shared-memory ring filled by one thread, so ~3 stages per CTA are always in flight and the compute reads
// shared memory.  A tile is released with one __syncthreads, after which thread 0 refills its stage.
#ifndef GMC_STREAM_STAGES
#define GMC_STREAM_STAGES 4
#endif
#ifndef GMC_BWD_TILE_ROWS
#define GMC_BWD_TILE_ROWS 8
#endif
constexpr int kStreamStages = GMC_STREAM_STAGES;
constexpr int kFwdTileRows = 32;
constexpr int kBwdTileRows = GMC_BWD_TILE_ROWS;

template <int NB, int NOUT>
__global__ void __launch_bounds__(512, 1)
skinny_fwd_b16_tma_kernel(const __grid_constant__ CUtensorMap tmH, const float* __restrict__ W, float* __restrict__ T,
                          int64_t ldt, int64_t n_rows, int n_in, int64_t n_tiles) {
    constexpr int TR = kFwdTileRows, NST = kStreamStages, R = 2, V = R * NOUT;
    constexpr int P = V <= 2 ? 2 : (V <= 4 ? 4 : 8);
    constexpr int SH = P == 2 ? 4 : (P == 4 ? 3 : 2);
    constexpr uint32_t BOX_BYTES = TR * 512, STAGE_BYTES = NB * BOX_BYTES;
    extern __shared__ __align__(128) uint8_t stream_smem[];
    const uint32_t sbase = tma::smem_u32(stream_smem);
    const uint32_t bar0 = sbase + NST * STAGE_BYTES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        tma::prefetch_map(&tmH);
        for (int s = 0; s < NST; ++s) tma::mbar_init(bar0 + 8 * s, 1);
        tma::mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int64_t t, int st) {
        const uint32_t bar = bar0 + 8 * st;
        tma::mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int q = 0; q < NB; ++q)
            tma::load_2d(sbase + st * STAGE_BYTES + q * BOX_BYTES, &tmH, q * 256, (int)(t * TR), bar);
    };
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (t < n_tiles) issue(t, s);
        }
    }
    float w[NB][8][NOUT];
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int col = q * 256 + lane * 8 + i;
#pragma unroll
            for (int k = 0; k < NOUT; ++k) w[q][i][k] = col < n_in ? __ldg(W + (int64_t)col * NOUT + k) : 0.f;
        }
    int st = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        tma::mbar_wait(bar0 + 8 * st, phase);
        const uint8_t* stage = stream_smem + st * STAGE_BYTES;
        float v[P];
#pragma unroll
        for (int i = 0; i < P; ++i) v[i] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const uint4 h = *reinterpret_cast<const uint4*>(stage + q * BOX_BYTES + (R * warp + r) * 512 + lane * 16);
                const uint32_t ww[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float lo = __uint_as_float(ww[i] << 16), hi = __uint_as_float(ww[i] & 0xffff0000u);
#pragma unroll
                    for (int k = 0; k < NOUT; ++k) {
                        v[r * NOUT + k] = fmaf(lo, w[q][2 * i][k], v[r * NOUT + k]);
                        v[r * NOUT + k] = fmaf(hi, w[q][2 * i + 1][k], v[r * NOUT + k]);
                    }
                }
            }
        const float tot = warp_multi_sum<P>(v, lane);
        const int idx = lane >> SH;
        if ((lane & ((1 << SH) - 1)) == 0 && idx < V) {
            const int r = idx / NOUT, k = idx - r * NOUT;
            const int64_t row = t * TR + R * warp + r;
            if (row < n_rows) T[row * ldt + k] = tot;
        }
        __syncthreads();                                   // every warp is done with the stage
        const int64_t tn = t + (int64_t)NST * gridDim.x;
        if (tid == 0 && tn < n_tiles) issue(tn, st);
        if (++st == NST) { st = 0; phase ^= 1; }
    }
}

// CTA b owns rows [b * rows_per, ...) with rows_per a multiple of the tile height, walks them in tiles of 8 rows;
// thread j owns columns 4j .. 4j+3.  The tile's dT rows (8 x NOUT floats) are fetched one tile ahead by the first
// 8 * NOUT threads and handed over through shared memory.  ws layout as skinny_bwd_kernel.
template <int NB, int NOUT>
__global__ void __launch_bounds__(128)
skinny_bwd_b16_tma_kernel(const __grid_constant__ CUtensorMap tmH, const float* __restrict__ dT, int64_t lddt,
                          const float* __restrict__ W, uint2* __restrict__ dH, int64_t lddh4, int64_t n_rows, int n_in,
                          int64_t rows_per, float* __restrict__ ws) {
    constexpr int TR = kBwdTileRows, NST = kStreamStages;
    constexpr uint32_t BOX_BYTES = TR * 512, STAGE_BYTES = NB * BOX_BYTES;
    extern __shared__ __align__(128) uint8_t stream_smem[];
    __shared__ float tsm[2][TR * NOUT];
    const uint32_t sbase = tma::smem_u32(stream_smem);
    const uint32_t bar0 = sbase + NST * STAGE_BYTES;
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = min(n_rows, r0 + rows_per);
    const int n_tiles = r1 > r0 ? (int)ceil_div<int64_t>(r1 - r0, TR) : 0;
    if (tid == 0) {
        tma::prefetch_map(&tmH);
        for (int s = 0; s < NST; ++s) tma::mbar_init(bar0 + 8 * s, 1);
        tma::mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int t, int st) {
        const uint32_t bar = bar0 + 8 * st;
        tma::mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int q = 0; q < NB; ++q)
            tma::load_2d(sbase + st * STAGE_BYTES + q * BOX_BYTES, &tmH, q * 256, (int)(r0 + (int64_t)t * TR), bar);
    };
    if (tid == 0)
        for (int s = 0; s < NST && s < n_tiles; ++s) issue(s, s);
    auto fetch_t = [&](int t) -> float {                   // thread tid < TR * NOUT: element (row tid / NOUT, k) of tile t
        const int64_t v = r0 + (int64_t)t * TR + tid / NOUT;
        return (tid < TR * NOUT && t < n_tiles && v < r1) ? __ldg(dT + v * lddt + (tid % NOUT)) : 0.f;
    };
    if (tid < TR * NOUT) tsm[0][tid] = fetch_t(0);

    const int j0 = tid * 4;
    const bool owner = j0 < n_in;
    const bool pad_owner = !owner && j0 < min((int)(lddh4 * 4), (n_in + 63) & ~63);   // zero-fills the row's pad columns
    // The kernel is issue-bound (ncu r01c: 71 % issue slots), and 24 of its ~40 instructions per row and thread were
    // scalar FMAs.  Columns are held in PAIRS and multiplied with packed fp32x2 FMAs (fma.rn.f32x2, sm_100): each lane
    // of a pair is an IEEE fma, so the results are bit-identical to the scalar form at half the FMA instructions.
    float2 wp[2][NOUT], dwp[2][NOUT];                      // columns (4j, 4j+1) and (4j+2, 4j+3)
    float db[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) db[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            wp[q][k].x = (j0 + 2 * q < n_in) ? __ldg(W + (int64_t)(j0 + 2 * q) * NOUT + k) : 0.f;
            wp[q][k].y = (j0 + 2 * q + 1 < n_in) ? __ldg(W + (int64_t)(j0 + 2 * q + 1) * NOUT + k) : 0.f;
            dwp[q][k] = make_float2(0.f, 0.f);
        }
    const uint32_t my_off = (uint32_t)(tid >> 6) * BOX_BYTES + (uint32_t)(tid & 63) * 8;   // box, then 8 bytes per thread
    __syncthreads();
    int st = 0;
    uint32_t phase = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const float t_next = fetch_t(t + 1);               // in flight during this tile
        tma::mbar_wait(bar0 + 8 * st, phase);
        const uint8_t* stage = stream_smem + st * STAGE_BYTES + my_off;
        const float* ts = tsm[t & 1];
        const int64_t vb = r0 + (int64_t)t * TR;
        if (owner) {
#pragma unroll
            for (int r = 0; r < TR; ++r) {
                const int64_t v = vb + r;
                if (v < r1) {                              // CTA-uniform
                    const uint2 h2 = *reinterpret_cast<const uint2*>(stage + r * 512);
                    const float2 hp[2] = {make_float2(__uint_as_float(h2.x << 16), __uint_as_float(h2.x & 0xffff0000u)),
                                          make_float2(__uint_as_float(h2.y << 16), __uint_as_float(h2.y & 0xffff0000u))};
                    float2 sp[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int k = 0; k < NOUT; ++k) {
                        const float tk = ts[r * NOUT + k];
                        const float2 tt = make_float2(tk, tk);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            sp[q] = __ffma2_rn(tt, wp[q][k], sp[q]);
                            dwp[q][k] = __ffma2_rn(hp[q], tt, dwp[q][k]);
                        }
                    }
                    const float o[4] = {hp[0].x > 0.f ? sp[0].x : 0.f, hp[0].y > 0.f ? sp[0].y : 0.f,
                                        hp[1].x > 0.f ? sp[1].x : 0.f, hp[1].y > 0.f ? sp[1].y : 0.f};
#pragma unroll
                    for (int i = 0; i < 4; ++i) db[i] += o[i];
                    uint2 pk;
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(o[1]), "f"(o[0]));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(o[3]), "f"(o[2]));
                    dH[v * lddh4 + tid] = pk;
                }
            }
        } else if (pad_owner) {
            // columns n_in .. next 128-byte line as zeros: a 1000-byte row in a 1024-byte pitch would end inside a sector
            // and turn the row's last write into a read-modify-write (profiles/r02r_write_pattern_probe: 3.9 vs 6.4 TB/s)
#pragma unroll
            for (int r = 0; r < TR; ++r)
                if (vb + r < r1) dH[(vb + r) * lddh4 + tid] = make_uint2(0u, 0u);
        }
        if (tid < TR * NOUT) tsm[(t + 1) & 1][tid] = t_next;
        __syncthreads();                                   // stage and tsm[t & 1] are free, tsm[(t + 1) & 1] is published
        if (tid == 0 && t + NST < n_tiles) issue(t + NST, st);
        if (++st == NST) { st = 0; phase ^= 1; }
    }
    if (owner) {
        float* my_ws = ws + (int64_t)blockIdx.x * n_in * (NOUT + 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j0 + i < n_in) {
#pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    my_ws[(int64_t)(j0 + i) * (NOUT + 1) + k] = (i & 1) ? dwp[i >> 1][k].y : dwp[i >> 1][k].x;
                my_ws[(int64_t)(j0 + i) * (NOUT + 1) + NOUT] = db[i];
            }
        }
    }
}

// TMA-streamed form of skinny_bwd_split_kernel (fp32 H in, NS stacked bf16 parts of row_scale . dHpre out): the same
// 4-stage ring of 8-row tiles as skinny_bwd_b16_tma_kernel with fp32 boxes of 128 columns (512-byte rows), thread j owns
// columns 4j .. 4j+3 (one LDS.128 per row), packed fp32x2 FMAs.  16 KB per stage at n_in = 512: three CTAs per SM.
template <int NB, int NOUT, int NS, bool F16 = false>
__global__ void __launch_bounds__(128)
skinny_bwd_split_tma_kernel(const __grid_constant__ CUtensorMap tmH, const float* __restrict__ dT, int64_t lddt,
                            const float* __restrict__ W, const float* __restrict__ row_scale, uint2* __restrict__ dH,
                            int64_t lddh4, int64_t split_rows, int64_t n_rows, int n_in, int64_t rows_per,
                            float* __restrict__ ws, float lo_up) {
    constexpr int TR = kBwdTileRows, NST = kStreamStages;
    constexpr uint32_t BOX_BYTES = TR * 512, STAGE_BYTES = NB * BOX_BYTES;
    extern __shared__ __align__(128) uint8_t stream_smem[];
    __shared__ float tsm[2][TR * (NOUT + 1)];              // dT rows + the row scale of the tile
    const uint32_t sbase = tma::smem_u32(stream_smem);
    const uint32_t bar0 = sbase + NST * STAGE_BYTES;
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = min(n_rows, r0 + rows_per);
    const int n_tiles = r1 > r0 ? (int)ceil_div<int64_t>(r1 - r0, TR) : 0;
    if (tid == 0) {
        tma::prefetch_map(&tmH);
        for (int s = 0; s < NST; ++s) tma::mbar_init(bar0 + 8 * s, 1);
        tma::mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int t, int st) {
        const uint32_t bar = bar0 + 8 * st;
        tma::mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int q = 0; q < NB; ++q)
            tma::load_2d(sbase + st * STAGE_BYTES + q * BOX_BYTES, &tmH, q * 128, (int)(r0 + (int64_t)t * TR), bar);
    };
    if (tid == 0)
        for (int s = 0; s < NST && s < n_tiles; ++s) issue(s, s);
    auto fetch_t = [&](int t) -> float {                   // thread tid < TR * (NOUT + 1): element (row, k) of tile t
        const int row = tid / (NOUT + 1), k = tid % (NOUT + 1);
        const int64_t v = r0 + (int64_t)t * TR + row;
        if (tid >= TR * (NOUT + 1) || t >= n_tiles || v >= r1) return 0.f;
        if (k < NOUT) return __ldg(dT + v * lddt + k);
        return row_scale ? __ldg(row_scale + v) : 1.f;
    };
    if (tid < TR * (NOUT + 1)) tsm[0][tid] = fetch_t(0);

    const int j0 = tid * 4;
    const bool owner = j0 < n_in;
    const bool pad_owner = !owner && j0 < min((int)(lddh4 * 4), (n_in + 63) & ~63);   // zero-fills the row's pad columns
    float2 wp[2][NOUT], dwp[2][NOUT];
    float db[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) db[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            wp[q][k].x = (j0 + 2 * q < n_in) ? __ldg(W + (int64_t)(j0 + 2 * q) * NOUT + k) : 0.f;
            wp[q][k].y = (j0 + 2 * q + 1 < n_in) ? __ldg(W + (int64_t)(j0 + 2 * q + 1) * NOUT + k) : 0.f;
            dwp[q][k] = make_float2(0.f, 0.f);
        }
    const uint32_t my_off = (uint32_t)(tid >> 5) * BOX_BYTES + (uint32_t)(tid & 31) * 16;   // box, then 16 bytes per thread
    __syncthreads();
    int st = 0;
    uint32_t phase = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const float t_next = fetch_t(t + 1);               // in flight during this tile
        tma::mbar_wait(bar0 + 8 * st, phase);
        const uint8_t* stage = stream_smem + st * STAGE_BYTES + my_off;
        const float* ts = tsm[t & 1];
        const int64_t vb = r0 + (int64_t)t * TR;
        if (owner) {
#pragma unroll
            for (int r = 0; r < TR; ++r) {
                const int64_t v = vb + r;
                if (v < r1) {                              // CTA-uniform
                    const float4 h4 = *reinterpret_cast<const float4*>(stage + r * 512);
                    const float2 hp[2] = {make_float2(h4.x, h4.y), make_float2(h4.z, h4.w)};
                    float2 sp[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int k = 0; k < NOUT; ++k) {
                        const float tk = ts[r * (NOUT + 1) + k];
                        const float2 tt = make_float2(tk, tk);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            sp[q] = __ffma2_rn(tt, wp[q][k], sp[q]);
                            dwp[q][k] = __ffma2_rn(hp[q], tt, dwp[q][k]);
                        }
                    }
                    const float rs = ts[r * (NOUT + 1) + NOUT];
                    float o[4] = {hp[0].x > 0.f ? sp[0].x : 0.f, hp[0].y > 0.f ? sp[0].y : 0.f,
                                  hp[1].x > 0.f ? sp[1].x : 0.f, hp[1].y > 0.f ? sp[1].y : 0.f};
#pragma unroll
                    for (int i = 0; i < 4; ++i) { db[i] += o[i]; o[i] *= rs; }
#pragma unroll
                    for (int sp_i = 0; sp_i < NS; ++sp_i) {
                        uint2 pk;
                        pk.x = pack_part<F16>(o[0], o[1]);
                        pk.y = pack_part<F16>(o[2], o[3]);
                        dH[((int64_t)sp_i * split_rows + v) * lddh4 + tid] = pk;
                        if (sp_i + 1 < NS) {
                            float p0, p1, p2, p3;
                            unpack_part<F16>(pk.x, p0, p1); unpack_part<F16>(pk.y, p2, p3);
                            o[0] = (o[0] - p0) * lo_up; o[1] = (o[1] - p1) * lo_up;
                            o[2] = (o[2] - p2) * lo_up; o[3] = (o[3] - p3) * lo_up;
                        }
                    }
                }
            }
        }
        if (pad_owner) {                                   // full 128-byte lines (see skinny_bwd_b16_tma_kernel)
#pragma unroll
            for (int r = 0; r < TR; ++r)
                if (vb + r < r1)
#pragma unroll
                    for (int sp_i = 0; sp_i < NS; ++sp_i) dH[((int64_t)sp_i * split_rows + vb + r) * lddh4 + tid] = make_uint2(0u, 0u);
        }
        if (tid < TR * (NOUT + 1)) tsm[(t + 1) & 1][tid] = t_next;
        __syncthreads();                                   // stage and tsm[t & 1] are free, tsm[(t + 1) & 1] is published
        if (tid == 0 && t + NST < n_tiles) issue(t + NST, st);
        if (++st == NST) { st = 0; phase ^= 1; }
    }
    if (owner) {
        float* my_ws = ws + (int64_t)blockIdx.x * n_in * (NOUT + 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j0 + i < n_in) {
#pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    my_ws[(int64_t)(j0 + i) * (NOUT + 1) + k] = (i & 1) ? dwp[i >> 1][k].y : dwp[i >> 1][k].x;
                my_ws[(int64_t)(j0 + i) * (NOUT + 1) + NOUT] = db[i];
            }
        }
    }
}

static bool stream_kernels_enabled() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("GMC_SKINNY_TMA"); cached = (e && e[0] == '0') ? 0 : 1; }
    return cached == 1;
}

__global__ void skinny_bwd_reduce_kernel(const float* __restrict__ ws, int n_ctas, int n_in, int nout,
                                         float* __restrict__ dW, float* __restrict__ dbias) {
    // block = 32 outputs (x) x 32 partial-groups (y): coalesced 128-byte reads, fixed summation order
    __shared__ float red[32][33];
    pdl_prologue();
    const int total = n_in * (nout + 1);
    const int i = blockIdx.x * 32 + threadIdx.x;               // over n_in * (nout + 1)
    float s = 0.f;
    if (i < total)
        for (int c = threadIdx.y; c < n_ctas; c += 32) s += ws[(int64_t)c * total + i];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && i < total) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += red[y][threadIdx.x];
        const int j = i / (nout + 1), k = i % (nout + 1);
        if (k < nout) dW[(int64_t)j * nout + k] = t;
        else if (dbias) dbias[j] = t;
    }
}

// ---- column sums ------------------------------------------------------------------------
// stage 1: CTA b reduces rows [b*rows_per, ...) into ws[b][0..C)
__global__ void __launch_bounds__(256)
colsum_stage1(const float* __restrict__ X, int64_t ldx, int64_t n_rows, int n_cols, float* __restrict__ ws) {
    __shared__ float red[256];
    const int64_t rows_per = ceil_div<int64_t>(n_rows, gridDim.x);
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = min(n_rows, r0 + rows_per);
    float* out = ws + (int64_t)blockIdx.x * n_cols;
    if (n_cols <= 8) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int64_t v = r0 + threadIdx.x; v < r1; v += blockDim.x)
            for (int k = 0; k < n_cols; ++k) acc[k] += __ldg(X + v * ldx + k);
        for (int k = 0; k < n_cols; ++k) {
            red[threadIdx.x] = acc[k];
            __syncthreads();
            for (int s = 128; s > 0; s >>= 1) {
                if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
                __syncthreads();
            }
            if (threadIdx.x == 0) out[k] = red[0];
            __syncthreads();
        }
    } else {
        for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
            float a = 0.f;
            for (int64_t v = r0; v < r1; ++v) a += __ldg(X + v * ldx + c);
            out[c] = a;
        }
    }
}

__global__ void colsum_stage2(const float* __restrict__ ws, int n_ctas, int n_cols, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    float s = 0.f;
    for (int b = 0; b < n_ctas; ++b) s += ws[(int64_t)b * n_cols + c];
    out[c] = s;
}

static int reduce_ctas() { return sm_count() * 4; }

// CTAs of the (non-streamed) skinny_bwd kernels: two waves of four per SM for large inputs, but never fewer than 8 rows per
// CTA -- every CTA leaves an [n_in, n_out + 1] partial for the reduce kernel, and a 500-node graph's step spent more time
// writing and adding 500 one-row partials than on the rows themselves
static int skinny_bwd_ctas(int64_t n_rows) {
    int n_ctas = reduce_ctas() * 2;
    const int64_t by_rows = n_rows > 8 ? (n_rows + 7) / 8 : 1;
    if ((int64_t)n_ctas > by_rows) n_ctas = (int)by_rows;
    return n_ctas;
}

}  // namespace gmc

extern "C" {

int gmc_skinny_fwd_f32(const float* H, int64_t ldh, const float* W, float* T, int64_t ldt, int64_t n_rows,
                       int32_t n_in, int32_t n_out, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(H && W && T, "gmc_skinny_fwd_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_in > 0 && n_out >= 1 && n_out <= kMaxClasses && ldh >= n_in && ldt >= n_out,
                "gmc_skinny_fwd_f32: bad sizes (n_out must be 1..8)");
    const int n_in_pad = (n_in + 3) & ~3;
    GMC_REQUIRE(n_in_pad * n_out <= kSkinnyMaxSmemFloats, "gmc_skinny_fwd_f32: n_in*n_out too large (%d)", n_in * n_out);
    if (n_rows == 0) return GMC_OK;
    const int vec = (n_in % 4 == 0) && (ldh % 4 == 0) && aligned16(H);
    const int threads = 256;
    int64_t blocks = ceil_div<int64_t>(n_rows, threads / 32);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)n_in_pad * n_out * sizeof(float);
    cudaStream_t s = as_stream(stream);
#define GMC_CASE(K) case K: skinny_fwd_kernel<K><<<(unsigned)blocks, threads, smem, s>>>(H, ldh, W, T, ldt, n_rows, n_in, vec); break;
    switch (n_out) { GMC_CASE(1) GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// T = H W with H stored in bf16 (ldh in elements, multiple of 8 covering n_in rounded up to 8; 16-byte aligned base),
// W and T fp32.  n_in <= 512, n_out <= 4 (W lives in registers); GMC_ERR_UNSUPPORTED otherwise.
int gmc_skinny_fwd_bf16(const void* H, int64_t ldh, const float* W, float* T, int64_t ldt, int64_t n_rows,
                        int32_t n_in, int32_t n_out, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(H && W && T, "gmc_skinny_fwd_bf16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_in > 0 && n_out >= 1 && ldh >= n_in && ldt >= n_out, "gmc_skinny_fwd_bf16: bad sizes");
    const int c8 = (n_in + 7) / 8;
    if (n_out > 4 || n_in > 512 || n_in % 4 != 0 || ldh % 8 != 0 || ldh < (int64_t)c8 * 8 || !aligned16(H)) {
        set_error("gmc_skinny_fwd_bf16: needs n_out <= 4, n_in <= 512, n_in %% 4 == 0, ldh %% 8 == 0 covering n_in rounded "
                  "up to 8 and a 16-byte aligned H");
        return GMC_ERR_UNSUPPORTED;
    }
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = as_stream(stream);
    if (stream_kernels_enabled() && n_rows >= 4096 && n_rows < (1ll << 31) - 64) {
        // TMA-streamed kernel: one persistent CTA per SM, 4 stages of 32 rows
        CUtensorMap tm;
        int rc = tma::make_bf16_map(&tm, H, (uint64_t)n_in, (uint64_t)n_rows, (uint64_t)ldh, 256, kFwdTileRows, "gmc_skinny_fwd_bf16");
        if (rc != GMC_OK) return rc;
        const int nb = n_in > 256 ? 2 : 1;
        const size_t smem = (size_t)kStreamStages * nb * kFwdTileRows * 512 + 64;
        const int64_t n_tiles = ceil_div<int64_t>(n_rows, kFwdTileRows);
        const unsigned grid = (unsigned)(n_tiles < sm_count() ? n_tiles : sm_count());
#define GMC_CASE(NB, K)                                                                                                \
        {                                                                                                              \
            static bool attr = false;                                                                                  \
            if (!attr) {                                                                                               \
                GMC_CUDA(cudaFuncSetAttribute(skinny_fwd_b16_tma_kernel<NB, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              kStreamStages * NB * kFwdTileRows * 512 + 64));                          \
                attr = true;                                                                                           \
            }                                                                                                          \
            skinny_fwd_b16_tma_kernel<NB, K><<<grid, 512, smem, s>>>(tm, W, T, ldt, n_rows, n_in, n_tiles);          \
        }
#define GMC_NB(K) if (nb == 2) GMC_CASE(2, K) else GMC_CASE(1, K)
        switch (n_out) {
            case 1: GMC_NB(1); break;
            case 2: GMC_NB(2); break;
            case 3: GMC_NB(3); break;
            default: GMC_NB(4); break;
        }
#undef GMC_NB
#undef GMC_CASE
        GMC_LAUNCH_CHECK();
        return GMC_OK;
    }
    const uint4* H8 = reinterpret_cast<const uint4*>(H);
    int64_t blocks = ceil_div<int64_t>(n_rows, 8 * 4);
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
#define GMC_CASE(K)                                                                                                   \
    case K:                                                                                                           \
        if (c8 <= 32) skinny_fwd_b16_kernel<1, K><<<(unsigned)blocks, 256, 0, s>>>(H8, ldh / 8, W, T, ldt, n_rows, n_in, c8); \
        else skinny_fwd_b16_kernel<2, K><<<(unsigned)blocks, 256, 0, s>>>(H8, ldh / 8, W, T, ldt, n_rows, n_in, c8);          \
        break;
    switch (n_out) { GMC_CASE(1) GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

size_t gmc_skinny_bwd_workspace_bytes(int32_t n_in, int32_t n_out) {
    return (size_t)gmc::reduce_ctas() * 2 * (size_t)n_in * (size_t)(n_out + 1) * sizeof(float);
}

int gmc_skinny_bwd_f32(const float* dT, int64_t lddt, const float* W, const float* H, int64_t ldh, float* dHpre,
                       int64_t lddh, float* dW, float* dbias, int64_t n_rows, int32_t n_in, int32_t n_out,
                       void* workspace, size_t workspace_bytes, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(dT && W && H && dHpre && dW, "gmc_skinny_bwd_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_in > 0 && n_out >= 1 && n_out <= kMaxClasses && ldh >= n_in && lddh >= n_in && lddt >= n_out,
                "gmc_skinny_bwd_f32: bad sizes (n_out must be 1..8)");
    GMC_REQUIRE(n_in % 4 == 0 && ldh % 4 == 0 && lddh % 4 == 0 && aligned16(H) && aligned16(dHpre),
                "gmc_skinny_bwd_f32: n_in and leading dimensions must be multiples of 4 with 16-byte aligned bases");
    const int n_ctas = skinny_bwd_ctas(n_rows);
    const size_t need = (size_t)n_ctas * n_in * (n_out + 1) * sizeof(float);
    if (!workspace || workspace_bytes < need) {
        set_error("gmc_skinny_bwd_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GMC_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    float* ws = reinterpret_cast<float*>(workspace);
    if (n_rows == 0) {
        GMC_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * n_in * n_out, s));
        if (dbias) GMC_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * n_in, s));
        return GMC_OK;
    }
#define GMC_CASE(K) case K: GMC_CUDA(launch_pdl(skinny_bwd_kernel<K>, n_ctas, 128, 0, s, dT, lddt, W, H, ldh, dHpre, lddh, n_rows, n_in, ws)); break;
    switch (n_out) { GMC_CASE(1) GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    const int total = n_in * (n_out + 1);
    GMC_CUDA(launch_pdl(skinny_bwd_reduce_kernel, ceil_div(total, 32), dim3(32, 32), 0, s, ws, n_ctas, n_in, n_out, dW, dbias));
    return GMC_OK;
}

// gmc_skinny_bwd_f32 with H and dHpre stored in bf16 (leading dimensions in elements, multiples of 4; 8-byte aligned
// bases): autograd of TrainingNeural.py:81-83 when the layer-1 activations are kept in bf16.  Same workspace.
int gmc_skinny_bwd_bf16(const float* dT, int64_t lddt, const float* W, const void* H, int64_t ldh, void* dHpre,
                        int64_t lddh, float* dW, float* dbias, int64_t n_rows, int32_t n_in, int32_t n_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(dT && W && H && dHpre && dW, "gmc_skinny_bwd_bf16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_in > 0 && n_out >= 1 && n_out <= kMaxClasses && ldh >= n_in && lddh >= n_in && lddt >= n_out,
                "gmc_skinny_bwd_bf16: bad sizes (n_out must be 1..8)");
    GMC_REQUIRE(n_in % 4 == 0 && ldh % 4 == 0 && lddh % 4 == 0 && (reinterpret_cast<uintptr_t>(H) & 7u) == 0 &&
                    (reinterpret_cast<uintptr_t>(dHpre) & 7u) == 0,
                "gmc_skinny_bwd_bf16: n_in and leading dimensions must be multiples of 4 with 8-byte aligned bases");
    const int n_ctas = skinny_bwd_ctas(n_rows);
    const size_t need = (size_t)n_ctas * n_in * (n_out + 1) * sizeof(float);
    if (!workspace || workspace_bytes < need) {
        set_error("gmc_skinny_bwd_bf16: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GMC_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    float* ws = reinterpret_cast<float*>(workspace);
    if (n_rows == 0) {
        GMC_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * n_in * n_out, s));
        if (dbias) GMC_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * n_in, s));
        return GMC_OK;
    }
    const uint2* H2 = reinterpret_cast<const uint2*>(H);
    uint2* dH2 = reinterpret_cast<uint2*>(dHpre);
    if (stream_kernels_enabled() && n_rows >= 65536 && n_rows < (1ll << 31) - 64 && n_in <= 512 && ldh % 8 == 0 && aligned16(H)) {
        // TMA-streamed kernel: 6 CTAs per SM (32 KB ring each + 1 KB reserved: 7 do not fit in 227 KB), one wave;
        // row ranges in multiples of the 8-row tile
        CUtensorMap tm;
        int rc = tma::make_bf16_map(&tm, H, (uint64_t)n_in, (uint64_t)n_rows, (uint64_t)ldh, 256, kBwdTileRows, "gmc_skinny_bwd_bf16");
        if (rc != GMC_OK) return rc;
        const int nb = n_in > 256 ? 2 : 1;
        const size_t smem = (size_t)kStreamStages * nb * kBwdTileRows * 512 + 64;
        int per_sm = (int)((226 * 1024) / (smem + 1024 + 256));          // 1 KB reserved per CTA + static shared memory
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        int nc = sm_count() * per_sm;
        if (nc > n_ctas) nc = n_ctas;
        const int64_t rows_per = ceil_div<int64_t>(ceil_div<int64_t>(n_rows, nc), kBwdTileRows) * kBwdTileRows;
        nc = (int)ceil_div<int64_t>(n_rows, rows_per);
#define GMC_CASE(NB, K)                                                                                                \
        {                                                                                                              \
            static bool attr = false;                                                                                  \
            if (!attr) {                                                                                               \
                GMC_CUDA(cudaFuncSetAttribute(skinny_bwd_b16_tma_kernel<NB, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              kStreamStages * NB * kBwdTileRows * 512 + 64));                          \
                GMC_CUDA(cudaFuncSetAttribute(skinny_bwd_b16_tma_kernel<NB, K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
                attr = true;                                                                                           \
            }                                                                                                          \
            skinny_bwd_b16_tma_kernel<NB, K><<<nc, 128, smem, s>>>(tm, dT, lddt, W, dH2, lddh / 4, n_rows, n_in, rows_per, ws); \
        }
#define GMC_NB(K) if (nb == 2) GMC_CASE(2, K) else GMC_CASE(1, K)
        switch (n_out) {
            case 1: GMC_NB(1); break;
            case 2: GMC_NB(2); break;
            case 3: GMC_NB(3); break;
            case 4: GMC_NB(4); break;
            case 5: GMC_NB(5); break;
            case 6: GMC_NB(6); break;
            case 7: GMC_NB(7); break;
            default: GMC_NB(8); break;
        }
#undef GMC_NB
#undef GMC_CASE
        GMC_LAUNCH_CHECK();
        const int total = n_in * (n_out + 1);
        skinny_bwd_reduce_kernel<<<ceil_div(total, 32), dim3(32, 32), 0, s>>>(ws, nc, n_in, n_out, dW, dbias);
        GMC_LAUNCH_CHECK();
        return GMC_OK;
    }
#define GMC_CASE(K) case K: GMC_CUDA(launch_pdl(skinny_bwd_b16_kernel<K>, n_ctas, 128, 0, s, dT, lddt, W, H2, ldh / 4, dH2, lddh / 4, n_rows, n_in, ws)); break;
    switch (n_out) { GMC_CASE(1) GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    const int total = n_in * (n_out + 1);
    GMC_CUDA(launch_pdl(skinny_bwd_reduce_kernel, ceil_div(total, 32), dim3(32, 32), 0, s, ws, n_ctas, n_in, n_out, dW, dbias));
    return GMC_OK;
}

// gmc_skinny_bwd_f32 whose dHpre leaves as n_split (2 or 3) stacked bf16 parts of row_scale[v] * dHpre[v, :] (row_scale
// nullable = 1): part s occupies rows [s * split_rows, s * split_rows + n_rows) of the bf16 matrix dH_split (lddh in
// elements, multiple of 4).  H is fp32.  Feeds gmc_gemm_bf16_split (op tn) -- autograd of TrainingNeural.py:80-83 at fp32
// grade on bf16 tensor cores.  Same workspace as gmc_skinny_bwd_f32.
int gmc_skinny_bwd_split(const float* dT, int64_t lddt, const float* W, const float* H, int64_t ldh, const float* row_scale,
                         void* dH_split, int64_t lddh, int64_t split_rows, int32_t n_split, int32_t lo_shift, float* dW,
                         float* dbias, int64_t n_rows, int32_t n_in, int32_t n_out, void* workspace, size_t workspace_bytes,
                         void* stream) {
    using namespace gmc;
    GMC_REQUIRE(lo_shift >= 0 && lo_shift <= 24, "gmc_skinny_bwd_split: lo_shift must be 0 (bf16 parts) or 1..24 (fp16 parts)");
    const bool f16 = lo_shift > 0;
    const float lo_up = f16 ? (float)(1u << lo_shift) : 1.0f;
    GMC_REQUIRE(dT && W && H && dH_split && dW, "gmc_skinny_bwd_split: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_in > 0 && n_out >= 1 && n_out <= 4 && ldh >= n_in && lddh >= n_in && lddt >= n_out,
                "gmc_skinny_bwd_split: bad sizes (n_out must be 1..4)");
    GMC_REQUIRE((n_split == 2 || n_split == 3) && split_rows >= n_rows, "gmc_skinny_bwd_split: n_split must be 2 or 3 and split_rows >= n_rows");
    GMC_REQUIRE(n_in % 4 == 0 && ldh % 4 == 0 && lddh % 4 == 0 && aligned16(H) && (reinterpret_cast<uintptr_t>(dH_split) & 7u) == 0,
                "gmc_skinny_bwd_split: n_in and leading dimensions must be multiples of 4 with aligned bases");
    const int n_ctas = skinny_bwd_ctas(n_rows);
    const size_t need = (size_t)n_ctas * n_in * (n_out + 1) * sizeof(float);
    if (!workspace || workspace_bytes < need) {
        set_error("gmc_skinny_bwd_split: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GMC_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    float* ws = reinterpret_cast<float*>(workspace);
    if (n_rows == 0) {
        GMC_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * n_in * n_out, s));
        if (dbias) GMC_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * n_in, s));
        return GMC_OK;
    }
    uint2* dH2 = reinterpret_cast<uint2*>(dH_split);
    if (stream_kernels_enabled() && n_rows >= 65536 && n_rows < (1ll << 31) - 64 && n_in <= 512 && n_out <= 4) {
        // TMA-streamed kernel: 16 KB stages, three CTAs per SM, one wave; row ranges in multiples of the 8-row tile
        CUtensorMap tm;
        int rc = tma::make_f32_map(&tm, H, (uint64_t)n_in, (uint64_t)n_rows, (uint64_t)ldh, 128, kBwdTileRows, "gmc_skinny_bwd_split");
        if (rc != GMC_OK) return rc;
        const int nb = (n_in + 127) / 128;
        int nc = sm_count() * (nb <= 2 ? 6 : 3);
        if (nc > n_ctas) nc = n_ctas;
        const int64_t rows_per = ceil_div<int64_t>(ceil_div<int64_t>(n_rows, nc), kBwdTileRows) * kBwdTileRows;
        nc = (int)ceil_div<int64_t>(n_rows, rows_per);
        const size_t smem = (size_t)kStreamStages * nb * kBwdTileRows * 512 + 64;
#define GMC_CASE(NB, K, NSV)                                                                                           \
        {                                                                                                              \
            static bool attr = false;                                                                                  \
            if (!attr) {                                                                                               \
                GMC_CUDA(cudaFuncSetAttribute(skinny_bwd_split_tma_kernel<NB, K, NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              kStreamStages * NB * kBwdTileRows * 512 + 64));                          \
                GMC_CUDA(cudaFuncSetAttribute(skinny_bwd_split_tma_kernel<NB, K, NSV>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
                attr = true;                                                                                           \
            }                                                                                                          \
            if (f16) {                                                                                                 \
                GMC_CUDA(cudaFuncSetAttribute(skinny_bwd_split_tma_kernel<NB, K, NSV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              kStreamStages * NB * kBwdTileRows * 512 + 64));                          \
                skinny_bwd_split_tma_kernel<NB, K, NSV, true><<<nc, 128, smem, s>>>(tm, dT, lddt, W, row_scale, dH2, lddh / 4, split_rows, n_rows, n_in, rows_per, ws, lo_up); \
            } else                                                                                                     \
            skinny_bwd_split_tma_kernel<NB, K, NSV><<<nc, 128, smem, s>>>(tm, dT, lddt, W, row_scale, dH2, lddh / 4, split_rows, n_rows, n_in, rows_per, ws, lo_up); \
        }
#define GMC_NS(NB, K) if (n_split == 2) GMC_CASE(NB, K, 2) else GMC_CASE(NB, K, 3)
#define GMC_NB(K) switch (nb) { case 1: GMC_NS(1, K); break; case 2: GMC_NS(2, K); break; case 3: GMC_NS(3, K); break; default: GMC_NS(4, K); break; }
        switch (n_out) {
            case 1: GMC_NB(1); break;
            case 2: GMC_NB(2); break;
            case 3: GMC_NB(3); break;
            default: GMC_NB(4); break;
        }
#undef GMC_NB
#undef GMC_NS
#undef GMC_CASE
        GMC_LAUNCH_CHECK();
        const int total = n_in * (n_out + 1);
        skinny_bwd_reduce_kernel<<<ceil_div(total, 32), dim3(32, 32), 0, s>>>(ws, nc, n_in, n_out, dW, dbias);
        GMC_LAUNCH_CHECK();
        return GMC_OK;
    }
#define GMC_CASE(K)                                                                                                           \
    case K:                                                                                                                   \
        if (f16 && n_split == 2) GMC_CUDA(launch_pdl(skinny_bwd_split_kernel<K, 2, true>, n_ctas, 128, 0, s, dT, lddt, W, H, ldh, row_scale, dH2, lddh / 4, split_rows, n_rows, n_in, ws, lo_up)); \
        else if (f16) GMC_CUDA(launch_pdl(skinny_bwd_split_kernel<K, 3, true>, n_ctas, 128, 0, s, dT, lddt, W, H, ldh, row_scale, dH2, lddh / 4, split_rows, n_rows, n_in, ws, lo_up)); \
        else if (n_split == 2) GMC_CUDA(launch_pdl(skinny_bwd_split_kernel<K, 2>, n_ctas, 128, 0, s, dT, lddt, W, H, ldh, row_scale, dH2, lddh / 4, split_rows, n_rows, n_in, ws, lo_up)); \
        else GMC_CUDA(launch_pdl(skinny_bwd_split_kernel<K, 3>, n_ctas, 128, 0, s, dT, lddt, W, H, ldh, row_scale, dH2, lddh / 4, split_rows, n_rows, n_in, ws, lo_up)); \
        break;
    switch (n_out) { GMC_CASE(1) GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    const int total = n_in * (n_out + 1);
    GMC_CUDA(launch_pdl(skinny_bwd_reduce_kernel, ceil_div(total, 32), dim3(32, 32), 0, s, ws, n_ctas, n_in, n_out, dW, dbias));
    return GMC_OK;
}

size_t gmc_colsum_workspace_bytes(int32_t n_cols) { return (size_t)gmc::reduce_ctas() * (size_t)n_cols * sizeof(float); }

int gmc_colsum_f32(const float* X, int64_t ldx, int64_t n_rows, int32_t n_cols, float* out, void* workspace,
                   size_t workspace_bytes, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(X && out, "gmc_colsum_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_colsum_f32: bad sizes");
    int n_ctas = reduce_ctas();
    if ((int64_t)n_ctas > n_rows) n_ctas = (int)(n_rows > 0 ? n_rows : 1);
    const size_t need = (size_t)n_ctas * n_cols * sizeof(float);
    if (!workspace || workspace_bytes < need) {
        set_error("gmc_colsum_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GMC_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    if (n_rows == 0) { GMC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_cols, s)); return GMC_OK; }
    float* ws = reinterpret_cast<float*>(workspace);
    colsum_stage1<<<n_ctas, 256, 0, s>>>(X, ldx, n_rows, n_cols, ws);
    GMC_LAUNCH_CHECK();
    colsum_stage2<<<ceil_div(n_cols, 256), 256, 0, s>>>(ws, n_ctas, n_cols, out);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

}  // extern "C"
