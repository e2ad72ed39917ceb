// gemm_tcgen05.cu -- (b) dense feature transforms on the 5th-generation tensor cores.
//
//   C[M,N] (+)= op(A) * op(B),  fp32 or bf16 operands in HBM, kind::tf32 / kind::f16 MMA, fp32 accumulation in TMEM,
//   fp32 output.
//     nn  X[M,K] * W1[K,N]          A K-major,  B MN-major      (GraphConv th.matmul, TrainingNeural.py:80)
//     tn  X[K,M]^T * dT1[K,N]       A MN-major, B MN-major      (dW1; K = all nodes of the batch -> split-K)
//     nt  dT1[M,K] * W1[N,K]^T      A K-major,  B K-major       (dX; only for trainable features)
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B boxes) into a 4-stage smem ring,
//               completion counted on mbarriers (expect_tx)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer: 128x256x8 TF32 MMAs, smem operands via
//               UMMA descriptors, tcgen05.commit releases smem stages / publishes accumulators
//   warps 2..5  epilogue: tcgen05.ld (32x32b.x32) TMEM -> registers -> global; two 256-column TMEM
//               accumulators so the epilogue of tile i overlaps the MMAs of tile i+1
//
// Tiles: BLOCK_M=128, BLOCK_N=256, BLOCK_K=32 fp32 (= one 128-byte swizzle row); 48 KB per stage.
// Out-of-range rows/columns/k are zero-filled by TMA and masked in the epilogue, so M, N, K are arbitrary
// (leading dimensions must be multiples of 4 floats: TMA needs 16-byte strides).
//
// Precision: GMC_GEMM_TF32 is one pass (operands truncated to 10 mantissa bits by the MMA); GMC_GEMM_TF32X3 three
// passes with materialised low-order parts (fp32-grade); gmc_gemm_bf16 takes bf16 operands (64-element stages, the same
// byte geometry: half the operand traffic per flop, twice the MMA rate).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gmc {
namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 32, UMMA_K = 8;
constexpr int STAGES = 4, ACC_STAGES = 2, TMEM_COLS = 512;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;              // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;              // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int CHUNK_BYTES = BLOCK_K * 128;                  // one MN-major box: 32 k-rows x 128 B
constexpr int THREADS = 192;
constexpr int EPI_BUF_BYTES = 32 * 32 * 4;                  // one 32-row x 32-column output block per epilogue warp
constexpr int EPI_BYTES = 4 /*warps*/ * 2 /*buffers*/ * EPI_BUF_BYTES;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (error returned to the host) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("gmc tcgen05 gemm: mbarrier timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask`, and each of those
// CTAs' mbarriers (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// two fp32 -> one word of two bf16 (round to nearest even); lo lands at the lower address
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t w;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    return w;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout), version 1.
// layout_type: 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B -- the only layout the
// hardware accepts for MN-major 32-bit (tf32) operands (4 k-rows x 128 B atoms, 32-byte swizzle chunks;
// the matching TMA mode is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, bool bf16 = false, int mma_n = BLOCK_N) {
    return (1u << 4)                              // D format  : F32
         | ((bf16 ? 1u : 2u) << 7) | ((bf16 ? 1u : 2u) << 10)   // A/B format: TF32 = 2 (kind::tf32), BF16 = 1 (kind::f16)
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

struct Params {
    float* C;                // output (or split-K workspace)
    int64_t M, N, K, ldc;
    int64_t split_stride;    // elements between split-K partials (0 when k_splits == 1)
    int64_t k_per_split;     // multiple of BLOCK_K
    int m_tiles, n_tiles, k_splits;
    int mg_tiles;            // m-tile groups of CLM tiles: ceil(m_tiles / CLM)
    int ng_tiles;            // n-tile groups of CLN tiles: ceil(n_tiles / CLN)
    int accumulate;
    int vec_ok;              // 16-byte aligned C rows
    int tma_store;           // epilogue through shared memory + cp.async.bulk.tensor stores (tmC valid)
    int c_bf16;              // C is a bf16 matrix (tmC describes it): accumulators rounded to nearest even on the way out
    const float* bias;       // c_bf16 only: fp32 bias[N] added to every row before the rounding (nullable)
    int relu;                // c_bf16 only: max(., 0) after the bias
    // c_bf16 only, optional skinny projection of the ROUNDED result: P[m, k] += sum_n bf16(C[m, n]) * projW[n][k], k < 4.
    // projW is [N rounded up to 64][4] fp32 (zero padded), P [M, ldp] fp32 zeroed by the launcher; the (at most two)
    // n-tiles of a row add their partial sums with atomics -- two addends into zero commute, so P is reproducible.
    const float4* proj_w;
    float* proj_out;
    int64_t ldp;
    int proj_k;
    // split-operand GEMMs (NS > 1): B is NS stacked bf16 matrices (hi, lo [, lo2] parts of an fp32 operand), part s
    // starting b_split_rows rows below part 0; the MMA sees them side by side along N and the epilogue adds them up
    int64_t b_split_rows;
    // fp32 epilogue (direct, non split-K outputs): C[m, n] = act(row_scale[m] * acc + bias[n]); each nullable / 0
    const float* row_scale;
    // split operands stored as fp16 (b_f16 = 1: A AND B are IEEE fp16 -- kind::f16 wants one 16-bit format for both operands,
    // a bf16 A with an fp16 B raises an illegal-instruction fault on B200) with the low-order parts scaled up by
    // 1 / part_scale per level (they would be fp16 subnormals otherwise); the epilogue multiplies part s by part_scale^s
    // before adding.  part_scale = 1 for bf16 parts.
    int b_f16;
    float part_scale;
    // fp32 epilogue + proj_w: every (row, n-tile) writes its partial projection sum_n C[m, n] projW[n][0..3] to
    // proj_part[n_tile][m] (one float4, written exactly once: deterministic); proj_reduce_kernel adds the n-tiles up
    float4* proj_part;
};

// One accumulator tile (this warp's 32 rows x RN real columns) from TMEM to C: the epilogue shared by the cluster kernel and
// the cta_group::2 kernel.  m0 = first row of this CTA's 128-row block, n0 = first real column, t_row = TMEM address of
// the warp's lane quarter in the accumulator stage, epi = the CTA's staging block, ebuf = the warp's buffer toggle.
// SC: bias and projection weights are read from the CTA's shared-memory copy `sconst` (kEpiConstCols bias floats, then
// kEpiConstCols float4 rows of projW, zero beyond N) instead of through __ldg.  The kernels run with ~225 KB of the SM's
// 256 KB as shared memory, so the "L1-resident" broadcast loads were L2 round trips: ncu (profiles/r02g_nn_epilogue.md)
// put 36 % of the epilogue warps' stall samples on the first use of those loads, and the epilogue -- one warp per
// scheduler -- is what a tile of the fused layer-1 GEMM waits for (tensor pipe 69 % active against 85 % without it).
constexpr int kEpiConstCols = 512;
constexpr int kEpiConstBytes = kEpiConstCols * 4 + kEpiConstCols * 16;

// EBUFS: staging buffers per epilogue warp (2 = the next chunk is staged while the TMA store of the previous one drains;
// 1 where the 10 KB of constants need the space: the NS = 2 cluster kernel, whose four 48 KB stages fill the SM).
template <int NS, bool SC = false, int EBUFS = 2>
__device__ __forceinline__ void epilogue_tile(const Params& p, const CUtensorMap* tmC_ptr, uint32_t epi, int warp, int lane,
                                              int q, int64_t m0, int n0, uint32_t t_row, float* Cs, int& ebuf,
                                              const float* sconst = nullptr) {
    constexpr int BN = NS == 3 ? 192 : BLOCK_N;
    constexpr int RN = BN / NS;
    const CUtensorMap& tmC = *tmC_ptr;
    const int64_t m = m0 + 32 * q + lane;
        if (NS == 1 && p.c_bf16) {
            // bf16 output: 64 accumulator columns -> 32 packed words = one 128-byte staging row per lane, same swizzle
            // and the same 32-row TMA store as the fp32 path (box 64 x 32 bf16); half the store traffic of the tile
            // projection accumulators as two fp32x2 pairs: packed FMAs (fma.rn.f32x2) halve the FMA instructions of the
            // fused projection -- ncu (profiles/r02d_ncu_full_summary.csv): with the projection the epilogue warps issue
            // 31 % of all slots and the tensor pipe drops from 85 % to 69 % active, i.e. the epilogue is what the tile waits for
            float2 pj01 = make_float2(0.f, 0.f), pj23 = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 64; ++c) {
                const int n = n0 + 64 * c;
                if (n >= p.N || m0 >= p.M) break;              // warp-uniform
                uint32_t ra[32], rb[32], wv[32];
                tmem_ld32(t_row + 64 * c, ra);
                tmem_ld32(t_row + 64 * c + 32, rb);
                if (p.bias) {                                  // same address in every lane: broadcast loads, L1-resident
                    const float4* b4 = reinterpret_cast<const float4*>((SC ? sconst : p.bias) + n);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (n + 4 * i < p.N) {                 // N % 4 == 0 (checked by the launcher)
                            const float4 b = SC ? b4[i] : __ldg(b4 + i);
                            ra[4 * i] = __float_as_uint(__uint_as_float(ra[4 * i]) + b.x);
                            ra[4 * i + 1] = __float_as_uint(__uint_as_float(ra[4 * i + 1]) + b.y);
                            ra[4 * i + 2] = __float_as_uint(__uint_as_float(ra[4 * i + 2]) + b.z);
                            ra[4 * i + 3] = __float_as_uint(__uint_as_float(ra[4 * i + 3]) + b.w);
                        }
                        if (n + 32 + 4 * i < p.N) {
                            const float4 b = SC ? b4[8 + i] : __ldg(b4 + 8 + i);
                            rb[4 * i] = __float_as_uint(__uint_as_float(rb[4 * i]) + b.x);
                            rb[4 * i + 1] = __float_as_uint(__uint_as_float(rb[4 * i + 1]) + b.y);
                            rb[4 * i + 2] = __float_as_uint(__uint_as_float(rb[4 * i + 2]) + b.z);
                            rb[4 * i + 3] = __float_as_uint(__uint_as_float(rb[4 * i + 3]) + b.w);
                        }
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        ra[i] = __float_as_uint(fmaxf(__uint_as_float(ra[i]), 0.f));
                        rb[i] = __float_as_uint(fmaxf(__uint_as_float(rb[i]), 0.f));
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    wv[i] = pack_bf16x2(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]));
                    wv[16 + i] = pack_bf16x2(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]));
                }
                if (p.proj_w) {                                // this row's 64 rounded values against projW[n .. n+63][0..3]
                    // same address in every lane: one broadcast wavefront (shared) / one sector (global)
                    const float4* w4 = (SC ? reinterpret_cast<const float4*>(sconst + kEpiConstCols) : p.proj_w) + n;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float lo = __uint_as_float(wv[i] << 16), hi = __uint_as_float(wv[i] & 0xffff0000u);
                        const float4 wa = SC ? w4[2 * i] : __ldg(w4 + 2 * i), wb = SC ? w4[2 * i + 1] : __ldg(w4 + 2 * i + 1);
                        const float2 l2 = make_float2(lo, lo), h2 = make_float2(hi, hi);
                        pj01 = __ffma2_rn(l2, make_float2(wa.x, wa.y), pj01);
                        pj23 = __ffma2_rn(l2, make_float2(wa.z, wa.w), pj23);
                        pj01 = __ffma2_rn(h2, make_float2(wb.x, wb.y), pj01);
                        pj23 = __ffma2_rn(h2, make_float2(wb.z, wb.w), pj23);
                    }
                }
                const uint32_t buf = epi + (uint32_t)((warp - 2) * EBUFS + (EBUFS == 2 ? ebuf : 0)) * EPI_BUF_BYTES;
                if (lane == 0) bulk_wait_read<EBUFS - 1>();
                __syncwarp();
                const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((i ^ (lane & 7)) << 4)),
                                 "r"(wv[4 * i]), "r"(wv[4 * i + 1]), "r"(wv[4 * i + 2]), "r"(wv[4 * i + 3]) : "memory");
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, buf, n, (int)(m0 + 32 * q));
                    bulk_commit();
                }
                ebuf ^= 1;
            }
            if (p.proj_w && m < p.M && m0 < p.M && n0 < p.N) {
                float* po = p.proj_out + m * p.ldp;
                const float pj[4] = {pj01.x, pj01.y, pj23.x, pj23.y};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < p.proj_k) atomicAdd(po + k, pj[k]);
            }
        } else {
        float2 pj01 = make_float2(0.f, 0.f), pj23 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = 0; c < RN / 32; ++c) {
            const int n = n0 + 32 * c;
            if (n >= p.N || m0 >= p.M) break;                  // warp-uniform
            uint32_t r[32];
            if (NS == 1) {
                tmem_ld32(t_row + 32 * c, r);
            } else {
                // split operand: column n of the product = sum of the NS accumulator groups (low-order parts first)
                uint32_t r2[32];
                tmem_ld32(t_row + (NS - 1) * RN + 32 * c, r);
                const float ps = p.part_scale;                     // Horner: ((p_{NS-1} ps + p_{NS-2}) ps + ...) + p_0
#pragma unroll
                for (int sp = NS - 2; sp >= 0; --sp) {
                    tmem_ld32(t_row + sp * RN + 32 * c, r2);
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(fmaf(__uint_as_float(r[i]), ps, __uint_as_float(r2[i])));
                }
            }
            if (p.row_scale) {
                const float rs = m < p.M ? __ldg(p.row_scale + m) : 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * rs);
            }
            if (p.bias) {                                      // same address in every lane: broadcast loads
                const float4* b4 = reinterpret_cast<const float4*>((SC ? sconst : p.bias) + n);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (n + 4 * i < p.N) {                     // N % 4 == 0 (checked by the launcher)
                        const float4 b = SC ? b4[i] : __ldg(b4 + i);
                        r[4 * i] = __float_as_uint(__uint_as_float(r[4 * i]) + b.x);
                        r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + b.y);
                        r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + b.z);
                        r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + b.w);
                    }
                }
            }
            if (p.relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.f));
            }
            if (p.proj_part) {                                 // this row's 32 values against projW[n .. n+31][0..3]
                // same address in every lane: broadcast; rows >= N of projW are zero padding (up to N rounded to 64)
                const float4* w4 = (SC ? reinterpret_cast<const float4*>(sconst + kEpiConstCols) : p.proj_w) + n;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = __uint_as_float(r[i]);
                    const float4 w = SC ? w4[i] : __ldg(w4 + i);
                    const float2 v2 = make_float2(v, v);
                    pj01 = __ffma2_rn(v2, make_float2(w.x, w.y), pj01);
                    pj23 = __ffma2_rn(v2, make_float2(w.z, w.w), pj23);
                }
            }
            if (p.tma_store) {
                // registers -> 128B-swizzled staging block (lane = row, 16-byte chunk i at i ^ (row & 7): conflict-free
                // STS.128) -> one TMA store of 32 full 128-byte row segments; rows / columns beyond M / N are clipped
                const uint32_t buf = epi + (uint32_t)((warp - 2) * EBUFS + (EBUFS == 2 ? ebuf : 0)) * EPI_BUF_BYTES;
                if (lane == 0) bulk_wait_read<EBUFS - 1>();    // the store that last read this buffer has drained it
                __syncwarp();
                const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((i ^ (lane & 7)) << 4)),
                                 "r"(r[4 * i]), "r"(r[4 * i + 1]), "r"(r[4 * i + 2]), "r"(r[4 * i + 3]) : "memory");
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, buf, n, (int)(m0 + 32 * q));
                    bulk_commit();
                }
                ebuf ^= 1;
            } else if (m < p.M) {
                float* dst = Cs + m * p.ldc + n;
                if (p.vec_ok && n + 32 <= p.N) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                               __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                        float4* d4 = reinterpret_cast<float4*>(dst) + i;
                        if (p.accumulate) { const float4 o = *d4; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        *d4 = v;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n + i < p.N) dst[i] = p.accumulate ? dst[i] + __uint_as_float(r[i]) : __uint_as_float(r[i]);
                }
            }
        }
        if (p.proj_part && m < p.M && m0 < p.M && n0 < p.N)
            p.proj_part[(int64_t)(n0 / RN) * p.M + m] = make_float4(pj01.x, pj01.y, pj23.x, pj23.y);
        }
}

// Thread-block cluster of CLM x CLN CTAs (rank = rm + CLM * rn) that owns CLM consecutive m-tiles x CLN consecutive
// n-tiles of one k-range.  The A tile of m-tile rm is needed by the CLN CTAs of that row: each loads 1/CLN of it and
// multicasts it to the row; the B tile of n-tile rn is needed by the CLM CTAs of that column: each loads 1/CLM and
// multicasts it to the column.  Why: the kernel is bound by L2 (LTS) throughput -- SM reads, DRAM fills and
// write-backs all count (config 3 nn, single CTA: 97.6 GB of SM reads in 9.7 ms; an L2 prefetch experiment made it
// slower).  SM reads per output scale with 1/(256 CLN) + 1/(128 CLM).  A stage may be refilled only when every CTA
// that receives one of this CTA's multicasts has consumed it, and symmetrically, so each MMA warp multicasts its
// tcgen05.commit to its row and column (arrival count CLM + CLN - 1).
//
// NS > 1 (bf16, MN-major B only): fp32-grade product with a SPLIT B operand.  B = B_0 + B_1 [+ B_2] with bf16 parts
// (hi / lo / lo2 of an fp32 matrix, 8 mantissa bits each), A exact in bf16 (small integers).  A tile covers RN = BN / NS
// real output columns; its B stage holds the NS parts of those columns side by side (BN = 256 for NS = 2, 192 for NS = 3),
// ONE MMA per k-step multiplies the A tile with all parts (the A tile is loaded once for NS products) and the epilogue
// adds the NS accumulator column groups.  Everything else -- ring, barriers, multicast, split-K -- is unchanged.
template <bool A_MN, bool B_MN, int CLM, int CLN, bool BF16, int NS = 1>
__global__ void __launch_bounds__(THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const Params p) {
    static_assert(NS == 1 || (BF16 && B_MN && CLN == 1), "split operands: bf16, MN-major B, 1-D clusters");
    constexpr int CL = CLM * CLN;
    constexpr int BN = NS == 3 ? 192 : BLOCK_N;                    // MMA N = columns of the B stage
    constexpr int RN = BN / NS;                                    // real output columns per tile
    constexpr int B_BYTES_T = BN * 128;                            // B stage bytes (BN columns x 128 B of k)
    constexpr int STAGE_T = A_BYTES + B_BYTES_T;
    // element-size dependent geometry; everything in BYTES is the same for fp32 (tf32) and bf16 operands: a stage row
    // is 128 B = 32 fp32 or 64 bf16 along k (K-major) or along m/n (MN-major chunk), one MMA eats 32 B of k
    constexpr int ELT = BF16 ? 2 : 4;
    constexpr int BK = 128 / ELT;                                  // k elements per stage
    constexpr int MNC = 128 / ELT;                                 // m/n elements per MN-major chunk
    constexpr int UK = 32 / ELT;                                   // k elements per MMA
    constexpr int CHUNK = BK * 128;                                // bytes of one MN-major chunk (BK k-rows x 128 B)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B atoms need 1024-byte alignment
    // split-operand kernels (NS > 1) carry the fused layer-1 epilogue: their bias / projection weights live in shared
    // memory behind the barriers (see epilogue_tile).  NS = 3 stages are 40 KB, so there is room; NS = 2 stages are 48 KB
    // and the staging block drops to one buffer per warp instead.
    constexpr bool kConsts = NS > 1;
    constexpr int EBUFS = NS == 2 ? 1 : 2;
    const uint32_t epi = tiles + STAGES * STAGE_T;                 // epilogue staging: 4 warps x EBUFS x 4 KB, 1024-aligned
    const uint32_t bars = epi + 4 * EBUFS * EPI_BUF_BYTES;
    static_assert(!kConsts || (size_t)STAGES * STAGE_T + 4 * EBUFS * EPI_BUF_BYTES + 256 + kEpiConstBytes + 1024 <= SMEM_BYTES,
                  "no room for the epilogue constants");
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
    const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 8 * ACC_STAGES;
    const uint32_t tmem_slot = tempty_bar + 8 * ACC_STAGES;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
    const uint32_t rm = rank % CLM, rn = rank / CLM;
    // row = the CLN CTAs that share my A tile, column = the CLM CTAs that share my B tile
    uint32_t row_mask = 0, col_mask = 0;
#pragma unroll
    for (int j = 0; j < CLN; ++j) row_mask |= 1u << (rm + CLM * j);
#pragma unroll
    for (int i = 0; i < CLM; ++i) col_mask |= 1u << (i + CLM * rn);
    const uint16_t MASK_A = (uint16_t)row_mask, MASK_B = (uint16_t)col_mask, MASK_ALL = (uint16_t)(row_mask | col_mask);

    if (warp == 0 && lane == 0) {
        prefetch_map(&tmA);
        prefetch_map(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, CLM + CLN - 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();          // peers' barriers are initialised before any multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // programmatic dependent launch: barrier init, TMEM allocation and descriptor prefetch above overlap the previous
    // kernel's tail; nothing before this line touches global memory
    pdl_wait();
    pdl_launch_dependents();
    const float* sconst = nullptr;
    if (kConsts && (p.bias || p.proj_w) && p.N <= kEpiConstCols) {  // uniform over the grid
        float* sb = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));
        float4* sp = reinterpret_cast<float4*>(sb + kEpiConstCols);
        const int prow = (int)((p.N + 63) / 64 * 64);              // rows of the padded projection matrix
        for (int i = threadIdx.x; i < kEpiConstCols; i += THREADS) {
            sb[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
            sp[i] = (p.proj_w && i < prow) ? __ldg(p.proj_w + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        sconst = sb;
        __syncthreads();
    }

    // cluster-level work items: (split, m-group, n-group); every CTA of a cluster walks the same list
    const int64_t tiles_gn = (int64_t)p.mg_tiles * p.ng_tiles;
    const int64_t n_work = tiles_gn * p.k_splits;
    const int64_t cw0 = blockIdx.x / CL, cw_step = gridDim.x / CL;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t w = cw0; w < n_work; w += cw_step) {
                const int split = (int)(w / tiles_gn);
                const int64_t rem = w - (int64_t)split * tiles_gn;
                const int m0 = ((int)(rem / p.ng_tiles) * CLM + (int)rm) * BLOCK_M;
                const int n0 = ((int)(rem % p.ng_tiles) * CLN + (int)rn) * RN;
                const int64_t kb = (int64_t)split * p.k_per_split;
                const int64_t ke = min(p.K, kb + p.k_per_split);
                for (int64_t k0 = kb; k0 < ke; k0 += BK) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = tiles + stage * STAGE_T, sb = sa + A_BYTES;
                    const uint32_t fb = full_bar + 8 * stage;
                    mbar_expect_tx(fb, STAGE_T);
                    if (CLN == 1) {
                        if (A_MN) {
#pragma unroll
                            for (int j = 0; j < BLOCK_M / MNC; ++j) tma_load_2d(sa + j * CHUNK, &tmA, m0 + MNC * j, (int)k0, fb);
                        } else {
                            tma_load_2d(sa, &tmA, (int)k0, m0, fb);
                        }
                    } else if (A_MN) {                             // my share of the 4 row chunks, to my cluster row
                        constexpr int PER = BLOCK_M / MNC / CLN;
#pragma unroll
                        for (int jj = 0; jj < PER; ++jj) {
                            const int j = (int)rn * PER + jj;
                            tma_load_2d_mc(sa + j * CHUNK, &tmA, m0 + MNC * j, (int)k0, fb, MASK_A);
                        }
                    } else {                                       // K-major A: my BLOCK_M / CLN rows, to my cluster row
                        constexpr int ROWS = BLOCK_M / CLN;
                        tma_load_2d_mc(sa + rn * (ROWS * 128), &tmA, (int)k0, m0 + (int)rn * ROWS, fb, MASK_A);
                    }
                    // chunk j of the B stage = column chunk (j % CPS) of split part (j / CPS): part s lives b_split_rows * s
                    // rows below part 0 in the stacked operand (NS == 1: CPS = all chunks, part 0 only)
                    constexpr int NCH = BN / MNC, CPS = NCH / NS;
                    if (CLM == 1) {
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < NCH; ++j)
                                tma_load_2d(sb + j * CHUNK, &tmB, n0 + MNC * (j % CPS),
                                            (int)(k0 + (NS > 1 ? (int64_t)(j / CPS) * p.b_split_rows : 0)), fb);
                        } else {
                            tma_load_2d(sb, &tmB, (int)k0, n0, fb);
                        }
                    } else if (B_MN) {                             // my share of the column chunks, to my cluster column
                        static_assert(NS == 1 || NCH % CLM == 0, "B chunks must divide over the cluster column");
                        constexpr int PER = NCH / CLM;
#pragma unroll
                        for (int jj = 0; jj < PER; ++jj) {
                            const int j = (int)rm * PER + jj;
                            tma_load_2d_mc(sb + j * CHUNK, &tmB, n0 + MNC * (j % CPS),
                                           (int)(k0 + (NS > 1 ? (int64_t)(j / CPS) * p.b_split_rows : 0)), fb, MASK_B);
                        }
                    } else {                                       // K-major B: my BLOCK_N / CLM rows, to my cluster column
                        constexpr int ROWS = BLOCK_N / CLM;
                        tma_load_2d_mc(sb + rm * (ROWS * 128), &tmB, (int)k0, n0 + (int)rm * ROWS, fb, MASK_B);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // A / B format fields (bits 7-9, 10-12): BF16 = 1 -> F16 = 0 for fp16 operands
        const uint32_t idesc = make_idesc(A_MN, B_MN, BF16, BN) & ~((NS > 1 && p.b_f16) ? ((1u << 7) | (1u << 10)) : 0u);
        // K-major : SWIZZLE_128B, rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1 = CUTLASS convention)
        // MN-major: SWIZZLE_128B_BASE32B, 32-element chunks 4096 B apart (LBO), 4-k-row atoms 512 B apart (SBO)
        // 16-bit MN-major: plain SWIZZLE_128B, 64-element chunks CHUNK apart (LBO), 8-k-row atoms 1024 B apart (SBO)
        // (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
        constexpr uint32_t MN_SBO = BF16 ? 1024 : 512, MN_LAY = BF16 ? 2 : 1, MN_STEP = UK * 128;
        const uint32_t a_lbo = A_MN ? CHUNK : 16, b_lbo = B_MN ? CHUNK : 16;
        const uint32_t a_sbo = A_MN ? MN_SBO : 1024, b_sbo = B_MN ? MN_SBO : 1024;
        const uint32_t a_lay = A_MN ? MN_LAY : 2, b_lay = B_MN ? MN_LAY : 2;
        const uint32_t a_step = A_MN ? MN_STEP : 32, b_step = B_MN ? MN_STEP : 32;
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int64_t w = cw0; w < n_work; w += cw_step) {
            const int split = (int)(w / tiles_gn);
            const int64_t kb = (int64_t)split * p.k_per_split;
            const int64_t ke = min(p.K, kb + p.k_per_split);
            mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            uint32_t first = 1;
            for (int64_t k0 = kb; k0 < ke; k0 += BK) {
                mbar_wait(full_bar + 8 * stage, phase);
                tc_fence_after();
                __syncwarp();
                if (elect_one()) {
                    const uint32_t sa = tiles + stage * STAGE_T, sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint64_t ad = make_desc(sa + k * a_step, a_lbo, a_sbo, a_lay);
                        const uint64_t bd = make_desc(sb + k * b_step, b_lbo, b_sbo, b_lay);
                        if (BF16) umma_bf16(d_tmem, ad, bd, idesc, first ? 0u : 1u); else umma_tf32(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                        first = 0;
                    }
                    // smem stage reusable once these MMAs retire -- in every CTA of the cluster
                    if (CL == 1) umma_commit(empty_bar + 8 * stage); else umma_commit_mc(empty_bar + 8 * stage, MASK_ALL);
                    if (k0 + BK >= ke) umma_commit(tfull_bar + 8 * acc);
                }
                __syncwarp();
                first = 0;
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                                    // TMEM lane quarter this warp may access
        int acc = 0; uint32_t acc_phase = 0;
        int ebuf = 0;
        for (int64_t w = cw0; w < n_work; w += cw_step) {
            const int split = (int)(w / tiles_gn);
            const int64_t rem = w - (int64_t)split * tiles_gn;
            const int64_t m0 = ((rem / p.ng_tiles) * CLM + rm) * BLOCK_M;
            const int n0 = ((int)(rem % p.ng_tiles) * CLN + (int)rn) * RN;
            float* Cs = p.C + (int64_t)split * p.split_stride;
            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + acc * BLOCK_N;
            if (kConsts && sconst) epilogue_tile<NS, true, EBUFS>(p, &tmC, epi, warp, lane, q, m0, n0, t_row, Cs, ebuf, sconst);
            else epilogue_tile<NS, false, EBUFS>(p, &tmC, epi, warp, lane, q, m0, n0, t_row, Cs, ebuf);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) bulk_wait_all();                            // outstanding TMA stores complete before the CTA exits
    }

    __syncwarp();                                                  // reconverge before the aligned barriers
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();          // nobody leaves while a peer may still multicast to us
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// =================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x 256 tile.  Each CTA stages its own
// 128 rows of A and its own 128 columns of B (32 KB per stage instead of 48 KB for the same MMA work), the
// leader CTA issues M=256 MMAs that read both CTAs' shared memory and write both CTAs' TMEM.  The single-CTA
// kernel is L2->SM bandwidth bound (48 KB per 512 MMA cycles = the ~42 B/clk/SM LTS limit); pairing cuts the
// operand traffic per SM by a third.
// =================================================================================================
constexpr int STAGES2 = 5;                                  // 6 before the epilogue constants moved into shared memory
constexpr int HALF = 128;                                   // rows of A / columns of B staged per CTA
constexpr int A2_BYTES = HALF * BLOCK_K * 4, B2_BYTES = HALF * BLOCK_K * 4, STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr size_t SMEM2_BYTES = (size_t)STAGES2 * STAGE2_BYTES + EPI_BYTES + 1024 + 256 + kEpiConstBytes;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// remote arrive with the DEFAULT (.release.cta) semantics, as CUTLASS' ClusterBarrier::arrive(cta_id) does.  The
// explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR per call; issued once per k-block by the peer
// producer it cost ~1750 cycles per k-block (ncu source page, profiles/r01_gemm_notes.md) -- the 2-CTA "stall".
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc2(bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);          // N = 256, M = 256 (pair)
}

__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc2x(bool a_mn, bool b_mn, bool bf16) {
    return (1u << 4) | ((bf16 ? 1u : 2u) << 7) | ((bf16 ? 1u : 2u) << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);          // N = 256, M = 256 (pair)
}

// fp32 (kind::tf32) or bf16 (kind::f16) operands; the epilogue is the cluster kernel's (epilogue_tile: fp32 / bf16 C, TMA
// stores through the swizzled staging block, bias / ReLU / row scale / fused projection), run by both CTAs on their own
// 128 rows.  Five 32 KB stages + the 32 KB staging block + 10 KB of epilogue constants (bias, projection weights).
template <bool A_MN, bool B_MN, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_umma2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const Params p) {
    constexpr int ELT = BF16 ? 2 : 4;
    constexpr int BK = 128 / ELT;                                  // k elements per stage
    constexpr int MNC = 128 / ELT;                                 // m/n elements per MN-major chunk
    constexpr int UK = 32 / ELT;                                   // k elements per MMA
    constexpr int CHUNK = BK * 128;                                // bytes of one MN-major chunk
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t epi = tiles + STAGES2 * STAGE2_BYTES;           // 4 warps x 2 x 4 KB, 1024-aligned
    const uint32_t bars = epi + EPI_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES2;
    const uint32_t tfull_bar = bars + 16 * STAGES2, tempty_bar = tfull_bar + 8 * ACC_STAGES;
    const uint32_t tmem_slot = tempty_bar + 8 * ACC_STAGES;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_map(&tmA);
        prefetch_map(&tmB);
        // full: ONE arrival (the leader's expect_tx for both CTAs' bytes); the peer's TMA only adds complete_tx
        for (int s = 0; s < STAGES2; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                                                    // see gemm_umma_kernel
    pdl_launch_dependents();

    // bias / projection weights of all <= 512 columns into shared memory, once per CTA (see epilogue_tile)
    const float* sconst = nullptr;
    if ((p.bias || p.proj_w) && p.N <= kEpiConstCols) {            // uniform over the grid
        float* sb = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));
        float4* sp = reinterpret_cast<float4*>(sb + kEpiConstCols);
        const int prow = (int)((p.N + 63) / 64 * 64);             // rows of the padded projection matrix
        for (int i = threadIdx.x; i < kEpiConstCols; i += THREADS) {
            sb[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
            sp[i] = (p.proj_w && i < prow) ? __ldg(p.proj_w + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        sconst = sb;
        __syncthreads();
    }

    const int64_t tiles_mn = (int64_t)p.m_tiles * p.n_tiles;       // tiles of 256 x 256
    const int64_t n_work = tiles_mn * p.k_splits;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs, own halves) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t w = pair; w < n_work; w += n_pairs) {
                const int split = (int)(w / tiles_mn);
                const int64_t rem = w - (int64_t)split * tiles_mn;
                const int m0 = (int)(rem / p.n_tiles) * 256 + HALF * (int)rank;
                const int n0 = (int)(rem % p.n_tiles) * 256 + HALF * (int)rank;
                const int64_t kb = (int64_t)split * p.k_per_split;
                const int64_t ke = min(p.K, kb + p.k_per_split);
                for (int64_t k0 = kb; k0 < ke; k0 += BK) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = tiles + stage * STAGE2_BYTES, sb = sa + A2_BYTES;
                    const uint32_t fb = mapa_u32(full_bar + 8 * stage, 0);          // the LEADER's full barrier
                    if (leader) mbar_expect_tx(full_bar + 8 * stage, 2 * STAGE2_BYTES);
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < HALF / MNC; ++j) tma_load_2d_2sm(sa + j * CHUNK, &tmA, m0 + MNC * j, (int)k0, fb);
                    } else {
                        tma_load_2d_2sm(sa, &tmA, (int)k0, m0, fb);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < HALF / MNC; ++j) tma_load_2d_2sm(sb + j * CHUNK, &tmB, n0 + MNC * j, (int)k0, fb);
                    } else {
                        tma_load_2d_2sm(sb, &tmB, (int)k0, n0, fb);
                    }
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t idesc = make_idesc2x(A_MN, B_MN, BF16);
            constexpr uint32_t MN_SBO = BF16 ? 1024 : 512, MN_LAY = BF16 ? 2 : 1, MN_STEP = UK * 128;
            const uint32_t a_lbo = A_MN ? CHUNK : 16, b_lbo = B_MN ? CHUNK : 16;
            const uint32_t a_sbo = A_MN ? MN_SBO : 1024, b_sbo = B_MN ? MN_SBO : 1024;
            const uint32_t a_lay = A_MN ? MN_LAY : 2, b_lay = B_MN ? MN_LAY : 2;
            const uint32_t a_step = A_MN ? MN_STEP : 32, b_step = B_MN ? MN_STEP : 32;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int64_t w = pair; w < n_work; w += n_pairs) {
                const int split = (int)(w / tiles_mn);
                const int64_t kb = (int64_t)split * p.k_per_split;
                const int64_t ke = min(p.K, kb + p.k_per_split);
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                uint32_t first = 1;
                for (int64_t k0 = kb; k0 < ke; k0 += BK) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    __syncwarp();
                    if (elect_one()) {
                        const uint32_t sa = tiles + stage * STAGE2_BYTES, sb = sa + A2_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            const uint64_t ad = make_desc(sa + k * a_step, a_lbo, a_sbo, a_lay);
                            const uint64_t bd = make_desc(sb + k * b_step, b_lbo, b_sbo, b_lay);
                            if (BF16) umma_bf16_2sm(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                            else umma_tf32_2sm(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                            first = 0;
                        }
                        umma_commit_2sm(empty_bar + 8 * stage, 3);      // frees the stage in BOTH CTAs
                        if (k0 + BK >= ke) umma_commit_2sm(tfull_bar + 8 * acc, 3);
                    }
                    __syncwarp();
                    first = 0;
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5, both CTAs: own 128 rows) =====================
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        int ebuf = 0;
        for (int64_t w = pair; w < n_work; w += n_pairs) {
            const int split = (int)(w / tiles_mn);
            const int64_t rem = w - (int64_t)split * tiles_mn;
            const int64_t m0 = (rem / p.n_tiles) * 256 + HALF * (int64_t)rank;
            const int n0 = (int)(rem % p.n_tiles) * 256;
            float* Cs = p.C + (int64_t)split * p.split_stride;
            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256;
            if (sconst) epilogue_tile<1, true>(p, &tmC, epi, warp, lane, q, m0, n0, t_row, Cs, ebuf, sconst);
            else epilogue_tile<1>(p, &tmC, epi, warp, lane, q, m0, n0, t_row, Cs, ebuf);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty_bar + 8 * acc, 0));   // leader's barrier, 8 arrivals
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) bulk_wait_all();                            // outstanding TMA stores complete before the CTA exits
    }

    __syncwarp();                                                 // reconverge before the .aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();                                           // nobody leaves while the peer may still touch us
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// T[m, k] = sum over the n-tiles of their partial projections, in tile order (deterministic)
__global__ void __launch_bounds__(256)
proj_reduce_kernel(const float4* __restrict__ part, int n_tiles, int64_t M, float* __restrict__ out, int64_t ldp, int k) {
    pdl_prologue();
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float4 s = __ldg(part + m);
    for (int t = 1; t < n_tiles; ++t) {
        const float4 v = __ldg(part + (int64_t)t * M + m);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = out + m * ldp;
    o[0] = s.x;
    if (k > 1) o[1] = s.y;
    if (k > 2) o[2] = s.z;
    if (k > 3) o[3] = s.w;
}

__global__ void tc_splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t MN, int64_t N,
                                        float* __restrict__ C, int64_t ldc, int accumulate) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * MN + i];
    float* c = C + (i / N) * ldc + (i % N);
    *c = accumulate ? *c + s : s;
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D tensor [outer rows, inner cols] of fp32 (elt = 4) or bf16 (elt = 2) with row pitch ld (elements);
// box = [box_outer, box_inner]
static int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                    uint32_t box_inner, uint32_t box_outer, bool mn_major, int elt = 4) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("gmc_gemm: cuTensorMapEncodeTiled entry point not available"); return GMC_ERR_UNSUPPORTED; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * (uint64_t)elt};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    // 32-bit MN-major operands need the 32-byte-atom swizzle; everything else (K-major, 16-bit MN-major) plain 128B
    const CUtensorMapSwizzle sw = (mn_major && elt == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(map, elt == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gmc_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r); return GMC_ERR_INVALID_ARG; }
    return GMC_OK;
}

// cluster shape of the multicast kernel: GMC_GEMM_CLUSTER = "1" | "2" | "4" | "8" (CLM x 1) | "2x2" | "4x2" (CLM x CLN).
// Default: 2x2 for tn (both operands streamed from HBM) when the problem has at least two n-tiles, else 4x1 --
// the best shapes measured at config 3 (profiles/r01_gemm_notes.md: nn 7.5 ms with 4x1, tn 7.0 ms with 2x2).
static void cluster_shape(bool a_mn, int64_t n_tiles, int* clm, int* cln, bool bf16 = false) {
    static int cm = -1, cn = -1;
    if (cm < 0) {
        const char* e = getenv("GMC_GEMM_CLUSTER");
        cm = 0; cn = 0;                                            // 0 = automatic
        if (e && e[0] >= '1' && e[0] <= '8') {
            cm = e[0] - '0';
            cn = (e[1] == 'x' && e[2] == '2') ? 2 : 1;
            const bool ok = (cn == 1 && (cm == 1 || cm == 2 || cm == 4 || cm == 8)) || (cn == 2 && (cm == 2 || cm == 4));
            if (!ok) { cm = 0; cn = 0; }
        }
    }
    if (cm > 0) { *clm = cm; *cln = cn; return; }
    // bf16 operands halve the L2 traffic per flop, so the sharing of the larger clusters buys nothing and their
    // co-residency loss costs: 2x1 clusters keep all 148 SMs busy (4x1: 132, 2x2: 128).  Config 3, scratch/gemm_bf16_cfg3.py:
    // tn 3.30 ms (2x2) -> 2.99 ms (2x1); nn 3.36 (4x1) -> 3.32 (2x1), 3.33 without clusters.
    if (bf16) { *clm = 2; *cln = 1; return; }
    if (a_mn && n_tiles >= 2) { *clm = 2; *cln = 2; return; }
    *clm = 4;
    *cln = 1;
}

// max_chain > 0 (fp32-grade split GEMMs): no accumulator sums more than ~max_chain k-rows -- the fp32 TMEM accumulation
// of a 10^6-term chain costs accuracy the split operands were meant to buy -- so a long K is cut into a multiple of the
// slot-filling split count; the partials (2 MB each at config 3) are added by the deterministic reduce kernel.
constexpr int64_t kSplitChain = 32768;
static int pick_splits_cl(int64_t cluster_tiles, int64_t K, int slots, int64_t max_chain = 0) {
    int64_t s = 1;
    // a split must be worth its reduce launch: >= 16 k-blocks each (a 500-node graph's dW1 GEMM, K = 500, ran as three
    // splits of 170 k-rows plus a 6.5 us reduce kernel in a 70 us step)
    if (cluster_tiles < slots && K >= 32 * BLOCK_K) {
        s = slots / cluster_tiles;
        const int64_t max_by_k = K / (16 * BLOCK_K);
        if (s > max_by_k) s = max_by_k;
        if (s < 1) s = 1;
    }
    if (max_chain > 0 && K > s * max_chain) s *= ceil_div<int64_t>(K, s * max_chain);
    return (int)s;
}

static bool no_tma_store() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("GMC_GEMM_NO_TMA_STORE"); cached = (e && e[0] == '1') ? 1 : 0; }
    return cached == 1;
}

// co-resident clusters of the kernel on this device: the GPC geometry may admit fewer than SMs / CL
template <bool A_MN, bool B_MN, int CLM, int CLN, bool BF16, int NS = 1>
static int cluster_slots() {
    constexpr int CL = CLM * CLN;
    static int cached = -1;
    if (cached < 0) {
        cudaFuncSetAttribute(gemm_umma_kernel<A_MN, B_MN, CLM, CLN, BF16, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        cached = sm_count() / CL;
        if (CL > 1) {
            cudaLaunchConfig_t probe = {};
            probe.blockDim = dim3(THREADS);
            probe.dynamicSmemBytes = SMEM_BYTES;
            probe.gridDim = dim3((sm_count() / CL) * CL);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            probe.attrs = attr;
            probe.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, gemm_umma_kernel<A_MN, B_MN, CLM, CLN, BF16, NS>, &probe) == cudaSuccess && n > 0) {
                if (n < cached) cached = n;
            } else {
                cudaGetLastError();
            }
        }
        if (getenv("GMC_GEMM_DEBUG")) fprintf(stderr, "gmc gemm: cluster %dx%d -> %d co-resident clusters\n", CLM, CLN, cached);
    }
    return cached;
}

// Output side of a launch, shared by the cluster kernel and the cta_group::2 kernel: where the accumulators go (C or the
// split-K workspace), the epilogue options and the TMA-store map.  p.n_tiles must be set (fused fp32 projection).
static int setup_output(Params& p, CUtensorMap* tmC_out, float* C, int64_t M, int64_t N, int64_t ldc, int accumulate, int splits,
                        void* workspace, size_t workspace_bytes, cudaStream_t s, int c_bf16, const float* bias, int relu,
                        const float* proj_w, float* proj_out, int64_t ldp, int proj_k, const float* row_scale) {
    if (splits == 1) {
        p.C = C; p.ldc = ldc; p.split_stride = 0; p.accumulate = accumulate;
        p.vec_ok = (ldc % (c_bf16 ? 8 : 4) == 0) && aligned16(C);
    } else {
        p.C = reinterpret_cast<float*>(workspace); p.ldc = N; p.split_stride = M * N; p.accumulate = 0;
        p.vec_ok = (N % 4 == 0) && aligned16(workspace);
    }
    // direct (non split-K, non accumulating) outputs leave through a 128B-swizzled staging block and TMA stores
    CUtensorMap& tmC = *tmC_out;
    memset(&tmC, 0, sizeof(tmC));
    int rc;
    p.tma_store = 0;
    if (c_bf16) {
        if (accumulate || !p.vec_ok) {
            set_error("gmc_gemm_bf16_bf16out: needs accumulate = 0, a 16-byte aligned C and ldc %% 8 == 0");
            return GMC_ERR_INVALID_ARG;
        }
        // the stored width runs to the end of the row's last 128-byte line when the pitch covers it: the columns past N
        // hold exact zeros (their accumulators see zero-filled B columns, no bias), and a row that ends inside a line
        // turns its last sector into a read-modify-write that costs a write stream 40 % of its bandwidth
        rc = make_map(&tmC, C, (uint64_t)(((N + 63) & ~(int64_t)63) <= ldc ? ((N + 63) & ~(int64_t)63) : N), (uint64_t)M,
                      (uint64_t)ldc, 64, 32, false, 2);
        if (rc) return rc;
        p.tma_store = 1;
        p.c_bf16 = 1;
        if (bias && (N % 4 != 0 || !aligned16(bias))) {
            set_error("gmc_gemm_bf16_bf16out: a bias needs N %% 4 == 0 and a 16-byte aligned pointer");
            return GMC_ERR_INVALID_ARG;
        }
        p.bias = bias;
        p.relu = relu;
        if (proj_w) {
            if (!proj_out || proj_k < 1 || proj_k > 4 || ldp < proj_k || N > 2 * BLOCK_N || !aligned16(proj_w)) {
                set_error("gmc_gemm_bf16_bf16out: the fused projection needs 1 <= n_proj <= 4, ldp >= n_proj, N <= 512 and a "
                          "16-byte aligned padded weight matrix");
                return GMC_ERR_INVALID_ARG;
            }
            GMC_CUDA(cudaMemset2DAsync(proj_out, (size_t)ldp * sizeof(float), 0, (size_t)proj_k * sizeof(float), (size_t)M, s));
            p.proj_w = reinterpret_cast<const float4*>(proj_w);
            p.proj_out = proj_out;
            p.ldp = ldp;
            p.proj_k = proj_k;
        }
    } else {
        if (row_scale || bias || relu || proj_w) {
            if (accumulate || (bias && (N % 4 != 0 || !aligned16(bias)))) {
                set_error("gmc_gemm: an fp32 epilogue (row scale / bias / ReLU) needs accumulate = 0 and, with a bias, "
                          "N %% 4 == 0 and a 16-byte aligned bias");
                return GMC_ERR_INVALID_ARG;
            }
            p.row_scale = row_scale;
            p.bias = bias;
            p.relu = relu;
        }
        if (proj_w) {
            const size_t need = (size_t)p.n_tiles * (size_t)M * sizeof(float4);
            // proj_out may be NULL: the partials then stay in the workspace for gmc_layer2_loss_fused_parts (no reduce launch)
            if (splits != 1 || accumulate || proj_k < 1 || proj_k > 4 || (proj_out && ldp < proj_k) || !aligned16(proj_w) ||
                !workspace || workspace_bytes < need || !aligned16(workspace)) {
                set_error("gmc_gemm: the fused fp32 projection needs a direct (non split-K, non accumulating) output, "
                          "1 <= n_proj <= 4, ldp >= n_proj, a 16-byte aligned padded weight matrix and a workspace of "
                          "n_tiles * M * 16 bytes (%zu)", need);
                return GMC_ERR_INVALID_ARG;
            }
            p.proj_w = reinterpret_cast<const float4*>(proj_w);
            p.proj_part = reinterpret_cast<float4*>(workspace);
        }
        if (splits == 1 && !accumulate && p.vec_ok && !no_tma_store()) {
            rc = make_map(&tmC, C, (uint64_t)(((N + 31) & ~(int64_t)31) <= ldc ? ((N + 31) & ~(int64_t)31) : N), (uint64_t)M,
                          (uint64_t)ldc, 32, 32, false);                         // full 128-byte lines, as above
            if (rc) return rc;
            p.tma_store = 1;
        }
    }
    return GMC_OK;
}

template <bool A_MN, bool B_MN, int CLM, int CLN, bool BF16, int NS = 1>
static int launch(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                  int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t s, int c_bf16 = 0,
                  const float* bias = nullptr, int relu = 0, const float* proj_w = nullptr, float* proj_out = nullptr,
                  int64_t ldp = 0, int proj_k = 0, int64_t b_split_rows = 0, const float* row_scale = nullptr, int lo_shift = 0) {
    constexpr int CL = CLM * CLN;
    constexpr int ELT = BF16 ? 2 : 4;
    constexpr int BK = 128 / ELT, MNC = 128 / ELT;                 // k elements per stage, m/n elements per MN-major chunk
    constexpr int RN = (NS == 3 ? 192 : BLOCK_N) / NS;             // real output columns per tile
    CUtensorMap tmA, tmB;
    int rc;
    if (A_MN) rc = make_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, MNC, BK, true, ELT);       // A[K rows, M cols]
    else      rc = make_map(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BLOCK_M / CLN, false, ELT);  // A[M rows, K cols]
    if (rc) return rc;
    // split operand: NS stacked [b_split_rows, N] parts; the last part ends at row (NS - 1) * b_split_rows + K
    if (B_MN) rc = make_map(&tmB, B, (uint64_t)N, (uint64_t)((NS - 1) * b_split_rows + K), (uint64_t)ldb, MNC, BK, true, ELT);  // B[K rows, N cols]
    else      rc = make_map(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BLOCK_N / CLM, false, ELT);  // B[N rows, K cols]
    if (rc) return rc;

    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    const int slots = cluster_slots<A_MN, B_MN, CLM, CLN, BF16, NS>();  // co-resident clusters (GPC geometry), <= SMs / CL

    Params p = {};
    p.M = M; p.N = N; p.K = K;
    p.b_split_rows = b_split_rows;
    p.b_f16 = lo_shift > 0 ? 1 : 0;
    p.part_scale = lo_shift > 0 ? 1.0f / (float)(1u << lo_shift) : 1.0f;
    p.m_tiles = (int)ceil_div<int64_t>(M, BLOCK_M);
    p.n_tiles = (int)ceil_div<int64_t>(N, RN);
    p.mg_tiles = ceil_div(p.m_tiles, CLM);
    p.ng_tiles = ceil_div(p.n_tiles, CLN);
    const int64_t ctiles = (int64_t)p.mg_tiles * p.ng_tiles;
    const bool fused_epilogue = c_bf16 || row_scale || bias || relu || proj_w;   // these leave straight from the accumulators
    int splits = fused_epilogue ? 1 : pick_splits_cl(ctiles, K, slots, NS > 1 ? kSplitChain : 0);
    if (splits > 1 && (!workspace || workspace_bytes < (size_t)splits * M * N * sizeof(float))) {
        splits = workspace ? (int)(workspace_bytes / ((size_t)M * N * sizeof(float))) : 1;
        if (splits < 1) splits = 1;
    }
    int64_t k_per = ceil_div<int64_t>(ceil_div<int64_t>(K, splits), BK) * BK;
    splits = (int)ceil_div<int64_t>(K, k_per);
    p.k_splits = splits;
    p.k_per_split = k_per;
    CUtensorMap tmC;
    rc = setup_output(p, &tmC, C, M, N, ldc, accumulate, splits, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w,
                      proj_out, ldp, proj_k, row_scale);
    if (rc) return rc;
    const int64_t n_work = ctiles * splits;
    const int grid = (int)(n_work < slots ? n_work : slots) * CL;
    cfg.gridDim = dim3(grid);
    GMC_CUDA(cudaLaunchKernelEx(&cfg, gemm_umma_kernel<A_MN, B_MN, CLM, CLN, BF16, NS>, tmA, tmB, tmC, p));
    if (splits > 1) {
        const int64_t MN = M * N;
        GMC_CUDA(launch_pdl(tc_splitk_reduce_kernel, (unsigned)ceil_div<int64_t>(MN, 256), 256, 0, s, p.C, splits, MN, N, C, ldc,
                            accumulate));
    }
    if (p.proj_part && proj_out) {
        GMC_CUDA(launch_pdl(proj_reduce_kernel, (unsigned)ceil_div<int64_t>(M, 256), 256, 0, s, p.proj_part, p.n_tiles, M,
                            proj_out, ldp, proj_k));
    }
    return GMC_OK;
}

template <bool A_MN, bool B_MN, bool BF16 = false>
static int launch_cl(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                     int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t s, int c_bf16 = 0,
                     const float* bias = nullptr, int relu = 0, const float* proj_w = nullptr, float* proj_out = nullptr,
                     int64_t ldp = 0, int proj_k = 0) {
    int clm, cln;
    cluster_shape(A_MN, ceil_div<int64_t>(N, BLOCK_N), &clm, &cln, BF16);
    if (BF16 && B_MN && clm == 8) clm = 4;                         // a bf16 MN-major B stage has four chunks to share out
#define GMC_GEMM_CASE(CM, CN)                                                                                       \
    if (clm == CM && cln == CN)                                                                                     \
        return launch<A_MN, B_MN, CM, CN, BF16>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
    GMC_GEMM_CASE(1, 1) GMC_GEMM_CASE(2, 1) GMC_GEMM_CASE(4, 1) GMC_GEMM_CASE(8, 1) GMC_GEMM_CASE(2, 2) GMC_GEMM_CASE(4, 2)
#undef GMC_GEMM_CASE
    return launch<A_MN, B_MN, 4, 1, BF16>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
}

static bool use_two_cta() {
    static int cached = -1;
    // Measured on B200 (profiles/r01_gemm_notes.md): the pair kernel moves 1.5x fewer bytes from L2 but its tensor
    // pipe is only 26 % active (1-CTA: 48 %), so it is opt-in (GMC_GEMM_2CTA=1) until the stall is understood.
    if (cached < 0) {
        const char* e = getenv("GMC_GEMM_2CTA");
        cached = (e && e[0] == '1') ? 1 : 0;
    }
    return cached == 1;
}

// bf16 operands through the cta_group::2 kernel: GMC_GEMM_BF16_2CTA = "1" (all ops) | "nn" | "tn" | "0".  Default "nn":
// interleaved A/B at config 3 under the power cap (profiles/r02_gemm_2cta.json): layer-1 GEMM with its full epilogue
// 3.6-5.3 ms (cluster kernel) vs 3.6-4.3 ms (pairs), plain bf16-out 3.4-4.6 vs 3.3-3.4 -- each CTA stages 32 KB per
// k-block instead of 48 KB, fewer shared-memory and L2 transactions per flop, so the pair kernel also holds its clocks
// under the 1 kW cap; tn (both operands MN-major, split-K) is slower in pairs (4.3-4.5 vs 3.2-3.3 ms) and keeps the
// 2x1 multicast clusters.
static int two_cta_bf16_mode() {
    const char* e = getenv("GMC_GEMM_BF16_2CTA");                  // read per call: benchmarks toggle it in-process
    if (!e) return 1;
    if (e[0] == '1') return 7;
    if (e[0] == 'n' && e[1] == 'n') return 1;
    if (e[0] == 't' && e[1] == 'n') return 4;
    return 0;
}
static bool use_two_cta_bf16(int op) { return (two_cta_bf16_mode() >> op) & 1; }

static int pick_splits2(int64_t tiles, int64_t K) {
    const int pairs = sm_count() / 2;
    if (tiles >= pairs || K < 32 * BLOCK_K) return 1;
    int64_t s = pairs / tiles;
    const int64_t max_by_k = K / (16 * BLOCK_K);
    if (s > max_by_k) s = max_by_k;
    return (int)(s < 1 ? 1 : s);
}

template <bool A_MN, bool B_MN, bool BF16>
static int launch2(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                   int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t s, int c_bf16 = 0,
                   const float* bias = nullptr, int relu = 0, const float* proj_w = nullptr, float* proj_out = nullptr,
                   int64_t ldp = 0, int proj_k = 0) {
    constexpr int ELT = BF16 ? 2 : 4;
    constexpr int BK = 128 / ELT, MNC = 128 / ELT;
    static bool attr_set = false;
    if (!attr_set) {
        GMC_CUDA(cudaFuncSetAttribute(gemm_umma2_kernel<A_MN, B_MN, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM2_BYTES));
        attr_set = true;
    }
    CUtensorMap tmA, tmB;
    int rc;
    if (A_MN) rc = make_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, MNC, BK, true, ELT);
    else      rc = make_map(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, HALF, false, ELT);
    if (rc) return rc;
    if (B_MN) rc = make_map(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, MNC, BK, true, ELT);
    else      rc = make_map(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, HALF, false, ELT);
    if (rc) return rc;

    Params p = {};
    p.M = M; p.N = N; p.K = K;
    p.m_tiles = (int)ceil_div<int64_t>(M, 256);
    p.n_tiles = (int)ceil_div<int64_t>(N, 256);
    const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
    const bool fused_epilogue = c_bf16 || bias || relu || proj_w;
    int splits = fused_epilogue ? 1 : pick_splits2(tiles, K);
    if (splits > 1 && (!workspace || workspace_bytes < (size_t)splits * M * N * sizeof(float))) {
        splits = workspace ? (int)(workspace_bytes / ((size_t)M * N * sizeof(float))) : 1;
        if (splits < 1) splits = 1;
    }
    int64_t k_per = ceil_div<int64_t>(ceil_div<int64_t>(K, splits), BK) * BK;
    splits = (int)ceil_div<int64_t>(K, k_per);
    p.k_splits = splits;
    p.k_per_split = k_per;
    CUtensorMap tmC;
    rc = setup_output(p, &tmC, C, M, N, ldc, accumulate, splits, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w,
                      proj_out, ldp, proj_k, nullptr);
    if (rc) return rc;
    const int64_t n_work = tiles * splits;
    const int pairs = sm_count() / 2;
    const int grid = 2 * (int)(n_work < pairs ? n_work : pairs);
    {
        cudaLaunchConfig_t cfg2 = {};
        cfg2.gridDim = dim3(grid);
        cfg2.blockDim = dim3(THREADS);
        cfg2.dynamicSmemBytes = SMEM2_BYTES;
        cfg2.stream = s;
        cudaLaunchAttribute attr2[1];
        attr2[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr2[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        cfg2.attrs = attr2;
        cfg2.numAttrs = 1;
        GMC_CUDA(cudaLaunchKernelEx(&cfg2, gemm_umma2_kernel<A_MN, B_MN, BF16>, tmA, tmB, tmC, p));
    }
    if (splits > 1) {
        const int64_t MN = M * N;
        GMC_CUDA(launch_pdl(tc_splitk_reduce_kernel, (unsigned)ceil_div<int64_t>(MN, 256), 256, 0, s, p.C, splits, MN, N, C, ldc,
                            accumulate));
    }
    return GMC_OK;
}

// ---- 3xTF32: low-order operand parts ------------------------------------------------------------
// The MMA reads the upper 19 bits of each fp32 operand (truncation), so feeding x itself IS feeding hi(x).
// lo(x) = x - hi(x) is exact in fp32; its own truncation to TF32 costs 2^-21 |x| at most.
__global__ void __launch_bounds__(256)
tf32_lo_kernel(const float* __restrict__ X, int64_t ldx, float* __restrict__ L, int64_t ldl, int64_t rows, int cols) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // one float4 of the (padded) output row
    const int c4 = (int)(ldl >> 2);
    const int64_t r = i / c4;
    const int c = (int)(i - r * c4) * 4;
    if (r >= rows) return;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols) {                                                      // cols % 4 == 0 is required by the TMA path
        const float4 x = *reinterpret_cast<const float4*>(X + r * ldx + c);
        out.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
        out.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
        out.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
        out.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
    }
    *reinterpret_cast<float4*>(L + r * ldl + c) = out;
}

static int64_t lo_ld(int64_t cols) { return (cols + 31) / 32 * 32; }     // 128-byte row pitch for the TMA boxes

static int make_lo(const float* X, int64_t ldx, float* L, int64_t rows, int64_t cols, cudaStream_t s) {
    const int64_t ldl = lo_ld(cols);
    const int64_t n4 = rows * (ldl / 4);
    if (n4 == 0) return GMC_OK;
    tf32_lo_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, s>>>(X, ldx, L, ldl, rows, (int)cols);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

}  // namespace tc

static size_t tc_splitk_bytes(int op, int64_t M, int64_t N, int64_t K, bool bf16 = false) {
    int splits;
    if ((tc::use_two_cta() && !bf16) || (bf16 && tc::use_two_cta_bf16(op))) {
        const int64_t tiles = ceil_div<int64_t>(M, 256) * ceil_div<int64_t>(N, 256);
        splits = tc::pick_splits2(tiles, K);
    } else {
        int clm, cln;
        tc::cluster_shape(op == 2, ceil_div<int64_t>(N, tc::BLOCK_N), &clm, &cln, bf16);
        const int64_t ctiles = ceil_div<int64_t>(ceil_div<int64_t>(M, tc::BLOCK_M), clm) *
                               ceil_div<int64_t>(ceil_div<int64_t>(N, tc::BLOCK_N), cln);
        splits = tc::pick_splits_cl(ctiles, K, sm_count() / (clm * cln));   // upper bound of what launch() picks
    }
    return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}

// operand shapes as stored: nn A[M,K] B[K,N]; nt A[M,K] B[N,K]; tn A[K,M] B[K,N]
static void operand_shapes(int op, int64_t M, int64_t N, int64_t K, int64_t* ra, int64_t* ca, int64_t* rb, int64_t* cb) {
    *ra = op == 2 ? K : M; *ca = op == 2 ? M : K;
    *rb = op == 1 ? N : K; *cb = op == 1 ? K : N;
}

size_t tc_workspace_bytes(int op, int64_t M, int64_t N, int64_t K, int precision) {
    size_t need = (tc_splitk_bytes(op, M, N, K) + 255) & ~(size_t)255;
    if (precision == GMC_GEMM_TF32X3) {
        int64_t ra, ca, rb, cb;
        operand_shapes(op, M, N, K, &ra, &ca, &rb, &cb);
        need += ((size_t)ra * tc::lo_ld(ca) * sizeof(float) + 255) & ~(size_t)255;
        need += ((size_t)rb * tc::lo_ld(cb) * sizeof(float) + 255) & ~(size_t)255;
    }
    return need;
}


static int tc_gemm_pass(int op, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                        int64_t ldb, int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes,
                        cudaStream_t s) {
    GMC_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && aligned16(A) && aligned16(B),
                "gmc_gemm(tf32): TMA needs 16-byte aligned bases and leading dimensions that are multiples of 4");
    GMC_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gmc_gemm(tf32): dimension exceeds int32 TMA coordinates");
    if (M == 0 || N == 0) return GMC_OK;
    if (K == 0) {
        if (!accumulate) GMC_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, s));
        return GMC_OK;
    }
    if (tc::use_two_cta()) {
        switch (op) {
            case 0: return tc::launch2<false, true, false>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
            case 1: return tc::launch2<false, false, false>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
            case 2: return tc::launch2<true, true, false>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
        }
    }
    switch (op) {
        case 0: return tc::launch_cl<false, true>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
        case 1: return tc::launch_cl<false, false>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
        case 2: return tc::launch_cl<true, true>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
    }
    set_error("gmc_gemm: bad op %d", op);
    return GMC_ERR_INVALID_ARG;
}

int tc_gemm(int op, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
            int64_t ldb, int64_t ldc, int accumulate, int precision, void* workspace, size_t workspace_bytes,
            cudaStream_t s) {
    if (precision == GMC_GEMM_TF32)
        return tc_gemm_pass(op, A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
    if (precision != GMC_GEMM_TF32X3) {
        set_error("gmc_gemm: unknown tensor-core precision %d", precision);
        return GMC_ERR_INVALID_ARG;
    }
    // 3xTF32 (fp32-grade): C = hi(A) hi(B) + lo(A) hi(B) + hi(A) lo(B); the lo*lo term (2^-22 relative) is dropped.
    // Three passes of the TF32 kernel, the low parts materialised in the caller's workspace.
    if (M == 0 || N == 0) return GMC_OK;
    if (K == 0) return tc_gemm_pass(op, A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
    GMC_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && aligned16(A) && aligned16(B),
                "gmc_gemm(tf32x3): TMA needs 16-byte aligned bases and leading dimensions that are multiples of 4");
    int64_t ra, ca, rb, cb;
    operand_shapes(op, M, N, K, &ra, &ca, &rb, &cb);
    GMC_REQUIRE(ca % 4 == 0 && cb % 4 == 0, "gmc_gemm(tf32x3): operand row lengths must be multiples of 4 floats");
    const size_t split_bytes = (tc_splitk_bytes(op, M, N, K) + 255) & ~(size_t)255;
    const size_t a_bytes = ((size_t)ra * tc::lo_ld(ca) * sizeof(float) + 255) & ~(size_t)255;
    const size_t b_bytes = ((size_t)rb * tc::lo_ld(cb) * sizeof(float) + 255) & ~(size_t)255;
    GMC_REQUIRE(workspace && workspace_bytes >= split_bytes + a_bytes + b_bytes,
                "gmc_gemm(tf32x3): workspace too small (gmc_gemm_workspace_bytes reports the need)");
    GMC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "gmc_gemm(tf32x3): workspace must be 256-byte aligned");
    char* wsb = reinterpret_cast<char*>(workspace);
    float* A_lo = reinterpret_cast<float*>(wsb + split_bytes);
    float* B_lo = reinterpret_cast<float*>(wsb + split_bytes + a_bytes);
    int rc = tc::make_lo(A, lda, A_lo, ra, ca, s);
    if (rc) return rc;
    rc = tc::make_lo(B, ldb, B_lo, rb, cb, s);
    if (rc) return rc;
    rc = tc_gemm_pass(op, A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, split_bytes, s);
    if (rc) return rc;
    rc = tc_gemm_pass(op, A_lo, B, C, M, N, K, tc::lo_ld(ca), ldb, ldc, 1, workspace, split_bytes, s);
    if (rc) return rc;
    return tc_gemm_pass(op, A, B_lo, C, M, N, K, lda, tc::lo_ld(cb), ldc, 1, workspace, split_bytes, s);
}

// bf16 operands (A, B point at __nv_bfloat16, leading dimensions in elements), fp32 accumulation and fp32 output:
// the same kernel with 64-element stages and tcgen05.mma kind::f16 -- half the operand bytes per flop through HBM,
// L2 and shared memory, twice the MMA rate.
size_t tc_bf16_workspace_bytes(int op, int64_t M, int64_t N, int64_t K) { return (tc_splitk_bytes(op, M, N, K, true) + 255) & ~(size_t)255; }

int tc_gemm_bf16(int op, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                 int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t s, int c_bf16,
                 const float* bias, int relu, const float* proj_w, float* proj_out, int64_t ldp, int proj_k) {
    GMC_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && aligned16(A) && aligned16(B),
                "gmc_gemm_bf16: TMA needs 16-byte aligned bases and leading dimensions that are multiples of 8 elements");
    GMC_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gmc_gemm_bf16: dimension exceeds int32 TMA coordinates");
    if (M == 0 || N == 0) return GMC_OK;
    const size_t celt = c_bf16 ? 2 : 4;
    if (K == 0) {
        if (!accumulate) GMC_CUDA(cudaMemset2DAsync(C, (size_t)ldc * celt, 0, (size_t)N * celt, (size_t)M, s));
        return GMC_OK;
    }
    float* Cf = reinterpret_cast<float*>(C);                       // reinterpreted by the kernel when c_bf16 is set
    if (tc::use_two_cta_bf16(op)) {
        switch (op) {
            case 0: return tc::launch2<false, true, true>(A, B, Cf, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
            case 1: return tc::launch2<false, false, true>(A, B, Cf, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
            case 2: return tc::launch2<true, true, true>(A, B, Cf, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
        }
    }
    switch (op) {
        case 0: return tc::launch_cl<false, true, true>(A, B, Cf, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
        case 1: return tc::launch_cl<false, false, true>(A, B, Cf, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
        case 2: return tc::launch_cl<true, true, true>(A, B, Cf, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, c_bf16, bias, relu, proj_w, proj_out, ldp, proj_k);
    }
    set_error("gmc_gemm_bf16: bad op %d", op);
    return GMC_ERR_INVALID_ARG;
}

// ---- split-operand bf16 GEMM: fp32-grade products at bf16 tensor-core rates ---------------------------------------
// C = op(A) * (B_0 + ... + B_{NS-1}): A exact in bf16 (the integer path counts d * A_hat X of a regular graph), B an fp32
// matrix split into NS bf16 parts by gmc_f32_split_bf16 (NS = 2: 16 mantissa bits, NS = 3: 24 = all of fp32), products
// accumulated in fp32 TMEM and added up in the epilogue.  op 0 (nn) and 2 (tn): B is MN-major in both.
namespace tc {
static void split_cluster(int ns, int* clm, bool a_mn = false) {
    // The B stage of a split GEMM is NS x as wide per real output column as a plain one, so sharing it pays (unlike the
    // plain bf16 nn kernel): every CTA of a CLM x 1 cluster loads 1 / CLM of the stage's 64-column chunks and multicasts
    // them.  NS = 2: four chunks -> CLM 1 | 2 | 4; NS = 3: three chunks -> CLM 1 | 3.  GMC_GEMM_SPLIT_CLUSTER overrides.
    // Measured at config 3 (profiles/r02_split_gemm.json): NS = 2 nn 8.6 / 6.3-7.0 / 7.5 ms for CLM 1 / 2 / 4, tn 10.6 /
    // 7.0-7.5 / 7.7; NS = 3 nn 12.9 / 11.6-11.8 for CLM 1 / 3, tn 17.8 / 21.2 (48 co-resident clusters of 3 against 8
    // m-tiles: idle CTAs) -> 2, 3 for nn, 2, 1 for tn.
    *clm = ns == 2 ? 2 : (a_mn ? 1 : 3);
    const char* e = getenv("GMC_GEMM_SPLIT_CLUSTER");
    if (e && e[0] >= '1' && e[0] <= '4') {
        const int c = e[0] - '0';
        if (ns == 2 && (c == 1 || c == 2 || c == 4)) *clm = c;
        if (ns == 3 && (c == 1 || c == 3)) *clm = c;
    }
}

static int64_t split_tiles(int ns, int64_t M, int64_t N, int clm) {
    const int rn = (ns == 3 ? 192 : BLOCK_N) / ns;
    return ceil_div<int64_t>(ceil_div<int64_t>(M, BLOCK_M), clm) * ceil_div<int64_t>(N, rn);
}
}  // namespace tc

size_t tc_bf16_split_workspace_bytes(int op, int64_t M, int64_t N, int64_t K, int n_split, int with_projection) {
    (void)op;
    if (n_split < 2 || n_split > 3) return 0;
    if (with_projection) {                                           // partial projections [n_tiles][M] float4, no split-K
        const int rn = (n_split == 3 ? 192 : tc::BLOCK_N) / n_split;
        return (size_t)ceil_div<int64_t>(N, rn) * (size_t)M * sizeof(float4);
    }
    int clm;
    tc::split_cluster(n_split, &clm, op == 2);
    const int splits = tc::pick_splits_cl(tc::split_tiles(n_split, M, N, clm), K, sm_count() / clm, tc::kSplitChain);
    return splits > 1 ? (((size_t)splits * M * N * sizeof(float) + 255) & ~(size_t)255) : 0;
}

int tc_gemm_bf16_split(int op, const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc, int n_split, int64_t b_split_rows, const float* row_scale,
                       const float* bias, int relu, const float* proj_w, float* proj_out, int64_t ldp, int proj_k,
                       int accumulate, int lo_shift, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    GMC_REQUIRE(lo_shift >= 0 && lo_shift <= 24, "gmc_gemm_bf16_split: lo_shift must be 0 (bf16 parts) or 1..24 (fp16 parts)");
    GMC_REQUIRE(op == 0 || op == 2, "gmc_gemm_bf16_split: op must be 0 (nn) or 2 (tn): the split operand is MN-major");
    GMC_REQUIRE(n_split == 2 || n_split == 3, "gmc_gemm_bf16_split: n_split must be 2 or 3");
    GMC_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && aligned16(A) && aligned16(B),
                "gmc_gemm_bf16_split: TMA needs 16-byte aligned bases and leading dimensions that are multiples of 8 elements");
    GMC_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && (int64_t)(n_split - 1) * b_split_rows + K < (1ll << 31),
                "gmc_gemm_bf16_split: dimension exceeds int32 TMA coordinates");
    GMC_REQUIRE(b_split_rows >= (K + 63) / 64 * 64, "gmc_gemm_bf16_split: b_split_rows must cover K rounded up to 64 "
                "(the rows between K and b_split_rows of every part must be zero)");
    if (M == 0 || N == 0) return GMC_OK;
    if (K == 0) {
        GMC_REQUIRE(!row_scale && !bias && !relu, "gmc_gemm_bf16_split: K = 0 with an epilogue");
        if (!accumulate) GMC_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, s));
        return GMC_OK;
    }
    int clm;
    tc::split_cluster(n_split, &clm, op == 2);
#define GMC_SPLIT_CASE(AMN, CM, NSV)                                                                                  \
    return tc::launch<AMN, true, CM, 1, true, NSV>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s, 0, \
                                                   bias, relu, proj_w, proj_out, ldp, proj_k, b_split_rows, row_scale, lo_shift);
#define GMC_SPLIT_OP(AMN)                                                                                             \
    if (n_split == 2) {                                                                                               \
        if (clm == 4) { GMC_SPLIT_CASE(AMN, 4, 2) } else if (clm == 2) { GMC_SPLIT_CASE(AMN, 2, 2) } else { GMC_SPLIT_CASE(AMN, 1, 2) } \
    }                                                                                                                 \
    if (clm == 3) { GMC_SPLIT_CASE(AMN, 3, 3) } else { GMC_SPLIT_CASE(AMN, 1, 3) }
    if (op == 0) { GMC_SPLIT_OP(false) }
    GMC_SPLIT_OP(true)
#undef GMC_SPLIT_OP
#undef GMC_SPLIT_CASE
}

}  // namespace gmc
