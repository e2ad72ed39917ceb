// spmm_slab.cu -- shared-memory-staged SpMM for block-diagonal batches of small graphs.
//
// The warp-per-row kernel (spmm.cu) reads every source row d times from L2: at n=1000, C=500,
// d=7 it moves 13 TB/s through L2 for 3.3 TB/s of HBM traffic and sits on the L2 roof (ncu,
// profiles/r01_v1_*).  Here one persistent CTA owns a (graph, column-slab) work item:
//
//   * the graph's source rows, restricted to a slab of W4 float4 columns, are staged ONCE into
//     shared memory: full 128-row boxes by TMA (cp.async.bulk.tensor.2d issued by one thread, mbarrier
//     completion -- no LSU wavefronts, no per-thread issue cost), the tail rows by cp.async.  The
//     default keeps TWO CTAs per SM (W4 = 7, 112 KB each) so one CTA's staging overlaps the other's
//     gathers; GMC_SLAB_VARIANT selects the other schedules that were measured (one CTA per SM with
//     224-byte slabs, or one CTA with two 112-byte buffers);
//   * LANES lanes own one output row and gather its neighbours' slab rows from shared memory with
//     LDS.128 -- every quarter-warp phase reads one contiguous row segment, so the gathers are
//     bank-conflict free;
//   * neighbour lists come from an ELL "plan" built once per static batch (gmc_spmm_plan_build):
//     8 uint16 LOCAL node ids (one 16-byte load) + ONE coefficient per row, padded slots point at an
//     all-zero shared-memory row; the next rows' slots are prefetched while the current row is
//     accumulated (also across work items);
//   * the output slab is written with 128-bit stores (row coefficient, bias and ReLU fused).
//
// L2->SM traffic drops from (d+1) x to ~1 x the matrix; HBM traffic stays the algorithmic
// 8*N*C + 4*nnz + 4*(N+1) bytes.  Used when every graph of the batch fits the slab buffer, has max
// degree <= 8 and uniform neighbour norms (regular graphs); otherwise gmc_spmm_batched_f32 runs the
// warp-per-row kernel.  The two kernels differ only in rounding order (sum-then-scale vs fused coef).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gmc {

constexpr int kEll = 8;                                   // padded neighbour slots per row
constexpr size_t kSlabSmemMax = 227 * 1024;
constexpr size_t kPlanHeader = 16;                        // int32 max_degree + padding

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- plan: ELL packing of a block-diagonal CSR -----------------------------------------------
// plan layout: [16-byte header: int32 max degree] [uint16 col[n_rows][8]] [float coef[n_rows]]
// col = LOCAL node ids, padded with n_g (the kernel's all-zero row).  A_hat[v,u] = nd[v] * ns[u] (GraphConv's
// norm='both'); the plan keeps ONE coefficient per row, coef[v] = nd[v] * ns[u], which requires every neighbour of
// v to carry the same ns -- true for the regular graphs the reference generates (GraphCreator 'reg') and for
// any graph whose neighbours of a node share a degree.  Other batches flag `overflow` and use the row kernel.
__global__ void ell_pack_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                const float* __restrict__ norm_src, const float* __restrict__ norm_dst,
                                const int32_t* __restrict__ graph_ptr, int n_graphs, int64_t n_rows,
                                int32_t* __restrict__ header, uint4* __restrict__ ell_col, float* __restrict__ coef,
                                int* __restrict__ overflow) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    const int g = find_graph(graph_ptr, n_graphs, v);
    const int base = __ldg(graph_ptr + g);
    const int n_g = __ldg(graph_ptr + g + 1) - base;
    const int e0 = __ldg(rowptr + v), deg = __ldg(rowptr + v + 1) - e0;
    if (deg > kEll || n_g > 65534) { *overflow = 1; return; }
    atomicMax(header, deg);
    uint32_t c[kEll];
    float ns0 = 0.f;
    bool uniform = true;
#pragma unroll
    for (int j = 0; j < kEll; ++j) {
        c[j] = (uint32_t)n_g;
        if (j < deg) {
            const int u = __ldg(colidx + e0 + j);
            const float ns = __ldg(norm_src + u);
            if (j == 0) ns0 = ns;
            uniform = uniform && (ns == ns0);
            c[j] = (uint32_t)(u - base);
        }
    }
    if (!uniform) { *overflow = 1; return; }
    ell_col[v] = make_uint4(c[0] | (c[1] << 16), c[2] | (c[3] << 16), c[4] | (c[5] << 16), c[6] | (c[7] << 16));
    coef[v] = __ldg(norm_dst + v) * ns0;
}

// ---- TMA / mbarrier wrappers (staging path) ----------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) { printf("gmc spmm slab: mbarrier timeout (block %d)\n", (int)blockIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int kBoxRows = 128;                             // rows per TMA box

// W4: float4 columns per slab; LANES: lanes per output row (power of two >= W4); MINB: CTAs per SM;
// DEPTH: unused since the batched list prefetch (kept in the instantiation list); TMA: stage full 128-row boxes with
// cp.async.bulk.tensor (one elected thread, mbarrier completion) and only the tail rows with cp.async
// PROJ: also emit this slab's share of the skinny projection T = Y W (W [n_cols, n_out <= 4]) -- the second GraphConv
// layer's th.matmul (TrainingNeural.py:83) -- as one float4 per (slab, row) into Tpart[n_slabs][n_rows]; a fixed-order
// reduction over the slabs (proj_reduce_kernel) finishes it, so the result stays deterministic.
struct SlabProj {
    const float* W;          // [n_cols, n_out] row-major
    float4* Tpart;           // [n_slabs][n_rows]
    int64_t n_rows;
    int n_out;
};

// XB: X and Y are bf16 matrices.  A 16-byte unit then holds 8 columns, so a 112-byte slab row covers 56 columns and
// the per-row work (id load, d shared-memory wavefronts) is amortised over twice the columns; c4 / ldx4 / ldy4 count
// 16-byte units, accumulation is fp32, the result is rounded to nearest even.  No bias / ReLU / projection.
template <int W4, int LANES, int THREADS, int MINB, int DEPTH, bool TMA, int NBUF, bool PROJ = false, bool XB = false>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_slab_kernel(const __grid_constant__ CUtensorMap tmX, const int32_t* __restrict__ header,
                 const uint4* __restrict__ ell_col, const float* __restrict__ plan_nd,
                 const int32_t* __restrict__ graph_ptr, const float4* __restrict__ X, float4* __restrict__ Y,
                 int n_graphs, int c4, int64_t ldx4, int64_t ldy4, const float4* __restrict__ bias, int relu,
                 int n_slabs, int rows_cap, int out_bf16, const SlabProj proj, int n_cols_b16) {
    extern __shared__ __align__(128) float4 sbuf_all[];   // NBUF x [rows_cap][W4]: rows, the all-zero row, one pad row; mbarriers
    constexpr int GROUPS = THREADS / LANES;
    constexpr uint32_t BOX_BYTES = kBoxRows * W4 * 16;
    const int tid = threadIdx.x;
    const int lg = tid & (LANES - 1);                     // lane within the row group
    const int gidx = tid / LANES;
    const int64_t n_items = (int64_t)n_graphs * n_slabs;
    const bool slot7 = __ldg(header) > 7;
    const uint32_t BUF_BYTES = (uint32_t)rows_cap * W4 * 16;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sbuf_all);
    const uint32_t bar0 = sbase + NBUF * BUF_BYTES;
    // PROJ: W slab transposed, wsm[k][W4] float4 (k < 4), 16 bytes past the barriers
    float4* wsm = reinterpret_cast<float4*>(reinterpret_cast<char*>(sbuf_all) + NBUF * BUF_BYTES + 16);
    uint32_t parity = 0;                                  // bit b = phase of buffer b's mbarrier
    if (TMA) {
        if (tid == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
            for (int b = 0; b < NBUF; ++b) mbar_init(bar0 + 8 * b, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // stage one (graph, slab) item into buffer b.  Full 128-row boxes go
    // through TMA (one thread), the tail rows (all rows without TMA) through cp.async: thread -> (row, float4)
    auto issue = [&](int64_t wi, int b0, int ng, int b) {
        const int col = (int)(wi % n_slabs) * W4;
        const int nvi = min(W4, c4 - col);
        const int n_box = TMA ? ng / kBoxRows : 0;
        if (TMA && tid == 0 && n_box > 0) {
            mbar_expect_tx(bar0 + 8 * b, (uint32_t)n_box * BOX_BYTES);
            for (int k = 0; k < n_box; ++k)
                tma_load_2d(sbase + b * BUF_BYTES + (uint32_t)k * BOX_BYTES, &tmX, col * (XB ? 8 : 4), b0 + k * kBoxRows, bar0 + 8 * b);
        }
        const int rt = n_box * kBoxRows;
        const float4* src = X + (int64_t)(b0 + rt) * ldx4 + col;
        float4* buf = sbuf_all + (size_t)b * rows_cap * W4;
        float4* dst = buf + rt * W4;
        const int n_t = ng - rt;
        int r = tid / nvi, q = tid - r * nvi;
        const int dr = THREADS / nvi, dq = THREADS - dr * nvi;
        while (r < n_t) {
            cp_async16(dst + r * W4 + q, src + (int64_t)r * ldx4 + q);
            r += dr; q += dq;
            if (q >= nvi) { q -= nvi; ++r; }
        }
        if (tid < W4) buf[ng * W4 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);   // target of padded slots
        cp_async_commit();
    };

    // Neighbour lists ("plan" rows: 8 uint16 ids + one coefficient).  A row group owns a CONTIGUOUS block of rows of the
    // item and fetches the lists of LANES rows with one load per lane (lane i holds row i of the batch: LANES x 16
    // contiguous bytes), then hands them to the group one row at a time by shuffle.  The batch after the current one
    // -- of the next item at the end of an item -- is always in flight, so a list is requested LANES..2 LANES rows
    // before its first use at the register cost of two rows.  (The earlier scheme loaded each row's list two rows
    // ahead: ncu showed 7.5-8.3 warps per issue stalled on exactly those loads once the bf16 gather got cheap.)
    struct Slots { uint4 c; float d; };
    Slots cur_s, nxt_s;
    cur_s.c = make_uint4(0u, 0u, 0u, 0u); cur_s.d = 0.f;
    nxt_s = cur_s;
    bool have_cur = false;

    int64_t w = blockIdx.x;
    int base = 0, n_g = 0, cur = 0;
    if (w < n_items) {
        const int g = (int)(w / n_slabs);
        base = __ldg(graph_ptr + g);
        n_g = __ldg(graph_ptr + g + 1) - base;
        if (NBUF == 2) issue(w, base, n_g, 0);
    }
    for (; w < n_items; w += gridDim.x) {
        const int s = (int)(w % n_slabs);
        const int col0 = s * W4;
        const int nv = min(W4, c4 - col0);
        float4* sbuf = sbuf_all + (size_t)cur * rows_cap * W4;
        if (NBUF == 1) issue(w, base, n_g, 0);
        // the next item's extent, fetched now so that its latency hides behind this item
        const int64_t wn = w + gridDim.x;
        int base_n = 0, n_g_n = 0;
        if (wn < n_items) {
            const int gn = (int)(wn / n_slabs);
            base_n = __ldg(graph_ptr + gn);
            n_g_n = __ldg(graph_ptr + gn + 1) - base_n;
        }
        if (NBUF == 2) {
            if (wn < n_items) issue(wn, base_n, n_g_n, cur ^ 1);
            else cp_async_commit();
        }
        const bool active = lg < nv;
        // batch j of an item with ng rows dealt in blocks of rpg rows: lane lg fetches row gidx * rpg + j * LANES + lg;
        // rows past the block or the graph get a list that points at the all-zero row with coefficient 0
        auto load_batch = [&](Slots& S, int b0, int ng, int rpg, int j) {
            const int i = j * LANES + lg;
            const int row = gidx * rpg + i;
            if (i < rpg && row < ng) {
                const int64_t v = (int64_t)b0 + row;
                S.c = __ldg(ell_col + v);
                S.d = __ldg(plan_nd + v);
            } else {
                const uint32_t pad = (uint32_t)ng | ((uint32_t)ng << 16);
                S.c = make_uint4(pad, pad, pad, pad);
                S.d = 0.f;
            }
        };
        const int rpg = ceil_div(n_g, GROUPS);
        const int rpg_n = ceil_div(n_g_n, GROUPS);
        const int n_batches = ceil_div(rpg, LANES);
        if (!have_cur) { load_batch(cur_s, base, n_g, rpg, 0); have_cur = true; }   // first item only: cold round trip
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), b4h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (XB) {                                         // a 16-byte unit = 8 columns = two float4 of the fp32 bias
            if (bias && active) {
                const int c = (col0 + lg) * 8;
                if (c < n_cols_b16) b4 = __ldg(bias + (col0 + lg) * 2);
                if (c + 4 < n_cols_b16) b4h = __ldg(bias + (col0 + lg) * 2 + 1);
            }
        } else if (bias && active) b4 = __ldg(bias + col0 + lg);
        if (PROJ) {                                       // W slab, transposed: wsm[k][column] (0 beyond n_out / n_cols)
            float* wf = reinterpret_cast<float*>(wsm);
            for (int i = tid; i < 4 * W4 * 4; i += THREADS) {
                const int k = i / (W4 * 4), c = i - k * (W4 * 4);
                const int col = col0 * 4 + c;
                wf[i] = (k < proj.n_out && col < c4 * 4) ? __ldg(proj.W + (int64_t)col * proj.n_out + k) : 0.f;
            }
        }

        if (NBUF == 2) cp_async_wait<1>(); else cp_async_wait<0>();
        if (TMA && n_g >= kBoxRows) { mbar_wait(bar0 + 8 * cur, (parity >> cur) & 1u); parity ^= 1u << cur; }
        __syncthreads();

        // padded slots point at the all-zero row, so they add exactly 0 and a non-finite source row can only reach
        // rows that really reference it.  Lanes >= nv read past their row (the pad row keeps that in bounds, and
        // those banks are free in the same LDS phase) and never store.
        const float4* sl = sbuf + lg;
        auto gather = [&](uint32_t u) { return sl[u * W4]; };
        auto add4 = [](float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; };
        auto row_out = [&](const Slots& S, int row, bool valid) {
            if constexpr (XB) {
                float a[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = 0.f;
                // add.rn.f32.bf16 (PTX 8.6, sm_100): fp32 accumulator += bf16 operand taken straight from a register half
                // (SASS FHADD.BF16 with .H0 / .H1 selectors) -- no unpack instructions, one issue slot per element
                auto add8 = [&](const float4& v) {
                    const uint32_t w[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\t"
                            "add.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
                            : "+f"(a[2 * i]), "+f"(a[2 * i + 1]) : "r"(w[i]));
                    }
                };
                {
                    const float4 v0 = gather(S.c.x & 0xffffu), v1 = gather(S.c.x >> 16), v2 = gather(S.c.y & 0xffffu),
                                 v3 = gather(S.c.y >> 16);
                    add8(v0); add8(v1); add8(v2); add8(v3);
                }
                {
                    const float4 v4 = gather(S.c.z & 0xffffu), v5 = gather(S.c.z >> 16), v6 = gather(S.c.w & 0xffffu);
                    add8(v4); add8(v5); add8(v6);
                }
                if (slot7) add8(gather(S.c.w >> 16));
                if (active && valid) {
                    const float bb[8] = {b4.x, b4.y, b4.z, b4.w, b4h.x, b4h.y, b4h.z, b4h.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        a[i] = fmaf(a[i], S.d, bb[i]);
                        if (relu) a[i] = fmaxf(a[i], 0.f);
                    }
                    __nv_bfloat162 o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
                    reinterpret_cast<uint4*>(Y)[(int64_t)(base + row) * ldy4 + col0 + lg] = *reinterpret_cast<const uint4*>(o);
                }
                return;
            }
            float4 acc = gather(S.c.x & 0xffffu);
            {
                const float4 v1 = gather(S.c.x >> 16), v2 = gather(S.c.y & 0xffffu), v3 = gather(S.c.y >> 16);
                add4(acc, v1); add4(acc, v2); add4(acc, v3);
            }
            {
                const float4 v4 = gather(S.c.z & 0xffffu), v5 = gather(S.c.z >> 16), v6 = gather(S.c.w & 0xffffu);
                add4(acc, v4); add4(acc, v5); add4(acc, v6);
            }
            if (slot7) add4(acc, gather(S.c.w >> 16));
            acc.x = fmaf(acc.x, S.d, b4.x); acc.y = fmaf(acc.y, S.d, b4.y);
            acc.z = fmaf(acc.z, S.d, b4.z); acc.w = fmaf(acc.w, S.d, b4.w);
            if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
            if (active && valid) {
                if (out_bf16) {                           // Y is a bf16 matrix: ldy4 counts 4-element (8-byte) units
                    __nv_bfloat162 o[2] = {__floats2bfloat162_rn(acc.x, acc.y), __floats2bfloat162_rn(acc.z, acc.w)};
                    reinterpret_cast<uint2*>(Y)[(int64_t)(base + row) * ldy4 + col0 + lg] = *reinterpret_cast<const uint2*>(o);
                } else {
                    Y[(int64_t)(base + row) * ldy4 + col0 + lg] = acc;
                }
            }
            if (PROJ) {
                // this lane's 4 columns against the W slab, then a sum over the row group's LANES lanes
                float p[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 wk = wsm[k * W4 + (active ? lg : 0)];
                    p[k] = active ? fmaf(acc.x, wk.x, fmaf(acc.y, wk.y, fmaf(acc.z, wk.z, acc.w * wk.w))) : 0.f;
                }
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) p[k] += __shfl_xor_sync(0xffffffffu, p[k], o);
                }
                if (lg == 0 && valid) proj.Tpart[(int64_t)s * proj.n_rows + base + row] = make_float4(p[0], p[1], p[2], p[3]);
            }
        };
        const int row_begin = gidx * rpg;
        for (int j = 0; j < n_batches; ++j) {
            if (j + 1 < n_batches) load_batch(nxt_s, base, n_g, rpg, j + 1);
            else if (wn < n_items) load_batch(nxt_s, base_n, n_g_n, rpg_n, 0);
#pragma unroll 2
            for (int k = 0; k < LANES; ++k) {
                const int i = j * LANES + k;
                if (i >= rpg) break;                      // CTA-uniform
                Slots sk;                                 // row k of the batch, from lane k of the group (all 32 lanes shuffle)
                sk.c.x = __shfl_sync(0xffffffffu, cur_s.c.x, k, LANES);
                sk.c.y = __shfl_sync(0xffffffffu, cur_s.c.y, k, LANES);
                sk.c.z = __shfl_sync(0xffffffffu, cur_s.c.z, k, LANES);
                sk.c.w = __shfl_sync(0xffffffffu, cur_s.c.w, k, LANES);
                sk.d = __shfl_sync(0xffffffffu, cur_s.d, k, LANES);
                const int row = row_begin + i;
                row_out(sk, row, row < n_g);
            }
            cur_s = nxt_s;
        }
        __syncthreads();                                  // the buffer may be overwritten by the next item
        base = base_n; n_g = n_g_n;
        if (NBUF == 2) cur ^= 1;
    }
    cp_async_wait<0>();
}

// T[row, k] = sum over slabs of Tpart[slab][row][k], slabs in index order (deterministic)
__global__ void __launch_bounds__(256)
proj_reduce_kernel(const float4* __restrict__ Tpart, int n_slabs, int64_t n_rows, int n_out, float* __restrict__ T, int64_t ldt) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n_slabs; ++i) {
        const float4 v = __ldg(Tpart + (int64_t)i * n_rows + row);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float o[4] = {s.x, s.y, s.z, s.w};
    for (int k = 0; k < n_out; ++k) T[row * ldt + k] = o[k];
}

static size_t plan_bytes(int64_t n_rows) { return kPlanHeader + (size_t)n_rows * (kEll * sizeof(uint16_t) + sizeof(float)); }

// GMC_SLAB_VARIANT (tuning aid): B (default) = two CTAs of 512 threads per SM, 112-byte slabs, TMA staging;
// A = one CTA per SM, 224-byte slabs, TMA; C = B with cp.async staging; D = B with prefetch depth 3; E = A with cp.async
static int slab_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GMC_SLAB_VARIANT");
        v = 1;
        if (e && e[0] >= 'A' && e[0] <= 'E') v = e[0] - 'A';
    }
    return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn slab_encode_fn() {
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

template <int W4, int LANES, int THREADS, int MINB, int DEPTH, bool TMA, int NBUF, bool PROJ = false, bool XB = false>
static int slab_launch_one(const void* plan, const int32_t* graph_ptr, int n_graphs, int max_nodes, const float4* X4,
                           float4* Y4, int64_t n_rows, int c4, int64_t ldx4, int64_t ldy4, const float4* b4, int relu,
                           cudaStream_t s, int* launched, int out_bf16 = 0, SlabProj proj = SlabProj{nullptr, nullptr, 0, 0},
                           int n_cols_b16 = 0) {
    const int rows_cap = (max_nodes + 2 + 7) & ~7;         // buffers stay 128-byte aligned (TMA destination)
    const size_t smem = (size_t)NBUF * rows_cap * W4 * sizeof(float4) + 16 + (PROJ ? 4 * W4 * sizeof(float4) : 0);
    if (smem + 1024 > (228 * 1024) / MINB || smem > kSlabSmemMax) return GMC_OK;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (TMA) {
        EncodeTiledFn enc = slab_encode_fn();
        if (!enc) return GMC_OK;                          // no driver entry point: let the caller take another kernel
        // XB: the real column count bounds the map, so the last unit's pad columns arrive as zeros
        cuuint64_t dims[2] = {XB ? (cuuint64_t)n_cols_b16 : (cuuint64_t)c4 * 4, (cuuint64_t)n_rows};
        cuuint64_t strides[1] = {(cuuint64_t)ldx4 * 16};
        cuuint32_t box[2] = {W4 * (XB ? 8 : 4), kBoxRows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, XB ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float4*>(X4), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("gmc_spmm_batched_f32: cuTensorMapEncodeTiled failed (%d)", (int)r); return GMC_ERR_INVALID_ARG; }
    }
    static bool attr = false;
    if (!attr) {
        GMC_CUDA(cudaFuncSetAttribute(spmm_slab_kernel<W4, LANES, THREADS, MINB, DEPTH, TMA, NBUF, PROJ, XB>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSlabSmemMax));
        GMC_CUDA(cudaFuncSetAttribute(spmm_slab_kernel<W4, LANES, THREADS, MINB, DEPTH, TMA, NBUF, PROJ, XB>,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr = true;
    }
    const int32_t* header = reinterpret_cast<const int32_t*>(plan);
    const uint4* ecol = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(plan) + kPlanHeader);
    const float* pnd = reinterpret_cast<const float*>(ecol + n_rows);
    const int n_slabs = ceil_div(c4, W4);
    const int64_t items = (int64_t)n_graphs * n_slabs;
    const int64_t slots = (int64_t)sm_count() * MINB;
    const int grid = (int)(items < slots ? items : slots);
    spmm_slab_kernel<W4, LANES, THREADS, MINB, DEPTH, TMA, NBUF, PROJ, XB><<<grid, THREADS, smem, s>>>(
        tm, header, ecol, pnd, graph_ptr, X4, Y4, n_graphs, c4, ldx4, ldy4, b4, relu, n_slabs, rows_cap, out_bf16, proj,
        n_cols_b16);
    GMC_LAUNCH_CHECK();
    *launched = 1;
    return GMC_OK;
}

static int slab_launch(const void* plan, const int32_t* graph_ptr, int n_graphs, int max_nodes, const float* X, void* Y,
                       int64_t n_rows, int n_cols, int64_t ldx, int64_t ldy, const float* bias, int relu,
                       cudaStream_t s, int* launched, int out_bf16 = 0) {
    *launched = 0;
    if (!plan || !graph_ptr || n_graphs <= 0 || max_nodes < 128 || n_cols < 16 || n_cols % 4 || ldx % 4 || ldy % 4)
        return GMC_OK;
    if (!aligned16(X) || (reinterpret_cast<uintptr_t>(Y) & (out_bf16 ? 7u : 15u)) || (bias && !aligned16(bias)) || !aligned16(plan))
        return GMC_OK;
    const int c4 = n_cols / 4;
    const float4* X4 = reinterpret_cast<const float4*>(X);
    float4* Y4 = reinterpret_cast<float4*>(Y);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#define GMC_SLAB(W4, LANES, THREADS, MINB, DEPTH, TMA, NBUF)                                                             \
    {                                                                                                              \
        const int rc = slab_launch_one<W4, LANES, THREADS, MINB, DEPTH, TMA, NBUF>(plan, graph_ptr, n_graphs,      \
                                                                             max_nodes, X4, Y4, n_rows, c4,       \
                                                                             ldx / 4, ldy / 4, b4, relu, s,       \
                                                                             launched, out_bf16);                  \
        if (rc != GMC_OK || *launched) return rc;                                                                  \
    }
    switch (slab_variant()) {
        case 0: GMC_SLAB(14, 16, 1024, 1, 2, true, 1) break;
        case 2: GMC_SLAB(7, 8, 1024, 1, 2, true, 2) break;
        case 3: GMC_SLAB(7, 8, 768, 1, 4, true, 2) break;
        case 4: GMC_SLAB(7, 8, 512, 2, 2, false, 1) break;
        default: break;
    }
    GMC_SLAB(7, 8, 512, 2, 2, true, 1)
    GMC_SLAB(14, 16, 1024, 1, 2, true, 1)
    GMC_SLAB(7, 8, 1024, 1, 2, true, 1)
    GMC_SLAB(4, 4, 1024, 1, 2, true, 1)
#undef GMC_SLAB
    return GMC_OK;
}

}  // namespace gmc

extern "C" {

size_t gmc_spmm_plan_bytes(int64_t n_rows) { return gmc::plan_bytes(n_rows); }

// Builds the ELL plan of a block-diagonal batch.  *overflow (device int32) is set to 1 when some row has more
// than 8 neighbours, in which case the plan must not be used (pass plan = NULL to gmc_spmm_batched_f32).
int gmc_spmm_plan_build(const int32_t* rowptr, const int32_t* colidx, const float* norm_src, const float* norm_dst,
                        const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, void* plan, int32_t* overflow,
                        void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && colidx && norm_src && norm_dst && graph_ptr && plan && overflow,
                "gmc_spmm_plan_build: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0, "gmc_spmm_plan_build: bad sizes");
    GMC_REQUIRE(aligned16(plan), "gmc_spmm_plan_build: plan must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    GMC_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int32_t), s));
    GMC_CUDA(cudaMemsetAsync(plan, 0, kPlanHeader, s));
    if (n_rows == 0) return GMC_OK;
    int32_t* header = reinterpret_cast<int32_t*>(plan);
    uint4* ecol = reinterpret_cast<uint4*>(reinterpret_cast<char*>(plan) + kPlanHeader);
    float* pcoef = reinterpret_cast<float*>(ecol + n_rows);
    ell_pack_kernel<<<(unsigned)ceil_div<int64_t>(n_rows, 256), 256, 0, s>>>(rowptr, colidx, norm_src, norm_dst,
                                                                            graph_ptr, n_graphs, n_rows, header, ecol,
                                                                            pcoef, overflow);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// Y = act(A_hat X + bias) for a block-diagonal batch, A_hat given by its per-edge values `coef`
// (gmc_edge_coef_f32).  `plan` (nullable) is the ELL plan of the same batch; with it, graphs that fit the
// shared-memory slab buffers take the staged kernel.  Results are identical to
// gmc_spmm_symnorm_f32(rowptr, colidx, coef, NULL, NULL, ...), only the schedule differs.
int gmc_spmm_batched_f32(const int32_t* rowptr, const int32_t* colidx, const float* coef, const int32_t* graph_ptr,
                         int32_t n_graphs, int32_t max_nodes, const void* plan, const float* X, float* Y,
                         int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy, const float* bias, int32_t relu,
                         void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && colidx && coef && X && Y, "gmc_spmm_batched_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols, "gmc_spmm_batched_f32: bad sizes");
    GMC_REQUIRE(X != Y, "gmc_spmm_batched_f32: in-place SpMM is not supported");
    if (n_rows == 0) return GMC_OK;
    int launched = 0;
    const int rc = slab_launch(plan, graph_ptr, n_graphs, max_nodes, X, Y, n_rows, n_cols, ldx, ldy, bias, relu,
                               as_stream(stream), &launched);
    if (rc != GMC_OK || launched) return rc;
    return gmc_spmm_symnorm_f32(rowptr, colidx, coef, nullptr, nullptr, X, Y, n_rows, n_cols, ldx, ldy, bias, relu, stream);
}

// Same product with the output rounded to bf16 on the way out (Y: bf16 matrix, ldy in elements): the B operand of
// the bf16 weight-gradient GEMM, written once instead of fp32 + a conversion pass.  Slab kernel only: returns
// GMC_ERR_UNSUPPORTED when the batch has no usable plan (the caller then runs the fp32 SpMM and gmc_f32_to_bf16).
int gmc_spmm_batched_bf16out(const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, const void* plan,
                             const float* X, void* Y, int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                             void* stream) {
    using namespace gmc;
    GMC_REQUIRE(graph_ptr && X && Y, "gmc_spmm_batched_bf16out: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols, "gmc_spmm_batched_bf16out: bad sizes");
    if (n_rows == 0) return GMC_OK;
    int launched = 0;
    const int rc = slab_launch(plan, graph_ptr, n_graphs, max_nodes, X, Y, n_rows, n_cols, ldx, ldy, nullptr, 0,
                               as_stream(stream), &launched, 1);
    if (rc != GMC_OK) return rc;
    if (!launched) {
        set_error("gmc_spmm_batched_bf16out: the batch cannot take the slab kernel (no plan, narrow or unaligned matrix)");
        return GMC_ERR_UNSUPPORTED;
    }
    return GMC_OK;
}

// Y = A_hat X with both matrices stored in bf16 (leading dimensions in elements, multiples of 8, covering n_cols
// rounded up to 8; pad columns of X must be finite -- the engine keeps them zero): Y = act(A_hat X + bias), the
// backward aggregation dT1 = A_hat dH1pre (no bias, no ReLU) and, with bias + ReLU, the forward H1 when the layer-1
// activations are kept in bf16.  fp32 accumulation and bias, result rounded to nearest even.
// Slab kernel only: GMC_ERR_UNSUPPORTED when the batch has no usable plan.
int gmc_spmm_batched_bf16(const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, const void* plan, const void* X,
                          void* Y, int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy, const float* bias,
                          int32_t relu, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(graph_ptr && X && Y, "gmc_spmm_batched_bf16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols, "gmc_spmm_batched_bf16: bad sizes");
    GMC_REQUIRE(X != Y, "gmc_spmm_batched_bf16: in-place SpMM is not supported");
    if (n_rows == 0) return GMC_OK;
    const int c8 = (n_cols + 7) / 8;
    const bool ok = plan && n_graphs > 0 && max_nodes >= 128 && n_cols >= 16 && ldx % 8 == 0 && ldy % 8 == 0 &&
                    ldx >= (int64_t)c8 * 8 && ldy >= (int64_t)c8 * 8 && aligned16(X) && aligned16(Y) && aligned16(plan) &&
                    n_cols % 4 == 0 && (!bias || aligned16(bias));
    int launched = 0;
    if (ok) {
        const float4* X4 = reinterpret_cast<const float4*>(X);
        float4* Y4 = reinterpret_cast<float4*>(Y);
        const SlabProj none{nullptr, nullptr, 0, 0};
        const float4* b4 = reinterpret_cast<const float4*>(bias);
        int rc = GMC_OK;
        static int variant = -1;                          // GMC_SLAB16_VARIANT = B | C (default) | D | E
        if (variant < 0) {
            const char* e = getenv("GMC_SLAB16_VARIANT");
            // default C (one CTA of 1024 threads per SM, two 112-byte slab buffers): 2.12 ms = 0.60 of the HBM roofline
            // inside the standard-layer-1 step against 2.37 ms = 0.54 for B (two CTAs of 512 threads), 1.90 vs 2.49 ms alone;
            // D / E = 64-column slabs, whose row pieces are full 128-byte lines: 2.13 ms
            variant = (e && e[0] == 'B') ? 0 : (e && e[0] == 'D') ? 2 : (e && e[0] == 'E') ? 3 : 1;
        }
        if (variant == 2) {                               // D: 64-column slabs (full 128-byte lines per row piece), one CTA per SM
            rc = slab_launch_one<8, 8, 1024, 1, 2, true, 1, false, true>(plan, graph_ptr, n_graphs, max_nodes, X4, Y4, n_rows, c8,
                                                                         ldx / 8, ldy / 8, b4, relu, as_stream(stream),
                                                                         &launched, 1, none, n_cols);
            if (rc != GMC_OK) return rc;
        }
        if (variant == 3) {                               // E: the same with 512 threads
            rc = slab_launch_one<8, 8, 512, 1, 2, true, 1, false, true>(plan, graph_ptr, n_graphs, max_nodes, X4, Y4, n_rows, c8,
                                                                        ldx / 8, ldy / 8, b4, relu, as_stream(stream),
                                                                        &launched, 1, none, n_cols);
            if (rc != GMC_OK) return rc;
        }
        if (variant == 1) {
            rc = slab_launch_one<7, 8, 1024, 1, 2, true, 2, false, true>(plan, graph_ptr, n_graphs, max_nodes, X4, Y4, n_rows, c8,
                                                                         ldx / 8, ldy / 8, b4, relu, as_stream(stream),
                                                                         &launched, 1, none, n_cols);
            if (rc != GMC_OK) return rc;
        }
        if (!launched)
        rc = slab_launch_one<7, 8, 512, 2, 2, true, 1, false, true>(plan, graph_ptr, n_graphs, max_nodes, X4, Y4, n_rows, c8,
                                                                        ldx / 8, ldy / 8, b4, relu, as_stream(stream),
                                                                        &launched, 1, none, n_cols);
        if (rc != GMC_OK) return rc;
        if (!launched) {
            rc = slab_launch_one<4, 4, 1024, 1, 2, true, 1, false, true>(plan, graph_ptr, n_graphs, max_nodes, X4, Y4, n_rows,
                                                                         c8, ldx / 8, ldy / 8, b4, relu, as_stream(stream),
                                                                         &launched, 1, none, n_cols);
            if (rc != GMC_OK) return rc;
        }
    }
    if (!launched) {
        set_error("gmc_spmm_batched_bf16: the batch cannot take the slab kernel (no plan, graphs too large, narrow or unaligned matrix)");
        return GMC_ERR_UNSUPPORTED;
    }
    return GMC_OK;
}

size_t gmc_spmm_batched_fused_workspace_bytes(int64_t n_rows, int32_t n_cols) {
    return (size_t)gmc::ceil_div(n_cols / 4, 7) * (size_t)n_rows * sizeof(float4);
}

// Y = act(A_hat X + bias) AND T = Y W (W [n_cols, n_out <= 4]) for a block-diagonal batch with an ELL plan: the slab
// kernel with the skinny projection of the second GraphConv layer (TrainingNeural.py:83) folded into its epilogue
// (per-slab partials in `workspace`, reduced in slab order: deterministic).  GMC_ERR_UNSUPPORTED when the batch cannot
// take the slab kernel or n_out > 4 -- gmc_spmm_fused_skinny_f32 is the general form.
int gmc_spmm_batched_fused_skinny_f32(const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, const void* plan,
                                      const float* X, float* Y, int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                                      const float* bias, int32_t relu, const float* W, int32_t n_out, float* T,
                                      int64_t ldt, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(graph_ptr && X && Y && W && T, "gmc_spmm_batched_fused_skinny_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols && ldt >= n_out && n_out >= 1,
                "gmc_spmm_batched_fused_skinny_f32: bad sizes");
    if (n_rows == 0) return GMC_OK;
    const bool ok = plan && n_out <= 4 && n_graphs > 0 && max_nodes >= 128 && n_cols >= 16 && n_cols % 4 == 0 &&
                    ldx % 4 == 0 && ldy % 4 == 0 && aligned16(X) && aligned16(Y) && (!bias || aligned16(bias)) &&
                    aligned16(plan) && workspace && aligned16(workspace) &&
                    workspace_bytes >= gmc_spmm_batched_fused_workspace_bytes(n_rows, n_cols);
    int launched = 0;
    if (ok) {
        SlabProj proj{W, reinterpret_cast<float4*>(workspace), n_rows, n_out};
        const int rc = slab_launch_one<7, 8, 512, 2, 2, true, 1, true>(
            plan, graph_ptr, n_graphs, max_nodes, reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(Y), n_rows,
            n_cols / 4, ldx / 4, ldy / 4, reinterpret_cast<const float4*>(bias), relu, as_stream(stream), &launched, 0, proj);
        if (rc != GMC_OK) return rc;
    }
    if (!launched) {
        set_error("gmc_spmm_batched_fused_skinny_f32: the batch cannot take the slab kernel (plan, n_out <= 4, graphs of "
                  "128..1006 nodes, 16-byte aligned rows, workspace)");
        return GMC_ERR_UNSUPPORTED;
    }
    const int n_slabs = ceil_div(n_cols / 4, 7);
    proj_reduce_kernel<<<(unsigned)ceil_div<int64_t>(n_rows, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(workspace), n_slabs, n_rows, n_out, T, ldt);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

}  // extern "C"
