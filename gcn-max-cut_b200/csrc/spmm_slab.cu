// spmm_slab.cu -- shared-memory-staged SpMM for block-diagonal batches of small graphs.
//
// The warp-per-row kernel (spmm.cu) reads every source row d times from L2: at n=1000, C=500,
// d=7 it moves 13 TB/s through L2 for 3.3 TB/s of HBM traffic and sits on the L2 roof (ncu,
// profiles/r01_v1_*).  Here one persistent CTA owns a (graph, column-slab) work item:
//
//   * the graph's source rows, restricted to a slab of W4 float4 columns, are staged ONCE into
//     shared memory with cp.async (16-byte LDGSTS, L1 bypass), double-buffered so the loads of the
//     next item overlap the gathers of the current one;
//   * a group of GROUP lanes owns one output row and gathers its neighbours' slab rows from shared
//     memory (LDS.128, one row group per 128-bit phase -> conflict-free);
//   * neighbour lists come from an ELL "plan" (8 padded (local col, coef) slots per row, built once per
//     static batch by gmc_spmm_plan_build): no rowptr->colidx dependent load chain, and the next row's
//     slots are prefetched while the current row is accumulated;
//   * the output slab is written with 128-bit stores (bias + ReLU fused).
//
// L2->SM traffic drops from (d+1) x to ~2 x the matrix; HBM traffic stays the algorithmic
// 8*N*C + 4*nnz + 4*(N+1) bytes.  Used when every graph of the batch fits the slab buffers and has
// max degree <= 8; otherwise gmc_spmm_batched_f32 runs the warp-per-row kernel (same results).
#include "common.cuh"

namespace gmc {

constexpr int kSlabThreads = 768;
constexpr int kEll = 8;                                   // padded neighbour slots per row
constexpr size_t kSlabSmemMax = 227 * 1024;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- plan: ELL packing of a block-diagonal CSR -----------------------------------------------
// plan layout: int32 col[n_rows][8] (LOCAL node ids, padded with n_g = the kernel's all-zero row) followed by
//              float coef[n_rows][8] (padded with 0).  One warp per 4 rows.
__global__ void ell_pack_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                const float* __restrict__ coef, const int32_t* __restrict__ graph_ptr, int n_graphs,
                                int64_t n_rows, int32_t* __restrict__ ell_col, float* __restrict__ ell_coef,
                                int* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // slot index
    if (i >= n_rows * kEll) return;
    const int64_t v = i / kEll;
    const int j = (int)(i % kEll);
    const int g = find_graph(graph_ptr, n_graphs, v);
    const int base = __ldg(graph_ptr + g);
    const int e0 = __ldg(rowptr + v), deg = __ldg(rowptr + v + 1) - e0;
    if (deg > kEll && j == 0) *overflow = 1;
    if (j < deg) {
        ell_col[i] = __ldg(colidx + e0 + j) - base;
        ell_coef[i] = __ldg(coef + e0 + j);
    } else {
        ell_col[i] = __ldg(graph_ptr + g + 1) - base;     // the slab buffers keep an all-zero row at index n_g
        ell_coef[i] = 0.f;
    }
}

// W4: float4 columns per slab, GROUP: lanes per output row (power of two >= W4, <= 8)
template <int W4, int GROUP>
__global__ void __launch_bounds__(kSlabThreads, 1)
spmm_slab_kernel(const int4* __restrict__ ell_col, const float4* __restrict__ ell_coef,
                 const int32_t* __restrict__ graph_ptr, const float4* __restrict__ X, float4* __restrict__ Y,
                 int n_graphs, int c4, int64_t ldx4, int64_t ldy4, const float4* __restrict__ bias, int relu,
                 int n_slabs, int rows_cap) {
    extern __shared__ float4 sbuf[];                      // [2][rows_cap][W4]
    const int tid = threadIdx.x;
    const int lg = tid & (GROUP - 1);                     // lane within the row group
    constexpr int GROUPS = kSlabThreads / GROUP;
    const int gidx = tid / GROUP;
    const int64_t n_items = (int64_t)n_graphs * n_slabs;

    auto issue = [&](int64_t w, int buf) {
        const int g = (int)(w / n_slabs), s = (int)(w % n_slabs);
        const int base = __ldg(graph_ptr + g);
        const int n_g = __ldg(graph_ptr + g + 1) - base;
        const int col0 = s * W4;
        const int nv = min(W4, c4 - col0);
        float4* dst = sbuf + (size_t)buf * rows_cap * W4;
        const float4* src = X + (int64_t)base * ldx4 + col0;
        // thread -> (row, float4) with the float4 index fastest, no integer division in the loop
        int r = tid / nv, q = tid - r * nv;
        const int dr = kSlabThreads / nv, dq = kSlabThreads - dr * nv;
        while (r < n_g) {
            cp_async16(dst + r * W4 + q, src + (int64_t)r * ldx4 + q);
            r += dr; q += dq;
            if (q >= nv) { q -= nv; ++r; }
        }
        if (tid < W4) dst[n_g * W4 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);   // zero row for padded slots
    };

    int buf = 0;
    int64_t w = blockIdx.x;
    if (w < n_items) issue(w, 0);
    cp_async_commit();
    for (; w < n_items; w += gridDim.x) {
        const int g = (int)(w / n_slabs), s = (int)(w % n_slabs);
        const int base = __ldg(graph_ptr + g);
        const int n_g = __ldg(graph_ptr + g + 1) - base;
        const int col0 = s * W4;
        const int nv = min(W4, c4 - col0);
        const bool active = lg < nv;
        // neighbour lists of this row group's first TWO rows: issued before waiting on the slab; every list is
        // re-loaded one full (two-row) iteration ahead of its use, which covers the L2 latency with 24 warps/SM
        struct Slots { int4 c_lo, c_hi; float4 a_lo, a_hi; };
        auto load_slots = [&](Slots& S, int row) {
            const int64_t slot = ((int64_t)base + row) * 2;
            S.c_lo = __ldg(ell_col + slot); S.c_hi = __ldg(ell_col + slot + 1);
            S.a_lo = __ldg(ell_coef + slot); S.a_hi = __ldg(ell_coef + slot + 1);
        };
        Slots S0, S1;
        const int r0 = gidx;
        if (r0 < n_g) load_slots(S0, r0);
        if (r0 + GROUPS < n_g) load_slots(S1, r0 + GROUPS);
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && active) b4 = __ldg(bias + col0 + lg);

        const int64_t wn = w + gridDim.x;
        if (wn < n_items) issue(wn, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();                               // everything but the newest group has landed
        __syncthreads();

        const float4* src = sbuf + (size_t)buf * rows_cap * W4 + lg;
        // padded slots carry coef 0 and point at the all-zero row, so they add exactly 0 and a non-finite
        // source row can only reach rows that really reference it
        auto row_out = [&](const Slots& S, int row) {
            if (!active) return;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v0 = src[S.c_lo.x * W4], v1 = src[S.c_lo.y * W4], v2 = src[S.c_lo.z * W4], v3 = src[S.c_lo.w * W4];
            const float4 v4 = src[S.c_hi.x * W4], v5 = src[S.c_hi.y * W4], v6 = src[S.c_hi.z * W4], v7 = src[S.c_hi.w * W4];
            fma4(acc, S.a_lo.x, v0); fma4(acc, S.a_lo.y, v1); fma4(acc, S.a_lo.z, v2); fma4(acc, S.a_lo.w, v3);
            fma4(acc, S.a_hi.x, v4); fma4(acc, S.a_hi.y, v5); fma4(acc, S.a_hi.z, v6); fma4(acc, S.a_hi.w, v7);
            acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
            if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
            Y[(int64_t)(base + row) * ldy4 + col0 + lg] = acc;
        };
        for (int r = r0; r < n_g; r += 2 * GROUPS) {
            row_out(S0, r);
            if (r + 2 * GROUPS < n_g) load_slots(S0, r + 2 * GROUPS);
            if (r + GROUPS < n_g) {
                row_out(S1, r + GROUPS);
                if (r + 3 * GROUPS < n_g) load_slots(S1, r + 3 * GROUPS);
            }
        }
        __syncthreads();                                  // buffer `buf` may be overwritten by the next issue
        buf ^= 1;
    }
    cp_async_wait<0>();
}

static size_t plan_bytes(int64_t n_rows) { return (size_t)n_rows * kEll * (sizeof(int32_t) + sizeof(float)); }

static int slab_launch(const void* plan, const int32_t* graph_ptr, int n_graphs, int max_nodes, const float* X, float* Y,
                       int64_t n_rows, int n_cols, int64_t ldx, int64_t ldy, const float* bias, int relu,
                       cudaStream_t s, int* launched) {
    *launched = 0;
    if (!plan || !graph_ptr || n_graphs <= 0 || max_nodes < 128 || n_cols < 64 || n_cols % 4 || ldx % 4 || ldy % 4)
        return GMC_OK;
    if (!aligned16(X) || !aligned16(Y) || (bias && !aligned16(bias)) || !aligned16(plan)) return GMC_OK;
    const int c4 = n_cols / 4;
    const int4* ecol = reinterpret_cast<const int4*>(plan);
    const float4* ecoef = reinterpret_cast<const float4*>(reinterpret_cast<const int32_t*>(plan) + n_rows * kEll);
    const float4* X4 = reinterpret_cast<const float4*>(X);
    float4* Y4 = reinterpret_cast<float4*>(Y);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#define GMC_SLAB(W4, GROUP)                                                                                     \
    {                                                                                                           \
        const size_t smem = (size_t)2 * (max_nodes + 1) * (W4) * sizeof(float4);                                      \
        if (smem <= kSlabSmemMax) {                                                                             \
            static bool attr = false;                                                                           \
            if (!attr) {                                                                                        \
                GMC_CUDA(cudaFuncSetAttribute(spmm_slab_kernel<W4, GROUP>,                                      \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSlabSmemMax)); \
                attr = true;                                                                                    \
            }                                                                                                   \
            const int n_slabs = ceil_div(c4, (W4));                                                             \
            const int64_t items = (int64_t)n_graphs * n_slabs;                                                  \
            const int grid = (int)(items < sm_count() ? items : sm_count());                                    \
            spmm_slab_kernel<W4, GROUP><<<grid, kSlabThreads, smem, s>>>(ecol, ecoef, graph_ptr, X4, Y4,        \
                                                                         n_graphs, c4, ldx / 4, ldy / 4, b4,    \
                                                                         relu, n_slabs, max_nodes + 1);        \
            GMC_LAUNCH_CHECK();                                                                                 \
            *launched = 1;                                                                                      \
            return GMC_OK;                                                                                      \
        }                                                                                                       \
    }
    GMC_SLAB(7, 8)
    GMC_SLAB(4, 4)
    GMC_SLAB(2, 2)
#undef GMC_SLAB
    return GMC_OK;
}

}  // namespace gmc

extern "C" {

size_t gmc_spmm_plan_bytes(int64_t n_rows) { return gmc::plan_bytes(n_rows); }

// Builds the ELL plan of a block-diagonal batch.  *overflow (device int32) is set to 1 when some row has more
// than 8 neighbours, in which case the plan must not be used (pass plan = NULL to gmc_spmm_batched_f32).
int gmc_spmm_plan_build(const int32_t* rowptr, const int32_t* colidx, const float* coef, const int32_t* graph_ptr,
                        int32_t n_graphs, int64_t n_rows, void* plan, int32_t* overflow, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && colidx && coef && graph_ptr && plan && overflow, "gmc_spmm_plan_build: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0, "gmc_spmm_plan_build: bad sizes");
    cudaStream_t s = as_stream(stream);
    GMC_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int32_t), s));
    if (n_rows == 0) return GMC_OK;
    int32_t* ecol = reinterpret_cast<int32_t*>(plan);
    float* ecoef = reinterpret_cast<float*>(ecol + n_rows * kEll);
    const int64_t total = n_rows * kEll;
    ell_pack_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, s>>>(rowptr, colidx, coef, graph_ptr, n_graphs,
                                                                           n_rows, ecol, ecoef, overflow);
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

}  // extern "C"
