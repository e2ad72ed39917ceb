// spmm_slab.cu -- shared-memory-staged SpMM for block-diagonal batches of small graphs.
//
// The warp-per-row kernel (spmm.cu) reads every source row d times from L2 (13 TB/s of L2
// traffic for 3.3 TB/s of HBM traffic at n=1000, C=500, d=7: L2-bandwidth bound, ncu
// profiles/r01_v1_*).  Here one persistent CTA owns a (graph, column-slab) work item:
//
//   * the graph's source rows, restricted to a slab of W4 float4 columns, are staged ONCE into
//     shared memory with cp.async (16-byte LDGSTS, L1 bypass), double-buffered so the loads of the
//     next item overlap the gathers of the current one;
//   * rows are then gathered from shared memory: a group of GROUP lanes owns one output row, its
//     neighbour (col, coef) pairs are loaded by the group's lanes and broadcast with shuffles;
//   * the output slab is written with 128-bit stores (bias + ReLU fused).
//
// L2->SM traffic drops from (d+1) x to ~2 x the matrix; HBM traffic stays the algorithmic
// 8*N*C + 4*nnz + 4*(N+1) bytes.  Used when every graph of the batch fits the slab buffers;
// otherwise gmc_spmm_symnorm_f32 falls back to the warp-per-row kernel (same results).
#include "common.cuh"

namespace gmc {

constexpr int kSlabThreads = 512;
constexpr size_t kSlabSmemMax = 227 * 1024;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// W4: float4 columns per slab, GROUP: lanes per output row (power of two >= W4, <= 8)
template <int W4, int GROUP>
__global__ void __launch_bounds__(kSlabThreads, 1)
spmm_slab_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                 const float* __restrict__ coef, const int32_t* __restrict__ graph_ptr, const float4* __restrict__ X,
                 float4* __restrict__ Y, int n_graphs, int c4, int64_t ldx4, int64_t ldy4,
                 const float4* __restrict__ bias, int relu, int n_slabs, int rows_cap) {
    extern __shared__ float4 sbuf[];                      // [2][rows_cap][W4]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int lg = lane & (GROUP - 1);                    // lane within the row group
    const int gsrc = lane & ~(GROUP - 1);                 // first lane of my group
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
        const int total = n_g * nv;
        for (int i = tid; i < total; i += kSlabThreads) {
            const int r = i / nv, q = i - r * nv;
            cp_async16(dst + r * W4 + q, src + (int64_t)r * ldx4 + q);
        }
    };

    int buf = 0;
    int64_t w = blockIdx.x;
    if (w < n_items) issue(w, 0);
    cp_async_commit();
    for (; w < n_items; w += gridDim.x) {
        const int64_t wn = w + gridDim.x;
        if (wn < n_items) issue(wn, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();                               // everything but the newest group has landed
        __syncthreads();

        const int g = (int)(w / n_slabs), s = (int)(w % n_slabs);
        const int base = __ldg(graph_ptr + g);
        const int n_g = __ldg(graph_ptr + g + 1) - base;
        const int col0 = s * W4;
        const int nv = min(W4, c4 - col0);
        const bool active = lg < nv;
        const float4* src = sbuf + (size_t)buf * rows_cap * W4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && active) b4 = __ldg(bias + col0 + lg);

        // warp-uniform trip counts so that the full-mask shuffles below are always convergent
        const int row_iters = (n_g + GROUPS - 1) / GROUPS;
        for (int it = 0; it < row_iters; ++it) {
            const int r = gidx + it * GROUPS;
            const bool valid = r < n_g;
            int e0 = 0, e1 = 0;
            if (valid) { e0 = __ldg(rowptr + base + r); e1 = __ldg(rowptr + base + r + 1); }
            const int deg = e1 - e0;
            const int deg_w = __reduce_max_sync(0xffffffffu, deg);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int eb = 0; eb < deg_w; eb += GROUP) {
                int my_c = 0;
                float my_a = 0.f;
                if (eb + lg < deg) {
                    my_c = __ldg(colidx + e0 + eb + lg) - base;
                    my_a = __ldg(coef + e0 + eb + lg);
                }
                const int cnt_w = min(GROUP, deg_w - eb);
#pragma unroll
                for (int j = 0; j < GROUP; ++j) {
                    if (j < cnt_w) {
                        const int c = __shfl_sync(0xffffffffu, my_c, gsrc + j);
                        const float a = __shfl_sync(0xffffffffu, my_a, gsrc + j);
                        if (active && eb + j < deg) fma4(acc, a, src[c * W4 + lg]);
                    }
                }
            }
            if (valid && active) {
                float4 o = acc;
                o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                Y[(int64_t)(base + r) * ldy4 + col0 + lg] = o;
            }
        }
        __syncthreads();                                  // buffer `buf` may be overwritten by the next issue
        buf ^= 1;
    }
    cp_async_wait<0>();
}

// returns 0 if the slab kernel was launched, 1 if the batch does not qualify (caller falls back)
int spmm_slab_try(const int32_t* rowptr, const int32_t* colidx, const float* coef, const int32_t* graph_ptr,
                  int n_graphs, int max_nodes, const float* X, float* Y, int64_t n_rows, int n_cols, int64_t ldx,
                  int64_t ldy, const float* bias, int relu, cudaStream_t s, int* launched) {
    *launched = 0;
    if (!graph_ptr || n_graphs <= 0 || max_nodes < 128 || n_cols < 64 || n_cols % 4 || ldx % 4 || ldy % 4) return GMC_OK;
    if (!aligned16(X) || !aligned16(Y) || (bias && !aligned16(bias))) return GMC_OK;
    const int c4 = n_cols / 4;
    const float4* X4 = reinterpret_cast<const float4*>(X);
    float4* Y4 = reinterpret_cast<float4*>(Y);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
    (void)n_rows;
#define GMC_SLAB(W4, GROUP)                                                                                     \
    {                                                                                                           \
        const size_t smem = (size_t)2 * max_nodes * (W4) * sizeof(float4);                                      \
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
            spmm_slab_kernel<W4, GROUP><<<grid, kSlabThreads, smem, s>>>(rowptr, colidx, coef, graph_ptr, X4,   \
                                                                         Y4, n_graphs, c4, ldx / 4, ldy / 4,    \
                                                                         b4, relu, n_slabs, max_nodes);        \
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

// Y = act(A_hat X + bias) for a block-diagonal batch, A_hat given by its per-edge values `coef`
// (gmc_edge_coef_f32).  graph_ptr / max_nodes describe the blocks; results are identical to
// gmc_spmm_symnorm_f32(rowptr, colidx, coef, NULL, NULL, ...), only the schedule differs.
extern "C" int gmc_spmm_batched_f32(const int32_t* rowptr, const int32_t* colidx, const float* coef,
                                    const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, const float* X,
                                    float* Y, int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                                    const float* bias, int32_t relu, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(rowptr && colidx && coef && X && Y, "gmc_spmm_batched_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && ldx >= n_cols && ldy >= n_cols, "gmc_spmm_batched_f32: bad sizes");
    GMC_REQUIRE(X != Y, "gmc_spmm_batched_f32: in-place SpMM is not supported");
    if (n_rows == 0) return GMC_OK;
    int launched = 0;
    const int rc = spmm_slab_try(rowptr, colidx, coef, graph_ptr, n_graphs, max_nodes, X, Y, n_rows, n_cols, ldx, ldy,
                                 bias, relu, as_stream(stream), &launched);
    if (rc != GMC_OK || launched) return rc;
    return gmc_spmm_symnorm_f32(rowptr, colidx, coef, nullptr, nullptr, X, Y, n_rows, n_cols, ldx, ldy, bias, relu, stream);
}
