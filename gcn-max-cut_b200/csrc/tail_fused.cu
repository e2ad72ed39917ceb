// tail_fused.cu -- everything between the two [N, hidden] passes of a training step, in ONE launch per batch:
//
//   Z   = A_hat T2 + b2                      second GraphConv layer's aggregation (TrainingNeural.py:83, dgl update_all)
//   P, s, loss_g, dZ                         softmax :84, override_fixed_nodes :87-94, apply_max_to_one_hot :96-106,
//                                            compute_loss :291-309 and their backward (loss.cu has the closed form)
//   db2 = colsum(dZ),  dT2 = A_hat dZ        backward of the aggregation (A_hat symmetric)
//
// The separate kernels (spmm_thread_row x2, cut_loss_kernel, colsum x2) move ~70 KB per 1000-node graph in five launches
// at 0.15-0.25 of the HBM roofline, and cut_loss recomputes three expf per NEIGHBOUR to get its hard label.  Graphs are
// independent and small, so one CTA owns one graph: its [n, K] rows live in shared memory (T2 -> node states -> dZ), every
// node's softmax / label is computed once, the graph's CSR is read from L1/L2 three times, and the per-graph loss and
// bias-gradient partial leave without atomics (bitwise reproducible).  Needs 3 * n * K floats of shared memory per graph
// (n <= 6000 at K = 3); larger graphs keep the separate kernels (GMC_ERR_UNSUPPORTED).
#include "common.cuh"

namespace gmc {

constexpr int kTailThreads = 256;
constexpr size_t kTailMaxSmem = 216 * 1024;

template <int K>
__global__ void __launch_bounds__(kTailThreads)
layer2_loss_fused_kernel(const float* __restrict__ T2, int64_t ldt, const int32_t* __restrict__ rowptr,
                         const int32_t* __restrict__ colidx, const float* __restrict__ coef, const float* __restrict__ vals,
                         const int32_t* __restrict__ graph_ptr, const float* __restrict__ bias2, int mode, int override_t,
                         float penalty, float C, float* __restrict__ Z_out, float* __restrict__ P_out,
                         double* __restrict__ loss, float* __restrict__ dZ_out, float* __restrict__ dT2, int64_t lddt,
                         float* __restrict__ db2_part, const float4* __restrict__ parts, int n_parts, int64_t part_stride,
                         float* __restrict__ T2_out) {
    extern __shared__ __align__(16) float tail_smem[];
    __shared__ double red_loss[kTailThreads / 32];
    __shared__ float red_db[kTailThreads / 32][K];
    pdl_prologue();
    const int g = blockIdx.x;
    const int base = __ldg(graph_ptr + g);
    const int n = __ldg(graph_ptr + g + 1) - base;
    float* sX = tail_smem;                  // T2 rows, later dZ rows
    float* sS = sX + (size_t)n * K;         // node states s (one-hot / probabilities / terminal override)
    float* sP = sS + (size_t)n * K;         // softmax probabilities
    const int tid = threadIdx.x;

    if (parts) {
        // T2 rows straight from the layer-1 GEMM's per-n-tile projection partials, added in tile order -- the arithmetic of
        // proj_reduce_kernel without its launch or the round trip of T2 through memory (K <= 4 here)
        for (int v = tid; v < n; v += kTailThreads) {
            float4 a = __ldg(parts + base + v);
            for (int t = 1; t < n_parts; ++t) {
                const float4 b = __ldg(parts + (int64_t)t * part_stride + base + v);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < 4) {
                    sX[v * K + k] = av[k];
                    if (T2_out) T2_out[(int64_t)(base + v) * ldt + k] = av[k];
                }
            }
        }
    } else {
        for (int i = tid; i < n * K; i += kTailThreads) {
            const int v = i / K, k = i - v * K;
            sX[i] = __ldg(T2 + (int64_t)(base + v) * ldt + k);
        }
    }
    __syncthreads();

    // ---- Z = A_hat T2 + b2, softmax, node state
    float b[K];
#pragma unroll
    for (int k = 0; k < K; ++k) b[k] = bias2 ? __ldg(bias2 + k) : 0.f;
    for (int v = tid; v < n; v += kTailThreads) {
        const int e0 = __ldg(rowptr + base + v), e1 = __ldg(rowptr + base + v + 1);
        float z[K];
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = 0.f;
        for (int e = e0; e < e1; ++e) {
            const int u = __ldg(colidx + e) - base;
            const float c = __ldg(coef + e);
#pragma unroll
            for (int k = 0; k < K; ++k) z[k] = fmaf(c, sX[u * K + k], z[k]);
        }
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) { z[k] += b[k]; m = fmaxf(m, z[k]); }
        float p[K], s = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { p[k] = expf(z[k] - m); s += p[k]; }
        int a = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) p[k] = p[k] / s;
#pragma unroll
        for (int k = 1; k < K; ++k) if (p[k] > p[a]) a = k;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            sP[v * K + k] = p[k];
            float st;
            if (override_t && v < 3) st = (k == v) ? 1.f : 0.f;
            else if (mode == GMC_LOSS_STE) st = (k == a) ? 1.f : 0.f;
            else st = p[k];
            sS[v * K + k] = st;
            if (Z_out) Z_out[(int64_t)(base + v) * K + k] = z[k];
            if (P_out) P_out[(int64_t)(base + v) * K + k] = p[k];
        }
    }
    __syncthreads();                        // every state is final; sX (T2) is dead from here on

    // ---- loss and dZ (closed form of loss.cu), dZ into sX
    double contrib = 0.0;
    float dbk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) dbk[k] = 0.f;
    const int nt = n < 3 ? n : 3;
    for (int v = tid; v < n; v += kTailThreads) {
        const int e0 = __ldg(rowptr + base + v), e1 = __ldg(rowptr + base + v + 1);
        float as[K];
#pragma unroll
        for (int k = 0; k < K; ++k) as[k] = 0.f;
        float wdeg = 0.f;
        for (int e = e0; e < e1; ++e) {
            const int u = __ldg(colidx + e) - base;
            const float w = vals ? __ldg(vals + e) : 1.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) as[k] = fmaf(w, sS[u * K + k], as[k]);
            wdeg += w;
        }
        float same = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) same = fmaf(sS[v * K + k], as[k], same);
        contrib += -(double)C * 0.5 * ((double)wdeg - (double)same);
        float gv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gv[k] = C * as[k];
        if (penalty != 0.f && v < 3) {
            float tot[K];
#pragma unroll
            for (int k = 0; k < K; ++k) tot[k] = 0.f;
            float pair = 0.f;
            for (int j = 0; j < nt; ++j) {
                if (j != v) {
#pragma unroll
                    for (int k = 0; k < K; ++k) tot[k] += sS[j * K + k];
                }
                for (int l = j + 1; l < nt; ++l) {
#pragma unroll
                    for (int k = 0; k < K; ++k) pair = fmaf(sS[j * K + k], sS[l * K + k], pair);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) gv[k] = fmaf(penalty, tot[k], gv[k]);
            if (v == 0) contrib += (double)penalty * (double)pair;
        }
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) dot = fmaf(sP[v * K + k], gv[k], dot);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float d = sP[v * K + k] * (gv[k] - dot);
            sX[v * K + k] = d;
            dbk[k] += d;
            if (dZ_out) dZ_out[(int64_t)(base + v) * K + k] = d;
        }
    }
    // per-graph loss and bias-gradient partial: warp sums, then a fixed-order sum over the warps
    contrib = warp_sum(contrib);
#pragma unroll
    for (int k = 0; k < K; ++k) dbk[k] = warp_sum(dbk[k]);
    if ((tid & 31) == 0) {
        red_loss[tid >> 5] = contrib;
#pragma unroll
        for (int k = 0; k < K; ++k) red_db[tid >> 5][k] = dbk[k];
    }
    __syncthreads();                        // also: every dZ row is in sX
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kTailThreads / 32; ++w) t += red_loss[w];
        loss[g] = t;
    }
    if (db2_part && tid < K) {
        float t = 0.f;
        for (int w = 0; w < kTailThreads / 32; ++w) t += red_db[w][tid];
        db2_part[(int64_t)g * K + tid] = t;
    }

    // ---- dT2 = A_hat dZ
    if (dT2) {
        for (int v = tid; v < n; v += kTailThreads) {
            const int e0 = __ldg(rowptr + base + v), e1 = __ldg(rowptr + base + v + 1);
            float t[K];
#pragma unroll
            for (int k = 0; k < K; ++k) t[k] = 0.f;
            for (int e = e0; e < e1; ++e) {
                const int u = __ldg(colidx + e) - base;
                const float c = __ldg(coef + e);
#pragma unroll
                for (int k = 0; k < K; ++k) t[k] = fmaf(c, sX[u * K + k], t[k]);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) dT2[(int64_t)(base + v) * lddt + k] = t[k];
        }
    }
}

// db2[k] = sum over graphs of their partials, in graph order
__global__ void __launch_bounds__(256)
tail_db2_reduce_kernel(const float* __restrict__ part, int n_graphs, int K, float* __restrict__ db2) {
    __shared__ float red[256];
    pdl_prologue();
    for (int k = 0; k < K; ++k) {
        // fixed assignment of graphs to threads and a fixed tree: deterministic
        float s = 0.f;
        for (int g = threadIdx.x; g < n_graphs; g += 256) s += part[(int64_t)g * K + k];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int h = 128; h > 0; h >>= 1) {
            if ((int)threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
            __syncthreads();
        }
        if (threadIdx.x == 0) db2[k] = red[0];
        __syncthreads();
    }
}

}  // namespace gmc

extern "C" {

size_t gmc_layer2_loss_fused_workspace_bytes(int32_t n_graphs, int32_t n_classes) {
    return (size_t)(n_graphs > 0 ? n_graphs : 0) * (size_t)n_classes * sizeof(float);
}

static int layer2_loss_impl(const float* T2, int64_t ldt, const float4* parts, int32_t n_parts, int64_t part_stride, float* T2_out,
                            const int32_t* rowptr, const int32_t* colidx, const float* coef,
                            const float* vals, const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, int64_t n_rows,
                            int32_t n_classes, const float* bias2, int32_t mode, int32_t override_terminals, float penalty,
                            float C, float* Z_out, float* P_out, double* loss_per_graph, float* dZ_out, float* dT2,
                            int64_t lddt, float* db2, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace gmc;
    GMC_REQUIRE((T2 || parts) && rowptr && colidx && coef && graph_ptr && loss_per_graph, "gmc_layer2_loss_fused: null pointer");
    GMC_REQUIRE(!parts || (n_parts >= 1 && part_stride >= n_rows && n_classes <= 4 && aligned16(parts)),
                "gmc_layer2_loss_fused_parts: needs n_parts >= 1, part_stride >= n_rows, n_classes <= 4, 16-byte aligned partials");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && max_nodes >= 0 && ldt >= n_classes, "gmc_layer2_loss_fused: bad sizes");
    GMC_REQUIRE(n_classes >= 2 && n_classes <= kMaxClasses, "gmc_layer2_loss_fused: n_classes must be 2..8");
    GMC_REQUIRE(mode == GMC_LOSS_STE || mode == GMC_LOSS_SOFT, "gmc_layer2_loss_fused: bad mode %d", mode);
    GMC_REQUIRE(!override_terminals || n_classes >= 3, "gmc_layer2_loss_fused: terminal override needs >= 3 classes");
    GMC_REQUIRE(!dT2 || lddt >= n_classes, "gmc_layer2_loss_fused: lddt too small");
    const size_t smem = (size_t)3 * (size_t)max_nodes * (size_t)n_classes * sizeof(float);
    if (smem > kTailMaxSmem) {
        set_error("gmc_layer2_loss_fused: a graph of %d nodes needs %zu bytes of shared memory (limit %zu); use the separate "
                  "kernels", max_nodes, smem, kTailMaxSmem);
        return GMC_ERR_UNSUPPORTED;
    }
    float* part = nullptr;
    if (db2) {
        const size_t need = gmc_layer2_loss_fused_workspace_bytes(n_graphs, n_classes);
        if (n_graphs > 0 && (!workspace || workspace_bytes < need)) {
            set_error("gmc_layer2_loss_fused: workspace too small (%zu < %zu)", workspace_bytes, need);
            return GMC_ERR_WORKSPACE;
        }
        // one graph: its partial IS db2 (the reduce of one term adds zeros: same bits), no second launch
        part = n_graphs == 1 ? db2 : reinterpret_cast<float*>(workspace);
    }
    cudaStream_t s = as_stream(stream);
    if (n_graphs == 0 || n_rows == 0) {
        if (db2) GMC_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * n_classes, s));
        if (n_graphs > 0) GMC_CUDA(cudaMemsetAsync(loss_per_graph, 0, sizeof(double) * (size_t)n_graphs, s));
        return GMC_OK;
    }
#define GMC_CASE(K)                                                                                                       \
    case K: {                                                                                                             \
        static size_t attr = 0;                                                                                           \
        if (smem > 48 * 1024 && smem > attr) {                                                                            \
            GMC_CUDA(cudaFuncSetAttribute(layer2_loss_fused_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailMaxSmem)); \
            attr = kTailMaxSmem;                                                                                          \
        }                                                                                                                 \
        GMC_CUDA(launch_pdl(layer2_loss_fused_kernel<K>, n_graphs, kTailThreads, smem, s, T2, ldt, rowptr, colidx, coef, vals, \
                            graph_ptr, bias2, mode, override_terminals, penalty, C, Z_out, P_out, loss_per_graph, dZ_out,  \
                            dT2, lddt, part, parts, n_parts, part_stride, T2_out));                                       \
    } break;
    switch (n_classes) { GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    if (db2 && n_graphs > 1) GMC_CUDA(launch_pdl(tail_db2_reduce_kernel, 1, 256, 0, s, part, n_graphs, n_classes, db2));
    return GMC_OK;
}

int gmc_layer2_loss_fused(const float* T2, int64_t ldt, const int32_t* rowptr, const int32_t* colidx, const float* coef,
                          const float* vals, const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, int64_t n_rows,
                          int32_t n_classes, const float* bias2, int32_t mode, int32_t override_terminals, float penalty,
                          float C, float* Z_out, float* P_out, double* loss_per_graph, float* dZ_out, float* dT2,
                          int64_t lddt, float* db2, void* workspace, size_t workspace_bytes, void* stream) {
    GMC_REQUIRE(T2, "gmc_layer2_loss_fused: null pointer");
    return layer2_loss_impl(T2, ldt, nullptr, 0, 0, nullptr, rowptr, colidx, coef, vals, graph_ptr, n_graphs, max_nodes, n_rows,
                            n_classes, bias2, mode, override_terminals, penalty, C, Z_out, P_out, loss_per_graph, dZ_out, dT2,
                            lddt, db2, workspace, workspace_bytes, stream);
}

// The same with T2 given as the projection partials a gmc_gemm_bf16_split call left in ITS workspace (proj_w set,
// proj_out = NULL): parts[t * part_stride + row] = float4 partial of n-tile t, n_parts = ceil(hidden / real columns per tile).
// T2_out (nullable, leading dimension ldt) receives the reduced rows.
int gmc_layer2_loss_fused_parts(const void* parts, int32_t n_parts, int64_t part_stride, float* T2_out, int64_t ldt,
                                const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                                const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, int64_t n_rows,
                                int32_t n_classes, const float* bias2, int32_t mode, int32_t override_terminals,
                                float penalty, float C, float* Z_out, float* P_out, double* loss_per_graph, float* dZ_out,
                                float* dT2, int64_t lddt, float* db2, void* workspace, size_t workspace_bytes, void* stream) {
    GMC_REQUIRE(parts, "gmc_layer2_loss_fused_parts: null pointer");
    GMC_REQUIRE(!T2_out || ldt >= n_classes, "gmc_layer2_loss_fused_parts: ldt too small");
    return layer2_loss_impl(nullptr, T2_out ? ldt : n_classes, reinterpret_cast<const float4*>(parts), n_parts, part_stride, T2_out,
                            rowptr, colidx, coef, vals, graph_ptr, n_graphs, max_nodes, n_rows, n_classes, bias2, mode,
                            override_terminals, penalty, C, Z_out, P_out, loss_per_graph, dZ_out, dT2, lddt, db2, workspace,
                            workspace_bytes, stream);
}

}  // extern "C"
