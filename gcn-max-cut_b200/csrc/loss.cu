// loss.cu -- (c) fused softmax + terminal override + argmax/STE + edge-wise max-cut loss + dL/dZ,
// one pass over the edge list, no dense n x n intermediate, no autograd graph.
//
// Replaces (reference python/Training/TrainingNeural.py): F.softmax :84, override_fixed_nodes
// :87-94, apply_max_to_one_hot :96-106 (a Python loop over rows), calculate_HC_vectorized
// :154-176 (dense [n,1000] products), compute_loss :291-309, terminal_independence_penalty
// :178-195, and the ~5n autograd nodes their backward creates.
//
// Closed form (verified against the reference's autograd in
// tests/golden/gcn_step.npz):  loss_g = -C * cut(s);  g_v = C * sum_u w_uv s_u;
// dZ_v = P_v .* (g_v - <P_v, g_v>) with the raw softmax output P (override and STE are
// identity for gradients on every row, terminals included).
//
// One thread per node.  A neighbour's hard label is recomputed from its logits (3 expf per
// neighbour -- cheaper than a second kernel and a label round trip through HBM).
// Algorithmic bytes per node: read Z 4K, CSR 4(d+1), write P and dZ 8K.
#include "common.cuh"

namespace gmc {

template <int K>
__device__ __forceinline__ void softmax_row(const float* __restrict__ z, float (&p)[K]) {
    float m = z[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = fmaxf(m, z[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { p[k] = expf(z[k] - m); s += p[k]; }
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = p[k] / s;
}

template <int K>
__device__ __forceinline__ int first_argmax(const float (&p)[K]) {
    int a = 0;
#pragma unroll
    for (int k = 1; k < K; ++k) if (p[k] > p[a]) a = k;
    return a;
}

// s-vector of node u (local index iu) under the selected mode
template <int K>
__device__ __forceinline__ void node_state(const float* __restrict__ Z, int64_t ldz, int64_t u, int iu, int mode,
                                           int override_t, float (&s)[K]) {
    if (override_t && iu < 3) {
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] = (k == iu) ? 1.f : 0.f;
        return;
    }
    float z[K], p[K];
#pragma unroll
    for (int k = 0; k < K; ++k) z[k] = __ldg(Z + u * ldz + k);
    softmax_row<K>(z, p);
    if (mode == GMC_LOSS_STE) {
        const int a = first_argmax<K>(p);
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] = (k == a) ? 1.f : 0.f;
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] = p[k];
    }
}

template <int K>
__global__ void __launch_bounds__(256)
cut_loss_kernel(const float* __restrict__ Z, int64_t ldz, const int32_t* __restrict__ rowptr,
                const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                const int32_t* __restrict__ graph_ptr, int n_graphs, int64_t n_rows, int mode, int override_t,
                float penalty, float C, float* __restrict__ P_out, double* __restrict__ loss, float* __restrict__ dZ) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v < n_rows;
    int g = -1;
    double contrib = 0.0;
    if (active) {
        g = find_graph(graph_ptr, n_graphs, v);
        const int base = __ldg(graph_ptr + g);
        const int iv = (int)(v - base);
        const int n_g = __ldg(graph_ptr + g + 1) - base;

        float z[K], p[K], sv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = __ldg(Z + v * ldz + k);
        softmax_row<K>(z, p);
        if (override_t && iv < 3) {
#pragma unroll
            for (int k = 0; k < K; ++k) sv[k] = (k == iv) ? 1.f : 0.f;
        } else if (mode == GMC_LOSS_STE) {
            const int a = first_argmax<K>(p);
#pragma unroll
            for (int k = 0; k < K; ++k) sv[k] = (k == a) ? 1.f : 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) sv[k] = p[k];
        }

        float as[K];
#pragma unroll
        for (int k = 0; k < K; ++k) as[k] = 0.f;
        float wdeg = 0.f;
        const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
        for (int e = e0; e < e1; ++e) {
            const int u = __ldg(colidx + e);
            const float w = vals ? __ldg(vals + e) : 1.0f;
            float su[K];
            node_state<K>(Z, ldz, u, u - base, mode, override_t, su);
#pragma unroll
            for (int k = 0; k < K; ++k) as[k] = fmaf(w, su[k], as[k]);
            wdeg += w;
        }
        float same = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) same = fmaf(sv[k], as[k], same);
        contrib = -(double)C * 0.5 * ((double)wdeg - (double)same);

        float gv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gv[k] = C * as[k];
        if (penalty != 0.f && iv < 3 && iv < n_g) {
            // terminal_independence_penalty on s rows 0..2: value once (thread of terminal 0),
            // gradient penalty * sum_{j != i} s_j on each terminal row
            float tot[K];
#pragma unroll
            for (int k = 0; k < K; ++k) tot[k] = 0.f;
            float pair = 0.f;
            const int nt = n_g < 3 ? n_g : 3;
            float st[3][K];
            for (int j = 0; j < nt; ++j) node_state<K>(Z, ldz, (int64_t)base + j, j, mode, override_t, st[j]);
            for (int j = 0; j < nt; ++j) {
                if (j != iv) {
#pragma unroll
                    for (int k = 0; k < K; ++k) tot[k] += st[j][k];
                }
                for (int l = j + 1; l < nt; ++l) {
#pragma unroll
                    for (int k = 0; k < K; ++k) pair = fmaf(st[j][k], st[l][k], pair);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) gv[k] = fmaf(penalty, tot[k], gv[k]);
            if (iv == 0) contrib += (double)penalty * (double)pair;
        }

        if (P_out) {
#pragma unroll
            for (int k = 0; k < K; ++k) P_out[v * K + k] = p[k];
        }
        if (dZ) {
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) dot = fmaf(p[k], gv[k], dot);
#pragma unroll
            for (int k = 0; k < K; ++k) dZ[v * K + k] = p[k] * (gv[k] - dot);
        }
    }
    // per-graph reduction: warp-uniform graph -> one double atomic per warp
    const int g0 = __shfl_sync(0xffffffffu, g, 0);
    const bool uniform = __all_sync(0xffffffffu, g == g0 || !active);
    if (uniform && g0 >= 0) {
        const double s = warp_sum(contrib);
        if ((threadIdx.x & 31) == 0) atomicAdd(loss + g0, s);
    } else if (active) {
        atomicAdd(loss + g, contrib);
    }
}

// standalone row softmax (inference / generic autograd use of GCNSoftmax.forward)
template <int K>
__global__ void softmax_fwd_kernel(const float* __restrict__ Z, int64_t ldz, int64_t n_rows, float* __restrict__ P) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    float z[K], p[K];
#pragma unroll
    for (int k = 0; k < K; ++k) z[k] = __ldg(Z + v * ldz + k);
    softmax_row<K>(z, p);
#pragma unroll
    for (int k = 0; k < K; ++k) P[v * K + k] = p[k];
}

// dZ = P .* (dP - <P, dP>)
template <int K>
__global__ void softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP, int64_t n_rows,
                                   float* __restrict__ dZ) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    float p[K], g[K], dot = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { p[k] = __ldg(P + v * K + k); g[k] = __ldg(dP + v * K + k); dot = fmaf(p[k], g[k], dot); }
#pragma unroll
    for (int k = 0; k < K; ++k) dZ[v * K + k] = p[k] * (g[k] - dot);
}

}  // namespace gmc

extern "C" int gmc_softmax_fwd_f32(const float* Z, int64_t ldz, int64_t n_rows, int32_t n_classes, float* P,
                                   void* stream) {
    using namespace gmc;
    GMC_REQUIRE(Z && P, "gmc_softmax_fwd_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_classes >= 1 && n_classes <= kMaxClasses && ldz >= n_classes,
                "gmc_softmax_fwd_f32: bad sizes (n_classes 1..8)");
    if (n_rows == 0) return GMC_OK;
    const unsigned blocks = (unsigned)ceil_div<int64_t>(n_rows, 256);
    cudaStream_t s = as_stream(stream);
#define GMC_CASE(K) case K: softmax_fwd_kernel<K><<<blocks, 256, 0, s>>>(Z, ldz, n_rows, P); break;
    switch (n_classes) { GMC_CASE(1) GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

extern "C" int gmc_softmax_bwd_f32(const float* P, const float* dP, int64_t n_rows, int32_t n_classes, float* dZ,
                                   void* stream) {
    using namespace gmc;
    GMC_REQUIRE(P && dP && dZ, "gmc_softmax_bwd_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_classes >= 1 && n_classes <= kMaxClasses, "gmc_softmax_bwd_f32: bad sizes");
    if (n_rows == 0) return GMC_OK;
    const unsigned blocks = (unsigned)ceil_div<int64_t>(n_rows, 256);
    cudaStream_t s = as_stream(stream);
#define GMC_CASE(K) case K: softmax_bwd_kernel<K><<<blocks, 256, 0, s>>>(P, dP, n_rows, dZ); break;
    switch (n_classes) { GMC_CASE(1) GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// dX[r, c] = Y[r, c] > 0 ? dY[r, c] : 0 -- backward of the ReLU between the two GraphConv layers (TrainingNeural.py:81)
// for the generic autograd path; leading dimensions in elements, one thread per 4 columns when everything is 16-byte aligned.
namespace gmc {
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ Y, int64_t ldy, float* __restrict__ dX,
                int64_t lddx, int64_t n_rows, int n_cols, int vec) {
    const int per = vec ? (n_cols + 3) / 4 : n_cols;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * per) return;
    const int64_t r = i / per;
    const int c = (int)(i - r * per);
    if (vec) {
        const float4 g = *reinterpret_cast<const float4*>(dY + r * lddy + 4 * c);
        const float4 y = *reinterpret_cast<const float4*>(Y + r * ldy + 4 * c);
        *reinterpret_cast<float4*>(dX + r * lddx + 4 * c) =
            make_float4(y.x > 0.f ? g.x : 0.f, y.y > 0.f ? g.y : 0.f, y.z > 0.f ? g.z : 0.f, y.w > 0.f ? g.w : 0.f);
    } else {
        dX[r * lddx + c] = Y[r * ldy + c] > 0.f ? dY[r * lddy + c] : 0.f;
    }
}
}  // namespace gmc

extern "C" int gmc_relu_bwd_f32(const float* dY, int64_t lddy, const float* Y, int64_t ldy, float* dX, int64_t lddx,
                                int64_t n_rows, int32_t n_cols, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(dY && Y && dX, "gmc_relu_bwd_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols > 0 && lddy >= n_cols && ldy >= n_cols && lddx >= n_cols, "gmc_relu_bwd_f32: bad sizes");
    if (n_rows == 0) return GMC_OK;
    const int vec = (n_cols % 4 == 0 && lddy % 4 == 0 && ldy % 4 == 0 && lddx % 4 == 0 && aligned16(dY) && aligned16(Y) &&
                     aligned16(dX)) ? 1 : 0;
    const int64_t total = n_rows * (vec ? n_cols / 4 : n_cols);
    relu_bwd_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(dY, lddy, Y, ldy, dX, lddx, n_rows,
                                                                                            n_cols, vec);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

extern "C" int gmc_softmax_cut_loss_fwd_bwd(const float* Z, int64_t ldz, const int32_t* rowptr, const int32_t* colidx,
                                            const float* vals, const int32_t* graph_ptr, int32_t n_graphs,
                                            int64_t n_rows, int32_t n_classes, int32_t mode,
                                            int32_t override_terminals, float penalty, float C, float* P_out,
                                            double* loss_per_graph, float* dZ_out, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(Z && rowptr && colidx && graph_ptr && loss_per_graph, "gmc_softmax_cut_loss_fwd_bwd: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && ldz >= n_classes, "gmc_softmax_cut_loss_fwd_bwd: bad sizes");
    GMC_REQUIRE(n_classes >= 2 && n_classes <= kMaxClasses, "gmc_softmax_cut_loss_fwd_bwd: n_classes must be 2..8");
    GMC_REQUIRE(mode == GMC_LOSS_STE || mode == GMC_LOSS_SOFT, "gmc_softmax_cut_loss_fwd_bwd: bad mode %d", mode);
    GMC_REQUIRE(!override_terminals || n_classes >= 3,
                "gmc_softmax_cut_loss_fwd_bwd: terminal override needs >= 3 classes (reference hard-codes 3, "
                "TrainingNeural.py:91-93)");
    cudaStream_t s = as_stream(stream);
    if (n_graphs > 0) GMC_CUDA(cudaMemsetAsync(loss_per_graph, 0, sizeof(double) * (size_t)n_graphs, s));
    if (n_rows == 0) return GMC_OK;
    // P_out / dZ_out are dense [n_rows, n_classes]
    const unsigned blocks = (unsigned)ceil_div<int64_t>(n_rows, 256);
#define GMC_CASE(K)                                                                                              \
    case K:                                                                                                      \
        cut_loss_kernel<K><<<blocks, 256, 0, s>>>(Z, ldz, rowptr, colidx, vals, graph_ptr, n_graphs, n_rows,    \
                                                  mode, override_terminals, penalty, C, P_out, loss_per_graph,  \
                                                  dZ_out);                                                       \
        break;
    switch (n_classes) { GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}
