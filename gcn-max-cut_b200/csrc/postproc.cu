// postproc.cu -- (e) integer post-processing over CSR: argmax labels, exact integer cut,
// P1 = best-of-`iters` categorical sampling, P2 = greedy best-improvement node moves.
//
// Replaces the pure-Python loops of the reference's python/Testing/TestingNeuralNetwork.py
// (simple_partition_assignment :100-122, calculate_cut_value :48-64, assign_partitions :18-46,
// post_processing_optimization :66-98) and the notebook heuristic `greedy_maxcut`
// ("Other Algorithms/huerestics_multi-max.ipynb":L5817-5854).  All results are bit-exact
// against oracle/postproc.c (integer arithmetic; the only floating-point step is the float32
// running sum of the class probabilities, done in the reference's order with no contraction).
#include "common.cuh"

namespace gmc {

// ---- labels ---------------------------------------------------------------------------
__global__ void argmax_labels_kernel(const float* __restrict__ P, int64_t ldp, const int32_t* __restrict__ graph_ptr,
                                     int n_graphs, int64_t n_rows, int K, int force, int32_t* __restrict__ labels) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    const float* p = P + v * ldp;
    int a = 0;
    float best = __ldg(p);
    for (int k = 1; k < K; ++k) {
        const float x = __ldg(p + k);
        if (x > best) { best = x; a = k; }
    }
    if (force) {
        const int g = find_graph(graph_ptr, n_graphs, v);
        const int base = __ldg(graph_ptr + g);
        const int iv = (int)(v - base);
        const int n_g = __ldg(graph_ptr + g + 1) - base;
        if (n_g >= 3 && iv < 3) a = iv;                  // `if len(partition_assignment) >= 3` (:116)
    }
    labels[v] = a;
}

// ---- exact integer cut ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cut_value_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ rowptr,
                 const int32_t* __restrict__ colidx, const int32_t* __restrict__ wts,
                 const int32_t* __restrict__ graph_ptr, int n_graphs, int64_t n_rows,
                 unsigned long long* __restrict__ twice) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v < n_rows;
    long long c = 0;
    int g = -1;
    if (active) {
        g = find_graph(graph_ptr, n_graphs, v);
        const int lv = __ldg(labels + v);
        const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
        for (int e = e0; e < e1; ++e)
            if (__ldg(labels + __ldg(colidx + e)) != lv) c += wts ? __ldg(wts + e) : 1;
    }
    const int g0 = __shfl_sync(0xffffffffu, g, 0);
    const bool uniform = __all_sync(0xffffffffu, g == g0 || !active);
    if (uniform && g0 >= 0) {
        const long long s = warp_sum(c);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(twice + g0, (unsigned long long)s);
    } else if (active && c) {
        atomicAdd(twice + g, (unsigned long long)c);
    }
}

__global__ void halve_kernel(int64_t* __restrict__ x, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] / 2;
}

// ---- many labelings of ONE graph (randomized baseline, RandomizedMaxCut.py:63-122) ------------------------
// grid = T labelings; labels uint8 [T][n]; cut[t] = exact integer cut of labeling t
__global__ void __launch_bounds__(256)
cut_value_multi_kernel(const uint8_t* __restrict__ labels, const int32_t* __restrict__ rowptr,
                       const int32_t* __restrict__ colidx, const int32_t* __restrict__ wts, int n,
                       int64_t* __restrict__ cut) {
    __shared__ long long red[8];
    const uint8_t* lab = labels + (int64_t)blockIdx.x * n;
    long long c = 0;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const int lv = lab[v];
        const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
        for (int e = e0; e < e1; ++e)
            if (lab[__ldg(colidx + e)] != lv) c += wts ? __ldg(wts + e) : 1;
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        cut[blockIdx.x] = s / 2;
    }
}

// ---- P1: categorical sampling ------------------------------------------------------------
// label of local node i (global v) in iteration `it` -- assign_partitions (:18-46)
__device__ __forceinline__ int sampled_label(const float* __restrict__ P, int64_t ldp, int64_t v, int i, int K,
                                             const double* __restrict__ Ug, int64_t per_iter, int it, int cmp_f32) {
    if (i < 3) return i;
    const double r = __ldg(Ug + (int64_t)it * per_iter + (i - 3));
    if (cmp_f32) {
        // numpy >= 2 (NEP 50): `0 + np.float32` stays float32 and the Python-float uniform is "weak": float32 running
        // sum, float32 comparison
        const float rf = (float)r;
        float cum = 0.0f;
        for (int k = 0; k < K; ++k) {
            cum = __fadd_rn(cum, __ldg(P + v * ldp + k));     // no contraction
            if (rf < cum) return k;
        }
        return K - 1;
    }
    // numpy 1.x (legacy promotion): the Python int 0 plus a float32 SCALAR promotes to float64, so the running sum is the
    // float64 sum of the float32 probabilities and the comparison is float64
    double cum = 0.0;
    for (int k = 0; k < K; ++k) {
        cum = __dadd_rn(cum, (double)__ldg(P + v * ldp + k));
        if (r < cum) return k;
    }
    return K - 1;
}

// grid (iters, n_graphs): CTA computes the cut of one sampling of one graph
__global__ void __launch_bounds__(256)
sample_cut_kernel(const float* __restrict__ P, int64_t ldp, const double* __restrict__ U,
                  const int64_t* __restrict__ u_ptr, const int32_t* __restrict__ rowptr,
                  const int32_t* __restrict__ colidx, const int32_t* __restrict__ wts,
                  const int32_t* __restrict__ graph_ptr, int K, int iters, int cmp_f32, int64_t* __restrict__ cuts) {
    __shared__ long long red[8];
    const int it = blockIdx.x, g = blockIdx.y;
    const int base = __ldg(graph_ptr + g);
    const int n_g = __ldg(graph_ptr + g + 1) - base;
    const int64_t per_iter = n_g > 3 ? n_g - 3 : 0;
    const double* Ug = U + __ldg(u_ptr + g);
    long long c = 0;
    for (int i = threadIdx.x; i < n_g; i += blockDim.x) {
        const int64_t v = (int64_t)base + i;
        const int lv = sampled_label(P, ldp, v, i, K, Ug, per_iter, it, cmp_f32);
        const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
        for (int e = e0; e < e1; ++e) {
            const int u = __ldg(colidx + e);
            const int lu = sampled_label(P, ldp, u, u - base, K, Ug, per_iter, it, cmp_f32);
            if (lu != lv) c += wts ? __ldg(wts + e) : 1;
        }
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        cuts[(int64_t)g * iters + it] = s / 2;
    }
}

// grid n_graphs: first iteration with the maximal cut (strictly-greater update rule, :94-96),
// then materialise that iteration's labels
__global__ void __launch_bounds__(256)
sample_pick_kernel(const float* __restrict__ P, int64_t ldp, const double* __restrict__ U,
                   const int64_t* __restrict__ u_ptr, const int32_t* __restrict__ graph_ptr, int K, int iters,
                   int cmp_f32, const int64_t* __restrict__ cuts, int32_t* __restrict__ best_labels,
                   int64_t* __restrict__ best_cut, int32_t* __restrict__ best_iter) {
    __shared__ long long red[8];
    __shared__ int s_best;
    const int g = blockIdx.x;
    // key = cut * 2^20 + (2^20 - 1 - it): max key == max cut, then smallest iteration
    long long key = -1;
    for (int it = threadIdx.x; it < iters; it += blockDim.x) {
        const long long k = (cuts[(int64_t)g * iters + it] << 20) | (long long)(0xFFFFF - it);
        key = k > key ? k : key;
    }
    key = warp_max(key);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = -1;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = red[w] > m ? red[w] : m;
        const int it = m < 0 ? -1 : (int)(0xFFFFF - (m & 0xFFFFF));
        s_best = it;
        best_iter[g] = it;
        best_cut[g] = m < 0 ? -1 : (m >> 20);
    }
    __syncthreads();
    const int it = s_best;
    if (it < 0) return;
    const int base = __ldg(graph_ptr + g);
    const int n_g = __ldg(graph_ptr + g + 1) - base;
    const int64_t per_iter = n_g > 3 ? n_g - 3 : 0;
    const double* Ug = U + __ldg(u_ptr + g);
    for (int i = threadIdx.x; i < n_g; i += blockDim.x)
        best_labels[base + i] = sampled_label(P, ldp, (int64_t)base + i, i, K, Ug, per_iter, it, cmp_f32);
}

// ---- P2: greedy best-improvement node moves ------------------------------------------------
// one CTA per graph; labels live in labels_out (global, L1/L2 resident for n <= ~10^5)
template <int K>
__global__ void __launch_bounds__(512)
greedy_kernel(const int32_t* __restrict__ labels_in, const int32_t* __restrict__ rowptr,
              const int32_t* __restrict__ colidx, const int32_t* __restrict__ wts,
              const int32_t* __restrict__ graph_ptr, int iters, int n_frozen, int32_t* labels_out,
              int64_t* __restrict__ cut_out, int32_t* __restrict__ moves_out) {
    __shared__ long long red[16];
    __shared__ long long s_key;
    const int g = blockIdx.x;
    const int base = __ldg(graph_ptr + g);
    const int n_g = __ldg(graph_ptr + g + 1) - base;
    int32_t* lab = labels_out + base;
    for (int i = threadIdx.x; i < n_g; i += blockDim.x) lab[i] = labels_in[base + i];
    __syncthreads();

    int moves = 0;
    for (int iter = 0; iter < iters; ++iter) {
        // key = gain * 2^36 + (2^32 - 1 - i) * 2^4 + (15 - c): max gain, then lowest node, then lowest class
        long long key = 0;
        for (int i = n_frozen + threadIdx.x; i < n_g; i += blockDim.x) {
            long long w[K];
#pragma unroll
            for (int c = 0; c < K; ++c) w[c] = 0;
            const int64_t v = (int64_t)base + i;
            const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
            for (int e = e0; e < e1; ++e) {
                const int u = __ldg(colidx + e) - base;
                if (u == i) continue;
                const int lu = lab[u];
                const long long we = wts ? __ldg(wts + e) : 1;
#pragma unroll
                for (int c = 0; c < K; ++c) w[c] += (lu == c) ? we : 0;
            }
            const int lv = lab[i];
            long long wl = 0;
#pragma unroll
            for (int c = 0; c < K; ++c) wl = (lv == c) ? w[c] : wl;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const long long gain = wl - w[c];
                if (c != lv && gain > 0) {
                    const long long k = (gain << 36) | ((long long)(0xFFFFFFFFll - i) << 4) | (long long)(15 - c);
                    key = k > key ? k : key;
                }
            }
        }
        key = warp_max(key);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long m = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = red[w] > m ? red[w] : m;
            s_key = m;
            if (m > 0) {
                const int i = (int)(0xFFFFFFFFll - ((m >> 4) & 0xFFFFFFFFll));
                const int c = 15 - (int)(m & 15);
                lab[i] = c;
            }
        }
        __syncthreads();
        if (s_key <= 0) break;
        ++moves;
        __syncthreads();
    }

    // final cut of this graph
    long long c2 = 0;
    for (int i = threadIdx.x; i < n_g; i += blockDim.x) {
        const int64_t v = (int64_t)base + i;
        const int lv = lab[i];
        const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
        for (int e = e0; e < e1; ++e)
            if (lab[__ldg(colidx + e) - base] != lv) c2 += wts ? __ldg(wts + e) : 1;
    }
    c2 = warp_sum(c2);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c2;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        cut_out[g] = s / 2;
        if (moves_out) moves_out[g] = moves;
    }
}

}  // namespace gmc

// ---- numpy's legacy generator on the device -----------------------------------------------------
// assign_partitions (TestingNeuralNetwork.py:18-46) draws one np.random.rand() per node and iteration: 2.27 M doubles
// for one pass over BASELINE config 2's test graphs -- 10 ms of host time (MT19937 is sequential) plus an 18 MB upload,
// half of the pass.  The same stream on the device: MT19937 as numpy's randomkit runs it (624-word state, regenerated
// in place; a double = (a >> 5) * 2^26 + (b >> 6) over 2^53 from two consecutive tempered words), one CTA.  The
// regeneration is a three-word chain per thread (word i from the old state, i + 227 and i + 454 from the word produced
// 227 places earlier -- by the same thread) plus one last word, so a block of 312 doubles costs three barriers instead
// of 624 dependent steps.  The caller hands in np.random.get_state() and puts the returned state back: the host generator
// ends exactly where the reference's own draws would have left it.
namespace gmc {
constexpr int kMtN = 624, kMtM = 397;

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far_word) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far_word ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
__device__ __forceinline__ double mt_double(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

__global__ void __launch_bounds__(256)
mt19937_uniform_kernel(const uint32_t* __restrict__ state_in, int pos_in, int64_t n, double* __restrict__ out,
                       uint32_t* __restrict__ state_out, int32_t* __restrict__ pos_out) {
    __shared__ uint32_t mt[kMtN];
    const int tid = threadIdx.x;
    for (int i = tid; i < kMtN; i += 256) mt[i] = state_in[i];
    __syncthreads();
    // everything below is uniform over the CTA
    int p = pos_in;
    int64_t remaining = n, oi = 0;
    bool carry = false;
    uint32_t cw = 0;
    while (remaining > 0) {
        if (p == kMtN) {
            // Thread i < 227 owns words i, i + 227 and (i < 169) i + 454: new[i] needs old words only, new[i + 227] needs
            // new[i], new[i + 454] needs new[i + 227] -- a chain inside the thread once the OLD neighbours are in registers.
            uint32_t o0 = 0, n0 = 0, far_w = 0, o1 = 0, n1 = 0, o2 = 0, n2 = 0, last = 0;
            if (tid < 227) {
                o0 = mt[tid]; n0 = mt[tid + 1]; far_w = mt[tid + kMtM];
                o1 = mt[tid + 227]; n1 = mt[tid + 228];
                if (tid < 169) { o2 = mt[tid + 454]; n2 = mt[tid + 455]; }
            }
            if (tid == 0) last = mt[kMtN - 1];
            __syncthreads();                                      // every read of the old state (and of its outputs) is done
            if (tid < 227) {
                const uint32_t v0 = mt_mix(o0, n0, far_w);
                const uint32_t v1 = mt_mix(o1, n1, v0);
                mt[tid] = v0;
                mt[tid + 227] = v1;
                if (tid < 169) mt[tid + 454] = mt_mix(o2, n2, v1);
            }
            __syncthreads();
            if (tid == 0) mt[kMtN - 1] = mt_mix(last, mt[0], mt[kMtM - 1]);   // 623: new[0] and new[396]
            __syncthreads();
            p = 0;
        }
        int start = p;
        if (carry) {                                              // the first word of this block completes a double
            if (tid == 0) out[oi] = mt_double(cw, mt_temper(mt[p]));
            ++oi; --remaining; start = p + 1; carry = false;
        }
        if (remaining == 0) { p = start; break; }
        const int64_t fit = (kMtN - start) / 2;
        const int pairs = (int)(remaining < fit ? remaining : fit);
        for (int t = tid; t < pairs; t += 256)
            out[oi + t] = mt_double(mt_temper(mt[start + 2 * t]), mt_temper(mt[start + 2 * t + 1]));
        oi += pairs; remaining -= pairs;
        int used = start + 2 * pairs;
        if (remaining > 0 && used < kMtN) {                       // one word left in the block: first half of the next double
            cw = mt_temper(mt[kMtN - 1]);
            carry = true;
            used = kMtN;
        }
        p = used;                                                 // (the regeneration's first barrier comes after these reads)
    }
    __syncthreads();
    for (int i = tid; i < kMtN; i += 256) state_out[i] = mt[i];
    if (tid == 0) *pos_out = p;
}
}  // namespace gmc

extern "C" {

// n doubles of numpy's legacy np.random.rand stream, continued from `state_in` (624 words, np.random.get_state()[1]) at
// position `pos` (0..624): out[0..n) on the device, and the generator state after those draws in state_out / pos_out.
int gmc_mt19937_uniform_f64(const uint32_t* state_in, int32_t pos, int64_t n, double* out, uint32_t* state_out,
                            int32_t* pos_out, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(state_in && state_out && pos_out && (out || n == 0), "gmc_mt19937_uniform_f64: null pointer");
    GMC_REQUIRE(pos >= 0 && pos <= kMtN && n >= 0, "gmc_mt19937_uniform_f64: pos must be 0..624, n >= 0");
    mt19937_uniform_kernel<<<1, 256, 0, as_stream(stream)>>>(state_in, pos, n, out, state_out, pos_out);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_argmax_labels(const float* P, int64_t ldp, const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows,
                      int32_t n_classes, int32_t force_terminals, int32_t* labels, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(P && labels && (graph_ptr || !force_terminals), "gmc_argmax_labels: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_classes >= 1 && ldp >= n_classes, "gmc_argmax_labels: bad sizes");
    if (n_rows == 0) return GMC_OK;
    argmax_labels_kernel<<<(unsigned)ceil_div<int64_t>(n_rows, 256), 256, 0, as_stream(stream)>>>(
        P, ldp, graph_ptr, n_graphs, n_rows, n_classes, force_terminals, labels);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_cut_value_i32(const int32_t* labels, const int32_t* rowptr, const int32_t* colidx, const int32_t* wts,
                      const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int64_t* cut, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(labels && rowptr && colidx && graph_ptr && cut, "gmc_cut_value_i32: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0, "gmc_cut_value_i32: bad sizes");
    cudaStream_t s = as_stream(stream);
    if (n_graphs == 0) return GMC_OK;
    GMC_CUDA(cudaMemsetAsync(cut, 0, sizeof(int64_t) * (size_t)n_graphs, s));
    if (n_rows == 0) return GMC_OK;
    cut_value_kernel<<<(unsigned)ceil_div<int64_t>(n_rows, 256), 256, 0, s>>>(
        labels, rowptr, colidx, wts, graph_ptr, n_graphs, n_rows, reinterpret_cast<unsigned long long*>(cut));
    GMC_LAUNCH_CHECK();
    halve_kernel<<<ceil_div(n_graphs, 256), 256, 0, s>>>(cut, n_graphs);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_cut_value_multi_u8(const uint8_t* labels, const int32_t* rowptr, const int32_t* colidx, const int32_t* wts,
                           int32_t n_nodes, int32_t n_labelings, int64_t* cut, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(labels && rowptr && colidx && cut, "gmc_cut_value_multi_u8: null pointer");
    GMC_REQUIRE(n_nodes >= 0 && n_labelings >= 0, "gmc_cut_value_multi_u8: bad sizes");
    if (n_labelings == 0) return GMC_OK;
    cut_value_multi_kernel<<<n_labelings, 256, 0, as_stream(stream)>>>(labels, rowptr, colidx, wts, n_nodes, cut);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_sample_best_cut(const float* P, int64_t ldp, const double* U, const int64_t* u_ptr, const int32_t* rowptr,
                        const int32_t* colidx, const int32_t* wts, const int32_t* graph_ptr, int32_t n_graphs,
                        int64_t n_rows, int32_t n_classes, int32_t iters, int32_t compare_f32, int64_t* cuts_ws,
                        int32_t* best_labels, int64_t* best_cut, int32_t* best_iter, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(P && U && u_ptr && rowptr && colidx && graph_ptr && cuts_ws && best_labels && best_cut && best_iter,
                "gmc_sample_best_cut: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_graphs <= 65535 && n_rows >= 0 && n_classes >= 1 && ldp >= n_classes,
                "gmc_sample_best_cut: bad sizes (n_graphs <= 65535 per call)");
    GMC_REQUIRE(iters >= 1 && iters < (1 << 20), "gmc_sample_best_cut: iters must be in [1, 2^20)");
    if (n_graphs == 0) return GMC_OK;
    cudaStream_t s = as_stream(stream);
    dim3 grid((unsigned)iters, (unsigned)n_graphs);
    sample_cut_kernel<<<grid, 256, 0, s>>>(P, ldp, U, u_ptr, rowptr, colidx, wts, graph_ptr, n_classes, iters,
                                           compare_f32, cuts_ws);
    GMC_LAUNCH_CHECK();
    sample_pick_kernel<<<n_graphs, 256, 0, s>>>(P, ldp, U, u_ptr, graph_ptr, n_classes, iters, compare_f32, cuts_ws,
                                                best_labels, best_cut, best_iter);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_greedy_node_move(const int32_t* labels_in, const int32_t* rowptr, const int32_t* colidx, const int32_t* wts,
                         const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_classes, int32_t iters,
                         int32_t n_frozen, int32_t* labels_out, int64_t* cut_out, int32_t* moves_out, void* stream) {
    using namespace gmc;
    GMC_REQUIRE(labels_in && rowptr && colidx && graph_ptr && labels_out && cut_out, "gmc_greedy_node_move: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && iters >= 0 && n_frozen >= 0, "gmc_greedy_node_move: bad sizes");
    GMC_REQUIRE(n_classes >= 2 && n_classes <= kMaxClasses, "gmc_greedy_node_move: n_classes must be 2..8");
    GMC_REQUIRE(labels_in != labels_out, "gmc_greedy_node_move: in-place labels not supported");
    if (n_graphs == 0) return GMC_OK;
    cudaStream_t s = as_stream(stream);
#define GMC_CASE(K)                                                                                               \
    case K:                                                                                                       \
        greedy_kernel<K><<<n_graphs, 512, 0, s>>>(labels_in, rowptr, colidx, wts, graph_ptr, iters, n_frozen,    \
                                                  labels_out, cut_out, moves_out);                               \
        break;
    switch (n_classes) { GMC_CASE(2) GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8) }
#undef GMC_CASE
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

}  // extern "C"
