/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's integer
 * post-processing (never linked into the product library).
 *
 * Built by oracle/Makefile (or oracle/postproc.py on demand) into
 * oracle/_build/liboracle_postproc.so and driven through ctypes from tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Graph layout: in-edge CSR of the undirected graph with BOTH directions stored
 * (dgl.from_networkx semantics, graphExtender.py:102-103), int32 indices, int32
 * edge weights (GraphCreator.py:88-90 writes weight=1; `wts == NULL` means all 1).
 *
 * All citations are relative to /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* python/Testing/TestingNeuralNetwork.py:48-64  calculate_cut_value
 * (== python/RandomAlgorithm/RandomizedMaxCut.py:48-60): sum of edge weights over
 * undirected edges whose end points carry different labels.  Each undirected
 * edge appears twice in the CSR, hence the exact /2 on an even integer. */
int64_t orc_cut_value(const int32_t* rowptr, const int32_t* colidx, const int32_t* wts,
                      const int32_t* labels, int32_t n)
{
    int64_t twice = 0;
    for (int32_t v = 0; v < n; ++v)
        for (int32_t e = rowptr[v]; e < rowptr[v + 1]; ++e)
            if (labels[v] != labels[colidx[e]]) twice += wts ? wts[e] : 1;
    return twice / 2;
}

/* python/Testing/TestingNeuralNetwork.py:18-46  assign_partitions.
 * Nodes 0,1,2 -> 0,1,2; every later node draws one uniform `r` and takes the first
 * class i with r < cumsum_i, else the last class.  The cumulative sum starts from
 * the Python int 0 and adds numpy float32 scalars.  Under numpy >= 2 (NEP 50: Python
 * scalars are "weak") the sum stays float32 and `rand_val < cumulative_prob` compares
 * in float32; under the reference's pinned numpy 1.x (legacy promotion, scalar with
 * scalar) `0 + np.float32` is float64, so the running sum is the float64 sum of the
 * float32 probabilities and the comparison is float64.  `compare_f32` selects which
 * (last bullet). */
void orc_assign_partitions(const float* P, int32_t n, int32_t K, const double* U,
                           int32_t compare_f32, int32_t* out)
{
    int32_t nt = n < 3 ? n : 3;
    for (int32_t t = 0; t < nt; ++t) out[t] = t;   /* the literal [0, 1, 2] at :31 */
    for (int32_t v = 3; v < n; ++v) {
        double r = U[v - 3];
        volatile float cum = 0.0f;
        volatile double cum64 = 0.0;
        int32_t pick = K - 1;
        for (int32_t i = 0; i < K; ++i) {
            cum = cum + P[(int64_t)v * K + i];
            cum64 = cum64 + (double)P[(int64_t)v * K + i];
            int hit = compare_f32 ? ((float)r < cum) : (r < cum64);
            if (hit) { pick = i; break; }
        }
        out[v] = pick;
    }
}

/* python/Testing/TestingNeuralNetwork.py:66-98  post_processing_optimization.
 * `iters` independent samplings, U holds iters*(n-3) uniforms in call order; keep
 * the first assignment whose cut is strictly greater than the best so far
 * (best starts at -inf, so iteration 0 is always taken when iters > 0).
 * Returns the best cut; best labels in out_labels; -1 and untouched labels if iters == 0
 * (the reference returns (None, -inf)). */
int64_t orc_sample_best_cut(const int32_t* rowptr, const int32_t* colidx, const int32_t* wts,
                            const float* P, int32_t n, int32_t K, const double* U, int32_t iters,
                            int32_t compare_f32, int32_t* out_labels, int32_t* out_best_iter)
{
    int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int64_t best = -1;
    int32_t best_it = -1;
    int32_t per = n > 3 ? n - 3 : 0;
    for (int32_t it = 0; it < iters; ++it) {
        orc_assign_partitions(P, n, K, U + (int64_t)it * per, compare_f32, cur);
        int64_t c = orc_cut_value(rowptr, colidx, wts, cur, n);
        if (best_it < 0 || c > best) {
            best = c; best_it = it;
            memcpy(out_labels, cur, sizeof(int32_t) * (size_t)n);
        }
    }
    free(cur);
    if (out_best_iter) *out_best_iter = best_it;
    return best;
}

/* North-star local search (P2 contract):
 * k-way generalisation of `greedy_maxcut`,
 * python/Other Algorithms/huerestics_multi-max.ipynb:L5817-5854 (cell 48): per
 * iteration evaluate every single-node move, take the maximum, FIRST index on ties
 * (`traversal_scores.index(best_score)`, :L5846), accept only if strictly better,
 * otherwise stop; at most `iters` iterations.  Terminals (nodes < n_frozen) are
 * skipped as in cell 8 `local_search_step` (:L99 `if node in {source, sink}: continue`).
 * Move order for tie-breaking: node ascending, then target class ascending.
 * gain(v,c) = w(v -> label(v)) - w(v -> c), w(v->c) = sum of weights of v's
 * neighbours currently labelled c.  Returns the final cut; *moves = accepted moves. */
int64_t orc_greedy_node_move(const int32_t* rowptr, const int32_t* colidx, const int32_t* wts,
                             const int32_t* labels_in, int32_t n, int32_t K, int32_t iters,
                             int32_t n_frozen, int32_t* labels_out, int32_t* moves)
{
    memcpy(labels_out, labels_in, sizeof(int32_t) * (size_t)n);
    int64_t* w = (int64_t*)malloc(sizeof(int64_t) * (size_t)(K > 0 ? K : 1));
    int32_t done = 0;
    for (int32_t it = 0; it < iters; ++it) {
        int64_t best_gain = 0; int32_t best_v = -1, best_c = -1;
        for (int32_t v = n_frozen; v < n; ++v) {
            for (int32_t c = 0; c < K; ++c) w[c] = 0;
            for (int32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
                int32_t u = colidx[e];
                if (u == v) continue;                 /* a self loop never changes the cut */
                w[labels_out[u]] += wts ? wts[e] : 1;
            }
            int32_t lv = labels_out[v];
            for (int32_t c = 0; c < K; ++c) {
                if (c == lv) continue;
                int64_t gain = w[lv] - w[c];
                if (gain > best_gain) { best_gain = gain; best_v = v; best_c = c; }
            }
        }
        if (best_v < 0) break;
        labels_out[best_v] = best_c;
        ++done;
    }
    free(w);
    if (moves) *moves = done;
    return orc_cut_value(rowptr, colidx, wts, labels_out, n);
}

/* simple_partition_assignment, python/Testing/TestingNeuralNetwork.py:100-122:
 * row-wise argmax (torch.max returns the FIRST maximal index on CPU), then nodes
 * 0,1,2 forced to 0,1,2 when n >= 3. */
void orc_simple_assignment(const float* P, int32_t n, int32_t K, int32_t* out)
{
    for (int32_t v = 0; v < n; ++v) {
        int32_t a = 0;
        for (int32_t i = 1; i < K; ++i)
            if (P[(int64_t)v * K + i] > P[(int64_t)v * K + a]) a = i;
        out[v] = a;
    }
    if (n >= 3) { out[0] = 0; out[1] = 1; out[2] = 2; }
}
