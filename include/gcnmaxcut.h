/* gcnmaxcut.h -- C ABI of libgcnmaxcut.so: the B200 (sm_100a) hot path of GCN-max-cut.
 *
 * The reference (MJavaadAkhtar/GCN-max-cut) is pure Python with NO FFI of its own;
 * every native instruction it executes lives in third-party wheels (dgl 2.0.0,
 * torch).  Each entry point below therefore names the reference Python call site
 * (file:line relative to the reference root) whose native work it replaces -- that
 * call site is "the interface it binds".  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers + sizes; all pointers are DEVICE pointers unless a parameter
 *     is documented as host; `stream` is a cudaStream_t passed as void*.
 *   - return 0 on success, a negative GMC_ERR_* for argument errors, or a positive
 *     cudaError_t; gmc_last_error() returns a thread-local description.
 *   - no allocation and no host synchronisation inside (callers pass workspaces),
 *     so every call is CUDA-graph capturable.
 *   - graphs: in-edge CSR of an undirected graph with BOTH directions stored
 *     (dgl.from_networkx semantics, graphExtender.py:102-103), int32 indices;
 *     a batch is block-diagonal: graph g owns nodes [graph_ptr[g], graph_ptr[g+1]).
 *   - matrices are row-major fp32 with an explicit leading dimension (elements).
 */
#ifndef GCNMAXCUT_H_
#define GCNMAXCUT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMC_ABI_VERSION 1

#define GMC_OK 0
#define GMC_ERR_INVALID_ARG (-1)
#define GMC_ERR_UNSUPPORTED (-2)
#define GMC_ERR_WORKSPACE (-3)

/* GEMM arithmetic */
#define GMC_GEMM_FP32 0      /* CUDA-core FFMA, exact fp32 (parity path)                       */
#define GMC_GEMM_TF32 1      /* tcgen05 kind::tf32, one pass, fp32 accumulate in TMEM          */
#define GMC_GEMM_TF32X3 2    /* tcgen05, 3 split passes (hi*hi + hi*lo + lo*hi): fp32-grade    */
/* With either tensor-core precision, problems with a dimension < 4 (the 3-class layer) run on the exact
 * CUDA-core kernel; other operands TMA cannot address (unaligned base / ld % 4 != 0) return an error. */

/* loss modes */
#define GMC_LOSS_STE 0       /* hard one-hot + straight-through (reference live path)          */
#define GMC_LOSS_SOFT 1      /* soft objective sum w_ij p_i.p_j (north-star variant)           */

int gmc_abi_version(void);
const char* gmc_last_error(void);
/* sm count / compute capability of the current device (host out-pointers). */
int gmc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- graph preparation ------------------------------------------------------------ */

/* norm[v] = max(deg(v),1)^-1/2 ; *zero_degree_count (device int32, nullable) receives the
 * number of rows with degree 0 (DGL raises on those: GraphConv allow_zero_in_degree=False).
 * Replaces: dgl GraphConv degree normalisation, TrainingNeural.py:80,83. */
int gmc_degree_norm_f32(const int32_t* rowptr, int64_t n_rows, float* norm,
                        int32_t* zero_degree_count, void* stream);

/* coef[e] = (vals ? vals[e] : 1) * norm_src[col[e]] * norm_dst[row(e)]  -- the values of
 * A_hat = D^-1/2 A D^-1/2, computed once per (static) batch. */
int gmc_edge_coef_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const float* norm_src, const float* norm_dst, int64_t n_rows,
                      float* coef, void* stream);

/* Dense padded adjacency rows X[v, u - graph_ptr[g]] = w_uv (zero elsewhere), X is
 * [n_rows, n_cols] with leading dimension ldx.  Device-side replacement for
 * commons.py:38-77 (gen_adj_matrix + qubo_dict_to_torch) + graphExtender.py:28-48. */
int gmc_csr_densify_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                        const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows,
                        int32_t n_cols, float* X, int64_t ldx, void* stream);

/* Strided device-to-device copy of an [n_rows, n_cols] fp32 matrix (cudaMemcpy2DAsync).  The engine keeps
 * its operands with leading dimensions padded to 128 bytes (TMA boxes and 128-bit gathers then never
 * straddle cache lines: measured +45 % on the tcgen05 GEMM) and uses this to refresh the padded copy of W1. */
int gmc_copy2d_f32(float* dst, int64_t lddst, const float* src, int64_t ldsrc, int64_t n_rows,
                   int32_t n_cols, void* stream);

/* ---- (a) symmetric-normalised CSR SpMM ---------------------------------------------- */

/* Y[v,:] = act( nd[v] * sum_{e in row v} w_e * ns[col_e] * X[col_e,:] + bias )
 * vals / norm_src / norm_dst / bias may be NULL (treated as 1 / 1 / 1 / 0); relu != 0
 * applies max(.,0).  Passing the precomputed gmc_edge_coef_f32 array as `vals` with both
 * norms NULL is the fast path.  A_hat is symmetric, so the backward SpMM is this call.
 * Replaces: dgl update_all(copy_u,sum) + norms + bias inside GraphConv.forward
 * (TrainingNeural.py:80 / :83) and F.relu (:81). */
int gmc_spmm_symnorm_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                         const float* norm_src, const float* norm_dst, const float* X, float* Y,
                         int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                         const float* bias, int32_t relu, void* stream);

/* Same product for a BLOCK-DIAGONAL batch whose A_hat values were precomputed (`coef` from
 * gmc_edge_coef_f32, coef_e = norm_dst[v] * norm_src[u] as in GraphConv norm='both').  With an ELL `plan`
 * of the batch (8 uint16 local neighbour ids + one coefficient per row, built once by
 * gmc_spmm_plan_build; needs max degree <= 8, graphs of < 65535 nodes and, per row, neighbours that
 * share one norm_src value -- regular graphs) and graphs that fit the on-chip slab buffer
 * ((max_nodes + 2) x 112 B, two CTAs per SM), the source rows of one (graph, 28-column slab) are staged
 * once into shared memory by TMA, gathered from there with conflict-free LDS.128 and scaled by the row
 * coefficient on the way out -- L2 traffic drops from (d+1)x to ~1x the matrix.  Otherwise (plan NULL,
 * large graphs, narrow matrices) it runs the warp-per-row kernel.  Both paths compute the same product;
 * they differ only in rounding order (sum-then-scale vs fused per-edge coef, 1e-7 relative). */
size_t gmc_spmm_plan_bytes(int64_t n_rows);
int gmc_spmm_plan_build(const int32_t* rowptr, const int32_t* colidx, const float* norm_src,
                        const float* norm_dst, const int32_t* graph_ptr, int32_t n_graphs,
                        int64_t n_rows, void* plan,
                        int32_t* overflow /* device, set to 1 if the plan cannot be used */, void* stream);
int gmc_spmm_batched_f32(const int32_t* rowptr, const int32_t* colidx, const float* coef,
                         const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes,
                         const void* plan, const float* X, float* Y, int64_t n_rows, int32_t n_cols,
                         int64_t ldx, int64_t ldy, const float* bias, int32_t relu, void* stream);

/* gmc_spmm_symnorm_f32 with the skinny second-layer projection fused into its epilogue:
 *   Y[v,:] = act(A_hat X + bias)[v,:]  and  T[v,k] = sum_j Y[v,j] W[j,k]   (W [n_cols, n_out], n_out <= 8)
 * One warp owns a whole row, so n_cols must be <= 512 (GMC_ERR_UNSUPPORTED otherwise).
 * Replaces GraphConv aggregation + bias (TrainingNeural.py:80), F.relu (:81) and conv2's th.matmul (:83)
 * in ONE pass: the separate read of H by gmc_skinny_fwd_f32 disappears. */
int gmc_spmm_fused_skinny_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                              const float* norm_src, const float* norm_dst, const float* X, float* Y,
                              int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                              const float* bias, int32_t relu, const float* W, int32_t n_out, float* T,
                              int64_t ldt, void* stream);

/* ---- (b) dense feature transforms --------------------------------------------------- */

/* C[M,N] = op(A) * op(B) (+ C if accumulate != 0), row-major.
 *   nn: A[M,K] lda, B[K,N] ldb      (X * W1,            TrainingNeural.py:80 th.matmul in GraphConv)
 *   nt: A[M,K] lda, B[N,K] ldb      (dT1 * W1^T,        autograd of :80, only for trainable features)
 *   tn: A[K,M] lda, B[K,N] ldb      (X^T * dT1 = dW1,   autograd of :80; K = all nodes -> split-K)
 * `workspace` is required when the implementation splits K (tn); query the size first.
 * precision: GMC_GEMM_*.  TF32 paths require sm_100a and 16-byte aligned rows. */
size_t gmc_gemm_workspace_bytes(int32_t op /*0 nn,1 nt,2 tn*/, int64_t M, int64_t N, int64_t K,
                                int32_t precision);
int gmc_gemm_nn(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldb, int64_t ldc, int32_t accumulate, int32_t precision,
                void* workspace, size_t workspace_bytes, void* stream);
int gmc_gemm_nt(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldb, int64_t ldc, int32_t accumulate, int32_t precision,
                void* workspace, size_t workspace_bytes, void* stream);
int gmc_gemm_tn(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldb, int64_t ldc, int32_t accumulate, int32_t precision,
                void* workspace, size_t workspace_bytes, void* stream);

/* Skinny second layer, forward: T[v,k] = sum_j H[v,j] * W[j,k], K_out <= 8.
 * Replaces th.matmul inside conv2 (TrainingNeural.py:83). */
int gmc_skinny_fwd_f32(const float* H, int64_t ldh, const float* W /*[n_in,n_out]*/, float* T,
                       int64_t ldt, int64_t n_rows, int32_t n_in, int32_t n_out, void* stream);

/* Skinny second layer, backward, one pass over H:
 *   dHpre[v,j] = (H[v,j] > 0) * sum_k dT[v,k] W[j,k]     (ReLU mask = sign of the output)
 *   dW[j,k]    = sum_v H[v,j] dT[v,k]
 *   dbias[j]   = sum_v dHpre[v,j]                        (gradient of the FIRST layer's bias)
 * Deterministic two-stage reduction through `workspace`
 * (gmc_skinny_bwd_workspace_bytes).  Replaces autograd of TrainingNeural.py:81-83. */
size_t gmc_skinny_bwd_workspace_bytes(int32_t n_in, int32_t n_out);
int gmc_skinny_bwd_f32(const float* dT, int64_t lddt, const float* W, const float* H, int64_t ldh,
                       float* dHpre, int64_t lddh, float* dW, float* dbias, int64_t n_rows,
                       int32_t n_in, int32_t n_out, void* workspace, size_t workspace_bytes,
                       void* stream);

/* out[j] = sum_v X[v,j]   (bias gradients); deterministic; workspace >= gmc_colsum_workspace_bytes. */
size_t gmc_colsum_workspace_bytes(int32_t n_cols);
int gmc_colsum_f32(const float* X, int64_t ldx, int64_t n_rows, int32_t n_cols, float* out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- (c) fused softmax + terminal override + max-cut loss + gradient ----------------- */

/* One pass over the edge list.  Per node v of graph g (local index i = v - graph_ptr[g]):
 *   P_v = softmax(Z_v);   s_v = onehot(argmax P_v) [STE] or P_v [SOFT];
 *   if override_terminals and i < 3: s_v = e_i             (override_fixed_nodes)
 *   loss_g = -C * 1/2 sum_{(u,v)} w_uv (1 - s_u . s_v)  (+ penalty * sum_{i<j<3} s_i . s_j)
 *   dZ_v   = P_v * (g_v - <P_v, g_v>),  g_v = C * sum_u w_uv s_u (+ penalty term on terminals)
 * P_out / dZ_out may be NULL.  loss_per_graph is float64 [n_graphs] (overwritten).
 * n_classes <= 8; override needs n_classes >= 3.
 * Replaces: F.softmax (TrainingNeural.py:84), override_fixed_nodes (:87-94),
 * apply_max_to_one_hot (:96-106), calculate_HC_vectorized/compute_loss (:154-176, :291-309),
 * terminal_independence_penalty (:178-195) and their autograd. */
int gmc_softmax_cut_loss_fwd_bwd(const float* Z, int64_t ldz, const int32_t* rowptr,
                                 const int32_t* colidx, const float* vals,
                                 const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows,
                                 int32_t n_classes, int32_t mode, int32_t override_terminals,
                                 float penalty, float C, float* P_out, double* loss_per_graph,
                                 float* dZ_out, void* stream);

/* Stand-alone row softmax P = softmax(Z) (dense [n_rows, n_classes] output) and its backward
 * dZ = P .* (dP - <P,dP>); used by GCNSoftmax.forward outside the fused training step
 * (inference: TestingNeuralNetwork.py:142-143; F.softmax TrainingNeural.py:84). */
int gmc_softmax_fwd_f32(const float* Z, int64_t ldz, int64_t n_rows, int32_t n_classes, float* P,
                        void* stream);
int gmc_softmax_bwd_f32(const float* P, const float* dP, int64_t n_rows, int32_t n_classes, float* dZ,
                        void* stream);
/* dX = (Y > 0) ? dY : 0 -- backward of F.relu (TrainingNeural.py:81) on the generic autograd path; leading dimensions in
 * elements. */
int gmc_relu_bwd_f32(const float* dY, int64_t lddy, const float* Y, int64_t ldy, float* dX, int64_t lddx,
                     int64_t n_rows, int32_t n_cols, void* stream);

/* (c') The whole second layer + loss + its backward for a block-diagonal batch of small graphs, ONE CTA per graph:
 *   Z = A_hat T2 + b2 (TrainingNeural.py:83), softmax / terminal override / STE / max-cut loss and dZ as
 *   gmc_softmax_cut_loss_fwd_bwd, db2 = colsum(dZ), dT2 = A_hat dZ -- the work of two gmc_spmm_* launches, the loss kernel
 *   and gmc_colsum_f32 with every node's label computed once and all [n, n_classes] intermediates in shared memory.
 * `coef` = gmc_edge_coef_f32 values (A_hat), `vals` = loss edge weights (nullable = 1).  Z_out / P_out / dZ_out (dense
 * [n_rows, n_classes]), dT2 ([n_rows, lddt]) and db2 ([n_classes]) are each nullable (inference passes dT2 = db2 = NULL).
 * Per-graph losses and db2 are sums in a fixed order: bitwise reproducible, no atomics.  Needs 3 * max_nodes * n_classes
 * floats of shared memory (max_nodes <= 6000 at 3 classes): GMC_ERR_UNSUPPORTED beyond that -- use the separate kernels.
 * workspace >= gmc_layer2_loss_fused_workspace_bytes when db2 is requested. */
size_t gmc_layer2_loss_fused_workspace_bytes(int32_t n_graphs, int32_t n_classes);
int gmc_layer2_loss_fused(const float* T2, int64_t ldt, const int32_t* rowptr, const int32_t* colidx, const float* coef,
                          const float* vals, const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes,
                          int64_t n_rows, int32_t n_classes, const float* bias2, int32_t mode,
                          int32_t override_terminals, float penalty, float C, float* Z_out, float* P_out,
                          double* loss_per_graph, float* dZ_out, float* dT2, int64_t lddt, float* db2, void* workspace,
                          size_t workspace_bytes, void* stream);
/* The same with T2 given as the per-n-tile projection partials that gmc_gemm_bf16_split leaves in its workspace when it is
 * called with proj_w and proj_out = NULL: parts[t * part_stride + row] (float4), n_parts = ceil(hidden / real columns per
 * tile: 128 for two parts, 64 for three).  Saves the reduce launch and T2's round trip through memory on the per-graph
 * training path (TrainingNeural.py:371-388: ~9 launches of a few microseconds per step).  T2_out (nullable) receives the
 * reduced rows. */
int gmc_layer2_loss_fused_parts(const void* parts, int32_t n_parts, int64_t part_stride, float* T2_out, int64_t ldt,
                                const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                                const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, int64_t n_rows,
                                int32_t n_classes, const float* bias2, int32_t mode, int32_t override_terminals,
                                float penalty, float C, float* Z_out, float* P_out, double* loss_per_graph, float* dZ_out,
                                float* dT2, int64_t lddt, float* db2, void* workspace, size_t workspace_bytes, void* stream);

/* ---- (d) fused multi-tensor Adam ----------------------------------------------------- */

/* torch.optim.Adam defaults (amsgrad=False, weight_decay=0): for each of n_tensors (<= 16)
 *   m += (g-m)(1-b1); v = v*b2 + (1-b2) g*g; p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * The four pointer arrays and `sizes` are HOST arrays of length n_tensors (copied by value
 * into the launch).  `step` is the 1-based step count t.  Replaces optimizer.step(),
 * TrainingNeural.py:386 (optimizer built at :336-337). */
int gmc_adam_multi(int32_t n_tensors, float* const* params, const float* const* grads,
                   float* const* exp_avg, float* const* exp_avg_sq, const int64_t* sizes,
                   double lr, double beta1, double beta2, double eps, int64_t step, void* stream);

/* Same, but t is read from (and incremented in) device memory so that a captured CUDA graph
 * can be replayed: *step_dev is the number of steps already taken. */
int gmc_adam_multi_devstep(int32_t n_tensors, float* const* params, const float* const* grads,
                           float* const* exp_avg, float* const* exp_avg_sq, const int64_t* sizes,
                           double lr, double beta1, double beta2, double eps, int64_t* step_dev,
                           void* stream);

/* gmc_adam_multi that also writes a bf16 copy of each updated parameter to shadows[t] (HOST array of device pointers,
 * entries nullable; element i of the copy = bf16(element i of the parameter)): learned node embeddings feed the next
 * step's bf16 GEMMs without a conversion pass (python/utils.py:184 `inputs = embed.weight`). */
int gmc_adam_multi_shadow(int32_t n_tensors, float* const* params, const float* const* grads,
                          float* const* exp_avg, float* const* exp_avg_sq, void* const* shadows, const int64_t* sizes,
                          double lr, double beta1, double beta2, double eps, int64_t step, void* stream);

/* The same with an 8-word device state (zero-initialised; word 0 = steps already taken): the launch advances the count
 * itself and leaves the bias-correction scalars of the NEXT step in the state, so a replayed per-graph training step
 * (TrainingNeural.py:371-388) carries neither a one-thread increment launch nor a double-precision pow() in front of
 * every CTA. */
int gmc_adam_multi_devstate(int32_t n_tensors, float* const* params, const float* const* grads,
                            float* const* exp_avg, float* const* exp_avg_sq, const int64_t* sizes,
                            double lr, double beta1, double beta2, double eps, int64_t* state, void* stream);

/* gmc_spmm_fused_skinny_f32 for a block-diagonal batch with an ELL plan: the TMA-staged slab kernel with the
 * skinny projection T = Y W (n_out <= 4) folded into its epilogue; per-slab partial sums go to `workspace`
 * (gmc_spmm_batched_fused_workspace_bytes) and are reduced in slab order, so T is deterministic.
 * GMC_ERR_UNSUPPORTED when the batch cannot take the slab kernel: use gmc_spmm_fused_skinny_f32. */
size_t gmc_spmm_batched_fused_workspace_bytes(int64_t n_rows, int32_t n_cols);
int gmc_spmm_batched_fused_skinny_f32(const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes,
                                      const void* plan, const float* X, float* Y, int64_t n_rows, int32_t n_cols,
                                      int64_t ldx, int64_t ldy, const float* bias, int32_t relu, const float* W,
                                      int32_t n_out, float* T, int64_t ldt, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* gmc_spmm_batched_f32 with the output rounded to bf16 on the way out (Y: bf16 matrix, ldy in elements): feeds
 * the B operand of the bf16 weight-gradient GEMM without an fp32 round trip.  Slab kernel only:
 * GMC_ERR_UNSUPPORTED when the batch has no usable plan. */
int gmc_spmm_batched_bf16out(const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, const void* plan,
                             const float* X, void* Y, int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                             void* stream);

/* bf16 operands (the north-star's "TF32/bf16 with fp32 accumulation"): A and B point at bf16 data in device
 * memory (leading dimensions in ELEMENTS, multiples of 8; 16-byte aligned bases), C is fp32.  Same tcgen05 / TMA /
 * cluster-multicast kernel as gmc_gemm_* with 64-element stages and kind::f16 MMAs: half the operand bytes per
 * flop, twice the MMA rate.  op: 0 nn, 1 nt, 2 tn.  gmc_f32_to_bf16 (round to nearest even) and
 * gmc_csr_densify_bf16 (0/1 adjacency rows are exact in bf16) produce the operands. */
size_t gmc_gemm_bf16_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K);
int gmc_gemm_bf16(int32_t op, const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K,
                  int64_t lda, int64_t ldb, int64_t ldc, int32_t accumulate, void* workspace,
                  size_t workspace_bytes, void* stream);
int gmc_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t n_rows, int32_t n_cols,
                    void* stream);
int gmc_csr_densify_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                         const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X,
                         int64_t ldx, void* stream);
/* Incremental gmc_csr_densify_bf16 for a reused feature buffer: clear = 0 writes the batch's adjacency entries into an
 * X that is zero everywhere else (no memset); clear = 1 writes zeros at the same positions (back to all-zero).
 * graphExtender.py:106-111 for a stream of graphs: nnz sectors per step instead of the whole matrix. */
int gmc_csr_scatter_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                         const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X,
                         int64_t ldx, int32_t clear, void* stream);
/* Pre-aggregated first-layer features XA = A_hat X (bf16) for X = the zero-padded adjacency rows of the batch and A_hat
 * given by `coef` (gmc_edge_coef_f32).  A_hat (X W1) = (A_hat X) W1 and X depends on the graph alone, so GraphConv
 * layer 1's aggregation (TrainingNeural.py:80) moves from the activations (every step, forward and backward) to the
 * features (once per graph): H1 = relu(XA W1 + b1) is then one GEMM with a bias/ReLU epilogue and
 * dW1 = XA^T dH1pre one GEMM.  ldx % 8 == 0, n_cols <= 6144; rows are written whole; bitwise reproducible.
 * `workspace` (nullable, gmc_csr_preaggregate_workspace_bytes = one byte per row) enables the counting kernel for rows
 * with unit weights and <= 8 neighbours of one degree and coefficient (every graph the reference generates); the
 * other rows, or all rows without a workspace, take the general kernel. */
size_t gmc_csr_preaggregate_workspace_bytes(int64_t n_rows);
int gmc_csr_preaggregate_bf16(const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                              const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X,
                              int64_t ldx, void* workspace, size_t workspace_bytes, void* stream);

/* ---- bf16 layer-1 activations (engine option activations='bf16', on top of the bf16 GEMM operands) ----------------
 * T1 = X W1, H1 = relu(A_hat T1 + b1), dH1pre and dT1 are [n_nodes, hidden] matrices that are each written once and
 * read once (or d times by a gather) per step; stored in bf16 they move half the bytes through HBM, L2 and shared
 * memory.  All arithmetic stays fp32 (TMEM accumulators, fp32 aggregation / bias / ReLU / projection / reductions);
 * values are rounded to nearest even when stored.  Leading dimensions are in ELEMENTS and multiples of 8 that cover
 * n_cols rounded up to 8; pad columns are written as zeros / must be finite on input.
 *
 * gmc_gemm_bf16_bf16out       : gmc_gemm_bf16 with a bf16 C (epilogue rounds the accumulators, TMA stores); no split-K,
 *                               no accumulate -- th.matmul of GraphConv layer 1 (TrainingNeural.py:80).  Optional fp32
 *                               bias[N] (N % 4 == 0) and ReLU in the epilogue (the whole first layer when the features
 *                               are pre-aggregated, gmc_csr_preaggregate_bf16).  Optional fused projection
 *                               P = bf16(C) projW (projW [N rounded up to 64][4] fp32 zero padded, n_proj <= 4,
 *                               N <= 512; P is zeroed by the call): T2 = H1 W2 (TrainingNeural.py:83) without a second
 *                               pass over H1.
 * gmc_spmm_fused_skinny_bf16  : gmc_spmm_fused_skinny_f32 with a bf16 X and an fp32 (y_bf16 = 0) or bf16 (y_bf16 = 1) Y;
 *                               with a bf16 Y the projection T = Y W uses the rounded Y (what the backward pass reads).
 * gmc_skinny_bwd_bf16         : gmc_skinny_bwd_f32 with bf16 H and dHpre (ldh, lddh multiples of 4).
 * gmc_spmm_batched_bf16       : gmc_spmm_batched_f32 (slab kernel only, plan required) with bf16 X and Y; fp32 bias.
 * gmc_skinny_fwd_bf16         : gmc_skinny_fwd_f32 with a bf16 H (n_out <= 4, n_in <= 512): the projection T = H1 W2
 *                               when the forward aggregation runs through the slab kernel. */
int gmc_gemm_bf16_bf16out(int32_t op, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K,
                          int64_t lda, int64_t ldb, int64_t ldc, const float* bias, int32_t relu, const float* proj_w,
                          float* proj_out, int64_t ldp, int32_t n_proj, void* stream);
int gmc_spmm_fused_skinny_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                               const float* norm_src, const float* norm_dst, const void* X, void* Y, int32_t y_bf16,
                               int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy, const float* bias,
                               int32_t relu, const float* W, int32_t n_out, float* T, int64_t ldt, void* stream);
int gmc_skinny_bwd_bf16(const float* dT, int64_t lddt, const float* W, const void* H, int64_t ldh, void* dHpre,
                        int64_t lddh, float* dW, float* dbias, int64_t n_rows, int32_t n_in, int32_t n_out,
                        void* workspace, size_t workspace_bytes, void* stream);
int gmc_spmm_batched_bf16(const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, const void* plan,
                          const void* X, void* Y, int64_t n_rows, int32_t n_cols, int64_t ldx, int64_t ldy,
                          const float* bias, int32_t relu, void* stream);
int gmc_skinny_fwd_bf16(const void* H, int64_t ldh, const float* W, float* T, int64_t ldt, int64_t n_rows,
                        int32_t n_in, int32_t n_out, void* stream);


/* ---- (b') layer-1 feature transform when the features ARE the zero-padded adjacency rows ----------------
 * The reference feeds `adjacency_matrix` [n, 1000] as input features (TrainingNeural.py:373; built by
 * DataGenerator/graphExtender.py:106-111).  For unit edge weights X W1 is a row gather of W1 and X^T dT1 a
 * per-local-node aggregation of dT1, so both dense GEMMs (and the dense X itself) can be skipped:
 *   fwd: T[v,:]  = sum_{u in N(v)} W[local(u),:]                      == gmc_gemm_nn(X, W)
 *   bwd: dW[j,:] = sum_graphs sum_{v in N_g(j)} dT[v,:], 0 for j >= n  == gmc_gemm_tn(X, dT)
 * `plan` is the batch's ELL plan (gmc_spmm_plan_build; only its neighbour ids are used).  Both return
 * GMC_ERR_UNSUPPORTED for shapes outside their reach (n_cols % 4, max_nodes > n_w_rows, n_w_rows > 1030,
 * max_nodes > 1024): the caller then keeps the dense tensor-core path.  bwd is deterministic. */
int gmc_adj_features_fwd_f32(const void* plan, const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes,
                             const float* W, int64_t ldw, int32_t n_w_rows, float* T, int64_t ldt,
                             int64_t n_rows, int32_t n_cols, void* stream);
size_t gmc_adj_features_bwd_workspace_bytes(int32_t n_w_rows, int32_t n_cols);
int gmc_adj_features_bwd_f32(const void* plan, const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes,
                             const float* dT, int64_t lddt, int64_t n_rows, int32_t n_cols, float* dW,
                             int64_t lddw, int32_t n_w_rows, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- (b'') fp32-grade layer-1 transforms at bf16 tensor-core rates (engine precision 'bf16x3') ----------------
 * The reference computes both GraphConv layers in fp32 (TORCH_DTYPE, TrainingNeural.py:33-34; th.matmul of layer 1 at
 * :80 through dgl GraphConv).  For a batch whose rows have ONE normalisation coefficient each (every regular graph) the
 * pre-aggregated features factor as (A_hat X)[v,:] = s_v * XI[v,:] with XI = sum over neighbours of the 0/1 adjacency
 * rows: small integers, exact in bf16 (gmc_csr_preaggregate_bf16 with unit coefficients builds XI; gmc_row_scale_f32
 * yields s and counts rows whose coefficients differ).  The remaining fp32 operand -- W1 forward, s . dH1pre backward --
 * is split into n_split bf16 parts hi + lo (+ lo2) (8 mantissa bits each: 16 / 24 bits), stored as stacked matrices
 * (part p at rows [p * split_rows, p * split_rows + n_rows); rows up to split_rows must be zero, split_rows >= K
 * rounded up to 64).  gmc_gemm_bf16_split multiplies the exact A tile with all parts in one tcgen05.mma per k-step
 * (A is staged once for n_split products), accumulates in fp32 TMEM and adds the partial accumulators in the epilogue:
 *   op 0 (nn): C[M,N] = act(row_scale[m] * (A[M,K] (B_0 + B_1 [+ B_2])) + bias[n])      H1 = relu(s . (XI W1) + b1)
 *   op 2 (tn): C[M,N] (+)= A[K,M]^T (B_0 + B_1 [+ B_2]), split-K over the workspace      dW1 = XI^T (s . dH1pre)
 * row_scale / bias (N % 4 == 0) / relu are nullable / 0 and need accumulate == 0 (they disable split-K).
 * proj_w (nullable; [N rounded up to 64][4] fp32 zero padded, as for gmc_gemm_bf16_bf16out): fused skinny projection
 * proj_out[m, 0..n_proj) = sum_n C[m, n] proj_w[n][.] of the fp32 result (T2 = H1 W2, TrainingNeural.py:83) -- every
 * (row, n-tile) writes one partial into the workspace (gmc_gemm_bf16_split_workspace_bytes with with_projection = 1) and
 * a second launch adds the tiles in order: deterministic, no atomics.
 * gmc_skinny_bwd_split is gmc_skinny_bwd_f32 (fp32 H) whose dHpre leaves as those stacked parts, pre-scaled by s. */
int gmc_f32_split_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t n_rows, int32_t n_cols,
                       int32_t n_split, int64_t split_rows, void* stream);
/* fp16 parts (11 significant bits each: TWO parts carry 22 of fp32's 24 bits -- fp32-grade at the cost of the two-part
 * bf16 GEMM).  Part p is stored multiplied by 2^(lo_shift * p) (the residuals would be fp16 subnormals otherwise);
 * gmc_gemm_bf16_split with the same lo_shift > 0 reads A AND B as fp16 (one 16-bit format per tcgen05.mma kind::f16: the
 * integer features come from gmc_csr_preaggregate_f16) and scales the parts back in its epilogue.  |src| must stay below
 * 65504 (gmc_skinny_bwd_split saturates). */
int gmc_f32_split_f16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t n_rows, int32_t n_cols,
                      int32_t n_split, int64_t split_rows, int32_t lo_shift, void* stream);
int gmc_row_scale_f32(const int32_t* rowptr, const float* coef, int64_t n_rows, float* row_scale,
                      int32_t* nonuniform_count /* device, nullable, incremented */, void* stream);
size_t gmc_gemm_bf16_split_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K, int32_t n_split,
                                           int32_t with_projection);
int gmc_gemm_bf16_split(int32_t op, const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K,
                        int64_t lda, int64_t ldb, int64_t ldc, int32_t n_split, int64_t b_split_rows,
                        const float* row_scale, const float* bias, int32_t relu, const float* proj_w, float* proj_out,
                        int64_t ldp, int32_t n_proj, int32_t accumulate, int32_t lo_shift, void* workspace,
                        size_t workspace_bytes, void* stream);
int gmc_skinny_bwd_split(const float* dT, int64_t lddt, const float* W, const float* H, int64_t ldh,
                         const float* row_scale, void* dH_split, int64_t lddh, int64_t split_rows, int32_t n_split,
                         int32_t lo_shift /* 0: bf16 parts; > 0: fp16 parts, part p scaled by 2^(lo_shift p) */,
                         float* dW, float* dbias, int64_t n_rows, int32_t n_in, int32_t n_out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Both of the above for a batch whose largest graph has max_nodes nodes (known to whoever built the batch): when
 * max_nodes <= n_cols (rounded up to 8) <= 2048 one CTA owns one graph and walks its 2-hop neighbourhoods from a
 * shared-memory copy of the graph's CSR slice (the row-parallel kernel pays an L2 round trip per level); out_f16 != 0
 * writes IEEE fp16.  Same bits as the row-parallel kernels. */
int gmc_csr_preaggregate_graphs(const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                                const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, int64_t n_rows,
                                int32_t n_cols, void* X, int64_t ldx, int32_t out_f16, void* stream);
/* gmc_csr_preaggregate_bf16 with the rows written as IEEE fp16: the A operand of the fp16-part GEMMs (both MMA operands
 * must share one 16-bit format; small integers are exact in either). */
int gmc_csr_preaggregate_f16(const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                             const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X,
                             int64_t ldx, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (csrc/peer.cu) -------------- */
/* The step's only collective -- all-reduce(sum) of the flat weight-gradient buffer between loss.backward() and
 * optimizer.step() (TrainingNeural.py:380-386 under data parallelism; the reference itself is single-process) -- as ONE
 * kernel: every rank's buffer is mapped into every other rank's address space (cudaIpc), rank r sums slice r of all
 * buffers in rank order and writes it to all of them; two flag barriers in the same kernel.  2 MB is pure latency:
 * 36.6 us per call on 8 B200s against 32.4 us for NCCL's all-reduce -- no gain, so the engine uses it only on request
 * (GMC_PEER_ALLREDUCE=1) and NCCL by default.
 *   gmc_peer_alloc / gmc_peer_free           cudaMalloc'ed, zero-filled memory whose cudaIpc handle may be exported
 *   gmc_ipc_get_handle / open / close        gmc_ipc_handle_bytes() opaque bytes per allocation, exchanged by the host
 *   gmc_peer_allreduce_f32                   bufs / flags: HOST arrays of `world` device pointers as mapped in THIS process
 *                                            (entry `rank` = the local allocation; flags: gmc_peer_flag_bytes() zeroed
 *                                            bytes each); `epoch` grows by one per call, identically on all ranks */
int gmc_peer_alloc(size_t bytes, void** ptr);
int gmc_peer_free(void* ptr);
size_t gmc_peer_flag_bytes(void);
int32_t gmc_ipc_handle_bytes(void);
int gmc_ipc_get_handle(const void* ptr, void* handle);
int gmc_ipc_open_handle(const void* handle, void** ptr);
int gmc_ipc_close_handle(void* ptr);
int gmc_peer_allreduce_f32(void* const* bufs, void* const* flags, int32_t world, int32_t rank, int64_t n, uint32_t epoch,
                           void* stream);

/* ---- (e) integer post-processing ------------------------------------------------------ */

/* labels[v] = first argmax_k P[v,k]; the first min(3,n_g) nodes of each graph are forced to
 * 0,1,2 when force_terminals != 0.  Replaces simple_partition_assignment,
 * TestingNeuralNetwork.py:100-122. */
int gmc_argmax_labels(const float* P, int64_t ldp, const int32_t* graph_ptr, int32_t n_graphs,
                      int64_t n_rows, int32_t n_classes, int32_t force_terminals, int32_t* labels,
                      void* stream);

/* cut[g] = sum of w over undirected edges of graph g whose end points differ (int64, exact).
 * wts: int32 per directed edge, NULL = 1.  Replaces calculate_cut_value,
 * TestingNeuralNetwork.py:48-64 (== RandomizedMaxCut.py:48-60). */
int gmc_cut_value_i32(const int32_t* labels, const int32_t* rowptr, const int32_t* colidx,
                      const int32_t* wts, const int32_t* graph_ptr, int32_t n_graphs,
                      int64_t n_rows, int64_t* cut, void* stream);

/* cut[t] for T labelings (uint8 [T][n_nodes]) of ONE graph -- the evaluation loop of the randomized k-way
 * baseline, RandomAlgorithm/RandomizedMaxCut.py:63-122 (one call evaluates a whole chunk of iterations). */
int gmc_cut_value_multi_u8(const uint8_t* labels, const int32_t* rowptr, const int32_t* colidx,
                           const int32_t* wts, int32_t n_nodes, int32_t n_labelings, int64_t* cut,
                           void* stream);

/* numpy's legacy MT19937 stream on the device: n doubles exactly as n calls of np.random.rand() would return them
 * (assign_partitions draws one per node and iteration, TestingNeuralNetwork.py:18-46), continued from the host generator's
 * state (np.random.get_state(): 624 words + position) -- state_out / pos_out go back into np.random.set_state(), so the
 * host generator ends where the reference's own draws would leave it.  All pointers are device pointers. */
int gmc_mt19937_uniform_f64(const uint32_t* state_in, int32_t pos, int64_t n, double* out, uint32_t* state_out,
                            int32_t* pos_out, void* stream);
/* P1: `iters` categorical samplings per graph, keep the FIRST iteration with the maximal cut.
 * U: float64 uniforms, graph g uses U[u_ptr[g] + it*(n_g-3) + (i-3)] for local node i >= 3 --
 * the host draws them with np.random.rand in the reference's call order.  compare_f32 selects
 * `float32(r) < cumsum` (numpy >= 2) or `r < float64(cumsum)` (numpy 1.x).  cuts_ws: int64
 * [n_graphs*iters].  Outputs: best_labels [n_rows], best_cut [n_graphs], best_iter [n_graphs].
 * Replaces assign_partitions + post_processing_optimization, TestingNeuralNetwork.py:18-46, 66-98. */
int gmc_sample_best_cut(const float* P, int64_t ldp, const double* U, const int64_t* u_ptr,
                        const int32_t* rowptr, const int32_t* colidx, const int32_t* wts,
                        const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows,
                        int32_t n_classes, int32_t iters, int32_t compare_f32, int64_t* cuts_ws,
                        int32_t* best_labels, int64_t* best_cut, int32_t* best_iter, void* stream);

/* P2: greedy best-improvement node-move local search, <= iters moves per graph, nodes with
 * local index < n_frozen never move; ties -> lowest node then lowest class; stops when no move
 * has positive gain.  k-way generalisation of greedy_maxcut,
 * "Other Algorithms/huerestics_multi-max.ipynb":L5817-5854. */
int gmc_greedy_node_move(const int32_t* labels_in, const int32_t* rowptr, const int32_t* colidx,
                         const int32_t* wts, const int32_t* graph_ptr, int32_t n_graphs,
                         int64_t n_rows, int32_t n_classes, int32_t iters, int32_t n_frozen,
                         int32_t* labels_out, int64_t* cut_out, int32_t* moves_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCNMAXCUT_H_ */
