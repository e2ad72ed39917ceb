// graph_ops.cu -- per-batch graph preparation: degree norms, A_hat edge coefficients,
// dense padded adjacency rows (device-side graphExtender).  HBM-bound integer/float work.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace gmc {

__global__ void degree_norm_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows,
                                   float* __restrict__ norm, int32_t* __restrict__ zero_count) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int is_zero = 0;
    if (v < n_rows) {
        int deg = rowptr[v + 1] - rowptr[v];
        is_zero = (deg == 0);
        // DGL: degs.clamp(min=1) then pow(-0.5)
        norm[v] = 1.0f / sqrtf((float)(deg < 1 ? 1 : deg));
    }
    if (zero_count) {
        unsigned m = __ballot_sync(0xffffffffu, is_zero);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(zero_count, __popc(m));
    }
}

// Rows of the reference's graphs have 6-8 edges: a warp per row left 3/4 of the lanes idle and cost 4 M warps per
// launch at config 3 (0.62 ms for edge_coef, 1.2 ms for the feature scatter inside the end-to-end step).  Eight
// lanes own a row (four rows per warp) and stride over its edges.
constexpr int kRowLanes = 8;

// graph of a row: the batches the reference builds have equal-sized graphs, so row / size(graph 0) is verified with two
// loads before falling back to the binary search
__device__ __forceinline__ int find_graph_guess(const int32_t* __restrict__ graph_ptr, int n_graphs, int64_t row) {
    const int n0 = __ldg(graph_ptr + 1) - __ldg(graph_ptr);
    if (n0 > 0) {
        // 32-bit division whenever the row index allows it: the 64-bit one is a ~100-instruction subroutine per call
        const int64_t g = row < (1ll << 31) ? (int64_t)((uint32_t)row / (uint32_t)n0) : row / n0;
        if (g < n_graphs && (int64_t)__ldg(graph_ptr + g) <= row && row < (int64_t)__ldg(graph_ptr + g + 1)) return (int)g;
    }
    return find_graph(graph_ptr, n_graphs, row);
}

__global__ void edge_coef_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 const float* __restrict__ vals, const float* __restrict__ ns,
                                 const float* __restrict__ nd, int64_t n_rows, float* __restrict__ coef) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    float d = nd ? nd[row] : 1.0f;
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        float w = vals ? vals[e] : 1.0f;
        float s = ns ? __ldg(ns + colidx[e]) : 1.0f;
        coef[e] = (w * s) * d;
    }
}

__global__ void densify_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                               const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr,
                               int n_graphs, int64_t n_rows, int n_cols, float* __restrict__ X, int64_t ldx,
                               int* __restrict__ bad) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int g = find_graph_guess(graph_ptr, n_graphs, row);
    int base = graph_ptr[g];
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        int local = colidx[e] - base;
        if (local < 0 || local >= n_cols) { if (bad) *bad = 1; continue; }
        X[row * ldx + local] = vals ? vals[e] : 1.0f;
    }
}

__global__ void densify_bf16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, int n_graphs,
                                    int64_t n_rows, int n_cols, __nv_bfloat16* __restrict__ X, int64_t ldx) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int g = find_graph_guess(graph_ptr, n_graphs, row);
    int base = graph_ptr[g];
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        int local = colidx[e] - base;
        if (local < 0 || local >= n_cols) continue;
        X[row * ldx + local] = __float2bfloat16_rn(vals ? vals[e] : 1.0f);
    }
}

// value != 0: X[row, local col] = bf16(vals ? vals[e] : 1) at every edge; value == 0: zeros at the same positions
__global__ void scatter_bf16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, int n_graphs,
                                    int64_t n_rows, int n_cols, __nv_bfloat16* __restrict__ X, int64_t ldx, int clear) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kRowLanes;
    int lane = threadIdx.x & (kRowLanes - 1);
    if (row >= n_rows) return;
    int g = find_graph_guess(graph_ptr, n_graphs, row);
    int base = graph_ptr[g];
    int e0 = rowptr[row], e1 = rowptr[row + 1];
    for (int e = e0 + lane; e < e1; e += kRowLanes) {
        int local = colidx[e] - base;
        if (local < 0 || local >= n_cols) continue;
        X[row * ldx + local] = __float2bfloat16_rn(clear ? 0.0f : (vals ? vals[e] : 1.0f));
    }
}

// Pre-aggregated first-layer features: XA = A_hat X with X the zero-padded (weighted) adjacency rows, written as bf16.
//   XA[v, local(j)] = sum_{u in N(v)} coef(v,u) * w(u,j)
// A_hat (X W1) = (A_hat X) W1, and X is a function of the graph alone, so GraphConv layer 1's aggregation
// (TrainingNeural.py:80, dgl update_all) can be applied to the features once instead of to the activations every step.
// One warp per row: the row is accumulated in a shared-memory fp32 buffer (neighbours u in CSR order, the lanes over
// N(u): distinct columns within one u, so plain adds in a fixed order -- deterministic), then written out whole as
// bf16 (zeros included: no memset of the [N, n_cols] matrix).
constexpr int kPreaggWarps = 8;

// F16OUT: the row is written as IEEE fp16 instead of bf16 (the integer features of the 'f16x2' GEMM path: small integers
// are exact in both, the MMA needs both operands in ONE 16-bit format)
template <bool F16OUT>
__global__ void __launch_bounds__(kPreaggWarps * 32)
preaggregate_bf16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                         const float* __restrict__ coef, const float* __restrict__ vals,
                         const int32_t* __restrict__ graph_ptr, int n_graphs, int64_t n_rows, int n_cols, int ncp,
                         __nv_bfloat16* __restrict__ X, int64_t ldx, const uint8_t* __restrict__ only, int pad_cols) {
    extern __shared__ __align__(16) float preagg_rows[];              // kPreaggWarps x ncp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* buf = preagg_rows + warp * ncp;
    const int64_t stride = (int64_t)gridDim.x * kPreaggWarps;
    // graph of a row without a division per row: equal-sized graphs (what the reference builds) give g = row / n0,
    // evaluated in float and verified against graph_ptr; anything else takes the binary search
    const int n0 = __ldg(graph_ptr + 1) - __ldg(graph_ptr);
    const float inv_n0 = n0 > 0 ? 1.0f / (float)n0 : 0.f;

    // ncu on the first version (gpurun_out/prof_preagg): ~690 warp instructions per row (a scan, an owner search and
    // ordered accumulation rounds for every row, a 64-bit division for the graph id), 57 % issue slots -- instruction
    // bound.  Software-pipelining the four dependent loads of a row (extent -> neighbours -> their extents -> 2-hop
    // columns) made it slower both before (4.2 -> 4.6 ms) and after (3.0 -> 3.7 ms) the diet below, so the loop is the
    // plain one and the common case is kept lean:
    //   fast row  = unit weights, <= 32 neighbours which all have one degree D and one coefficient c, <= 64 two-hop
    //               pairs: pair p belongs to neighbour p / D and every addend is the same c, so the adds can be
    //               unordered shared-memory atomics and still reproduce bit for bit (equal addends commute exactly);
    //   other rows = neighbours one after the other in CSR order, lanes over N(u) (distinct columns): fixed order.
    // Config 3 (4.1 M rows, 8.4 GB written): 3.0 ms.
    // ncu (round 2, profiles/r02_preagg_ncu.txt): the shared-memory pipe was the limiter at 70 % -- per row 40 store + 85
    // load + 23 atomic wavefronts: full zeroing of the 4 KB row buffer, float4 reads at a 32-byte lane stride (two-way
    // bank conflicts) and atomicAdd(float) on shared memory, which is a compare-and-swap loop (ATOMS.CAST.SPIN).  Now:
    //  * fast rows COUNT with native integer atomics into the mantissa of the float 2^23 (bits 0x4B000000 + k is the
    //    float 2^23 + k), so an entry is read back as fmaf(f, c, -2^23 c) = k c with one rounding;
    //  * the buffer is initialised once and a fast row restores only the <= 64 entries it touched;
    //  * the read-out uses a 16-byte lane stride (conflict-free LDS.128) and 8-byte stores.
    constexpr uint32_t kMagicBits = 0x4B000000u;                       // 8388608.0f
    const float magic = __uint_as_float(kMagicBits);
    for (int c = lane * 4; c < ncp; c += 128) *reinterpret_cast<float4*>(buf + c) = make_float4(magic, magic, magic, magic);
    __syncwarp();
    for (int64_t row = (int64_t)blockIdx.x * kPreaggWarps + warp; row < n_rows; row += stride) {
        if (only && !__ldg(only + row)) continue;                      // second pass: rows the counting kernel left
        const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
        int g = (int)(((float)row + 0.5f) * inv_n0);
        if (g >= n_graphs) g = n_graphs - 1;
        int base = __ldg(graph_ptr + g);
        if (n0 <= 0 || (int64_t)base > row || row >= (int64_t)__ldg(graph_ptr + g + 1)) {
            g = find_graph(graph_ptr, n_graphs, row);
            base = __ldg(graph_ptr + g);
        }
        const int cnt = e1 - e0;
        int my_f0 = 0, my_deg = 0;
        float my_c = 0.f;
        if (lane < min(cnt, 32)) {
            const int u = __ldg(colidx + e0 + lane);
            my_c = __ldg(coef + e0 + lane);
            my_f0 = __ldg(rowptr + u);
            my_deg = __ldg(rowptr + u + 1) - my_f0;
        }
        const int D0 = __shfl_sync(0xffffffffu, my_deg, 0);
        const float c0 = __shfl_sync(0xffffffffu, my_c, 0);
        const bool same = __all_sync(0xffffffffu, lane >= cnt || (my_deg == D0 && my_c == c0));
        const bool fast = same && !vals && cnt > 0 && cnt <= 32 && D0 > 0 && cnt * D0 <= 64;
        int touched[2] = {-1, -1};
        if (!fast)                                                     // general rows accumulate floats from zero
            for (int c = lane * 4; c < ncp; c += 128) *reinterpret_cast<float4*>(buf + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        if (fast) {
            const int total = cnt * D0;
            const float inv_d = 1.0f / (float)D0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int p = 32 * j + lane;
                const int owner = p < total ? (int)(((float)p + 0.5f) * inv_d) : 0;      // p / D0, exact for p < 64
                const int o_f0 = __shfl_sync(0xffffffffu, my_f0, owner);
                if (p < total) {
                    const int local = __ldg(colidx + o_f0 + (p - owner * D0)) - base;
                    if (local >= 0 && local < n_cols) { atomicAdd(reinterpret_cast<unsigned int*>(buf) + local, 1u); touched[j] = local; }
                }
            }
        } else {
            for (int eb = e0; eb < e1; eb += 32) {
                const int n_here = min(32, e1 - eb);
                if (eb > e0) {
                    my_f0 = 0; my_deg = 0; my_c = 0.f;
                    if (lane < n_here) {
                        const int u = __ldg(colidx + eb + lane);
                        my_c = __ldg(coef + eb + lane);
                        my_f0 = __ldg(rowptr + u);
                        my_deg = __ldg(rowptr + u + 1) - my_f0;
                    }
                }
                for (int q = 0; q < n_here; ++q) {
                    const int f0 = __shfl_sync(0xffffffffu, my_f0, q), dg = __shfl_sync(0xffffffffu, my_deg, q);
                    const float c = __shfl_sync(0xffffffffu, my_c, q);
                    for (int f = f0 + lane; f < f0 + dg; f += 32) {
                        const int local = __ldg(colidx + f) - base;
                        if (local >= 0 && local < n_cols) buf[local] += c * (vals ? __ldg(vals + f) : 1.0f);
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        __nv_bfloat16* xr = X + row * ldx;
        // lane l reads float4 number l + 32 i (16-byte stride: conflict-free LDS.128) and stores its four bf16 as 8 bytes
        const float sc = fast ? c0 : 1.0f, off = fast ? -magic * c0 : 0.0f;   // 2^23 c0 is exact: one rounding of k c0
        for (int c4 = lane; c4 * 4 < ncp; c4 += 32) {
            const float4 a = *reinterpret_cast<const float4*>(buf + c4 * 4);
            const float v0 = fmaf(a.x, sc, off), v1 = fmaf(a.y, sc, off), v2 = fmaf(a.z, sc, off), v3 = fmaf(a.w, sc, off);
            uint2 o;
            if (F16OUT) {
                const __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
                o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
            } else {
                const __nv_bfloat162 b0 = __floats2bfloat162_rn(v0, v1), b1 = __floats2bfloat162_rn(v2, v3);
                o.x = *reinterpret_cast<const uint32_t*>(&b0); o.y = *reinterpret_cast<const uint32_t*>(&b1);
            }
            *reinterpret_cast<uint2*>(xr + c4 * 4) = o;
        }
        // pad columns up to the 128-byte line boundary as zeros: a row that ends inside a line (2000 of 2048 bytes) makes
        // its last sector a read-modify-write and costs the whole stream 40 % of its bandwidth (profiles/r02r_write_pattern_probe)
        for (int c = ncp + lane * 4; c < pad_cols; c += 128) *reinterpret_cast<uint2*>(xr + c) = make_uint2(0u, 0u);
        __syncwarp();                                                  // everyone has read the row
        if (fast) {
            if (touched[0] >= 0) buf[touched[0]] = magic;
            if (touched[1] >= 0) buf[touched[1]] = magic;
        } else {
            for (int c = lane * 4; c < ncp; c += 128) *reinterpret_cast<float4*>(buf + c) = make_float4(magic, magic, magic, magic);
        }
        __syncwarp();
    }
}

// Per-GRAPH form of preaggregate_bf16_kernel for batches of small graphs (every training / test set of the reference:
// n <= 1000).  A row of A_hat X is a 2-hop walk -- row extent -> neighbours -> THEIR extents -> 2-hop columns -- and in the
// row-parallel kernel above each of those four dependent levels is an L2 round trip.  Here one CTA owns one graph: its CSR slice (local row pointers,
// local 16-bit column ids: 4 (n + 1) + 2 nnz bytes = 18 KB at n = 1000, d = 7) is staged in shared memory with one
// coalesced pass, and the whole walk runs on shared-memory latency; only the per-row coefficient (needed at read-out) and
// the 2 KB row store touch global memory.  Row arithmetic, order and rounding are those of the row-parallel kernel, so the
// two produce the same bits.  Graphs that do not fit the staging buffers are walked from global memory by the same code.
constexpr int kPgWarps = 8;

static bool graph_kernel_enabled() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("GMC_PREAGG_GRAPH"); cached = (e && e[0] == '0') ? 0 : 1; }
    return cached == 1;
}

// Rows of one graph, one warp per row.  STAGED: the CSR slice is in shared memory (s_rp relative row pointers, s_ci 16-bit
// local column ids, 0xFFFF = outside), else it is read from global memory.  Compile-time so that the accessors are plain
// loads: as run-time selects inside lambdas they cost a third of the kernel's 358 instructions per row and put their
// closure in local memory (ncu: 4 LDL + 3 STL per row).  The row's coefficients -- the only global load on the walk -- are
// requested one row ahead (37 % of the stall samples sat on their first use).
template <bool F16OUT, bool STAGED>
__device__ __forceinline__ void pg_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                        const float* __restrict__ coef, const int32_t* __restrict__ s_rp,
                                        const uint16_t* __restrict__ s_ci, float* __restrict__ buf, int base, int n,
                                        int e_base, int n_cols, int ncp, __nv_bfloat16* __restrict__ X, int64_t ldx,
                                        int lane, int warp, int pad_cols) {
    constexpr uint32_t kMagicBits = 0x4B000000u;                       // 8388608.0f: counts live in its mantissa
    const float magic = __uint_as_float(kMagicBits);
#define PG_RP(i) (STAGED ? s_rp[(i)] : (__ldg(rowptr + base + (i)) - e_base))
    auto ci = [=](int e) -> int {
        if (STAGED) { const int v = s_ci[e]; return v == 0xFFFF ? -1 : v; }
        const int v = __ldg(colidx + e_base + e) - base;
        return (v >= 0 && v < n) ? v : -1;
    };
    int v = warp;
    int e0 = 0, cnt = 0;
    float my_c = 0.f;
    if (v < n) {
        e0 = PG_RP(v); cnt = PG_RP(v + 1) - e0;
        if (lane < min(cnt, 32)) my_c = __ldg(coef + e_base + e0 + lane);
    }
    for (; v < n; v += kPgWarps) {
        // next row's extent and coefficients: in flight while this row is built
        const int vn = v + kPgWarps;
        int e0_n = 0, cnt_n = 0;
        float c_n = 0.f;
        if (vn < n) {
            e0_n = PG_RP(vn); cnt_n = PG_RP(vn + 1) - e0_n;
            if (lane < min(cnt_n, 32)) c_n = __ldg(coef + e_base + e0_n + lane);
        }
        int my_f0 = 0, my_deg = 0;
        if (lane < min(cnt, 32)) {
            const int u = ci(e0 + lane);
            if (u >= 0) { my_f0 = PG_RP(u); my_deg = PG_RP(u + 1) - my_f0; }
        }
        const int D0 = __shfl_sync(0xffffffffu, my_deg, 0);
        const float c0 = __shfl_sync(0xffffffffu, my_c, 0);
        const bool same = __all_sync(0xffffffffu, lane >= cnt || (my_deg == D0 && my_c == c0));
        const bool fast = same && cnt > 0 && cnt <= 32 && D0 > 0 && cnt * D0 <= 64;
        int touched0 = -1, touched1 = -1;
        if (!fast)
            for (int c = lane * 4; c < ncp; c += 128) *reinterpret_cast<float4*>(buf + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        if (fast) {
            const int total = cnt * D0;
            const float inv_d = 1.0f / (float)D0;
            {
                const int p = lane;
                const int owner = p < total ? (int)(((float)p + 0.5f) * inv_d) : 0;
                const int o_f0 = __shfl_sync(0xffffffffu, my_f0, owner);
                if (p < total) {
                    const int local = ci(o_f0 + (p - owner * D0));
                    if (local >= 0 && local < n_cols) { atomicAdd(reinterpret_cast<unsigned int*>(buf) + local, 1u); touched0 = local; }
                }
            }
            {
                const int p = 32 + lane;
                const int owner = p < total ? (int)(((float)p + 0.5f) * inv_d) : 0;
                const int o_f0 = __shfl_sync(0xffffffffu, my_f0, owner);
                if (p < total) {
                    const int local = ci(o_f0 + (p - owner * D0));
                    if (local >= 0 && local < n_cols) { atomicAdd(reinterpret_cast<unsigned int*>(buf) + local, 1u); touched1 = local; }
                }
            }
        } else {
            // neighbours one after the other in CSR order, lanes over N(u) (distinct columns): fixed order
            float cc = my_c;
            for (int eb = e0; eb < e0 + cnt; eb += 32) {
                const int n_here = min(32, e0 + cnt - eb);
                if (eb > e0) {
                    my_f0 = 0; my_deg = 0; cc = 0.f;
                    if (lane < n_here) {
                        cc = __ldg(coef + e_base + eb + lane);
                        const int u = ci(eb + lane);
                        if (u >= 0) { my_f0 = PG_RP(u); my_deg = PG_RP(u + 1) - my_f0; }
                    }
                }
                for (int q = 0; q < n_here; ++q) {
                    const int f0 = __shfl_sync(0xffffffffu, my_f0, q), dg = __shfl_sync(0xffffffffu, my_deg, q);
                    const float c = __shfl_sync(0xffffffffu, cc, q);
                    for (int f = f0 + lane; f < f0 + dg; f += 32) {
                        const int local = ci(f);
                        if (local >= 0 && local < n_cols) buf[local] += c;
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        __nv_bfloat16* xr = X + (int64_t)(base + v) * ldx;
        const float sc = fast ? c0 : 1.0f, off = fast ? -magic * c0 : 0.0f;
        for (int c4 = lane; c4 * 4 < ncp; c4 += 32) {
            const float4 a = *reinterpret_cast<const float4*>(buf + c4 * 4);
            const float v0 = fmaf(a.x, sc, off), v1 = fmaf(a.y, sc, off), v2 = fmaf(a.z, sc, off), v3 = fmaf(a.w, sc, off);
            uint2 o;
            if (F16OUT) {
                const __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
                o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
            } else {
                const __nv_bfloat162 b0 = __floats2bfloat162_rn(v0, v1), b1 = __floats2bfloat162_rn(v2, v3);
                o.x = *reinterpret_cast<const uint32_t*>(&b0); o.y = *reinterpret_cast<const uint32_t*>(&b1);
            }
            *reinterpret_cast<uint2*>(xr + c4 * 4) = o;
        }
        for (int c = ncp + lane * 4; c < pad_cols; c += 128) *reinterpret_cast<uint2*>(xr + c) = make_uint2(0u, 0u);   // full lines
        __syncwarp();
        if (fast) {
            if (touched0 >= 0) buf[touched0] = magic;
            if (touched1 >= 0) buf[touched1] = magic;
        } else {
            for (int c = lane * 4; c < ncp; c += 128) *reinterpret_cast<float4*>(buf + c) = make_float4(magic, magic, magic, magic);
        }
        __syncwarp();
        e0 = e0_n; cnt = cnt_n; my_c = c_n;
    }
#undef PG_RP
}

template <bool F16OUT>
__global__ void __launch_bounds__(kPgWarps * 32)
preaggregate_graph_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                          const float* __restrict__ coef, const int32_t* __restrict__ graph_ptr, int n_graphs, int n_cols,
                          int ncp, int cap_nodes, int cap_edges, __nv_bfloat16* __restrict__ X, int64_t ldx, int pad_cols) {
    extern __shared__ __align__(16) float pg_smem[];                   // kPgWarps x ncp row buffers, then the CSR slice
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* buf = pg_smem + warp * ncp;
    int32_t* s_rp = reinterpret_cast<int32_t*>(pg_smem + kPgWarps * ncp);
    uint16_t* s_ci = reinterpret_cast<uint16_t*>(s_rp + cap_nodes + 1);
    const float magic = __uint_as_float(0x4B000000u);
    for (int c = lane * 4; c < ncp; c += 128) *reinterpret_cast<float4*>(buf + c) = make_float4(magic, magic, magic, magic);
    for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
        const int base = __ldg(graph_ptr + g);
        const int n = __ldg(graph_ptr + g + 1) - base;
        const int e_base = __ldg(rowptr + base);
        const int nnz = __ldg(rowptr + base + n) - e_base;
        const bool staged = n <= cap_nodes && nnz <= cap_edges && n <= 65535;      // block-uniform
        __syncthreads();                                               // the previous graph's slice is no longer read
        if (staged) {
            for (int i = threadIdx.x; i <= n; i += kPgWarps * 32) s_rp[i] = __ldg(rowptr + base + i) - e_base;
            for (int e = threadIdx.x; e < nnz; e += kPgWarps * 32) {
                const int local = __ldg(colidx + e_base + e) - base;   // a neighbour outside the graph cannot be staged
                s_ci[e] = (local >= 0 && local < n) ? (uint16_t)local : (uint16_t)0xFFFF;
            }
        }
        __syncthreads();
        if (staged) pg_rows<F16OUT, true>(rowptr, colidx, coef, s_rp, s_ci, buf, base, n, e_base, n_cols, ncp, X, ldx, lane, warp, pad_cols);
        else pg_rows<F16OUT, false>(rowptr, colidx, coef, s_rp, s_ci, buf, base, n, e_base, n_cols, ncp, X, ldx, lane, warp, pad_cols);
    }
}

// Counting form of the fast rows (unit weights, <= 8 neighbours which all have one degree D <= 8 and one coefficient
// c): every 2-hop addend is the same c, so a row of A_hat X is c x (number of 2-hop paths into each column).  Eight
// lanes own a row and count the paths in BYTES (1 KB per row instead of a 4 KB fp32 buffer: 32 rows per CTA, ~190 rows
// in flight per SM instead of 48, which is what the dependent loads of a row need), then map counts to bf16 through a
// 65-entry table k -> bf16(k c) and write the row whole.  Rows that do not qualify are flagged in `slow` and left to
// preaggregate_bf16_kernel (second launch, masked).
constexpr int kPcLanes = 8, kPcRows = 32, kPcLut = 72;

__global__ void __launch_bounds__(kPcLanes * kPcRows)
preaggregate_counts_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                           const float* __restrict__ coef, const int32_t* __restrict__ graph_ptr, int n_graphs,
                           int64_t n_rows, int n_cols, int ncp, __nv_bfloat16* __restrict__ X, int64_t ldx,
                           uint8_t* __restrict__ slow) {
    extern __shared__ __align__(16) uint32_t pc_smem[];               // kPcRows x wstride count words, then the tables
    const int wstride = ((ncp >> 2) + 3) & ~3;                         // words per row, 16-byte multiple
    const int sub = threadIdx.x & (kPcLanes - 1), grp = threadIdx.x / kPcLanes;
    uint32_t* cnt = pc_smem + grp * wstride;
    uint16_t* lut = reinterpret_cast<uint16_t*>(pc_smem + kPcRows * wstride) + grp * kPcLut;
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~(kPcLanes - 1));
    const int n0 = __ldg(graph_ptr + 1) - __ldg(graph_ptr);
    const float inv_n0 = n0 > 0 ? 1.0f / (float)n0 : 0.f;
    for (int64_t row = (int64_t)blockIdx.x * kPcRows + grp; row < n_rows; row += (int64_t)gridDim.x * kPcRows) {
        for (int w = sub * 4; w < wstride; w += kPcLanes * 4) *reinterpret_cast<uint4*>(cnt + w) = make_uint4(0u, 0u, 0u, 0u);
        const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
        int g = (int)(((float)row + 0.5f) * inv_n0);
        if (g >= n_graphs) g = n_graphs - 1;
        int base = __ldg(graph_ptr + g);
        if (n0 <= 0 || (int64_t)base > row || row >= (int64_t)__ldg(graph_ptr + g + 1))
            base = __ldg(graph_ptr + find_graph(graph_ptr, n_graphs, row));
        const int deg = e1 - e0;
        int my_f0 = 0, my_deg = 0;
        float my_c = 0.f;
        if (sub < min(deg, kPcLanes)) {
            const int u = __ldg(colidx + e0 + sub);
            my_c = __ldg(coef + e0 + sub);
            my_f0 = __ldg(rowptr + u);
            my_deg = __ldg(rowptr + u + 1) - my_f0;
        }
        const int D0 = __shfl_sync(gmask, my_deg, 0, kPcLanes);
        const float c0 = __shfl_sync(gmask, my_c, 0, kPcLanes);
        const bool mine_ok = sub >= deg || (my_deg == D0 && my_c == c0);
        const bool fast = deg <= kPcLanes && D0 <= kPcLanes && (__ballot_sync(gmask, mine_ok) & gmask) == gmask;
        if (!fast) {                                                   // group-uniform
            if (sub == 0) slow[row] = 1;
            __syncwarp(gmask);
            continue;
        }
        __syncwarp(gmask);                                             // the count row is zero
#pragma unroll
        for (int q = 0; q < kPcLanes; ++q) {
            const int f0 = __shfl_sync(gmask, my_f0, q, kPcLanes);
            if (q < deg && sub < D0) {
                const int local = __ldg(colidx + f0 + sub) - base;
                if (local >= 0 && local < n_cols) atomicAdd(cnt + (local >> 2), 1u << ((local & 3) * 8));
            }
        }
        for (int k = sub; k < kPcLut; k += kPcLanes) lut[k] = __bfloat16_as_ushort(__float2bfloat16_rn((float)k * c0));
        __syncwarp(gmask);
        __nv_bfloat16* xr = X + row * ldx;
        for (int c8 = sub; c8 * 8 < ncp; c8 += kPcLanes) {
            const uint2 w = *reinterpret_cast<const uint2*>(cnt + c8 * 2);
            uint4 o;
            o.x = (uint32_t)lut[w.x & 0xffu] | ((uint32_t)lut[(w.x >> 8) & 0xffu] << 16);
            o.y = (uint32_t)lut[(w.x >> 16) & 0xffu] | ((uint32_t)lut[w.x >> 24] << 16);
            o.z = (uint32_t)lut[w.y & 0xffu] | ((uint32_t)lut[(w.y >> 8) & 0xffu] << 16);
            o.w = (uint32_t)lut[(w.y >> 16) & 0xffu] | ((uint32_t)lut[w.y >> 24] << 16);
            *reinterpret_cast<uint4*>(xr + c8 * 8) = o;
        }
        if (sub == 0) slow[row] = 0;
        __syncwarp(gmask);                                             // counts and table are free for the next row
    }
}

// 8 elements per thread: two 128-bit loads, one 128-bit store
__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ src, int64_t lds, __nv_bfloat16* __restrict__ dst, int64_t ldd, int64_t n_rows,
                   int n_cols) {
    const int c8 = (n_cols + 7) / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = i / c8;
    const int c = (int)(i - r * c8) * 8;
    if (r >= n_rows) return;
    const float* s = src + r * lds + c;
    __nv_bfloat16* d = dst + r * ldd + c;
    if (c + 8 <= n_cols && ((lds | ldd) % 8 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
        (reinterpret_cast<uintptr_t>(dst) % 16 == 0)) {
        const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
        __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                               __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
        *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(o);
    } else {
        for (int k = 0; k < 8 && c + k < n_cols; ++k) d[k] = __float2bfloat16_rn(s[k]);
    }
}

}  // namespace gmc

extern "C" {

int gmc_degree_norm_f32(const int32_t* rowptr, int64_t n_rows, float* norm, int32_t* zero_degree_count,
                        void* stream) {
    GMC_REQUIRE(rowptr && norm && n_rows >= 0, "gmc_degree_norm_f32: null pointer or negative size");
    cudaStream_t s = gmc::as_stream(stream);
    if (zero_degree_count) GMC_CUDA(cudaMemsetAsync(zero_degree_count, 0, sizeof(int32_t), s));
    if (n_rows == 0) return GMC_OK;
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows, threads);
    gmc::degree_norm_kernel<<<(unsigned)blocks, threads, 0, s>>>(rowptr, n_rows, norm, zero_degree_count);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_edge_coef_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, const float* norm_src,
                      const float* norm_dst, int64_t n_rows, float* coef, void* stream) {
    GMC_REQUIRE(rowptr && colidx && coef && n_rows >= 0, "gmc_edge_coef_f32: null pointer or negative size");
    if (n_rows == 0) return GMC_OK;
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::edge_coef_kernel<<<(unsigned)blocks, threads, 0, gmc::as_stream(stream)>>>(rowptr, colidx, vals, norm_src,
                                                                                    norm_dst, n_rows, coef);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_csr_densify_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* graph_ptr,
                        int32_t n_graphs, int64_t n_rows, int32_t n_cols, float* X, int64_t ldx, void* stream) {
    GMC_REQUIRE(rowptr && colidx && graph_ptr && X, "gmc_csr_densify_f32: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_csr_densify_f32: bad sizes");
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = gmc::as_stream(stream);
    // rows are stored back to back: one flat memset (pad columns included) runs at write bandwidth, the pitched 2-D
    // form does not
    GMC_CUDA(cudaMemsetAsync(X, 0, (size_t)ldx * sizeof(float) * (size_t)(n_rows - 1) + (size_t)n_cols * sizeof(float), s));
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::densify_kernel<<<(unsigned)blocks, threads, 0, s>>>(rowptr, colidx, vals, graph_ptr, n_graphs, n_rows,
                                                             n_cols, X, ldx, nullptr);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// bf16 twin of gmc_csr_densify_f32 (0/1 adjacency entries are exact in bf16): the features of the bf16 GEMM path
int gmc_csr_densify_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* graph_ptr,
                         int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X, int64_t ldx, void* stream) {
    GMC_REQUIRE(rowptr && colidx && graph_ptr && X, "gmc_csr_densify_bf16: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_csr_densify_bf16: bad sizes");
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = gmc::as_stream(stream);
    GMC_CUDA(cudaMemsetAsync(X, 0, (size_t)ldx * 2 * (size_t)(n_rows - 1) + (size_t)n_cols * 2, s));
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::densify_bf16_kernel<<<(unsigned)blocks, threads, 0, s>>>(rowptr, colidx, vals, graph_ptr, n_graphs, n_rows, n_cols,
                                                                  reinterpret_cast<__nv_bfloat16*>(X), ldx);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// Incremental form of gmc_csr_densify_bf16 for a feature buffer that is reused step after step: clear = 0 writes the
// adjacency entries of the batch into an X that is zero everywhere (no memset), clear = 1 writes zeros at the same
// positions, restoring the all-zero state.  Touches nnz 32-byte sectors instead of the whole n_rows x ldx matrix
// (config 3: 0.9 GB instead of 8.4 GB) -- graphExtender.py:106-111 for a stream of graphs.
int gmc_csr_scatter_bf16(const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* graph_ptr,
                         int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X, int64_t ldx, int32_t clear,
                         void* stream) {
    GMC_REQUIRE(rowptr && colidx && graph_ptr && X, "gmc_csr_scatter_bf16: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_csr_scatter_bf16: bad sizes");
    if (n_rows == 0) return GMC_OK;
    int threads = 256;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows * gmc::kRowLanes, threads);
    gmc::scatter_bf16_kernel<<<(unsigned)blocks, threads, 0, gmc::as_stream(stream)>>>(
        rowptr, colidx, vals, graph_ptr, n_graphs, n_rows, n_cols, reinterpret_cast<__nv_bfloat16*>(X), ldx, clear);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

// XA = A_hat X in bf16 for X = the zero-padded (weighted) adjacency rows of the batch (graphExtender.py:106-111) and
// A_hat given by its per-edge values `coef` (gmc_edge_coef_f32): the features of the pre-aggregated first layer,
// H1 = relu(XA W1 + b1) == relu(A_hat (X W1) + b1) (TrainingNeural.py:80-81).  Every row is written whole (ldx % 8 == 0,
// 16-byte aligned base); rows are built in a fixed order, so the result is bitwise reproducible.
size_t gmc_csr_preaggregate_workspace_bytes(int64_t n_rows) { return n_rows > 0 ? (size_t)n_rows : 0; }

static int preaggregate_impl(const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                             const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X,
                             int64_t ldx, void* workspace, size_t workspace_bytes, void* stream, bool f16,
                             int32_t max_nodes = -1) {
    GMC_REQUIRE(rowptr && colidx && coef && graph_ptr && X, "gmc_csr_preaggregate_bf16: null pointer");
    GMC_REQUIRE(n_graphs >= 0 && n_rows >= 0 && n_cols > 0 && ldx >= n_cols, "gmc_csr_preaggregate_bf16: bad sizes");
    const int ncp = (n_cols + 7) & ~7;
    GMC_REQUIRE(ldx % 8 == 0 && ldx >= ncp && gmc::aligned16(X) && ncp <= 6144,
                "gmc_csr_preaggregate_bf16: needs ldx %% 8 == 0 covering n_cols rounded up to 8, a 16-byte aligned X and "
                "n_cols <= 6144");
    if (n_rows == 0) return GMC_OK;
    cudaStream_t s = gmc::as_stream(stream);
    __nv_bfloat16* Xb = reinterpret_cast<__nv_bfloat16*>(X);
    // gmc_csr_preaggregate_graphs (max_nodes >= 0) also owns the pad columns of a row up to the next 128-byte line (zeros)
    int pad_cols = ncp;
    if (max_nodes >= 0) {
        const int64_t line = ((int64_t)ncp + 63) & ~(int64_t)63;
        pad_cols = (int)(line <= ldx ? line : ncp);
    }
    // counting kernel first when the caller passes the row-mask workspace (unit weights and a count row that fits): it
    // flags the rows it leaves.  Measured at config 3: 3.24 ms for the pair of launches against 3.00 ms for the general
    // kernel alone -- byte counters lift the occupancy limit (160 rows in flight per SM instead of 48) without making
    // the kernel faster, so the Python wrapper does not pass a workspace by default.
    const uint8_t* only = nullptr;
    const int wstride = ((ncp >> 2) + 3) & ~3;
    const size_t pc_smem = (size_t)gmc::kPcRows * wstride * 4 + (size_t)gmc::kPcRows * gmc::kPcLut * 2;
    if (!f16 && !vals && workspace && workspace_bytes >= (size_t)n_rows && pc_smem <= 100 * 1024) {
        static bool attr2 = false;
        if (!attr2) {
            GMC_CUDA(cudaFuncSetAttribute(gmc::preaggregate_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr2 = true;
        }
        int per_sm = (int)((220 * 1024) / (pc_smem + 1024));
        if (per_sm > 8) per_sm = 8;
        int64_t blocks = gmc::ceil_div<int64_t>(n_rows, gmc::kPcRows);
        const int64_t cap = (int64_t)gmc::sm_count() * per_sm;
        if (blocks > cap) blocks = cap;
        uint8_t* slow = reinterpret_cast<uint8_t*>(workspace);
        gmc::preaggregate_counts_kernel<<<(unsigned)blocks, gmc::kPcLanes * gmc::kPcRows, pc_smem, s>>>(
            rowptr, colidx, coef, graph_ptr, n_graphs, n_rows, n_cols, ncp, Xb, ldx, slow);
        GMC_LAUNCH_CHECK();
        only = slow;
    }
    // batches of small graphs (the caller states the largest one: no more nodes than feature columns; unit weights): one
    // CTA per graph with its CSR slice in shared memory.  A graph whose EDGES exceed the staging buffer (average degree
    // above 8) is walked from global memory by its CTA -- bounded work, since it has at most ncp rows.
    if (!only && !vals && n_graphs > 0 && max_nodes > 0 && max_nodes <= ncp && ncp <= 2048 && gmc::graph_kernel_enabled()) {
        const int cap_nodes = ncp, cap_edges = 8 * ncp + 64;
        const size_t g_smem = (size_t)gmc::kPgWarps * ncp * sizeof(float) + (size_t)(cap_nodes + 1) * 4 + (size_t)cap_edges * 2 + 16;
        static bool attr3 = false;
        if (!attr3) {
            GMC_CUDA(cudaFuncSetAttribute(gmc::preaggregate_graph_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            GMC_CUDA(cudaFuncSetAttribute(gmc::preaggregate_graph_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            attr3 = true;
        }
        int per_sm = (int)((220 * 1024) / (g_smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        int64_t blocks = n_graphs;
        const int64_t cap = (int64_t)gmc::sm_count() * per_sm;
        if (blocks > cap) blocks = cap;
        if (f16)
            gmc::preaggregate_graph_kernel<true><<<(unsigned)blocks, gmc::kPgWarps * 32, g_smem, s>>>(
                rowptr, colidx, coef, graph_ptr, n_graphs, n_cols, ncp, cap_nodes, cap_edges, Xb, ldx, pad_cols);
        else
            gmc::preaggregate_graph_kernel<false><<<(unsigned)blocks, gmc::kPgWarps * 32, g_smem, s>>>(
                rowptr, colidx, coef, graph_ptr, n_graphs, n_cols, ncp, cap_nodes, cap_edges, Xb, ldx, pad_cols);
        GMC_LAUNCH_CHECK();
        return GMC_OK;
    }
    const size_t smem = (size_t)gmc::kPreaggWarps * ncp * sizeof(float);
    static bool attr = false;
    if (!attr) {
        GMC_CUDA(cudaFuncSetAttribute(gmc::preaggregate_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 6144 * 4));
        GMC_CUDA(cudaFuncSetAttribute(gmc::preaggregate_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 6144 * 4));
        attr = true;
    }
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t blocks = gmc::ceil_div<int64_t>(n_rows, gmc::kPreaggWarps);
    const int64_t cap = (int64_t)gmc::sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (f16)
        gmc::preaggregate_bf16_kernel<true><<<(unsigned)blocks, gmc::kPreaggWarps * 32, smem, s>>>(
            rowptr, colidx, coef, vals, graph_ptr, n_graphs, n_rows, n_cols, ncp, Xb, ldx, only, pad_cols);
    else
        gmc::preaggregate_bf16_kernel<false><<<(unsigned)blocks, gmc::kPreaggWarps * 32, smem, s>>>(
            rowptr, colidx, coef, vals, graph_ptr, n_graphs, n_rows, n_cols, ncp, Xb, ldx, only, pad_cols);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_csr_preaggregate_bf16(const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                              const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X,
                              int64_t ldx, void* workspace, size_t workspace_bytes, void* stream) {
    return preaggregate_impl(rowptr, colidx, coef, vals, graph_ptr, n_graphs, n_rows, n_cols, X, ldx, workspace,
                             workspace_bytes, stream, false);
}

// the same rows written as IEEE fp16 (the 'f16x2' GEMM path needs both MMA operands in one 16-bit format)
int gmc_csr_preaggregate_f16(const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                             const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t n_cols, void* X,
                             int64_t ldx, void* stream) {
    return preaggregate_impl(rowptr, colidx, coef, vals, graph_ptr, n_graphs, n_rows, n_cols, X, ldx, nullptr, 0, stream, true);
}

// gmc_csr_preaggregate_bf16 / _f16 for a batch whose largest graph has max_nodes nodes (a host-side fact of every batch
// the callers build): batches of small graphs (max_nodes <= n_cols rounded up to 8, <= 2048) take the per-graph kernel
// with the graph's CSR slice in shared memory; everything else the row-parallel one.  out_f16 != 0 writes IEEE fp16.
int gmc_csr_preaggregate_graphs(const int32_t* rowptr, const int32_t* colidx, const float* coef, const float* vals,
                                const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, int64_t n_rows, int32_t n_cols,
                                void* X, int64_t ldx, int32_t out_f16, void* stream) {
    return preaggregate_impl(rowptr, colidx, coef, vals, graph_ptr, n_graphs, n_rows, n_cols, X, ldx, nullptr, 0, stream,
                             out_f16 != 0, max_nodes);
}

// dst[r, c] = bf16(src[r, c]), round to nearest even; leading dimensions in elements
int gmc_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t n_rows, int32_t n_cols, void* stream) {
    GMC_REQUIRE(src && dst, "gmc_f32_to_bf16: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols >= 0 && lds >= n_cols && ldd >= n_cols, "gmc_f32_to_bf16: bad sizes");
    if (n_rows == 0 || n_cols == 0) return GMC_OK;
    const int64_t total = n_rows * ((n_cols + 7) / 8);
    gmc::f32_to_bf16_kernel<<<(unsigned)gmc::ceil_div<int64_t>(total, 256), 256, 0, gmc::as_stream(stream)>>>(
        src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, n_rows, n_cols);
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

int gmc_copy2d_f32(float* dst, int64_t lddst, const float* src, int64_t ldsrc, int64_t n_rows, int32_t n_cols,
                   void* stream) {
    GMC_REQUIRE(dst && src, "gmc_copy2d_f32: null pointer");
    GMC_REQUIRE(n_rows >= 0 && n_cols >= 0 && lddst >= n_cols && ldsrc >= n_cols, "gmc_copy2d_f32: bad sizes");
    if (n_rows == 0 || n_cols == 0) return GMC_OK;
    GMC_CUDA(cudaMemcpy2DAsync(dst, (size_t)lddst * sizeof(float), src, (size_t)ldsrc * sizeof(float),
                               (size_t)n_cols * sizeof(float), (size_t)n_rows, cudaMemcpyDeviceToDevice,
                               gmc::as_stream(stream)));
    return GMC_OK;
}

}  // extern "C"
