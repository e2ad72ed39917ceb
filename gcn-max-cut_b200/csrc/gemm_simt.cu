// gemm_simt.cu -- (b) dense feature transforms, exact-fp32 CUDA-core path (GMC_GEMM_FP32).
//
// This is the PARITY path: plain FFMA with fp32 accumulation, so logits agree with the
// reference's MKL sgemm to ~1e-6 and the discontinuous STE loss sees no label flips.
// The tensor-core path (tcgen05 kind::tf32, gemm_tcgen05.cu) is the throughput path.
//
//   nn: C = A[M,K] * B[K,N]       X * W1             (GraphConv th.matmul, TrainingNeural.py:80)
//   nt: C = A[M,K] * B[N,K]^T     dT1 * W1^T         (only when features are trainable)
//   tn: C = A[K,M]^T * B[K,N]     X^T * dT1 = dW1    (K = all nodes of the batch -> split-K,
//                                                     deterministic two-stage reduction)
//
// 128x128x8 CTA tile, 256 threads, 8x8 register tile per thread, double-buffered shared
// memory with register prefetch; global loads are 128-bit when rows are 16-byte aligned.
#include "common.cuh"

namespace gmc {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;

struct Frag4 { float v[4]; };

template <bool KMAJOR>   // KMAJOR: element(r, k) at P[r*ld + k]; else at P[k*ld + r]     (r = m or n)
__device__ __forceinline__ Frag4 load_tile_frag(const float* __restrict__ P, int64_t ld, int64_t r0, int64_t R,
                                                int64_t kt, int64_t kend, int tid, bool vec) {
    Frag4 f;
    f.v[0] = f.v[1] = f.v[2] = f.v[3] = 0.f;
    if (KMAJOR) {
        const int64_t r = r0 + (tid >> 1);
        const int64_t k = kt + ((tid & 1) << 2);
        if (r < R) {
            const float* p = P + r * ld + k;
            if (vec && k + 3 < kend) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (k + i < kend) f.v[i] = __ldg(p + i);
            }
        }
    } else {
        const int64_t k = kt + (tid >> 5);
        const int64_t r = r0 + ((tid & 31) << 2);
        if (k < kend) {
            const float* p = P + k * ld + r;
            if (vec && r + 3 < R) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (r + i < R) f.v[i] = __ldg(p + i);
            }
        }
    }
    return f;
}

template <bool KMAJOR>
__device__ __forceinline__ void store_tile_frag(float (*S)[BM + PAD], const Frag4& f, int tid) {
    if (KMAJOR) {
        const int r = tid >> 1, k = (tid & 1) << 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) S[k + i][r] = f.v[i];
    } else {
        const int k = tid >> 5, r = (tid & 31) << 2;
        *reinterpret_cast<float4*>(&S[k][r]) = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
    }
}

template <bool A_KMAJOR, bool B_NMAJOR>
__global__ void __launch_bounds__(256, 2)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int64_t M, int64_t N,
             int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t k_per_split, int64_t split_stride,
             int accumulate, int vecA, int vecB, int vecC) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n0 = (int64_t)blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = min(K, kbeg + k_per_split);
    C += (int64_t)blockIdx.z * split_stride;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    Frag4 fa = load_tile_frag<A_KMAJOR>(A, lda, m0, M, kbeg, kend, tid, vecA);
    Frag4 fb = load_tile_frag<!B_NMAJOR>(B, ldb, n0, N, kbeg, kend, tid, vecB);
    store_tile_frag<A_KMAJOR>(As[0], fa, tid);
    store_tile_frag<!B_NMAJOR>(Bs[0], fb, tid);
    __syncthreads();

    int buf = 0;
    for (int64_t kt = kbeg; kt < kend; kt += BK) {
        const bool has_next = kt + BK < kend;
        if (has_next) {
            fa = load_tile_frag<A_KMAJOR>(A, lda, m0, M, kt + BK, kend, tid, vecA);
            fb = load_tile_frag<!B_NMAJOR>(B, ldb, n0, N, kt + BK, kend, tid, vecB);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) {
            store_tile_frag<A_KMAJOR>(As[buf ^ 1], fa, tid);
            store_tile_frag<!B_NMAJOR>(Bs[buf ^ 1], fb, tid);
            __syncthreads();
            buf ^= 1;
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t n = n0 + (h ? 64 + tx * 4 : tx * 4);
            float* c = C + m * ldc + n;
            float4 r = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            if (vecC && n + 3 < N) {
                if (accumulate) {
                    const float4 o = *reinterpret_cast<const float4*>(c);
                    r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
                }
                *reinterpret_cast<float4*>(c) = r;
            } else {
                const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) c[j] = accumulate ? c[j] + rr[j] : rr[j];
            }
        }
    }
}

// C[m,n] (+)= sum_z ws[z][m][n], fixed summation order -> deterministic
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t MN, int64_t N, float* __restrict__ C,
                                     int64_t ldc, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * MN + i];
    float* c = C + (i / N) * ldc + (i % N);
    *c = accumulate ? *c + s : s;
}

static int pick_splits(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ceil_div<int64_t>(M, BM) * ceil_div<int64_t>(N, BN);
    const int64_t target = (int64_t)sm_count() * 4;
    if (tiles >= target || K <= 4096) return 1;
    int64_t s = ceil_div<int64_t>(target, tiles);
    const int64_t max_by_k = K / 2048 > 0 ? K / 2048 : 1;
    if (s > max_by_k) s = max_by_k;
    if (s > 64) s = 64;
    return (int)(s < 1 ? 1 : s);
}

template <bool A_KMAJOR, bool B_NMAJOR>
static int sgemm_launch(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                        int64_t ldb, int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes,
                        cudaStream_t s) {
    if (M == 0 || N == 0) return GMC_OK;
    const int vecA = (lda % 4 == 0) && aligned16(A);
    const int vecB = (ldb % 4 == 0) && aligned16(B);
    int splits = pick_splits(M, N, K);
    if (splits > 1 && (!workspace || workspace_bytes < (size_t)splits * M * N * sizeof(float))) {
        // fall back to fewer splits that fit (1 = no workspace needed)
        splits = workspace ? (int)(workspace_bytes / ((size_t)M * N * sizeof(float))) : 1;
        if (splits < 1) splits = 1;
    }
    int64_t k_per = ceil_div<int64_t>(K, splits);
    k_per = ceil_div<int64_t>(k_per, BK) * BK;
    if (k_per < BK) k_per = BK;
    splits = (int)ceil_div<int64_t>(K > 0 ? K : 1, k_per);
    dim3 grid((unsigned)ceil_div<int64_t>(M, BM), (unsigned)ceil_div<int64_t>(N, BN), (unsigned)splits);
    GMC_REQUIRE(grid.y <= 65535, "gmc_gemm: N too large for this kernel");
    if (splits == 1) {
        const int vecC = (ldc % 4 == 0) && aligned16(C);
        sgemm_kernel<A_KMAJOR, B_NMAJOR><<<grid, 256, 0, s>>>(A, B, C, M, N, K, lda, ldb, ldc, k_per, 0, accumulate,
                                                              vecA, vecB, vecC);
        GMC_LAUNCH_CHECK();
    } else {
        float* ws = reinterpret_cast<float*>(workspace);
        const int vecC = (N % 4 == 0) && aligned16(ws);
        sgemm_kernel<A_KMAJOR, B_NMAJOR><<<grid, 256, 0, s>>>(A, B, ws, M, N, K, lda, ldb, N, k_per, M * N, 0, vecA,
                                                              vecB, vecC);
        GMC_LAUNCH_CHECK();
        const int64_t MN = M * N;
        splitk_reduce_kernel<<<(unsigned)ceil_div<int64_t>(MN, 256), 256, 0, s>>>(ws, splits, MN, N, C, ldc, accumulate);
        GMC_LAUNCH_CHECK();
    }
    return GMC_OK;
}

size_t simt_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    const int splits = pick_splits(M, N, K);
    return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}

int simt_gemm(int op, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
              int64_t ldb, int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    switch (op) {
        case 0: return sgemm_launch<true, true>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
        case 1: return sgemm_launch<true, false>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
        case 2: return sgemm_launch<false, true>(A, B, C, M, N, K, lda, ldb, ldc, accumulate, workspace, workspace_bytes, s);
    }
    set_error("gmc_gemm: bad op %d", op);
    return GMC_ERR_INVALID_ARG;
}

}  // namespace gmc
