// common.cuh -- shared helpers for libgcnmaxcut (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gcnmaxcut.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgcnmaxcut targets sm_100a (B200) only"
#endif

namespace gmc {

constexpr int kWarp = 32;
constexpr int kMaxClasses = 8;
constexpr int kNumSMsB200 = 148;

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();                      // cached cudaDevAttrMultiProcessorCount of the current device

#define GMC_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            gmc::set_error(__VA_ARGS__);               \
            return GMC_ERR_INVALID_ARG;                \
        }                                              \
    } while (0)

#define GMC_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            gmc::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                           cudaGetErrorString(e__));                                    \
            return (int)e__;                                                            \
        }                                                                               \
    } while (0)

#define GMC_LAUNCH_CHECK() GMC_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch ------------------------------------------------------
// The reference trains one small graph per optimiser step (TrainingNeural.py:371-388): ~10 dependent launches of a few
// microseconds each, so the gaps between kernels are a third of the step.  Kernels on that chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and begin with pdl_prologue(): `launch_dependents` lets the NEXT
// kernel's CTAs be scheduled (and run their own prologue) as soon as every CTA of this grid is resident, `wait` blocks
// until the PREVIOUS grid has completed and flushed -- executed by every thread before its first global access, so
// completion stays transitive along the chain (C waits for B, which waited for A).  A kernel launched without the
// attribute simply waits for full completion, so converted and unconverted kernels mix freely.  Opt-in (GMC_PDL=1):
// without the attribute the device-side instructions are no-ops.
bool pdl_enabled();

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
__host__ __device__ constexpr inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// wait FIRST: a kernel that triggers before it waits lets the whole chain behind it become resident at once (each early
// CTA triggers its own dependents), and the spinning CTAs keep the 225 KB GEMM CTAs off their SMs -- measured 59 us per
// 500-node step against 49 us without early launches
__device__ __forceinline__ void pdl_prologue() { pdl_wait(); pdl_launch_dependents(); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_max(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// Sums P (power of two <= 32) per-lane values over the warp in P - 1 + log2(32 / P) shuffles instead of 5 P: every
// halving step exchanges the half a lane does not keep.  Returns the total of value index (lane >> log2(32 / P)),
// valid in every lane (lanes that share an index hold the same total).
template <int P>
__device__ __forceinline__ float warp_multi_sum(float (&v)[P], int lane) {
    int m = 16;
#pragma unroll
    for (int n = P; n > 1; n >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
#pragma unroll
    for (; m > 0; m >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
    return v[0];
}

// graph id of node v: largest g with graph_ptr[g] <= v   (graph_ptr has n_graphs+1 entries)
__device__ __forceinline__ int find_graph(const int32_t* __restrict__ graph_ptr, int n_graphs, int64_t v) {
    int lo = 0, hi = n_graphs;          // invariant: graph_ptr[lo] <= v < graph_ptr[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(graph_ptr + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// streaming 128-bit load that does not pollute L1 (data touched once per CTA)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void fma4(float4& acc, float a, const float4& x) {
    acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y);
    acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
}

#endif  // __CUDACC__

}  // namespace gmc
