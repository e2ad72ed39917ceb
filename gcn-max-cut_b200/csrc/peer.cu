// peer.cu -- the data-parallel gradient exchange as ONE kernel over NVLink peer memory.
//
// The only collective of the training step is the all-reduce(sum) of the flat weight-gradient buffer [dW1|db1|dW2|db2]
// (502 003 floats = 2.0 MB at F = 1000, H = 500, K = 3; the reference has no distributed code: the step being exchanged
// is python/Training/TrainingNeural.py:380-386, loss.backward() ... optimizer.step()).  At 2 MB the collective is pure
// latency.  Every rank keeps its gradient buffer in memory the other ranks of the node have mapped (cudaIpc handles
// exchanged once through torch.distributed), and the exchange is a
// one-shot reduce-scatter + all-gather:
//   barrier 1   every rank's gradients are complete (flags written with release.sys after the producing kernels)
//   reduce      rank r adds slice r of all W buffers in rank order 0..W-1 (peer loads over NVLink, L1 bypassed) and
//               writes the sum into slice r of ALL W buffers (slices are disjoint: no rank reads what another writes)
//   barrier 2   every slice of my buffer has arrived -> Adam may read it
// The sum has one fixed order for every element, so all ranks hold bit-identical gradients -- and therefore weights --
// after every step (bench.py dp_check.weights_identical_on_all_ranks).
// Measured on 8 B200s, back to back (scratch/ar_probe.py): 36.6 us per call; NCCL 2.28's all-reduce of the same buffer: 32.4 us.
// The exchange is therefore NOT what data parallelism costs here (the 0.2-0.6 ms "all-reduce interval" of a step is the
// wait for that step's slowest rank), and this path is opt-in (GMC_PEER_ALLREDUCE=1); NCCL stays the default.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gmc {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 512;
constexpr int kPeerFlagWords = 64;        // per rank: words [0, 16) barrier 1, [16, 32) barrier 2 (one per peer), word 32 = local block counter

struct PeerPtrs {
    float4* buf[kPeerMaxWorld];
    uint32_t* flag[kPeerMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float4* p) {              // not cached in L1: the line is rewritten remotely
    float4 r;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ bool reached(uint32_t v, uint32_t epoch) { return (int32_t)(v - epoch) >= 0; }

template <int W>
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_kernel(PeerPtrs p, int world, int rank, int64_t n4, uint32_t epoch) {
    const int w_n = W > 0 ? W : world;
    uint32_t* my_flags = p.flag[rank];
    // ---- barrier 1: the producing kernels of this rank finished before this launch (stream order); tell everyone, then
    // wait until everyone has told me.  Every block waits for itself (no grid-wide sync needed before the loads).
    if (blockIdx.x == 0 && (int)threadIdx.x < w_n) st_release_sys(p.flag[threadIdx.x] + rank, epoch);
    if ((int)threadIdx.x < w_n)
        while (!reached(ld_acquire_sys(my_flags + threadIdx.x), epoch)) __nanosleep(20);
    __syncthreads();

    // ---- my slice: sum in rank order, result to every rank
    const int64_t per = (n4 + w_n - 1) / w_n;
    const int64_t lo = (int64_t)rank * per, hi = min(n4, lo + per);
    for (int64_t i = lo + (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i < hi; i += (int64_t)gridDim.x * kPeerThreads) {
        float4 v[W > 0 ? W : kPeerMaxWorld];
#pragma unroll
        for (int q = 0; q < (W > 0 ? W : kPeerMaxWorld); ++q)
            if (q < w_n) v[q] = ld_peer4(p.buf[q] + i);
        float4 s = v[0];
#pragma unroll
        for (int q = 1; q < (W > 0 ? W : kPeerMaxWorld); ++q)
            if (q < w_n) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
#pragma unroll
        for (int q = 0; q < (W > 0 ? W : kPeerMaxWorld); ++q)
            if (q < w_n) p.buf[q][i] = s;
    }

    // ---- barrier 2: the LAST block of this rank to finish its stores tells everyone; everyone waits for all W ranks
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* counter = my_flags + 2 * kPeerMaxWorld;
        if (atomicAdd(counter, 1u) == gridDim.x - 1) {
            *counter = 0;                                              // ready for the next launch (stream-ordered)
            __threadfence_system();
            for (int q = 0; q < w_n; ++q) st_release_sys(p.flag[q] + kPeerMaxWorld + rank, epoch);
        }
    }
    if ((int)threadIdx.x < w_n)
        while (!reached(ld_acquire_sys(my_flags + kPeerMaxWorld + threadIdx.x), epoch)) __nanosleep(20);
    __syncthreads();
}

}  // namespace gmc

extern "C" {

// Device memory other processes of the node can map: plain cudaMalloc (legacy cudaIpc handles do not cover the stream-
// ordered / virtual-memory pools a framework allocator may use), zero-filled.
int gmc_peer_alloc(size_t bytes, void** ptr) {
    GMC_REQUIRE(ptr && bytes > 0, "gmc_peer_alloc: bad arguments");
    GMC_CUDA(cudaMalloc(ptr, bytes));
    GMC_CUDA(cudaMemset(*ptr, 0, bytes));
    return GMC_OK;
}

int gmc_peer_free(void* ptr) {
    if (ptr) GMC_CUDA(cudaFree(ptr));
    return GMC_OK;
}

size_t gmc_peer_flag_bytes(void) { return (size_t)gmc::kPeerFlagWords * sizeof(uint32_t); }
int32_t gmc_ipc_handle_bytes(void) { return (int32_t)sizeof(cudaIpcMemHandle_t); }

int gmc_ipc_get_handle(const void* ptr, void* handle) {
    GMC_REQUIRE(ptr && handle, "gmc_ipc_get_handle: null pointer");
    cudaIpcMemHandle_t h;
    GMC_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle, &h, sizeof(h));
    return GMC_OK;
}

int gmc_ipc_open_handle(const void* handle, void** ptr) {
    GMC_REQUIRE(ptr && handle, "gmc_ipc_open_handle: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    GMC_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GMC_OK;
}

int gmc_ipc_close_handle(void* ptr) {
    if (ptr) GMC_CUDA(cudaIpcCloseMemHandle(ptr));
    return GMC_OK;
}

// In-place all-reduce(sum) of n floats held at the same offset of `world` peer-mapped buffers.  bufs / flags: HOST arrays of
// `world` device pointers as seen from THIS process (entry `rank` is the local allocation); every buffer holds n rounded up
// to 4 floats, every flag block gmc_peer_flag_bytes() zero-initialised bytes.  All ranks call with the same `epoch`, which
// must grow by one per call.  Blocks until nothing: the kernel itself waits for the peers on `stream`.
int gmc_peer_allreduce_f32(void* const* bufs, void* const* flags, int32_t world, int32_t rank, int64_t n, uint32_t epoch,
                           void* stream) {
    using namespace gmc;
    GMC_REQUIRE(bufs && flags, "gmc_peer_allreduce_f32: null pointer");
    GMC_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world && n >= 0,
                "gmc_peer_allreduce_f32: world must be 1..%d, 0 <= rank < world", kPeerMaxWorld);
    if (n == 0 || world == 1) return GMC_OK;
    PeerPtrs p;
    for (int q = 0; q < world; ++q) {
        GMC_REQUIRE(bufs[q] && flags[q] && aligned16(bufs[q]), "gmc_peer_allreduce_f32: null / unaligned peer pointer %d", q);
        p.buf[q] = reinterpret_cast<float4*>(bufs[q]);
        p.flag[q] = reinterpret_cast<uint32_t*>(flags[q]);
    }
    const int64_t n4 = (n + 3) / 4;
    const int64_t per = (n4 + world - 1) / world;
    int blocks = (int)ceil_div<int64_t>(per, kPeerThreads);
    if (blocks > 64) blocks = 64;                                      // all blocks must be co-resident (they spin)
    if (blocks < 1) blocks = 1;
    cudaStream_t s = as_stream(stream);
    switch (world) {
        case 2: peer_allreduce_kernel<2><<<blocks, kPeerThreads, 0, s>>>(p, world, rank, n4, epoch); break;
        case 4: peer_allreduce_kernel<4><<<blocks, kPeerThreads, 0, s>>>(p, world, rank, n4, epoch); break;
        case 8: peer_allreduce_kernel<8><<<blocks, kPeerThreads, 0, s>>>(p, world, rank, n4, epoch); break;
        default: peer_allreduce_kernel<0><<<blocks, kPeerThreads, 0, s>>>(p, world, rank, n4, epoch); break;
    }
    GMC_LAUNCH_CHECK();
    return GMC_OK;
}

}  // extern "C"
