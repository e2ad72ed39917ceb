// Raw tcgen05.mma kind::tf32 issue-rate probe: every SM loops over MMAs on one resident smem stage (no TMA, no
// epilogue).  Prints TFLOP/s for cta_group::1 128x256x8.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lay) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)lay << 61;
    return d;
}
template <int KIND>   // 0 = tf32 (K=8 per MMA), 1 = bf16 (K=16 per MMA)
__global__ void __launch_bounds__(128, 1) probe(int iters, unsigned long long* cycles) {
    extern __shared__ uint8_t raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(raw + (base - smem_u32(raw)))[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        // K-major SWIZZLE_128B operands: A 128 rows x 128 B, B 256 rows x 128 B
        const uint32_t idesc = (1u << 4) | ((KIND == 0 ? 2u : 1u) << 7) | ((KIND == 0 ? 2u : 1u) << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sa = base, sb = base + 16 * 1024;
        unsigned long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_desc(sa + k * 32, 16, 1024, 2), bd = make_desc(sb + k * 32, 16, 1024, 2);
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + (it & 1) * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + (it & 1) * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int KIND>
static void run(const char* name, int kk) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* d;
    cudaMalloc(&d, sms * 8);
    cudaFuncSetAttribute(probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        probe<KIND><<<sms, 128, 64 * 1024>>>(iters, d);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        unsigned long long c0 = 0;
        cudaMemcpy(&c0, d, 8, cudaMemcpyDeviceToHost);
        const double flops = 2.0 * 128 * 256 * kk * 4.0 * iters * sms;
        printf("%s: %d SMs, %.3f ms, %.1f TFLOP/s, %.1f cycles per 128x256x%d MMA (%s)\n", name, sms, ms, flops / ms / 1e9,
               (double)c0 / (4.0 * iters), kk, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    run<0>("tf32", 8);
    run<1>("bf16", 16);
    return 0;
}
