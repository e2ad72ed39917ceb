// adam.cu -- (d) fused multi-tensor Adam, one launch for all parameter tensors.
//
// Replaces optimizer.step() (reference python/Training/TrainingNeural.py:386; optimizer built at
// :336-337 as torch.optim.Adam(params, lr) with defaults betas=(0.9,0.999), eps=1e-8,
// weight_decay=0, amsgrad=False).  Formula and operation order follow torch's single-tensor
// Adam so results agree to fp32 rounding:
//     m.lerp_(g, 1-b1);  v.mul_(b2).addcmul_(g, g, 1-b2)
//     denom = sqrt(v)/sqrt(1-b2^t) + eps;  p.addcdiv_(m, denom, -lr/(1-b1^t))
// HBM-bound: 28 B per parameter per step (read p,g,m,v; write p,m,v).
#include <math.h>

#include "common.cuh"

namespace gmc {

constexpr int kAdamMaxTensors = 16;
constexpr int kAdamChunk = 1024 * 4;       // elements per CTA (256 threads x 4 x float4): large tensors (embedding tables)
constexpr int kAdamChunkSmall = 1024;      // one float4 per thread: the 502 003 weights of the reference model are latency-
                                           // bound (126 CTAs x 4 dependent round trips = 10.7 us; 491 CTAs x 1 = one trip)
constexpr int64_t kAdamSmallTotal = 1 << 22;

struct AdamArgs {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    int64_t n[kAdamMaxTensors];
    uint16_t* shadow[kAdamMaxTensors];     // nullable: bf16 copy of the updated parameter, same element index (the operand of
                                           // next step's bf16 GEMMs: learned node embeddings, TrainingNeural.py:332 / utils.py:184)
    int block_start[kAdamMaxTensors + 1];  // first CTA of each tensor
    int n_tensors;
};

struct AdamScalars { float one_minus_b1, b2, one_minus_b2, eps, step_size, bc2_sqrt; };

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamScalars& c) {
    m = fmaf(g - m, c.one_minus_b1, m);                 // lerp, weight < 0.5 branch
    v = fmaf(c.one_minus_b2 * g, g, v * c.b2);          // v*b2 + (1-b2)*g*g
    const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
    p = p - c.step_size * (m / denom);
}

// Device-side step state of the graph-replayable variants.
//   plain (gmc_adam_multi_devstep): state[0] = steps taken; a one-thread kernel increments it after the update.
//   cached (gmc_adam_multi_devstate, 8 x int64): state[0] = steps taken, state[1] = the step number whose bias-correction
//   scalars state[2] holds (step_size | bc2_sqrt as float bits) for the (lr, beta1, beta2) in state[4..6], state[3] = CTA
//   ticket.  One extra CTA per launch advances state[0] and computes the NEXT step's scalars while the others work, so
//   the double-precision pow() pair is off the critical path of every CTA and the increment needs no launch of its own
//   (a 500-node graph's step is ~10 launches of a few microseconds each: TrainingNeural.py:371-388).
constexpr int kAdamStateWords = 8;

__device__ __forceinline__ void adam_bias_scalars(double lr, double b1, double b2, double t, float& step_size, float& bc2_sqrt) {
    const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
    step_size = (float)(lr / bc1);
    bc2_sqrt = (float)sqrt(bc2);
}

template <int ITER>
__global__ void __launch_bounds__(256)
adam_kernel(AdamArgs a, AdamScalars c, double lr, double b1, double b2, int64_t* step_dev, int cached) {
    pdl_prologue();
    if (step_dev) {                                      // graph-replayable variants: t lives on the device
        __shared__ float sc[2];
        if (threadIdx.x == 0) {
            const int64_t t = step_dev[0] + 1;
            if (cached && step_dev[1] == t && step_dev[4] == __double_as_longlong(lr) &&
                step_dev[5] == __double_as_longlong(b1) && step_dev[6] == __double_as_longlong(b2)) {
                const unsigned long long pk = (unsigned long long)step_dev[2];
                sc[0] = __uint_as_float((uint32_t)pk);
                sc[1] = __uint_as_float((uint32_t)(pk >> 32));
            } else {
                adam_bias_scalars(lr, b1, b2, (double)t, sc[0], sc[1]);
            }
        }
        __syncthreads();
        c.step_size = sc[0];
        c.bc2_sqrt = sc[1];
    }
    if (cached) {
        // every working CTA has read state[0..6] above before it takes its ticket.  The one extra CTA (the last block
        // index, so everything it waits for was dispatched before it) computes the next step's scalars meanwhile and is
        // the only writer, once all tickets are in.
        const unsigned int n_work = gridDim.x - 1;
        if (blockIdx.x < n_work) {
            if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(step_dev + 3), 1ull);
        } else {
            if (threadIdx.x == 0) {
                const int64_t t = step_dev[0] + 1;       // the step this launch takes
                float ss, bs;
                adam_bias_scalars(lr, b1, b2, (double)(t + 1), ss, bs);
                volatile unsigned long long* ticket = reinterpret_cast<volatile unsigned long long*>(step_dev + 3);
                while (*ticket < (unsigned long long)n_work) __nanosleep(64);
                step_dev[2] = (int64_t)((unsigned long long)__float_as_uint(ss) | ((unsigned long long)__float_as_uint(bs) << 32));
                step_dev[4] = __double_as_longlong(lr);
                step_dev[5] = __double_as_longlong(b1);
                step_dev[6] = __double_as_longlong(b2);
                step_dev[1] = t + 1;
                step_dev[0] = t;
                step_dev[3] = 0;
            }
            return;
        }
    }
    int t = 0;
    while (t + 1 < a.n_tensors && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
    const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * (ITER * 1024);
    const int64_t n = a.n[t];
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    uint16_t* __restrict__ sh = a.shadow[t];
    const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0 && (((uintptr_t)sh) & 7u) == 0;
#pragma unroll
    for (int r = 0; r < ITER; ++r) {
        const int64_t i = base + ((int64_t)r * 256 + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec && i + 3 < n) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            adam_elem(pp.x, gg.x, mm.x, vv.x, c); adam_elem(pp.y, gg.y, mm.y, vv.y, c);
            adam_elem(pp.z, gg.z, mm.z, vv.z, c); adam_elem(pp.w, gg.w, mm.w, vv.w, c);
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
            if (sh) {
                uint2 pk;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(pp.y), "f"(pp.x));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(pp.w), "f"(pp.z));
                *reinterpret_cast<uint2*>(sh + i) = pk;
            }
        } else {
            for (int64_t j = i; j < n && j < i + 4; ++j) {
                float pp = p[j], mm = m[j], vv = v[j];
                adam_elem(pp, g[j], mm, vv, c);
                p[j] = pp; m[j] = mm; v[j] = vv;
                if (sh) {
                    uint32_t w;
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(0.f), "f"(pp));
                    sh[j] = (uint16_t)(w & 0xffffu);
                }
            }
        }
    }
}

__global__ void adam_step_inc_kernel(int64_t* step_dev) { *step_dev += 1; }

static int adam_launch(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const int64_t* sizes, double lr, double beta1, double beta2,
                       double eps, int64_t step, int64_t* step_dev, void* stream, int cached = 0,
                       void* const* shadows = nullptr) {
    GMC_REQUIRE(n_tensors >= 0 && n_tensors <= kAdamMaxTensors, "gmc_adam_multi: n_tensors must be 0..%d", kAdamMaxTensors);
    GMC_REQUIRE(n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && sizes), "gmc_adam_multi: null array");
    GMC_REQUIRE(step_dev || step >= 1, "gmc_adam_multi: step is 1-based");
    AdamArgs a;
    a.n_tensors = 0;
    int blocks = 0;
    int64_t total = 0;
    for (int t = 0; t < n_tensors; ++t) total += sizes[t] > 0 ? sizes[t] : 0;
    const bool small = total <= kAdamSmallTotal;
    const int chunk = small ? kAdamChunkSmall : kAdamChunk;
    for (int t = 0; t < n_tensors; ++t) {
        GMC_REQUIRE(sizes[t] >= 0, "gmc_adam_multi: negative size");
        if (sizes[t] == 0) continue;
        GMC_REQUIRE(params[t] && grads[t] && exp_avg[t] && exp_avg_sq[t], "gmc_adam_multi: null tensor %d", t);
        const int k = a.n_tensors++;
        a.p[k] = params[t]; a.g[k] = grads[t]; a.m[k] = exp_avg[t]; a.v[k] = exp_avg_sq[t]; a.n[k] = sizes[t];
        a.shadow[k] = shadows ? reinterpret_cast<uint16_t*>(shadows[t]) : nullptr;
        a.block_start[k] = blocks;
        blocks += (int)ceil_div<int64_t>(sizes[t], chunk);
    }
    a.block_start[a.n_tensors] = blocks;
    AdamScalars c;
    c.one_minus_b1 = (float)(1.0 - beta1);
    c.b2 = (float)beta2;
    c.one_minus_b2 = (float)(1.0 - beta2);
    c.eps = (float)eps;
    c.step_size = 0.f; c.bc2_sqrt = 1.f;
    if (!step_dev) {
        const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
        c.step_size = (float)(lr / bc1);
        c.bc2_sqrt = (float)sqrt(bc2);
    }
    cudaStream_t s = as_stream(stream);
    if (blocks > 0) {
        if (small) GMC_CUDA(launch_pdl(adam_kernel<1>, blocks + (cached ? 1 : 0), 256, 0, s, a, c, lr, beta1, beta2, step_dev, cached));
        else GMC_CUDA(launch_pdl(adam_kernel<4>, blocks + (cached ? 1 : 0), 256, 0, s, a, c, lr, beta1, beta2, step_dev, cached));
    }
    if (step_dev && (!cached || blocks == 0)) {
        adam_step_inc_kernel<<<1, 1, 0, s>>>(step_dev);
        GMC_LAUNCH_CHECK();
    }
    return GMC_OK;
}

}  // namespace gmc

extern "C" {

int gmc_adam_multi(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* sizes, double lr, double beta1, double beta2, double eps,
                   int64_t step, void* stream) {
    return gmc::adam_launch(n_tensors, params, grads, exp_avg, exp_avg_sq, sizes, lr, beta1, beta2, eps, step, nullptr,
                            stream);
}

int gmc_adam_multi_devstep(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                           float* const* exp_avg_sq, const int64_t* sizes, double lr, double beta1, double beta2,
                           double eps, int64_t* step_dev, void* stream) {
    GMC_REQUIRE(step_dev, "gmc_adam_multi_devstep: null step pointer");
    return gmc::adam_launch(n_tensors, params, grads, exp_avg, exp_avg_sq, sizes, lr, beta1, beta2, eps, 0, step_dev,
                            stream);
}


// gmc_adam_multi that also leaves a bf16 copy of every updated parameter in shadows[t] (nullable per tensor; same element
// index as the parameter): the next step's bf16 GEMM operand without a separate 6-bytes-per-element conversion pass.
int gmc_adam_multi_shadow(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                          float* const* exp_avg_sq, void* const* shadows, const int64_t* sizes, double lr, double beta1,
                          double beta2, double eps, int64_t step, void* stream) {
    GMC_REQUIRE(n_tensors == 0 || shadows, "gmc_adam_multi_shadow: null shadow array");
    return gmc::adam_launch(n_tensors, params, grads, exp_avg, exp_avg_sq, sizes, lr, beta1, beta2, eps, step, nullptr,
                            stream, 0, shadows);
}

// step state of kAdamStateWords (8) int64 words, zero-initialised by the caller (word 0 may be preset to the number of
// steps already taken); no separate increment launch, bias-correction scalars precomputed by the previous launch.
int gmc_adam_multi_devstate(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                            float* const* exp_avg_sq, const int64_t* sizes, double lr, double beta1, double beta2,
                            double eps, int64_t* state, void* stream) {
    GMC_REQUIRE(state, "gmc_adam_multi_devstate: null state pointer");
    return gmc::adam_launch(n_tensors, params, grads, exp_avg, exp_avg_sq, sizes, lr, beta1, beta2, eps, 0, state,
                            stream, 1);
}

}  // extern "C"
