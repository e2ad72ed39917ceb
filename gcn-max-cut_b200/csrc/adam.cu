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
constexpr int kAdamChunk = 1024 * 4;       // elements per CTA (256 threads x 4 x float4)

struct AdamArgs {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    int64_t n[kAdamMaxTensors];
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

__global__ void __launch_bounds__(256)
adam_kernel(AdamArgs a, AdamScalars c, double lr, double b1, double b2, int64_t* step_dev) {
    if (step_dev) {                                      // graph-replayable variant: t lives on the device
        __shared__ float sc[2];
        if (threadIdx.x == 0) {
            const double t = (double)(*step_dev + 1);
            const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
            sc[0] = (float)(lr / bc1);
            sc[1] = (float)sqrt(bc2);
        }
        __syncthreads();
        c.step_size = sc[0];
        c.bc2_sqrt = sc[1];
    }
    int t = 0;
    while (t + 1 < a.n_tensors && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
    const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * kAdamChunk;
    const int64_t n = a.n[t];
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
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
        } else {
            for (int64_t j = i; j < n && j < i + 4; ++j) {
                float pp = p[j], mm = m[j], vv = v[j];
                adam_elem(pp, g[j], mm, vv, c);
                p[j] = pp; m[j] = mm; v[j] = vv;
            }
        }
    }
}

__global__ void adam_step_inc_kernel(int64_t* step_dev) { *step_dev += 1; }

static int adam_launch(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const int64_t* sizes, double lr, double beta1, double beta2,
                       double eps, int64_t step, int64_t* step_dev, void* stream) {
    GMC_REQUIRE(n_tensors >= 0 && n_tensors <= kAdamMaxTensors, "gmc_adam_multi: n_tensors must be 0..%d", kAdamMaxTensors);
    GMC_REQUIRE(n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && sizes), "gmc_adam_multi: null array");
    GMC_REQUIRE(step_dev || step >= 1, "gmc_adam_multi: step is 1-based");
    AdamArgs a;
    a.n_tensors = 0;
    int blocks = 0;
    for (int t = 0; t < n_tensors; ++t) {
        GMC_REQUIRE(sizes[t] >= 0, "gmc_adam_multi: negative size");
        if (sizes[t] == 0) continue;
        GMC_REQUIRE(params[t] && grads[t] && exp_avg[t] && exp_avg_sq[t], "gmc_adam_multi: null tensor %d", t);
        const int k = a.n_tensors++;
        a.p[k] = params[t]; a.g[k] = grads[t]; a.m[k] = exp_avg[t]; a.v[k] = exp_avg_sq[t]; a.n[k] = sizes[t];
        a.block_start[k] = blocks;
        blocks += (int)ceil_div<int64_t>(sizes[t], kAdamChunk);
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
        adam_kernel<<<blocks, 256, 0, s>>>(a, c, lr, beta1, beta2, step_dev);
        GMC_LAUNCH_CHECK();
    }
    if (step_dev) {
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

}  // extern "C"
