// api_misc.cu -- version / error / device queries of the C ABI (include/gcnmaxcut.h).
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace gmc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static int cached = -1;
    // opt-in: measured on config 1 (500-node graphs, 9 launches per step) 49.7 us per step with, 50.7 without --
    // the step is bound by the kernels' own few-microsecond durations, not by the gaps between them
    if (cached < 0) { const char* e = getenv("GMC_PDL"); cached = (e && e[0] == '1') ? 1 : 0; }
    return cached == 1;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = kNumSMsB200;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsB200;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace gmc

extern "C" {

int gmc_abi_version(void) { return GMC_ABI_VERSION; }

const char* gmc_last_error(void) { return gmc::g_err; }

int gmc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    GMC_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) { GMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (cc_major) { GMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { GMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return GMC_OK;
}

}  // extern "C"
