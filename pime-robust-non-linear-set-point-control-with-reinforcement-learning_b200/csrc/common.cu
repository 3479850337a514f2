// common.cu -- error plumbing and device discovery for libpime_b200.
#include <cstdio>

#include "pime_common.cuh"

namespace pime {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    g_last_error = buf;
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? PIME_ENODEV : PIME_ECUDA;
}

// There is no CPU fallback: a compute entry point called without an sm_100 device fails loudly.
int require_device() {
    static thread_local int cached = 1;  // 1 = unknown
    if (cached <= 0) {
        if (cached < 0) g_last_error = "no CUDA device with compute capability 10.x is available (libpime_b200 has no CPU fallback)";
        return cached;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        cached = PIME_ENODEV;
        g_last_error = "no CUDA device with compute capability 10.x is available (libpime_b200 has no CPU fallback)";
        return cached;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        cached = PIME_ENODEV;
        g_last_error = "libpime_b200 is built for sm_100a only; current device is not compute capability 10.x";
        return cached;
    }
    cached = PIME_OK;
    return cached;
}

int device_sm_count() {
    static thread_local int cached[64] = {0};
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
    if (cached[dev] == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMs;
        cached[dev] = v;
    }
    return cached[dev];
}

}  // namespace pime

extern "C" {

int pime_abi_version(void) { return PIME_B200_ABI_VERSION; }

const char *pime_last_error(void) { return pime::g_last_error.c_str(); }

int pime_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        pime::set_error("no CUDA device");
        return PIME_ENODEV;
    }
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (sm_count) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sm_count = v; }
    if (cc_major) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev); *cc_major = v; }
    if (cc_minor) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev); *cc_minor = v; }
    return PIME_OK;
}

}  // extern "C"
