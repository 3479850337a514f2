// pime_common.cuh -- shared device/host helpers for libpime_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <string>

#include "../../include/pime_b200.h"

namespace pime {

// --------------------------------------------------------------------------------------------- errors
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
int require_device();  // PIME_OK or PIME_ENODEV (message set)
int device_sm_count();  // multiprocessors of the current device (148 on B200); kNumSMs when the query fails

#define PIME_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return ::pime::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define PIME_REQUIRE(cond, msg)                                                \
    do {                                                                       \
        if (!(cond)) {                                                         \
            ::pime::set_error(std::string("invalid argument: ") + (msg));      \
            return PIME_EINVAL;                                                \
        }                                                                      \
    } while (0)

#define PIME_LAUNCH_CHECK() PIME_CUDA(cudaGetLastError())

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline int grid_for(int64_t n, int block, int waves_cap = 16) {
    int64_t want = (n + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * waves_cap;
    if (want <= cap) return (int)(want < 1 ? 1 : want);
    return (int)cap;  // grid-stride loop covers the rest
}

// --------------------------------------------------------------------------------------------- Philox4x32-10
// Counter-based RNG: key = seed, counter = (env index lo, env index hi, tick, stream).  Same constants and
// round structure as Random123; the tests compare the device stream with an independent CPU restatement bit for bit.
struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t seed, uint64_t index, uint32_t tick, uint32_t stream) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = tick, c3 = stream;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 r;
    r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
    return r;
}

enum : uint32_t {
    kStreamReset0 = 0,  // streams 0,1,2: the six reset uniforms (two 53-bit uniforms per block)
    kStreamStep = 3     // per-step block: words 0,1 -> process noise pair, words 2,3 -> exploration noise
};

__host__ __device__ __forceinline__ double u01_53(uint32_t lo, uint32_t hi) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// six uniforms in [0,1): order a1,a2,Kp,h1,h2,r (water tank) / qww_V,qc_V,x,r,-,- (pH)
__device__ __forceinline__ void reset_uniforms(uint64_t seed, uint64_t index, uint32_t episode, double u[6]) {
#pragma unroll
    for (uint32_t s = 0; s < 3; ++s) {
        Philox4 w = philox4x32_10(seed, index, episode, kStreamReset0 + s);
        u[2 * s] = u01_53(w.v[0], w.v[1]);
        u[2 * s + 1] = u01_53(w.v[2], w.v[3]);
    }
}

// Box-Muller on two 32-bit words -> two N(0,1) floats (fp32 is ample for noise terms).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &z0, float &z1) {
    float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

// --------------------------------------------------------------------------------------------- arithmetic policy
// f64 follows the numpy reference operation by operation: every product and sum is rounded separately
// (__dmul_rn/__dadd_rn are never contracted into an FMA).  f32 is the throughput path: FMA contraction and
// MUFU approximations are allowed there.
template <typename T> struct Num;

template <> struct Num<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
};

template <> struct Num<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ float sqrt(float a) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
};

template <typename T> __device__ __forceinline__ T clip_lo0(T v) { return v < (T)0 ? (T)0 : v; }

template <typename T> __device__ __forceinline__ T nan_of() { return (T)__int_as_float(0x7fc00000); }
template <typename T> __device__ __forceinline__ T clampT(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }

// reward of |achieved - goal| (nonlinear_watertank.py:486-514, ph.py:202-225)
template <typename T> __device__ __forceinline__ T reward_of(int reward_type, T d, T z1, T thr) {
    using N = Num<T>;
    if (reward_type == PIME_REWARD_SPARSE) return d > thr ? (T)-1 : (T)-0.0;
    if (reward_type == PIME_REWARD_DISTANCE) return N::mul(-d, z1);
    return N::mul(-N::mul(d, d), z1);
}

}  // namespace pime
