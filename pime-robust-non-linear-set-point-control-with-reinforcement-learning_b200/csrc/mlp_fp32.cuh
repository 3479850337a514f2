// mlp_fp32.cuh -- the FIDELITY mode of the actor / critic forward: every layer in fp32 on the CUDA cores, tanhf / fmaxf
// activations, straight from the fp32 state_dict parameters (kept behind the fp16 tiles of the pack).  Same arithmetic type
// as the reference's torch modules (elegantrl/net_residual.py:138-205, :6-66, net.py:274-277); only the summation order
// differs from an fp32 GEMM, so |a_avg - torch fp32| is ~1e-6 instead of the ~1e-3 of the fp16-operand tensor-core engine.
// It exists for parity (closed-loop comparisons against the reference with a non-trivial actor, the critic values the
// learner consumes); it is ~25x slower than the tcgen05 engine and is never the bench path.
//
// One CTA = 256 threads = kFR (32) rows.  Activations live in shared memory TRANSPOSED, [unit][row] with a row stride of 33
// floats (conflict-free both ways); thread n owns output unit n of a layer and keeps the 32 row sums in registers.
#pragma once

#include "tc_mlp.cuh"

namespace pime {
namespace f32 {

constexpr int kFR = 32;          // rows (envs) per CTA
constexpr int kFThreads = 256;   // one thread per output unit (H <= 256)
constexpr int kRS = 33;          // row stride of the transposed activation tiles
constexpr int kTile = 256 * kRS; // floats of one activation tile
constexpr int kSmemFloats = 2 * kTile + 32 * kRS + 64;   // two activation tiles, the observation tile, the outputs
constexpr int kSmemBytes = kSmemFloats * 4;

enum { ACT_TANH = 0, ACT_RELU = 1 };

// out[n][r] = act(b[n] + sum_k W[n][k] * in[k][r]) for n < N; W is the torch layout [N][ldw] in global memory (L1 / L2)
template <int ACT>
__device__ __forceinline__ void dense(const float *__restrict__ W, int ldw, const float *__restrict__ b, int N, int K, const float *in,
                                      float *out) {
    const int n = threadIdx.x;
    if (n < N) {
        float acc[kFR];
        const float bias = __ldg(b + n);
#pragma unroll
        for (int r = 0; r < kFR; ++r) acc[r] = bias;
        const float *w = W + (size_t)n * ldw;
        for (int k = 0; k < K; ++k) {
            const float wk = __ldg(w + k);
            const float *a = in + k * kRS;
#pragma unroll
            for (int r = 0; r < kFR; ++r) acc[r] = fmaf(wk, a[r], acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kFR; ++r) out[n * kRS + r] = ACT == ACT_RELU ? fmaxf(acc[r], 0.0f) : tanhf(acc[r]);
    }
}

// Linear(K -> 1): thread r < kFR owns row r
__device__ __forceinline__ void dense_out(const float *__restrict__ w, float b, int K, const float *in, float *out) {
    const int r = threadIdx.x;
    if (r < kFR) {
        float acc = b;
        for (int k = 0; k < K; ++k) acc = fmaf(__ldg(w + k), in[k * kRS + r], acc);
        out[r] = acc;
    }
}

// net(obs) for the kFR rows whose observations sit in sObs[k][r] (k < S); result in sOut[r].  `p` = the fp32 parameters in
// state_dict order (PackLayout::src offsets).  All 256 threads must call it.
__device__ __forceinline__ void forward(const tc::PackLayout &L, const float *__restrict__ p, const float *sObs, float *tA, float *tB,
                                        float *sOut) {
    const int H = L.H, S = L.S;
    if (L.kind == PIME_ACTOR_MODULAR) {                       // net_residual.py:150-170
        const int So = S - L.D, Hh = H / 2;
        dense<ACT_TANH>(p + L.src[0], So, p + L.src[1], H, So, sObs, tA);                    // other_net.0
        __syncthreads();
        dense<ACT_TANH>(p + L.src[2], H, p + L.src[3], Hh, H, tA, tB);                       // other_net.2 -> cat[0:H/2]
        __syncthreads();
        dense<ACT_TANH>(p + L.src[4], L.D, p + L.src[5], H, L.D, sObs + So * kRS, tA);       // integrator_net.0
        __syncthreads();
        dense<ACT_TANH>(p + L.src[6], H, p + L.src[7], Hh, H, tA, tB + Hh * kRS);            // integrator_net.2 -> cat[H/2:H]
        __syncthreads();
        dense<ACT_TANH>(p + L.src[8], H, p + L.src[9], H, H, tB, tA);                        // net.0
        __syncthreads();
        dense_out(p + L.src[10], __ldg(p + L.src[11]), H, tA, sOut);                         // net.2
    } else if (L.kind == PIME_CRITIC_ADV) {                   // net.py:274-277
        dense<ACT_RELU>(p + L.src[0], S, p + L.src[1], H, S, sObs, tA);
        __syncthreads();
        dense<ACT_RELU>(p + L.src[2], H, p + L.src[3], H, H, tA, tB);
        __syncthreads();
        dense<ACT_RELU>(p + L.src[4], H, p + L.src[5], H, H, tB, tA);
        __syncthreads();
        dense_out(p + L.src[6], __ldg(p + L.src[7]), H, tA, sOut);
    } else {                                                  // net_residual.py:6-66
        dense<ACT_TANH>(p + L.src[0], S, p + L.src[1], H, S, sObs, tA);
        __syncthreads();
        dense<ACT_TANH>(p + L.src[2], H, p + L.src[3], H, H, tA, tB);
        __syncthreads();
        dense<ACT_TANH>(p + L.src[4], H, p + L.src[5], H, H, tB, tA);
        __syncthreads();
        dense_out(p + L.src[6], __ldg(p + L.src[7]), H, tA, sOut);
    }
    __syncthreads();
}

}  // namespace f32
}  // namespace pime
