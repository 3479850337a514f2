// actor.cu -- weight packing and the stand-alone actor / critic forward on tcgen05 (see tc_mlp.cuh).
#include "mlp_fp32.cuh"

namespace pime {

// fp32 torch-layout parameters -> kernel image: the block program (header) + fp16 weight blocks in tcgen05 operand
// layout ([K/8 core-matrix columns][N rows][8 fp16]).  One CTA per block.
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ tc::PackLayout L, const float *__restrict__ p,
                                                   uint8_t *__restrict__ out) {
    const int b = blockIdx.x;
    if (b > L.nblk) {    // fp32 copy of the parameters behind the fp16 blocks (fidelity-mode forward)
        float *dst = reinterpret_cast<float *>(out + L.f32_off);
        for (int j = (b - L.nblk - 1) * blockDim.x + threadIdx.x; j < L.param_count; j += (gridDim.x - L.nblk - 1) * blockDim.x) dst[j] = p[j];
        return;
    }
    if (b == L.nblk) {   // fp32 output layer: weight vector + bias into the header
        float *ow = reinterpret_cast<float *>(out + tc::kOutWOff);
        for (int j = threadIdx.x; j < L.H; j += blockDim.x) ow[j] = p[L.out_w + j];
        if (threadIdx.x == 0) ow[L.H] = p[L.out_b];
        if (L.kind == PIME_ACTOR_MODULAR) {   // integrator_net.0 (one input): interleaved (w, b) pairs, fp32
            float *wb = reinterpret_cast<float *>(out + tc::kL1iOff);
            for (int j = threadIdx.x; j < L.H; j += blockDim.x) { wb[2 * j] = p[L.l1i_w + j]; wb[2 * j + 1] = p[L.l1i_b + j]; }
        }
        return;
    }
    const tc::Blk B = L.blk[b];
    const tc::BlkSrc s = L.bsrc[b];
    if (threadIdx.x == 0) reinterpret_cast<tc::Blk *>(out)[b] = B;
    const int NB = B.nb8 * 8, K = B.k16s * 16;
    __half *dst = reinterpret_cast<__half *>(out + tc::kHeaderBytes + B.src_off);
    for (int e = threadIdx.x; e < NB * K; e += blockDim.x) {
        const int n = e / K, k = e % K;
        float v = 0.0f;
        bool lo = false;
        if (n < s.n_real) {
            const int u = s.n0 + n;
            if (s.type == tc::SRC_HID) {
                v = p[s.w_off + (int64_t)u * s.ld + (s.k0 + k + s.k_rot) % s.ld];
            } else if (s.type == tc::SRC_BIAS) {
                if (k < 2) { v = p[s.b_off + u]; lo = k == 1; }
            } else {  // SRC_L1: [W_hi | W_hi | W_lo (3 terms) | b_hi b_lo] against [in_hi | in_lo | in_hi | 1 1]
                const int kk = s.k0 + k, nw = L.nterms * L.nin;
                if (kk < nw) {
                    const int t = kk / L.nin, c = kk % L.nin;
                    if (c >= s.c0 && c < s.c0 + s.cN) { v = p[s.w_off + (int64_t)u * s.ld + (c - s.c0)]; lo = t == 2; }
                } else if (kk < nw + 2) {
                    v = p[s.b_off + u];
                    lo = kk == nw + 1;
                }
            }
        }
        __half h = __float2half_rn(v);
        if (lo) h = __float2half_rn(v - __half2float(h));
        dst[(size_t)(k / 8) * (NB * 8) + (size_t)n * 8 + (k % 8)] = h;
    }
}

// Stand-alone forward over n rows (1 pass per group): the owners load the row-major observations and store net(obs).
template <int KIND, int H>
__global__ void __launch_bounds__(tc::kThreads, 1) actor_forward_kernel(tc::MlpParams mp, int64_t n, const float *__restrict__ obs,
                                                                        float *__restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    tc::Engine<KIND, H> eng;
    eng.setup(smem, mp);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp < 8) {
        eng.worker_loop(2);
    } else if (warp < 12) {
        const int row = tid - tc::kWorkerThreads;
        float o[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = 0.0f;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            const int64_t i = (int64_t)blockIdx.x * tc::kTileEnvs + g * tc::kRows + row;
            const int64_t ii = i < n ? i : n - 1;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < mp.S) o[k] = obs[ii * mp.S + k];
            eng.write_obs(row, g, o);
        }
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            const int64_t i = (int64_t)blockIdx.x * tc::kTileEnvs + g * tc::kRows + row;
            const float a = eng.read_out(row, g);
            if (i < n) out[i] = a;
        }
    } else if (warp == tc::kMmaWarp) {
        eng.mma_loop(2);
    } else if ((tid & 31) == 0) {
        eng.producer_loop(2);
    }
    eng.teardown();
}

// Fidelity mode: fp32 CUDA-core forward (mlp_fp32.cuh), 32 rows per CTA.
__global__ void __launch_bounds__(f32::kFThreads) actor_forward_fp32_kernel(const __grid_constant__ tc::PackLayout L,
                                                                            const float *__restrict__ params, int64_t n,
                                                                            const float *__restrict__ obs, float *__restrict__ out) {
    extern __shared__ __align__(16) float sm32[];
    float *tA = sm32, *tB = tA + f32::kTile, *sObs = tB + f32::kTile, *sOut = sObs + 32 * f32::kRS;
    const int64_t r0 = (int64_t)blockIdx.x * f32::kFR;
    for (int j = threadIdx.x; j < L.S * f32::kFR; j += blockDim.x) {
        const int r = j / L.S, k = j % L.S;
        const int64_t i = r0 + r < n ? r0 + r : n - 1;
        sObs[k * f32::kRS + r] = obs[i * L.S + k];
    }
    __syncthreads();
    f32::forward(L, params, sObs, tA, tB, sOut);
    if (threadIdx.x < f32::kFR && r0 + threadIdx.x < n) out[r0 + threadIdx.x] = sOut[threadIdx.x];
}

static int launch_forward_fp32(const tc::PackLayout &L, const void *pack, int64_t n, const float *obs, float *out, cudaStream_t stream) {
    PIME_CUDA(cudaFuncSetAttribute(actor_forward_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f32::kSmemBytes));
    const int64_t grid = (n + f32::kFR - 1) / f32::kFR;
    PIME_REQUIRE(grid <= 0x7fffffffLL, "too many rows for one launch");
    actor_forward_fp32_kernel<<<(unsigned)grid, f32::kFThreads, f32::kSmemBytes, stream>>>(
        L, reinterpret_cast<const float *>((const uint8_t *)pack + L.f32_off), n, obs, out);
    PIME_LAUNCH_CHECK();
    return PIME_OK;
}

template <int KIND, int H>
static int launch_forward_kh(const tc::PackLayout &L, const void *pack, int64_t n, const float *obs, float *out, cudaStream_t stream) {
    using G = tc::Geo<KIND, H>;
    auto kern = actor_forward_kernel<KIND, H>;
    PIME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SmemBytes));
    const int64_t grid = (n + tc::kTileEnvs - 1) / tc::kTileEnvs;
    PIME_REQUIRE(grid <= 0x7fffffffLL, "too many rows for one launch");
    kern<<<(unsigned)grid, tc::kThreads, G::SmemBytes, stream>>>(tc::make_mlp_params(L, pack), n, obs, out);
    PIME_LAUNCH_CHECK();
    return PIME_OK;
}

template <int KIND>
static int launch_forward_k(const tc::PackLayout &L, const void *pack, int64_t n, const float *obs, float *out, cudaStream_t stream) {
    switch (L.H) {
        case 32: return launch_forward_kh<KIND, 32>(L, pack, n, obs, out, stream);
        case 64: return launch_forward_kh<KIND, 64>(L, pack, n, obs, out, stream);
        case 128: return launch_forward_kh<KIND, 128>(L, pack, n, obs, out, stream);
        case 256: return launch_forward_kh<KIND, 256>(L, pack, n, obs, out, stream);
    }
    set_error("mid_dim must be 32, 64, 128 or 256");
    return PIME_EINVAL;
}

}  // namespace pime

using namespace pime;

extern "C" {

int64_t pime_actor_param_count(const pime_actor_config *cfg) {
    tc::PackLayout L;
    if (!cfg || !tc::make_pack_layout(*cfg, L)) return -1;
    return L.param_count;
}

int64_t pime_actor_pack_bytes(const pime_actor_config *cfg) {
    tc::PackLayout L;
    if (!cfg || !tc::make_pack_layout(*cfg, L)) return -1;
    return L.total_bytes;
}

int32_t pime_actor_block_list(const pime_actor_config *cfg, int32_t *out, int32_t max_blocks) {
    tc::PackLayout L;
    if (!cfg || !tc::make_pack_layout(*cfg, L)) return -1;
    for (int b = 0; b < L.nblk && b < max_blocks && out; ++b) {
        out[4 * b + 0] = L.blk[b].nb8 * 8;
        out[4 * b + 1] = L.blk[b].k16s;
        out[4 * b + 2] = L.blk[b].d_col;
        out[4 * b + 3] = L.blk[b].bytes16 * 16;
    }
    return L.nblk;
}

int pime_actor_pack(const pime_actor_config *cfg, const float *params, void *pack, void *stream) {
    PIME_REQUIRE(cfg && params && pack, "null pointer");
    tc::PackLayout L;
    PIME_REQUIRE(tc::make_pack_layout(*cfg, L), "unsupported actor dimensions (H in {32,64,128,256}, S <= 32, modular: S-1 <= 3)");
    PIME_REQUIRE(((uintptr_t)pack & 127) == 0, "pack must be 128-byte aligned");
    if (int rc = require_device()) return rc;
    pack_kernel<<<L.nblk + 1 + 16, 256, 0, (cudaStream_t)stream>>>(L, params, (uint8_t *)pack);
    PIME_LAUNCH_CHECK();
    return PIME_OK;
}

int pime_actor_forward(const pime_actor_config *cfg, const void *pack, int64_t n, const float *obs, float *a_avg, void *stream) {
    PIME_REQUIRE(cfg && pack && obs && a_avg, "null pointer");
    tc::PackLayout L;
    PIME_REQUIRE(tc::make_pack_layout(*cfg, L), "unsupported actor dimensions");
    if (int rc = require_device()) return rc;
    if (n <= 0) return PIME_OK;
    if (cfg->precision == PIME_PRECISION_FP32) return launch_forward_fp32(L, pack, n, obs, a_avg, (cudaStream_t)stream);
    PIME_REQUIRE(cfg->precision == PIME_PRECISION_TC, "precision");
    switch (cfg->kind) {
        case PIME_ACTOR_PLAIN: return launch_forward_k<PIME_ACTOR_PLAIN>(L, pack, n, obs, a_avg, (cudaStream_t)stream);
        case PIME_ACTOR_MODULAR: return launch_forward_k<PIME_ACTOR_MODULAR>(L, pack, n, obs, a_avg, (cudaStream_t)stream);
        case PIME_CRITIC_ADV: return launch_forward_k<PIME_CRITIC_ADV>(L, pack, n, obs, a_avg, (cudaStream_t)stream);
    }
    set_error("unknown actor kind");
    return PIME_EINVAL;
}

}  // extern "C"
