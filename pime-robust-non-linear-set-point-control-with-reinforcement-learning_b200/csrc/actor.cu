// actor.cu -- weight packing and the stand-alone actor / critic forward on tcgen05 (see tc_mlp.cuh).
#include "tc_mlp.cuh"

namespace pime {

struct PackKernelArgs {
    tc::PackLayout L;
};

// fp32 torch-layout parameters -> kernel image (fp32 vectors + fp16 weight blocks in tcgen05 operand layout)
__global__ void __launch_bounds__(256) pack_kernel(PackKernelArgs a, const float *__restrict__ p, uint8_t *__restrict__ out) {
    const tc::PackLayout &L = a.L;
    const int H = L.H, Hh = H / 2, S = L.S;
    float *f32 = reinterpret_cast<float *>(out);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // ---- fp32 section
    for (int64_t j = tid0; j < L.f32_floats; j += stride) {
        float v = 0.0f;
        const int jj = (int)j;
        if (L.kind == PIME_ACTOR_MODULAR) {
            const int So = S - L.D;
            if (jj < 4 * H) {            // l1o: (w0,w1,w2,b) per hidden unit
                const int u = jj / 4, c = jj % 4;
                v = c == 3 ? p[L.src[1] + u] : (c < So ? p[L.src[0] + u * So + c] : 0.0f);
            } else if (jj < 6 * H) {     // l1i: (w,b)
                const int u = (jj - 4 * H) / 2, c = (jj - 4 * H) % 2;
                v = c == 0 ? p[L.src[4] + u] : p[L.src[5] + u];
            } else if (jj < 7 * H) {     // b1 = cat(bo1, bi1)
                const int c = jj - 6 * H;
                v = c < Hh ? p[L.src[3] + c] : p[L.src[7] + c - Hh];
            } else if (jj < 9 * H) {     // ep2: (bn0[c], Wn1[c])
                const int c = (jj - 7 * H) / 2, w = (jj - 7 * H) % 2;
                v = w == 0 ? p[L.src[9] + c] : p[L.src[10] + c];
            } else if (jj == 9 * H) {
                v = p[L.src[11]];
            }
        } else {
            if (jj < H) v = p[L.src[1] + jj];
            else if (jj < 2 * H) v = p[L.src[3] + jj - H];
            else if (jj < 4 * H) {
                const int c = (jj - 2 * H) / 2, w = (jj - 2 * H) % 2;
                v = w == 0 ? p[L.src[5] + c] : p[L.src[6] + c];
            } else if (jj == 4 * H) v = p[L.src[7]];
        }
        f32[j] = v;
    }
    // ---- fp16 section
    __half *f16 = reinterpret_cast<__half *>(out + L.f16_off);
    for (int ph = 0; ph < 3; ++ph) {
        const int N = L.phN[ph], K = L.phK[ph], NB = L.phNB[ph], nbn = N / NB;
        const int64_t total = (int64_t)N * K;
        for (int64_t e = tid0; e < total; e += stride) {
            const int n = (int)(e / K), k = (int)(e % K);
            float v;
            if (L.kind != PIME_ACTOR_MODULAR && ph == 0) {  // [W0 | W0 | 0]: multiplies [obs_hi | obs_lo | 0]
                v = k < S ? p[L.src_w[0] + n * S + k] : (k < 2 * S ? p[L.src_w[0] + n * S + (k - S)] : 0.0f);
            } else {
                v = p[L.src_w[ph] + (int64_t)n * L.src_ld[ph] + k];
            }
            const int nb = n / NB, nl = n % NB, kb = k / tc::KB, kl = k % tc::KB;
            const int64_t blk = (int64_t)kb * nbn + nb;
            const int64_t byte_off = (int64_t)L.phOff[ph] + blk * ((int64_t)NB * tc::KB * 2) + (int64_t)(kl / 8) * (NB * 16) +
                                     (int64_t)nl * 16 + (kl % 8) * 2;
            f16[byte_off / 2] = __float2half_rn(v);
        }
    }
}

template <int KIND, int H, int MINB>
__global__ void __launch_bounds__(tc::kThreads, MINB) actor_forward_kernel(tc::MlpParams mp, int64_t n, const float *__restrict__ obs,
                                                                           float *__restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    tc::Engine<KIND, H> eng;
    eng.setup(smem, mp);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp < 4) {
        const int64_t i = (int64_t)blockIdx.x * tc::kRows + tid;
        const bool live = i < n;
        const int64_t ii = live ? i : n - 1;
        float o[32];
        for (int k = 0; k < mp.S; ++k) o[k] = obs[ii * mp.S + k];
        const float a = eng.forward(tid, o);
        if (live) out[i] = a;
    } else if ((tid & 31) == 0) {
        if (warp == 4) eng.mma_loop(1);
        else eng.producer_loop(1);
    }
    eng.teardown();
}

template <int KIND, int H>
static int launch_forward_kh(const tc::PackLayout &L, const void *pack, int64_t n, const float *obs, float *out, cudaStream_t stream) {
    using G = tc::Geo<KIND, H>;
    constexpr int MINB = G::SmemBytes <= 113 * 1024 ? 2 : 1;
    auto kern = actor_forward_kernel<KIND, H, MINB>;
    PIME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SmemBytes));
    const int64_t grid = (n + tc::kRows - 1) / tc::kRows;
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

int pime_actor_pack(const pime_actor_config *cfg, const float *params, void *pack, void *stream) {
    PIME_REQUIRE(cfg && params && pack, "null pointer");
    PackKernelArgs a;
    PIME_REQUIRE(tc::make_pack_layout(*cfg, a.L), "unsupported actor dimensions (H in {32,64,128,256}, S <= 32, modular: S-1 <= 3)");
    PIME_REQUIRE(((uintptr_t)pack & 127) == 0, "pack must be 128-byte aligned");
    if (int rc = require_device()) return rc;
    pack_kernel<<<kNumSMs, 256, 0, (cudaStream_t)stream>>>(a, params, (uint8_t *)pack);
    PIME_LAUNCH_CHECK();
    return PIME_OK;
}

int pime_actor_forward(const pime_actor_config *cfg, const void *pack, int64_t n, const float *obs, float *a_avg, void *stream) {
    PIME_REQUIRE(cfg && pack && obs && a_avg, "null pointer");
    tc::PackLayout L;
    PIME_REQUIRE(tc::make_pack_layout(*cfg, L), "unsupported actor dimensions");
    if (int rc = require_device()) return rc;
    if (n <= 0) return PIME_OK;
    switch (cfg->kind) {
        case PIME_ACTOR_PLAIN: return launch_forward_k<PIME_ACTOR_PLAIN>(L, pack, n, obs, a_avg, (cudaStream_t)stream);
        case PIME_ACTOR_MODULAR: return launch_forward_k<PIME_ACTOR_MODULAR>(L, pack, n, obs, a_avg, (cudaStream_t)stream);
        case PIME_CRITIC_ADV: return launch_forward_k<PIME_CRITIC_ADV>(L, pack, n, obs, a_avg, (cudaStream_t)stream);
    }
    set_error("unknown actor kind");
    return PIME_EINVAL;
}

}  // extern "C"
