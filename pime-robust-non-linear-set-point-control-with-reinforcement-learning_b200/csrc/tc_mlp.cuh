// tc_mlp.cuh -- the residual-actor MLP on Blackwell tensor cores (tcgen05 + TMEM + TMA bulk copies), written for
// the fused rollout: 128 envs per CTA = the M=128 rows of every tcgen05.mma = the 128 TMEM lanes, so that thread r
// of the CTA owns env r end to end (plant state in registers, its activations in "its" TMEM lane, its row of the A
// operand in shared memory).  No cross-thread traffic other than the tensor-core operands.
//
//   warps 0-3 (128 threads) : env threads.  Produce A (fp16, canonical no-swizzle K-major core-matrix layout),
//                             run the epilogues TMEM -> regs -> (+bias, tanh) -> fp16 A of the next layer.
//   warp 4, lane 0          : MMA issuer.  tcgen05.mma.cta_group::1.kind::f16, M=128, N<=128 per instruction,
//                             accumulators in TMEM, tcgen05.commit -> mbarriers.
//   warp 5, lane 0          : TMA producer.  cp.async.bulk streams the pre-tiled fp16 weight blocks (8 KB) from
//                             the L2-resident pack through a kStages-deep shared-memory ring.
//
// Layers (reference elegantrl/net_residual.py):
//   modular (:138-205): other_net L1 and integrator_net L1 on CUDA cores (K = 3 / 1), then
//                       G0: [128 x H] x Wo1^T -> D[:, 0:H/2],  G1: [128 x H] x Wi1^T -> D[:, H/2:H],
//                       G2: tanh(D + b)[128 x H] x Wn0^T -> D[:, 0:H],  out = tanh(D + bn0) . Wn1 + bn1.
//   plain (:6-66) / CriticAdv (net.py:274-277): the S-wide first layer also runs on the tensor core by splitting
//                       the fp32 observation into fp16 hi + lo parts (K = 2S padded to 32/64), then two H x H GEMMs.
#pragma once

#include "pime_common.cuh"

namespace pime {
namespace tc {

constexpr int kRows = 128;
constexpr int kEnvThreads = 128;
constexpr int kThreads = 192;
constexpr int KB = 32;               // K extent of one streamed weight block (two K=16 MMAs)
constexpr int kStages = 4;
constexpr int kMaxBlkBytes = 128 * KB * 2;
constexpr int kChunkBytes = kRows * 16;  // one K core-matrix column (8 fp16) for all 128 rows = 2048 B

// ------------------------------------------------------------------------------------------------ pack layout
struct PackLayout {
    int kind, H, S, D, KP;
    int param_count;
    int f32_floats;
    int off_l1o, off_l1i, off_b0, off_b1, off_ep2, off_sc;   // float offsets in the fp32 section
    int f16_off;                                             // byte offset of the fp16 section
    int phN[3], phK[3], phNB[3], phCol[3], phOff[3];         // per GEMM phase; phOff in bytes within fp16 section
    int src[12];                                             // offsets of the state_dict tensors inside `params`
    int src_w[3], src_ld[3];                                 // fp32 source matrix offset / leading dim in params
    int blocks_per_step;
    int total_bytes;
};

inline bool make_pack_layout(const pime_actor_config &c, PackLayout &L) {
    const int H = c.mid_dim, S = c.state_dim, D = c.integrator_dim;
    if (!(H == 32 || H == 64 || H == 128 || H == 256)) return false;
    if (S < 1 || S > 32) return false;
    L = PackLayout{};
    L.kind = c.kind; L.H = H; L.S = S; L.D = D;
    const int Hh = H / 2;
    if (c.kind == PIME_ACTOR_MODULAR) {
        const int So = S - D;
        if (D != 1 || So < 1 || So > 3) return false;
        // state_dict order: other_net.0.{w,b} other_net.2.{w,b} integrator_net.0.{w,b} integrator_net.2.{w,b}
        //                   net.0.{w,b} net.2.{w,b}
        const int sizes[12] = {H * So, H, Hh * H, Hh, H * D, H, Hh * H, Hh, H * H, H, H, 1};
        int o = 0;
        for (int j = 0; j < 12; ++j) { L.src[j] = o; o += sizes[j]; }
        L.param_count = o;
        L.KP = 0;
        L.off_l1o = 0; L.off_l1i = 4 * H; L.off_b1 = 6 * H; L.off_ep2 = 7 * H; L.off_sc = 9 * H; L.off_b0 = 0;
        L.f32_floats = 9 * H + 4;
        L.phN[0] = Hh; L.phN[1] = Hh; L.phN[2] = H;
        L.phK[0] = H; L.phK[1] = H; L.phK[2] = H;
        L.phCol[0] = 0; L.phCol[1] = Hh; L.phCol[2] = 0;
        L.src_w[0] = L.src[2]; L.src_w[1] = L.src[6]; L.src_w[2] = L.src[8];
        L.src_ld[0] = H; L.src_ld[1] = H; L.src_ld[2] = H;
    } else {
        // state_dict order: net.0.{w,b} net.2.{w,b} net.4.{w,b} net.6.{w,b}
        const int sizes[8] = {H * S, H, H * H, H, H * H, H, H, 1};
        int o = 0;
        for (int j = 0; j < 8; ++j) { L.src[j] = o; o += sizes[j]; }
        L.param_count = o;
        L.KP = (2 * S <= 32) ? 32 : 64;
        L.off_b0 = 0; L.off_b1 = H; L.off_ep2 = 2 * H; L.off_sc = 4 * H; L.off_l1o = 0; L.off_l1i = 0;
        L.f32_floats = 4 * H + 4;
        L.phN[0] = H; L.phN[1] = H; L.phN[2] = H;
        L.phK[0] = L.KP; L.phK[1] = H; L.phK[2] = H;
        L.phCol[0] = 0; L.phCol[1] = 0; L.phCol[2] = 0;
        L.src_w[0] = L.src[0]; L.src_w[1] = L.src[2]; L.src_w[2] = L.src[4];
        L.src_ld[0] = S; L.src_ld[1] = H; L.src_ld[2] = H;
    }
    L.f16_off = ((L.f32_floats * 4 + 127) / 128) * 128;
    int off = 0;
    L.blocks_per_step = 0;
    for (int p = 0; p < 3; ++p) {
        L.phNB[p] = L.phN[p] < 128 ? L.phN[p] : 128;
        L.phOff[p] = off;
        off += L.phN[p] * L.phK[p] * 2;
        L.blocks_per_step += (L.phN[p] / L.phNB[p]) * (L.phK[p] / KB);
    }
    L.total_bytes = L.f16_off + off;
    return true;
}

// what the kernels need of the layout (passed by value)
struct MlpParams {
    const uint8_t *pack;
    int f32_floats, f16_off;
    int phK[3], phOff[3];
    int S, KP;
    int off_l1o, off_l1i, off_b0, off_b1, off_ep2, off_sc;
};

inline MlpParams make_mlp_params(const PackLayout &L, const void *pack) {
    MlpParams m;
    m.pack = (const uint8_t *)pack;
    m.f32_floats = L.f32_floats; m.f16_off = L.f16_off;
    for (int p = 0; p < 3; ++p) { m.phK[p] = L.phK[p]; m.phOff[p] = L.phOff[p]; }
    m.S = L.S; m.KP = L.KP;
    m.off_l1o = L.off_l1o; m.off_l1i = L.off_l1i; m.off_b0 = L.off_b0; m.off_b1 = L.off_b1; m.off_ep2 = L.off_ep2; m.off_sc = L.off_sc;
    return m;
}

template <int KIND, int H> struct Geo {
    static constexpr bool kModular = KIND == PIME_ACTOR_MODULAR;
    static constexpr bool kRelu = KIND == PIME_CRITIC_ADV;
    static constexpr int Hh = H / 2;
    static constexpr int ACols = H < 64 ? 64 : H;             // A tile columns (plain phase 0 may need 64)
    static constexpr int ABytes = kRows * ACols * 2;
    static constexpr int TmemCols = H < 32 ? 32 : H;
    static constexpr int F32Floats = kModular ? 9 * H + 4 : 4 * H + 4;
    static constexpr int RingBytes = kStages * kMaxBlkBytes;
    static constexpr int SmemBytes = ABytes + RingBytes + ((F32Floats * 4 + 15) / 16) * 16 + 256;
    __host__ __device__ static constexpr int phN(int p) { return kModular ? (p == 2 ? H : Hh) : H; }
    __host__ __device__ static constexpr int phNB(int p) { return phN(p) < 128 ? phN(p) : 128; }
    __host__ __device__ static constexpr int phCol(int p) { return kModular ? (p == 1 ? Hh : 0) : 0; }
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a descriptor or protocol bug must fault, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// smem matrix descriptor, SWIZZLE_NONE, K-major: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version(1)<<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor kind::f16: D=f32 (bit 4), A=B=f16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of warp w gets row 32*(w%4)+t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t *u = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
          "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
          "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool RELU> __device__ __forceinline__ float act_fn(float x) { return RELU ? fmaxf(x, 0.0f) : tanh_fast(x); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ------------------------------------------------------------------------------------------------ the engine
template <int KIND, int H> struct Engine {
    using G = Geo<KIND, H>;
    uint8_t *sA, *sRing;
    float *sF;
    uint64_t *full, *empty, *a_ready, *d_ready;
    uint32_t *tmem_slot;
    uint32_t tmem_base;
    uint32_t dph;  // parity of the next d_ready completion (env threads)
    MlpParams mp;

    // All kThreads threads.  Carves shared memory, initialises barriers, allocates TMEM, loads the fp32 vectors.
    __device__ __forceinline__ void setup(uint8_t *smem, const MlpParams &p) {
        mp = p;
        sA = smem;
        sRing = smem + G::ABytes;
        sF = reinterpret_cast<float *>(smem + G::ABytes + G::RingBytes);
        uint8_t *tail = smem + G::ABytes + G::RingBytes + ((G::F32Floats * 4 + 15) / 16) * 16;
        full = reinterpret_cast<uint64_t *>(tail);
        empty = full + kStages;
        a_ready = empty + kStages;
        d_ready = a_ready + 1;
        tmem_slot = reinterpret_cast<uint32_t *>(d_ready + 1);
        dph = 0;
        const int tid = threadIdx.x;
        if (tid == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            mbar_init(a_ready, kEnvThreads);
            mbar_init(d_ready, 1);
            fence_barrier_init();
        }
        if (tid / 32 == 4) tmem_alloc(tmem_slot, G::TmemCols);
        const float *src = reinterpret_cast<const float *>(mp.pack);
        for (int j = tid; j < G::F32Floats; j += kThreads) sF[j] = __ldg(src + j);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = *tmem_slot;
    }

    __device__ __forceinline__ void teardown() {
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x / 32 == 4) tmem_dealloc(tmem_base, G::TmemCols);
    }

    // ---- warp 5 lane 0: stream every weight block of every step through the ring
    __device__ __forceinline__ void producer_loop(int steps) {
        const uint8_t *w16 = mp.pack + mp.f16_off;
        uint32_t it = 0;
        for (int s = 0; s < steps; ++s) {
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const int NB = G::phNB(p), nbn = G::phN(p) / NB;
                const uint32_t bytes = (uint32_t)NB * KB * 2;
                const int nblk = nbn * (mp.phK[p] / KB);
                const uint8_t *src = w16 + mp.phOff[p];
                for (int b = 0; b < nblk; ++b, ++it) {
                    const uint32_t st = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&full[st], bytes);
                    bulk_g2s(sRing + st * kMaxBlkBytes, src + (size_t)b * bytes, bytes, &full[st]);
                }
            }
        }
    }

    // ---- warp 4 lane 0: issue the MMAs of every phase of every step
    __device__ __forceinline__ void mma_loop(int steps) {
        uint32_t it = 0, aph = 0;
        const uint32_t a_base = smem_u32(sA);
        for (int s = 0; s < steps; ++s) {
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const int NB = G::phNB(p), nbn = G::phN(p) / NB;
                const uint32_t idesc = make_idesc(kRows, NB);
                const int nkb = mp.phK[p] / KB;
                mbar_wait(a_ready, aph);
                aph ^= 1;
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    for (int nb = 0; nb < nbn; ++nb, ++it) {
                        const uint32_t st = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait(&full[st], ph);
                        tc_fence_after();
                        const uint32_t b_base = smem_u32(sRing + st * kMaxBlkBytes);
                        const uint32_t d_addr = tmem_base + (uint32_t)(G::phCol(p) + nb * NB);
#pragma unroll
                        for (int kk = 0; kk < KB / 16; ++kk) {
                            const int k16 = kb * (KB / 16) + kk;
                            const uint64_t adesc = make_desc(a_base + (uint32_t)k16 * 2 * kChunkBytes, kChunkBytes, 128);
                            const uint64_t bdesc = make_desc(b_base + (uint32_t)kk * 2 * (NB * 16), NB * 16, 128);
                            mma_f16(d_addr, adesc, bdesc, idesc, k16 > 0 ? 1u : 0u);
                        }
                        mma_commit(&empty[st]);  // frees the ring slot once these MMAs have read it
                    }
                }
                mma_commit(d_ready);  // accumulator of this phase complete (and A no longer being read)
            }
        }
    }

    // ---- env threads
    __device__ __forceinline__ void a_store8(int row, int kchunk, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
        *reinterpret_cast<uint4 *>(sA + (size_t)kchunk * kChunkBytes + row * 16) = make_uint4(p0, p1, p2, p3);
    }
    __device__ __forceinline__ void signal_a() {
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(a_ready);
    }
    __device__ __forceinline__ void wait_d() {
        mbar_wait(d_ready, dph);
        dph ^= 1;
        tc_fence_after();
    }

    // epilogue of a hidden GEMM: A[:, c] = act(D[:, c] + bias[c]) for c in [0, ncols)
    __device__ __forceinline__ void epilogue_to_a(int row, int ncols, const float *bias) {
        const uint32_t taddr = tmem_base + ((uint32_t)((row / 32) * 32) << 16);
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 b0 = *reinterpret_cast<const float4 *>(bias + c0 + q * 8);
                const float4 b1 = *reinterpret_cast<const float4 *>(bias + c0 + q * 8 + 4);
                const float x0 = act_fn<G::kRelu>(v[q * 8 + 0] + b0.x), x1 = act_fn<G::kRelu>(v[q * 8 + 1] + b0.y);
                const float x2 = act_fn<G::kRelu>(v[q * 8 + 2] + b0.z), x3 = act_fn<G::kRelu>(v[q * 8 + 3] + b0.w);
                const float x4 = act_fn<G::kRelu>(v[q * 8 + 4] + b1.x), x5 = act_fn<G::kRelu>(v[q * 8 + 5] + b1.y);
                const float x6 = act_fn<G::kRelu>(v[q * 8 + 6] + b1.z), x7 = act_fn<G::kRelu>(v[q * 8 + 7] + b1.w);
                a_store8(row, c0 / 8 + q, pack_h2(x0, x1), pack_h2(x2, x3), pack_h2(x4, x5), pack_h2(x6, x7));
            }
        }
    }

    // last layer fused into the epilogue: sum_c act(D[:, c] + b[c]) * w[c] + b_out   (ep2 = interleaved (b, w))
    __device__ __forceinline__ float epilogue_dot(int row) {
        const uint32_t taddr = tmem_base + ((uint32_t)((row / 32) * 32) << 16);
        const float *ep2 = sF + mp.off_ep2;
        float acc0 = 0.f, acc1 = 0.f;
        for (int c0 = 0; c0 < H; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float4 bw = *reinterpret_cast<const float4 *>(ep2 + 2 * (c0 + j));
                acc0 = fmaf(act_fn<G::kRelu>(v[j] + bw.x), bw.y, acc0);
                acc1 = fmaf(act_fn<G::kRelu>(v[j + 1] + bw.z), bw.w, acc1);
            }
        }
        return acc0 + acc1 + sF[mp.off_sc];
    }

    // Full forward for the env owned by this thread; obs = float32 observation (S values).  All 128 env threads of
    // the CTA call this together.  Returns net(obs) (pre-tanh, pre-prior).
    __device__ __forceinline__ float forward(int row, const float *obs) {
        if constexpr (G::kModular) {
            const int So = mp.S - 1;
            const float o0 = obs[0], o1 = obs[1], o2 = So > 2 ? obs[2] : 0.0f, oi = obs[So];
            {   // other_net[0..1]: tanh(Wo0 obs_other + bo0)  (net_residual.py:151)
                const float4 *l1 = reinterpret_cast<const float4 *>(sF + mp.off_l1o);
#pragma unroll 2
                for (int kc = 0; kc < H / 8; ++kc) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 wb = l1[kc * 8 + j];
                        x[j] = tanh_fast(fmaf(wb.x, o0, fmaf(wb.y, o1, fmaf(wb.z, o2, wb.w))));
                    }
                    a_store8(row, kc, pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]), pack_h2(x[6], x[7]));
                }
            }
            signal_a();
            wait_d();  // G0 done: D[:, 0:H/2] = other hidden x Wo1^T, A free again
            {   // integrator_net[0..1]: tanh(Wi0 I + bi0)  (net_residual.py:154)
                const float4 *l1 = reinterpret_cast<const float4 *>(sF + mp.off_l1i);
#pragma unroll 2
                for (int kc = 0; kc < H / 8; ++kc) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const float4 wb = l1[(kc * 8 + j) / 2];  // (w_j, b_j, w_{j+1}, b_{j+1})
                        x[j] = tanh_fast(fmaf(wb.x, oi, wb.y));
                        x[j + 1] = tanh_fast(fmaf(wb.z, oi, wb.w));
                    }
                    a_store8(row, kc, pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]), pack_h2(x[6], x[7]));
                }
            }
            signal_a();
            wait_d();  // G1 done: D[:, H/2:H]
            epilogue_to_a(row, H, sF + mp.off_b1);  // cat(tanh(.. + bo1), tanh(.. + bi1))  (:152,:155,:170)
            signal_a();
            wait_d();  // G2 done: D[:, 0:H] = cat x Wn0^T
            return epilogue_dot(row);  // net[1..2]: tanh(. + bn0) . Wn1 + bn1  (:157-158)
        } else {
            {   // first layer on the tensor core: A = [fp16(obs) | fp16(obs - fp16(obs)) | 0], B = [W0 | W0 | 0]
                __align__(16) __half hl[64];
                const int S = mp.S;
#pragma unroll 1
                for (int k = 0; k < 64; ++k) hl[k] = __float2half_rn(0.0f);
#pragma unroll 1
                for (int k = 0; k < S; ++k) {
                    const __half hi = __float2half_rn(obs[k]);
                    hl[k] = hi;
                    hl[S + k] = __float2half_rn(obs[k] - __half2float(hi));
                }
                const uint4 *q = reinterpret_cast<const uint4 *>(hl);
                for (int kc = 0; kc < mp.KP / 8; ++kc) {
                    const uint4 u = q[kc];
                    a_store8(row, kc, u.x, u.y, u.z, u.w);
                }
            }
            signal_a();
            wait_d();
            epilogue_to_a(row, H, sF + mp.off_b0);
            signal_a();
            wait_d();
            epilogue_to_a(row, H, sF + mp.off_b1);
            signal_a();
            wait_d();
            return epilogue_dot(row);
        }
    }
};

}  // namespace tc
}  // namespace pime
