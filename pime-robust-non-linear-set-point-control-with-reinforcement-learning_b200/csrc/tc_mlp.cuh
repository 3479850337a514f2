// tc_mlp.cuh -- the residual-actor MLP on Blackwell tensor cores (tcgen05 + TMEM + TMA bulk copies), written for
// the fused rollout: 128 envs per CTA = the M=128 rows of every tcgen05.mma = the 128 TMEM lanes, so that thread r
// of the CTA owns env r end to end (plant state in registers, its activations in "its" TMEM lane, its row of the A
// operand in shared memory).  No cross-thread traffic other than the tensor-core operands.
//
//   warps 0-3 (128 threads) : env threads.  Write the observation operand, run the epilogues
//                             TMEM -> regs -> tanh/relu -> fp16 A operand of the next layer.
//   warp 4, lane 0          : MMA issuer.  Interprets the block program of the pack: one or more
//                             tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=128, K=16) per streamed weight block,
//                             accumulators in TMEM, tcgen05.commit -> mbarriers.
//   warp 5, lane 0          : TMA producer.  cp.async.bulk streams the pre-tiled fp16 weight blocks (<= 8 KB) from
//                             the L2-resident pack through a kStages-deep shared-memory ring.
//
// EVERY layer runs on the tensor core, so that an env thread spends ~1.7 instructions per hidden activation
// (tcgen05.ld/32 + MUFU.TANH + F2FP/2 + STS/8) and the kernel is bound by the MUFU pipe:
//   * first layers (K = 1..30 inputs): the fp32 observation is split into fp16 hi + lo parts and the fp32 weight
//     into hi + lo parts, A = [in_hi | in_lo | in_hi | 1 1], B = [W_hi | W_hi | W_lo | b_hi b_lo]  (3 terms, ~2^-22
//     relative: fp32-grade first layer; 2 terms when 3S+2 > 32);
//   * hidden layers: A = fp16 activations written by the previous epilogue, B = fp16 weights; the bias enters
//     through one extra K=16 MMA against a constant "ones" operand, B = [b_hi b_lo 0 ...];
//   * output layer Linear(H -> 1): an N=16 MMA whose row 0 is the weight vector; column 0 of TMEM is net(obs).
//
// Layers (reference elegantrl/net_residual.py):
//   modular (:138-205): P0 other_net.0 -> D[0:H];  P1 other_net.2 -> D[0:H/2], integrator_net.0 (units 0..H/2) ->
//                       D[H/2:H];  P2 integrator_net.0 (units H/2..H) -> D[H/2:H];  P3 integrator_net.2 -> D[H/2:H];
//                       P4 net.0 on cat(D[0:H]) -> D[0:H];  P5 net.2 -> D[0:16].
//   plain (:6-66) / CriticAdv (net.py:274-277): P0 net.0, P1 net.2, P2 net.4, P3 net.6.
#pragma once

#include "pime_common.cuh"

namespace pime {
namespace tc {

constexpr int kRows = 128;
constexpr int kEnvThreads = 128;
constexpr int kThreads = 192;
constexpr int kStages = 4;
constexpr int kMaxBlkBytes = 8192;
constexpr int kChunkBytes = kRows * 16;    // one K core-matrix column (8 fp16) for all 128 rows = 2048 B
constexpr int kK16Bytes = 2 * kChunkBytes; // one K=16 slice of an A operand = 4096 B
constexpr int kMaxBlocks = 56;
constexpr int kHeaderBytes = 1024;         // the block program at the head of the pack (kMaxBlocks x 16 B)
constexpr int kMaxKP = 80;                 // widest first-layer operand: 2 x 32 inputs + 2 -> 80

// ------------------------------------------------------------------------------------------------ block program
enum : uint32_t { BLK_FRESH = 1u, BLK_WAIT_A = 2u, BLK_COMMIT_D = 4u };

struct Blk {             // one streamed weight block = k16s MMAs of shape 128 x (8*nb8) x 16
    uint32_t src_off;    // byte offset inside the fp16 section of the pack
    uint16_t bytes16;    // block bytes / 16
    uint8_t nb8;         // N / 8
    uint8_t k16s;        // K / 16
    uint16_t a_off16;    // byte offset / 16 of the A operand inside the CTA's operand area
    uint16_t d_col;      // first TMEM column of the accumulator
    uint32_t flags;      // BLK_*
};
static_assert(sizeof(Blk) == 16, "Blk is copied as uint4");

enum { SRC_HID = 0, SRC_BIAS = 1, SRC_L1 = 2 };
struct BlkSrc {          // how pack_kernel fills the block from the fp32 state_dict parameters
    int type;
    int w_off, ld;       // weight matrix [.., ld] at params + w_off
    int b_off;           // bias vector
    int n0, n_real;      // first output unit of the block, number of real (non-padding) rows
    int k0;              // first K index (SRC_HID: input unit; SRC_L1: position in the split operand)
    int c0, cN;          // SRC_L1: the weight matrix covers inputs [c0, c0+cN) of the nin-wide input vector
};

__host__ __device__ constexpr int geo_acols(int H) { return H < 64 ? 64 : H; }
__host__ __device__ constexpr int geo_abytes(int H) { return kRows * geo_acols(H) * 2; }
__host__ __device__ constexpr int geo_obs_off(int kind, int H) { return kind == PIME_ACTOR_MODULAR ? geo_abytes(H) : 0; }
__host__ __device__ constexpr int geo_ones_off(int kind, int H) { return geo_abytes(H) + (kind == PIME_ACTOR_MODULAR ? kK16Bytes : 0); }

struct PackLayout {
    int kind, H, S, D;
    int param_count;
    int nin, nterms, KP;   // first-layer operand: inputs, split terms (3 or 2), K padded to a multiple of 16
    int nblk;
    int f16_bytes, total_bytes;
    int src[12];           // offsets of the state_dict tensors inside `params`
    Blk blk[kMaxBlocks];
    BlkSrc bsrc[kMaxBlocks];
};

struct GemmSpec {
    int type, N, n_real, n0, K16, a_off, d_col, w_off, ld, b_off, c0, cN;
};

inline bool emit_gemm(PackLayout &L, const GemmSpec &g) {
    const int NB = g.N < 128 ? g.N : 128, nbn = g.N / NB;
    int kpb = kMaxBlkBytes / (NB * 32);
    if (kpb < 1) kpb = 1;
    if (kpb > g.K16) kpb = g.K16;
    for (int k = 0; k < g.K16; k += kpb) {
        const int kk = g.K16 - k < kpb ? g.K16 - k : kpb;
        for (int nb = 0; nb < nbn; ++nb) {
            if (L.nblk >= kMaxBlocks) return false;
            Blk &b = L.blk[L.nblk];
            BlkSrc &s = L.bsrc[L.nblk];
            b.src_off = (uint32_t)L.f16_bytes;
            b.bytes16 = (uint16_t)(NB * kk * 32 / 16);
            b.nb8 = (uint8_t)(NB / 8);
            b.k16s = (uint8_t)kk;
            b.a_off16 = (uint16_t)((g.a_off + k * kK16Bytes) / 16);
            b.d_col = (uint16_t)(g.d_col + nb * NB);
            b.flags = k == 0 ? BLK_FRESH : 0u;
            s.type = g.type; s.w_off = g.w_off; s.ld = g.ld; s.b_off = g.b_off;
            s.n0 = g.n0 + nb * NB;
            s.n_real = g.n_real - nb * NB;
            s.k0 = k * 16; s.c0 = g.c0; s.cN = g.cN;
            L.f16_bytes += NB * kk * 32;
            ++L.nblk;
        }
    }
    return true;
}

inline bool make_pack_layout(const pime_actor_config &c, PackLayout &L) {
    const int H = c.mid_dim, S = c.state_dim, D = c.integrator_dim;
    if (!(H == 32 || H == 64 || H == 128 || H == 256)) return false;
    if (S < 1 || S > 32) return false;
    if (!(c.kind == PIME_ACTOR_MODULAR || c.kind == PIME_ACTOR_PLAIN || c.kind == PIME_CRITIC_ADV)) return false;
    L = PackLayout{};
    L.kind = c.kind; L.H = H; L.S = S; L.D = D;
    const int Hh = H / 2, HK = H / 16;
    const int a_obs = geo_obs_off(c.kind, H), a_ones = geo_ones_off(c.kind, H);
    bool ok = true;
    int first = 0;
    auto phase_end = [&]() {
        L.blk[first].flags |= BLK_WAIT_A;
        L.blk[L.nblk - 1].flags |= BLK_COMMIT_D;
        first = L.nblk;
    };
    auto bias = [&](int N, int n_real, int b_off, int d_col) {
        ok = ok && emit_gemm(L, GemmSpec{SRC_BIAS, N, n_real, 0, 1, a_ones, d_col, 0, 0, b_off, 0, 0});
    };
    auto hid = [&](int N, int n_real, int w_off, int d_col) {   // accumulates on top of the bias block
        const int nb0 = L.nblk;
        ok = ok && emit_gemm(L, GemmSpec{SRC_HID, N, n_real, 0, HK, 0, d_col, w_off, H, 0, 0, 0});
        for (int j = nb0; j < L.nblk; ++j) L.blk[j].flags &= ~BLK_FRESH;
    };
    auto l1 = [&](int N, int n0, int w_off, int ld, int b_off, int c0, int cN, int d_col) {
        ok = ok && emit_gemm(L, GemmSpec{SRC_L1, N, N, n0, L.KP / 16, a_obs, d_col, w_off, ld, b_off, c0, cN});
    };
    if (c.kind == PIME_ACTOR_MODULAR) {
        const int So = S - D;
        if (D != 1 || So < 1 || So > 3) return false;
        // state_dict order: other_net.0.{w,b} other_net.2.{w,b} integrator_net.0.{w,b} integrator_net.2.{w,b}
        //                   net.0.{w,b} net.2.{w,b}
        const int sizes[12] = {H * So, H, Hh * H, Hh, H * D, H, Hh * H, Hh, H * H, H, H, 1};
        int o = 0;
        for (int j = 0; j < 12; ++j) { L.src[j] = o; o += sizes[j]; }
        L.param_count = o;
        L.nin = 4; L.nterms = 3; L.KP = 16;   // inputs (o0,o1,o2,I): 3 x 4 + 2 = 14 <= 16
        l1(H, 0, L.src[0], So, L.src[1], 0, So, 0);                     phase_end();  // P0 other_net.0
        bias(Hh, Hh, L.src[3], 0); hid(Hh, Hh, L.src[2], 0);                          // P1 other_net.2
        l1(Hh, 0, L.src[4], 1, L.src[5], 3, 1, Hh);                     phase_end();  //    integrator_net.0, units [0,H/2)
        l1(Hh, Hh, L.src[4], 1, L.src[5], 3, 1, Hh);                    phase_end();  // P2 integrator_net.0, units [H/2,H)
        bias(Hh, Hh, L.src[7], Hh); hid(Hh, Hh, L.src[6], Hh);          phase_end();  // P3 integrator_net.2
        bias(H, H, L.src[9], 0); hid(H, H, L.src[8], 0);                phase_end();  // P4 net.0
        bias(16, 1, L.src[11], 0); hid(16, 1, L.src[10], 0);            phase_end();  // P5 net.2
    } else {
        // state_dict order: net.0.{w,b} net.2.{w,b} net.4.{w,b} net.6.{w,b}
        const int sizes[8] = {H * S, H, H * H, H, H * H, H, H, 1};
        int o = 0;
        for (int j = 0; j < 8; ++j) { L.src[j] = o; o += sizes[j]; }
        L.param_count = o;
        L.nin = S;
        L.nterms = 3 * S + 2 <= 32 ? 3 : 2;
        L.KP = ((L.nterms * S + 2 + 15) / 16) * 16;
        if (L.KP * kRows * 2 > geo_abytes(H)) return false;  // the observation operand aliases the A tile
        l1(H, 0, L.src[0], S, L.src[1], 0, S, 0);                       phase_end();  // P0 net.0
        bias(H, H, L.src[3], 0); hid(H, H, L.src[2], 0);                phase_end();  // P1 net.2
        bias(H, H, L.src[5], 0); hid(H, H, L.src[4], 0);                phase_end();  // P2 net.4
        bias(16, 1, L.src[7], 0); hid(16, 1, L.src[6], 0);              phase_end();  // P3 net.6
    }
    if (!ok) return false;
    L.total_bytes = kHeaderBytes + L.f16_bytes;
    return true;
}

// what the kernels need of the layout (passed by value)
struct MlpParams {
    const uint8_t *pack;
    int nblk, S, nin, nterms, KP;
};

inline MlpParams make_mlp_params(const PackLayout &L, const void *pack) {
    MlpParams m;
    m.pack = (const uint8_t *)pack;
    m.nblk = L.nblk; m.S = L.S; m.nin = L.nin; m.nterms = L.nterms; m.KP = L.KP;
    return m;
}

template <int KIND, int H> struct Geo {
    static constexpr bool kModular = KIND == PIME_ACTOR_MODULAR;
    static constexpr bool kRelu = KIND == PIME_CRITIC_ADV;
    static constexpr int Hh = H / 2;
    static constexpr int ABytes = geo_abytes(H);
    static constexpr int ObsOff = geo_obs_off(KIND, H);
    static constexpr int OnesOff = geo_ones_off(KIND, H);
    static constexpr int RingOff = OnesOff + kK16Bytes;
    static constexpr int TblOff = RingOff + kStages * kMaxBlkBytes;
    static constexpr int BarOff = TblOff + kMaxBlocks * 16;
    static constexpr int SmemBytes = BarOff + 128;
    static constexpr int TmemCols = H < 32 ? 32 : H;
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


// 32 lanes x 16 / x1 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t *u = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t u;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(u) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(u);
}

__device__ __forceinline__ uint32_t pack_hh(__half a, __half b) {
    __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
// x = hi + lo with hi, lo in fp16 (relative error of hi + lo ~ 2^-22)
__device__ __forceinline__ void split_h(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// ------------------------------------------------------------------------------------------------ the engine
template <int KIND, int H> struct Engine {
    using G = Geo<KIND, H>;
    uint8_t *sA, *sRing;
    const Blk *tbl;
    uint64_t *full, *empty, *a_ready, *d_ready;
    uint32_t *tmem_slot;
    uint32_t tmem_base;
    uint32_t dph;  // parity of the next d_ready completion (env threads)
    MlpParams mp;

    // All kThreads threads.  Carves shared memory, initialises barriers, allocates TMEM, loads the block program and
    // writes the constant "ones" operand.
    __device__ __forceinline__ void setup(uint8_t *smem, const MlpParams &p) {
        mp = p;
        sA = smem;
        sRing = smem + G::RingOff;
        tbl = reinterpret_cast<const Blk *>(smem + G::TblOff);
        full = reinterpret_cast<uint64_t *>(smem + G::BarOff);
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
        const uint4 *src = reinterpret_cast<const uint4 *>(mp.pack);
        uint4 *dst = reinterpret_cast<uint4 *>(smem + G::TblOff);
        for (int j = tid; j < mp.nblk; j += kThreads) dst[j] = __ldg(src + j);
        // ones operand: K=16 slice whose first two columns are 1.0 (fp16 0x3C00): multiplies [b_hi b_lo 0 ...]
        for (int j = tid; j < 2 * kRows; j += kThreads)
            *reinterpret_cast<uint4 *>(smem + G::OnesOff + j * 16) = j < kRows ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
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
        const uint8_t *w16 = mp.pack + kHeaderBytes;
        const int nblk = mp.nblk;
        uint32_t it = 0;
        for (int s = 0; s < steps; ++s) {
            for (int b = 0; b < nblk; ++b, ++it) {
                const Blk B = tbl[b];
                const uint32_t st = it % kStages, ph = (it / kStages) & 1;
                const uint32_t bytes = (uint32_t)B.bytes16 * 16u;
                mbar_wait(&empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&full[st], bytes);
                bulk_g2s(sRing + st * kMaxBlkBytes, w16 + B.src_off, bytes, &full[st]);
            }
        }
    }

    // ---- warp 4 lane 0: interpret the block program, once per step
    __device__ __forceinline__ void mma_loop(int steps) {
        const int nblk = mp.nblk;
        uint32_t it = 0, aph = 0;
        const uint32_t a_base = smem_u32(sA);
        for (int s = 0; s < steps; ++s) {
            for (int b = 0; b < nblk; ++b, ++it) {
                const Blk B = tbl[b];
                if (B.flags & BLK_WAIT_A) {
                    mbar_wait(a_ready, aph);
                    aph ^= 1;
                    tc_fence_after();
                }
                const uint32_t st = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t nbytes16 = (uint32_t)B.nb8 * 8u * 16u;  // bytes of one K core-matrix column of B
                const uint32_t idesc = make_idesc(kRows, (int)B.nb8 * 8);
                const uint32_t b_base = smem_u32(sRing + st * kMaxBlkBytes);
                const uint32_t a_addr = a_base + (uint32_t)B.a_off16 * 16u;
                const uint32_t d_addr = tmem_base + (uint32_t)B.d_col;
                const uint32_t k16s = B.k16s;
                uint32_t acc = (B.flags & BLK_FRESH) ? 0u : 1u;
                for (uint32_t kk = 0; kk < k16s; ++kk) {
                    const uint64_t adesc = make_desc(a_addr + kk * kK16Bytes, kChunkBytes, 128);
                    const uint64_t bdesc = make_desc(b_base + kk * 2u * nbytes16, nbytes16, 128);
                    mma_f16(d_addr, adesc, bdesc, idesc, acc);
                    acc = 1u;
                }
                mma_commit(&empty[st]);  // frees the ring slot once these MMAs have read it
                if (B.flags & BLK_COMMIT_D) mma_commit(d_ready);  // accumulators of this phase complete, A free again
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

    // epilogue of a layer: A[:, acol + c] = act(D[:, dcol + c]) for c in [0, ncols); the bias is already in D
    template <int NCOLS> __device__ __forceinline__ void epilogue_to_a(int row, int dcol, int acol) {
        const uint32_t taddr = tmem_base + ((uint32_t)((row / 32) * 32) << 16) + (uint32_t)dcol;
        if constexpr (NCOLS % 32 == 0) {
#pragma unroll 2
            for (int c0 = 0; c0 < NCOLS; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = act_fn<G::kRelu>(v[j]);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    a_store8(row, (acol + c0) / 8 + q, pack_h2(v[q * 8 + 0], v[q * 8 + 1]), pack_h2(v[q * 8 + 2], v[q * 8 + 3]),
                             pack_h2(v[q * 8 + 4], v[q * 8 + 5]), pack_h2(v[q * 8 + 6], v[q * 8 + 7]));
            }
        } else {
            static_assert(NCOLS % 16 == 0, "layer width");
#pragma unroll
            for (int c0 = 0; c0 < NCOLS; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = act_fn<G::kRelu>(v[j]);
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    a_store8(row, (acol + c0) / 8 + q, pack_h2(v[q * 8 + 0], v[q * 8 + 1]), pack_h2(v[q * 8 + 2], v[q * 8 + 3]),
                             pack_h2(v[q * 8 + 4], v[q * 8 + 5]), pack_h2(v[q * 8 + 6], v[q * 8 + 7]));
            }
        }
    }

    // first-layer operand of this env: [in_hi | in_lo | in_hi (3 terms) | 1 1 | 0 ...]  (see the file header)
    __device__ __forceinline__ void write_obs(int row, const float *obs) {
        uint8_t *dst = sA + G::ObsOff + row * 16;
        if constexpr (G::kModular) {
            const int So = mp.S - 1;
            __half h[4], l[4];
            split_h(obs[0], h[0], l[0]);
            split_h(So > 1 ? obs[1] : 0.0f, h[1], l[1]);
            split_h(So > 2 ? obs[2] : 0.0f, h[2], l[2]);
            split_h(obs[So], h[3], l[3]);
            const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
            *reinterpret_cast<uint4 *>(dst) = make_uint4(pack_hh(h[0], h[1]), pack_hh(h[2], h[3]), pack_hh(l[0], l[1]), pack_hh(l[2], l[3]));
            *reinterpret_cast<uint4 *>(dst + kChunkBytes) =
                make_uint4(pack_hh(h[0], h[1]), pack_hh(h[2], h[3]), pack_hh(one, one), pack_hh(zero, zero));
        } else {
            __align__(16) __half hl[kMaxKP];
            const int S = mp.S, KP = mp.KP, nt = mp.nterms;
#pragma unroll 1
            for (int k = 0; k < KP; ++k) hl[k] = __float2half_rn(0.0f);
#pragma unroll 1
            for (int k = 0; k < S; ++k) {
                __half hi, lo;
                split_h(obs[k], hi, lo);
                hl[k] = hi;
                hl[S + k] = lo;
                if (nt == 3) hl[2 * S + k] = hi;
            }
            hl[nt * S] = __float2half_rn(1.0f);
            hl[nt * S + 1] = __float2half_rn(1.0f);
            const uint4 *q = reinterpret_cast<const uint4 *>(hl);
            for (int kc = 0; kc < KP / 8; ++kc) *reinterpret_cast<uint4 *>(dst + (size_t)kc * kChunkBytes) = q[kc];
        }
    }

    // Full forward for the env owned by this thread; obs = float32 observation (S values).  All 128 env threads of
    // the CTA call this together.  Returns net(obs) (pre-tanh, pre-prior).
    __device__ __forceinline__ float forward(int row, const float *obs) {
        constexpr int Hh = G::Hh;
        write_obs(row, obs);
        signal_a();
        if constexpr (G::kModular) {
            wait_d();                               // P0: D[0:H] = other_net.0 pre-activation (net_residual.py:151)
            epilogue_to_a<H>(row, 0, 0);
            signal_a();
            wait_d();                               // P1: D[0:H/2] = other_net.2 (:152), D[H/2:H] = integrator_net.0 units [0,H/2) (:154)
            epilogue_to_a<Hh>(row, Hh, 0);
            signal_a();
            wait_d();                               // P2: D[H/2:H] = integrator_net.0 units [H/2,H)
            epilogue_to_a<Hh>(row, Hh, Hh);
            signal_a();
            wait_d();                               // P3: D[H/2:H] = integrator_net.2 (:155)
            epilogue_to_a<H>(row, 0, 0);            // cat(tanh(other), tanh(integrator)) (:170)
            signal_a();
            wait_d();                               // P4: D[0:H] = net.0 (:157)
            epilogue_to_a<H>(row, 0, 0);
            signal_a();
            wait_d();                               // P5: D[:, 0] = net.2 (:158)
        } else {
            wait_d();
            epilogue_to_a<H>(row, 0, 0);
            signal_a();
            wait_d();
            epilogue_to_a<H>(row, 0, 0);
            signal_a();
            wait_d();
            epilogue_to_a<H>(row, 0, 0);
            signal_a();
            wait_d();
        }
        const float out = tmem_ld1(tmem_base + ((uint32_t)((row / 32) * 32) << 16));
        tc_fence_before();  // orders this read before the next step's MMAs (they follow the a_ready arrive)
        return out;
    }
};

}  // namespace tc
}  // namespace pime
