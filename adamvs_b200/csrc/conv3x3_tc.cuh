// 3x3 stride-1 convolutions of the recurrent regulariser on the 5th-generation tensor cores: tcgen05.mma kind::tf32,
// accumulators in TMEM, at fp32 accuracy through an exact hi/lo operand split ("3xTF32", here all four partial products).
//
// Implicit GEMM without an im2col copy ("padded linear"): a tile's input (TH+2 rows x 34 columns, halo included) sits
// in shared memory as [8-channel chunk][hi|lo][channel quad][position][4 channels] - positions enumerate the padded
// tile row-major with pitch IPO = 34, 16 bytes per position, which is the canonical K-major no-swizzle UMMA operand
// (core matrix = 8 positions x 16 B, SBO = 128 B, LBO = one quad plane).  Output position m = oy*IPO + ox reads tap
// (ky,kx) at position m + ky*IPO + kx, so every tap is the SAME operand at a 16-byte-granular start-address offset:
// per 128 output positions and 8 input channels, 9 taps x {hi, lo} MMAs of M=128 x N x K=8.  Output positions with
// ox >= 32 are garbage rows of D (6 %) that the epilogue skips.
//
// fp32 accuracy: kind::tf32 reads the top 19 bits of each fp32 word.  Activations and weights are split into
// hi = rna_tf32(x), lo = rna_tf32(x - hi) (|x - hi - lo| <= 2^-23 |x|); the B operand carries [W_hi rows | W_lo rows]
// (N = 2*Cout), A_hi and A_lo are multiplied with it in turn and the epilogue adds the two column halves:
// (A_hi + A_lo)(W_hi + W_lo), fp32 accumulation in TMEM.  An M=128 x K=8 tf32 MMA is bound by the 4 KB shared-memory
// read of its A tile (32 clk) for every N <= 64, so the doubled N and the lo*lo term are free.
// PREC_TF32 drops the A_lo pass (activations rounded to tf32, weights still exact): half the MMA time, reported
// separately with its own tolerance.
//
// Warp-specialised persistent CTA (one per SM, 288 threads), static round-robin tile schedule:
//   warps 4-7       converters: read the tile's input (halo included, out-of-image positions = 0 = conv padding)
//                   straight from global memory - 8 channels x 5 positions per thread, requested one chunk ahead so
//                   that the DRAM latency hides behind the MMAs of the previous chunk - split every value and write
//                   the hi/lo quad-interleaved operand stage (2 + 2 STS.128 per position, conflict-free).  Shared
//                   memory is the contended resource (the MMAs read ~92 B/clk of operands), so the input makes no
//                   detour through it: an earlier TMA box ring + LDS transpose cost 2.2x the LSU/TMA traffic
//   warp 8 lane 0   MMA issuer: waits for operand stages, issues MT*9*{hi,lo} MMAs per chunk, tcgen05.commit hands the
//                   stage back to the converters and, after a tile's last chunk, the accumulator to the epilogue
//   warps 0-3       epilogue of the PREVIOUS tile while the MMAs of the current tile run: tcgen05.ld of their TMEM lane
//                   quarter; the GRU state / gate operands are requested one step ahead (the first step's before the
//                   accumulator is awaited), bias + gate non-linearity + GRU blend, coalesced row stores
// Accumulators are double buffered in TMEM (2 x MT x 2*Cout columns).
// Measured (tools/umma_rate_probe.cu): one M=128 x N<=64 x K=8 kind::tf32 MMA with both operands in shared memory takes
// 49 clk (operand fetch bound; 64 clk at N=128, 128 at N=256), i.e. 72 x 49 = 3.5 kclk per 8-channel chunk of a
// 32x15 tile - that is the floor of this formulation.
#pragma once
#include "conv3x3.cuh"

namespace adamvs {

enum { PREC_FP32X3 = 0, PREC_TF32 = 1 };

// Debug builds only (ADAMVS_TC_TRACE=1 python adamvs_b200/build.py --force): block 0 of the GRU-1 gate convolution
// records clock64() stamps of its three roles; tools/tc_trace.py prints the timeline.
#ifdef ADAMVS_TC_TRACE
__device__ long long g_tc_trace[3][512][4];
#define TC_TRACE(role, idx, f) do { if (CA == 8 && CB == 8 && COUT == 16 && PREC == PREC_FP32X3 && blockIdx.x == 0 && (idx) < 512) g_tc_trace[role][idx][f] = clock64(); } while (0)
#else
#define TC_TRACE(role, idx, f) do { } while (0)
#endif

template <int MT_>
struct TcGeom {
    static constexpr int MT = MT_;                               // M tiles (128 output positions each) per tile
    static constexpr int TW = 32, IPO = TW + 2;                  // operand pitch: tile + left/right halo
    static constexpr int TH = MT * 128 / IPO;                    // 15 rows (MT = 4) | 7 rows (MT = 2)
    static constexpr int IH = TH + 2;
    static constexpr int NPOS = (MT * 128 + 2 * IPO + 2 + 7) / 8 * 8;   // positions any tap of any M row can touch
    static_assert(IH * IPO <= NPOS, "operand plane too small");
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                                      // descriptor version 1 (Blackwell); no swizzle, base offset 0
    return d;
}
// D = f32, A = B = tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a broken pipeline traps (kills the context with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
// exact two-term tf32 split, round to nearest on both terms
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    const float r = v - hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
    lo = __uint_as_float(l);
}

template <int CA, int CB, int COUT, int MT, int PREC>
struct TcCfg {
    using G = TcGeom<MT>;
    static constexpr int CIN = CA + CB, NCH = CIN / CK;           // 8-channel chunks = K steps per tap
    static constexpr int NB = 2 * COUT;                            // MMA N: [W_hi rows | W_lo rows]
    static constexpr int NHL = PREC == PREC_FP32X3 ? 2 : 1;        // operand passes per stage (hi, lo)
    static constexpr int PLANE_BYTES = G::NPOS * 16;               // one channel quad of one pass
    static constexpr int STAGE_BYTES = NHL * 2 * PLANE_BYTES;      // [hi|lo][2 quads][NPOS][4]
    static constexpr int B_STEP_BYTES = 2 * NB * 16;               // [2 quads][NB rows][4]
    static constexpr int B_BYTES = 9 * NCH * B_STEP_BYTES;
    static constexpr int BUDGET = 227 * 1024 - 512 - B_BYTES;
    static constexpr int NA_FIT = BUDGET / STAGE_BYTES;
    static constexpr int NA = NA_FIT > 4 ? 4 : NA_FIT;             // operand stages
    static constexpr int ACC_COLS = MT * NB;                       // TMEM columns of one accumulator buffer
    static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;
    static constexpr int NBAR = 2 * NA + 4;
    static constexpr size_t SMEM = (size_t)NA * STAGE_BYTES + B_BYTES + 8 * NBAR + 16;
    static_assert(CA % CK == 0 && CB % CK == 0, "channel groups must be chunk aligned");
    static_assert(NB % 16 == 0 && NB <= 256, "M = 128 MMAs need N % 16 == 0");
    static_assert(NA >= 2, "needs two operand stages");
    static_assert(2 * ACC_COLS <= 512, "does not fit TMEM");
    static_assert(SMEM <= 227 * 1024, "does not fit shared memory");
    // one CTA per SM is what keeps a 512-column allocation from blocking a co-resident CTA forever
    static_assert(TMEM_COLS <= 256 || SMEM > 114 * 1024, "512-column configurations must be alone on their SM");
};

constexpr int kTcThreads = 288;     // warps 0-3 epilogue (TMEM lane quarters 0-3), 4-7 converters, 8 MMA issuer

template <int CA, int CB, int COUT, int EPI, int MT, int PREC>
__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_tc_kernel(ConvArgs a, TileGrid tg) {
    using C = TcCfg<CA, CB, COUT, MT, PREC>;
    using G = TcGeom<MT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;                                            // [NA] operand stages
    unsigned char* sB = sA + C::NA * C::STAGE_BYTES;                         // [9][NCH][2][NB][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::B_BYTES);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + C::NA;
    uint64_t* d_full = a_empty + C::NA;
    uint64_t* d_empty = d_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
    __shared__ float sBias[COUT];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < C::NA; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 4); }
        fence_mbar_init();
    }
    if (tid < COUT) sBias[tid] = (EPI == EPI_RELU || a.bias == nullptr) ? 0.f : __ldg(a.bias + tid);
    // resident weights: [ci][tap][co] -> B operand [tap][chunk][quad][W_hi co | W_lo co][4 ci], split once per CTA
    for (int i = tid; i < 9 * C::NCH * 2 * COUT * 4; i += kTcThreads) {
        const int j = i & 3, n = (i >> 2) % COUT, kq = (i / (4 * COUT)) & 1, s = (i / (8 * COUT)) % C::NCH, t = i / (8 * COUT * C::NCH);
        const float v = __ldg(a.wpk + ((size_t)(8 * s + 4 * kq + j) * 9 + t) * COUT + n);
        float hi, lo;
        split_tf32(v, hi, lo);
        float* dst = reinterpret_cast<float*>(sB + (size_t)(t * C::NCH + s) * C::B_STEP_BYTES + kq * C::NB * 16);
        dst[n * 4 + j] = hi;
        dst[(COUT + n) * 4 + j] = lo;
    }
    // positions past the tile are read only by garbage rows of D; give them finite values once
    for (int i = tid; i < C::NA * C::STAGE_BYTES / 16; i += kTcThreads) reinterpret_cast<float4*>(sA)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int my_tiles = (tg.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * C::NCH;                                      // chunk stream of this CTA
    const int tiles_per_item = tg.tiles_x * tg.tiles_y;

    if (warp == 8) {
        // ===== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(C::NB);
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            int g = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int acc = ti & 1;
                if (ti >= 2) mbar_wait_bounded(&d_empty[acc], ((ti >> 1) - 1) & 1);
                for (int c = 0; c < C::NCH; ++c, ++g) {
                    const int st = g % C::NA;
                    TC_TRACE(1, g, 0);
                    mbar_wait_bounded(&a_full[st], (g / C::NA) & 1);
                    tc_fence_after();
                    TC_TRACE(1, g, 1);
                    // descriptors differ only in their 14-bit start-address field (16-byte units; shared memory is
                    // < 256 KB, so adding offsets never carries out of the field)
                    const uint64_t bd0 = umma_desc(b_base + (uint32_t)(c * C::B_STEP_BYTES), C::NB * 16, 128);
                    // M tile innermost: consecutive MMAs accumulate into different TMEM tiles.  Back-to-back MMAs into
                    // the SAME accumulator serialise on its read-after-write latency (44-49 clk each for any N <= 64,
                    // tools/umma_rate_probe.cu); MT independent chains hide it.
                    const uint64_t ad00 = umma_desc(a_base + (uint32_t)st * C::STAGE_BYTES, C::PLANE_BYTES, 128);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint64_t bd = bd0 + (uint64_t)((t * C::NCH * C::B_STEP_BYTES) >> 4);
#pragma unroll
                        for (int hl = 0; hl < C::NHL; ++hl) {
                            const uint64_t ad0 = ad00 + (uint64_t)((hl * 2 * C::PLANE_BYTES + ((t / 3) * G::IPO + (t % 3)) * 16) >> 4);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                const uint32_t d = tmem + (uint32_t)(acc * C::ACC_COLS + mt * C::NB);
                                umma_tf32(d, ad0 + (uint64_t)((mt * 128 * 16) >> 4), bd, idesc, (t | hl) ? 1u : (c ? 1u : 0u));
                            }
                        }
                    }
                    umma_commit(&a_empty[st]);                                 // operand stage free once these MMAs retire
                    TC_TRACE(1, g, 2);
                }
                umma_commit(&d_full[acc]);                                     // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== converters: global memory -> registers (one chunk ahead) -> hi/lo quad-interleaved operand stage
        const int ct = tid - 128;
        constexpr int NPP = G::IH * G::IPO;                                  // positions of one quad plane
        constexpr int NJ = (NPP + 127) / 128;
        int pr[NJ], pc[NJ];                                                  // tile-relative (row, column) of this thread's positions
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int pos = ct + 128 * j;
            pr[j] = pos / G::IPO;
            pc[j] = pos - pr[j] * G::IPO;
        }
        const size_t in_plane = (size_t)a.hin * a.win;
        float pv[NJ][CK];                                                    // the chunk in flight
        auto fetch = [&](int g) {
            const int ti = g / C::NCH, c = g - ti * C::NCH;
            const int tile = blockIdx.x + ti * gridDim.x;
            const int b = tile / tiles_per_item, r = tile - b * tiles_per_item;
            const int ix0 = (r % tg.tiles_x) * G::TW - 1, iy0 = (r / tg.tiles_x) * G::TH - 1;
            const bool fromA = c * CK < CA;
            const float* base = fromA ? a.inA + (size_t)a.k * in_plane + (size_t)b * a.strideA_b + (size_t)(c * CK) * a.strideA_c
                                      : a.inB + (size_t)b * a.strideB_b + (size_t)(c * CK - CA) * a.strideB_c;
            const size_t cs = fromA ? (size_t)a.strideA_c : (size_t)a.strideB_c;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int gy = iy0 + pr[j], gx = ix0 + pc[j];
                const bool in = (j < NJ - 1 || ct + 128 * j < NPP) && gy >= 0 && gy < a.hin && gx >= 0 && gx < a.win;
                const float* p = base + (in ? (size_t)gy * a.win + gx : 0);
#pragma unroll
                for (int e = 0; e < CK; ++e) pv[j][e] = in ? __ldg(p + e * cs) : 0.f;
            }
        };
        if (total > 0) fetch(0);
#pragma unroll 1
        for (int g = 0; g < total; ++g) {
            const int st = g % C::NA;
            if (ct == 0) TC_TRACE(0, g, 0);
            if (g >= C::NA) mbar_wait_bounded(&a_empty[st], ((g / C::NA) - 1) & 1);
            if (ct == 0) TC_TRACE(0, g, 1);
            unsigned char* stage = sA + (size_t)st * C::STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int pos = ct + 128 * j;
                if (j == NJ - 1 && pos >= NPP) break;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float4 h4, l4;
                    split_tf32(pv[j][4 * q + 0], h4.x, l4.x); split_tf32(pv[j][4 * q + 1], h4.y, l4.y);
                    split_tf32(pv[j][4 * q + 2], h4.z, l4.z); split_tf32(pv[j][4 * q + 3], h4.w, l4.w);
                    float4* dst = reinterpret_cast<float4*>(stage + (size_t)q * C::PLANE_BYTES) + pos;
                    *dst = h4;
                    if (PREC == PREC_FP32X3) *(dst + 2 * G::NPOS) = l4;
                }
            }
            fence_proxy_async();                                               // operand writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[st]);
            if (ct == 0) TC_TRACE(0, g, 2);
            if (g + 1 < total) fetch(g + 1);                                   // lands while this thread waits for the next stage
            if (ct == 0) TC_TRACE(0, g, 3);
        }
    } else {
        // ===== epilogue (warps 0-3 = TMEM lane quarters 0-3): runs one tile behind the MMAs
        const size_t plane = (size_t)a.hout * a.wout;
        constexpr int CG = COUT < 16 ? COUT : 16;                            // output channels per TMEM load pair
        constexpr int NCG = COUT / CG, NG = MT * NCG;                        // (M tile, channel group) steps per tile
        constexpr int HC = COUT / 2;
#pragma unroll 1
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int acc = ti & 1;
            const int tile = blockIdx.x + ti * gridDim.x;
            const int b = tile / tiles_per_item, rr = tile - b * tiles_per_item;
            const int ox0 = (rr % tg.tiles_x) * G::TW, oy0 = (rr / tg.tiles_x) * G::TH;
            auto geom = [&](int mt, bool& valid, size_t& pix) {
                const int m = mt * 128 + warp * 32 + lane;
                const int ry = m / G::IPO, rx = m - ry * G::IPO;
                const int oy = oy0 + ry, ox = ox0 + rx;
                valid = ry < G::TH && rx < G::TW && oy < a.hout && ox < a.wout;
                pix = valid ? (size_t)oy * a.wout + ox : 0;
            };
            // GRU state / gate operands of step gi: requested one step ahead (the first step's before the accumulator
            // is awaited), so their DRAM latency overlaps the MMAs / the previous step instead of stalling every step
            auto preload = [&](int gi, float (&hs)[CG], float (&us)[CG]) {
                if (EPI != EPI_GATES && EPI != EPI_CAND) return;
                const int mt = gi / NCG, c0 = (gi - mt * NCG) * CG;
                bool valid; size_t pix;
                geom(mt, valid, pix);
                if (EPI == EPI_GATES) {
#pragma unroll
                    for (int c = 0; c < CG; ++c)
                        hs[c] = (valid && c0 + c < HC) ? __ldg(a.hstate + ((size_t)b * HC + c0 + c) * plane + pix) : 0.f;
                } else {
#pragma unroll
                    for (int c = 0; c < CG; ++c) {
                        const size_t o = ((size_t)b * COUT + c0 + c) * plane + pix;
                        us[c] = valid ? __ldg(a.ugate + o) : 0.f;
                        hs[c] = valid ? a.hstate[o] : 0.f;
                    }
                }
            };
            auto process = [&](int gi, const float (&hs)[CG], const float (&us)[CG]) {
                const int mt = gi / NCG, c0 = (gi - mt * NCG) * CG;
                bool valid; size_t pix;
                geom(mt, valid, pix);
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * C::ACC_COLS + mt * C::NB);
                uint32_t rh[16], rl[16];
                float v[CG];
                if (COUT >= 16) {
                    tmem_ld16_issue(taddr + c0, rh);                           // columns of A x W_hi
                    tmem_ld16_issue(taddr + COUT + c0, rl);                    // columns of A x W_lo
                } else {                                                       // COUT == 8: both halves in one 16-column load
                    tmem_ld16_issue(taddr, rh);
                }
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < CG; ++c)
                    v[c] = COUT >= 16 ? __uint_as_float(rh[c]) + __uint_as_float(rl[c]) : __uint_as_float(rh[c]) + __uint_as_float(rh[(COUT + c) & 15]);
                if (!valid) return;
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    const int co = c0 + c;
                    if (EPI == EPI_GATES) {
                        const float s = sigmoid_f(v[c] + sBias[co]);
                        if (co < HC) a.out0[((size_t)b * HC + co) * plane + pix] = s * hs[c];      // reset gate -> r*h
                        else a.out1[((size_t)b * HC + (co - HC)) * plane + pix] = s;
                    } else if (EPI == EPI_CAND) {
                        a.out0[((size_t)b * COUT + co) * plane + pix] = us[c] * hs[c] + (1.f - us[c]) * tanh_f(v[c] + sBias[co]);
                    } else if (EPI == EPI_RELU) {
                        a.out0[((size_t)b * COUT + co) * plane + pix] = fmaxf(v[c], 0.f);
                    } else {                                                   // EPI_BIAS
                        const float y = v[c] + sBias[co];
                        a.out0[((size_t)b * COUT + co) * plane + pix] = a.relu ? fmaxf(y, 0.f) : y;
                    }
                }
            };
            float hs0[CG], us0[CG], hs1[CG], us1[CG];
            preload(0, hs0, us0);
            if (tid == 0) TC_TRACE(2, ti, 0);
            mbar_wait_bounded(&d_full[acc], (ti >> 1) & 1);
            tc_fence_after();
            if (tid == 0) TC_TRACE(2, ti, 1);
#pragma unroll
            for (int gi = 0; gi < NG; gi += 2) {
                if (gi + 1 < NG) preload(gi + 1, hs1, us1);
                process(gi, hs0, us0);
                if (gi + 1 < NG) {
                    if (gi + 2 < NG) preload(gi + 2, hs0, us0);
                    process(gi + 1, hs1, us1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[acc]);
            if (tid == 0) TC_TRACE(2, ti, 2);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
}

template <int CA, int CB, int COUT, int EPI>
struct TcLayer {
    // plane sizes below ~2 waves of 15-row tiles use 7-row tiles (MT = 2)
    static int choose_mt(int hout, int wout, int B) {
        const long long t4 = (long long)((wout + 31) / 32) * ((hout + TcGeom<4>::TH - 1) / TcGeom<4>::TH) * B;
        return t4 >= 2LL * sm_count() ? 4 : 2;
    }
    static bool plan(ConvPlan& p, const ConvArgs& a, int B, int depthA) {
        (void)depthA;
        p.args = a;
        p.cfg = choose_mt(a.hout, a.wout, B);
        return true;
    }
    template <int MT, int PREC>
    static cudaError_t launch_cfg(ConvPlan& p, int B, cudaStream_t st) {
        using C = TcCfg<CA, CB, COUT, MT, PREC>;
        auto kern = conv3x3_tc_kernel<CA, CB, COUT, EPI, MT, PREC>;
        static bool ready[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        dev = dev < 64 ? dev : 63;
        if (!ready[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            if (e != cudaSuccess) return e;
            ready[dev] = true;
        }
        p.tg.tiles_x = (p.args.wout + TcGeom<MT>::TW - 1) / TcGeom<MT>::TW;
        p.tg.tiles_y = (p.args.hout + TcGeom<MT>::TH - 1) / TcGeom<MT>::TH;
        p.tg.ntiles = p.tg.tiles_x * p.tg.tiles_y * B;
        int ctas = sm_count();
        if (ctas > p.tg.ntiles) ctas = p.tg.ntiles;
        kern<<<dim3(ctas, 1, 1), kTcThreads, C::SMEM, st>>>(p.args, p.tg);
        return cudaGetLastError();
    }
    static cudaError_t launch(ConvPlan& p, int B, int prec, cudaStream_t st) {
        if (p.cfg == 4) return prec == PREC_TF32 ? launch_cfg<4, PREC_TF32>(p, B, st) : launch_cfg<4, PREC_FP32X3>(p, B, st);
        return prec == PREC_TF32 ? launch_cfg<2, PREC_TF32>(p, B, st) : launch_cfg<2, PREC_FP32X3>(p, B, st);
    }
};

}  // namespace adamvs
