// 3x3 stride-1 convolutions of the recurrent regulariser on the 5th-generation tensor cores: tcgen05.mma kind::tf32,
// accumulators in TMEM, at fp32 accuracy through a hi/lo operand split ("3xTF32": the three partial products that matter).
//
// Implicit GEMM without an im2col copy.  A tile's input (4*MT + 2 rows x 32 columns, halo included) sits in shared
// memory as [8-channel chunk][hi|lo][channel quad][position][4 channels], positions row-major with pitch 32, 16 bytes per
// position - the canonical K-major no-swizzle UMMA operand (core matrix = 8 positions x 16 B, SBO = 128 B, LBO = one
// quad plane).  The three ROW taps are operand address offsets: M tile mt, tap row ky reads positions
// mt*128 + ky*32 + [0,128) - a core-matrix aligned start address.  The three COLUMN taps sit in the MMA's N:
//     D'[p][kx][co] = sum_ky sum_ci in[p + 32 ky][ci] * W[co][ci][ky][kx]            (N = 3 x 2*Cout)
//     out[p][co]    = D'[p][0][co] + D'[p+1][1][co] + D'[p+2][2][co]
// and the shift-and-add over kx is two warp shuffles per value in the epilogue (a warp = one 32-position row of D';
// its lanes 30, 31 are the halo columns, so a tile is 30 pixels wide).
// Why: one kind::tf32 MMA of M = 128, K = 8 costs 44-49 clk for ANY N <= ~96 (tools/umma_rate_probe.cu,
// tools/umma_ts_probe.cu: same with A in TMEM, same with interleaved accumulators) and N/2 clk above that.  With every
// tap its own MMA (9 taps x {hi,lo} x N = 2*Cout <= 64) the layers ran at 18 x 49 clk per 128 positions and 8 channels,
// operand-issue bound at 13-34 % tensor-pipe activity (profiles/r01q, r01t).  Folding kx into N issues 3x fewer MMAs
// (6 per 128 positions and chunk: 49 clk each at N = 48 / 96, 96 clk at N = 192).
//
// fp32 accuracy: kind::tf32 reads the top 19 bits of each fp32 word.  Activations and weights are split into
// hi = rna_tf32(x), lo = rna_tf32(x - hi) (|x - hi - lo| <= 2^-23 |x|); the B operand carries [W_hi rows of the three
// kx | W_lo rows of the three kx]; A_hi is multiplied with all of it, A_lo with the W_hi rows only (a smaller N on the
// same operand) and the epilogue adds the two column halves: A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 accumulation in
// TMEM.  The dropped A_lo W_lo term is <= 2^-22 |a||w| per product, below the rounding of the fp32 accumulation itself
// (round 2: it had been computed too - all four products; dropping it shortens the second pass from N = 6 Cout to
// 3 Cout: 11-15 % less tensor-pipe time and operand traffic).  PREC_TF32 drops the A_lo pass altogether (activations
// rounded to tf32, weights still exact), reported separately with its own tolerance.
//
// Warp-specialised persistent CTA (one per SM, 512 threads), static round-robin tile schedule:
//   warp 8 lane 0   TMA producer: planar [8 ch][4*MT+2][36] boxes (out-of-image elements zero-filled = conv padding)
//                   into a ring deep enough to cover the DRAM latency
//   warp 10 lane 0  TMA producer of the GRU epilogue operands (state h for the reset gates, update gate u and h for the
//                   candidate blend) of the OUTPUT tile, into their own ring (released by the epilogue)
//   warps 4-7       converters: planar slot -> split -> hi/lo quad-interleaved operand stage (4 conflict-free LDS.32,
//                   2 + 2 STS.128 per position; a warp = one row of 32 positions)
//   warp 9 lane 0   MMA issuer: per chunk 3 (ky) x {hi,lo} x MT MMAs; tcgen05.commit hands the operand stage back to
//                   the converters and, after a tile's last chunk, the accumulator to the epilogue
//   warps 0-3,12-15 epilogue of the PREVIOUS tile while the MMAs of the current tile run (accumulators double buffered
//                   in TMEM, 2 x MT x 3 x 2*Cout columns): tcgen05.ld of their lane quarter (= image row mt*4 + warp%4),
//                   kx shift-and-add by shuffle, bias + gate non-linearity + GRU blend with operands from the
//                   producer's ring, all values first and all row stores afterwards; the two warp sets take alternate
//                   (M tile, 8-channel group) steps
#pragma once
#include "conv3x3.cuh"

namespace adamvs {

enum { PREC_FP32X3 = 0, PREC_TF32 = 1 };

// Debug builds only (ADAMVS_TC_TRACE=1 python adamvs_b200/build.py --force): block 0 of the GRU-1 gate convolution
// records clock64() stamps of its three roles; tools/tc_trace.py prints the timeline.
#ifdef ADAMVS_TC_TRACE
static __device__ long long g_tc_trace[4][512][4];
#define TC_TRACE(role, idx, f) do { if (CA == 8 && CB == 8 && COUT == 16 && PREC == PREC_FP32X3 && blockIdx.x == 0 && (idx) < 512) g_tc_trace[role][idx][f] = clock64(); } while (0)
#else
#define TC_TRACE(role, idx, f) do { } while (0)
#endif

template <int MT_>
struct TcGeom {
    static constexpr int MT = MT_;                               // M tiles (128 positions = 4 rows x 32) per tile
    static constexpr int TW = 30, PW = 32;                       // output columns per tile | positions per row (+ 2 halo)
    static constexpr int TH = 4 * MT, IH = TH + 2;
    static constexpr int BOXW = 36;                              // TMA box: the start column is rounded down to 4 floats
    static constexpr int BOX_FLOATS = CK * IH * BOXW;
    static constexpr int NPOS = IH * PW;                         // positions of one quad plane; the MMAs read exactly these
    static_assert((BOX_FLOATS * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");
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
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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

// Activation split for the converters (runs 8 x 10^9 times per depth map): hi = x rounded to tf32 by integer
// arithmetic (add half an ulp, clear the 13 low bits: 2 instructions; cvt.rna.tf32 compiles to ~5), lo = x - hi (exact in
// fp32).  lo is handed over unrounded: the tensor core drops its low 13 bits, an error <= 2^-21 |x|.
__device__ __forceinline__ void split_fast(float v, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
    lo = v - hi;
}

template <int CA, int CB, int COUT, int MT, int PREC, int EPI>
struct TcCfg {
    using G = TcGeom<MT>;
    static constexpr int CIN = CA + CB, NCH = CIN / CK;           // 8-channel chunks = K steps per tap row
    static constexpr int NB = 2 * COUT;                            // [W_hi rows | W_lo rows] of one kx
    static constexpr int N3 = 3 * NB;                              // MMA N: [W_hi: kx0 kx1 kx2 | W_lo: kx0 kx1 kx2] x COUT rows
    static constexpr int NLO = (3 * COUT + 15) / 16 * 16;          // N of the A_lo pass: the W_hi rows
    static constexpr int NHL = PREC == PREC_FP32X3 ? 2 : 1;        // operand passes per stage (hi, lo)
    static constexpr int PLANE_BYTES = G::NPOS * 16;               // one channel quad of one pass
    static constexpr int STAGE_BYTES = NHL * 2 * PLANE_BYTES;      // [hi|lo][2 quads][NPOS][4]
    static constexpr int B_STEP_BYTES = 2 * N3 * 16;               // [2 quads][N3 rows][4] of one (ky, chunk)
    static constexpr int B_BYTES = 3 * NCH * B_STEP_BYTES;
    static constexpr int SLOT_BYTES = G::BOX_FLOATS * 4;
    // operand stages: four where at least four TMA slots still fit beside them (the converters waited ~200 clk per chunk
    // for a free stage with three), else three
    static constexpr int EPI_TILE_BYTES0 = (EPI == EPI_GATES ? COUT / 2 : (EPI == EPI_CAND ? 2 * COUT : 0)) * G::TH * 32 * 4;
    static constexpr int NA = (227 * 1024 - 1024 - B_BYTES - 4 * STAGE_BYTES - 3 * EPI_TILE_BYTES0) / SLOT_BYTES >= 4 ? 4 : 3;
    // GRU epilogue operands (h for the reset gates; u and h for the candidate blend) travel like the input: the producer
    // streams [channels][4*MT rows][32] boxes of the output tile into a ring EPI_R tiles deep
    static constexpr int NG_STEPS = MT * (COUT / 8);               // (M tile, 8-channel group) epilogue steps per tile
    static constexpr int EPI_CH = EPI == EPI_GATES ? COUT / 2 : (EPI == EPI_CAND ? 2 * COUT : 0);
    static constexpr int EPI_PLANE = G::TH * 32;                   // floats per channel of a box
    static constexpr int EPI_TILE_BYTES = EPI_CH * EPI_PLANE * 4;
    static constexpr int BUDGET0 = 227 * 1024 - 1024 - B_BYTES - NA * STAGE_BYTES;
    static constexpr int EPI_R = EPI_CH == 0 ? 0 : ((BUDGET0 - 3 * EPI_TILE_BYTES) / SLOT_BYTES >= 4 ? 3 : 2);
    static constexpr int EPI_BYTES = EPI_R * EPI_TILE_BYTES;
    static constexpr int BUDGET = BUDGET0 - EPI_BYTES;
    static constexpr int NSLOT_FIT = BUDGET / SLOT_BYTES;
    static constexpr int NSLOT = NSLOT_FIT > 6 ? 6 : NSLOT_FIT;    // planar TMA ring
    static constexpr int ACC_COLS = MT * N3;                       // TMEM columns of one accumulator buffer
    static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;
    static constexpr int NBAR = 2 * NSLOT + 2 * NA + 4 + 2 * 3;
    static constexpr size_t SMEM_USED = (size_t)NSLOT * SLOT_BYTES + (size_t)NA * STAGE_BYTES + B_BYTES + EPI_BYTES + 8 * NBAR + 16;
    // every configuration allocates more than half of TMEM: ask for more than half of the shared memory too, so that a
    // second CTA can never become resident on the SM and spin forever in tcgen05.alloc
    static constexpr size_t SMEM = SMEM_USED > 120 * 1024 ? SMEM_USED : 120 * 1024;
    static_assert(CA % CK == 0 && CB % CK == 0, "channel groups must be chunk aligned");
    static_assert(N3 % 16 == 0 && N3 <= 256, "M = 128 MMAs need N % 16 == 0, N <= 256");
    static_assert(NSLOT >= 2, "needs a TMA ring of at least two boxes");
    static_assert(2 * ACC_COLS <= 512, "does not fit TMEM");
    static_assert(SMEM <= 227 * 1024, "does not fit shared memory");
};

// warps 0-3 and 12.. : epilogue, NSET sets of four warps (TMEM lane quarter = warp % 4; set s takes steps s, s + NSET, ..),
// 4-7 converters, 8 input TMA producer, 9 MMA issuer, 10 epilogue-operand TMA producer, 11 idle.  An epilogue step is a latency chain (TMEM load -> shuffles ->
// MUFU -> stores, ~1 kclk for ~150 instructions).  MEASURED (round 2, bench B = 32, profiles/r2d_k3_epilogue_sets.txt):
// NSET = 2 / 3 / 4 (512 / 640 / 768 threads, no spills) give K3 stage 1/2/3 = 10.97/21.83/19.61, 10.94/22.08/19.84,
// 10.97/22.28/20.05 ms - no gain, so the epilogue's warp count is not what bounds these kernels; 2 sets stay.
constexpr int kTcDefaultSets = 2;
constexpr int kTcConverters = 4;                              // converter warps 4-7; 8 (adding warps 16-19, a 640-thread CTA) measured 1.7 % SLOWER (profiles/r2d)
constexpr int tc_threads(int nset) { return (12 + 4 * (nset - 1) + (kTcConverters > 4 ? 4 : 0)) * 32; }

template <int CA, int CB, int COUT, int EPI, int MT, int PREC, int NSET>
__global__ void __launch_bounds__(tc_threads(NSET), 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmH, ConvArgs a, TileGrid tg) {
    using C = TcCfg<CA, CB, COUT, MT, PREC, EPI>;
    using G = TcGeom<MT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sSlot = smem_raw;                                         // [NSLOT] planar [8][IH][36]
    unsigned char* sA = sSlot + C::NSLOT * C::SLOT_BYTES;                    // [NA] operand stages
    unsigned char* sB = sA + C::NA * C::STAGE_BYTES;                         // [3 ky][NCH][2 quads][N3][4]
    float* sEpi = reinterpret_cast<float*>(sB + C::B_BYTES);                 // [EPI_R][u channels | h channels][TH][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::B_BYTES + C::EPI_BYTES);
    uint64_t* slot_full = bars;
    uint64_t* slot_empty = slot_full + C::NSLOT;
    uint64_t* a_full = slot_empty + C::NSLOT;
    uint64_t* a_empty = a_full + C::NA;
    uint64_t* d_full = a_empty + C::NA;
    uint64_t* d_empty = d_full + 2;
    uint64_t* e_full = d_empty + 2;
    uint64_t* e_empty = e_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_empty + 3);
    __shared__ float sBias[COUT];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < C::NSLOT; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], kTcConverters); }
        for (int i = 0; i < C::NA; ++i) { mbar_init(&a_full[i], kTcConverters); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 4 * NSET); }
        for (int i = 0; i < 3; ++i) { mbar_init(&e_full[i], 1); mbar_init(&e_empty[i], 4 * NSET); }
        fence_mbar_init();
    }
    const int wrow = a.wpk_cout > 0 ? a.wpk_cout : COUT, ocout = a.out_cout > 0 ? a.out_cout : COUT;   // slice of a wider layer
    if (tid < COUT) sBias[tid] = (EPI == EPI_RELU || a.bias == nullptr) ? 0.f : __ldg(a.bias + a.co_off + tid);
    // resident weights: [ci][tap][co] -> B operand [ky][chunk][quad][kx: W_hi co | W_lo co][4 ci], split once per CTA
    for (int i = tid; i < 9 * C::NCH * 2 * COUT * 4; i += tc_threads(NSET)) {
        int r = i;
        const int j = r & 3; r >>= 2;
        const int n = r % COUT; r /= COUT;
        const int kx = r % 3; r /= 3;
        const int kq = r & 1; r >>= 1;
        const int s = r % C::NCH, ky = r / C::NCH;
        const float v = __ldg(a.wpk + ((size_t)(8 * s + 4 * kq + j) * 9 + ky * 3 + kx) * wrow + a.co_off + n);
        float hi, lo;
        split_tf32(v, hi, lo);
        float* dst = reinterpret_cast<float*>(sB + (size_t)(ky * C::NCH + s) * C::B_STEP_BYTES + kq * C::N3 * 16);
        dst[(kx * COUT + n) * 4 + j] = hi;
        dst[((3 + kx) * COUT + n) * 4 + j] = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    pdl_wait();                                                              // the prologue above overlapped the previous kernel
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int my_tiles = (tg.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * C::NCH;                                      // chunk stream of this CTA
    const int tiles_per_item = tg.tiles_x * tg.tiles_y;

    if (warp == 8) {
        // ===== TMA producer
        if (lane == 0) {
            for (int g = 0; g < total; ++g) {
                const int slot = g % C::NSLOT;
                if (g >= C::NSLOT) mbar_wait_bounded(&slot_empty[slot], ((g / C::NSLOT) - 1) & 1);
                const int ti = g / C::NCH, c = g - ti * C::NCH;
                const int tile = tg.at(blockIdx.x + ti * gridDim.x);
                const int b = tg.by_item.div(tile), r = tile - b * tiles_per_item;
                const int ty = tg.by_x.div(r);
                const int ox0 = (r - ty * tg.tiles_x) * G::TW, oy0 = ty * G::TH;
                const bool fromA = c * CK < CA;
                const int plane = fromA ? b * a.planesA + c * CK : b * a.planesB + (c * CK - CA);
                mbar_expect_tx(&slot_full[slot], C::SLOT_BYTES);
                tma_load_4d(sSlot + slot * C::SLOT_BYTES, fromA ? &tmA : &tmB, &slot_full[slot], (ox0 - 1) & ~3, oy0 - 1, fromA ? a.k : 0, plane);
            }
        }
    } else if (warp == 10) {
        // ===== TMA producer of the GRU epilogue operands (state h for the reset gates; update gate u and h for the blend)
        // of the OUTPUT tile.  Its own thread: the ring is EPI_R tiles deep and is released by the epilogue, which runs
        // two to three tiles behind the input stream - issued from the input producer's loop (as it was) the wait for a
        // free ring entry held the INPUT boxes back too (converters waited ~680 clk per chunk for data, profiles/r2j).
        // Same-box A/B with the fourth operand stage below: K3 44.2 -> 43.7 ms per 32 maps (gpurun session r2z7).
        if (C::EPI_R > 0 && lane == 0) {
            constexpr int ER = C::EPI_R > 0 ? C::EPI_R : 1;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int es = ti % ER;
                if (ti >= ER) mbar_wait_bounded(&e_empty[es], ((ti / ER) - 1) & 1);
                const int tile = tg.at(blockIdx.x + ti * gridDim.x);
                const int b = tg.by_item.div(tile), r = tile - b * tiles_per_item;
                const int ty = tg.by_x.div(r);
                const int ox0 = (r - ty * tg.tiles_x) * G::TW, oy0 = ty * G::TH;
                float* dstE = sEpi + (size_t)es * (C::EPI_TILE_BYTES / 4);
                mbar_expect_tx(&e_full[es], C::EPI_TILE_BYTES);
                if (EPI == EPI_GATES) {
                    tma_load_4d(dstE, &tmH, &e_full[es], ox0 & ~3, oy0, 0, b * (COUT / 2));
                } else {
                    tma_load_4d(dstE, &tmU, &e_full[es], ox0 & ~3, oy0, 0, b * COUT);
                    tma_load_4d(dstE + COUT * C::EPI_PLANE, &tmH, &e_full[es], ox0 & ~3, oy0, 0, b * COUT);
                }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer
        if (lane == 0) {
            // A_hi meets all of [W_hi | W_lo]; A_lo only the W_hi rows (rounded up to the MMA's N granularity: with 8 or 24
            // output channels that includes eight W_lo rows, a harmless part of the dropped lo x lo term)
            const uint32_t idesc_hl[2] = {umma_idesc_tf32(C::N3), umma_idesc_tf32(C::NLO)};
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            int g = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int acc = ti & 1;
                TC_TRACE(1, g, 3);
                if (ti >= 2) mbar_wait_bounded(&d_empty[acc], ((ti >> 1) - 1) & 1);
                for (int c = 0; c < C::NCH; ++c, ++g) {
                    const int st = g % C::NA;
                    TC_TRACE(1, g, 0);
                    mbar_wait_bounded(&a_full[st], (g / C::NA) & 1);
                    tc_fence_after();
                    TC_TRACE(1, g, 1);
                    // descriptors differ only in their 14-bit start-address field (16-byte units; shared memory is
                    // < 256 KB, so adding offsets never carries out of the field)
                    const uint64_t ad00 = umma_desc(a_base + (uint32_t)st * C::STAGE_BYTES, C::PLANE_BYTES, 128);
                    const uint64_t bd00 = umma_desc(b_base + (uint32_t)(c * C::B_STEP_BYTES), C::N3 * 16, 128);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint64_t bd = bd00 + (uint64_t)((ky * C::NCH * C::B_STEP_BYTES) >> 4);
#pragma unroll
                        for (int hl = 0; hl < C::NHL; ++hl) {
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                const uint32_t d = tmem + (uint32_t)(acc * C::ACC_COLS + mt * C::N3);
                                const uint64_t ad = ad00 + (uint64_t)((hl * 2 * C::PLANE_BYTES + (mt * 128 + ky * G::PW) * 16) >> 4);
                                umma_tf32(d, ad, bd, idesc_hl[hl], (ky | hl) ? 1u : (c ? 1u : 0u));
                            }
                        }
                    }
                    umma_commit(&a_empty[st]);                                 // operand stage free once these MMAs retire
                    TC_TRACE(1, g, 2);
                }
                umma_commit(&d_full[acc]);                                     // accumulator complete
            }
        }
    } else if ((warp >= 4 && warp < 8) || (kTcConverters > 4 && warp >= 12 + 4 * (NSET - 1))) {
        // ===== converters: planar TMA slot -> hi/lo quad-interleaved operand stage; warp cw converts rows cw, cw+4, ...
        const int cw = warp < 8 ? warp - 4 : warp - (12 + 4 * (NSET - 1)) + 4;
        constexpr int NJ = (G::IH + kTcConverters - 1) / kTcConverters;
#pragma unroll 1
        for (int g = 0; g < total; ++g) {
            const int slot = g % C::NSLOT, st = g % C::NA;
            const int ti = g / C::NCH;
            const int tile = tg.at(blockIdx.x + ti * gridDim.x);
            const int rr = tile - tg.by_item.div(tile) * tiles_per_item;
            const int ox0 = (rr - tg.by_x.div(rr) * tg.tiles_x) * G::TW;
            const int off = (ox0 - 1) - ((ox0 - 1) & ~3);                     // column of position 0 inside the box
            if (cw == 0 && lane == 0) TC_TRACE(0, g, 0);
            mbar_wait_bounded(&slot_full[slot], (g / C::NSLOT) & 1);
            if (cw == 0 && lane == 0) TC_TRACE(0, g, 3);
            if (g >= C::NA) mbar_wait_bounded(&a_empty[st], ((g / C::NA) - 1) & 1);
            if (cw == 0 && lane == 0) TC_TRACE(0, g, 1);
            const float* pl = reinterpret_cast<const float*>(sSlot + slot * C::SLOT_BYTES) + off + lane;
            unsigned char* stage = sA + (size_t)st * C::STAGE_BYTES;
            // all loads first (the compiler cannot move a shared-memory load above an earlier shared-memory store)
            float e[NJ][CK];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int r = cw + kTcConverters * j;
                const bool on = j < NJ - 1 || r < G::IH;                       // warp-uniform
#pragma unroll
                for (int ch = 0; ch < CK; ++ch) e[j][ch] = on ? pl[(ch * G::IH + (on ? r : 0)) * G::BOXW] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int r = cw + kTcConverters * j;
                if (j == NJ - 1 && r >= G::IH) break;                          // warp-uniform
                const int pos = r * G::PW + lane;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float4 h4, l4;
                    split_fast(e[j][4 * q + 0], h4.x, l4.x); split_fast(e[j][4 * q + 1], h4.y, l4.y);
                    split_fast(e[j][4 * q + 2], h4.z, l4.z); split_fast(e[j][4 * q + 3], h4.w, l4.w);
                    float4* dst = reinterpret_cast<float4*>(stage + (size_t)q * C::PLANE_BYTES) + pos;
                    *dst = h4;
                    if (PREC == PREC_FP32X3) *(dst + 2 * G::NPOS) = l4;
                }
            }
            fence_proxy_async();                                               // operand writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) { mbar_arrive(&a_full[st]); mbar_arrive(&slot_empty[slot]); }
            if (cw == 0 && lane == 0) TC_TRACE(0, g, 2);
        }
    } else if (warp < 4 || (warp >= 12 && warp < 12 + 4 * (NSET - 1))) {
        // ===== epilogue (TMEM lane quarter = warp % 4 = image row mt*4 + quarter): runs one tile behind the MMAs.  One warp
        // per scheduler is latency bound (TMEM load -> shuffles -> MUFU -> stores: ~4.4 clk per instruction measured),
        // so two warps per quarter take alternate (M tile, channel group) steps.
        const int quarter = warp & 3, eset = warp < 4 ? 0 : ((warp - 12) >> 2) + 1;
        constexpr int NGW = (C::NG_STEPS + NSET - 1) / NSET;                 // steps per warp and tile (the last may be empty)
        const size_t plane = (size_t)a.hout * a.wout;
        constexpr int CG = 8;                                                // output channels per step
        constexpr int NCG = COUT / CG, NG = MT * NCG;                        // (M tile, channel group) steps per tile
        constexpr int HC = COUT / 2;
#pragma unroll 1
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int acc = ti & 1;
            const int tile = tg.at(blockIdx.x + ti * gridDim.x);
            const int b = tg.by_item.div(tile), rr = tile - b * tiles_per_item;
            const int ty = tg.by_x.div(rr);
            const int ox = (rr - ty * tg.tiles_x) * G::TW + lane, oy0 = ty * G::TH + quarter;
            float st_s[2] = {0.f, 0.f}, st_q[2] = {0.f, 0.f};               // RAW_STATS: this lane's moments of the tile, per channel half
            auto process = [&](int sI) {
                const int gi = eset + NSET * sI;
                const int mt = gi / NCG, c0 = (gi - mt * NCG) * CG;
                // GRU operands of this pixel: box row 4*mt + quarter, column (ox0 - aligned box start) + lane
                const float* opnd = sEpi + (size_t)(ti % (C::EPI_R > 0 ? C::EPI_R : 1)) * (C::EPI_TILE_BYTES / 4)
                                  + (4 * mt + quarter) * 32 + (ox - lane - ((ox - lane) & ~3)) + lane;
                const int oy = oy0 + 4 * mt;
                const bool valid = lane < G::TW && oy < a.hout && ox < a.wout;
                const size_t pix = valid ? (size_t)oy * a.wout + ox : 0;
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * C::ACC_COLS + mt * C::N3);
                float v[CG];
                // all six loads in flight before the one wait (kx x {A W_hi, A W_lo} columns of this channel group)
                uint32_t rh[3][8], rl[3][8];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    tmem_ld8_issue(taddr + kx * COUT + c0, rh[kx]);
                    tmem_ld8_issue(taddr + (3 + kx) * COUT + c0, rl[kx]);
                }
                tmem_ld_wait();
                if (tid == 0 && gi == 0) TC_TRACE(3, ti, 0);
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    // out[p] = D'[p][0] + D'[p+1][1] + D'[p+2][2]: the neighbours' partial sums come by shuffle
                    v[c] = __uint_as_float(rh[0][c]) + __uint_as_float(rl[0][c]);
                    v[c] += __shfl_down_sync(0xffffffffu, __uint_as_float(rh[1][c]) + __uint_as_float(rl[1][c]), 1);
                    v[c] += __shfl_down_sync(0xffffffffu, __uint_as_float(rh[2][c]) + __uint_as_float(rl[2][c]), 2);
                }
                if (tid == 0 && gi == 0) TC_TRACE(3, ti, 1);
                if (EPI != EPI_RAW_STATS && !valid) return;                   // (RAW_STATS: every lane stays for the warp reduction below)
                // All values first, all stores afterwards: the operands below come from (generic-address) shared memory,
                // and the compiler will not move such a load above an earlier global store - with load, MUFU and store
                // interleaved per channel the eight ~180-clk dependency chains ran one after the other (profiles/r02e).
                float res[CG];
                static_assert(EPI != EPI_GATES || HC % CG == 0, "a step's channel group lies in one gate half");
                const bool reset_half = c0 < HC;                             // warp-uniform: the whole step is reset or update gates
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    const int co = c0 + c;
                    if (EPI == EPI_GATES) {
                        const float s = sigmoid_f(v[c] + sBias[co]);
                        res[c] = reset_half ? s * opnd[(reset_half ? co : 0) * C::EPI_PLANE] : s;   // reset gate -> r*h | update gate
                    } else if (EPI == EPI_CAND) {
                        const float u = opnd[co * C::EPI_PLANE], hv = opnd[(COUT + co) * C::EPI_PLANE];
                        res[c] = u * hv + (1.f - u) * tanh_f(v[c] + sBias[co]);
                    } else if (EPI == EPI_RELU) {
                        res[c] = fmaxf(v[c], 0.f);
                    } else if (EPI == EPI_RAW_STATS) {                         // raw conv + bias; GroupNorm moments below
                        res[c] = v[c] + sBias[co];
                    } else {                                                   // EPI_BIAS
                        const float y = v[c] + sBias[co];
                        res[c] = a.relu ? fmaxf(y, 0.f) : y;
                    }
                }
                {
                    float* dst;
                    if (EPI == EPI_GATES) dst = c0 < HC ? a.out0 + ((size_t)b * HC + c0) * plane + pix : a.out1 + ((size_t)b * HC + (c0 - HC)) * plane + pix;
                    else dst = a.out0 + ((size_t)b * ocout + a.co_off + c0) * plane + pix;
                    if (EPI == EPI_RAW_STATS) {
                        float s = 0.f, q = 0.f;
                        if (valid) {
#pragma unroll
                            for (int c = 0; c < CG; ++c) { dst[(size_t)c * plane] = res[c]; s += res[c]; q = fmaf(res[c], res[c], q); }
                        }
                        const bool upper = a.stats_split && (a.co_off + c0) >= ocout / 2;       // warp-uniform
                        st_s[0] += upper ? 0.f : s; st_q[0] += upper ? 0.f : q;
                        st_s[1] += upper ? s : 0.f; st_q[1] += upper ? q : 0.f;
                    } else {
#pragma unroll
                        for (int c = 0; c < CG; ++c) dst[(size_t)c * plane] = res[c];
                    }
                }
                if (tid == 0 && gi == 0) TC_TRACE(3, ti, 2);
            };
            if (tid == 0) TC_TRACE(2, ti, 0);
            mbar_wait_bounded(&d_full[acc], (ti >> 1) & 1);
            tc_fence_after();
            if (C::EPI_R > 0) mbar_wait_bounded(&e_full[ti % (C::EPI_R > 0 ? C::EPI_R : 1)], (ti / (C::EPI_R > 0 ? C::EPI_R : 1)) & 1);
            if (tid == 0) TC_TRACE(2, ti, 1);
#pragma unroll
            for (int sI = 0; sI < NGW; ++sI)
                if (eset + NSET * sI < NG) process(sI);                        // steps eset, eset + NSET, ... of this warp set
            if (EPI == EPI_RAW_STATS) {
                // per-item moments of the raw output for the one-group GroupNorm that follows (msrednet.cu): warp partials
                // in fp64 -> atomics on stats[b][half][sum, sum of squares], like the FFMA kernels' EPI_RAW_STATS
#pragma unroll
                for (int gpart = 0; gpart < 2; ++gpart) {
                    double ds = (double)st_s[gpart], dq = (double)st_q[gpart];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dq += __shfl_xor_sync(0xffffffffu, dq, o); }
                    if (lane == 0 && (gpart == 0 || a.stats_split)) {
                        atomicAdd(a.stats + ((size_t)b * 2 + gpart) * 2, ds);
                        atomicAdd(a.stats + ((size_t)b * 2 + gpart) * 2 + 1, dq);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&d_empty[acc]);
                if (C::EPI_R > 0) mbar_arrive(&e_empty[ti % (C::EPI_R > 0 ? C::EPI_R : 1)]);
            }
            if (tid == 0) TC_TRACE(2, ti, 2);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
}

template <int CA, int CB, int COUT, int EPI>
struct TcLayer {
    // M tiles per tile: as many as the double-buffered accumulator (2 x MT x 6*Cout columns) fits into TMEM
    static constexpr int MT = COUT <= 8 ? (EPI == EPI_CAND ? 2 : 4) : (COUT <= 16 ? 2 : 1);   // COUT 24, 32: 2 x 144 / 192 columns
    using G = TcGeom<MT>;
    static bool plan(ConvPlan& p, const ConvArgs& a, int B, int depthA) {
        p.args = a;
        p.cfg = MT;
        if (!make_tmap_4d(&p.tA, a.inA, a.win, a.hin, depthA, (long long)B * a.planesA, G::BOXW, G::IH, CK)) return false;
        if (CB > 0) { if (!make_tmap_4d(&p.tB, a.inB, a.win, a.hin, 1, (long long)B * a.planesB, G::BOXW, G::IH, CK)) return false; }
        else p.tB = p.tA;
        p.tU = p.tA; p.tH = p.tA;
        if (EPI == EPI_GATES) {
            if (!make_tmap_4d(&p.tH, a.hstate, a.wout, a.hout, 1, (long long)B * (COUT / 2), 32, G::TH, COUT / 2)) return false;
        } else if (EPI == EPI_CAND) {
            if (!make_tmap_4d(&p.tU, a.ugate, a.wout, a.hout, 1, (long long)B * COUT, 32, G::TH, COUT)) return false;
            if (!make_tmap_4d(&p.tH, a.hstate, a.wout, a.hout, 1, (long long)B * COUT, 32, G::TH, COUT)) return false;
        }
        return true;
    }
    template <int PREC, int NSET>
    static cudaError_t launch_set(ConvPlan& p, int B, cudaStream_t st) {
        using C = TcCfg<CA, CB, COUT, MT, PREC, EPI>;
        auto kern = conv3x3_tc_kernel<CA, CB, COUT, EPI, MT, PREC, NSET>;
        static bool ready[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        dev = dev < 64 ? dev : 63;
        if (!ready[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            if (e != cudaSuccess) return e;
            ready[dev] = true;
        }
        p.tg.tiles_x = (p.args.wout + G::TW - 1) / G::TW;
        p.tg.tiles_y = (p.args.hout + G::TH - 1) / G::TH;
        p.tg.ntiles = p.tg.tiles_x * p.tg.tiles_y * B;
        if (p.tg.ntiles >= (1 << 26)) return cudaErrorInvalidValue;
        p.tg.by_x = FastDiv(p.tg.tiles_x); p.tg.by_item = FastDiv(p.tg.tiles_x * p.tg.tiles_y);
        int ctas = sm_count();
        if (ctas > p.tg.ntiles) ctas = p.tg.ntiles;
        return launch_pdl(kern, dim3(ctas, 1, 1), dim3(tc_threads(NSET)), C::SMEM, st, p.tA, p.tB, p.tU, p.tH, p.args, p.tg);
    }
    template <int PREC>
    static cudaError_t launch_prec(ConvPlan& p, int B, cudaStream_t st) {
        return launch_set<PREC, kTcDefaultSets>(p, B, st);
    }
    static cudaError_t launch(ConvPlan& p, int B, int prec, cudaStream_t st) {
        return prec == PREC_TF32 ? launch_prec<PREC_TF32>(p, B, st) : launch_prec<PREC_FP32X3>(p, B, st);
    }
};

}  // namespace adamvs
