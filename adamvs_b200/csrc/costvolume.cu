// K1 adamvs_pair_score_f32 and K2 adamvs_fused_volume_f32: homography warp + bilinear gather fused
// with the matching product and the per-view weighted aggregation.  No per-view warped volume, no
// sampling grid and no hypothesis tensor is ever written (the reference materialises all three:
// models/module.py:543-566, models/adamvs.py:258,269-301).
//
// Thread mapping: one thread per (reference pixel, depth plane), x fastest, channels innermost with
// the tap sets of all source views held in registers.
#include "common.cuh"
#include "tma.cuh"
#include <limits.h>
#include <stdlib.h>
#include <type_traits>

namespace adamvs {

constexpr int kMaxViews = 8;      // source views per reference view handled by one launch
constexpr int kTileX = 32;        // one warp = 32 x-consecutive reference pixels
constexpr int kTileY = 2;         // pixel rows per block
constexpr int kTileK = 4;         // depth planes per block (adjacent planes share most of their source footprint in L1)
constexpr int kCvThreads = kTileX * kTileY * kTileK;

// Thread mapping (both kernels): lane = x, warp = (pixel row, plane).  A warp's tap load for one
// (view, channel) touches 1-2 cache lines of one source row; its output store is one coalesced
// 128-byte row of the [.., k, y, :] plane.  The block's 8 warps cover 2 pixel rows x 4 adjacent
// planes, whose source footprints overlap, so the L2->L1 traffic per output stays ~1.5x instead of
// the 4.6x (4 views x halo) of a one-plane block.  The 16 tap reads per output make the LSU/L1
// gather rate (128 B/clk/SM), not HBM, the binding resource: DESIGN.md §3.

// Tap weights pre-multiplied by a per-view factor; offsets are 32-bit element offsets into one
// [h,w] plane so that loads are `uniform channel base + per-thread offset`.
struct WTaps {
    unsigned o00, o01, o10, o11;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ WTaps scaled_taps(const Ray& r, float d, int h, int w, float scale) {
    const Taps t = make_taps(r, d, h, w);
    WTaps o;
    o.o00 = (unsigned)t.o00; o.o01 = (unsigned)t.o01; o.o10 = (unsigned)t.o10; o.o11 = (unsigned)t.o11;
    o.w00 = t.w00 * scale; o.w01 = t.w01 * scale; o.w10 = t.w10 * scale; o.w11 = t.w11 * scale;
    return o;
}

__device__ __forceinline__ float gather4(const float* __restrict__ p, const WTaps& t, float acc) {
    acc = fmaf(t.w00, __ldg(p + t.o00), acc);
    acc = fmaf(t.w01, __ldg(p + t.o01), acc);
    acc = fmaf(t.w10, __ldg(p + t.o10), acc);
    acc = fmaf(t.w11, __ldg(p + t.o11), acc);
    return acc;
}

// ------------------------------------------------------------------------------------------------
// K1: score[b,v,k,y,x] = mean_c ref[c] * warp_v[c,k]         (grid.z = b * Vs + v)
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kCvThreads)
pair_score_kernel(const float* __restrict__ feat, const float* __restrict__ relproj, HypSpec hs,
                  float* __restrict__ score, int V, int D, int h, int w) {
    const int hw = h * w;
    const int tiles_x = (w + kTileX - 1) / kTileX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x = tx * kTileX + lane;
    const int y = ty * kTileY + (wid % kTileY);
    const int k = blockIdx.y * kTileK + wid / kTileY;
    const int Vs = V - 1;
    const int b = blockIdx.z / Vs, v = blockIdx.z - b * Vs;
    if (x >= w || y >= h || k >= D) return;
    const int pix = y * w + x;
    const float* ref = feat + ((size_t)b * V) * C * hw + pix;
    const float* src = feat + ((size_t)b * V + v + 1) * C * hw;
    const Ray ray = make_ray(relproj + ((size_t)b * Vs + v) * 12, (float)x, (float)y);
    const HypLine line = hyp_line(hs, b, pix, hw, D);
    const WTaps t = scaled_taps(ray, hyp_at(line, k), h, w, 1.f);
    float acc = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        const float s = gather4(src + (size_t)c * hw, t, 0.f);
        acc = fmaf(__ldg(ref + (size_t)c * hw), s, acc);
    }
    score[(((size_t)b * Vs + v) * D + k) * hw + pix] = acc / (float)C;
}

// ------------------------------------------------------------------------------------------------
// K2: volume[b,c,k,y,x] = view-weighted mean of ref[c]*warp_v[c,k] (two epsilon conventions)
//   numerator eps:   (1e-5 + sum_v w_v ref warp_v) / sum_v w_v
//   denominator eps:  sum_v w_v ref warp_v / (1e-5 + sum_v w_v)
// evaluated as (start + ref[c] * sum_v sum_tap (tapw*w_v) * tex) * (1/denom): 16 FFMA + 2 per output.
// ------------------------------------------------------------------------------------------------
template <int C, int VS>
__global__ void __launch_bounds__(kCvThreads)
fused_volume_kernel(const float* __restrict__ feat, const float* __restrict__ relproj, HypSpec hs,
                    const float* __restrict__ weights, int eps_mode,
                    float* __restrict__ volume, int D, int h, int w) {
    constexpr int V = VS + 1;
    const int hw = h * w;
    const int tiles_x = (w + kTileX - 1) / kTileX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x = tx * kTileX + lane;
    const int y = ty * kTileY + (wid % kTileY);
    const int k = blockIdx.y * kTileK + wid / kTileY;
    const int b = blockIdx.z;
    if (x >= w || y >= h || k >= D) return;
    const int pix = y * w + x;
    const float* ref = feat + ((size_t)b * V) * C * hw + pix;
    const float* src0 = feat + ((size_t)b * V + 1) * C * hw;

    const HypLine line = hyp_line(hs, b, pix, hw, D);
    const float d = hyp_at(line, k);
    WTaps t[VS];
    float wsum = 0.f;            // reference: weight_sum = 0 + w_0 + w_1 ... (left to right)
#pragma unroll
    for (int v = 0; v < VS; ++v) {
        const float wv = __ldg(weights + ((size_t)b * VS + v) * hw + pix);
        wsum += wv;
        const Ray ray = make_ray(relproj + ((size_t)b * VS + v) * 12, (float)x, (float)y);
        t[v] = scaled_taps(ray, d, h, w, wv);
    }
    const bool eps_num = (eps_mode == ADAMVS_EPS_NUMERATOR);
    const float inv = 1.f / (eps_num ? wsum : (1e-5f + wsum));
    const float start = eps_num ? 1e-5f : 0.f;
    float* out = volume + (((size_t)b * C) * D + k) * hw + pix;
#pragma unroll 2
    for (int c = 0; c < C; ++c) {
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < VS; ++v) s = gather4(src0 + ((size_t)v * C + c) * hw, t[v], s);
        const float rc = __ldg(ref + (size_t)c * hw);
        __stcs(out + (size_t)c * D * hw, fmaf(rc, s, start) * inv);     // streaming store: written once, read by K3 later
    }
}

template <int C>
static int launch_pair_score(const float* feat, const float* relproj, const HypSpec& hs, float* score,
                             int B, int V, int D, int h, int w, cudaStream_t st) {
    const int tiles = ((w + kTileX - 1) / kTileX) * ((h + kTileY - 1) / kTileY);
    dim3 grid(tiles, (D + kTileK - 1) / kTileK, B * (V - 1));
    pair_score_kernel<C><<<grid, kCvThreads, 0, st>>>(feat, relproj, hs, score, V, D, h, w);
    ADAMVS_LAUNCH_RESULT();
}

template <int C, int VS>
static int launch_fused_volume(const float* feat, const float* relproj, const HypSpec& hs, const float* weights,
                               int eps_mode, float* volume, int B, int D, int h, int w, cudaStream_t st) {
    const int tiles = ((w + kTileX - 1) / kTileX) * ((h + kTileY - 1) / kTileY);
    dim3 grid(tiles, (D + kTileK - 1) / kTileK, B);
    fused_volume_kernel<C, VS><<<grid, kCvThreads, 0, st>>>(feat, relproj, hs, weights, eps_mode, volume, D, h, w);
    ADAMVS_LAUNCH_RESULT();
}

// ================================================================================================
// TMA-staged variant (the fast path; needs w % 4 == 0).
//
// A block owns a (32/G) x 8 pixel tile and 4*G adjacent depth planes of one batch item; a thread owns one pixel and
// kKC = 4 of those planes.  Every thread first computes the bilinear cells and weights of its pixel for its 4 planes x
// VS views; a block-wide min/max gives, per source view, the bounding box of all cells.  When every box fits
// BW x BH texels (always for the near-fronto-parallel geometry of aerial blocks; otherwise the block takes the
// global-gather path below) the source footprint is brought into shared memory by the TMA unit,
// 4 channels x VS views per stage, double buffered: out-of-image texels arrive as zeros, which is
// exactly grid_sample's per-corner zero padding, so taps need no clamping or masking.  The gather
// then is `LDS [cell + immediate]` (no address arithmetic per tap): 16 LDS + 16 FFMA per output.
// The binding resource is the shared-memory gather rate (16 taps x 4 B per 4-byte output at
// 128 B/clk/SM = one 32-lane wavefront per clock), not HBM: DESIGN.md §3.
//
// Lane mapping (round 2).  A warp is (32/G) x-consecutive pixels x G ADJACENT depth planes (lane = plane lane * 32/G +
// pixel) and the box pitch is 32 words; a thread's planes are k0 + G*kk + its plane lane, so at every tap load the G
// plane lanes of one pixel work on adjacent planes.  Round 1 used 32 pixels of one plane per warp with pitch 64: 44 % of
// all shared-memory wavefronts were bank conflicts (profiles/r03c) - 33-word spans when the source magnification exceeds
// 1 and, in bench, the SCATTER of neighbouring pixels' samples (an untrained stage hands down a white-noise depth map:
// adjacent pixels sample +-6..12 px apart).  Adjacent planes of ONE pixel, however, sample ~0.15 px apart whatever the
// depth map looks like: the same texel (a broadcast, free) or its neighbour.  So a warp touches 32/G scattered word groups
// instead of 32 scattered words.  Measured in bench (profiles/r2b_r2f_k2_stage2_in_bench_b32.txt): G = 2 is the fastest
// (stage 1/2/3 2.60/4.95/3.67 ms against round 1's 2.90/5.20/3.75; still 43 % conflicts at stage 2, where the handed-down
// depth is noise); G = 4 (8 pixels x 4 planes) 2.87/5.28/- ms; lanes along the depth axis only (32 planes of one pixel per
// warp, results transposed through shared memory) removed the conflicts (2.6 %) but cost 53 % more instructions and ran
// latency-bound: 4.58/7.47/5.44 ms.  G = 2 ships.
// The same box serves 4*G planes and is 32 texels wide: L2 -> shared-memory traffic per output is half of round 1's.
// ================================================================================================
// Source box per (view, channel).  48 columns, not 32: the box is also the shared-memory pitch (TMA writes dense rows), and
// the two planes (half-warps) of a warp usually sample adjacent source ROWS at almost the same columns - with a pitch of
// 32 words those are the same banks (a 2-way conflict on most loads), with 48 they are 16 banks apart.  Same-box A/B at
// the bench batch: stage 2 / 3 10.03 / 7.45 -> 9.11 / 6.48 ms, shared wavefronts 2186 M -> 1793 M (conflicts 932 M -> 539 M),
// profiles/r2zc_* vs r2zf_*.  74 KB of stage buffers per CTA, still two CTAs per SM.
constexpr int kBW = 48, kBH = 12;
constexpr int kBox = kBW * kBH;
constexpr int kCK = 4;                      // channels per pipeline stage
// Rough depth (an untrained or noisy previous stage, oblique geometry) spreads a tile's footprint over more source
// rows.  The stage buffer is re-cut at run time, per block, into fewer channels of taller boxes: configuration j holds
// kCK >> j channels of kBH << j rows per view (4x12, 2x24, 1x48) - same bytes, same occupancy, same gather code with
// another channel stride - and only blocks whose footprint exceeds 32 x 48 texels take the global-gather path.
constexpr int kNCfg = 3;
constexpr int kKC = 4;                      // depth planes per thread
constexpr int kPY = 8;                      // pixel rows per block (one warp each); the tile is 32/G pixels wide
constexpr int kWvThreads = 32 * kPY;

enum { MODE_FUSED = 0, MODE_SCORE = 1, MODE_VARIANCE = 2 };

// Predicated streaming store: keeps the gather loop free of branches (the value is always computed).
__device__ __forceinline__ void st_cs_pred(float* p, float v, bool pred) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.cs.f32 [%0], %1; }" ::"l"(p), "f"(v), "r"((int)pred) : "memory");
}

struct WarpVolArgs {
    const float* feat;        // [B,V,C,h,w]
    const float* relproj;     // [B,V-1,12]
    HypSpec hs;
    const float* weights;     // MODE_FUSED: [B,V-1,h,w]
    int eps_mode;
    float* out;               // FUSED/VARIANCE: [B,C,D,h,w]; SCORE: [B,V-1,D,h,w]
    int D, h, w;
};

// Unclamped bilinear cell (x0,y0 may be -1) and corner weights; false when the sample lies entirely
// outside the source image or is not finite (contributes zero).
__device__ __forceinline__ bool cell_taps(const Ray& r, float d, int h, int w, float scale, int& x0, int& y0, float (&wt)[4]) {
    const float X = __fadd_rn(__fmul_rn(r.qx, d), r.tx);
    const float Y = __fadd_rn(__fmul_rn(r.qy, d), r.ty);
    const float Z = __fadd_rn(__fmul_rn(r.qz, d), r.tz);
    // Correctly rounded divisions like the reference's xy / z.  Its normalise -> grid_sample un-normalise round trip
    // is an identity whose fp32 roundings move a sample by a few ulp of u (~1e-4 px at x ~ 700); so do the fp32
    // torch.inverse behind the relative projection and the GEMM order of rot @ xyz (measured: DESIGN.md §5), so no
    // independent implementation can track the reference's sample positions closer than that and none of it is
    // imitated here.
    const float u = __fdiv_rn(X, Z);
    const float v = __fdiv_rn(Y, Z);
    const bool ok = (u > -1.f) && (u < (float)w) && (v > -1.f) && (v < (float)h);      // false for NaN / inf; finite Z < 0 as the reference
    const float fu = floorf(ok ? u : 0.f), fv = floorf(ok ? v : 0.f);
    x0 = (int)fu; y0 = (int)fv;
    const float ax = (ok ? u : 0.f) - fu, ay = (ok ? v : 0.f) - fv;
    const float bx = 1.f - ax, by = 1.f - ay;
    const float sc = ok ? scale : 0.f;
    wt[0] = bx * by * sc; wt[1] = ax * by * sc; wt[2] = bx * ay * sc; wt[3] = ax * ay * sc;
    return ok;
}

// Global-gather path for one pixel and the block's planes (any geometry).  Same arithmetic as the
// fast path; used by blocks whose source footprint does not fit the shared-memory box.
template <int C, int VS, int MODE>
__device__ __noinline__ void warp_volume_slow(const WarpVolArgs& a, int b, int x, int y, int k0, int kstep) {
    constexpr int V = VS + 1;
    const int hw = a.h * a.w, pix = y * a.w + x;
    const float* ref = a.feat + ((size_t)b * V) * C * hw + pix;
    const float* src0 = a.feat + ((size_t)b * V + 1) * C * hw;
    const HypLine line = hyp_line(a.hs, b, pix, hw, a.D);
    float wv[VS], wsum = 0.f;
#pragma unroll
    for (int v = 0; v < VS; ++v) {
        wv[v] = MODE == MODE_FUSED ? __ldg(a.weights + ((size_t)b * VS + v) * hw + pix) : 1.f;
        wsum += wv[v];
    }
    const bool eps_num = (a.eps_mode == ADAMVS_EPS_NUMERATOR);
    const float inv = 1.f / (eps_num ? wsum : (1e-5f + wsum));
    const float start = eps_num ? 1e-5f : 0.f;
    for (int kk = 0; kk < kKC; ++kk) {
        const int k = k0 + kk * kstep;
        if (k >= a.D) break;
        const float d = hyp_at(line, k);
        WTaps t[VS];
#pragma unroll
        for (int v = 0; v < VS; ++v)
            t[v] = scaled_taps(make_ray(a.relproj + ((size_t)b * VS + v) * 12, (float)x, (float)y), d, a.h, a.w, wv[v]);
        float acc[VS];
#pragma unroll
        for (int v = 0; v < VS; ++v) acc[v] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float rc = __ldg(ref + (size_t)c * hw);
            if (MODE == MODE_FUSED) {
                float s = 0.f;
#pragma unroll
                for (int v = 0; v < VS; ++v) s = gather4(src0 + ((size_t)v * C + c) * hw, t[v], s);
                __stcs(a.out + (((size_t)b * C + c) * a.D + k) * hw + pix, fmaf(rc, s, start) * inv);
            } else if (MODE == MODE_SCORE) {
#pragma unroll
                for (int v = 0; v < VS; ++v) acc[v] = fmaf(rc, gather4(src0 + ((size_t)v * C + c) * hw, t[v], 0.f), acc[v]);
            } else {
                float sum = rc, sq = rc * rc;
#pragma unroll
                for (int v = 0; v < VS; ++v) {
                    const float sv = gather4(src0 + ((size_t)v * C + c) * hw, t[v], 0.f);
                    sum += sv; sq = fmaf(sv, sv, sq);
                }
                const float m = sum / (float)V;
                __stcs(a.out + (((size_t)b * C + c) * a.D + k) * hw + pix, sq / (float)V - m * m);
            }
        }
        if (MODE == MODE_SCORE) {
#pragma unroll
            for (int v = 0; v < VS; ++v) a.out[(((size_t)b * VS + v) * a.D + k) * hw + pix] = acc[v] / (float)C;
        }
    }
}

template <int C, int VS, int MODE, int G>
__global__ void __launch_bounds__(kWvThreads, 2)
warp_volume_tma_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                       const __grid_constant__ CUtensorMap tm2, WarpVolArgs a, int nk, int force_cfg) {
    constexpr int V = VS + 1;
    constexpr int STAGE = VS * kCK * kBox;             // floats per pipeline stage
    static_assert(C % (2 * kCK) == 0, "the stage loop is unrolled by the two pipeline buffers");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sbuf = reinterpret_cast<float*>(smem_raw);                  // [2][VS][kCK][kBH][kBW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbuf + 2 * STAGE);    // [2]
    int* sbox = reinterpret_cast<int*>(bars + 2);                      // [VS][4] minx, miny, maxx, maxy

    constexpr int PXW = 32 / G;                                        // tile width
    const int tid = threadIdx.x, lane = tid & 31, row = tid >> 5;
    const int px = lane % PXW, pl = lane / PXW;                        // pixel and plane lane (see "Lane mapping")
    const int kchunk = blockIdx.x % nk, tile = blockIdx.x / nk;        // plane chunk fastest: blocks that share a
    const int tiles_x = (a.w + PXW - 1) / PXW;                         // source footprint run together (L2 hits)
    const int x = (tile % tiles_x) * PXW + px, y = (tile / tiles_x) * kPY + row;
    const int b = blockIdx.y;
    const int k0 = kchunk * (kKC * G) + pl;                            // this thread's planes: k0 + G*kk
    const int hw = a.h * a.w;
    const bool inside = (x < a.w) && (y < a.h);
    const int pix = inside ? y * a.w + x : 0;

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (tid < VS * 4) sbox[tid] = (tid & 2) ? INT_MIN : INT_MAX;
    __syncthreads();

    // ---- phase 1: cells + weights of this pixel for kKC planes x VS views; bounding boxes
    float wt[kKC][VS][4];
    int cell[kKC][VS];                   // (y0+1) << 16 | (x0+1); later the box-relative float offset
    float inv = 0.f, start = 0.f;
    {
        const HypLine line = hyp_line(a.hs, b, pix, hw, a.D);
        float wsum = 0.f;
#pragma unroll
        for (int v = 0; v < VS; ++v) {
            const float wv = (MODE == MODE_FUSED && inside) ? __ldg(a.weights + ((size_t)b * VS + v) * hw + pix) : 1.f;
            wsum += wv;
            const Ray ray = make_ray(a.relproj + ((size_t)b * VS + v) * 12, (float)x, (float)y);
            int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
#pragma unroll
            for (int kk = 0; kk < kKC; ++kk) {
                int x0, y0;
                const bool ok = cell_taps(ray, hyp_at(line, k0 + G * kk), a.h, a.w, wv, x0, y0, wt[kk][v]) && inside && (k0 + G * kk < a.D);
                if (ok) { mnx = min(mnx, x0); mxx = max(mxx, x0); mny = min(mny, y0); mxy = max(mxy, y0); }
                else { x0 = INT_MAX; wt[kk][v][0] = wt[kk][v][1] = wt[kk][v][2] = wt[kk][v][3] = 0.f; }
                cell[kk][v] = ok ? (((y0 + 1) << 16) | (x0 + 1)) : -1;
            }
            mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
            mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
            if (lane == 0) {
                atomicMin(&sbox[v * 4 + 0], mnx); atomicMin(&sbox[v * 4 + 1], mny);
                atomicMax(&sbox[v * 4 + 2], mxx); atomicMax(&sbox[v * 4 + 3], mxy);
            }
        }
        const bool eps_num = (a.eps_mode == ADAMVS_EPS_NUMERATOR);
        inv = 1.f / (eps_num ? wsum : (1e-5f + wsum));
        start = eps_num ? 1e-5f : 0.f;
    }
    __syncthreads();
    int bx[VS], by[VS];
    bool fits = true;
    int rows = 0;                                      // source rows the tallest view footprint needs
#pragma unroll
    for (int v = 0; v < VS; ++v) {
        const int mnx = sbox[v * 4 + 0], mny = sbox[v * 4 + 1], mxx = sbox[v * 4 + 2], mxy = sbox[v * 4 + 3];
        const bool any = mnx != INT_MAX;
        bx[v] = any ? (mnx & ~3) : 0;                  // TMA: innermost start coordinate 16-byte aligned
        by[v] = any ? mny : 0;
        fits = fits && (!any || (mxx + 1 - bx[v] < kBW));
        rows = max(rows, any ? mxy + 2 - by[v] : 0);
    }
    int cfg = rows <= kBH ? 0 : (rows <= 2 * kBH ? 1 : 2);
    if (force_cfg > 0 && force_cfg < kNCfg && cfg < force_cfg) cfg = force_cfg;     // test hook: exercise the tall-box cuts
    if (!fits || rows > (kBH << (kNCfg - 1)) || force_cfg >= kNCfg) {               // block-uniform
        if (inside) warp_volume_slow<C, VS, MODE>(a, b, x, y, k0, G);
        return;
    }
#pragma unroll
    for (int kk = 0; kk < kKC; ++kk)
#pragma unroll
        for (int v = 0; v < VS; ++v) {
            const int c = cell[kk][v];
            cell[kk][v] = c < 0 ? 0 : (((c >> 16) - 1 - by[v]) * kBW + ((c & 0xffff) - 1 - bx[v]));
        }

    // ---- phase 2: double-buffered channel stages
    const float* ref = a.feat + ((size_t)b * V) * C * hw + pix;
    // per-thread gather bases: every tap address below is base + compile-time constant (stage, view, channel, corner)
    const float* tap[kKC][VS];
#pragma unroll
    for (int kk = 0; kk < kKC; ++kk)
#pragma unroll
        for (int v = 0; v < VS; ++v) tap[kk][v] = sbuf + cell[kk][v];
    float acc[MODE == MODE_SCORE ? kKC : 1][MODE == MODE_SCORE ? VS : 1];
    if (MODE == MODE_SCORE) {
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk)
#pragma unroll
            for (int v = 0; v < VS; ++v) acc[MODE == MODE_SCORE ? kk : 0][MODE == MODE_SCORE ? v : 0] = 0.f;
    }
    const size_t plane_stride = (size_t)a.D * hw;      // output channel stride
    float* outp = a.out + ((size_t)b * C * a.D + k0) * hw + pix;
    bool kvalid[kKC];
#pragma unroll
    for (int kk = 0; kk < kKC; ++kk) kvalid[kk] = inside && (k0 + G * kk < a.D);

    auto run = [&](auto cfg_tag) {
        constexpr int CFG = decltype(cfg_tag)::value;
        constexpr int CKc = kCK >> CFG;                // channels per stage
        constexpr int CBOX = kBox << CFG;              // floats per (view, channel) box
        constexpr int NCH = C / CKc;
        const CUtensorMap* tm = CFG == 0 ? &tm0 : (CFG == 1 ? &tm1 : &tm2);
        auto issue = [&](int ch) {
            const int s = ch & 1;
            fence_proxy_async();
            mbar_expect_tx(&bars[s], STAGE * 4);
#pragma unroll
            for (int v = 0; v < VS; ++v)
                tma_load_4d(sbuf + s * STAGE + v * kCK * kBox, tm, &bars[s], bx[v], by[v], 0, (b * V + v + 1) * C + ch * CKc);
        };
        if (tid == 0) issue(0);
        auto stage_body = [&](auto stage_tag, int ch) {
            constexpr int S = decltype(stage_tag)::value;
            if (tid == 0 && ch + 1 < NCH) issue(ch + 1);
            float rc[CKc];
#pragma unroll
            for (int cc = 0; cc < CKc; ++cc) rc[cc] = inside ? __ldg(ref + (size_t)(ch * CKc + cc) * hw) : 0.f;
            mbar_wait(&bars[S], (ch >> 1) & 1);
#pragma unroll
            for (int cc = 0; cc < CKc; ++cc) {
#pragma unroll
                for (int kk = 0; kk < kKC; ++kk) {
                    // one independent 4-tap chain per view (ILP), combined afterwards
                    float sv[VS];
#pragma unroll
                    for (int v = 0; v < VS; ++v) {
                        const float* p = tap[kk][v] + (S * STAGE + v * kCK * kBox + cc * CBOX);
                        float t = wt[kk][v][0] * p[0];
                        t = fmaf(wt[kk][v][1], p[1], t);
                        t = fmaf(wt[kk][v][2], p[kBW], t);
                        sv[v] = fmaf(wt[kk][v][3], p[kBW + 1], t);
                    }
                    if (MODE == MODE_SCORE) {
#pragma unroll
                        for (int v = 0; v < VS; ++v)
                            acc[MODE == MODE_SCORE ? kk : 0][MODE == MODE_SCORE ? v : 0] =
                                fmaf(rc[cc], sv[v], acc[MODE == MODE_SCORE ? kk : 0][MODE == MODE_SCORE ? v : 0]);
                    } else {
                        float r;
                        if (MODE == MODE_FUSED) {
                            float sacc = sv[0];
#pragma unroll
                            for (int v = 1; v < VS; ++v) sacc += sv[v];
                            r = fmaf(rc[cc], sacc, start) * inv;
                        } else {
                            float sum = rc[cc], sq = rc[cc] * rc[cc];
#pragma unroll
                            for (int v = 0; v < VS; ++v) { sum += sv[v]; sq = fmaf(sv[v], sv[v], sq); }
                            const float m = sum / (float)V;
                            r = sq / (float)V - m * m;
                        }
                        st_cs_pred(outp + (size_t)cc * plane_stride + (size_t)(G * kk) * hw, r, kvalid[kk]);
                    }
                }
            }
            outp += (size_t)CKc * plane_stride;
            __syncthreads();                           // stage S fully consumed before chunk ch+2 lands in it
        };
#pragma unroll 1
        for (int ch = 0; ch < NCH; ch += 2) {
            stage_body(std::integral_constant<int, 0>{}, ch);
            stage_body(std::integral_constant<int, 1>{}, ch + 1);
        }
    };
    if (cfg == 0) run(std::integral_constant<int, 0>{});
    else if (cfg == 1) run(std::integral_constant<int, 1>{});
    else run(std::integral_constant<int, 2>{});
    if (MODE == MODE_SCORE) {
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk)
#pragma unroll
            for (int v = 0; v < VS; ++v)
                if (kvalid[kk])
                    a.out[(((size_t)b * VS + v) * a.D + k0 + G * kk) * hw + pix] =
                        acc[MODE == MODE_SCORE ? kk : 0][MODE == MODE_SCORE ? v : 0] / (float)C;
    }
}

// Any width / alignment / view count: the global-gather path alone.
template <int C, int VS, int MODE>
__global__ void __launch_bounds__(kWvThreads)
warp_volume_plain_kernel(WarpVolArgs a, int nk) {
    const int tid = threadIdx.x, lane = tid & 31, row = tid >> 5;            // 32 x 8 pixels, kKC contiguous planes
    const int kchunk = blockIdx.x % nk, tile = blockIdx.x / nk;
    const int tiles_x = (a.w + 31) / 32;
    const int x = (tile % tiles_x) * 32 + lane, y = (tile / tiles_x) * 8 + row;
    if (x < a.w && y < a.h) warp_volume_slow<C, VS, MODE>(a, blockIdx.y, x, y, kchunk * kKC, 1);
}

template <int C, int VS, int MODE>
static int launch_warp_volume_plain(const WarpVolArgs& a, int B, cudaStream_t st) {
    const int nk = (a.D + kKC - 1) / kKC;
    const long long tiles = (long long)((a.w + 31) / 32) * ((a.h + 7) / 8);
    if (tiles * nk > 0x7fffffffLL) return ADAMVS_EINVAL;
    dim3 grid((unsigned)(tiles * nk), B, 1);
    warp_volume_plain_kernel<C, VS, MODE><<<grid, 256, 0, st>>>(a, nk);
    ADAMVS_LAUNCH_RESULT();
}

template <int C, int VS, int MODE, int G>
static int launch_warp_volume_tma_g(const WarpVolArgs& a, int B, const CUtensorMap* tm, int force_cfg, cudaStream_t st) {
    constexpr size_t smem = sizeof(float) * 2 * VS * kCK * kBox + 2 * sizeof(uint64_t) + VS * 4 * sizeof(int);
    auto kern = warp_volume_tma_kernel<C, VS, MODE, G>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int nk = (a.D + kKC * G - 1) / (kKC * G);
    const long long tiles = (long long)((a.w + 32 / G - 1) / (32 / G)) * ((a.h + kPY - 1) / kPY);
    if (tiles * nk > 0x7fffffffLL) return ADAMVS_EINVAL;
    dim3 grid((unsigned)(tiles * nk), B, 1);
    kern<<<grid, kWvThreads, smem, st>>>(tm[0], tm[1], tm[2], a, nk, force_cfg);
    ADAMVS_LAUNCH_RESULT();
}

template <int C, int VS, int MODE>
static int launch_warp_volume_tma(const WarpVolArgs& a, int B, cudaStream_t st) {
    CUtensorMap tm[kNCfg];
    for (int j = 0; j < kNCfg; ++j)
        if (!make_tmap_4d(&tm[j], a.feat, a.w, a.h, 1, (long long)B * (VS + 1) * C, kBW, kBH << j, kCK >> j)) return -100;
    // test hook (read once per process, like ADAMVS_CONV_CFG): ADAMVS_WARP_CFG=1|2 forces the taller box cuts, 3 the
    // global-gather path, so that parity tests cover every path on smooth synthetic depth
    static const int force_cfg = [] { const char* e = getenv("ADAMVS_WARP_CFG"); return (e && *e >= '0' && *e <= '3') ? (*e - '0') : 0; }();
    return launch_warp_volume_tma_g<C, VS, MODE, 2>(a, B, tm, force_cfg, st);
}

// TMA needs 16-byte aligned rows: w % 4 == 0 and a 16-byte aligned base; C a multiple of the stage depth.
static bool tma_eligible(const float* feat, int w, int h) {
    return (w % 4 == 0) && (reinterpret_cast<uintptr_t>(feat) % 16 == 0) && h < 65535 && w < 65535;
}

template <int MODE>
static int dispatch_warp_volume_tma(const WarpVolArgs& a, int B, int V, int C, cudaStream_t st) {
#define ADAMVS_WV(CC, VV) return launch_warp_volume_tma<CC, VV, MODE>(a, B, st)
#define ADAMVS_WV_C(VV) switch (C) { case 8: ADAMVS_WV(8, VV); case 16: ADAMVS_WV(16, VV); case 32: ADAMVS_WV(32, VV); default: return -100; }
    switch (V - 1) {
        case 2: ADAMVS_WV_C(2)
        case 3: ADAMVS_WV_C(3)
        case 4: ADAMVS_WV_C(4)
        case 5: ADAMVS_WV_C(5)
        case 6: ADAMVS_WV_C(6)
        default: return -100;                          // other view counts: global-gather kernels
    }
#undef ADAMVS_WV_C
#undef ADAMVS_WV
}

static int check_hyp(int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range) {
    if (!hyp_src) return ADAMVS_EINVAL;
    if (hyp_mode == ADAMVS_HYP_PLANES) return hyp_ncol >= 2 ? 0 : ADAMVS_EINVAL;
    if (hyp_mode == ADAMVS_HYP_PER_PIXEL) return half_range ? 0 : ADAMVS_EINVAL;
    return ADAMVS_EINVAL;
}

}  // namespace adamvs

using namespace adamvs;

extern "C" int adamvs_pair_score_f32(const float* feat, const float* relproj,
                                     int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                     float* score, int B, int V, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(feat && relproj && score && B > 0 && (long long)B * (V - 1) <= 65535 && V >= 2 && V - 1 <= kMaxViews);
    ADAMVS_CHECK_ARG(D >= 2 && h > 0 && w > 0);
    if (int e = check_hyp(hyp_mode, hyp_src, hyp_ncol, half_range)) return e;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    cudaStream_t st = (cudaStream_t)stream;
    if (tma_eligible(feat, w, h)) {
        const WarpVolArgs a{feat, relproj, hs, nullptr, ADAMVS_EPS_DENOMINATOR, score, D, h, w};
        const int rc = dispatch_warp_volume_tma<MODE_SCORE>(a, B, V, C, st);
        if (rc != -100) return rc;
    }
    switch (C) {
        case 8:  return launch_pair_score<8>(feat, relproj, hs, score, B, V, D, h, w, st);
        case 16: return launch_pair_score<16>(feat, relproj, hs, score, B, V, D, h, w, st);
        case 32: return launch_pair_score<32>(feat, relproj, hs, score, B, V, D, h, w, st);
        default: return ADAMVS_EINVAL;
    }
}

extern "C" int adamvs_fused_volume_f32(const float* feat, const float* relproj,
                                       int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                       const float* weights, int eps_mode,
                                       float* volume, int B, int V, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(feat && relproj && weights && volume && B > 0 && B <= 65535 && D >= 2 && h > 0 && w > 0);
    ADAMVS_CHECK_ARG(eps_mode == ADAMVS_EPS_NUMERATOR || eps_mode == ADAMVS_EPS_DENOMINATOR);
    if (int e = check_hyp(hyp_mode, hyp_src, hyp_ncol, half_range)) return e;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    cudaStream_t st = (cudaStream_t)stream;
    if (tma_eligible(feat, w, h)) {
        const WarpVolArgs a{feat, relproj, hs, weights, eps_mode, volume, D, h, w};
        const int rc = dispatch_warp_volume_tma<MODE_FUSED>(a, B, V, C, st);
        if (rc != -100) return rc;
    }
#define ADAMVS_FV(CC, VV) return launch_fused_volume<CC, VV>(feat, relproj, hs, weights, eps_mode, volume, B, D, h, w, st)
#define ADAMVS_FV_C(VV) switch (C) { case 8: ADAMVS_FV(8, VV); case 16: ADAMVS_FV(16, VV); case 32: ADAMVS_FV(32, VV); default: return ADAMVS_EINVAL; }
    switch (V - 1) {
        case 1: ADAMVS_FV_C(1)
        case 2: ADAMVS_FV_C(2)
        case 3: ADAMVS_FV_C(3)
        case 4: ADAMVS_FV_C(4)
        case 5: ADAMVS_FV_C(5)
        case 6: ADAMVS_FV_C(6)
        default: return ADAMVS_EINVAL;
    }
#undef ADAMVS_FV_C
#undef ADAMVS_FV
}

extern "C" int adamvs_variance_volume_f32(const float* feat, const float* relproj,
                                          int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                          float* volume, int B, int V, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(feat && relproj && volume && B > 0 && B <= 65535 && D >= 2 && h > 0 && w > 0);
    if (int e = check_hyp(hyp_mode, hyp_src, hyp_ncol, half_range)) return e;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    cudaStream_t st = (cudaStream_t)stream;
    const WarpVolArgs a{feat, relproj, hs, nullptr, ADAMVS_EPS_DENOMINATOR, volume, D, h, w};
    if (tma_eligible(feat, w, h)) {
        const int rc = dispatch_warp_volume_tma<MODE_VARIANCE>(a, B, V, C, st);
        if (rc != -100) return rc;
    }
#define ADAMVS_VV(CC, VV) return launch_warp_volume_plain<CC, VV, MODE_VARIANCE>(a, B, st)
#define ADAMVS_VV_C(VV) switch (C) { case 8: ADAMVS_VV(8, VV); case 16: ADAMVS_VV(16, VV); case 32: ADAMVS_VV(32, VV); default: return ADAMVS_EINVAL; }
    switch (V - 1) {
        case 1: ADAMVS_VV_C(1)
        case 2: ADAMVS_VV_C(2)
        case 3: ADAMVS_VV_C(3)
        case 4: ADAMVS_VV_C(4)
        case 5: ADAMVS_VV_C(5)
        case 6: ADAMVS_VV_C(6)
        default: return ADAMVS_EINVAL;
    }
#undef ADAMVS_VV_C
#undef ADAMVS_VV
}
