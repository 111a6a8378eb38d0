// K1 adamvs_pair_score_f32 and K2 adamvs_fused_volume_f32: homography warp + bilinear gather fused
// with the matching product and the per-view weighted aggregation.  No per-view warped volume, no
// sampling grid and no hypothesis tensor is ever written (the reference materialises all three:
// models/module.py:543-566, models/adamvs.py:258,269-301).
//
// Thread mapping: one thread per (reference pixel, depth plane), x fastest, channels innermost with
// the tap sets of all source views held in registers.
#include "common.cuh"

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
