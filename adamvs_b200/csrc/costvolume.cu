// K1 adamvs_pair_score_f32 and K2 adamvs_fused_volume_f32: homography warp + bilinear gather fused
// with the matching product and the per-view weighted aggregation.  No per-view warped volume, no
// sampling grid and no hypothesis tensor is ever written (the reference materialises all three:
// models/module.py:543-566, models/adamvs.py:258,269-301).
//
// Thread mapping: one thread per reference pixel, x fastest, so that the D*C output planes are
// written as fully coalesced 128-byte rows; source taps of neighbouring lanes are neighbouring
// pixels of the same source row (NCHW planes), so a warp's tap load touches 1-2 cache lines.
#include "common.cuh"

namespace adamvs {

constexpr int kMaxViews = 8;      // source views per reference view handled by one launch
constexpr int kPixThreads = 128;

// ------------------------------------------------------------------------------------------------
// K1: score[b,v,k,y,x] = mean_c ref[c] * warp_v[c,k]
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kPixThreads)
pair_score_kernel(const float* __restrict__ feat, const float* __restrict__ relproj, HypSpec hs,
                  float* __restrict__ score, int V, int D, int h, int w) {
    const int hw = h * w;
    const int pix = blockIdx.x * kPixThreads + threadIdx.x;
    const int v = blockIdx.y;           // source view index 0..V-2
    const int b = blockIdx.z;
    if (pix >= hw) return;
    const int y = pix / w, x = pix - y * w;
    const float* ref = feat + ((size_t)b * V) * C * hw + pix;
    const float* src = feat + ((size_t)b * V + v + 1) * C * hw;
    float r[C];
#pragma unroll
    for (int c = 0; c < C; ++c) r[c] = __ldg(ref + (size_t)c * hw);
    const Ray ray = make_ray(relproj + ((size_t)b * (V - 1) + v) * 12, (float)x, (float)y);
    const HypLine line = hyp_line(hs, b, pix, hw, D);
    float* out = score + (((size_t)b * (V - 1) + v) * D) * hw + pix;
    for (int k = 0; k < D; ++k) {
        const Taps t = make_taps(ray, hyp_at(line, k), h, w);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float* p = src + (size_t)c * hw;
            const float s = t.w00 * __ldg(p + t.o00) + t.w01 * __ldg(p + t.o01)
                          + t.w10 * __ldg(p + t.o10) + t.w11 * __ldg(p + t.o11);
            acc += r[c] * s;
        }
        out[(size_t)k * hw] = acc / (float)C;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: volume[b,c,k,y,x] = view-weighted mean of ref[c]*warp_v[c,k] (two epsilon conventions)
// ------------------------------------------------------------------------------------------------
template <int C, int VS>
__global__ void __launch_bounds__(kPixThreads)
fused_volume_kernel(const float* __restrict__ feat, const float* __restrict__ relproj, HypSpec hs,
                    const float* __restrict__ weights, int eps_mode,
                    float* __restrict__ volume, int D, int h, int w) {
    constexpr int V = VS + 1;
    const int hw = h * w;
    const int pix = blockIdx.x * kPixThreads + threadIdx.x;
    const int b = blockIdx.z;
    if (pix >= hw) return;
    const int y = pix / w, x = pix - y * w;
    const float* ref = feat + ((size_t)b * V) * C * hw + pix;
    const float* src0 = feat + ((size_t)b * V + 1) * C * hw;

    Ray ray[VS];
    float wv[VS];
    float wsum = 0.f;            // reference: weight_sum = 0 + w_0 + w_1 ... (left to right)
#pragma unroll
    for (int v = 0; v < VS; ++v) {
        ray[v] = make_ray(relproj + ((size_t)b * VS + v) * 12, (float)x, (float)y);
        wv[v] = __ldg(weights + ((size_t)b * VS + v) * hw + pix);
        wsum += wv[v];
    }
    const bool eps_num = (eps_mode == ADAMVS_EPS_NUMERATOR);
    const float denom = eps_num ? wsum : (1e-5f + wsum);   // predict class: 1e-5 + w_0 + w_1 ...
    const float start = eps_num ? 1e-5f : 0.f;
    const HypLine line = hyp_line(hs, b, pix, hw, D);
    float* out = volume + ((size_t)b * C * D) * hw + pix;

    for (int k = 0; k < D; ++k) {
        const float d = hyp_at(line, k);
        Taps t[VS];
#pragma unroll
        for (int v = 0; v < VS; ++v) t[v] = make_taps(ray[v], d, h, w);
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            const float rc = __ldg(ref + (size_t)c * hw);
            float acc = start;
#pragma unroll
            for (int v = 0; v < VS; ++v) {
                const float* p = src0 + ((size_t)v * C + c) * hw;
                const float s = t[v].w00 * __ldg(p + t[v].o00) + t[v].w01 * __ldg(p + t[v].o01)
                              + t[v].w10 * __ldg(p + t[v].o10) + t[v].w11 * __ldg(p + t[v].o11);
                acc += (rc * s) * wv[v];
            }
            out[((size_t)c * D + k) * hw] = acc / denom;
        }
    }
}

template <int C>
static int launch_pair_score(const float* feat, const float* relproj, const HypSpec& hs, float* score,
                             int B, int V, int D, int h, int w, cudaStream_t st) {
    dim3 grid((h * w + kPixThreads - 1) / kPixThreads, V - 1, B);
    pair_score_kernel<C><<<grid, kPixThreads, 0, st>>>(feat, relproj, hs, score, V, D, h, w);
    ADAMVS_LAUNCH_RESULT();
}

template <int C, int VS>
static int launch_fused_volume(const float* feat, const float* relproj, const HypSpec& hs, const float* weights,
                               int eps_mode, float* volume, int B, int D, int h, int w, cudaStream_t st) {
    dim3 grid((h * w + kPixThreads - 1) / kPixThreads, 1, B);
    fused_volume_kernel<C, VS><<<grid, kPixThreads, 0, st>>>(feat, relproj, hs, weights, eps_mode, volume, D, h, w);
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
    ADAMVS_CHECK_ARG(feat && relproj && score && B > 0 && B <= 65535 && V >= 2 && V - 1 <= kMaxViews);
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
