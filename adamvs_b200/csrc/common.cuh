// Shared device helpers for the adamvs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/adamvs_b200.h"

#define ADAMVS_CHECK_ARG(cond) do { if (!(cond)) return ADAMVS_EINVAL; } while (0)
#define ADAMVS_LAUNCH_RESULT() do { cudaError_t e__ = cudaGetLastError(); return e__ == cudaSuccess ? 0 : (int)e__; } while (0)

namespace adamvs {

// ---------------------------------------------------------------------------------------------
// Depth hypotheses (reference: models/module.py:628-663).  Evaluated on the fly; the [B,D,h,w]
// tensor of the reference is never materialised.
// ---------------------------------------------------------------------------------------------
struct HypSpec {
    int mode;                 // ADAMVS_HYP_*
    const float* src;         // depth_values [B,ncol] or cur_depth [B,h,w]
    int ncol;
    const float* half_range;  // device scalar (PER_PIXEL only)
};

// lo/step for one (batch item, pixel). d_k = lo + k*step with the reference's two roundings
// (mul then add, no contraction).
struct HypLine { float lo, step; };

__device__ __forceinline__ HypLine hyp_line(const HypSpec& hs, int b, int pix /* y*w+x */, int hw, int D) {
    HypLine l;
    if (hs.mode == ADAMVS_HYP_PLANES) {
        const float lo = __ldg(hs.src + (size_t)b * hs.ncol);
        const float hi = __ldg(hs.src + (size_t)b * hs.ncol + 1);
        l.lo = lo;
        l.step = __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1));
    } else {
        const float cur = __ldg(hs.src + (size_t)b * hw + pix);
        const float hr = __ldg(hs.half_range);
        const float lo = __fsub_rn(cur, hr);
        const float hi = __fadd_rn(cur, hr);
        l.lo = lo;
        l.step = __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1));
    }
    return l;
}

__device__ __forceinline__ float hyp_at(const HypLine& l, int k) {
    return __fadd_rn(l.lo, __fmul_rn((float)k, l.step));
}

// ---------------------------------------------------------------------------------------------
// Homography (reference: models/module.py:539-556).  P = {r00..r22, t0,t1,t2}.
// q = R*(x,y,1); (X,Y,Z) = q*d + t; u = X/Z, v = Y/Z in source pixel coordinates.  The reference's
// normalise/un-normalise round trip is an identity up to ~1e-6 relative and is skipped.
// ---------------------------------------------------------------------------------------------
struct Ray { float qx, qy, qz, tx, ty, tz; };

__device__ __forceinline__ Ray make_ray(const float* __restrict__ P, float x, float y) {
    Ray r;
    // same association as torch.matmul(rot, [x,y,1]): r0*x + r1*y + r2*1, accumulated left to right
    r.qx = __fadd_rn(__fadd_rn(__fmul_rn(P[0], x), __fmul_rn(P[1], y)), P[2]);
    r.qy = __fadd_rn(__fadd_rn(__fmul_rn(P[3], x), __fmul_rn(P[4], y)), P[5]);
    r.qz = __fadd_rn(__fadd_rn(__fmul_rn(P[6], x), __fmul_rn(P[7], y)), P[8]);
    r.tx = P[9]; r.ty = P[10]; r.tz = P[11];
    return r;
}

// Bilinear tap set with per-corner zero padding (grid_sample, padding_mode='zeros',
// align_corners=True).  Offsets are clamped into the image so that a load is always
// legal; out-of-image corners carry weight 0.  A sample behind the source camera (finite Z < 0) is treated as the
// reference treats it: X/Z, Y/Z are taken as they come and sampled where they land (models/module.py:553-556 has
// no sign test).  Only non-finite coordinates (Z = 0, NaN inputs), for which the reference's grid_sample result is
// platform dependent (NaN on CPU), give all-zero weights here — the one documented deviation.
struct Taps {
    int o00, o01, o10, o11;   // y*w+x offsets
    float w00, w01, w10, w11;
};

__device__ __forceinline__ Taps make_taps(const Ray& r, float d, int h, int w) {
    const float X = __fadd_rn(__fmul_rn(r.qx, d), r.tx);
    const float Y = __fadd_rn(__fmul_rn(r.qy, d), r.ty);
    const float Z = __fadd_rn(__fmul_rn(r.qz, d), r.tz);
    float u = __fdiv_rn(X, Z);
    float v = __fdiv_rn(Y, Z);
    Taps t;
    const bool ok = (u > -1.f) && (u < (float)w) && (v > -1.f) && (v < (float)h);      // false for NaN / inf
    if (!ok) { u = -2.f; v = -2.f; }
    const float fu = floorf(u), fv = floorf(v);
    const int x0 = (int)fu, y0 = (int)fv;
    const float ax = u - fu, ay = v - fv;
    const float bx = 1.f - ax, by = 1.f - ay;
    const bool x0in = (x0 >= 0) & (x0 < w), x1in = (x0 + 1 >= 0) & (x0 + 1 < w);
    const bool y0in = (y0 >= 0) & (y0 < h), y1in = (y0 + 1 >= 0) & (y0 + 1 < h);
    const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x0 + 1, 0), w - 1);
    const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y0 + 1, 0), h - 1);
    t.o00 = cy0 * w + cx0; t.o01 = cy0 * w + cx1; t.o10 = cy1 * w + cx0; t.o11 = cy1 * w + cx1;
    t.w00 = (ok && x0in && y0in) ? bx * by : 0.f;
    t.w01 = (ok && x1in && y0in) ? ax * by : 0.f;
    t.w10 = (ok && x0in && y1in) ? bx * ay : 0.f;
    t.w11 = (ok && x1in && y1in) ? ax * ay : 0.f;
    return t;
}

// ---------------------------------------------------------------------------------------------
// align_corners=False bilinear source index (ATen area_pixel_compute_source_index, used by
// upsample_bilinear2d): src = scale*(dst+0.5)-0.5 clamped at 0; i1 = min(i0+1, n-1).
// ---------------------------------------------------------------------------------------------
struct Lerp { int i0, i1; float l0, l1; };

__device__ __forceinline__ Lerp lerp_index(int dst, float scale, int n_in) {
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    Lerp r;
    r.i0 = min((int)s, n_in - 1);
    r.i1 = min(r.i0 + 1, n_in - 1);
    r.l1 = s - (float)r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}

// Gate non-linearities on the SFU: exactly one MUFU.EX2 and one MUFU.RCP each (ex2.approx.ftz / rcp.approx.ftz through
// inline PTX).  __expf / __fdividef compile to the same two MUFU operations plus a denormal-range guard (FSETP + two
// predicated FMULs per exp) that the gates do not need - flushing exp(-x) < 2^-126 to zero gives sigmoid = 1, tanh = +-1
// exactly as the limit does - and that was a sixth of the GRU epilogue's instructions (they run 64x per thread per tile).
// Absolute error ~1e-7 on values in [0,1] / [-1,1], far below the parity bar.  Saturation: ex2 -> +inf gives rcp -> 0.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_f(float x) { return fmaf(-2.f, rcp_ftz(1.f + ex2_ftz(2.8853900817779268f * x)), 1.f); }

}  // namespace adamvs
