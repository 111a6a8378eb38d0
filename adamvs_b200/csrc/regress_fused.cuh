// Output layer + online depth regression, and weight packing — shared by regnet.cu and msrednet.cu.
#pragma once
#include "common.cuh"

namespace adamvs {

// ------------------------------------------------------------------------------------------------
// 8: logit + online regression.  State per output pixel: (m, s, ws) for the softmax convention
// (running max logit, sum exp(l-m), sum d*exp(l-m)) or (emax, esum, dsum) for the reference's
// un-shifted predict convention.  Plane 0 initialises, plane D-1 finalises into depth/conf.
// ------------------------------------------------------------------------------------------------
// The 72+1 output-layer scalars ([8,1,3,3] ConvTranspose2d and [1,8,3,3] Conv2d are both ci*9+tap)
// are staged in shared memory by each block straight from the reference-layout tensors.
struct OutWeights { const float* w; const float* b; };

__device__ __forceinline__ void stage_out_weights(const OutWeights& ow, float* s) {
    if (threadIdx.x < 72) s[threadIdx.x] = __ldg(ow.w + threadIdx.x);
    if (threadIdx.x == 72) s[72] = __ldg(ow.b);
    __syncthreads();
}

struct RegressState { float* s0; float* s1; float* s2; };

__device__ __forceinline__ void regress_update(const RegressState& st, size_t o, float logit, float dval, int k, int D,
                                               int prob_mode, float* depth, float* conf) {
    float a0, a1, a2;
    if (k == 0) { a0 = prob_mode == ADAMVS_PROB_SOFTMAX ? -INFINITY : 0.f; a1 = 0.f; a2 = 0.f; }
    else { a0 = st.s0[o]; a1 = st.s1[o]; a2 = st.s2[o]; }
    if (prob_mode == ADAMVS_PROB_SOFTMAX) {
        const float m = fmaxf(a0, logit);
        const float scale = expf(a0 - m);           // 0 when a0 = -inf
        const float e = expf(logit - m);
        a1 = a1 * scale + e;
        a2 = a2 * scale + dval * e;
        a0 = m;
        if (k == D - 1) { depth[o] = a2 / a1; conf[o] = 1.f / a1; return; }
    } else {
        const float e = expf(logit);
        a0 = (a0 < e) ? e : a0;                     // adamvs.py:518-519
        a2 = dval * e + a2;                         // adamvs.py:524
        a1 = a1 + e;                                // adamvs.py:527
        if (k == D - 1) { const float den = a1 + 1e-10f; depth[o] = a2 / den; conf[o] = a0 / den; return; }
    }
    st.s0[o] = a0; st.s1[o] = a1; st.s2[o] = a2;
}

// The same update in three phases, for kernels that handle several pixels per thread: the compiler does not move a
// global load above an earlier global store, so load -> exp -> store per pixel serialises the pixels' L2 round trips;
// all loads first, all stores last keeps them in flight together.
struct RegressAcc { float a0, a1, a2; };

__device__ __forceinline__ RegressAcc regress_load(const RegressState& st, size_t o, int k, int prob_mode) {
    RegressAcc r;
    if (k == 0) { r.a0 = prob_mode == ADAMVS_PROB_SOFTMAX ? -INFINITY : 0.f; r.a1 = 0.f; r.a2 = 0.f; }
    else { r.a0 = st.s0[o]; r.a1 = st.s1[o]; r.a2 = st.s2[o]; }
    return r;
}
// exp on the SFU (ex2.approx after a multiply by log2 e): relative error <= 2^-21 + |x| 2^-24, i.e. ~1e-6 on the
// un-shifted exp(logit) of the predict convention at |logit| = 20 and less on the shifted softmax terms (arguments <= 0) -
// two orders below the 1e-4 probability tolerance, at a fifth of the instructions of expf.
__device__ __forceinline__ void regress_step(RegressAcc& r, float logit, float dval, int prob_mode) {
    if (prob_mode == ADAMVS_PROB_SOFTMAX) {
        const float m = fmaxf(r.a0, logit);
        const float scale = __expf(r.a0 - m);       // 0 when a0 = -inf
        const float e = __expf(logit - m);
        r.a1 = r.a1 * scale + e;
        r.a2 = r.a2 * scale + dval * e;
        r.a0 = m;
    } else {
        const float e = __expf(logit);
        r.a0 = (r.a0 < e) ? e : r.a0;               // adamvs.py:518-519
        r.a2 = dval * e + r.a2;                     // adamvs.py:524
        r.a1 = r.a1 + e;                            // adamvs.py:527
    }
}
__device__ __forceinline__ void regress_store(const RegressState& st, size_t o, const RegressAcc& r, int k, int D, int prob_mode,
                                              float* depth, float* conf) {
    if (k == D - 1) {
        if (prob_mode == ADAMVS_PROB_SOFTMAX) { depth[o] = r.a2 / r.a1; conf[o] = 1.f / r.a1; }
        else { const float den = r.a1 + 1e-10f; depth[o] = r.a2 / den; conf[o] = r.a0 / den; }
        return;
    }
    st.s0[o] = r.a0; st.s1[o] = r.a1; st.s2[o] = r.a2;
}

// logit = conv3x3(y [+ y2]; 8->1) + b at the same resolution.  `flip` reads the taps mirrored, which turns the
// correlation into PyTorch's stride-1 ConvTranspose2d (MS-REDNet's output layer, models/msrednet.py:351).
template <bool HAS_Y2, bool FLIP>
static __global__ void __launch_bounds__(128)
out_conv_regress_kernel(const float* __restrict__ y, const float* __restrict__ y2, OutWeights ow, HypSpec hs,
                        int prob_mode, RegressState st,
                        float* __restrict__ depth, float* __restrict__ conf, float* __restrict__ logits_out,
                        int k, int D, int h, int w) {
    __shared__ float sw[73];
    stage_out_weights(ow, sw);
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= w) return;
    const size_t hw = (size_t)h * w;
    float acc = sw[72];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
        const float* p = y + ((size_t)b * 8 + ci) * hw;
        const float* p2 = HAS_Y2 ? y2 + ((size_t)b * 8 + ci) * hw : nullptr;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int gy = yy + ky - 1;
            if (gy < 0 || gy >= h) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int gx = x + kx - 1;
                if (gx < 0 || gx >= w) continue;
                float v = __ldg(p + (size_t)gy * w + gx);
                if (HAS_Y2) v += __ldg(p2 + (size_t)gy * w + gx);
                const int tap = ky * 3 + kx;
                acc = fmaf(v, sw[ci * 9 + (FLIP ? 8 - tap : tap)], acc);
            }
        }
    }
    const int pix = yy * w + x;
    const size_t o = (size_t)b * hw + pix;
    if (logits_out) logits_out[((size_t)b * D + k) * hw + pix] = acc;
    const HypLine line = hyp_line(hs, b, pix, (int)hw, D);
    regress_update(st, o, acc, hyp_at(line, k), k, D, prob_mode, depth, conf);
}

// ------------------------------------------------------------------------------------------------
// weight packing: reference layouts -> [ci][tap][co]
// ------------------------------------------------------------------------------------------------
// `neg_first` input channels are packed with the opposite sign: conv(-x, W) = conv(x, -W) (MS-REDNet feeds -cost).
static __global__ void pack_conv_kernel(const float* __restrict__ w, float* __restrict__ pk, int cout, int cin, int transposed,
                                        int neg_first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * cin * 9) return;
    const int co = i % cout, t = (i / cout) % 9, ci = i / (cout * 9);
    const float v = transposed ? w[((size_t)ci * cout + co) * 9 + t] : w[((size_t)co * cin + ci) * 9 + t];
    pk[i] = ci < neg_first ? -v : v;
}

}  // namespace adamvs
