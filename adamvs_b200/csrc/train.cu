// Training path: the backward kernels of the cascade cost-volume hot path, so that the reference's
// train_whu.py --mode train runs against the drop-in models.
//
//   K1 / K2 backward   adamvs_pair_score_bwd_f32, adamvs_fused_volume_bwd_f32: the sampling grid is built under
//                      torch.no_grad() in the reference (models/module.py:538), so the only gradients are with respect to the
//                      features - a bilinear scatter-add into the source maps, a gather into the reference map - and, for K2,
//                      the per-view aggregation weights.
//   K3 (recurrent regulariser, BPTT over the planes) is composed in adamvs_b200/autograd.py from three convolution kernels
//                      with run-time channel counts: adamvs_conv2d_f32 (3x3 conv stride 1 / 2 and the stride-2 transposed conv;
//                      every data gradient is one of these with the weight tensor re-read in the transposed role) and
//                      adamvs_conv2d_wgrad_f32 (the weight gradient, a reduction over batch and pixels).
//   K4 backward        adamvs_softmax_expect_f32 / adamvs_softmax_expect_bwd_f32: softmax over D, expectation over a
//                      materialised hypothesis tensor, max; gradients to the logits and to the hypotheses (through which the
//                      previous stage's depth map receives its gradient - the reference does not detach it, adamvs.py:365).
// Plain FFMA kernels: correctness and coverage first; the inference kernels (costvolume.cu, regnet.cu) stay the fast path.
#include "common.cuh"

namespace adamvs {

// ------------------------------------------------------------------------------------------------------------------
// y[n,co,oy,ox] = act(bias[co] + sum_ci sum_ky,kx x[n,ci,iy,ix] * W)
//   conv (TRANSPOSED = false):  iy = oy*STRIDE - 1 + ky, W = w[((co*Cin + ci)*3 + ky)*3 + kx]      (Conv2d [Cout,Cin,3,3])
//   transposed conv (stride 2, padding 1, output_padding 1):  oy + 1 - ky = 2*iy,
//                               W = w[((ci*Cout + co)*3 + ky)*3 + kx]                              (ConvTranspose2d [Cin,Cout,3,3])
// Block = 16x16 output pixels x 8 output channels of one batch item; input channels in chunks of 8 through shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGT = 16, kGCO = 8, kGCI = 8;

template <int STRIDE, bool TRANSPOSED>
__global__ void __launch_bounds__(kGT * kGT)
conv2d_generic_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ y, int Cin, int Cout, int hin, int win, int hout, int wout, int relu) {
    constexpr int IH = TRANSPOSED ? kGT / 2 + 1 : kGT * STRIDE + 2;
    constexpr int IP = IH + 1;
    __shared__ float sIn[kGCI][IH][IP];
    __shared__ float sW[kGCI][9][kGCO];
    const int tid = threadIdx.x, tx = tid % kGT, ty = tid / kGT;
    const int tiles_x = (wout + kGT - 1) / kGT;
    const int ox0 = (blockIdx.x % tiles_x) * kGT, oy0 = (blockIdx.x / tiles_x) * kGT;
    const int co0 = blockIdx.y * kGCO, n = blockIdx.z;
    const int ix0 = TRANSPOSED ? ox0 / 2 : ox0 * STRIDE - 1, iy0 = TRANSPOSED ? oy0 / 2 : oy0 * STRIDE - 1;
    const int ox = ox0 + tx, oy = oy0 + ty;
    float acc[kGCO];
#pragma unroll
    for (int c = 0; c < kGCO; ++c) acc[c] = 0.f;
    for (int ci0 = 0; ci0 < Cin; ci0 += kGCI) {
        __syncthreads();
        for (int i = tid; i < kGCI * IH * IH; i += kGT * kGT) {
            const int cx = i % IH, r = (i / IH) % IH, c = i / (IH * IH);
            const int gy = iy0 + r, gx = ix0 + cx, ci = ci0 + c;
            sIn[c][r][cx] = (ci < Cin && gy >= 0 && gy < hin && gx >= 0 && gx < win)
                                ? __ldg(x + (((size_t)n * Cin + ci) * hin + gy) * win + gx) : 0.f;
        }
        for (int i = tid; i < kGCI * 9 * kGCO; i += kGT * kGT) {
            const int co = i % kGCO, t = (i / kGCO) % 9, c = i / (kGCO * 9);
            const int ci = ci0 + c, cog = co0 + co;
            float v = 0.f;
            if (ci < Cin && cog < Cout) v = TRANSPOSED ? __ldg(w + ((size_t)ci * Cout + cog) * 9 + t) : __ldg(w + ((size_t)cog * Cin + ci) * 9 + t);
            sW[c][t][co] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < kGCI; ++c) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float v;
                    if (TRANSPOSED) {
                        const int t_y = oy + 1 - ky, t_x = ox + 1 - kx;
                        const bool ok = !(t_y & 1) && !(t_x & 1) && t_y >= 0 && t_x >= 0;
                        const int ly = (t_y >> 1) - iy0, lx = (t_x >> 1) - ix0;       // zero beyond the image: loaded as 0
                        v = (ok && ly >= 0 && ly < IH && lx >= 0 && lx < IH) ? sIn[c][ly][lx] : 0.f;
                    } else {
                        v = sIn[c][ty * STRIDE + ky][tx * STRIDE + kx];
                    }
#pragma unroll
                    for (int co = 0; co < kGCO; ++co) acc[co] = fmaf(v, sW[c][ky * 3 + kx][co], acc[co]);
                }
            }
        }
    }
    if (ox >= wout || oy >= hout) return;
#pragma unroll
    for (int co = 0; co < kGCO; ++co) {
        if (co0 + co >= Cout) break;
        float r = acc[co] + (bias ? __ldg(bias + co0 + co) : 0.f);
        if (relu) r = fmaxf(r, 0.f);
        y[(((size_t)n * Cout + co0 + co) * hout + oy) * wout + ox] = r;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of a 3x3 convolution (padding 1, stride 1 | 2):
//     gw[co,ci,ky,kx] += sum_{n,oy,ox} gy[n,co,oy,ox] * x[n,ci,oy*STRIDE - 1 + ky, ox*STRIDE - 1 + kx]
// The stride-2 transposed convolution's weight gradient is the same sum with the roles exchanged (x' = its output
// gradient, gy' = its input), which yields the [Cin,Cout,3,3] layout directly.  A block owns one 16x16 tile of gy of one
// batch item and walks all (8 co) x (8 ci) chunk pairs; its 256 threads share the 576 (co, ci, tap) sums of a pair and
// add them to gw with atomics (gw is zeroed by the caller; fp32 summation order is not deterministic).
// ------------------------------------------------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(256)
conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw,
                    int Cin, int Cout, int hin, int win, int hout, int wout) {
    constexpr int IH = kGT * STRIDE + 2, IP = IH + 1;
    __shared__ float sX[kGCI][IH][IP];
    __shared__ float sG[kGCO][kGT * kGT + 1];
    const int tid = threadIdx.x;
    const int tiles_x = (wout + kGT - 1) / kGT;
    const int ox0 = (blockIdx.x % tiles_x) * kGT, oy0 = (blockIdx.x / tiles_x) * kGT, n = blockIdx.y;
    const int ix0 = ox0 * STRIDE - 1, iy0 = oy0 * STRIDE - 1;
    // this thread's (co, ci, tap) triples of a chunk pair: t, t + 256, t + 512 (< 576)
    for (int ci0 = 0; ci0 < Cin; ci0 += kGCI) {
        __syncthreads();
        for (int i = tid; i < kGCI * IH * IH; i += 256) {
            const int cx = i % IH, r = (i / IH) % IH, c = i / (IH * IH);
            const int yy = iy0 + r, xx = ix0 + cx, ci = ci0 + c;
            sX[c][r][cx] = (ci < Cin && yy >= 0 && yy < hin && xx >= 0 && xx < win)
                               ? __ldg(x + (((size_t)n * Cin + ci) * hin + yy) * win + xx) : 0.f;
        }
        for (int co0 = 0; co0 < Cout; co0 += kGCO) {
            __syncthreads();
            for (int i = tid; i < kGCO * kGT * kGT; i += 256) {
                const int p = i % (kGT * kGT), c = i / (kGT * kGT);
                const int oy = oy0 + p / kGT, ox = ox0 + p % kGT, co = co0 + c;
                sG[c][p] = (co < Cout && oy < hout && ox < wout) ? __ldg(gy + (((size_t)n * Cout + co) * hout + oy) * wout + ox) : 0.f;
            }
            __syncthreads();
#pragma unroll 1
            for (int t = tid; t < kGCO * kGCI * 9; t += 256) {
                const int tap = t % 9, c = (t / 9) % kGCI, co = t / (9 * kGCI);
                const int ky = tap / 3, kx = tap % 3;
                float s = 0.f;
#pragma unroll 4
                for (int p = 0; p < kGT * kGT; ++p)
                    s = fmaf(sG[co][p], sX[c][(p / kGT) * STRIDE + ky][(p % kGT) * STRIDE + kx], s);
                if (co0 + co < Cout && ci0 + c < Cin) atomicAdd(gw + ((size_t)(co0 + co) * Cin + ci0 + c) * 9 + tap, s);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K1 backward.  score[b,v,k,p] = (1/C) sum_c ref[c,p] * warp_v[c,k,p]  (adamvs.py:270-272)
//   g_ref[c,p]        += (g/C) * warp_v[c,k,p]
//   g_src_v[c, tap]   += (g/C) * ref[c,p] * tap weight            (scatter-add over the 4 bilinear corners)
// One thread per (b, v, k, p); g_feat [B,V,C,h,w] is zeroed by the caller.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pair_score_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ relproj, HypSpec hs,
                      const float* __restrict__ g_score, float* __restrict__ g_feat, int V, int C, int D, int h, int w) {
    const int hw = h * w, Vs = V - 1;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y, bv = blockIdx.z, b = bv / Vs, v = bv - b * Vs;
    if (pix >= hw) return;
    const int x = pix % w, y = pix / w;
    const float g = g_score[(((size_t)b * Vs + v) * D + k) * hw + pix] / (float)C;
    if (g == 0.f) return;
    const Ray ray = make_ray(relproj + ((size_t)b * Vs + v) * 12, (float)x, (float)y);
    const Taps t = make_taps(ray, hyp_at(hyp_line(hs, b, pix, hw, D), k), h, w);
    const float* ref = feat + ((size_t)b * V) * C * hw + pix;
    const float* src = feat + ((size_t)b * V + v + 1) * C * hw;
    float* gref = g_feat + ((size_t)b * V) * C * hw + pix;
    float* gsrc = g_feat + ((size_t)b * V + v + 1) * C * hw;
    for (int c = 0; c < C; ++c) {
        const float* sc = src + (size_t)c * hw;
        const float wv = t.w00 * __ldg(sc + t.o00) + t.w01 * __ldg(sc + t.o01) + t.w10 * __ldg(sc + t.o10) + t.w11 * __ldg(sc + t.o11);
        atomicAdd(gref + (size_t)c * hw, g * wv);
        const float gr = g * __ldg(ref + (size_t)c * hw);
        float* gc = gsrc + (size_t)c * hw;
        if (t.w00 != 0.f) atomicAdd(gc + t.o00, gr * t.w00);
        if (t.w01 != 0.f) atomicAdd(gc + t.o01, gr * t.w01);
        if (t.w10 != 0.f) atomicAdd(gc + t.o10, gr * t.w10);
        if (t.w11 != 0.f) atomicAdd(gc + t.o11, gr * t.w11);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K2 backward.  F[c,k,p] = (start + R[c,p] * S[c,k,p]) / den,  S = sum_v w_v[p] * W_v[c,k,p],  den = sum_v w_v (numerator
// epsilon: start = 1e-5) or 1e-5 + sum_v w_v (denominator epsilon: start = 0)            (adamvs.py:285-301 / 495-512)
//   g_R[c,p]      += gF * S / den
//   g_W_v[c,k,p]   = gF * R * w_v / den        -> scattered to the source map through the bilinear weights
//   g_w_v[p]      += gF * (R * W_v - F) / den  (d(1/den)/dw_v = -1/den^2 for both conventions)
// One thread per (b, k, p).
// ------------------------------------------------------------------------------------------------------------------
template <int VS>
__global__ void __launch_bounds__(128)
fused_volume_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ relproj, HypSpec hs,
                        const float* __restrict__ weights, int eps_mode, const float* __restrict__ g_vol,
                        float* __restrict__ g_feat, float* __restrict__ g_weights, int C, int D, int h, int w) {
    constexpr int V = VS + 1;
    const int hw = h * w;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y, b = blockIdx.z;
    if (pix >= hw) return;
    const int x = pix % w, y = pix / w;
    const float d = hyp_at(hyp_line(hs, b, pix, hw, D), k);
    Taps t[VS];
    float wv[VS], wsum = 0.f, gw[VS];
#pragma unroll
    for (int v = 0; v < VS; ++v) {
        wv[v] = __ldg(weights + ((size_t)b * VS + v) * hw + pix);
        wsum += wv[v];
        gw[v] = 0.f;
        t[v] = make_taps(make_ray(relproj + ((size_t)b * VS + v) * 12, (float)x, (float)y), d, h, w);
    }
    const bool eps_num = (eps_mode == ADAMVS_EPS_NUMERATOR);
    const float inv = 1.f / (eps_num ? wsum : (1e-5f + wsum));
    const float start = eps_num ? 1e-5f : 0.f;
    const float* ref = feat + ((size_t)b * V) * C * hw + pix;
    float* gref = g_feat + ((size_t)b * V) * C * hw + pix;
    for (int c = 0; c < C; ++c) {
        const float gF = g_vol[(((size_t)b * C + c) * D + k) * hw + pix];
        const float R = __ldg(ref + (size_t)c * hw);
        float Wv[VS], S = 0.f;
#pragma unroll
        for (int v = 0; v < VS; ++v) {
            const float* sc = feat + (((size_t)b * V + v + 1) * C + c) * hw;
            Wv[v] = t[v].w00 * __ldg(sc + t[v].o00) + t[v].w01 * __ldg(sc + t[v].o01) + t[v].w10 * __ldg(sc + t[v].o10) + t[v].w11 * __ldg(sc + t[v].o11);
            S = fmaf(wv[v], Wv[v], S);
        }
        const float F = fmaf(R, S, start) * inv;
        atomicAdd(gref + (size_t)c * hw, gF * S * inv);
#pragma unroll
        for (int v = 0; v < VS; ++v) {
            gw[v] = fmaf(gF * inv, R * Wv[v] - F, gw[v]);
            const float gs = gF * R * wv[v] * inv;
            float* gc = g_feat + (((size_t)b * V + v + 1) * C + c) * hw;
            if (t[v].w00 != 0.f) atomicAdd(gc + t[v].o00, gs * t[v].w00);
            if (t[v].w01 != 0.f) atomicAdd(gc + t[v].o01, gs * t[v].w01);
            if (t[v].w10 != 0.f) atomicAdd(gc + t[v].o10, gs * t[v].w10);
            if (t[v].w11 != 0.f) atomicAdd(gc + t[v].o11, gs * t[v].w11);
        }
    }
#pragma unroll
    for (int v = 0; v < VS; ++v) atomicAdd(g_weights + ((size_t)b * VS + v) * hw + pix, gw[v]);
}

// ------------------------------------------------------------------------------------------------------------------
// K4 (training form): p = softmax_D(logits); depth = sum_k p_k * hyp_k; conf = max_k p_k   (adamvs.py:306-310, module.py:617-625)
// with a materialised hypothesis tensor at the logits' resolution, and its backward:
//   g_logit_j = p_j * ((hyp_j - depth) * g_depth + g_conf * p_m * ([j = m] / p_j - 1)),   g_hyp_j = p_j * g_depth
// One thread per output pixel, lanes along x, D coalesced loads per pass.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
softmax_expect_kernel(const float* __restrict__ logits, const float* __restrict__ hyp, float* __restrict__ depth,
                      float* __restrict__ conf, int D, int hw) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (pix >= hw) return;
    const float* p = logits + (size_t)n * D * hw + pix;
    const float* q = hyp + (size_t)n * D * hw + pix;
    float m = -INFINITY;
    for (int k = 0; k < D; ++k) m = fmaxf(m, __ldg(p + (size_t)k * hw));
    float s = 0.f;
    for (int k = 0; k < D; ++k) s += expf(__ldg(p + (size_t)k * hw) - m);
    float dsum = 0.f, pmax = 0.f;
    for (int k = 0; k < D; ++k) {
        const float pk = expf(__ldg(p + (size_t)k * hw) - m) / s;
        dsum += pk * __ldg(q + (size_t)k * hw);
        pmax = fmaxf(pmax, pk);
    }
    depth[(size_t)n * hw + pix] = dsum;
    conf[(size_t)n * hw + pix] = pmax;
}

__global__ void __launch_bounds__(128)
softmax_expect_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ hyp, const float* __restrict__ depth,
                          const float* __restrict__ g_depth, const float* __restrict__ g_conf,
                          float* __restrict__ g_logits, float* __restrict__ g_hyp, int D, int hw) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (pix >= hw) return;
    const float* p = logits + (size_t)n * D * hw + pix;
    const float* q = hyp + (size_t)n * D * hw + pix;
    float m = -INFINITY;
    int am = 0;
    for (int k = 0; k < D; ++k) { const float l = __ldg(p + (size_t)k * hw); if (l > m) { m = l; am = k; } }   // first maximum, like torch.max
    float s = 0.f;
    for (int k = 0; k < D; ++k) s += expf(__ldg(p + (size_t)k * hw) - m);
    const float gd = g_depth ? g_depth[(size_t)n * hw + pix] : 0.f;
    const float gc = g_conf ? g_conf[(size_t)n * hw + pix] : 0.f;
    const float dep = depth[(size_t)n * hw + pix];
    const float pm = 1.f / s;                                         // exp(m - m) / s
    for (int k = 0; k < D; ++k) {
        const float pk = expf(__ldg(p + (size_t)k * hw) - m) / s;
        const float hk = __ldg(q + (size_t)k * hw);
        g_logits[(size_t)n * D * hw + (size_t)k * hw + pix] = pk * (hk - dep) * gd + gc * pm * ((k == am ? 1.f : 0.f) - pk);
        if (g_hyp) g_hyp[(size_t)n * D * hw + (size_t)k * hw + pix] = pk * gd;
    }
}

}  // namespace adamvs

using namespace adamvs;

extern "C" int adamvs_conv2d_f32(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int Cout,
                                 int hin, int win, int stride, int transposed, int relu, void* stream) {
    ADAMVS_CHECK_ARG(x && w && y && N > 0 && N <= 65535 && Cin > 0 && Cout > 0 && hin > 0 && win > 0);
    ADAMVS_CHECK_ARG(transposed ? stride == 2 : (stride == 1 || (stride == 2 && hin % 2 == 0 && win % 2 == 0)));
    const int hout = transposed ? 2 * hin : hin / stride, wout = transposed ? 2 * win : win / stride;
    dim3 grid(((wout + kGT - 1) / kGT) * ((hout + kGT - 1) / kGT), (Cout + kGCO - 1) / kGCO, N);
    cudaStream_t st = (cudaStream_t)stream;
    if (transposed) conv2d_generic_kernel<2, true><<<grid, kGT * kGT, 0, st>>>(x, w, bias, y, Cin, Cout, hin, win, hout, wout, relu);
    else if (stride == 1) conv2d_generic_kernel<1, false><<<grid, kGT * kGT, 0, st>>>(x, w, bias, y, Cin, Cout, hin, win, hout, wout, relu);
    else conv2d_generic_kernel<2, false><<<grid, kGT * kGT, 0, st>>>(x, w, bias, y, Cin, Cout, hin, win, hout, wout, relu);
    ADAMVS_LAUNCH_RESULT();
}

extern "C" int adamvs_conv2d_wgrad_f32(const float* x, const float* gy, float* gw, int N, int Cin, int Cout,
                                       int hin, int win, int stride, void* stream) {
    ADAMVS_CHECK_ARG(x && gy && gw && N > 0 && N <= 65535 && Cin > 0 && Cout > 0 && hin > 0 && win > 0);
    ADAMVS_CHECK_ARG(stride == 1 || (stride == 2 && hin % 2 == 0 && win % 2 == 0));
    const int hout = hin / stride, wout = win / stride;
    dim3 grid(((wout + kGT - 1) / kGT) * ((hout + kGT - 1) / kGT), N, 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1) conv2d_wgrad_kernel<1><<<grid, 256, 0, st>>>(x, gy, gw, Cin, Cout, hin, win, hout, wout);
    else conv2d_wgrad_kernel<2><<<grid, 256, 0, st>>>(x, gy, gw, Cin, Cout, hin, win, hout, wout);
    ADAMVS_LAUNCH_RESULT();
}

static int check_hyp_train(int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range) {
    if (!hyp_src) return ADAMVS_EINVAL;
    if (hyp_mode == ADAMVS_HYP_PLANES) return hyp_ncol >= 2 ? 0 : ADAMVS_EINVAL;
    if (hyp_mode == ADAMVS_HYP_PER_PIXEL) return half_range ? 0 : ADAMVS_EINVAL;
    return ADAMVS_EINVAL;
}

extern "C" int adamvs_pair_score_bwd_f32(const float* feat, const float* relproj,
                                         int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                         const float* g_score, float* g_feat, int B, int V, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(feat && relproj && g_score && g_feat && B > 0 && V >= 2 && (long long)B * (V - 1) <= 65535 && D >= 2 && D <= 65535);
    if (int e = check_hyp_train(hyp_mode, hyp_src, hyp_ncol, half_range)) return e;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    dim3 grid((h * w + 127) / 128, D, B * (V - 1));
    pair_score_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(feat, relproj, hs, g_score, g_feat, V, C, D, h, w);
    ADAMVS_LAUNCH_RESULT();
}

extern "C" int adamvs_fused_volume_bwd_f32(const float* feat, const float* relproj,
                                           int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                           const float* weights, int eps_mode, const float* g_volume,
                                           float* g_feat, float* g_weights, int B, int V, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(feat && relproj && weights && g_volume && g_feat && g_weights && B > 0 && B <= 65535 && D >= 2 && D <= 65535);
    ADAMVS_CHECK_ARG(eps_mode == ADAMVS_EPS_NUMERATOR || eps_mode == ADAMVS_EPS_DENOMINATOR);
    if (int e = check_hyp_train(hyp_mode, hyp_src, hyp_ncol, half_range)) return e;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    dim3 grid((h * w + 127) / 128, D, B);
    cudaStream_t st = (cudaStream_t)stream;
#define ADAMVS_FVB(VV) fused_volume_bwd_kernel<VV><<<grid, 128, 0, st>>>(feat, relproj, hs, weights, eps_mode, g_volume, g_feat, g_weights, C, D, h, w)
    switch (V - 1) {
        case 1: ADAMVS_FVB(1); break;
        case 2: ADAMVS_FVB(2); break;
        case 3: ADAMVS_FVB(3); break;
        case 4: ADAMVS_FVB(4); break;
        case 5: ADAMVS_FVB(5); break;
        case 6: ADAMVS_FVB(6); break;
        default: return ADAMVS_EINVAL;
    }
#undef ADAMVS_FVB
    ADAMVS_LAUNCH_RESULT();
}

extern "C" int adamvs_softmax_expect_f32(const float* logits, const float* hyp, float* depth, float* conf,
                                         int N, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(logits && hyp && depth && conf && N > 0 && N <= 65535 && D >= 2 && h > 0 && w > 0);
    dim3 grid((h * w + 127) / 128, N, 1);
    softmax_expect_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(logits, hyp, depth, conf, D, h * w);
    ADAMVS_LAUNCH_RESULT();
}

extern "C" int adamvs_softmax_expect_bwd_f32(const float* logits, const float* hyp, const float* depth,
                                             const float* g_depth, const float* g_conf, float* g_logits, float* g_hyp,
                                             int N, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(logits && hyp && depth && g_logits && N > 0 && N <= 65535 && D >= 2 && h > 0 && w > 0);
    dim3 grid((h * w + 127) / 128, N, 1);
    softmax_expect_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(logits, hyp, depth, g_depth, g_conf, g_logits, g_hyp, D, h * w);
    ADAMVS_LAUNCH_RESULT();
}
