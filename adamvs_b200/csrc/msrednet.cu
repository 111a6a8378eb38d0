// K5/K6 — MS-REDNet (BASELINE config 5): variance cost volume and the four-level GroupNorm conv-GRU
// recurrent regulariser swept over the depth planes with the online regression folded into the last layer.
//
// Reference: slice_RED_Regularization.forward / RED_Regularization.forward (models/msrednet.py:355-372,
// 150-181), ConvGRUCell2 (models/module.py:54-106), ConvReLU / ConvTransReLU (module.py:264-270, 294-301),
// InferDepthNet.forward / DepthNet.forward (msrednet.py:379-436, 203-242).
//
// Per depth plane (x = -cost is folded into the packed weights: conv(-x, W) = conv(x, -W)):
//   c1 = relu(conv s2(x; C->16))  c2 = relu(conv s2(c1; 16->32))  c3 = relu(conv s2(c2; 32->64))
//   s4 = GRU4(c3, s4);  u3 = relu(convT s2(s4; 64->32))
//   s3 = GRU3(c2, s3);  u2 = relu(convT s2(u3 + s3; 32->16))
//   s2 = GRU2(c1, s2);  u1 = relu(convT s2(u2 + s2; 16->8))
//   s1 = GRU1(x,  s1);  logit = convT s1(u1 + s1; 8->1) + b;  online softmax / expectation / max
// GRU(x,h) with GroupNorm(1, HC) on reset gate, update gate and candidate:
//   f = conv(cat(x,h)) + b;  r = sig(GN_r(f[:HC]));  u = sig(GN_u(f[HC:]));
//   o = conv(cat(x, r*h)) + b;  h' = u*h + (1-u)*tanh(GN_o(o))
// A GroupNorm with one group normalises over the whole [HC,h,w] plane of a batch item, i.e. it needs the
// moments of a conv output before any of it can be consumed.  Each conv therefore writes its raw output
// and accumulates per-item moments in fp64 (warp partials -> atomics); a light elementwise kernel then
// applies normalisation, activation and the GRU blend.  All statistics buffers of a sweep are zeroed
// once up front, so the plane loop contains kernel launches only.
#include "conv3x3.cuh"
#include "conv3x3_tc.cuh"
#include <string.h>
#include "regress_fused.cuh"

namespace adamvs {

// ---- elementwise halves of the GroupNorm GRU ------------------------------------------------------
struct Moments { float mean, rstd; };

__device__ __forceinline__ Moments moments_of(const double* st, double n) {
    const double mean = st[0] / n;
    double var = st[1] / n - mean * mean;              // biased, like nn.GroupNorm
    var = var < 0.0 ? 0.0 : var;
    Moments m;
    m.mean = (float)mean;
    m.rstd = (float)(1.0 / sqrt(var + 1e-5));
    return m;
}

// f [B,2HC,hw] raw gate conv output -> rh = sig(GN_r(f[:HC])) * h,  u = sig(GN_u(f[HC:]))
static __global__ void __launch_bounds__(256)
gn_gates_kernel(const float* __restrict__ f, const double* __restrict__ stats, const float* __restrict__ gr,
                const float* __restrict__ br, const float* __restrict__ gu, const float* __restrict__ bu,
                const float* __restrict__ h, float* __restrict__ rh, float* __restrict__ u, int HC, int hw) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const double n = (double)HC * hw;
    const Moments mr = moments_of(stats + ((size_t)b * 2 + 0) * 2, n), mu = moments_of(stats + ((size_t)b * 2 + 1) * 2, n);
    const float ar = mr.rstd * __ldg(gr + c), cr = __ldg(br + c) - mr.mean * ar;
    const float au = mu.rstd * __ldg(gu + c), cu = __ldg(bu + c) - mu.mean * au;
    const float fr = f[((size_t)b * 2 * HC + c) * hw + i];
    const float fu = f[((size_t)b * 2 * HC + HC + c) * hw + i];
    const size_t o = ((size_t)b * HC + c) * hw + i;
    rh[o] = sigmoid_f(fmaf(fr, ar, cr)) * h[o];
    u[o] = sigmoid_f(fmaf(fu, au, cu));
}

// o [B,HC,hw] raw candidate conv output -> h = u*h + (1-u)*tanh(GN_o(o))   (in place on h)
static __global__ void __launch_bounds__(256)
gn_cand_kernel(const float* __restrict__ o, const double* __restrict__ stats, const float* __restrict__ g,
               const float* __restrict__ bt, const float* __restrict__ u, float* __restrict__ h, int HC, int hw) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const Moments m = moments_of(stats + ((size_t)b * 2) * 2, (double)HC * hw);
    const float a = m.rstd * __ldg(g + c), cc = __ldg(bt + c) - m.mean * a;
    const size_t idx = ((size_t)b * HC + c) * hw + i;
    const float uv = u[idx];
    h[idx] = uv * h[idx] + (1.f - uv) * tanh_f(fmaf(o[idx], a, cc));
}

// ---- y = relu(convT3x3 s2 p1 op1 (inA [+ inB]; CIN -> COUT)), no bias (ConvTransReLU) -------------------
// One thread per input pixel and block of 8 output channels: it owns the 2x2 outputs that (iy,ix) is the
// top-left contributor of (tap table in regnet.cu / SURVEY.md Appendix B).  wpk is [ci][tap][COUT].
template <int CIN, int COUT>
static __global__ void __launch_bounds__(128)
upconv_relu_kernel(const float* __restrict__ inA, const float* __restrict__ inB, const float* __restrict__ wpk,
                   float* __restrict__ out, int hin, int win) {
    constexpr int COB = 8;
    __shared__ float sW[CIN * 9 * COB];
    const int cob = blockIdx.y % (COUT / COB), iy = blockIdx.y / (COUT / COB);
    for (int i = threadIdx.x; i < CIN * 9 * COB; i += blockDim.x) sW[i] = __ldg(wpk + (size_t)(i / COB) * COUT + cob * COB + (i % COB));
    __syncthreads();
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    if (ix >= win) return;
    const size_t ip = (size_t)hin * win;
    const int wout = 2 * win;
    const size_t op = (size_t)4 * ip;
    const bool hx = ix + 1 < win, hy = iy + 1 < hin;
    float acc[4][COB];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < COB; ++c) acc[q][c] = 0.f;
    const size_t base = (size_t)b * CIN * ip + (size_t)iy * win + ix;
#pragma unroll 2
    for (int ci = 0; ci < CIN; ++ci) {
        const float* p = inA + base + (size_t)ci * ip;
        float v00 = __ldg(p);
        float v01 = hx ? __ldg(p + 1) : 0.f;
        float v10 = hy ? __ldg(p + win) : 0.f;
        float v11 = (hx && hy) ? __ldg(p + win + 1) : 0.f;
        if (inB) {
            const float* q = inB + base + (size_t)ci * ip;
            v00 += __ldg(q);
            if (hx) v01 += __ldg(q + 1);
            if (hy) v10 += __ldg(q + win);
            if (hx && hy) v11 += __ldg(q + win + 1);
        }
        const float* w = sW + ci * 9 * COB;
#pragma unroll
        for (int c = 0; c < COB; ++c) {
            acc[0][c] += v00 * w[4 * COB + c];
            acc[1][c] += v01 * w[3 * COB + c] + v00 * w[5 * COB + c];
            acc[2][c] += v10 * w[1 * COB + c] + v00 * w[7 * COB + c];
            acc[3][c] += v11 * w[0 * COB + c] + v10 * w[2 * COB + c] + v01 * w[6 * COB + c] + v00 * w[8 * COB + c];
        }
    }
#pragma unroll
    for (int c = 0; c < COB; ++c) {
        const size_t o = ((size_t)b * COUT + cob * COB + c) * op + (size_t)(2 * iy) * wout + 2 * ix;
        *reinterpret_cast<float2*>(out + o) = make_float2(fmaxf(acc[0][c], 0.f), fmaxf(acc[1][c], 0.f));
        *reinterpret_cast<float2*>(out + o + wout) = make_float2(fmaxf(acc[2][c], 0.f), fmaxf(acc[3][c], 0.f));
    }
}

template <int CIN, int COUT>
static cudaError_t launch_upconv(const float* inA, const float* inB, const float* wpk, float* out, int B, int hin, int win,
                                 cudaStream_t st) {
    dim3 grid((win + 127) / 128, hin * (COUT / 8), B);
    upconv_relu_kernel<CIN, COUT><<<grid, 128, 0, st>>>(inA, inB, wpk, out, hin, win);
    return cudaGetLastError();
}

// ---- one conv layer with its launch paths (tensor cores / TMA persistent FFMA / generic tiles) -------------
// Round 2: the GRU convolutions of the levels whose weights fit the tensor-core kernel's shared memory (levels 1-3:
// 8/16/32 hidden channels) run on conv3x3_tc.cuh like Ada-MVS's regulariser - same hi/lo tf32 split, fp32 accuracy -
// with an EPI_RAW_STATS epilogue (raw output + bias, GroupNorm moments as fp64 atomics).  A 64-channel gate
// convolution (N = 3 x 2 x 64 exceeds one MMA) runs as two 32-channel slices.  Planes below ~2 tiles per SM and level 4
// (128 input channels: 295 KB of split weights) keep the FFMA kernels.  ADAMVS_K3_MATH=ffma|tc forces one path (tests).
static int msred_math() {
    static const int m = [] {
        const char* e = getenv("ADAMVS_K3_MATH");
        return (e && !strcmp(e, "ffma")) ? 0 : (e && !strcmp(e, "tc")) ? 2 : 1;      // 0 FFMA, 1 auto, 2 tensor cores wherever possible
    }();
    return m;
}

template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI>
struct Layer {
    using L = ConvLayer<CA, CB, COUT, COB, STRIDE, EPI>;
    static constexpr bool TC_OK = STRIDE == 1 && EPI == EPI_RAW_STATS && CA + CB <= 64 && COUT <= 64;
    static constexpr int TC_COUT = COUT > 32 ? 32 : COUT;          // slice width
    using T = TcLayer<CA, CB, TC_OK ? TC_COUT : 8, TC_OK ? EPI : EPI_RELU>;
    ConvPlan plan{}, tplan{};
    ConvArgs args;
    bool tma = false, tc = false;
    void setup(const ConvArgs& a, int B, int depthA) {
        args = a;
        const bool aligned = ((reinterpret_cast<uintptr_t>(a.inA) | reinterpret_cast<uintptr_t>(a.inB) |
                               reinterpret_cast<uintptr_t>(a.out0) | reinterpret_cast<uintptr_t>(a.out1) |
                               reinterpret_cast<uintptr_t>(a.hstate) | reinterpret_cast<uintptr_t>(a.ugate)) % 16) == 0;
        tma = aligned && (a.win % 4 == 0) && (a.wout % 4 == 0) && L::plan(plan, a, B, depthA);
        if constexpr (TC_OK) {
            const long long px = (long long)a.hout * a.wout * B;
            if (tma && msred_math() != 0 && (msred_math() == 2 || px >= 30000)) {
                ConvArgs t = a;
                if (TC_COUT != COUT) { t.wpk_cout = COUT; t.out_cout = COUT; }
                tc = T::plan(tplan, t, B, depthA);
            }
        }
    }
    void set_stats(double* p) { args.stats = p; plan.args.stats = p; tplan.args.stats = p; }
    cudaError_t run(int B, int k, cudaStream_t st) {
        if constexpr (TC_OK) {
            if (tc) {
                tplan.args.k = k;
                for (int co = 0; co < COUT; co += TC_COUT) {
                    tplan.args.co_off = co;
                    cudaError_t e = T::launch(tplan, B, PREC_FP32X3, st);
                    if (e != cudaSuccess) return e;
                }
                return cudaSuccess;
            }
        }
        if (tma) { plan.args.k = k; return L::launch(plan, B, st); }
        ConvArgs a = args;
        a.inA = args.inA + (size_t)k * args.hin * args.win;        // plane k of a [.., D, h, w] volume (k = 0 otherwise)
        constexpr size_t smem = sizeof(float) * ((CA + CB) * 9 * COB + CK * TileGeom<STRIDE, 16, 16>::IH * TileGeom<STRIDE, 16, 16>::IP);
        auto kern = conv3x3_kernel<CA, CB, COUT, COB, STRIDE, EPI, 16, 16>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        dim3 grid(((a.wout + 15) / 16) * ((a.hout + 15) / 16), COUT / COB, B);
        kern<<<grid, TileGeom<STRIDE, 16, 16>::GROUP * (COB / COT), smem, st>>>(a);
        return cudaGetLastError();
    }
};

struct MsWorkspace {
    float *pk_c1, *pk_c2, *pk_c3, *pk_g[4], *pk_o[4], *pk_u3, *pk_u2, *pk_u1;
    float *c1, *c2, *c3, *s[4], *f[4], *o[4], *rh[4], *u[4], *up3, *up2, *up1, *r0, *r1, *r2;
    double* stats;            // [D][8][B][2][2]
    size_t total;
};

static MsWorkspace ms_carve(float* base, int B, int C, int D, int h, int w) {
    MsWorkspace ws;
    size_t off = 0;
    auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return p; };
    const size_t px[4] = {(size_t)h * w, (size_t)(h / 2) * (w / 2), (size_t)(h / 4) * (w / 4), (size_t)(h / 8) * (w / 8)};
    const int hc[4] = {8, 16, 32, 64};
    const int xin[4] = {C, 16, 32, 64};
    ws.pk_c1 = take((size_t)C * 9 * 16); ws.pk_c2 = take(16 * 9 * 32); ws.pk_c3 = take(32 * 9 * 64);
    for (int l = 0; l < 4; ++l) { ws.pk_g[l] = take((size_t)(xin[l] + hc[l]) * 9 * 2 * hc[l]); ws.pk_o[l] = take((size_t)(xin[l] + hc[l]) * 9 * hc[l]); }
    ws.pk_u3 = take(64 * 9 * 32); ws.pk_u2 = take(32 * 9 * 16); ws.pk_u1 = take(16 * 9 * 8);
    ws.c1 = take(B * 16 * px[1]); ws.c2 = take(B * 32 * px[2]); ws.c3 = take(B * 64 * px[3]);
    for (int l = 0; l < 4; ++l) {
        ws.s[l] = take(B * hc[l] * px[l]); ws.f[l] = take(B * 2 * hc[l] * px[l]); ws.o[l] = take(B * hc[l] * px[l]);
        ws.rh[l] = take(B * hc[l] * px[l]); ws.u[l] = take(B * hc[l] * px[l]);
    }
    ws.up3 = take(B * 32 * px[2]); ws.up2 = take(B * 16 * px[1]); ws.up1 = take(B * 8 * px[0]);
    ws.r0 = take(B * px[0]); ws.r1 = take(B * px[0]); ws.r2 = take(B * px[0]);
    ws.stats = reinterpret_cast<double*>(take((size_t)D * 8 * B * 4 * 2));
    ws.total = off;
    return ws;
}

}  // namespace adamvs

using namespace adamvs;

extern "C" size_t adamvs_regnet_msred_workspace_floats(int B, int C, int D, int h, int w) {
    if (B <= 0 || C <= 0 || D <= 0 || h <= 0 || w <= 0) return 0;
    return ms_carve(nullptr, B, C, D, h, w).total;
}

#define ADAMVS_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

template <int C>
static int run_msred(const float* volume, const adamvs_msred_weights* wts, const HypSpec& hs, int prob_mode,
                     MsWorkspace& ws, float* depth, float* conf, float* logits_out, int B, int D, int h, int w, cudaStream_t st) {
    const int hh[4] = {h, h / 2, h / 4, h / 8}, wwv[4] = {w, w / 2, w / 4, w / 8};
    size_t px[4];
    for (int l = 0; l < 4; ++l) px[l] = (size_t)hh[l] * wwv[l];
    const int hc[4] = {8, 16, 32, 64};
    const int xin[4] = {C, 16, 32, 64};

    auto pack = [&](const float* src, float* dst, int cout, int cin, int tr, int neg) {
        const int n = cout * cin * 9;
        pack_conv_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, cout, cin, tr, neg);
    };
    pack(wts->conv1_w, ws.pk_c1, 16, C, 0, C);          // conv1 and GRU1 see -cost
    pack(wts->conv2_w, ws.pk_c2, 32, 16, 0, 0);
    pack(wts->conv3_w, ws.pk_c3, 64, 32, 0, 0);
    for (int l = 0; l < 4; ++l) {
        pack(wts->gate_w[l], ws.pk_g[l], 2 * hc[l], xin[l] + hc[l], 0, l == 0 ? C : 0);
        pack(wts->out_w[l], ws.pk_o[l], hc[l], xin[l] + hc[l], 0, l == 0 ? C : 0);
    }
    pack(wts->up3_w, ws.pk_u3, 32, 64, 1, 0);
    pack(wts->up2_w, ws.pk_u2, 16, 32, 1, 0);
    pack(wts->up1_w, ws.pk_u1, 8, 16, 1, 0);
    ADAMVS_TRY(cudaGetLastError());
    for (int l = 0; l < 4; ++l) ADAMVS_TRY(cudaMemsetAsync(ws.s[l], 0, sizeof(float) * B * hc[l] * px[l], st));
    ADAMVS_TRY(cudaMemsetAsync(ws.stats, 0, sizeof(double) * (size_t)D * 8 * B * 4, st));

    // ---- layers (arguments fixed for the whole sweep; only the volume plane index and the stats slot change)
    Layer<C, 0, 16, 16, 2, EPI_RELU> Lc1;
    Layer<16, 0, 32, 16, 2, EPI_RELU> Lc2;
    Layer<32, 0, 64, 16, 2, EPI_RELU> Lc3;
    Layer<64, 64, 128, 16, 1, EPI_RAW_STATS> Lg4; Layer<64, 64, 64, 16, 1, EPI_RAW_STATS> Lo4;
    Layer<32, 32, 64, 16, 1, EPI_RAW_STATS> Lg3;  Layer<32, 32, 32, 16, 1, EPI_RAW_STATS> Lo3;
    Layer<16, 16, 32, 16, 1, EPI_RAW_STATS> Lg2;  Layer<16, 16, 16, 16, 1, EPI_RAW_STATS> Lo2;
    Layer<C, 8, 16, 16, 1, EPI_RAW_STATS> Lg1;    Layer<C, 8, 8, 8, 1, EPI_RAW_STATS> Lo1;

    auto conv_args = [&](const float* inA, int planesA, long long strideA_c, long long strideA_b, const float* inB, int planesB,
                         int lvl_in, int lvl_out, const float* wpk, const float* bias, float* out) {
        ConvArgs a{};
        a.inA = inA; a.planesA = planesA; a.strideA_c = strideA_c; a.strideA_b = strideA_b;
        a.inB = inB; a.planesB = planesB; a.strideB_c = (long long)px[lvl_in]; a.strideB_b = (long long)planesB * px[lvl_in];
        a.wpk = wpk; a.bias = bias; a.out0 = out;
        a.hin = hh[lvl_in]; a.win = wwv[lvl_in]; a.hout = hh[lvl_out]; a.wout = wwv[lvl_out];
        return a;
    };
    const long long vol_c = (long long)D * px[0], vol_b = (long long)C * D * px[0];
    Lc1.setup(conv_args(volume, C, vol_c, vol_b, nullptr, 0, 0, 1, ws.pk_c1, nullptr, ws.c1), B, D);
    Lc2.setup(conv_args(ws.c1, 16, px[1], 16 * px[1], nullptr, 0, 1, 2, ws.pk_c2, nullptr, ws.c2), B, 1);
    Lc3.setup(conv_args(ws.c2, 32, px[2], 32 * px[2], nullptr, 0, 2, 3, ws.pk_c3, nullptr, ws.c3), B, 1);
    auto gate = [&](auto& L, const float* x, int xp, long long xc, long long xb, int l, int depthA) {
        ConvArgs a = conv_args(x, xp, xc, xb, ws.s[l], hc[l], l, l, ws.pk_g[l], wts->gate_b[l], ws.f[l]);
        a.stats_split = 1;
        L.setup(a, B, depthA);
    };
    auto cand = [&](auto& L, const float* x, int xp, long long xc, long long xb, int l, int depthA) {
        ConvArgs a = conv_args(x, xp, xc, xb, ws.rh[l], hc[l], l, l, ws.pk_o[l], wts->out_b[l], ws.o[l]);
        a.stats_split = 0;
        L.setup(a, B, depthA);
    };
    gate(Lg4, ws.c3, 64, px[3], 64 * px[3], 3, 1); cand(Lo4, ws.c3, 64, px[3], 64 * px[3], 3, 1);
    gate(Lg3, ws.c2, 32, px[2], 32 * px[2], 2, 1); cand(Lo3, ws.c2, 32, px[2], 32 * px[2], 2, 1);
    gate(Lg2, ws.c1, 16, px[1], 16 * px[1], 1, 1); cand(Lo2, ws.c1, 16, px[1], 16 * px[1], 1, 1);
    gate(Lg1, volume, C, vol_c, vol_b, 0, D);      cand(Lo1, volume, C, vol_c, vol_b, 0, D);

    const OutWeights ow{wts->prob_w, wts->prob_b};
    const RegressState rs{ws.r0, ws.r1, ws.r2};

    auto gru = [&](auto& Lg, auto& Lo, int l, int k, int kvol) -> int {
        double* sg = ws.stats + ((size_t)k * 8 + 2 * l) * B * 4;
        double* so = ws.stats + ((size_t)k * 8 + 2 * l + 1) * B * 4;
        Lg.set_stats(sg);
        Lo.set_stats(so);
        ADAMVS_TRY(Lg.run(B, kvol, st));
        const int hw4 = (int)px[l];
        dim3 grid((hw4 + 255) / 256, hc[l], B);
        gn_gates_kernel<<<grid, 256, 0, st>>>(ws.f[l], sg, wts->rnorm_w[l], wts->rnorm_b[l], wts->unorm_w[l], wts->unorm_b[l],
                                              ws.s[l], ws.rh[l], ws.u[l], hc[l], hw4);
        ADAMVS_TRY(Lo.run(B, kvol, st));
        gn_cand_kernel<<<grid, 256, 0, st>>>(ws.o[l], so, wts->onorm_w[l], wts->onorm_b[l], ws.u[l], ws.s[l], hc[l], hw4);
        return 0;
    };

    for (int k = 0; k < D; ++k) {
        ADAMVS_TRY(Lc1.run(B, k, st));
        ADAMVS_TRY(Lc2.run(B, 0, st));
        ADAMVS_TRY(Lc3.run(B, 0, st));
        if (int e = gru(Lg4, Lo4, 3, k, 0)) return e;
        ADAMVS_TRY((launch_upconv<64, 32>(ws.s[3], nullptr, ws.pk_u3, ws.up3, B, hh[3], wwv[3], st)));
        if (int e = gru(Lg3, Lo3, 2, k, 0)) return e;
        ADAMVS_TRY((launch_upconv<32, 16>(ws.up3, ws.s[2], ws.pk_u2, ws.up2, B, hh[2], wwv[2], st)));
        if (int e = gru(Lg2, Lo2, 1, k, 0)) return e;
        ADAMVS_TRY((launch_upconv<16, 8>(ws.up2, ws.s[1], ws.pk_u1, ws.up1, B, hh[1], wwv[1], st)));
        if (int e = gru(Lg1, Lo1, 0, k, k)) return e;
        dim3 grid((w + 127) / 128, h, B);
        out_conv_regress_kernel<true, true><<<grid, 128, 0, st>>>(ws.up1, ws.s[0], ow, hs, prob_mode, rs, depth, conf, logits_out, k, D, h, w);
        ADAMVS_TRY(cudaGetLastError());
    }
    return 0;
}

extern "C" int adamvs_regnet_msred_f32(const float* volume, const adamvs_msred_weights* wts,
                                       int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                       int prob_mode, float* workspace, size_t workspace_floats,
                                       float* depth, float* conf, float* logits_out,
                                       int B, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(volume && wts && workspace && depth && conf && hyp_src);
    ADAMVS_CHECK_ARG(B > 0 && B <= 65535 && D >= 2 && h > 0 && w > 0 && (h % 8) == 0 && (w % 8) == 0 && h <= 8190);
    ADAMVS_CHECK_ARG(C == 8 || C == 16 || C == 32);
    ADAMVS_CHECK_ARG(prob_mode == ADAMVS_PROB_SOFTMAX || prob_mode == ADAMVS_PROB_EXP_EPS);
    ADAMVS_CHECK_ARG(hyp_mode == ADAMVS_HYP_PLANES ? hyp_ncol >= 2 : (hyp_mode == ADAMVS_HYP_PER_PIXEL && half_range));
    ADAMVS_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 16 == 0 && reinterpret_cast<uintptr_t>(volume) % 16 == 0);
    MsWorkspace ws = ms_carve(workspace, B, C, D, h, w);
    if (ws.total > workspace_floats) return ADAMVS_ENOSPACE;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 8: return run_msred<8>(volume, wts, hs, prob_mode, ws, depth, conf, logits_out, B, D, h, w, st);
        case 16: return run_msred<16>(volume, wts, hs, prob_mode, ws, depth, conf, logits_out, B, D, h, w, st);
        default: return run_msred<32>(volume, wts, hs, prob_mode, ws, depth, conf, logits_out, B, D, h, w, st);
    }
}
