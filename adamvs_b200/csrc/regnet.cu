// K3 (+K4 fused): the recurrent conv-GRU encoder-decoder regulariser, swept over the D depth planes,
// with the softmax depth regression folded into its last layer.
//
// Reference: CostRegNetRED.forward / SliceCostRegNetRED.forward (models/adamvs.py:172-195, 415-424),
// ConvGRUCell.forward (models/module.py:24-52), ConvReLU (module.py:264-270), regression
// (adamvs.py:306-310, 516-531; module.py:617-625).
//
// fp32 parity (depth 1e-4 rel / prob 1e-4 abs) rules out single-pass TF32/BF16 operands here
// (SURVEY.md §0), and with N = 8..32 output channels a 3xTF32 tcgen05 formulation is bound by the
// shared-memory read of the im2col A operand at ~1.3x the FFMA peak at best, so this
// is a register-tiled FFMA direct convolution: every thread owns a 4x2 pixel patch x 8 output
// channels (64 accumulators), input planes are staged in shared memory 8 channels at a time, weights
// sit in shared memory as [ci][tap][co] and are read as warp-wide broadcasts.
//
// Per depth plane (all on one stream, states/intermediates stay L2-resident):
//   1 conv1   x1  = relu(conv3x3(F_k))                                  C  -> 8
//   2 gates1  r,u = sigmoid(conv3x3(cat(x1,h1)) + b);  rh1 = r*h1        16 -> 16
//   3 cand1   h1  = u*h1 + (1-u)*tanh(conv3x3(cat(x1,rh1)) + b)          16 -> 8
//   4 conv2   x2  = relu(conv3x3 stride 2 (h1))                          8  -> 16
//   5 gates2 / 6 cand2 at half resolution                                32 -> 32 / 32 -> 16
//   7 up1     y   = relu(convT3x3 s2 (h2) + b + h1)                      16 -> 8
//   8 out     logit = convT3x3 s2 (y) + b  (stages 1-2) | conv3x3(y)+b (stage 3); online softmax update
#include "common.cuh"

namespace adamvs {

constexpr int PX = 4;     // thread patch width  (pixels, along x)
constexpr int PY = 2;     // thread patch height
constexpr int COT = 8;    // output channels per thread
constexpr int CK = 8;     // input channels staged per shared-memory chunk

enum { EPI_RELU = 0, EPI_GATES = 1, EPI_CAND = 2 };

struct ConvArgs {
    const float* inA; long long strideA_c, strideA_b;   // first  CA input channels: base, channel stride, batch stride
    const float* inB; long long strideB_c, strideB_b;   // next   CB input channels
    const float* wpk;      // packed weights [CIN][9][COUT]
    const float* bias;     // [COUT] or nullptr
    float* out0;           // RELU: out [B,COUT,hout,wout] | GATES: rh [B,HC,h,w] | CAND: h (read-modify-write)
    float* out1;           // GATES: u [B,HC,h,w]
    const float* hstate;   // GATES / CAND: h [B,HC,h,w]
    const float* ugate;    // CAND: u [B,HC,h,w]
    int hin, win, hout, wout;
};

template <int STRIDE, int TW, int TH>
struct TileGeom {
    static constexpr int IH = STRIDE == 1 ? TH + 2 : 2 * TH + 1;
    static constexpr int IW = STRIDE == 1 ? TW + 2 : 2 * TW + 1;
    static constexpr int IP = STRIDE == 1 ? TW + 4 : 2 * TW + 4;     // row pitch, multiple of 4 floats
    static constexpr int GROUP = (TW / PX) * (TH / PY);              // threads per output-channel group
};

template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI, int TW, int TH>
__global__ void __launch_bounds__((TW / PX) * (TH / PY) * (COB / COT))
conv3x3_kernel(ConvArgs a) {
    using G = TileGeom<STRIDE, TW, TH>;
    constexpr int CIN = CA + CB;
    constexpr int NT = G::GROUP * (COB / COT);
    static_assert(CA % CK == 0 && CB % CK == 0, "channel groups must be chunk aligned");
    static_assert(COB % COT == 0 && COUT % COB == 0, "bad output channel blocking");
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                                  // [CIN][9][COB]
    float* sIn = smem + CIN * 9 * COB;                 // [CK][IH][IP]

    const int tid = threadIdx.x;
    const int cog = tid / G::GROUP;                    // output-channel group inside the block
    const int t = tid - cog * G::GROUP;
    const int tx = t % (TW / PX), ty = t / (TW / PX);
    const int tiles_x = (a.wout + TW - 1) / TW;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int cob = blockIdx.y;                        // output-channel block
    const int b = blockIdx.z;
    const int ox0 = tile_x * TW, oy0 = tile_y * TH;
    const int ix0 = ox0 * STRIDE - 1, iy0 = oy0 * STRIDE - 1;

    for (int i = tid; i < CIN * 9 * COB; i += NT) {
        const int col = i % COB, ct = i / COB;
        sW[i] = __ldg(a.wpk + (size_t)ct * COUT + cob * COB + col);
    }

    float acc[PY][PX][COT];
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int c = 0; c < COT; ++c) acc[j][p][c] = 0.f;

    for (int chunk = 0; chunk < CIN / CK; ++chunk) {
        const int ci0 = chunk * CK;
        const float* base;
        long long cstride;
        if (ci0 < CA) { base = a.inA + (long long)b * a.strideA_b + (long long)ci0 * a.strideA_c; cstride = a.strideA_c; }
        else { base = a.inB + (long long)b * a.strideB_b + (long long)(ci0 - CA) * a.strideB_c; cstride = a.strideB_c; }
        __syncthreads();                               // previous chunk fully consumed (and sW visible)
        for (int i = tid; i < CK * G::IH * G::IW; i += NT) {
            const int col = i % G::IW, rc = i / G::IW;
            const int row = rc % G::IH, c = rc / G::IH;
            const int gy = iy0 + row, gx = ix0 + col;
            float v = 0.f;
            if (gy >= 0 && gy < a.hin && gx >= 0 && gx < a.win) v = __ldg(base + c * cstride + (long long)gy * a.win + gx);
            sIn[(c * G::IH + row) * G::IP + col] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < CK; ++c) {
            const float* wrow = sW + ((ci0 + c) * 9) * COB + cog * COT;
            const float* irow = sIn + (c * G::IH) * G::IP;
            if (STRIDE == 1) {
#pragma unroll
                for (int r = 0; r < PY + 2; ++r) {
                    const float* ip = irow + (PY * ty + r) * G::IP + PX * tx;
                    const float4 v0 = *reinterpret_cast<const float4*>(ip);
                    const float2 v1 = *reinterpret_cast<const float2*>(ip + 4);
                    const float in[6] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y};
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int j = r - ky;
                        if (j < 0 || j >= PY) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float4 w0 = *reinterpret_cast<const float4*>(wrow + (ky * 3 + kx) * COB);
                            const float4 w1 = *reinterpret_cast<const float4*>(wrow + (ky * 3 + kx) * COB + 4);
                            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                            for (int p = 0; p < PX; ++p)
#pragma unroll
                                for (int co = 0; co < COT; ++co) acc[j][p][co] = fmaf(in[p + kx], wv[co], acc[j][p][co]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 2 * PY + 1; ++r) {
                    const float* ip = irow + (2 * PY * ty + r) * G::IP + 2 * PX * tx;
                    float in[2 * PX + 1];
#pragma unroll
                    for (int q = 0; q < 2 * PX + 1; ++q) in[q] = ip[q];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int jj = r - ky;
                        if (jj < 0 || (jj & 1) || jj / 2 >= PY) continue;
                        const int j = jj / 2;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float4 w0 = *reinterpret_cast<const float4*>(wrow + (ky * 3 + kx) * COB);
                            const float4 w1 = *reinterpret_cast<const float4*>(wrow + (ky * 3 + kx) * COB + 4);
                            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                            for (int p = 0; p < PX; ++p)
#pragma unroll
                                for (int co = 0; co < COT; ++co) acc[j][p][co] = fmaf(in[2 * p + kx], wv[co], acc[j][p][co]);
                        }
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------------ epilogue
    const int co_base = cob * COB + cog * COT;          // first global output channel of this thread
    const size_t plane = (size_t)a.hout * a.wout;
#pragma unroll
    for (int j = 0; j < PY; ++j) {
        const int oy = oy0 + PY * ty + j;
        if (oy >= a.hout) continue;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int ox = ox0 + PX * tx + p;
            if (ox >= a.wout) continue;
            const size_t pix = (size_t)oy * a.wout + ox;
#pragma unroll
            for (int c = 0; c < COT; ++c) {
                const int co = co_base + c;
                float v = acc[j][p][c];
                if (EPI == EPI_RELU) {
                    a.out0[((size_t)b * COUT + co) * plane + pix] = fmaxf(v, 0.f);
                } else if (EPI == EPI_GATES) {
                    constexpr int HC = COUT / 2;
                    v = sigmoid_f(v + __ldg(a.bias + co));
                    if (co < HC) {                      // reset gate -> r*h
                        const size_t o = ((size_t)b * HC + co) * plane + pix;
                        a.out0[o] = v * a.hstate[o];
                    } else {                            // update gate
                        a.out1[((size_t)b * HC + (co - HC)) * plane + pix] = v;
                    }
                } else {
                    const size_t o = ((size_t)b * COUT + co) * plane + pix;
                    const float cand = tanhf(v + __ldg(a.bias + co));
                    const float u = a.ugate[o];
                    a.out0[o] = u * a.hstate[o] + (1.f - u) * cand;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 7: y = relu(convT3x3 s2 p1 op1 (h2; 16->8) + b + h1).  One thread per half-resolution pixel: it owns
// the 2x2 full-resolution outputs that (iy,ix) is the top-left contributor of.
//   out(2iy  ,2ix  ) = in(iy,ix) W11
//   out(2iy  ,2ix+1) = in(iy,ix+1) W10 + in(iy,ix) W12
//   out(2iy+1,2ix  ) = in(iy+1,ix) W01 + in(iy,ix) W21
//   out(2iy+1,2ix+1) = in(iy+1,ix+1) W00 + in(iy+1,ix) W02 + in(iy,ix+1) W20 + in(iy,ix) W22
// (PyTorch ConvTranspose2d scatter form, SURVEY.md Appendix B.)
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(128)
upconv_add_relu_kernel(const float* __restrict__ in, const float* __restrict__ wpk, const float* __restrict__ bias,
                       const float* __restrict__ skip, float* __restrict__ out, int hin, int win) {
    __shared__ float sW[CIN * 9 * COUT];
    for (int i = threadIdx.x; i < CIN * 9 * COUT; i += blockDim.x) sW[i] = __ldg(wpk + i);
    __syncthreads();
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y;
    const int b = blockIdx.z;
    if (ix >= win) return;
    const size_t ip = (size_t)hin * win;
    const int wout = 2 * win;
    const size_t op = (size_t)4 * ip;
    const bool hx = ix + 1 < win, hy = iy + 1 < hin;
    float acc[4][COUT];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[q][c] = 0.f;
    const float* pin = in + (size_t)b * CIN * ip + (size_t)iy * win + ix;
#pragma unroll 2
    for (int ci = 0; ci < CIN; ++ci) {
        const float* p = pin + (size_t)ci * ip;
        const float v00 = __ldg(p);
        const float v01 = hx ? __ldg(p + 1) : 0.f;
        const float v10 = hy ? __ldg(p + win) : 0.f;
        const float v11 = (hx && hy) ? __ldg(p + win + 1) : 0.f;
        const float* w = sW + ci * 9 * COUT;
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            acc[0][c] += v00 * w[4 * COUT + c];
            acc[1][c] += v01 * w[3 * COUT + c] + v00 * w[5 * COUT + c];
            acc[2][c] += v10 * w[1 * COUT + c] + v00 * w[7 * COUT + c];
            acc[3][c] += v11 * w[0 * COUT + c] + v10 * w[2 * COUT + c] + v01 * w[6 * COUT + c] + v00 * w[8 * COUT + c];
        }
    }
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
        const float bc = __ldg(bias + c);
        const size_t o = ((size_t)b * COUT + c) * op + (size_t)(2 * iy) * wout + 2 * ix;
        const float2 s0 = *reinterpret_cast<const float2*>(skip + o);
        const float2 s1 = *reinterpret_cast<const float2*>(skip + o + wout);
        float2 r0, r1;
        r0.x = fmaxf(acc[0][c] + bc + s0.x, 0.f); r0.y = fmaxf(acc[1][c] + bc + s0.y, 0.f);
        r1.x = fmaxf(acc[2][c] + bc + s1.x, 0.f); r1.y = fmaxf(acc[3][c] + bc + s1.y, 0.f);
        *reinterpret_cast<float2*>(out + o) = r0;
        *reinterpret_cast<float2*>(out + o + wout) = r1;
    }
}

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

// stage 3: logit = conv3x3(y; 8->1) + b at the same resolution
__global__ void __launch_bounds__(128)
out_conv_regress_kernel(const float* __restrict__ y, OutWeights ow, HypSpec hs, int prob_mode, RegressState st,
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
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int gy = yy + ky - 1;
            if (gy < 0 || gy >= h) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int gx = x + kx - 1;
                if (gx < 0 || gx >= w) continue;
                acc = fmaf(__ldg(p + (size_t)gy * w + gx), sw[ci * 9 + ky * 3 + kx], acc);
            }
        }
    }
    const int pix = yy * w + x;
    const size_t o = (size_t)b * hw + pix;
    if (logits_out) logits_out[((size_t)b * D + k) * hw + pix] = acc;
    const HypLine line = hyp_line(hs, b, pix, (int)hw, D);
    regress_update(st, o, acc, hyp_at(line, k), k, D, prob_mode, depth, conf);
}

// stages 1-2: logit = convT3x3 s2 (y; 8->1) + b at twice the resolution; the hypothesis of an output
// pixel is the align_corners=False bilinear upsample of the plane's hypotheses (module.py:622).
__global__ void __launch_bounds__(128)
out_upconv_regress_kernel(const float* __restrict__ y, OutWeights ow, HypSpec hs, int prob_mode, RegressState st,
                          float* __restrict__ depth, float* __restrict__ conf, float* __restrict__ logits_out,
                          int k, int D, int h, int w) {
    __shared__ float sw[73];
    stage_out_weights(ow, sw);
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y;
    const int b = blockIdx.z;
    if (ix >= w) return;
    const size_t hw = (size_t)h * w;
    const bool hx = ix + 1 < w, hy = iy + 1 < h;
    float l00 = sw[72], l01 = sw[72], l10 = sw[72], l11 = sw[72];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
        const float* p = y + ((size_t)b * 8 + ci) * hw + (size_t)iy * w + ix;
        const float v00 = __ldg(p);
        const float v01 = hx ? __ldg(p + 1) : 0.f;
        const float v10 = hy ? __ldg(p + w) : 0.f;
        const float v11 = (hx && hy) ? __ldg(p + w + 1) : 0.f;
        const float* wt = sw + ci * 9;
        l00 += v00 * wt[4];
        l01 += v01 * wt[3] + v00 * wt[5];
        l10 += v10 * wt[1] + v00 * wt[7];
        l11 += v11 * wt[0] + v10 * wt[2] + v01 * wt[6] + v00 * wt[8];
    }
    const int Ho = 2 * h, Wo = 2 * w;
    const size_t ohw = (size_t)Ho * Wo;
    const float lg[4] = {l00, l01, l10, l11};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int oy = 2 * iy + (q >> 1), ox = 2 * ix + (q & 1);
        float dval;
        if (hs.mode == ADAMVS_HYP_PLANES) {
            dval = hyp_at(hyp_line(hs, b, 0, (int)hw, D), k);
        } else {
            const Lerp ly = lerp_index(oy, 0.5f, h), lx = lerp_index(ox, 0.5f, w);
            const float d00 = hyp_at(hyp_line(hs, b, ly.i0 * w + lx.i0, (int)hw, D), k);
            const float d01 = hyp_at(hyp_line(hs, b, ly.i0 * w + lx.i1, (int)hw, D), k);
            const float d10 = hyp_at(hyp_line(hs, b, ly.i1 * w + lx.i0, (int)hw, D), k);
            const float d11 = hyp_at(hyp_line(hs, b, ly.i1 * w + lx.i1, (int)hw, D), k);
            dval = ly.l0 * (lx.l0 * d00 + lx.l1 * d01) + ly.l1 * (lx.l0 * d10 + lx.l1 * d11);
        }
        const size_t o = (size_t)b * ohw + (size_t)oy * Wo + ox;
        if (logits_out) logits_out[((size_t)b * D + k) * ohw + (size_t)oy * Wo + ox] = lg[q];
        regress_update(st, o, lg[q], dval, k, D, prob_mode, depth, conf);
    }
}

// ------------------------------------------------------------------------------------------------
// weight packing: reference layouts -> [ci][tap][co]
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_kernel(const float* __restrict__ w, float* __restrict__ pk, int cout, int cin, int transposed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * cin * 9) return;
    const int co = i % cout, t = (i / cout) % 9, ci = i / (cout * 9);
    pk[i] = transposed ? w[((size_t)ci * cout + co) * 9 + t] : w[((size_t)co * cin + ci) * 9 + t];
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI, int TW, int TH>
static cudaError_t launch_conv(const ConvArgs& a, int B, cudaStream_t st) {
    using G = TileGeom<STRIDE, TW, TH>;
    constexpr int NT = G::GROUP * (COB / COT);
    constexpr size_t smem = sizeof(float) * ((CA + CB) * 9 * COB + CK * G::IH * G::IP);
    auto kern = conv3x3_kernel<CA, CB, COUT, COB, STRIDE, EPI, TW, TH>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int tiles = ((a.wout + TW - 1) / TW) * ((a.hout + TH - 1) / TH);
    dim3 grid(tiles, COUT / COB, B);
    kern<<<grid, NT, smem, st>>>(a);
    return cudaGetLastError();
}

// Tile choice: the largest tile that still gives every SM at least ~2 blocks.
static int pick_tile(int hout, int wout, int B, int co_blocks) {
    const long long want = 2LL * 148;
    auto blocks = [&](int tw, int th) { return (long long)((wout + tw - 1) / tw) * ((hout + th - 1) / th) * B * co_blocks; };
    if (blocks(32, 32) >= want) return 0;
    if (blocks(32, 16) >= want) return 1;
    return 2;
}

template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI>
static cudaError_t launch_conv_auto(const ConvArgs& a, int B, cudaStream_t st) {
    switch (pick_tile(a.hout, a.wout, B, COUT / COB)) {
        case 0: return launch_conv<CA, CB, COUT, COB, STRIDE, EPI, 32, 32>(a, B, st);
        case 1: return launch_conv<CA, CB, COUT, COB, STRIDE, EPI, 32, 16>(a, B, st);
        default: return launch_conv<CA, CB, COUT, COB, STRIDE, EPI, 16, 16>(a, B, st);
    }
}

struct Workspace {
    float *pk_conv1, *pk_gates1, *pk_cand1, *pk_conv2, *pk_gates2, *pk_cand2, *pk_up1;
    float *x1, *h1, *rh1, *u1, *x2, *h2, *rh2, *u2, *y, *s0, *s1, *s2;
    size_t total;
};

static Workspace carve(float* base, int B, int C, int h, int w, int out_up) {
    Workspace ws;
    size_t off = 0;
    auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return p; };
    const size_t hw = (size_t)h * w, hw2 = (size_t)(h / 2) * (w / 2), ohw = out_up ? 4 * hw : hw;
    ws.pk_conv1 = take((size_t)C * 9 * 8);
    ws.pk_gates1 = take(16 * 9 * 16);
    ws.pk_cand1 = take(16 * 9 * 8);
    ws.pk_conv2 = take(8 * 9 * 16);
    ws.pk_gates2 = take(32 * 9 * 32);
    ws.pk_cand2 = take(32 * 9 * 16);
    ws.pk_up1 = take(16 * 9 * 8);
    ws.x1 = take(B * 8 * hw);  ws.h1 = take(B * 8 * hw);  ws.rh1 = take(B * 8 * hw);  ws.u1 = take(B * 8 * hw);
    ws.x2 = take(B * 16 * hw2); ws.h2 = take(B * 16 * hw2); ws.rh2 = take(B * 16 * hw2); ws.u2 = take(B * 16 * hw2);
    ws.y = take(B * 8 * hw);
    ws.s0 = take(B * ohw); ws.s1 = take(B * ohw); ws.s2 = take(B * ohw);
    ws.total = off;
    return ws;
}

template <int C>
static cudaError_t run_conv1(const ConvArgs& a, int B, cudaStream_t st) {
    return launch_conv_auto<C, 0, 8, 8, 1, EPI_RELU>(a, B, st);
}

}  // namespace adamvs

using namespace adamvs;

extern "C" size_t adamvs_regnet_red_workspace_floats(int B, int C, int D, int h, int w, int out_up) {
    (void)D;
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return 0;
    return carve(nullptr, B, C, h, w, out_up).total;
}

#define ADAMVS_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

extern "C" int adamvs_regnet_red_f32(const float* volume, const adamvs_regnet_weights* hwts,
                                     int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                     int out_up, int prob_mode,
                                     float* workspace, size_t workspace_floats,
                                     float* depth, float* conf, float* logits_out,
                                     int B, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(volume && hwts && workspace && depth && conf && hyp_src);
    ADAMVS_CHECK_ARG(B > 0 && B <= 65535 && D >= 2 && h > 0 && w > 0 && (h % 2) == 0 && (w % 2) == 0 && h <= 65535);
    ADAMVS_CHECK_ARG(C == 8 || C == 16 || C == 32);
    ADAMVS_CHECK_ARG(prob_mode == ADAMVS_PROB_SOFTMAX || prob_mode == ADAMVS_PROB_EXP_EPS);
    ADAMVS_CHECK_ARG(hyp_mode == ADAMVS_HYP_PLANES ? hyp_ncol >= 2 : (hyp_mode == ADAMVS_HYP_PER_PIXEL && half_range));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws = carve(workspace, B, C, h, w, out_up);
    if (ws.total > workspace_floats) return ADAMVS_ENOSPACE;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    const int h2 = h / 2, w2 = w / 2;
    const size_t hw = (size_t)h * w, hw2 = (size_t)h2 * w2;

    // one-off per call: weights into [ci][tap][co]; states to zero (adamvs.py:175-176 / 448-449)
    auto pack = [&](const float* src, float* dst, int cout, int cin, int tr) {
        const int n = cout * cin * 9;
        pack_conv_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, cout, cin, tr);
    };
    pack(hwts->conv1_w, ws.pk_conv1, 8, C, 0);
    pack(hwts->gates1_w, ws.pk_gates1, 16, 16, 0);
    pack(hwts->cand1_w, ws.pk_cand1, 8, 16, 0);
    pack(hwts->conv2_w, ws.pk_conv2, 16, 8, 0);
    pack(hwts->gates2_w, ws.pk_gates2, 32, 32, 0);
    pack(hwts->cand2_w, ws.pk_cand2, 16, 32, 0);
    pack(hwts->up1_w, ws.pk_up1, 8, 16, 1);
    ADAMVS_TRY(cudaGetLastError());
    ADAMVS_TRY(cudaMemsetAsync(ws.h1, 0, sizeof(float) * B * 8 * hw, st));
    ADAMVS_TRY(cudaMemsetAsync(ws.h2, 0, sizeof(float) * B * 16 * hw2, st));
    const OutWeights ow{hwts->out_w, hwts->out_b};
    const RegressState rs{ws.s0, ws.s1, ws.s2};

    for (int k = 0; k < D; ++k) {
        ConvArgs a{};
        // 1 conv1: plane k of the volume, channel stride D*h*w
        a.inA = volume + (size_t)k * hw; a.strideA_c = (long long)D * hw; a.strideA_b = (long long)C * D * hw;
        a.wpk = ws.pk_conv1; a.out0 = ws.x1; a.hin = h; a.win = w; a.hout = h; a.wout = w;
        if (C == 8) ADAMVS_TRY(run_conv1<8>(a, B, st));
        else if (C == 16) ADAMVS_TRY(run_conv1<16>(a, B, st));
        else ADAMVS_TRY(run_conv1<32>(a, B, st));
        // 2 gates1
        a = ConvArgs{};
        a.inA = ws.x1; a.strideA_c = hw; a.strideA_b = 8 * hw; a.inB = ws.h1; a.strideB_c = hw; a.strideB_b = 8 * hw;
        a.wpk = ws.pk_gates1; a.bias = hwts->gates1_b; a.out0 = ws.rh1; a.out1 = ws.u1; a.hstate = ws.h1;
        a.hin = h; a.win = w; a.hout = h; a.wout = w;
        ADAMVS_TRY((launch_conv_auto<8, 8, 16, 16, 1, EPI_GATES>(a, B, st)));
        // 3 cand1 (+ state update in place)
        a.inB = ws.rh1; a.wpk = ws.pk_cand1; a.bias = hwts->cand1_b; a.out0 = ws.h1; a.out1 = nullptr; a.ugate = ws.u1;
        ADAMVS_TRY((launch_conv_auto<8, 8, 8, 8, 1, EPI_CAND>(a, B, st)));
        // 4 conv2 (stride 2)
        a = ConvArgs{};
        a.inA = ws.h1; a.strideA_c = hw; a.strideA_b = 8 * hw; a.wpk = ws.pk_conv2; a.out0 = ws.x2;
        a.hin = h; a.win = w; a.hout = h2; a.wout = w2;
        ADAMVS_TRY((launch_conv_auto<8, 0, 16, 16, 2, EPI_RELU>(a, B, st)));
        // 5 gates2
        a = ConvArgs{};
        a.inA = ws.x2; a.strideA_c = hw2; a.strideA_b = 16 * hw2; a.inB = ws.h2; a.strideB_c = hw2; a.strideB_b = 16 * hw2;
        a.wpk = ws.pk_gates2; a.bias = hwts->gates2_b; a.out0 = ws.rh2; a.out1 = ws.u2; a.hstate = ws.h2;
        a.hin = h2; a.win = w2; a.hout = h2; a.wout = w2;
        ADAMVS_TRY((launch_conv_auto<16, 16, 32, 16, 1, EPI_GATES>(a, B, st)));
        // 6 cand2
        a.inB = ws.rh2; a.wpk = ws.pk_cand2; a.bias = hwts->cand2_b; a.out0 = ws.h2; a.out1 = nullptr; a.ugate = ws.u2;
        ADAMVS_TRY((launch_conv_auto<16, 16, 16, 16, 1, EPI_CAND>(a, B, st)));
        // 7 up1 + skip + relu
        {
            dim3 grid((w2 + 127) / 128, h2, B);
            upconv_add_relu_kernel<16, 8><<<grid, 128, 0, st>>>(ws.h2, ws.pk_up1, hwts->up1_b, ws.h1, ws.y, h2, w2);
        }
        // 8 output layer + regression
        {
            dim3 grid((w + 127) / 128, h, B);
            if (out_up) out_upconv_regress_kernel<<<grid, 128, 0, st>>>(ws.y, ow, hs, prob_mode, rs, depth, conf, logits_out, k, D, h, w);
            else out_conv_regress_kernel<<<grid, 128, 0, st>>>(ws.y, ow, hs, prob_mode, rs, depth, conf, logits_out, k, D, h, w);
        }
        ADAMVS_TRY(cudaGetLastError());
    }
    return 0;
}
