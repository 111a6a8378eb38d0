// adamvs_conv3x3_f32 — the 3x3 convolutions of the two 2-D networks around the cost-volume path (FeatureNet0,
// reference models/adamvs.py:57-104, and the stage-1 pair U-Net CostRegNet2D, :198-238), run on the same
// persistent TMA-fed FFMA kernels as the recurrent regulariser instead of cuDNN's fp32 paths: eval-mode
// BatchNorm is folded into weights and bias by the caller, ReLU is fused, and a channel concatenation in front
// of the conv (DeConv2dFuse, module.py:506-524) is read as two tensors instead of being materialised.
#include "conv3x3.cuh"
#include "conv3x3_tc.cuh"
#include <string.h>

namespace adamvs {

template <int CA, int CB, int COUT, int STRIDE>
static int run_conv(const ConvArgs& a, int N, cudaStream_t st, bool tma_only = false) {
    constexpr int COB = COUT % 16 == 0 ? 16 : 8;
    using L = ConvLayer<CA, CB, COUT, COB, STRIDE, EPI_BIAS>;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.inA) | reinterpret_cast<uintptr_t>(a.inB) |
                           reinterpret_cast<uintptr_t>(a.out0)) % 16) == 0;
    // Stride-1 layers with enough pixels for ~2 tiles per SM run on the tensor cores (conv3x3_tc.cuh: kind::tf32 with the
    // exact hi/lo operand split, fp32 accuracy): 4-5x the FFMA kernels on the 16- and 32-channel layers.
    // ADAMVS_CONV2D_MATH=ffma (or a forced ADAMVS_CONV_CFG) keeps the FFMA kernels - test / measurement hook.
    static const bool tc_allowed = [] {
        const char* e = getenv("ADAMVS_CONV2D_MATH");
        return !(e && !strcmp(e, "ffma")) && getenv("ADAMVS_CONV_CFG") == nullptr;
    }();
    if constexpr (STRIDE == 1 && COUT == 48) {                         // N = 3 x 2 x 48 exceeds one MMA: two 24-channel slices
        if (tc_allowed && aligned && a.win % 4 == 0 && (long long)N * a.hout * a.wout >= 30000) {
            using T = TcLayer<CA, CB, 24, EPI_BIAS>;
            ConvArgs h = a;
            h.wpk_cout = 48; h.out_cout = 48;
            ConvPlan p{};
            if (T::plan(p, h, N, 1)) {
                for (int half = 0; half < 2; ++half) {
                    p.args.co_off = 24 * half;
                    cudaError_t e = T::launch(p, N, PREC_FP32X3, st);
                    if (e != cudaSuccess) return (int)e;
                }
                return 0;
            }
        }
    }
    if constexpr (STRIDE == 1 && COUT <= 32) {
        if (tc_allowed && aligned && a.win % 4 == 0 && (long long)N * a.hout * a.wout >= 30000) {
            using T = TcLayer<CA, CB, COUT, EPI_BIAS>;
            ConvPlan p{};
            if (T::plan(p, a, N, 1)) {
                cudaError_t e = T::launch(p, N, PREC_FP32X3, st);
                return e == cudaSuccess ? 0 : (int)e;
            }
        }
    }
    if (aligned && a.win % 4 == 0 && a.wout % 4 == 0) {
        ConvPlan p{};
        if (L::plan(p, a, N, 1)) {
            cudaError_t e = L::launch(p, N, st);
            return e == cudaSuccess ? 0 : (int)e;
        }
    }
    if (tma_only) return ADAMVS_EINVAL;                // the pointer-arithmetic kernel would read past a narrower tensor
    using G = TileGeom<STRIDE, 16, 16>;
    constexpr size_t smem = sizeof(float) * ((CA + CB) * 9 * COB + CK * G::IH * G::IP);
    auto kern = conv3x3_kernel<CA, CB, COUT, COB, STRIDE, EPI_BIAS, 16, 16>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(((a.wout + 15) / 16) * ((a.hout + 15) / 16), COUT / COB, N);
    kern<<<grid, G::GROUP * (COB / COT), smem, st>>>(a);
    ADAMVS_LAUNCH_RESULT();
}

// ------------------------------------------------------------------------------------------------------------------
// adamvs_context_head_f32 - FeatureNet0's three output heads (reference models/adamvs.py:112-149):
//     out = conv1x1(cat(upsample(a), upsample(c), x))          a, c: pooled-context maps at 1/4 and 1/8 of x's size
// in one pass over x: the reference materialises both bilinear upsamplings and the concatenation (4.5x the bytes of x)
// before a cuDNN 1x1 convolution; here a thread owns PX x-adjacent pixels and all COUT outputs, interpolates the
// CCTX + CCTX context channels from the (L1/L2-resident) small maps in registers and reads x once.
// Bilinear weights as ATen's upsample_bilinear2d, align_corners = False.
// ------------------------------------------------------------------------------------------------------------------
template <int CX, int CCTX, int COUT, int PX>
__global__ void __launch_bounds__(128)
context_head_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ c,
                    const float* __restrict__ wgt, float* __restrict__ out, int h, int w, int ha, int wa, int hc, int wc) {
    constexpr int CIN = 2 * CCTX + CX;
    __shared__ float sW[CIN * COUT];                                   // [ci][co]
    for (int i = threadIdx.x; i < CIN * COUT; i += 128) sW[i] = __ldg(wgt + (size_t)(i % COUT) * CIN + i / COUT);
    __syncthreads();
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * PX, y = blockIdx.y, n = blockIdx.z;
    if (x0 >= w) return;                                               // w % PX == 0: a group is all in or all out
    float acc[PX][COUT];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[p][co] = 0.f;
    // context channels: cat order is (a, c, x)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const float* src = m == 0 ? a : c;
        const int hs = m == 0 ? ha : hc, ws = m == 0 ? wa : wc;
        const Lerp ly = lerp_index(y, (float)hs / (float)h, hs);
        const float* base = src + (size_t)n * CCTX * hs * ws;
        Lerp lx[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) lx[p] = lerp_index(x0 + p, (float)ws / (float)w, ws);
        const int cmin = lx[0].i0;
        if (lx[PX - 1].i1 - cmin <= 2) {
            // the PX pixels of a thread sample at most three source columns (maps at 1/4 and 1/8 of x's width: always):
            // load those once per channel and row - 6 gathers instead of 4 * PX - and give every pixel three column
            // weights (its two lerp weights on its two columns, zero on the third)
            float wcol[PX][3];
#pragma unroll
            for (int p = 0; p < PX; ++p)
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    wcol[p][d] = (lx[p].i0 - cmin == d ? lx[p].l0 : 0.f) + (lx[p].i1 - cmin == d ? lx[p].l1 : 0.f);
            const int c1 = min(cmin + 1, ws - 1), c2 = min(cmin + 2, ws - 1);
#pragma unroll
            for (int j = 0; j < CCTX; ++j) {
                const float* r0 = base + (size_t)j * hs * ws + ly.i0 * ws;
                const float* r1 = base + (size_t)j * hs * ws + ly.i1 * ws;
                const float a0 = __ldg(r0 + cmin), a1 = __ldg(r0 + c1), a2 = __ldg(r0 + c2);
                const float b0 = __ldg(r1 + cmin), b1 = __ldg(r1 + c1), b2 = __ldg(r1 + c2);
                const float* wr = sW + (m * CCTX + j) * COUT;
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const float t0 = fmaf(wcol[p][2], a2, fmaf(wcol[p][1], a1, wcol[p][0] * a0));
                    const float t1 = fmaf(wcol[p][2], b2, fmaf(wcol[p][1], b1, wcol[p][0] * b0));
                    const float v = ly.l0 * t0 + ly.l1 * t1;
#pragma unroll
                    for (int co = 0; co < COUT; ++co) acc[p][co] = fmaf(v, wr[co], acc[p][co]);
                }
            }
        } else {
#pragma unroll
            for (int p = 0; p < PX; ++p) {
#pragma unroll
                for (int j = 0; j < CCTX; ++j) {
                    const float* pl = base + (size_t)j * hs * ws;
                    const float v00 = __ldg(pl + ly.i0 * ws + lx[p].i0), v01 = __ldg(pl + ly.i0 * ws + lx[p].i1);
                    const float v10 = __ldg(pl + ly.i1 * ws + lx[p].i0), v11 = __ldg(pl + ly.i1 * ws + lx[p].i1);
                    const float v = ly.l0 * (lx[p].l0 * v00 + lx[p].l1 * v01) + ly.l1 * (lx[p].l0 * v10 + lx[p].l1 * v11);
                    const float* wr = sW + (m * CCTX + j) * COUT;
#pragma unroll
                    for (int co = 0; co < COUT; ++co) acc[p][co] = fmaf(v, wr[co], acc[p][co]);
                }
            }
        }
    }
    const size_t hw = (size_t)h * w;
    const float* px = x + (size_t)n * CX * hw + (size_t)y * w + x0;
#pragma unroll 4
    for (int ci = 0; ci < CX; ++ci) {
        float v[PX];
        if (PX == 4) { const float4 t = *reinterpret_cast<const float4*>(px + (size_t)ci * hw); v[0] = t.x; v[1] = t.y; v[2 % PX] = t.z; v[3 % PX] = t.w; }
        else { const float2 t = *reinterpret_cast<const float2*>(px + (size_t)ci * hw); v[0] = t.x; v[1] = t.y; }
        const float* wr = sW + (2 * CCTX + ci) * COUT;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            const float wv = wr[co];
#pragma unroll
            for (int p = 0; p < PX; ++p) acc[p][co] = fmaf(v[p], wv, acc[p][co]);
        }
    }
    float* po = out + (size_t)n * COUT * hw + (size_t)y * w + x0;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        if (PX == 4) *reinterpret_cast<float4*>(po + (size_t)co * hw) = make_float4(acc[0][co], acc[1][co], acc[2 % PX][co], acc[3 % PX][co]);
        else *reinterpret_cast<float2*>(po + (size_t)co * hw) = make_float2(acc[0][co], acc[1][co]);
    }
}

template <int CX, int CCTX, int COUT, int PX>
static int launch_context_head(const float* x, const float* a, const float* c, const float* wgt, float* out,
                               int N, int h, int w, int ha, int wa, int hc, int wc, cudaStream_t st) {
    dim3 grid((w / PX + 127) / 128, h, N);
    context_head_kernel<CX, CCTX, COUT, PX><<<grid, 128, 0, st>>>(x, a, c, wgt, out, h, w, ha, wa, hc, wc);
    ADAMVS_LAUNCH_RESULT();
}

}  // namespace adamvs

using namespace adamvs;

extern "C" int adamvs_conv3x3_supported(int CA, int CB, int COUT, int stride) {
    if (stride == 1 && CA == 3 && CB == 0 && COUT == 8) return 1;     // the 3-channel image: one 8-channel chunk, see adamvs_conv3x3_f32
    if (stride == 1) {
        if (CB == 0) return (CA == 8 && COUT == 8) || (CA == 16 && COUT == 16) || (CA == 32 && COUT == 32) || (CA == 48 && COUT == 48) ||
                            (CA == 32 && COUT == 16) || (CA == 64 && COUT == 32);     // the 5x5 stride-2 convs in polyphase form
        return (CA == 16 && CB == 16 && COUT == 16) || (CA == 8 && CB == 8 && COUT == 8);
    }
    return stride == 2 && CB == 0 && CA == 48 && COUT == 48;
}

extern "C" int adamvs_conv3x3_f32(const float* inA, int CA, const float* inB, int CB, const float* wpk, const float* bias,
                                  int relu, int stride, float* out, int N, int COUT, int hin, int win, void* stream) {
    ADAMVS_CHECK_ARG(inA && wpk && bias && out && N > 0 && hin > 0 && win > 0 && (CB == 0 || inB));
    ADAMVS_CHECK_ARG(adamvs_conv3x3_supported(CA, CB, COUT, stride));
    ADAMVS_CHECK_ARG(stride == 1 || (hin % 2 == 0 && win % 2 == 0));
    const int hout = hin / stride, wout = win / stride;
    const size_t hw = (size_t)hin * win;
    ConvArgs a{};
    a.inA = inA; a.strideA_c = (long long)hw; a.strideA_b = (long long)CA * hw; a.planesA = CA;
    a.inB = inB; a.strideB_c = (long long)hw; a.strideB_b = (long long)CB * hw; a.planesB = CB;
    a.wpk = wpk; a.bias = bias; a.out0 = out; a.relu = relu;
    a.hin = hin; a.win = win; a.hout = hout; a.wout = wout;
    cudaStream_t st = (cudaStream_t)stream;
    if (CA == 3) {
        // FeatureNet0's first layer reads the 3-channel image in place (no zero-padded copy): the TMA box of the one
        // 8-channel chunk starts at plane 3n; its planes 3..7 are the next image's (finite) channels - or the tensor map's
        // zero fill behind the last image - and meet zero weights (wpk rows 3..7 are zero by contract).
        ADAMVS_CHECK_ARG(win % 4 == 0 && reinterpret_cast<uintptr_t>(inA) % 16 == 0);
        a.strideA_b = 3LL * (long long)hw; a.planesA = 3;
        return run_conv<8, 0, 8, 1>(a, N, st, true);
    }
    if (stride == 2) return run_conv<48, 0, 48, 2>(a, N, st);
    if (CB == 16) return run_conv<16, 16, 16, 1>(a, N, st);
    if (CB == 8) return run_conv<8, 8, 8, 1>(a, N, st);
    switch (CA) {
        case 8: return run_conv<8, 0, 8, 1>(a, N, st);
        case 16: return run_conv<16, 0, 16, 1>(a, N, st);
        case 32: return COUT == 32 ? run_conv<32, 0, 32, 1>(a, N, st) : run_conv<32, 0, 16, 1>(a, N, st);
        case 64: return run_conv<64, 0, 32, 1>(a, N, st);
        default: return run_conv<48, 0, 48, 1>(a, N, st);
    }
}

// ---- y = act(convT3x3 s2 p1 op1 (x; CIN -> COUT) + bias): Deconv2d + folded BatchNorm + ReLU (module.py:202-245) and
// CostRegNet2D's up blocks (adamvs.py:212-225).  One thread per input pixel and block of 8 output channels; it owns the
// 2x2 outputs that (iy,ix) is the top-left contributor of (tap table in regnet.cu).  wpk is [ci][tap][COUT].
namespace adamvs {
template <int CIN, int COUT>
static __global__ void __launch_bounds__(128)
deconv3x3_kernel(const float* __restrict__ in, const float* __restrict__ wpk, const float* __restrict__ bias, int relu,
                 const float* __restrict__ residual, float* __restrict__ out, int hin, int win) {
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
    const bool hx = ix + 1 < win, hy = iy + 1 < hin;
    float acc[4][COB];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < COB; ++c) acc[q][c] = 0.f;
    const float* pin = in + (size_t)b * CIN * ip + (size_t)iy * win + ix;
#pragma unroll 4
    for (int ci = 0; ci < CIN; ++ci) {
        const float* p = pin + (size_t)ci * ip;
        const float v00 = __ldg(p);
        const float v01 = hx ? __ldg(p + 1) : 0.f;
        const float v10 = hy ? __ldg(p + win) : 0.f;
        const float v11 = (hx && hy) ? __ldg(p + win + 1) : 0.f;
        const float* w = sW + ci * 9 * COB;
#pragma unroll
        for (int c = 0; c < COB; ++c) {
            acc[0][c] = fmaf(v00, w[4 * COB + c], acc[0][c]);
            acc[1][c] = fmaf(v01, w[3 * COB + c], fmaf(v00, w[5 * COB + c], acc[1][c]));
            acc[2][c] = fmaf(v10, w[1 * COB + c], fmaf(v00, w[7 * COB + c], acc[2][c]));
            acc[3][c] = fmaf(v11, w[0 * COB + c], fmaf(v10, w[2 * COB + c], fmaf(v01, w[6 * COB + c], fmaf(v00, w[8 * COB + c], acc[3][c]))));
        }
    }
#pragma unroll
    for (int c = 0; c < COB; ++c) {
        const float bc = __ldg(bias + cob * COB + c);
        float r[4] = {acc[0][c] + bc, acc[1][c] + bc, acc[2][c] + bc, acc[3][c] + bc};
        if (relu) { r[0] = fmaxf(r[0], 0.f); r[1] = fmaxf(r[1], 0.f); r[2] = fmaxf(r[2], 0.f); r[3] = fmaxf(r[3], 0.f); }
        const size_t o = ((size_t)b * COUT + cob * COB + c) * 4 * ip + (size_t)(2 * iy) * wout + 2 * ix;
        if (residual) {                                  // skip connection added after the activation (adamvs.py:233-235)
            const float2 s0 = *reinterpret_cast<const float2*>(residual + o), s1 = *reinterpret_cast<const float2*>(residual + o + wout);
            r[0] += s0.x; r[1] += s0.y; r[2] += s1.x; r[3] += s1.y;
        }
        *reinterpret_cast<float2*>(out + o) = make_float2(r[0], r[1]);
        *reinterpret_cast<float2*>(out + o + wout) = make_float2(r[2], r[3]);
    }
}

// The same transposed convolution, TMA-fed and persistent (win % 4 == 0).  The kernel above re-reads its input from
// global memory for every block of 8 output channels, with nothing in flight while it computes (20-23 TFLOP/s).  Here a
// CTA owns one block of 8 output channels (its weights resident in shared memory) and walks over tiles of 32 x 8 input
// pixels; the input arrives as [8 channels][9 rows][36 columns] boxes (right / bottom neighbours included, image borders
// zero-filled) through a ring that stays three boxes ahead.  A thread owns two vertically adjacent input pixels (2 x 2x2
// outputs) and four channels: per input channel 6 input loads + 9 broadcast weight vectors feed 72 FFMA (the K3 tail's
// mapping, regnet.cu).  Same FFMA order per output as above: bit-identical results.
constexpr int kDcW = 32, kDcH = 8, kDcBoxW = 36, kDcBoxH = kDcH + 1, kDcStages = 4;
constexpr int kDcStageFloats = 8 * kDcBoxH * kDcBoxW;

template <int CIN>
constexpr size_t deconv_tma_smem() { return sizeof(float) * (kDcStages * kDcStageFloats + CIN * 72) + kDcStages * sizeof(uint64_t); }

template <int CIN, int COUT>
static __global__ void __launch_bounds__(256, 2)
deconv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const float* __restrict__ wpk, const float* __restrict__ bias, int relu,
                     const float* __restrict__ residual, float* __restrict__ out, int hin, int win, TileGrid tg) {
    constexpr int NCH = CIN / 8;
    static_assert((kDcStageFloats * 4) % 128 == 0, "TMA destinations are 128-byte aligned");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sIn = reinterpret_cast<float*>(smem_raw);                       // [kDcStages][8][9][36]
    float* sW = sIn + kDcStages * kDcStageFloats;                          // [CIN][9][8] of this CTA's channel block
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + CIN * 72);
    const int tid = threadIdx.x, cob = blockIdx.y;
    if (tid == 0) { for (int i = 0; i < kDcStages; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    for (int i = tid; i < CIN * 72; i += 256) sW[i] = __ldg(wpk + (size_t)(i / 8) * COUT + cob * 8 + (i % 8));
    // this thread: input rows 2*qi, 2*qi + 1 of the tile, column j, output channels cob*8 + c0 .. + 3
    const int half = tid >> 7, pr = tid & 127, qi = pr >> 5, j = pr & 31, c0 = 4 * half;
    float bc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) bc[c] = __ldg(bias + cob * 8 + c0 + c);
    __syncthreads();

    const int my_tiles = (tg.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * NCH;                                      // chunk stream of this CTA
    const int tiles_per_item = tg.tiles_x * tg.tiles_y;
    auto tile_origin = [&](int ti, int& b, int& ix0, int& iy0) {
        const int tile = blockIdx.x + ti * gridDim.x;
        b = tg.by_item.div(tile);
        const int r = tile - b * tiles_per_item, ty = tg.by_x.div(r);
        ix0 = (r - ty * tg.tiles_x) * kDcW; iy0 = ty * kDcH;
    };
    auto issue = [&](int g) {                                              // one thread
        const int ti = g / NCH, c = g - ti * NCH, s = g % kDcStages;
        int b, ix0, iy0;
        tile_origin(ti, b, ix0, iy0);
        fence_proxy_async();
        mbar_expect_tx(&bars[s], kDcStageFloats * 4);
        tma_load_4d(sIn + s * kDcStageFloats, &tmIn, &bars[s], ix0, iy0, 0, b * CIN + c * 8);
    };
    if (tid == 0)
        for (int g = 0; g < kDcStages - 1 && g < total; ++g) issue(g);

    const size_t ip = (size_t)hin * win;
    const int wout = 2 * win;
    int g = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        int b, ix0, iy0;
        tile_origin(ti, b, ix0, iy0);
        float acc[2][4][4];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[q][k][c] = 0.f;
        for (int ch = 0; ch < NCH; ++ch, ++g) {
            const int s = g % kDcStages;
            if (tid == 0 && g + kDcStages - 1 < total) issue(g + kDcStages - 1);   // its slot was consumed in iteration g - 1
            mbar_wait(&bars[s], (g / kDcStages) & 1);
            const float* st = sIn + s * kDcStageFloats + (2 * qi) * kDcBoxW + j;
#pragma unroll 4
            for (int cc = 0; cc < 8; ++cc) {
                const float* p = st + cc * kDcBoxH * kDcBoxW;
                float v[3][2];
#pragma unroll
                for (int r = 0; r < 3; ++r) { v[r][0] = p[r * kDcBoxW]; v[r][1] = p[r * kDcBoxW + 1]; }
                const float* w = sW + (ch * 8 + cc) * 72 + c0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float v00 = v[q][0], v01 = v[q][1], v10 = v[q + 1][0], v11 = v[q + 1][1];
                        acc[q][0][c] = fmaf(v00, w[4 * 8 + c], acc[q][0][c]);
                        acc[q][1][c] = fmaf(v01, w[3 * 8 + c], fmaf(v00, w[5 * 8 + c], acc[q][1][c]));
                        acc[q][2][c] = fmaf(v10, w[1 * 8 + c], fmaf(v00, w[7 * 8 + c], acc[q][2][c]));
                        acc[q][3][c] = fmaf(v11, w[0 * 8 + c], fmaf(v10, w[2 * 8 + c], fmaf(v01, w[6 * 8 + c], fmaf(v00, w[8 * 8 + c], acc[q][3][c]))));
                    }
            }
            __syncthreads();                                               // stage s consumed: the next issue may refill it
        }
        const int ix = ix0 + j;
        if (ix < win) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int iy = iy0 + 2 * qi + q;
                if (iy >= hin) break;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float r[4] = {acc[q][0][c] + bc[c], acc[q][1][c] + bc[c], acc[q][2][c] + bc[c], acc[q][3][c] + bc[c]};
                    if (relu) { r[0] = fmaxf(r[0], 0.f); r[1] = fmaxf(r[1], 0.f); r[2] = fmaxf(r[2], 0.f); r[3] = fmaxf(r[3], 0.f); }
                    const size_t o = ((size_t)b * COUT + cob * 8 + c0 + c) * 4 * ip + (size_t)(2 * iy) * wout + 2 * ix;
                    if (residual) {                      // skip connection added after the activation (adamvs.py:233-235)
                        const float2 s0 = *reinterpret_cast<const float2*>(residual + o), s1 = *reinterpret_cast<const float2*>(residual + o + wout);
                        r[0] += s0.x; r[1] += s0.y; r[2] += s1.x; r[3] += s1.y;
                    }
                    *reinterpret_cast<float2*>(out + o) = make_float2(r[0], r[1]);
                    *reinterpret_cast<float2*>(out + o + wout) = make_float2(r[2], r[3]);
                }
            }
        }
    }
}

template <int CIN, int COUT>
static int launch_deconv(const float* in, const float* wpk, const float* bias, int relu, const float* residual, float* out, int N, int hin, int win, cudaStream_t st) {
    // ADAMVS_DECONV_CFG=0 keeps the plain kernel (test / measurement hook, read once per process)
    static const bool tma_allowed = [] { const char* e = getenv("ADAMVS_DECONV_CFG"); return !(e && *e == '0'); }();
    if (tma_allowed && win % 4 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0 && hin < 65535 && win < 65535) {
        CUtensorMap tm;
        TileGrid tg{};
        tg.tiles_x = (win + kDcW - 1) / kDcW;
        tg.tiles_y = (hin + kDcH - 1) / kDcH;
        const long long nt = (long long)tg.tiles_x * tg.tiles_y * N;
        if (nt < (1 << 26) && make_tmap_4d(&tm, in, win, hin, 1, (long long)N * CIN, kDcBoxW, kDcBoxH, 8)) {
            tg.ntiles = (int)nt;
            tg.by_x = FastDiv(tg.tiles_x); tg.by_item = FastDiv(tg.tiles_x * tg.tiles_y);
            auto kern = deconv3x3_tma_kernel<CIN, COUT>;
            constexpr size_t smem = deconv_tma_smem<CIN>();
            static bool ready[64] = {false};
            int dev = 0;
            cudaGetDevice(&dev);
            dev = dev < 64 ? dev : 63;
            if (!ready[dev]) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return (int)e;
                ready[dev] = true;
            }
            constexpr int ncob = COUT / 8;
            int ctas = 2 * sm_count() / ncob;                // two CTAs per SM over all channel blocks
            if (ctas < 1) ctas = 1;
            if (ctas > tg.ntiles) ctas = tg.ntiles;
            kern<<<dim3(ctas, ncob), 256, smem, st>>>(tm, wpk, bias, relu, residual, out, hin, win, tg);
            ADAMVS_LAUNCH_RESULT();
        }
    }
    if ((long long)hin * (COUT / 8) > 65535 || N > 65535) return ADAMVS_EINVAL;
    dim3 grid((win + 127) / 128, hin * (COUT / 8), N);
    deconv3x3_kernel<CIN, COUT><<<grid, 128, 0, st>>>(in, wpk, bias, relu, residual, out, hin, win);
    ADAMVS_LAUNCH_RESULT();
}
}  // namespace adamvs

extern "C" int adamvs_deconv3x3_supported(int CIN, int COUT) {
    return (CIN == 32 && COUT == 16) || (CIN == 16 && COUT == 8) || (CIN == 48 && COUT == 48);
}

extern "C" int adamvs_deconv3x3_res_f32(const float* in, const float* wpk, const float* bias, int relu, const float* residual,
                                        float* out, int N, int CIN, int COUT, int hin, int win, void* stream) {
    ADAMVS_CHECK_ARG(in && wpk && bias && out && N > 0 && hin > 0 && win > 0 && adamvs_deconv3x3_supported(CIN, COUT));
    ADAMVS_CHECK_ARG((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(residual)) % 8 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (CIN == 32) return launch_deconv<32, 16>(in, wpk, bias, relu, residual, out, N, hin, win, st);
    if (CIN == 16) return launch_deconv<16, 8>(in, wpk, bias, relu, residual, out, N, hin, win, st);
    return launch_deconv<48, 48>(in, wpk, bias, relu, residual, out, N, hin, win, st);
}

extern "C" int adamvs_deconv3x3_f32(const float* in, const float* wpk, const float* bias, int relu, float* out,
                                    int N, int CIN, int COUT, int hin, int win, void* stream) {
    return adamvs_deconv3x3_res_f32(in, wpk, bias, relu, nullptr, out, N, CIN, COUT, hin, win, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// adamvs_context_pool_f32 - the two pooled-context branches in front of a FeatureNet0 output head (reference
// models/adamvs.py:112-147: AvgPool2d(4) / AvgPool2d(8) -> 1x1 conv -> BatchNorm -> ReLU), eval-mode BatchNorm folded:
//     a = relu(Wa * avgpool4(x) + ba)   [N,CO,h/4,w/4]        c = relu(Wc * avgpool8(x) + bc)   [N,CO,h/8,w/8]
// One pass over x.  A thread owns one 4x4 block (four float4 row loads per channel, coalesced across the warp) and keeps
// the C pooled values in registers; the four threads of an 8x8 cell sit in lanes l, l^1 (x), l^2 (y) and combine by
// shuffle (the mean of four equal-sized means).  The 1x1 convolutions run from shared-memory weights.
// ------------------------------------------------------------------------------------------------------------------
template <int C, int CO>
__global__ void __launch_bounds__(128)
context_pool_kernel(const float* __restrict__ x, const float* __restrict__ wa, const float* __restrict__ ba,
                    const float* __restrict__ wc, const float* __restrict__ bc, float* __restrict__ a, float* __restrict__ c,
                    int h, int w) {
    __shared__ float sWa[C * CO], sWc[C * CO], sBa[CO], sBc[CO];
    for (int i = threadIdx.x; i < C * CO; i += 128) { sWa[i] = __ldg(wa + i); sWc[i] = __ldg(wc + i); }     // [co][ci]
    if (threadIdx.x < CO) { sBa[threadIdx.x] = __ldg(ba + threadIdx.x); sBc[threadIdx.x] = __ldg(bc + threadIdx.x); }
    __syncthreads();
    const int h4 = h / 4, w4 = w / 4, h8 = h / 8, w8 = w / 8;
    // lane -> (x bit, y bit, cell): lanes l, l^1, l^2, l^3 are the 2x2 blocks of one 8x8 cell
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int cell = (blockIdx.x * 4 + wrp) * 8 + (lane >> 2);         // 8x8 cell along x
    const int cy = blockIdx.y, n = blockIdx.z;
    const int bx = 2 * cell + (lane & 1), by = 2 * cy + ((lane >> 1) & 1);   // 4x4 block coordinates
    const bool in = cell < w8;
    float p4[C];
    const float* px = x + (size_t)n * C * h * w + (size_t)(4 * by) * w + 4 * bx;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        float s = 0.f;
        if (in) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(px + (size_t)ch * h * w + (size_t)r * w));
                s += (v.x + v.y) + (v.z + v.w);
            }
        }
        p4[ch] = s * (1.f / 16.f);
    }
    if (in) {
#pragma unroll 2
        for (int co = 0; co < CO; ++co) {
            float acc = sBa[co];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) acc = fmaf(sWa[co * C + ch], p4[ch], acc);
            a[(((size_t)n * CO + co) * h4 + by) * w4 + bx] = fmaxf(acc, 0.f);
        }
    }
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        float s = p4[ch];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        p4[ch] = s * 0.25f;
    }
    if (in && (lane & 3) == 0) {
        for (int co = 0; co < CO; ++co) {
            float acc = sBc[co];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) acc = fmaf(sWc[co * C + ch], p4[ch], acc);
            c[(((size_t)n * CO + co) * h8 + cy) * w8 + cell] = fmaxf(acc, 0.f);
        }
    }
}

template <int C, int CO>
static int launch_context_pool(const float* x, const float* wa, const float* ba, const float* wc, const float* bc,
                               float* a, float* c, int N, int h, int w, cudaStream_t st) {
    const int w8 = w / 8;
    dim3 grid((w8 + 31) / 32, h / 8, N);
    context_pool_kernel<C, CO><<<grid, 128, 0, st>>>(x, wa, ba, wc, bc, a, c, h, w);
    ADAMVS_LAUNCH_RESULT();
}

extern "C" int adamvs_context_pool_supported(int C, int CO) {
    return (C == 32 && CO == 16) || (C == 16 && CO == 8) || (C == 8 && CO == 4);
}

extern "C" int adamvs_context_pool_f32(const float* x, const float* wa, const float* ba, const float* wc, const float* bc,
                                       float* a, float* c, int N, int C, int CO, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(x && wa && ba && wc && bc && a && c && N > 0 && N <= 65535 && h > 0 && w > 0 && h % 8 == 0 && w % 8 == 0);
    ADAMVS_CHECK_ARG(h / 8 <= 65535 && adamvs_context_pool_supported(C, CO) && reinterpret_cast<uintptr_t>(x) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 32) return launch_context_pool<32, 16>(x, wa, ba, wc, bc, a, c, N, h, w, st);
    if (C == 16) return launch_context_pool<16, 8>(x, wa, ba, wc, bc, a, c, N, h, w, st);
    return launch_context_pool<8, 4>(x, wa, ba, wc, bc, a, c, N, h, w, st);
}

extern "C" int adamvs_context_head_supported(int CX, int CCTX, int COUT) {
    return (CX == 32 && CCTX == 16 && COUT == 32) || (CX == 16 && CCTX == 8 && COUT == 16) || (CX == 8 && CCTX == 4 && COUT == 8);
}

extern "C" int adamvs_context_head_f32(const float* x, const float* ctx_a, const float* ctx_c, const float* weight, float* out,
                                       int N, int CX, int CCTX, int COUT, int h, int w, int ha, int wa, int hc, int wc, void* stream) {
    ADAMVS_CHECK_ARG(x && ctx_a && ctx_c && weight && out && N > 0 && N <= 65535 && h > 0 && h <= 65535 && w > 0);
    ADAMVS_CHECK_ARG(ha > 0 && wa > 0 && hc > 0 && wc > 0 && w % 4 == 0);
    ADAMVS_CHECK_ARG(adamvs_context_head_supported(CX, CCTX, COUT));
    ADAMVS_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (CX == 32) return launch_context_head<32, 16, 32, 2>(x, ctx_a, ctx_c, weight, out, N, h, w, ha, wa, hc, wc, st);
    if (CX == 16) return launch_context_head<16, 8, 16, 4>(x, ctx_a, ctx_c, weight, out, N, h, w, ha, wa, hc, wc, st);
    return launch_context_head<8, 4, 8, 4>(x, ctx_a, ctx_c, weight, out, N, h, w, ha, wa, hc, wc, st);
}
