// K3 + K4: the recurrent conv-GRU encoder-decoder regulariser, swept over the D depth planes, followed by the softmax
// depth regression over its logit volume (one more kernel in the same call).
//
// Reference: CostRegNetRED.forward / SliceCostRegNetRED.forward (models/adamvs.py:172-195, 415-424),
// ConvGRUCell.forward (models/module.py:24-52), ConvReLU (module.py:264-270), regression
// (adamvs.py:306-310, 516-531; module.py:617-625).
//
// fp32 parity (depth 1e-4 rel / prob 1e-4 abs) rules out single-pass TF32/BF16 operands here.  The five
// stride-1 convolutions run on tcgen05 with an exact hi/lo tf32 operand split (conv3x3_tc.cuh) when the plane is large
// enough, else - like conv2 and the tail always - on register-tiled FFMA kernels (conv3x3.cuh): DESIGN.md §3.
//
// Per depth plane (all on one stream; at the bench batch every intermediate is far larger than the L2 - 300 MB per 8-channel
// tensor at stage 3, B = 32 - and is re-read from HBM by the next kernel, which measurements show is NOT what bounds the
// chain: DESIGN.md 3, "Round 2: what bounds these kernels"):
//   1 conv1   x1  = relu(conv3x3(F_k))                                  C  -> 8
//   2 gates1  r,u = sigmoid(conv3x3(cat(x1,h1)) + b);  rh1 = r*h1        16 -> 16
//   3 cand1   h1  = u*h1 + (1-u)*tanh(conv3x3(cat(x1,rh1)) + b)          16 -> 8
//   4 conv2   x2  = relu(conv3x3 stride 2 (h1))                          8  -> 16
//   5 gates2 / 6 cand2 at half resolution                                32 -> 32 / 32 -> 16
//   7 up1     y   = relu(convT3x3 s2 (h2) + b + h1)                      16 -> 8
//   8 out     logit = convT3x3 s2 (y) + b  (stages 1-2) | conv3x3(y)+b (stage 3)   -> logits[:, k]      (7 + 8: one kernel)
// After the sweep: depth, confidence = regression over logits [B,D,Ho,Wo] (regress_volume_kernel).
#include "conv3x3.cuh"
#include "conv3x3_tc.cuh"
#include "regress_fused.cuh"
#include <string.h>

namespace adamvs {

// ------------------------------------------------------------------------------------------------
// 7+8 fused: y = relu(convT3x3 s2 p1 op1 (h2; 16->8) + b + h1) is produced per 32x16 tile (with the 1-pixel halo the
// output layer needs, recomputed) straight into shared memory and consumed there by the output layer; y never goes to
// global memory (saves a 32 B/px write and read per plane and one launch).  One thread per half-resolution quad owns the
// 2x2 full-resolution pixels that (iy,ix) is the top-left contributor of (PyTorch ConvTranspose2d scatter form):
//   out(2iy  ,2ix  ) = in(iy,ix) W11
//   out(2iy  ,2ix+1) = in(iy,ix+1) W10 + in(iy,ix) W12
//   out(2iy+1,2ix  ) = in(iy+1,ix) W01 + in(iy,ix) W21
//   out(2iy+1,2ix+1) = in(iy+1,ix+1) W00 + in(iy+1,ix) W02 + in(iy,ix+1) W20 + in(iy,ix) W22
//   UP = false (stage 3): logit = conv3x3(y) + b on the tile            -> y halo on all four sides
//   UP = true  (stages 1-2): logit(2y+a, 2x+b) = convT3x3 s2 (y) + b     -> y halo on the right and bottom
// ------------------------------------------------------------------------------------------------
constexpr int kTailW = 32, kTailH = 16;

// (lo, step) of every pixel's hypothesis line, once per sweep: the x2-upsampling tail blends four neighbours' hypotheses per
// output pixel and plane, and evaluating hyp_line() there (two loads and an IEEE division per corner) was half of its
// instructions.  Same values, computed once.
__global__ void __launch_bounds__(256)
hyp_lines_kernel(HypSpec hs, float2* __restrict__ out, int B, int hw, int D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * hw) return;
    const HypLine l = hyp_line(hs, i / hw, i % hw, hw, D);
    out[i] = make_float2(l.lo, l.step);
}

template <bool UP>
struct TailGeom {
    static constexpr int LO = UP ? 0 : 1;                       // halo before the tile
    static constexpr int YH = kTailH + LO + 1, YW = kTailW + LO + 1;      // y rows / cols held (18x34 | 17x33)
    static constexpr int YP = YW + 3;                            // pitch
    static constexpr int QH = UP ? kTailH / 2 + 1 : kTailH / 2 + 2;       // half-resolution quads covering them (9 | 10)
    static constexpr int QW = UP ? kTailW / 2 + 1 : kTailW / 2 + 2;       // (17 | 18)
    static constexpr int HP = QW + 1 + 1;                        // h2 patch: QW + 1 columns, padded
    static constexpr int HH = QH + 1;
};

template <bool UP>
__global__ void __launch_bounds__(256, 3)
tail_regress_kernel(const float* h2s, const float* __restrict__ wpk_up, const float* __restrict__ up_b,
                    const float* h1, OutWeights ow, float* __restrict__ logits, int k, int D, int h, int w) {
    // h2s / h1 were written by kernels this one may overlap with (programmatic dependent launch: no L1 invalidation in
    // between, and the read-only path's contract does not hold): they are read with ld.global.cg, never __ldg / restrict.
    using G = TailGeom<UP>;
    __shared__ float sH2[16 * G::HH * G::HP];
    __shared__ float sY[8 * G::YH * G::YP];
    __shared__ float sWu[16 * 9 * 8];
    __shared__ float sWo[73];
    const int tid = threadIdx.x;
    const int tiles_x = (w + kTailW - 1) / kTailW;
    const int ox0 = (blockIdx.x % tiles_x) * kTailW, oy0 = (blockIdx.x / tiles_x) * kTailH;
    const int b = blockIdx.z;
    const int h2 = h / 2, w2 = w / 2;
    const size_t hw = (size_t)h * w, hw2 = (size_t)h2 * w2;
    const int qy0 = oy0 / 2 - G::LO, qx0 = ox0 / 2 - G::LO;     // first half-resolution quad

    pdl_launch_dependents();
    for (int i = tid; i < 16 * 9 * 8; i += 256) sWu[i] = __ldg(wpk_up + i);
    if (tid < 72) sWo[tid] = __ldg(ow.w + tid);
    if (tid == 72) sWo[72] = __ldg(ow.b);
    pdl_wait();                                                  // weights above overlapped the previous kernel's tail
    {   // h2 patch: all of a thread's loads in flight before its first shared-memory store
        constexpr int NE = 16 * G::HH * (G::QW + 1), NI = (NE + 255) / 256;
        float hv[NI];
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int i = tid + 256 * j;
            const int cx = i % (G::QW + 1), rc = i / (G::QW + 1);
            const int ry = rc % G::HH, c = rc / G::HH;
            const int gy = qy0 + ry, gx = qx0 + cx;
            hv[j] = (i < NE && gy >= 0 && gy < h2 && gx >= 0 && gx < w2) ? __ldcg(h2s + ((size_t)b * 16 + c) * hw2 + (size_t)gy * w2 + gx) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int i = tid + 256 * j;
            if (i >= NE) break;
            const int cx = i % (G::QW + 1), rc = i / (G::QW + 1);
            const int ry = rc % G::HH, c = rc / G::HH;
            sH2[(c * G::HH + ry) * G::HP + cx] = hv[j];
        }
    }
    __syncthreads();

    // ---- phase A: one half-resolution quad per thread -> 2x2 pixels x 8 channels of y
    if (tid < G::QH * G::QW) {
        const int qy = tid / G::QW, qx = tid - qy * G::QW;
        float acc[4][8];
        // bias + skip connection h1 of the four pixels: requested now, consumed after the FFMA loop
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int fy = 2 * (qy0 + qy) + (q >> 1), fx = 2 * (qx0 + qx) + (q & 1);
            const bool in = fy >= 0 && fy < h && fx >= 0 && fx < w;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                acc[q][c] = in ? __ldg(up_b + c) + __ldcg(h1 + ((size_t)b * 8 + c) * hw + (size_t)fy * w + fx) : 0.f;
        }
#pragma unroll 4
        for (int ci = 0; ci < 16; ++ci) {
            const float* p = sH2 + (ci * G::HH + qy) * G::HP + qx;
            const float v00 = p[0], v01 = p[1], v10 = p[G::HP], v11 = p[G::HP + 1];
            const float* wt = sWu + ci * 72;
#pragma unroll
            for (int c = 0; c < 8; ++c) {                                                 // 9 FFMA, nothing else
                acc[0][c] = fmaf(v00, wt[4 * 8 + c], acc[0][c]);
                acc[1][c] = fmaf(v00, wt[5 * 8 + c], fmaf(v01, wt[3 * 8 + c], acc[1][c]));
                acc[2][c] = fmaf(v00, wt[7 * 8 + c], fmaf(v10, wt[1 * 8 + c], acc[2][c]));
                acc[3][c] = fmaf(v00, wt[8 * 8 + c], fmaf(v01, wt[6 * 8 + c], fmaf(v10, wt[2 * 8 + c], fmaf(v11, wt[0 * 8 + c], acc[3][c]))));
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int fy = 2 * (qy0 + qy) + (q >> 1), fx = 2 * (qx0 + qx) + (q & 1);     // full-resolution pixel
            const int ry = fy - (oy0 - G::LO), rx = fx - (ox0 - G::LO);                  // position inside sY
            if (ry < 0 || ry >= G::YH || rx < 0 || rx >= G::YW) continue;
            const bool in = fy >= 0 && fy < h && fx >= 0 && fx < w;
#pragma unroll
            for (int c = 0; c < 8; ++c)                                                   // outside the image: the output layer's zero padding
                sY[(c * G::YH + ry) * G::YP + rx] = in ? fmaxf(acc[q][c], 0.f) : 0.f;
        }
    }
    __syncthreads();

    // ---- phase B: output layer, two x-adjacent y pixels per thread; logits go to the [B,D,Ho,Wo] volume
    const int ty = tid / 16, tx = (tid % 16) * 2;
    const int y0 = oy0 + ty;
    if (y0 >= h || ox0 + tx >= w) return;                        // w is even: both pixels are in or out
    if (!UP) {
        float lg[2] = {sWo[72], sWo[72]};
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* p = sY + (ci * G::YH + ty + ky) * G::YP + tx;
                const float v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
                const float* wt = sWo + ci * 9 + ky * 3;
                lg[0] = fmaf(v0, wt[0], fmaf(v1, wt[1], fmaf(v2, wt[2], lg[0])));
                lg[1] = fmaf(v1, wt[0], fmaf(v2, wt[1], fmaf(v3, wt[2], lg[1])));
            }
        *reinterpret_cast<float2*>(logits + ((size_t)b * D + k) * hw + (size_t)y0 * w + ox0 + tx) = make_float2(lg[0], lg[1]);
    } else {
        const int Wo = 2 * w;
        const size_t ohw = 4 * hw;
        float lg[2][4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float l00 = sWo[72], l01 = sWo[72], l10 = sWo[72], l11 = sWo[72];
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                const float* p = sY + (ci * G::YH + ty) * G::YP + tx + e;
                const float v00 = p[0], v01 = p[1], v10 = p[G::YP], v11 = p[G::YP + 1];   // zero beyond the image (phase A)
                const float* wt = sWo + ci * 9;
                l00 = fmaf(v00, wt[4], l00);
                l01 = fmaf(v00, wt[5], fmaf(v01, wt[3], l01));
                l10 = fmaf(v00, wt[7], fmaf(v10, wt[1], l10));
                l11 = fmaf(v00, wt[8], fmaf(v01, wt[6], fmaf(v10, wt[2], fmaf(v11, wt[0], l11))));
            }
            lg[e][0] = l00; lg[e][1] = l01; lg[e][2] = l10; lg[e][3] = l11;
        }
        // output rows 2*y0, 2*y0 + 1; columns 2*(ox0 + tx) .. + 3 (16-byte aligned: tx is even)
        float* po = logits + ((size_t)b * D + k) * ohw + (size_t)(2 * y0) * Wo + 2 * (ox0 + tx);
        *reinterpret_cast<float4*>(po) = make_float4(lg[0][0], lg[0][1], lg[1][0], lg[1][1]);
        *reinterpret_cast<float4*>(po + Wo) = make_float4(lg[0][2], lg[0][3], lg[1][2], lg[1][3]);
    }
}

// ------------------------------------------------------------------------------------------------
// The same tail, TMA-fed and persistent (w % 8 == 0).  The kernel above spends most of a tile waiting: h2 patch from
// global -> barrier -> skip connection from global -> FFMA -> barrier -> output layer, with 60-70 % of the threads of a
// 3-CTA SM idle in the first phase (FFMA pipe 18 % busy, 8.9 kclk per tile and SM against ~3.2 kclk of issue slots).
// Here a CTA walks over tiles; one thread requests the NEXT tile's two boxes (h2 patch, h1 = skip connection over the
// quads' whole footprint, image borders zero-filled by the tensor map) while the CTA computes the current one, y is
// written in place over the h1 box (every element is read and written by the one thread that owns its quad), and the
// output layer runs with lanes along x (conflict-free rows, 128 / 256-byte row stores).  Same arithmetic, same order of
// the FFMA chains: bit-identical logits.
// ------------------------------------------------------------------------------------------------
template <bool UP, int TH>
struct TailTma {
    static constexpr int LO = UP ? 0 : 1;
    static constexpr int QH = (TH / 2 + 1 + LO + 1) / 2 * 2, QW = kTailW / 2 + 1 + LO;   // quads (rows, rounded up to pairs, x cols)
    static constexpr int NPAIR = QH / 2 * QW;                                      // vertical quad pairs; x 2 channel halves = threads
    // TMA boxes start on 16-byte boundaries of the global rows: the quads of the stage-3 form begin one quad left of the
    // tile (x = 16t - 1 at half, 32t - 2 at full resolution), so its boxes begin 3 / 2 columns further left
    static constexpr int XH = UP ? 0 : 3, XY = UP ? 0 : 2;                         // first used column of the h2 / h1 box
    static constexpr int YH = 2 * QH, YP = UP ? 36 : 40;                           // y / h1 box: rows, pitch (XY + 2*QW <= YP)
    static constexpr int HH = QH + 1, HP = UP ? 20 : 24;                           // h2 box: rows, pitch (XH + QW + 1 <= HP)
    static constexpr int H2F = 16 * HH * HP, H1F = 8 * YH * YP, STAGE = H2F + H1F; // floats; both multiples of 32
    static constexpr int R = TH / 8;                                               // output rows per thread (8 warps)
    static constexpr size_t SMEM = sizeof(float) * (2 * STAGE + 16 * 72 + 8 * 3 * 4 + 4) + 2 * sizeof(uint64_t);
    static_assert(2 * NPAIR <= 256 && TH % 8 == 0 && XY + 2 * QW <= YP && XH + QW + 1 <= HP, "tile geometry");
    static_assert(2 * SMEM + 2048 <= 227 * 1024, "two CTAs per SM");
    static_assert((H2F * 4) % 128 == 0 && (H1F * 4) % 128 == 0, "TMA destinations are 128-byte aligned");
};

template <bool UP, int TH>
__global__ void __launch_bounds__(256, 2)
tail_tma_kernel(const __grid_constant__ CUtensorMap tmH2, const __grid_constant__ CUtensorMap tmH1,
                const float* __restrict__ wpk_up, const float* __restrict__ up_b, OutWeights ow,
                float* __restrict__ logits, int k, int D, int h, int w, TileGrid tg) {
    using G = TailTma<UP, TH>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sIn = reinterpret_cast<float*>(smem_raw);               // [2][h2 box | h1 box]
    float* sWu = sIn + 2 * G::STAGE;                               // [16][9][8]
    float* sWo = sWu + 16 * 72;                                    // [8][3][4] (kx padded), then the bias
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWo + 8 * 3 * 4 + 4);
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const size_t hw = (size_t)h * w;

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    pdl_launch_dependents();
    for (int i = tid; i < 16 * 72; i += 256) sWu[i] = __ldg(wpk_up + i);
    if (tid < 96) { const int kx = tid & 3; sWo[tid] = kx < 3 ? __ldg(ow.w + (tid >> 2) * 3 + kx) : 0.f; }
    if (tid == 96) sWo[96] = __ldg(ow.b);
    // phase-A role of this thread (the same for every tile): quad pair (rows 2*pa_qi, 2*pa_qi + 1; column pa_qx), channels pa_c0 ..+3
    const bool pa_on = tid < 2 * G::NPAIR;
    const int pa_half = tid >= G::NPAIR, pa_pr = tid - pa_half * G::NPAIR;
    const int pa_qi = pa_pr / G::QW, pa_qx = pa_pr - pa_qi * G::QW, pa_c0 = 4 * pa_half;
    float ub[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ub[c] = __ldg(up_b + pa_c0 + c);
    __syncthreads();
    pdl_wait();                                                    // everything above overlapped the previous kernel's tail

    auto origin = [&](int item, int& b, int& ox0, int& oy0) {
        item = tg.at(item);
        b = tg.by_item.div(item);
        const int t = item - b * (tg.tiles_x * tg.tiles_y);
        const int tyi = tg.by_x.div(t);
        ox0 = (t - tyi * tg.tiles_x) * kTailW; oy0 = tyi * TH;
    };
    auto issue = [&](int item, int s) {                           // one thread
        int b, ox0, oy0;
        origin(item, b, ox0, oy0);
        const int qy0 = oy0 / 2 - G::LO, qx0 = ox0 / 2 - G::LO;
        float* dst = sIn + s * G::STAGE;
        fence_proxy_async();                                       // the stage was last written by generic stores (y in place)
        mbar_expect_tx(&bars[s], G::STAGE * 4);
        tma_load_4d(dst, &tmH2, &bars[s], qx0 - G::XH, qy0, 0, b * 16);
        tma_load_4d(dst + G::H2F, &tmH1, &bars[s], 2 * qx0 - G::XY, 2 * qy0, 0, b * 8);
    };

    if (tid == 0 && (int)blockIdx.x < tg.ntiles) issue(blockIdx.x, 0);
    int n = 0;
    for (int item = blockIdx.x; item < tg.ntiles; item += gridDim.x, ++n) {
        const int s = n & 1;
        if (tid == 0 && item + (int)gridDim.x < tg.ntiles) issue(item + gridDim.x, s ^ 1);
        int b, ox0, oy0;
        origin(item, b, ox0, oy0);
        const int qy0 = oy0 / 2 - G::LO, qx0 = ox0 / 2 - G::LO;
        const float* sH2 = sIn + s * G::STAGE;
        float* sY = sIn + s * G::STAGE + G::H2F;
        mbar_wait(&bars[s], (n >> 1) & 1);

        // ---- phase A: y in place over h1.  A thread owns two vertically adjacent half-resolution quads (2 x 2x2 pixels)
        // and four of the eight channels: per input channel 6 input loads + 9 broadcast weight vectors feed 72 FFMA
        // (one quad x eight channels needed 4 + 18: the kernel was bound by shared-memory wavefronts, l1tex 86 %)
        if (pa_on) {
            const int fx0 = 2 * (qx0 + pa_qx);
            const bool inx = fx0 >= 0 && fx0 < w;                  // h, w even: a quad's rows / columns are in or out pairwise
            float* py = sY + (4 * pa_qi) * G::YP + 2 * pa_qx + G::XY + pa_c0 * G::YH * G::YP;
            float acc[2][4][4];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float2 t0 = *reinterpret_cast<const float2*>(py + c * G::YH * G::YP + (2 * q) * G::YP);
                    const float2 t1 = *reinterpret_cast<const float2*>(py + c * G::YH * G::YP + (2 * q + 1) * G::YP);
                    acc[q][0][c] = ub[c] + t0.x; acc[q][1][c] = ub[c] + t0.y; acc[q][2][c] = ub[c] + t1.x; acc[q][3][c] = ub[c] + t1.y;
                }
#pragma unroll 4
            for (int ci = 0; ci < 16; ++ci) {
                const float* p = sH2 + (ci * G::HH + 2 * pa_qi) * G::HP + pa_qx + G::XH;
                float v[3][2];
#pragma unroll
                for (int r = 0; r < 3; ++r) { v[r][0] = p[r * G::HP]; v[r][1] = p[r * G::HP + 1]; }
                const float* wt = sWu + ci * 72 + pa_c0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float v00 = v[q][0], v01 = v[q][1], v10 = v[q + 1][0], v11 = v[q + 1][1];
                        acc[q][0][c] = fmaf(v00, wt[4 * 8 + c], acc[q][0][c]);
                        acc[q][1][c] = fmaf(v00, wt[5 * 8 + c], fmaf(v01, wt[3 * 8 + c], acc[q][1][c]));
                        acc[q][2][c] = fmaf(v00, wt[7 * 8 + c], fmaf(v10, wt[1 * 8 + c], acc[q][2][c]));
                        acc[q][3][c] = fmaf(v00, wt[8 * 8 + c], fmaf(v01, wt[6 * 8 + c], fmaf(v10, wt[2 * 8 + c], fmaf(v11, wt[0 * 8 + c], acc[q][3][c]))));
                    }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int fy0 = 2 * (qy0 + 2 * pa_qi + q);
                const bool in0 = inx && fy0 >= 0 && fy0 < h, in1 = inx && fy0 + 1 >= 0 && fy0 + 1 < h;   // outside the image: the output layer's zero padding
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    *reinterpret_cast<float2*>(py + c * G::YH * G::YP + (2 * q) * G::YP) =
                        in0 ? make_float2(fmaxf(acc[q][0][c], 0.f), fmaxf(acc[q][1][c], 0.f)) : make_float2(0.f, 0.f);
                    *reinterpret_cast<float2*>(py + c * G::YH * G::YP + (2 * q + 1) * G::YP) =
                        in1 ? make_float2(fmaxf(acc[q][2][c], 0.f), fmaxf(acc[q][3][c], 0.f)) : make_float2(0.f, 0.f);
                }
            }
            // these generic-proxy stores share their buffer with the TMA write of the tile after next: order them against
            // the async proxy here, in the writing thread (a fence in the issuing thread alone left a rare run-to-run
            // difference in test_forward_is_bit_reproducible)
            fence_proxy_async();
        }
        __syncthreads();

        // ---- phase B: output layer; lane = column, warp = R consecutive rows
        const int ty0 = wp * G::R, x = ox0 + lane;
        const float ob = sWo[96];
        if (!UP) {
            float lg[G::R];
#pragma unroll
            for (int r = 0; r < G::R; ++r) lg[r] = ob;
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                float4 wk[3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) wk[ky] = *reinterpret_cast<const float4*>(sWo + (ci * 3 + ky) * 4);
#pragma unroll
                for (int rr = 0; rr < G::R + 2; ++rr) {            // y row oy0 + ty0 + rr - 1 = box row ty0 + rr + 1
                    const float* p = sY + (ci * G::YH + ty0 + rr + 1) * G::YP + lane + 1 + G::XY;
                    const float v0 = p[0], v1 = p[1], v2 = p[2];
#pragma unroll
                    for (int r = 0; r < G::R; ++r) {
                        const int ky = rr - r;
                        if (ky >= 0 && ky < 3) lg[r] = fmaf(v0, wk[ky].x, fmaf(v1, wk[ky].y, fmaf(v2, wk[ky].z, lg[r])));
                    }
                }
            }
            if (x < w) {
#pragma unroll
                for (int r = 0; r < G::R; ++r)
                    if (oy0 + ty0 + r < h) logits[((size_t)b * D + k) * hw + (size_t)(oy0 + ty0 + r) * w + x] = lg[r];
            }
        } else {
            float lg[G::R][4];
#pragma unroll
            for (int r = 0; r < G::R; ++r) lg[r][0] = lg[r][1] = lg[r][2] = lg[r][3] = ob;
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                float4 wk[3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) wk[ky] = *reinterpret_cast<const float4*>(sWo + (ci * 3 + ky) * 4);
                float va[G::R + 1], vb[G::R + 1];                  // y(row, x), y(row, x + 1); zero beyond the image (phase A)
#pragma unroll
                for (int rr = 0; rr < G::R + 1; ++rr) {
                    const float* p = sY + (ci * G::YH + ty0 + rr) * G::YP + lane;
                    va[rr] = p[0]; vb[rr] = p[1];
                }
#pragma unroll
                for (int r = 0; r < G::R; ++r) {
                    const float v00 = va[r], v01 = vb[r], v10 = va[r + 1], v11 = vb[r + 1];
                    lg[r][0] = fmaf(v00, wk[1].y, lg[r][0]);
                    lg[r][1] = fmaf(v00, wk[1].z, fmaf(v01, wk[1].x, lg[r][1]));
                    lg[r][2] = fmaf(v00, wk[2].y, fmaf(v10, wk[0].y, lg[r][2]));
                    lg[r][3] = fmaf(v00, wk[2].z, fmaf(v01, wk[2].x, fmaf(v10, wk[0].z, fmaf(v11, wk[0].x, lg[r][3]))));
                }
            }
            if (x < w) {
                const int Wo = 2 * w;
#pragma unroll
                for (int r = 0; r < G::R; ++r) {
                    const int y0 = oy0 + ty0 + r;
                    if (y0 >= h) break;
                    float* po = logits + ((size_t)b * D + k) * (4 * hw) + (size_t)(2 * y0) * Wo + 2 * x;
                    *reinterpret_cast<float2*>(po) = make_float2(lg[r][0], lg[r][1]);
                    *reinterpret_cast<float2*>(po + Wo) = make_float2(lg[r][2], lg[r][3]);
                }
            }
        }
        __syncthreads();                                           // y fully consumed before the tile after next lands here
    }
}

template <bool UP, int TH>
static cudaError_t launch_tail_tma(const CUtensorMap& tmH2, const CUtensorMap& tmH1, const float* wpk_up, const float* up_b,
                                   OutWeights ow, float* logits, int k, int D, int h, int w, int B, int flip, cudaStream_t st) {
    using G = TailTma<UP, TH>;
    auto kern = tail_tma_kernel<UP, TH>;
    static bool ready[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = dev < 64 ? dev : 63;
    if (!ready[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        if (e != cudaSuccess) return e;
        ready[dev] = true;
    }
    TileGrid tg{};
    tg.tiles_x = (w + kTailW - 1) / kTailW;
    tg.tiles_y = (h + TH - 1) / TH;
    tg.ntiles = tg.tiles_x * tg.tiles_y * B;
    if (tg.ntiles >= (1 << 26)) return cudaErrorInvalidValue;
    tg.by_x = FastDiv(tg.tiles_x); tg.by_item = FastDiv(tg.tiles_x * tg.tiles_y);
    tg.flip = flip;
    int ctas = 2 * sm_count();
    if (ctas > tg.ntiles) ctas = tg.ntiles;
    return launch_pdl(kern, dim3(ctas), dim3(256), G::SMEM, st, tmH2, tmH1, wpk_up, up_b, ow, logits, k, D, h, w, tg);
}

// ------------------------------------------------------------------------------------------------
// Regression over the logit volume, once per stage: one thread per output pixel walks the D planes (every load a
// coalesced row segment) with the running softmax / un-shifted-exp state in registers.
// The state used to live in global memory and was updated by the tail of every plane: 24 B read + written per output
// pixel and plane - 29 GB per step at stage 2 (B = 16), three times the bytes of writing the logits once and reading
// them once here, and the reason tail<UP> was the top kernel of the step (profiles/r02p_launches_bench_b16.txt).
// Hypotheses as the reference takes them: planes, per pixel, or per pixel bilinearly upsampled x2 (align_corners =
// False) together with the logits (module.py:622 / adamvs.py:522).
// ------------------------------------------------------------------------------------------------
template <bool UP>
__global__ void __launch_bounds__(256)
regress_volume_kernel(const float* __restrict__ logits, HypSpec hs, const float2* __restrict__ hlines, int prob_mode,
                      float* __restrict__ depth, float* __restrict__ conf, int D, int h, int w) {
    const int Wo = UP ? 2 * w : w, Ho = UP ? 2 * h : h;
    const size_t ohw = (size_t)Ho * Wo;
    const int hw = h * w;
    const size_t o = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (o >= ohw) return;
    const int oy = (int)(o / Wo), ox = (int)(o - (size_t)oy * Wo);
    auto line_at = [&](int p) {
        if (hlines == nullptr) return hyp_line(hs, b, p, hw, D);
        const float2 t = __ldg(hlines + (size_t)b * hw + p);
        return HypLine{t.x, t.y};
    };
    HypLine c00, c01, c10, c11;
    Lerp ly{0, 0, 1.f, 0.f}, lx{0, 0, 1.f, 0.f};
    const bool blend = UP && hs.mode == ADAMVS_HYP_PER_PIXEL;
    if (blend) {
        ly = lerp_index(oy, 0.5f, h); lx = lerp_index(ox, 0.5f, w);
        c00 = line_at(ly.i0 * w + lx.i0); c01 = line_at(ly.i0 * w + lx.i1);
        c10 = line_at(ly.i1 * w + lx.i0); c11 = line_at(ly.i1 * w + lx.i1);
    } else {
        c00 = hs.mode == ADAMVS_HYP_PLANES ? hyp_line(hs, b, 0, hw, D) : line_at(oy * w + ox);
        c01 = c10 = c11 = c00;
    }
    const float* p = logits + (size_t)b * D * ohw + o;
    const RegressState none{nullptr, nullptr, nullptr};
    RegressAcc r = regress_load(none, 0, 0, prob_mode);
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
        const float lg = __ldg(p + (size_t)k * ohw);
        float dval;
        if (blend) dval = ly.l0 * (lx.l0 * hyp_at(c00, k) + lx.l1 * hyp_at(c01, k)) + ly.l1 * (lx.l0 * hyp_at(c10, k) + lx.l1 * hyp_at(c11, k));
        else dval = hyp_at(c00, k);
        regress_step(r, lg, dval, prob_mode);
    }
    regress_store(none, (size_t)b * ohw + o, r, D - 1, D, prob_mode, depth, conf);
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers (non-TMA fallback: any even h, w; fixed 16x16 tiles)
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

template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI>
static cudaError_t launch_conv_auto(const ConvArgs& a, int B, cudaStream_t st) {
    return launch_conv<CA, CB, COUT, COB, STRIDE, EPI, 16, 16>(a, B, st);
}

using Gates1 = ConvLayer<8, 8, 16, 16, 1, EPI_GATES>;
using Cand1 = ConvLayer<8, 8, 8, 8, 1, EPI_CAND>;
using Conv2 = ConvLayer<8, 0, 16, 16, 2, EPI_RELU>;
using Gates2 = ConvLayer<16, 16, 32, 16, 1, EPI_GATES>;
using Cand2 = ConvLayer<16, 16, 16, 16, 1, EPI_CAND>;
template <int C> using Conv1 = ConvLayer<C, 0, 8, 8, 1, EPI_RELU>;
// the same layers on the tensor cores (conv3x3_tc.cuh); conv2 (stride 2) and the tail stay on the FFMA kernels
using TcGates1 = TcLayer<8, 8, 16, EPI_GATES>;
using TcCand1 = TcLayer<8, 8, 8, EPI_CAND>;
using TcGates2 = TcLayer<16, 16, 32, EPI_GATES>;
using TcCand2 = TcLayer<16, 16, 16, EPI_CAND>;
template <int C> using TcConv1 = TcLayer<C, 0, 8, EPI_RELU>;

struct Workspace {
    float *pk_conv1, *pk_gates1, *pk_cand1, *pk_conv2, *pk_gates2, *pk_cand2, *pk_up1;
    float *x1, *h1, *rh1, *u1, *x2, *h2, *rh2, *u2, *y, *logits;
    size_t total;
};

static Workspace carve(float* base, int B, int C, int D, int h, int w, int out_up) {
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
    ws.logits = take((size_t)B * D * ohw);                      // [B,D,Ho,Wo], unused when the caller passes logits_out
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
    if (B <= 0 || C <= 0 || D <= 0 || h <= 0 || w <= 0) return 0;
    return carve(nullptr, B, C, D, h, w, out_up).total;
}

#ifdef ADAMVS_TC_TRACE
extern "C" int adamvs_tc_trace_read(long long* host) {
    return (int)cudaMemcpyFromSymbol(host, adamvs::g_tc_trace, sizeof(adamvs::g_tc_trace));
}
#endif

#define ADAMVS_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

// Default arithmetic of adamvs_regnet_red_f32.  ADAMVS_K3_MATH=ffma|tc|tf32 is a test / measurement hook read once per
// process (like ADAMVS_CONV_CFG); production callers pick a mode explicitly through adamvs_regnet_red_ex_f32.
static int default_math() {
    static const int m = [] {
        const char* e = getenv("ADAMVS_K3_MATH");
        if (e && !strcmp(e, "ffma")) return ADAMVS_MATH_FFMA;
        if (e && !strcmp(e, "tc")) return ADAMVS_MATH_TC_FP32;
        if (e && !strcmp(e, "tf32")) return ADAMVS_MATH_TC_TF32;
        if (e && !strcmp(e, "auto")) return ADAMVS_MATH_AUTO;
        return ADAMVS_MATH_DEFAULT;
    }();
    return m;
}

extern "C" int adamvs_regnet_red_f32(const float* volume, const adamvs_regnet_weights* hwts,
                                     int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                     int out_up, int prob_mode,
                                     float* workspace, size_t workspace_floats,
                                     float* depth, float* conf, float* logits_out,
                                     int B, int C, int D, int h, int w, void* stream) {
    return adamvs_regnet_red_ex_f32(volume, hwts, hyp_mode, hyp_src, hyp_ncol, half_range, out_up, prob_mode, default_math(),
                                    workspace, workspace_floats, depth, conf, logits_out, B, C, D, h, w, stream);
}

// L2-resident sub-batching.  One plane step touches, per batch item, four 8-channel full-resolution tensors, four
// 16-channel half-resolution ones, one plane of the volume and one of the logits; with the whole batch in one launch
// (B = 32: 300 MB per 8-channel tensor at stage 3) every kernel of the chain reads its predecessor's output back
// from HBM.  The sweep is therefore run over sub-batches whose working set fits the 126 MB L2, re-using the SAME
// workspace slots for every sub-batch, so x1 / r*h / u / the GRU states stay in L2 from the kernel that writes them
// to the kernels that read them.
// MEASURED (B200, bench B = 32, profiles/r2c_k3_l2_subbatch_sweep.txt): it is a LOSS - stage 1/2/3 10.9/21.9/19.6 ms with
// whole-batch launches against 14.4/33.8/28.8 ms at a 72 MB budget (48 MB: 16.9/41.8/28.8; 110 MB: 12.6/29.1/28.8): the
// launches shrink to ~10 us, where per-launch prologue (weight split, TMEM allocation) and tile-quantisation tails cost
// more than the L2 hits return.  Default therefore OFF (0 = one launch over the whole batch); ADAMVS_K3_L2_MB=<MB>
// enables it for measurements.
static int l2_budget_mb() {
    static const int mb = [] { const char* e = getenv("ADAMVS_K3_L2_MB"); return e ? atoi(e) : 0; }();
    return mb;
}

static int regnet_sweep(const float* volume, const adamvs_regnet_weights* hwts, const Workspace& ws, const HypSpec& hs,
                        const float2* hlines, int out_up, int prob_mode, int math_mode, float* logits,
                        float* depth, float* conf, int B, int C, int D, int h, int w, cudaStream_t st);

extern "C" int adamvs_regnet_red_ex_f32(const float* volume, const adamvs_regnet_weights* hwts,
                                        int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                        int out_up, int prob_mode, int math_mode,
                                        float* workspace, size_t workspace_floats,
                                        float* depth, float* conf, float* logits_out,
                                        int B, int C, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(volume && hwts && workspace && depth && conf && hyp_src);
    ADAMVS_CHECK_ARG(math_mode == ADAMVS_MATH_FFMA || math_mode == ADAMVS_MATH_TC_FP32 || math_mode == ADAMVS_MATH_TC_TF32 ||
                     math_mode == ADAMVS_MATH_AUTO);
    ADAMVS_CHECK_ARG(B > 0 && B <= 65535 && D >= 2 && h > 0 && w > 0 && (h % 2) == 0 && (w % 2) == 0 && h <= 65535);
    ADAMVS_CHECK_ARG(C == 8 || C == 16 || C == 32);
    ADAMVS_CHECK_ARG(logits_out == nullptr || reinterpret_cast<uintptr_t>(logits_out) % 16 == 0);
    ADAMVS_CHECK_ARG(prob_mode == ADAMVS_PROB_SOFTMAX || prob_mode == ADAMVS_PROB_EXP_EPS);
    ADAMVS_CHECK_ARG(hyp_mode == ADAMVS_HYP_PLANES ? hyp_ncol >= 2 : (hyp_mode == ADAMVS_HYP_PER_PIXEL && half_range));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws = carve(workspace, B, C, D, h, w, out_up);
    if (ws.total > workspace_floats) return ADAMVS_ENOSPACE;
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    const size_t hw = (size_t)h * w, hw2 = (size_t)(h / 2) * (w / 2), ohw = out_up ? 4 * hw : hw;

    // one-off per call: weights into [ci][tap][co]
    auto pack = [&](const float* src, float* dst, int cout, int cin, int tr) {
        const int n = cout * cin * 9;
        pack_conv_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, cout, cin, tr, 0);
    };
    pack(hwts->conv1_w, ws.pk_conv1, 8, C, 0);
    pack(hwts->gates1_w, ws.pk_gates1, 16, 16, 0);
    pack(hwts->cand1_w, ws.pk_cand1, 8, 16, 0);
    pack(hwts->conv2_w, ws.pk_conv2, 16, 8, 0);
    pack(hwts->gates2_w, ws.pk_gates2, 32, 32, 0);
    pack(hwts->cand2_w, ws.pk_cand2, 16, 32, 0);
    pack(hwts->up1_w, ws.pk_up1, 8, 16, 1);
    ADAMVS_TRY(cudaGetLastError());
    float* logits = logits_out ? logits_out : ws.logits;
    const float2* hlines = nullptr;                       // per-pixel hypothesis lines for the regression (in ws.y: y itself
    if (hyp_mode == ADAMVS_HYP_PER_PIXEL) {               // never leaves shared memory since the tail was fused)
        hyp_lines_kernel<<<(unsigned)(((size_t)B * hw + 255) / 256), 256, 0, st>>>(hs, reinterpret_cast<float2*>(ws.y), B, (int)hw, D);
        ADAMVS_TRY(cudaGetLastError());
        hlines = reinterpret_cast<const float2*>(ws.y);
    }

    // sub-batches sized to the L2 budget, balanced
    const double item_bytes = 4.0 * (4.0 * 8 * hw + 4.0 * 16 * hw2 + (double)C * hw + (double)ohw);
    int Bs = B;
    if (l2_budget_mb() > 0) {
        long long fit = (long long)(l2_budget_mb() * 1048576.0 / item_bytes);
        if (fit < 1) fit = 1;
        if (fit < B) { const int nsub = (int)((B + fit - 1) / fit); Bs = (B + nsub - 1) / nsub; }
    }
    for (int b0 = 0; b0 < B; b0 += Bs) {
        const int nb = (B - b0 < Bs) ? B - b0 : Bs;
        HypSpec hsub = hs;
        hsub.src = hyp_src + (hyp_mode == ADAMVS_HYP_PLANES ? (size_t)b0 * hyp_ncol : (size_t)b0 * hw);
        const int rc = regnet_sweep(volume + (size_t)b0 * C * D * hw, hwts, ws, hsub, hlines ? hlines + (size_t)b0 * hw : nullptr,
                                    out_up, prob_mode, math_mode, logits + (size_t)b0 * D * ohw, depth + (size_t)b0 * ohw,
                                    conf + (size_t)b0 * ohw, nb, C, D, h, w, st);
        if (rc != 0) return rc;
    }
    return 0;
}

// One sweep over the D planes for B (sub-batch) items in workspace slots 0 .. B-1, then their regression.
static int regnet_sweep(const float* volume, const adamvs_regnet_weights* hwts, const Workspace& ws, const HypSpec& hs,
                        const float2* hlines, int out_up, int prob_mode, int math_mode, float* logits,
                        float* depth, float* conf, int B, int C, int D, int h, int w, cudaStream_t st) {
    const int h2 = h / 2, w2 = w / 2;
    const size_t hw = (size_t)h * w, hw2 = (size_t)h2 * w2;
    // states to zero (adamvs.py:175-176 / 448-449)
    ADAMVS_TRY(cudaMemsetAsync(ws.h1, 0, sizeof(float) * B * 8 * hw, st));
    ADAMVS_TRY(cudaMemsetAsync(ws.h2, 0, sizeof(float) * B * 16 * hw2, st));
    const OutWeights ow{hwts->out_w, hwts->out_b};
    float* workspace = ws.pk_conv1;                         // alignment check below: the carve starts here

    // ---- per-layer arguments (fixed for the whole sweep; only conv1's plane index changes)
    ConvArgs a1{}, a2{}, a3{}, a4{}, a5{}, a6{};
    a1.inA = volume; a1.strideA_c = (long long)D * hw; a1.strideA_b = (long long)C * D * hw; a1.planesA = C;
    a1.wpk = ws.pk_conv1; a1.out0 = ws.x1; a1.hin = h; a1.win = w; a1.hout = h; a1.wout = w;
    a2.inA = ws.x1; a2.strideA_c = hw; a2.strideA_b = 8 * hw; a2.planesA = 8;
    a2.inB = ws.h1; a2.strideB_c = hw; a2.strideB_b = 8 * hw; a2.planesB = 8;
    a2.wpk = ws.pk_gates1; a2.bias = hwts->gates1_b; a2.out0 = ws.rh1; a2.out1 = ws.u1; a2.hstate = ws.h1;
    a2.hin = h; a2.win = w; a2.hout = h; a2.wout = w;
    a3 = a2; a3.inB = ws.rh1; a3.wpk = ws.pk_cand1; a3.bias = hwts->cand1_b; a3.out0 = ws.h1; a3.out1 = nullptr; a3.ugate = ws.u1;
    a4.inA = ws.h1; a4.strideA_c = hw; a4.strideA_b = 8 * hw; a4.planesA = 8; a4.wpk = ws.pk_conv2; a4.out0 = ws.x2;
    a4.hin = h; a4.win = w; a4.hout = h2; a4.wout = w2;
    a5.inA = ws.x2; a5.strideA_c = hw2; a5.strideA_b = 16 * hw2; a5.planesA = 16;
    a5.inB = ws.h2; a5.strideB_c = hw2; a5.strideB_b = 16 * hw2; a5.planesB = 16;
    a5.wpk = ws.pk_gates2; a5.bias = hwts->gates2_b; a5.out0 = ws.rh2; a5.out1 = ws.u2; a5.hstate = ws.h2;
    a5.hin = h2; a5.win = w2; a5.hout = h2; a5.wout = w2;
    a6 = a5; a6.inB = ws.rh2; a6.wpk = ws.pk_cand2; a6.bias = hwts->cand2_b; a6.out0 = ws.h2; a6.out1 = nullptr; a6.ugate = ws.u2;

    // TMA needs 16-byte global row strides at both resolutions
    bool tma = (w % 8 == 0) && ((reinterpret_cast<uintptr_t>(volume) | reinterpret_cast<uintptr_t>(workspace)) % 16 == 0);
    ConvPlan p1{}, p2{}, p3{}, p4{}, p5{}, p6{};
    if (tma) {
        bool ok = (C == 8 ? Conv1<8>::plan(p1, a1, B, D) : C == 16 ? Conv1<16>::plan(p1, a1, B, D) : Conv1<32>::plan(p1, a1, B, D));
        ok = ok && Gates1::plan(p2, a2, B, 1) && Cand1::plan(p3, a3, B, 1) && Conv2::plan(p4, a4, B, 1)
                && Gates2::plan(p5, a5, B, 1) && Cand2::plan(p6, a6, B, 1);
        tma = ok;
    }
    // Which layers run on the tensor cores.  AUTO: measured per layer on B200 (profiles/r02g): with >= ~2 tiles per SM the
    // tcgen05 kernels beat the FFMA kernels on every layer (stage 3, B = 8: 53/131/111/105/65 us against 69/279/155/267/161);
    // smaller planes keep the split-K FFMA kernels, which spread one plane over all SMs.
    const int prec = math_mode == ADAMVS_MATH_TC_TF32 ? PREC_TF32 : PREC_FP32X3;
    const long long px = (long long)h * w * B;
    const bool all_tc = tma && (math_mode == ADAMVS_MATH_TC_FP32 || math_mode == ADAMVS_MATH_TC_TF32);
    const bool auto_tc = tma && math_mode == ADAMVS_MATH_AUTO;
    const bool tc_full = all_tc || (auto_tc && px >= 30000), tc_half = all_tc || (auto_tc && px / 4 >= 30000);
    const bool tc1 = tc_full, tc2 = tc_full, tc3 = tc_full, tc5 = tc_half, tc6 = tc_half;
    ConvPlan q1{}, q2{}, q3{}, q5{}, q6{};
    {
        bool ok = true;
        if (tc1) ok = ok && (C == 8 ? TcConv1<8>::plan(q1, a1, B, D) : C == 16 ? TcConv1<16>::plan(q1, a1, B, D) : TcConv1<32>::plan(q1, a1, B, D));
        if (tc2) ok = ok && TcGates1::plan(q2, a2, B, 1);
        if (tc3) ok = ok && TcCand1::plan(q3, a3, B, 1);
        if (tc5) ok = ok && TcGates2::plan(q5, a5, B, 1);
        if (tc6) ok = ok && TcCand2::plan(q6, a6, B, 1);
        if (!ok) return ADAMVS_EINVAL;
    }

    // tail: TMA-fed persistent kernel when the rows are TMA-addressable; 32x24 tiles when there are enough of them to
    // keep 2 CTAs per SM busy for several rounds, else 32x16 (ADAMVS_TAIL_CFG=0|16|24: test / measurement hook)
    int tail_th = 0;
    CUtensorMap tmH2, tmH1;
    if (tma) {
        static const int forced = [] {
            const char* e = getenv("ADAMVS_TAIL_CFG");
            return !e ? -1 : !strcmp(e, "0") ? 0 : !strcmp(e, "16") ? 16 : !strcmp(e, "24") ? 24 : -1;
        }();
        const long long items24 = (long long)((w + kTailW - 1) / kTailW) * ((h + 23) / 24) * B;
        tail_th = forced >= 0 ? forced : (items24 >= 8LL * 2 * sm_count() ? 24 : 16);
        if (!out_up && tail_th == 24) tail_th = 16;             // the stage-3 form's wider boxes: two CTAs per SM only at 32x16
        if (tail_th) {
            int hp, hh, yp, yh;
            auto geom = [&](auto g) { using T = decltype(g); hp = T::HP; hh = T::HH; yp = T::YP; yh = T::YH; };
            if (out_up) { if (tail_th == 24) geom(TailTma<true, 24>{}); else geom(TailTma<true, 16>{}); }
            else geom(TailTma<false, 16>{});
            if (!make_tmap_4d(&tmH2, ws.h2, w2, h2, 1, (long long)B * 16, hp, hh, 16) ||
                !make_tmap_4d(&tmH1, ws.h1, w, h, 1, (long long)B * 8, yp, yh, 8)) tail_th = 0;
        }
    }

    // Tile order of consecutive kernels alternates (first-to-last, last-to-first, ...): a persistent kernel leaves the END
    // of the tensor it wrote in the L2; its consumer starts there.  ADAMVS_K3_ZIGZAG=0 switches it off (measurement hook).
    static const bool zigzag = [] { const char* e = getenv("ADAMVS_K3_ZIGZAG"); return !(e && *e == '0'); }();
    int launches = 0;
    auto zz = [&]() { return zigzag ? (launches++ & 1) : 0; };
    for (int k = 0; k < D; ++k) {
        if (tma) {
            p1.tg.flip = q1.tg.flip = zz();
            p2.tg.flip = q2.tg.flip = zz();
            p3.tg.flip = q3.tg.flip = zz();
            p4.tg.flip = zz();
            p5.tg.flip = q5.tg.flip = zz();
            p6.tg.flip = q6.tg.flip = zz();
            if (tc1) {
                q1.args.k = k;
                if (C == 8) ADAMVS_TRY(TcConv1<8>::launch(q1, B, prec, st));
                else if (C == 16) ADAMVS_TRY(TcConv1<16>::launch(q1, B, prec, st));
                else ADAMVS_TRY(TcConv1<32>::launch(q1, B, prec, st));
            } else {
                p1.args.k = k;
                if (C == 8) ADAMVS_TRY(Conv1<8>::launch(p1, B, st));
                else if (C == 16) ADAMVS_TRY(Conv1<16>::launch(p1, B, st));
                else ADAMVS_TRY(Conv1<32>::launch(p1, B, st));
            }
            if (tc2) ADAMVS_TRY(TcGates1::launch(q2, B, prec, st)); else ADAMVS_TRY(Gates1::launch(p2, B, st));
            if (tc3) ADAMVS_TRY(TcCand1::launch(q3, B, prec, st)); else ADAMVS_TRY(Cand1::launch(p3, B, st));
            ADAMVS_TRY(Conv2::launch(p4, B, st));
            if (tc5) ADAMVS_TRY(TcGates2::launch(q5, B, prec, st)); else ADAMVS_TRY(Gates2::launch(p5, B, st));
            if (tc6) ADAMVS_TRY(TcCand2::launch(q6, B, prec, st)); else ADAMVS_TRY(Cand2::launch(p6, B, st));
        } else {
            a1.inA = volume + (size_t)k * hw;
            if (C == 8) ADAMVS_TRY(run_conv1<8>(a1, B, st));
            else if (C == 16) ADAMVS_TRY(run_conv1<16>(a1, B, st));
            else ADAMVS_TRY(run_conv1<32>(a1, B, st));
            ADAMVS_TRY((launch_conv_auto<8, 8, 16, 16, 1, EPI_GATES>(a2, B, st)));
            ADAMVS_TRY((launch_conv_auto<8, 8, 8, 8, 1, EPI_CAND>(a3, B, st)));
            ADAMVS_TRY((launch_conv_auto<8, 0, 16, 16, 2, EPI_RELU>(a4, B, st)));
            ADAMVS_TRY((launch_conv_auto<16, 16, 32, 16, 1, EPI_GATES>(a5, B, st)));
            ADAMVS_TRY((launch_conv_auto<16, 16, 16, 16, 1, EPI_CAND>(a6, B, st)));
        }
        // 7+8: up1 + skip + relu -> output layer -> logits[:, k], one launch
        if (tail_th == 24) {
            ADAMVS_TRY((launch_tail_tma<true, 24>(tmH2, tmH1, ws.pk_up1, hwts->up1_b, ow, logits, k, D, h, w, B, zz(), st)));
        } else if (tail_th == 16) {
            if (out_up) ADAMVS_TRY((launch_tail_tma<true, 16>(tmH2, tmH1, ws.pk_up1, hwts->up1_b, ow, logits, k, D, h, w, B, zz(), st)));
            else ADAMVS_TRY((launch_tail_tma<false, 16>(tmH2, tmH1, ws.pk_up1, hwts->up1_b, ow, logits, k, D, h, w, B, zz(), st)));
        } else {
            dim3 grid(((w + kTailW - 1) / kTailW) * ((h + kTailH - 1) / kTailH), 1, B);
            if (out_up) ADAMVS_TRY(launch_pdl(tail_regress_kernel<true>, grid, dim3(256), 0, st, (const float*)ws.h2, (const float*)ws.pk_up1, hwts->up1_b, (const float*)ws.h1, ow, logits, k, D, h, w));
            else ADAMVS_TRY(launch_pdl(tail_regress_kernel<false>, grid, dim3(256), 0, st, (const float*)ws.h2, (const float*)ws.pk_up1, hwts->up1_b, (const float*)ws.h1, ow, logits, k, D, h, w));
        }
    }
    // K4: softmax / expectation / max over the logit volume
    {
        const size_t ohw = out_up ? 4 * hw : hw;
        dim3 grid((unsigned)((ohw + 255) / 256), B, 1);
        if (out_up) regress_volume_kernel<true><<<grid, 256, 0, st>>>(logits, hs, hlines, prob_mode, depth, conf, D, h, w);
        else regress_volume_kernel<false><<<grid, 256, 0, st>>>(logits, hs, hlines, prob_mode, depth, conf, D, h, w);
        ADAMVS_TRY(cudaGetLastError());
    }
    return 0;
}
