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
#include "tma.cuh"

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
    int planesA, planesB;  // channels per batch item of the tensors behind inA / inB (TMA plane coordinate)
    int k;                 // depth-plane coordinate of inA (conv1 reads plane k of the cost volume)
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
// TMA-fed variant (the fast path; needs w % 4 == 0).  The whole input tile of the block — all CIN
// channels with the 1-pixel halo, zero-filled outside the image by the TMA unit — is requested up
// front as CIN/8 boxes, each signalling its own mbarrier, so the FFMA loop on chunk c overlaps the
// arrival of chunks c+1.. and no thread spends instructions on address arithmetic or bounds checks.
// Small planes (stage 1, or B = 1) get more parallelism from PY = 1 patches and from splitting the
// input-channel chunks over KSPLIT warps groups, reduced through shared memory.
// ------------------------------------------------------------------------------------------------
template <int CA, int CB, int COUT, int COB, int STRIDE, int TH, int PY, int KSPLIT>
struct TmaCfg {
    static constexpr int TW = 32;
    static constexpr int CIN = CA + CB, NCHUNK = CIN / CK, NCOG = COB / COT;
    static constexpr int IH = STRIDE == 1 ? TH + 2 : 2 * TH + 1;
    // the innermost TMA start coordinate must be 16-byte aligned (tools/tma_probe.cu), so the box
    // starts 4 columns left of the tile; the 1-pixel halo column is tile column 3
    static constexpr int IP = STRIDE == 1 ? TW + 8 : 2 * TW + 8;
    static constexpr int GROUP = (TW / PX) * (TH / PY);
    static constexpr int NT = GROUP * NCOG * KSPLIT;
    static constexpr int CHUNK_FLOATS = CK * IH * IP;
    static constexpr int NACC = PY * PX * COT;
    static constexpr int RED_FLOATS = (KSPLIT - 1) * GROUP * NCOG * NACC;
    static constexpr int W_FLOATS = CIN * 9 * COB;
    static constexpr size_t SMEM = sizeof(float) * (size_t)(NCHUNK * CHUNK_FLOATS + W_FLOATS + RED_FLOATS) + 8 * NCHUNK;
    static_assert(NCHUNK % KSPLIT == 0, "KSPLIT must divide the chunk count");
    static_assert((GROUP * NCOG) % 32 == 0, "warps must be uniform in (k-slice, channel group)");
    static_assert(TH % PY == 0 && CA % CK == 0 && CB % CK == 0 && COB % COT == 0 && COUT % COB == 0, "bad blocking");
};

template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI, int TH, int PY, int KSPLIT>
__global__ void __launch_bounds__(TmaCfg<CA, CB, COUT, COB, STRIDE, TH, PY, KSPLIT>::NT)
conv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvArgs a) {
    using G = TmaCfg<CA, CB, COUT, COB, STRIDE, TH, PY, KSPLIT>;
    constexpr int TW = G::TW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sIn = reinterpret_cast<float*>(smem_raw);                 // [NCHUNK][CK][IH][IP]
    float* sW = sIn + G::NCHUNK * G::CHUNK_FLOATS;                   // [CIN][9][COB]
    float* sRed = sW + G::W_FLOATS;                                  // [KSPLIT-1][NACC][GROUP*NCOG]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + G::RED_FLOATS);

    const int tid = threadIdx.x;
    const int ks = tid / (G::GROUP * G::NCOG);
    const int gt = tid - ks * (G::GROUP * G::NCOG);                  // thread index inside the k-slice
    const int cog = gt / G::GROUP;
    const int t = gt - cog * G::GROUP;
    const int tx = t % (TW / PX), ty = t / (TW / PX);
    const int tiles_x = (a.wout + TW - 1) / TW;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int cob = blockIdx.y;
    const int b = blockIdx.z;
    const int ox0 = tile_x * TW, oy0 = tile_y * TH;

    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < G::NCHUNK; ++c) mbar_init(&bars[c], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < G::NCHUNK; ++c) {
            mbar_expect_tx(&bars[c], G::CHUNK_FLOATS * 4);
            const bool fromA = c * CK < CA;
            const int plane = fromA ? b * a.planesA + c * CK : b * a.planesB + (c * CK - CA);
            tma_load_4d(sIn + c * G::CHUNK_FLOATS, fromA ? &tmA : &tmB, &bars[c],
                        ox0 * STRIDE - 4, oy0 * STRIDE - 1, fromA ? a.k : 0, plane);
        }
    }
    for (int i = tid; i < G::W_FLOATS; i += G::NT) {
        const int col = i % COB, ct = i / COB;
        sW[i] = __ldg(a.wpk + (size_t)ct * COUT + cob * COB + col);
    }
    __syncthreads();

    float acc[PY][PX][COT];
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int c = 0; c < COT; ++c) acc[j][p][c] = 0.f;

#pragma unroll 1
    for (int chunk = ks; chunk < G::NCHUNK; chunk += KSPLIT) {
        mbar_wait(&bars[chunk], 0);
        const float* sC = sIn + chunk * G::CHUNK_FLOATS;
#pragma unroll 2
        for (int c = 0; c < CK; ++c) {
            const float* wrow = sW + ((chunk * CK + c) * 9) * COB + cog * COT;
            const float* irow = sC + (c * G::IH) * G::IP;
            if (STRIDE == 1) {
#pragma unroll
                for (int r = 0; r < PY + 2; ++r) {
                    const float* ip = irow + (PY * ty + r) * G::IP + PX * tx;
                    const float4 v0 = *reinterpret_cast<const float4*>(ip);
                    const float4 v1 = *reinterpret_cast<const float4*>(ip + 4);
                    const float4 v2 = *reinterpret_cast<const float4*>(ip + 8);
                    const float in[6] = {v0.w, v1.x, v1.y, v1.z, v1.w, v2.x};      // tile columns 4tx+3 .. 4tx+8
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
                    const float4 v0 = *reinterpret_cast<const float4*>(ip);
                    const float4 v1 = *reinterpret_cast<const float4*>(ip + 4);
                    const float4 v2 = *reinterpret_cast<const float4*>(ip + 8);
                    const float in[9] = {v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};   // columns 8tx+3 .. 8tx+11
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

    if (KSPLIT > 1) {                                   // reduce the k-slices into slice 0
        constexpr int GN = G::GROUP * G::NCOG;
        if (ks > 0) {
            float* dst = sRed + (size_t)(ks - 1) * G::NACC * GN + gt;
#pragma unroll
            for (int j = 0; j < PY; ++j)
#pragma unroll
                for (int p = 0; p < PX; ++p)
#pragma unroll
                    for (int c = 0; c < COT; ++c) dst[((j * PX + p) * COT + c) * GN] = acc[j][p][c];
        }
        __syncthreads();
        if (ks > 0) return;
#pragma unroll
        for (int s = 0; s < KSPLIT - 1; ++s) {
            const float* src = sRed + (size_t)s * G::NACC * GN + gt;
#pragma unroll
            for (int j = 0; j < PY; ++j)
#pragma unroll
                for (int p = 0; p < PX; ++p)
#pragma unroll
                    for (int c = 0; c < COT; ++c) acc[j][p][c] += src[((j * PX + p) * COT + c) * GN];
        }
    }

    // ------------------------------------------------------------------ epilogue (float4 along x)
    const int co_base = cob * COB + cog * COT;
    const size_t plane = (size_t)a.hout * a.wout;
    const int ox = ox0 + PX * tx;
    if (ox >= a.wout) return;                           // wout % 4 == 0: a float4 is all in or all out
    if (EPI == EPI_GATES) {
        // r*h needs h at the output pixel: it is input channel CA + (co % HC), already in the tile
        constexpr int HC = COUT / 2;
        if (co_base < HC) {
#pragma unroll
            for (int c = 0; c < G::NCHUNK; ++c) if (c * CK >= CA) mbar_wait(&bars[c], 0);
        }
    }
#pragma unroll
    for (int j = 0; j < PY; ++j) {
        const int oy = oy0 + PY * ty + j;
        if (oy >= a.hout) continue;
        const size_t pix = (size_t)oy * a.wout + ox;
#pragma unroll
        for (int c = 0; c < COT; ++c) {
            const int co = co_base + c;
            float v[4] = {acc[j][0][c], acc[j][1][c], acc[j][2][c], acc[j][3][c]};
            if (EPI == EPI_RELU) {
#pragma unroll
                for (int p = 0; p < 4; ++p) v[p] = fmaxf(v[p], 0.f);
                *reinterpret_cast<float4*>(a.out0 + ((size_t)b * COUT + co) * plane + pix) = make_float4(v[0], v[1], v[2], v[3]);
            } else if (EPI == EPI_GATES) {
                constexpr int HC = COUT / 2;
                const float bc = __ldg(a.bias + co);
#pragma unroll
                for (int p = 0; p < 4; ++p) v[p] = sigmoid_f(v[p] + bc);
                if (co < HC) {
                    const int ci = CA + co;
                    const float4 hh = *reinterpret_cast<const float4*>(
                        sIn + (ci / CK) * G::CHUNK_FLOATS + ((ci % CK) * G::IH + PY * ty + j + 1) * G::IP + PX * tx + 4);
                    v[0] *= hh.x; v[1] *= hh.y; v[2] *= hh.z; v[3] *= hh.w;
                    *reinterpret_cast<float4*>(a.out0 + ((size_t)b * HC + co) * plane + pix) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    *reinterpret_cast<float4*>(a.out1 + ((size_t)b * HC + (co - HC)) * plane + pix) = make_float4(v[0], v[1], v[2], v[3]);
                }
            } else {
                const size_t o = ((size_t)b * COUT + co) * plane + pix;
                const float bc = __ldg(a.bias + co);
                const float4 u = *reinterpret_cast<const float4*>(a.ugate + o);
                const float4 hh = *reinterpret_cast<const float4*>(a.hstate + o);
                const float uu[4] = {u.x, u.y, u.z, u.w}, hv[4] = {hh.x, hh.y, hh.z, hh.w};
#pragma unroll
                for (int p = 0; p < 4; ++p) v[p] = uu[p] * hv[p] + (1.f - uu[p]) * tanhf(v[p] + bc);
                *reinterpret_cast<float4*>(a.out0 + o) = make_float4(v[0], v[1], v[2], v[3]);
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

// Fallback launcher (any even h, w): fixed 16x16 tiles.
template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI>
static cudaError_t launch_conv_auto(const ConvArgs& a, int B, cudaStream_t st) {
    return launch_conv<CA, CB, COUT, COB, STRIDE, EPI, 16, 16>(a, B, st);
}

// ---- TMA path: per-layer plan (configuration + tensor maps), built once per regulariser call -----
struct ConvPlan {
    int cfg;                 // 0 BIG (TH16,PY2), 1 MID (TH8,PY1), 2 SMALL (TH8,PY1, split-K over all chunks)
    CUtensorMap tA, tB;
    ConvArgs args;
};

template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI, int TH, int PY, int KSPLIT>
static cudaError_t launch_tma(const ConvPlan& p, int B, cudaStream_t st) {
    using G = TmaCfg<CA, CB, COUT, COB, STRIDE, TH, PY, KSPLIT>;
    auto kern = conv3x3_tma_kernel<CA, CB, COUT, COB, STRIDE, EPI, TH, PY, KSPLIT>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    const int tiles = ((p.args.wout + 31) / 32) * ((p.args.hout + TH - 1) / TH);
    dim3 grid(tiles, COUT / COB, B);
    kern<<<grid, G::NT, G::SMEM, st>>>(p.tA, p.tB, p.args);
    return cudaGetLastError();
}

template <int CA, int CB, int COUT, int COB, int STRIDE, int EPI>
struct ConvLayer {
    static constexpr int NCHUNK = (CA + CB) / CK;
    static int choose_cfg(int hout, int wout, int B) {
        const long long want = 148LL * 768;                       // ~24 warps per SM
        const long long px = (long long)hout * wout * B;
        const long long t_big = px / 8 * (COUT / COT), t_mid = px / 4 * (COUT / COT);
        if (t_big >= want) return 0;
        if (t_mid >= want || NCHUNK == 1) return 1;
        return 2;
    }
    static bool plan(ConvPlan& p, const ConvArgs& a, int B, int depthA) {
        p.args = a;
        p.cfg = choose_cfg(a.hout, a.wout, B);
        const int TH = p.cfg == 0 ? 16 : 8;
        const int IH = STRIDE == 1 ? TH + 2 : 2 * TH + 1;
        const int IP = STRIDE == 1 ? 40 : 72;
        if (!make_tmap_4d(&p.tA, a.inA, a.win, a.hin, depthA, (long long)B * a.planesA, IP, IH, CK)) return false;
        if (CB > 0) { if (!make_tmap_4d(&p.tB, a.inB, a.win, a.hin, 1, (long long)B * a.planesB, IP, IH, CK)) return false; }
        else p.tB = p.tA;
        return true;
    }
    static cudaError_t launch(const ConvPlan& p, int B, cudaStream_t st) {
        switch (p.cfg) {
            case 0: return launch_tma<CA, CB, COUT, COB, STRIDE, EPI, 16, 2, 1>(p, B, st);
            case 1: return launch_tma<CA, CB, COUT, COB, STRIDE, EPI, 8, 1, 1>(p, B, st);
            default: return launch_tma<CA, CB, COUT, COB, STRIDE, EPI, 8, 1, NCHUNK>(p, B, st);
        }
    }
};

using Gates1 = ConvLayer<8, 8, 16, 16, 1, EPI_GATES>;
using Cand1 = ConvLayer<8, 8, 8, 8, 1, EPI_CAND>;
using Conv2 = ConvLayer<8, 0, 16, 16, 2, EPI_RELU>;
using Gates2 = ConvLayer<16, 16, 32, 16, 1, EPI_GATES>;
using Cand2 = ConvLayer<16, 16, 16, 16, 1, EPI_CAND>;
template <int C> using Conv1 = ConvLayer<C, 0, 8, 8, 1, EPI_RELU>;

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
    ConvPlan p1, p2, p3, p4, p5, p6;
    if (tma) {
        bool ok = (C == 8 ? Conv1<8>::plan(p1, a1, B, D) : C == 16 ? Conv1<16>::plan(p1, a1, B, D) : Conv1<32>::plan(p1, a1, B, D));
        ok = ok && Gates1::plan(p2, a2, B, 1) && Cand1::plan(p3, a3, B, 1) && Conv2::plan(p4, a4, B, 1)
                && Gates2::plan(p5, a5, B, 1) && Cand2::plan(p6, a6, B, 1);
        tma = ok;
    }

    for (int k = 0; k < D; ++k) {
        if (tma) {
            p1.args.k = k;
            if (C == 8) ADAMVS_TRY(Conv1<8>::launch(p1, B, st));
            else if (C == 16) ADAMVS_TRY(Conv1<16>::launch(p1, B, st));
            else ADAMVS_TRY(Conv1<32>::launch(p1, B, st));
            ADAMVS_TRY(Gates1::launch(p2, B, st));
            ADAMVS_TRY(Cand1::launch(p3, B, st));
            ADAMVS_TRY(Conv2::launch(p4, B, st));
            ADAMVS_TRY(Gates2::launch(p5, B, st));
            ADAMVS_TRY(Cand2::launch(p6, B, st));
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
