// Probe for the next step named in DESIGN.md §9: the stride-2 transposed 3x3 convolution of the regulariser's tail
// (ConvTranspose2d(16 -> 8, k=3, s=2, p=1, output_padding=1), reference models/adamvs.py:166-170) as a 2x2-tap implicit GEMM on
// tcgen05 (kind::tf32, exact hi/lo split), checked against an fp64 CPU reference.  NOT part of the library; compile-checked
// in the build container, to be run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tc_upconv tools/tc_upconv_probe.cu && timeout 60 /tmp/tc_upconv
//
// Formulation.  out(2iy+a, 2ix+b, co) = sum over the input pixels (iy+dy, ix+dx), dy <= a, dx <= b, of in[ci] * W[ci][co][ky][kx]
// with ky = (a == 0 ? 1 : (dy == 1 ? 0 : 2)), kx likewise (PyTorch's scatter form).  With the 4 output phases (a,b) stacked
// into 32 "phase channels" this is a 2x2-tap convolution at INPUT resolution:
//   D'[p][dx][phase*8+co] = sum_dy sum_ci in[p + 32 dy][ci] * B[dy][dx][ci][phase*8+co]       (rows: operand address offsets)
//   out_phase[p]          = D'[p][0] + D'[p+1][1]                                               (columns: one warp shuffle)
// Positions have pitch 32 (31 pixels + 1 halo column on the right), an M tile is 4 rows x 32; N = 2 (dx) x 2 (W_hi | W_lo) x 32
// = 128, K = 8 channels per MMA: 2 chunks x 2 dy x {A_hi, A_lo} = 8 MMAs of 64 clk per 128 input positions = 512 output
// pixels x 8 channels - against 1152 FFMA per input pixel on the FFMA tail today.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int CIN = 16, COUT = 8, PH = 4, NPC = PH * COUT;          // 32 phase channels
constexpr int PW = 32, TWV = 31, ROWS = 4, IH = ROWS + 1, NPOS = IH * PW;
constexpr int NB = 2 * NPC, N2 = 2 * NB;                             // [W_hi | W_lo] per dx; both dx: N = 128
constexpr int PLANE_BYTES = NPOS * 16, STAGE_BYTES = 2 * 2 * PLANE_BYTES;   // [hi|lo][2 quads][NPOS][4]
constexpr int B_STEP_BYTES = 2 * N2 * 16;                            // [2 quads][N2 rows][4] of one (dy, chunk)
constexpr int NCH = CIN / 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    const float r = v - hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
    lo = __uint_as_float(l);
}
// tap of weight W[ci][co][ky][kx] that input offset d in {0,1} feeds for output parity a in {0,1}; -1 = none
__host__ __device__ inline int tap_of(int a, int d) { return d > a ? -1 : (a == 0 ? 1 : (d == 1 ? 0 : 2)); }

// in [CIN][h][w], wgt [CIN][COUT][3][3] (ConvTranspose2d layout), out [COUT][2h][2w]; one CTA, tiles in sequence
__global__ void __launch_bounds__(128) upconv_tc(const float* __restrict__ in, const float* __restrict__ wgt, float* __restrict__ out, int h, int w) {
    extern __shared__ __align__(128) unsigned char raw[];
    unsigned char* sA = raw;                                          // [NCH] stages
    unsigned char* sB = sA + NCH * STAGE_BYTES;                       // [2 dy][NCH][2 quads][N2][4]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * NCH * B_STEP_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // B operand: n = dx*NB + {0: W_hi, 1: W_lo}*NPC + phase*8 + co
    for (int i = tid; i < 2 * NCH * 2 * 2 * NPC * 4; i += 128) {
        int r = i;
        const int j = r & 3; r >>= 2;
        const int pc = r % NPC; r /= NPC;
        const int dx = r & 1; r >>= 1;
        const int kq = r & 1; r >>= 1;
        const int s = r % NCH, dy = r / NCH;
        const int phase = pc / COUT, co = pc % COUT, a = phase >> 1, b = phase & 1;
        const int ky = tap_of(a, dy), kx = tap_of(b, dx), ci = 8 * s + 4 * kq + j;
        const float v = (ky < 0 || kx < 0) ? 0.f : wgt[((size_t)ci * COUT + co) * 9 + ky * 3 + kx];
        float hi, lo;
        split_tf32(v, hi, lo);
        float* dst = reinterpret_cast<float*>(sB + (size_t)(dy * NCH + s) * B_STEP_BYTES + kq * N2 * 16) + dx * NB * 4;
        dst[pc * 4 + j] = hi;
        dst[(NPC + pc) * 4 + j] = lo;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int tiles_x = (w + TWV - 1) / TWV, tiles_y = (h + ROWS - 1) / ROWS;
    uint32_t parity = 0;
    for (int tile = 0; tile < tiles_x * tiles_y; ++tile) {
        const int ix0 = (tile % tiles_x) * TWV, iy0 = (tile / tiles_x) * ROWS;
        // operand stages: position (r, c) = input pixel (iy0 + r, ix0 + c), zero outside the image
        for (int i = tid; i < NCH * 2 * NPOS; i += 128) {
            const int pos = i % NPOS, q = (i / NPOS) & 1, s = i / (2 * NPOS);
            const int gy = iy0 + pos / PW, gx = ix0 + pos % PW;
            float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f), l4 = h4;
            if (gy < h && gx < w) {
                const float* p = in + (size_t)(8 * s + 4 * q) * h * w + (size_t)gy * w + gx;
                split_tf32(p[0], h4.x, l4.x); split_tf32(p[(size_t)h * w], h4.y, l4.y);
                split_tf32(p[(size_t)2 * h * w], h4.z, l4.z); split_tf32(p[(size_t)3 * h * w], h4.w, l4.w);
            }
            float4* dst = reinterpret_cast<float4*>(sA + (size_t)s * STAGE_BYTES + (size_t)q * PLANE_BYTES) + pos;
            *dst = h4;
            *(dst + 2 * NPOS) = l4;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            for (int s = 0; s < NCH; ++s)
                for (int dy = 0; dy < 2; ++dy)
                    for (int hl = 0; hl < 2; ++hl) {
                        const uint64_t ad = make_desc(smem_u32(sA) + s * STAGE_BYTES + hl * 2 * PLANE_BYTES + dy * PW * 16, PLANE_BYTES, 128);
                        const uint64_t bd = make_desc(smem_u32(sB) + (dy * NCH + s) * B_STEP_BYTES, N2 * 16, 128);
                        const uint32_t acc = (s | dy | hl) ? 1u : 0u;
                        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                    }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
            parity ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: warp = input row, lane = input column; 8 phase channels per step
        const int iy = iy0 + warp, ix = ix0 + lane;
        const bool valid = lane < TWV && iy < h && ix < w;
        for (int c0 = 0; c0 < NPC; c0 += 8) {
            uint32_t r[4][8];                                        // [dx*2 + {hi,lo}]
            for (int dx = 0; dx < 2; ++dx)
                for (int hl = 0; hl < 2; ++hl) {
                    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dx * NB + hl * NPC + c0);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(r[dx * 2 + hl][0]), "=r"(r[dx * 2 + hl][1]), "=r"(r[dx * 2 + hl][2]), "=r"(r[dx * 2 + hl][3]),
                                   "=r"(r[dx * 2 + hl][4]), "=r"(r[dx * 2 + hl][5]), "=r"(r[dx * 2 + hl][6]), "=r"(r[dx * 2 + hl][7])
                                 : "r"(ta) : "memory");
                }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 8; ++c) {
                float v = __uint_as_float(r[0][c]) + __uint_as_float(r[1][c]);
                v += __shfl_down_sync(0xffffffffu, __uint_as_float(r[2][c]) + __uint_as_float(r[3][c]), 1);
                const int pc = c0 + c, phase = pc / COUT, co = pc % COUT;
                if (valid) out[((size_t)co * 2 * h + 2 * iy + (phase >> 1)) * 2 * w + 2 * ix + (phase & 1)] = v;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
    const int h = 23, w = 70;                                        // ragged: 3 tiles across (31 + 31 + 8), 6 down (4 x 5 + 3)
    std::vector<float> in((size_t)CIN * h * w), wg((size_t)CIN * COUT * 9), out((size_t)COUT * 4 * h * w, -777.f);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.f - 1.f; };
    for (auto& v : in) v = rnd() * 3.f;
    for (auto& v : wg) v = rnd() * 0.4f;
    float *d_in, *d_w, *d_out;
    cudaMalloc(&d_in, in.size() * 4); cudaMalloc(&d_w, wg.size() * 4); cudaMalloc(&d_out, out.size() * 4);
    cudaMemcpy(d_in, in.data(), in.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_w, wg.data(), wg.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_out, out.data(), out.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)NCH * STAGE_BYTES + 2 * NCH * B_STEP_BYTES + 64;
    cudaFuncSetAttribute(upconv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    upconv_tc<<<1, 128, smem>>>(d_in, d_w, d_out, h, w);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
    // fp64 reference: scatter form of ConvTranspose2d(k=3, s=2, p=1, output_padding=1)
    std::vector<double> ref((size_t)COUT * 4 * h * w, 0.0);
    for (int ci = 0; ci < CIN; ++ci)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                for (int co = 0; co < COUT; ++co)
                    for (int ky = 0; ky < 3; ++ky)
                        for (int kx = 0; kx < 3; ++kx) {
                            const int oy = 2 * y - 1 + ky, ox = 2 * x - 1 + kx;
                            if (oy < 0 || oy >= 2 * h || ox < 0 || ox >= 2 * w) continue;
                            ref[((size_t)co * 2 * h + oy) * 2 * w + ox] += (double)in[((size_t)ci * h + y) * w + x] * wg[((size_t)ci * COUT + co) * 9 + ky * 3 + kx];
                        }
    double worst = 0, mx = 0;
    size_t unwritten = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        if (out[i] == -777.f) ++unwritten;
        worst = fmax(worst, fabs(ref[i] - (double)out[i]));
        mx = fmax(mx, fabs(ref[i]));
    }
    printf("upconv 16->8 s2 on tcgen05 (%dx%d -> %dx%d): %s, smem %zu B, max |err| %.3e (max |ref| %.2f), %zu outputs never written\n",
           h, w, 2 * h, 2 * w, cudaGetErrorString(e), smem, worst, mx, unwritten);
    return (e == cudaSuccess && worst < 1e-4 * mx && unwritten == 0) ? 0 : 1;
}
