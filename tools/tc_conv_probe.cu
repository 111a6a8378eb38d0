// Probe: 3x3 convolution (16 -> 16 channels, zero padding) as an implicit GEMM on tcgen05 (kind::tf32), TMEM accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tc_probe tools/tc_conv_probe.cu && timeout 60 /tmp/tc_probe
//
// Layout trick ("padded linear"): the input tile (with halo) sits in shared memory as [channel quad][position][4 ch],
// positions enumerating an 18 x 40 padded tile row-major, 16 bytes per position.  An output position m = oy*40 + ox reads
// tap (ky,kx) at position m + ky*40 + kx + 3, i.e. every tap is the SAME K-major operand at a 16-byte-granular start
// address offset: 9 taps x (Cin/8) k-steps of M=128 x N=16 x K=8 MMAs per 128 output positions, no im2col copy.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int CIN = 16, COUT = 16, TW = 32, TH = 16, IP = 40, IH = TH + 2;
constexpr int NPOS = IH * IP + 8;                     // + slack: garbage output positions read a few rows past the tile
constexpr int MT = TH * IP / 128;                     // 5 M-tiles of 128 output positions
static_assert(TH * IP % 128 == 0, "tile must be a whole number of M tiles");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    return d;                                         // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

template <int MODE>
__global__ void __launch_bounds__(128) conv_tc(const float* __restrict__ x, const float* __restrict__ wgt, float* __restrict__ out, int H, int W) {
    extern __shared__ __align__(128) unsigned char raw[];
    constexpr bool SPLIT = MODE == 1 || MODE == 2;
    constexpr int A_FLOATS = (CIN / 4) * NPOS * 4, B_FLOATS = 9 * (CIN / 8) * 2 * COUT * 4;
    float* sA = reinterpret_cast<float*>(raw);                         // [hi|lo][CIN/4][NPOS][4]
    float* sB = sA + (SPLIT ? 2 : 1) * A_FLOATS;                       // [tap][kstep][kquad][hi n | lo n][4]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (SPLIT ? 2 : 1) * B_FLOATS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (W + TW - 1) / TW;
    const int ox0 = (blockIdx.x % tiles_x) * TW, oy0 = (blockIdx.x / tiles_x) * TH;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // input tile -> [quad][pos][4]; position (r, c) holds image pixel (oy0 - 1 + r, ox0 - 4 + c)
    for (int i = tid; i < (CIN / 4) * NPOS; i += 128) {
        const int q = i / NPOS, pos = i % NPOS;
        const int r = pos / IP, c = pos % IP;
        const int gy = oy0 - 1 + r, gx = ox0 - 4 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < IH && gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t o = (size_t)gy * W + gx, p = (size_t)H * W;
            v = make_float4(x[(4 * q + 0) * p + o], x[(4 * q + 1) * p + o], x[(4 * q + 2) * p + o], x[(4 * q + 3) * p + o]);
        }
        if (SPLIT) {                                                    // hi = top 19 bits (what kind::tf32 reads), lo = exact remainder
            float4 hi, lo;
            hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
            hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
            hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
            hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
            reinterpret_cast<float4*>(sA)[i] = hi;
            reinterpret_cast<float4*>(sA + A_FLOATS)[i] = lo;
        } else {
            reinterpret_cast<float4*>(sA)[i] = v;
        }
    }
    // weights -> [tap][kstep][kquad][n][4]:  B[n][k] = W[n][ci = 8*s + 4*kq + j][tap]
    // split layout: per (tap, kstep, kquad) 2*COUT rows: rows 0..COUT-1 = hi, COUT..2COUT-1 = lo  (one N = 2*COUT operand)
    for (int i = tid; i < B_FLOATS; i += 128) {
        const int j = i & 3, n = (i >> 2) % COUT, kq = (i / (4 * COUT)) % 2, s = (i / (8 * COUT)) % (CIN / 8), t = i / (8 * COUT * (CIN / 8));
        const float v = wgt[((size_t)n * CIN + 8 * s + 4 * kq + j) * 9 + t];
        if (SPLIT) {
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            const int base = (((t * (CIN / 8) + s) * 2 + kq) * 2 * COUT) * 4;
            sB[base + n * 4 + j] = hi;
            sB[base + (COUT + n) * 4 + j] = v - hi;
        } else {
            sB[i] = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        // idesc: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
        auto idesc_n = [](int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        const uint32_t a_lbo = NPOS * 16;
        const uint32_t brows = SPLIT ? 2 * COUT : COUT, b_lbo = brows * 16;
        auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                         " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                         ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        };
        constexpr int DCOLS = MODE == 2 ? 2 * COUT : COUT;
        if (MODE != 3)
        for (int mt = 0; mt < MT; ++mt) {
            for (int t = 0; t < 9; ++t) {
                const int ky = t / 3, kx = t % 3;
                for (int s = 0; s < CIN / 8; ++s) {
                    const uint32_t aoff = (uint32_t)(mt * 128 + ky * IP + kx + 3) * 16 + (uint32_t)s * 2 * a_lbo;
                    const uint64_t ah = make_desc(a_base + aoff, a_lbo, 128);
                    const uint64_t al = make_desc(a_base + A_FLOATS * 4 + aoff, a_lbo, 128);
                    const uint32_t boff = (uint32_t)((t * (CIN / 8) + s) * 2 * brows * 16);
                    const uint64_t bh = make_desc(b_base + boff, b_lbo, 128);                 // rows 0.. (hi, or hi|lo when N = 2*COUT)
                    const uint64_t bl = make_desc(b_base + boff + COUT * 16, b_lbo, 128);     // rows COUT.. (lo)
                    const uint32_t acc = (t | s) ? 1u : 0u;
                    const uint32_t d = tmem + (uint32_t)(mt * DCOLS);
                    if (MODE == 0) mma(d, ah, bh, idesc_n(COUT), acc);
                    if (MODE == 1) { mma(d, ah, bh, idesc_n(COUT), acc); mma(d, al, bh, idesc_n(COUT), 1u); mma(d, ah, bl, idesc_n(COUT), 1u); }
                    if (MODE == 2) { mma(d, ah, bh, idesc_n(2 * COUT), acc); mma(d, al, bh, idesc_n(COUT), 1u); }
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // wait for the MMAs (bounded spin: a wrong descriptor must not hang the box)
    {
        uint32_t done = 0;
        for (long it = 0; it < 20000000L && !done; ++it)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)) : "memory");
        if (!done) { if (tid == 0) printf("block %d: MMA completion never arrived\n", blockIdx.x); asm volatile("trap;"); }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int mt = 0; mt < MT; ++mt) {
        uint32_t r[16];
        constexpr int DCOLS2 = MODE == 2 ? 2 * COUT : COUT;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * DCOLS2);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (MODE == 2) {                                                   // add the a_hi * b_lo half
            uint32_t q[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                           "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
                         : "r"(taddr + COUT) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 16; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(q[c]));
        }
        const int m = mt * 128 + warp * 32 + lane;
        const int oy = oy0 + m / IP, ox = ox0 + m % IP;
        if (m % IP < TW && oy < H && ox < W)
            for (int c = 0; c < COUT; ++c) out[((size_t)c * H + oy) * W + ox] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int MODE>
static int run(const char* name, int H, int W, bool check) {
    std::vector<float> hx((size_t)CIN * H * W), hw((size_t)COUT * CIN * 9), ho((size_t)COUT * H * W);
    srand(1);
    for (auto& v : hx) v = (rand() % 20001 - 10000) / 10000.f;
    for (auto& v : hw) v = (rand() % 20001 - 10000) / 40000.f;
    float *dx, *dw, *dout;
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&dout, ho.size() * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xff, ho.size() * 4);
    constexpr int SP = (MODE == 1 || MODE == 2) ? 2 : 1;
    const size_t smem = (size_t)SP * ((CIN / 4) * NPOS * 16 + 9 * (CIN / 8) * 2 * COUT * 16) + 16;
    cudaFuncSetAttribute(conv_tc<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    conv_tc<MODE><<<tiles, 128, smem>>>(dx, dw, dout, H, W);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) conv_tc<MODE><<<tiles, 128, smem>>>(dx, dw, dout, H, W);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double fl = 2.0 * 9 * CIN * COUT * (double)H * W;
    printf("%-28s %dx%d: %s, smem %zu B, %d tiles, %.3f ms, %.1f TFLOP/s (fp32-equivalent)\n", name, H, W, cudaGetErrorString(e), smem, tiles, ms, fl / ms * 1e-9);
    if (e != cudaSuccess) return 1;
    if (check) {
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0, mx = 0; size_t nan = 0;
        for (int co = 0; co < COUT; ++co)
            for (int y = 0; y < H; ++y)
                for (int xx = 0; xx < W; ++xx) {
                    double a = 0;
                    for (int ci = 0; ci < CIN; ++ci)
                        for (int ky = 0; ky < 3; ++ky)
                            for (int kx = 0; kx < 3; ++kx) {
                                const int gy = y + ky - 1, gx = xx + kx - 1;
                                if (gy >= 0 && gy < H && gx >= 0 && gx < W) a += (double)hx[((size_t)ci * H + gy) * W + gx] * hw[((size_t)co * CIN + ci) * 9 + ky * 3 + kx];
                            }
                    const float o = ho[((size_t)co * H + y) * W + xx];
                    if (!(o == o)) { ++nan; continue; }
                    worst = fmax(worst, fabs((double)o - a)); mx = fmax(mx, fabs(a));
                }
        printf("    max abs err vs fp64 %.3e (max |ref| %.3f), %zu NaN\n", worst, mx, nan);
    }
    cudaFree(dx); cudaFree(dw); cudaFree(dout);
    return 0;
}

int main() {
    if (run<0>("tf32 single pass", 48, 96, true)) return 1;
    if (run<1>("3xTF32 (3 MMAs)", 48, 96, true)) return 1;
    if (run<2>("3xTF32 merged (2 MMAs)", 48, 96, true)) return 1;
    run<3>("no MMA (load + epilogue)", 1536, 3072, false);
    run<0>("tf32 single pass", 1536, 3072, false);
    run<1>("3xTF32 (3 MMAs)", 1536, 3072, false);
    run<2>("3xTF32 merged (2 MMAs)", 1536, 3072, false);
    return 0;
}
