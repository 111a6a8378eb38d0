// Probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written there with tcgen05.st) and B in shared
// memory: (1) does D = A x B^T come out right with A[row = lane][k = column], (2) how many clocks per M=128 x N x K=8 MMA
// (the shared-memory A operand costs 49 clk for any N <= 64, tools/umma_rate_probe.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_ts tools/umma_ts_probe.cu && timeout 60 /tmp/umma_ts
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// out: [128][N] result of the check; clk: clocks of the timed loop
template <int N, int ILV>
__global__ void __launch_bounds__(128) probe(float* out, long long* clk, int iters) {
    __shared__ __align__(128) float sB[2 * N * 4];             // [k quad][n][4]
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // B[n][k] = (n % 5) - 2 + 0.25 * k   (tf32 exact)
    for (int i = tid; i < 2 * N * 4; i += 128) {
        const int j = i & 3, n = (i >> 2) % N, kq = i / (4 * N);
        sB[i] = (float)(n % 5) - 2.f + 0.25f * (float)(kq * 4 + j);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // A[m][k] = (m % 7) - 3 + 0.5 * k  -> thread m writes its row: 8 columns starting at column 256
    {
        uint32_t r[8];
        for (int k = 0; k < 8; ++k) r[k] = __float_as_uint((float)(tid % 7) - 3.f + 0.5f * (float)k);
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bd = make_desc(smem_u32(sB), N * 16, 128);
    if (tid == 0) {
        mma_ts(tmem, tmem + 256u, bd, idesc, 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 8; ++c) out[tid * N + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // rate: 18 accumulating MMAs per group from 18 different A column blocks (like 9 taps x hi/lo), alternating accumulators
    if (tid == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int t = 0; t < 18; ++t)
                mma_ts(tmem + (uint32_t)((ILV > 1 ? (t % ILV) : (it & 1)) * N), tmem + 256u + (uint32_t)(t * 8), bd, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar[1])), "r"(0u) : "memory");
        clk[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int ILV = 1>
static void run() {
    float* d_out; long long* d_clk;
    cudaMalloc(&d_out, sizeof(float) * 128 * N);
    cudaMalloc(&d_clk, sizeof(long long));
    const int iters = 2000;
    probe<N, ILV><<<1, 128>>>(d_out, d_clk, iters);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[128 * 256];
    long long clk = 0;
    cudaMemcpy(h, d_out, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost);
    cudaMemcpy(&clk, d_clk, sizeof(long long), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 8; ++k) ref += ((double)(m % 7) - 3.0 + 0.5 * k) * ((double)(n % 5) - 2.0 + 0.25 * k);
            worst = fmax(worst, fabs(ref - (double)h[m * N + n]));
        }
    printf("N=%3d accumulators=%d A in TMEM: %s, max |err| %.3g (D[5][3] = %.4f), %.1f clk per MMA\n", N, ILV, cudaGetErrorString(e), worst, h[5 * N + 3], (double)clk / (iters * 18.0));
    cudaFree(d_out); cudaFree(d_clk);
}

int main() {
    run<16>(); run<32>(); run<64>(); run<128>();
    run<16, 2>(); run<16, 4>(); run<32, 2>(); run<32, 4>(); run<64, 2>(); run<64, 4>();
    return 0;
}
