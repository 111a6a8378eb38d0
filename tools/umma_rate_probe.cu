// Probe: issue rate of back-to-back tcgen05.mma kind::tf32 (M = 128, K = 8, A and B from shared memory, no swizzle,
// K-major) for the N the regulariser layers use.  Answers "how many clocks does one M=128 x N x K=8 MMA really take"
// - the pipe-active metric shows 128*N/256 clocks, the shared-memory read of the 4 KB A tile suggests >= 32.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate_probe.cu && timeout 60 /tmp/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}

template <int N, int SAME_A, int ILV>
__global__ void __launch_bounds__(128) rate(long long* out, int iters) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* sA = reinterpret_cast<float*>(raw);                 // 2 quads x 1200 positions x 4
    float* sB = sA + 2 * 1200 * 4;                             // 2 quads x N rows x 4
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * 256 * 4);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * 1200 * 4 + 2 * 256 * 4; i += 128) sA[i] = 0.f;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bd = make_desc(smem_u32(sB), N * 16, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int t = 0; t < 18; ++t) {                    // 9 taps x {hi, lo}: shifted A start addresses like the conv kernel
                const uint32_t off = SAME_A ? 0u : (uint32_t)(((t % 9) / 3 * 34 + (t % 3)) * 16 + (t / 9) * 0);
                const uint64_t ad = make_desc(smem_u32(sA) + off + (uint32_t)((it & 3) * 128 * 16), 1200 * 16, 128);
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                             ::"r"(tmem + (uint32_t)((ILV > 1 ? (t % ILV) : (it & 1)) * N)), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int SAME_A, int ILV = 1>
static void run(int blocks) {
    long long* d;
    cudaMalloc(&d, sizeof(long long) * blocks);
    const size_t smem = (2 * 1200 * 4 + 2 * 256 * 4) * 4 + 64;
    cudaFuncSetAttribute(rate<N, SAME_A, ILV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    rate<N, SAME_A, ILV><<<blocks, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148] = {0};
    cudaMemcpy(h, d, sizeof(long long) * (blocks < 148 ? blocks : 148), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < blocks && i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%3d %s accumulators=%d blocks=%3d: %s, %.1f clk per MMA (M=128, K=8)\n", N, SAME_A ? "same A " : "shifted", ILV, blocks, cudaGetErrorString(e), (double)mx / (iters * 18.0));
    cudaFree(d);
}

int main() {
    run<16, 0>(1); run<32, 0>(1); run<64, 0>(1); run<128, 0>(1); run<256, 0>(1);
    run<32, 1>(1);
    run<16, 0>(148); run<32, 0>(148); run<64, 0>(148);
    // consecutive MMAs into different accumulators (no read-after-write chain on D)
    run<16, 0, 2>(1); run<16, 0, 4>(1); run<32, 0, 2>(1); run<32, 0, 4>(1); run<64, 0, 2>(1); run<64, 0, 4>(1); run<32, 0, 4>(148);
    return 0;
}
