// Micro-probe: sustained FFMA vs packed FFMA2 (fma.rn.f32x2) throughput on sm_100a, with an
// SGEMM-like register pattern (acc[p][c] += a[p]*b[c]).  Decides the inner-loop form of the K3 kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_probe tools/ffma_probe.cu && /tmp/ffma_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ITERS>
__global__ void __launch_bounds__(256) k_ffma(float* out, const float* in) {
    float a[8], b[8], acc[8][8];
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; }
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 1e-6f; }
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%0, %1};\n"
                 "fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc; }"
                 : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

template <int ITERS>
__global__ void __launch_bounds__(256) k_ffma2(float* out, const float* in) {
    float a[8]; float2 b[4], acc[8][4];
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + i];
    for (int i = 0; i < 4; ++i) b[i] = make_float2(in[threadIdx.x + 8 + 2 * i], in[threadIdx.x + 9 + 2 * i]);
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) ffma2(acc[i][j], aa, b[j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 1e-6f; }
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    constexpr int ITERS = 4096;
    const int blocks = 148 * 8, threads = 256;
    float *out, *in;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_ffma<ITERS><<<blocks, threads>>>(out, in); else k_ffma2<ITERS><<<blocks, threads>>>(out, in);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double flops = 2.0 * 64 * ITERS * (double)blocks * threads;
        printf("%s: %.3f ms  %.1f TFLOP/s  (err=%s)\n", which == 0 ? "FFMA " : "FFMA2", best, flops / best * 1e-9,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
