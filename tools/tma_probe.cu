// Probe: 4-D TMA tile load with negative start coordinates / boxes larger than the tensor (zero fill).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_probe tools/tma_probe.cu && /tmp/tma_probe
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../adamvs_b200/csrc/tma.cuh"
using namespace adamvs;

template <int IP, int IH>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int x0, int y0, int k, int plane) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* s = reinterpret_cast<float*>(raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s + 8 * IH * IP);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(bar, 8 * IH * IP * 4); tma_load_4d(s, &tm, bar, x0, y0, k, plane); }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 8 * IH * IP; i += blockDim.x) out[i] = s[i];
}

template <int IP, int IH>
int run(int w, int h, int D, int planes, int x0, int y0) {
    std::vector<float> hsrc((size_t)planes * D * h * w);
    for (size_t i = 0; i < hsrc.size(); ++i) hsrc[i] = (float)(i % 9973) + 1.f;
    float *src, *out;
    cudaMalloc(&src, hsrc.size() * 4); cudaMalloc(&out, 8 * IH * IP * 4);
    cudaMemcpy(src, hsrc.data(), hsrc.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    bool ok = make_tmap_4d(&tm, src, w, h, D, planes, IP, IH, 8);
    printf("IP=%d IH=%d w=%d h=%d D=%d planes=%d x0=%d y0=%d encode=%d ", IP, IH, w, h, D, planes, x0, y0, (int)ok);
    if (!ok) { printf("\n"); return 1; }
    const int k = D - 1, p0 = planes - 8;
    probe<IP, IH><<<1, 128, 8 * IH * IP * 4 + 16>>>(tm, out, x0, y0, k, p0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch=%s ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<float> hout(8 * IH * IP);
    cudaMemcpy(hout.data(), out, hout.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < 8; ++c) for (int r = 0; r < IH; ++r) for (int x = 0; x < IP; ++x) {
        const int gy = y0 + r, gx = x0 + x;
        float want = 0.f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) want = hsrc[(((size_t)(p0 + c) * D + k) * h + gy) * w + gx];
        if (hout[(c * IH + r) * IP + x] != want) ++bad;
    }
    printf("mismatches=%d\n", bad);
    return bad;
}

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    switch (v) {
        case 0: return run<32, 8>(64, 32, 1, 8, 0, 0);
        case 1: return run<32, 8>(64, 32, 2, 16, 0, 0);
        case 2: return run<32, 8>(64, 32, 2, 16, -1, -1);
        case 3: return run<36, 10>(64, 32, 2, 16, 0, 0);
        case 4: return run<36, 10>(64, 32, 2, 16, -1, -1);
        case 5: return run<36, 10>(192, 96, 3, 32, -1, -1);
        case 6: return run<36, 10>(24, 16, 2, 32, -1, -1);
        case 7: return run<36, 10>(12, 8, 1, 16, -1, -1);
        case 8: return run<68, 17>(192, 96, 1, 8, -1, -1);
        case 9: return run<40, 10>(64, 32, 2, 16, -4, -1);
        case 10: return run<40, 10>(24, 16, 2, 32, -4, -1);
        case 11: return run<72, 17>(12, 8, 1, 16, -4, -1);
        case 12: return run<40, 10>(64, 32, 2, 16, 60, 28);
        case 13: return run<40, 10>(64, 32, 2, 16, 0, -1);
    }
    return 0;
}
