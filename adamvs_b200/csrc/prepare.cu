// adamvs_cascade_prepare / adamvs_resize_bilinear_f32 — tiny set-up kernels that take the host out
// of the per-plane loop (the reference calls torch.inverse 544 times per depth map, each with a
// host sync: models/module.py:539).
#include "common.cuh"

namespace adamvs {

// 4x4 inverse by Gauss-Jordan with partial pivoting in fp64, then src*inv(ref); one thread per
// (stage, batch item, source view).
__global__ void relproj_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                               const float* __restrict__ p3, int B, int V, float* __restrict__ out) {
    const int Vs = V - 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3 * B * Vs) return;
    const int s = idx / (B * Vs);
    const int b = (idx / Vs) % B;
    const int v = idx % Vs + 1;
    const float* proj = s == 0 ? p1 : (s == 1 ? p2 : p3);
    const float* ref = proj + ((size_t)b * V) * 16;
    const float* src = proj + ((size_t)b * V + v) * 16;
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = (double)ref[i * 4 + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(a[c][c]);
        for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); piv = r; }
        if (piv != c) for (int j = 0; j < 8; ++j) { double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double inv = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= inv;
        for (int r = 0; r < 4; ++r) if (r != c) {
            const double f = a[r][c];
            for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    float* o = out + (size_t)idx * 12;
    for (int i = 0; i < 3; ++i) {
        double m[4];
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += (double)src[i * 4 + k] * a[k][4 + j];
            m[j] = acc;
        }
        o[i * 3 + 0] = (float)m[0]; o[i * 3 + 1] = (float)m[1]; o[i * 3 + 2] = (float)m[2];
        o[9 + i] = (float)m[3];
    }
}

struct CascadeScalars { int ndepth[3]; double ratio[3]; };

__global__ void half_range_kernel(const float* __restrict__ dv, int ncol, int interval_mode, int num_depth,
                                  CascadeScalars cs, float* __restrict__ half_range) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double interval;
    if (interval_mode == ADAMVS_INTERVAL_LAST_COLUMN) interval = (double)dv[ncol - 1];
    else interval = ((double)dv[ncol - 1] - (double)dv[0]) / (double)num_depth;
    for (int s = 0; s < 3; ++s) {
        // python: ndepth / 2 * (ratio * interval), all in double, then cast to the tensor's fp32
        const double pix = cs.ratio[s] * interval;
        half_range[s] = (float)(((double)cs.ndepth[s] / 2.0) * pix);
    }
}

__global__ void resize_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int hi, int wi,
                              int ho, int wo, float sy, float sx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    if (x >= wo) return;
    const Lerp ly = lerp_index(y, sy, hi), lx = lerp_index(x, sx, wi);
    const float* p = in + (size_t)n * hi * wi;
    const float v00 = __ldg(p + ly.i0 * wi + lx.i0), v01 = __ldg(p + ly.i0 * wi + lx.i1);
    const float v10 = __ldg(p + ly.i1 * wi + lx.i0), v11 = __ldg(p + ly.i1 * wi + lx.i1);
    out[((size_t)n * ho + y) * wo + x] = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
}

}  // namespace adamvs

using namespace adamvs;

extern "C" int adamvs_abi_version(void) { return ADAMVS_ABI_VERSION; }

extern "C" int adamvs_cascade_prepare(const float* proj_s1, const float* proj_s2, const float* proj_s3,
                                      const float* depth_values, int ncol, int B, int V,
                                      int interval_mode, int num_depth,
                                      const int* host_ndepths, const double* host_ratios,
                                      float* relproj, float* half_range, void* stream) {
    ADAMVS_CHECK_ARG(proj_s1 && proj_s2 && proj_s3 && depth_values && relproj && half_range);
    ADAMVS_CHECK_ARG(host_ndepths && host_ratios && B > 0 && V >= 2 && ncol >= 2);
    ADAMVS_CHECK_ARG(interval_mode == ADAMVS_INTERVAL_LAST_COLUMN || (interval_mode == ADAMVS_INTERVAL_FROM_RANGE && num_depth > 0));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = 3 * B * (V - 1);
    relproj_kernel<<<(n + 63) / 64, 64, 0, st>>>(proj_s1, proj_s2, proj_s3, B, V, relproj);
    CascadeScalars cs;
    for (int s = 0; s < 3; ++s) { cs.ndepth[s] = host_ndepths[s]; cs.ratio[s] = host_ratios[s]; }
    half_range_kernel<<<1, 32, 0, st>>>(depth_values, ncol, interval_mode, num_depth, cs, half_range);
    ADAMVS_LAUNCH_RESULT();
}

extern "C" int adamvs_resize_bilinear_f32(const float* in, float* out, int N, int hi, int wi, int ho, int wo,
                                          void* stream) {
    ADAMVS_CHECK_ARG(in && out && N > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && ho <= 65535 && N <= 65535);
    dim3 grid((wo + 127) / 128, ho, N);
    // ATen: scale = (float)in / out  (align_corners=False, no user scale factor)
    resize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(in, out, N, hi, wi, ho, wo, (float)hi / (float)ho,
                                                          (float)wi / (float)wo);
    ADAMVS_LAUNCH_RESULT();
}
