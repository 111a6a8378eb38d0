// adamvs_conv3x3_f32 — the 3x3 convolutions of the two 2-D networks around the cost-volume path (FeatureNet0,
// reference models/adamvs.py:57-104, and the stage-1 pair U-Net CostRegNet2D, :198-238), run on the same
// persistent TMA-fed FFMA kernels as the recurrent regulariser instead of cuDNN's fp32 paths: eval-mode
// BatchNorm is folded into weights and bias by the caller, ReLU is fused, and a channel concatenation in front
// of the conv (DeConv2dFuse, module.py:506-524) is read as two tensors instead of being materialised.
#include "conv3x3.cuh"

namespace adamvs {

template <int CA, int CB, int COUT, int STRIDE>
static int run_conv(const ConvArgs& a, int N, cudaStream_t st) {
    constexpr int COB = COUT % 16 == 0 ? 16 : 8;
    using L = ConvLayer<CA, CB, COUT, COB, STRIDE, EPI_BIAS>;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.inA) | reinterpret_cast<uintptr_t>(a.inB) |
                           reinterpret_cast<uintptr_t>(a.out0)) % 16) == 0;
    if (aligned && a.win % 4 == 0 && a.wout % 4 == 0) {
        ConvPlan p;
        if (L::plan(p, a, N, 1)) {
            cudaError_t e = L::launch(p, N, st);
            return e == cudaSuccess ? 0 : (int)e;
        }
    }
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

}  // namespace adamvs

using namespace adamvs;

extern "C" int adamvs_conv3x3_supported(int CA, int CB, int COUT, int stride) {
    if (stride == 1) {
        if (CB == 0) return (CA == 8 && COUT == 8) || (CA == 16 && COUT == 16) || (CA == 32 && COUT == 32) || (CA == 48 && COUT == 48);
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
    if (stride == 2) return run_conv<48, 0, 48, 2>(a, N, st);
    if (CB == 16) return run_conv<16, 16, 16, 1>(a, N, st);
    if (CB == 8) return run_conv<8, 8, 8, 1>(a, N, st);
    switch (CA) {
        case 8: return run_conv<8, 0, 8, 1>(a, N, st);
        case 16: return run_conv<16, 0, 16, 1>(a, N, st);
        case 32: return run_conv<32, 0, 32, 1>(a, N, st);
        default: return run_conv<48, 0, 48, 1>(a, N, st);
    }
}
