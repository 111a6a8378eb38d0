// K4 standalone: softmax over D + expectation + max for a materialised logit volume [N,D,h,w]
// (the stage-1 pair branch: F.softmax / depth_regression / max at models/adamvs.py:274-283, 481-489;
// models/module.py:617-625).  D is the slow dimension of the layout, so the reduction is a per-thread
// loop over D with lanes along x (every load is a coalesced 128-byte row segment); a warp-shuffle
// reduction would only apply if D were the fastest dimension.  Two passes over the logits (max, then
// sum) — the second hits L2.
#include "common.cuh"

namespace adamvs {

__global__ void __launch_bounds__(128)
softmax_regress_kernel(const float* __restrict__ logits, HypSpec hs, int prob_mode, float* __restrict__ depth,
                       float* __restrict__ conf, int n_per_batch, int D, int h, int w) {
    const int hw = h * w;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (pix >= hw) return;
    const float* p = logits + (size_t)n * D * hw + pix;
    const HypLine line = hyp_line(hs, n / n_per_batch, pix, hw, D);
    if (prob_mode == ADAMVS_PROB_SOFTMAX) {
        float m = -INFINITY;
        for (int k = 0; k < D; ++k) m = fmaxf(m, __ldg(p + (size_t)k * hw));
        float s = 0.f;
        for (int k = 0; k < D; ++k) s += expf(__ldg(p + (size_t)k * hw) - m);
        // reference: p_k = e_k / s (softmax), depth = sum_k p_k * d_k, conf = max_k p_k
        float dsum = 0.f, pmax = 0.f;
        for (int k = 0; k < D; ++k) {
            const float pk = expf(__ldg(p + (size_t)k * hw) - m) / s;
            dsum += pk * hyp_at(line, k);
            pmax = fmaxf(pmax, pk);
        }
        depth[(size_t)n * hw + pix] = dsum;
        conf[(size_t)n * hw + pix] = pmax;
    } else {
        float emax = 0.f, esum = 0.f, dsum = 0.f;
        for (int k = 0; k < D; ++k) {
            const float e = expf(__ldg(p + (size_t)k * hw));
            emax = (emax < e) ? e : emax;
            dsum = hyp_at(line, k) * e + dsum;
            esum += e;
        }
        const float den = esum + 1e-10f;
        depth[(size_t)n * hw + pix] = dsum / den;
        conf[(size_t)n * hw + pix] = emax / den;
    }
}

}  // namespace adamvs

using namespace adamvs;

extern "C" int adamvs_softmax_regress_f32(const float* logits,
                                          int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                          int prob_mode, float* depth, float* conf,
                                          int N, int n_per_batch, int D, int h, int w, void* stream) {
    ADAMVS_CHECK_ARG(logits && depth && conf && hyp_src && N > 0 && N <= 65535 && n_per_batch > 0 && D >= 2 && h > 0 && w > 0);
    ADAMVS_CHECK_ARG(prob_mode == ADAMVS_PROB_SOFTMAX || prob_mode == ADAMVS_PROB_EXP_EPS);
    ADAMVS_CHECK_ARG(hyp_mode == ADAMVS_HYP_PLANES ? hyp_ncol >= 2 : (hyp_mode == ADAMVS_HYP_PER_PIXEL && half_range));
    const HypSpec hs{hyp_mode, hyp_src, hyp_ncol, half_range};
    dim3 grid((h * w + 127) / 128, N, 1);
    softmax_regress_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(logits, hs, prob_mode, depth, conf, n_per_batch, D, h, w);
    ADAMVS_LAUNCH_RESULT();
}
