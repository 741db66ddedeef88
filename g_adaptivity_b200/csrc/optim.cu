// Adam on the flat parameter vector of the deformer (144 + L floats): `torch.optim.Adam`
// semantics of the reference training loop (src/run_GNN.py:88,128,131), as one single-CTA kernel
// so that the whole training step is CUDA-graph replayable with no host involvement.
#include "common.cuh"

namespace gad {
namespace {

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g,
                                              float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                              float b1, float b2, float eps, float wd, float gscale,
                                              int64_t* __restrict__ step) {
    const int64_t t = step[0] + 1;
    const double bc1 = 1.0 - pow((double)b1, (double)t);
    const double bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float gi = g[i] * gscale;
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = m[i] + (1.f - b1) * (gi - m[i]);          // lerp, as torch does
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
    }
    __syncthreads();
    if (threadIdx.x == 0) step[0] = t;
}

}  // namespace
}  // namespace gad

using namespace gad;

extern "C" int gad_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                             float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                             int64_t* step, void* stream) {
    GAD_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step && n > 0, "gad_adam_step: bad arguments");
    k_adam<<<1, 256, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                             weight_decay, grad_scale, step);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}
