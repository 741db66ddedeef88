// Adam on the flat parameter vector of the deformer (144 + L floats): `torch.optim.Adam`
// semantics of the reference training loop (src/run_GNN.py:88,128,131), as one single-CTA kernel
// so that the whole training step is CUDA-graph replayable with no host involvement.
#include "common.cuh"
#include "tail_math.cuh"

namespace gad {
namespace {

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g,
                                              float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                              float b1, float b2, float eps, float wd, float gscale,
                                              int64_t* __restrict__ step) {
    const int64_t t = step[0] + 1;
    tail::adam(p, g, m, v, n, lr, b1, b2, eps, wd, gscale, t);
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
