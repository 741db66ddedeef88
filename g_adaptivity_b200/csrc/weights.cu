// Dense per-layer projection, folded (derivation and device routines: tail_math.cuh).
//
// gad_prepare_weights: (Wq, bq, Wk) -> Mu = {M = c Wq^T Wk, u = c Wk^T bq}, c = log2(e)/(sqrt(C) T);
// gad_weight_grads: the chain rule from dL/d(M, u) back to the Linear parameters of
// src/GRAND_plus.py:146-147 (d/d lin_key.bias == 0: softmax shift invariance).
#include "common.cuh"
#include "tail_math.cuh"

namespace gad {
namespace {

__global__ void k_prepare_weights(const float* __restrict__ Wq, const float* __restrict__ bq,
                                  const float* __restrict__ Wk, int Lw, int C, int CE, float inv_temp,
                                  float* __restrict__ Mu) {
    tail::prepare_weights(Wq, bq, Wk, Lw, C, CE, tail::fold_scale(inv_temp, C), Mu);
}

__global__ void k_weight_grads(const float* __restrict__ Wq, const float* __restrict__ bq,
                               const float* __restrict__ Wk, const float* __restrict__ gMu, int Lw, int C, int CE,
                               float inv_temp, float* __restrict__ gWq, float* __restrict__ gbq,
                               float* __restrict__ gWk, float* __restrict__ gbk) {
    tail::weight_grads(Wq, bq, Wk, gMu, Lw, C, CE, tail::fold_scale(inv_temp, C), gWq, gbq, gWk, gbk);
}

}  // namespace
}  // namespace gad

using namespace gad;

extern "C" int gad_prepare_weights(const float* Wq, const float* bq, const float* Wk, int Lw, int C, int CE,
                                   float inv_temp, float* Mu, void* stream) {
    GAD_CHECK_ARG(Wq && bq && Wk && Mu, "gad_prepare_weights: null pointer");
    GAD_CHECK_ARG(Lw > 0 && C > 0 && (CE == 2 || CE == 4 || CE == 8), "gad_prepare_weights: Lw=%d C=%d CE=%d", Lw, C, CE);
    k_prepare_weights<<<1, 128, 0, as_stream(stream)>>>(Wq, bq, Wk, Lw, C, CE, inv_temp, Mu);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

extern "C" int gad_weight_grads(const float* Wq, const float* bq, const float* Wk, const float* gMu, int Lw,
                                int C, int CE, float inv_temp, float* gWq, float* gbq, float* gWk, float* gbk,
                                void* stream) {
    GAD_CHECK_ARG(Wq && bq && Wk && gMu && gWq && gbq && gWk && gbk, "gad_weight_grads: null pointer");
    GAD_CHECK_ARG(Lw > 0 && C > 0 && (CE == 2 || CE == 4 || CE == 8), "gad_weight_grads: Lw=%d C=%d CE=%d", Lw, C, CE);
    k_weight_grads<<<1, 256, 0, as_stream(stream)>>>(Wq, bq, Wk, gMu, Lw, C, CE, inv_temp, gWq, gbq, gWk, gbk);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}
