// CTA-wide device routines for the small dense steps around the deformer kernels: the folded
// projection (M, u), its chain rule back to the Linear parameters, the fixed-order reduction of the
// per-tile partials and Adam.  Each is callable from a stand-alone single-CTA kernel (weights.cu,
// optim.cu, ell_api.cu) or from the tail of the one-launch training kernel (ell_kernels.cuh), where
// the last CTA to finish runs them back to back so that a whole training step is one launch.
//
// The reference computes q = x Wq^T + bq and k = x Wk^T + bk for every node and layer
// (src/GRAND_plus.py:225-226, two [N,C]x[C,C] addmm calls) and then <q_i, k_j>/sqrt(C) per edge
// (:279).  Only min(in_dim, C) input channels are ever non-zero (identity encoder,
// src/GNN.py:75-83), and every term of <q_i, k_j> that does not depend on j cancels in the
// segment softmax (:333).  What remains is the bilinear form
//     s_e = x_i^T M x_j + u^T x_j,   M = c Wq^T Wk  (CE x CE),  u = c Wk^T bq,  c = log2(e)/(sqrt(C) T)
// (log2 domain: the kernels exponentiate with ex2), so the "projection GEMM" shrinks to one
// CE x CE x C product per weight set and step -- a few hundred FMAs, done in fp64 by one CTA.  That
// is why no tcgen05 tile is issued for it: a 128 x 8 x 8 tf32 MMA would need a 3xTF32 split for the
// 1e-5 parity bar (SURVEY hazard 10) plus a TMEM round trip per 32 B of node state, for work that
// no longer exists per node.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gad {
namespace tail {

constexpr double LOG2E = 1.4426950408889634074;

// Mu[l] = { M (CE x CE, row-major M[a][b]), u (CE) } for l < Lw.
__host__ __device__ __forceinline__ double fold_scale(float inv_temp, int C) {
    return LOG2E * (double)inv_temp / sqrt((double)C);   // logits live in the log2 domain
}

__device__ __forceinline__ void prepare_weights(const float* Wq, const float* bq,
                                                const float* Wk, int Lw, int C, int CE, double c,
                                                float* Mu) {
    const int live = CE < C ? CE : C;
    const int musz = CE * CE + CE;
    for (int idx = threadIdx.x; idx < Lw * musz; idx += blockDim.x) {
        const int l = idx / musz, r = idx % musz;
        const float* wq = Wq + (size_t)l * C * C;
        const float* wk = Wk + (size_t)l * C * C;
        const float* b = bq + (size_t)l * C;
        double acc = 0.0;
        if (r < CE * CE) {
            const int a = r / CE, bcol = r % CE;
            if (a < live && bcol < live)
                for (int o = 0; o < C; ++o) acc += (double)wq[o * C + a] * (double)wk[o * C + bcol];
        } else {
            const int bcol = r - CE * CE;
            if (bcol < live)
                for (int o = 0; o < C; ++o) acc += (double)b[o] * (double)wk[o * C + bcol];
        }
        Mu[idx] = (float)(c * acc);
    }
}

// dWq[o,a] = c sum_b Wk[o,b] G_M[a,b];  dWk[o,b] = c (sum_a Wq[o,a] G_M[a,b] + bq[o] G_u[b]);
// dbq[o] = c sum_b Wk[o,b] G_u[b];  dbk = 0.   (gMu = dL/d(M, u) in the log2 domain: same c.)
__device__ __forceinline__ void weight_grads(const float* Wq, const float* bq,
                                             const float* Wk, const float* gMu, int Lw, int C, int CE,
                                             double c, float* gWq, float* gbq,
                                             float* gWk, float* gbk) {
    const int live = CE < C ? CE : C;
    const int musz = CE * CE + CE;
    for (int idx = threadIdx.x; idx < Lw * C * C; idx += blockDim.x) {
        const int l = idx / (C * C), r = idx % (C * C);
        const int o = r / C, a = r % C;
        const float* wq = Wq + (size_t)l * C * C;
        const float* wk = Wk + (size_t)l * C * C;
        const float* b = bq + (size_t)l * C;
        const float* GM = gMu + (size_t)l * musz;
        const float* Gu = GM + CE * CE;
        double dq = 0.0, dk = 0.0;
        if (a < live) {
            for (int bb = 0; bb < live; ++bb) dq += (double)wk[o * C + bb] * (double)GM[a * CE + bb];
            for (int aa = 0; aa < live; ++aa) dk += (double)wq[o * C + aa] * (double)GM[aa * CE + a];
            dk += (double)b[o] * (double)Gu[a];
        }
        gWq[idx] = (float)(c * dq);
        gWk[idx] = (float)(c * dk);
    }
    for (int idx = threadIdx.x; idx < Lw * C; idx += blockDim.x) {
        const int l = idx / C, o = idx % C;
        const float* wk = Wk + (size_t)l * C * C;
        const float* Gu = gMu + (size_t)l * musz + CE * CE;
        double d = 0.0;
        for (int bb = 0; bb < live; ++bb) d += (double)wk[o * C + bb] * (double)Gu[bb];
        gbq[idx] = (float)(c * d);
        gbk[idx] = 0.0f;
    }
}

// b^t for integer t >= 0 by repeated squaring (exact to a few ulp; ~2 log2(t) multiplies instead of
// a double-precision pow() call on the serial tail of the training kernel).
__device__ __forceinline__ double ipow(double b, long long t) {
    double r = 1.0;
    while (t > 0) {
        if (t & 1) r *= b;
        b *= b;
        t >>= 1;
    }
    return r;
}

// torch.optim.Adam semantics (run_GNN.py:88,128,131): L2 weight decay added to the gradient,
// bias-corrected moments.
struct AdamCoef {
    float step_size, bc2_sqrt, b1, b2, eps, wd, gscale;
};

__device__ __forceinline__ AdamCoef adam_coef(float lr, float b1, float b2, float eps, float wd, float gscale,
                                              long long t) {
    const double bc1 = 1.0 - ipow((double)b1, t);
    const double bc2 = 1.0 - ipow((double)b2, t);
    AdamCoef c;
    c.step_size = (float)((double)lr / bc1);
    c.bc2_sqrt = (float)sqrt(bc2);
    c.b1 = b1;
    c.b2 = b2;
    c.eps = eps;
    c.wd = wd;
    c.gscale = gscale;
    return c;
}

// one element; returns the new parameter, updates (m, v) in place
__device__ __forceinline__ float adam_update(const AdamCoef& c, float pi, float g, float& m, float& v) {
    float gi = g * c.gscale;
    if (c.wd != 0.f) gi = fmaf(c.wd, pi, gi);
    const float mi = m + (1.f - c.b1) * (gi - m);          // lerp, as torch does
    const float vi = c.b2 * v + (1.f - c.b2) * gi * gi;
    m = mi;
    v = vi;
    const float denom = sqrtf(vi) / c.bc2_sqrt + c.eps;
    return pi - c.step_size * (mi / denom);
}

__device__ __forceinline__ void adam_one(const AdamCoef& c, float* p, float g, float* m, float* v, long long i) {
    float mi = m[i], vi = v[i];
    p[i] = adam_update(c, p[i], g, mi, vi);
    m[i] = mi;
    v[i] = vi;
}

// `t` = step number of THIS update (1-based); the caller stores it back to the device counter.
__device__ __forceinline__ void adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1,
                                     float b2, float eps, float wd, float gscale, int64_t t) {
    const AdamCoef c = adam_coef(lr, b1, b2, eps, wd, gscale, t);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) adam_one(c, p, g[i], m, v, i);
}

// Fixed-order sum of the per-tile partials, one warp per output column (lanes stride over tiles,
// then a butterfly): columns [0, slots*nacc) are (G_M, G_u[, g_tau]) per weight slot, then L
// step-size columns (shared weights) and one loss column.  Reads bypass L1 (written by other CTAs).
__device__ __forceinline__ void reduce_partials(const float* partials, int T, int slots, int nacc, int musz,
                                                float* gMu, const float* tau_partials, int L,
                                                float* g_tau, const float* loss_partials,
                                                float loss_scale, float* loss, int warp, int nwarps) {
    const int lane = threadIdx.x & 31;
    const int ncol = slots * nacc;
    for (int w = warp; w < ncol + L + 1; w += nwarps) {
        const float* src;
        int stride;
        if (w < ncol) {
            src = partials + w;
            stride = ncol;
        } else if (w < ncol + L) {
            if (!(tau_partials && g_tau)) continue;
            src = tau_partials + (w - ncol);
            stride = L;
        } else {
            if (!(loss_partials && loss)) continue;
            src = loss_partials;
            stride = 1;
        }
        double sd = 0.0;   // fp64: the per-tile partials cancel heavily across meshes
        for (int t = lane; t < T; t += 32) sd += (double)__ldcg(src + (size_t)t * stride);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, d);
        if (lane != 0) continue;
        const float s = (float)sd;
        if (w < ncol) {
            const int l = w / nacc, a = w % nacc;
            if (a < musz) gMu[(size_t)l * musz + a] = s;
            else if (g_tau && slots > 1) g_tau[l] = s;
        } else if (w < ncol + L) {
            g_tau[w - ncol] = s;
        } else {
            loss[0] = s * loss_scale;
        }
    }
}


// ---- pieces of the single-pass tail of the training kernel (ell_kernels.cuh: train_tail) -------
// Asynchronous coalesced copy of n floats global -> shared (cp.async.cg: 16-byte lines straight
// from L2, no register round trip), so that SEVERAL arrays can be in flight at once and the whole
// staging costs one L2 latency; finish with stage_wait() + __syncthreads().  The (< 4 float)
// remainder and unaligned arrays go through registers with ld.global.cg.
// (The tail runs once per launch on cold instruction-cache lines, so its cost is dominated by code
// size: one out-of-line copy, loops not unrolled.)
static __device__ __noinline__ void stage_async(float* __restrict__ dst, const float* __restrict__ src, int n) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const int n4 = n >> 2;
        const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
        const size_t g0 = __cvta_generic_to_global(src);
#pragma unroll 1
        for (int i = tid; i < n4; i += nthr)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * (uint32_t)i), "l"(g0 + 16ull * (size_t)i)
                         : "memory");
#pragma unroll 1
        for (int i = (n4 << 2) + tid; i < n; i += nthr) dst[i] = __ldcg(src + i);
    } else {
#pragma unroll 1
        for (int i = tid; i < n; i += nthr) dst[i] = __ldcg(src + i);
    }
}
__device__ __forceinline__ void stage_wait() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Chain rule of the fold for ONE entry (row lo = l * C + o, column a < live) of Wq / Wk: same
// arithmetic, same order as weight_grads above, with the live-channel loops unrolled so that every
// operand load is issued up front.  wq / wk point at the row, bo = bq[lo], GM at the layer's (G_M, G_u).
template <int CE>
__device__ __forceinline__ void weight_grad_entry(const float* wq, const float* wk, float bo, const float* GM,
                                                  int live, int a, double c, float& dWq, float& dWk) {
    float wkv[CE], wqv[CE], gq[CE], gk[CE];
#pragma unroll
    for (int bb = 0; bb < CE; ++bb) {
        const bool on = bb < live;
        wkv[bb] = on ? wk[bb] : 0.f;
        wqv[bb] = on ? wq[bb] : 0.f;
        gq[bb] = on ? GM[a * CE + bb] : 0.f;
        gk[bb] = on ? GM[bb * CE + a] : 0.f;
    }
    const float gu = GM[CE * CE + a];
    double dq = 0.0, dk = 0.0;
#pragma unroll
    for (int bb = 0; bb < CE; ++bb)
        if (bb < live) dq += (double)wkv[bb] * (double)gq[bb];
#pragma unroll
    for (int aa = 0; aa < CE; ++aa)
        if (aa < live) dk += (double)wqv[aa] * (double)gk[aa];
    dk += (double)bo * (double)gu;
    dWq = (float)(c * dq);
    dWk = (float)(c * dk);
}

// prepare_weights with CE known at compile time and a rolled reduction loop (small code).
template <int CE>
__device__ __forceinline__ void prepare_weights_t(const float* Wq, const float* bq, const float* Wk, int Lw, int C,
                                                  double c, float* Mu) {
    constexpr int MUSZ = CE * CE + CE;
    const int live = CE < C ? CE : C;
#pragma unroll 1
    for (int idx = threadIdx.x; idx < Lw * MUSZ; idx += blockDim.x) {
        const int l = idx / MUSZ, r = idx % MUSZ;
        const bool isM = r < CE * CE;
        const int acol = isM ? r / CE : 0, bcol = isM ? r % CE : r - CE * CE;
        const float* wk = Wk + (size_t)l * C * C + bcol;
        const float* lhs = isM ? Wq + (size_t)l * C * C + acol : bq + (size_t)l * C;   // Wq[o, a] or bq[o]
        const int ls = isM ? C : 1;
        double acc = 0.0;
        if (acol < live && bcol < live) {
#pragma unroll 2
            for (int o = 0; o < C; ++o) acc += (double)lhs[(size_t)o * ls] * (double)wk[(size_t)o * C];
        }
        Mu[idx] = (float)(c * acc);
    }
}

}  // namespace tail
}  // namespace gad
