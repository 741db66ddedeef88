// Per-node arithmetic of one attention-diffusion F-evaluation and of its backward, shared by the
// streaming kernels (state in global memory / L2) and the mesh-resident kernels (state in shared
// memory).  Everything is templated on CE, the number of live channels (2, 4, 8), so a node row is
// one or two 128-bit (or one 64-bit) accesses.
//
// Forward, per destination node i (src/GRAND_plus.py:225-343 with value = identity):
//     p_i   = M^T x_i + u                          (folded q/k projection, see weights.cu)
//     s_e   = <p_i, x_j>                           for every in-edge e = (j -> i), CSR row order
//     m_i   = max_e s_e,  w_e = exp(s_e - m_i),  Z_i = sum_e w_e
//     o_i   = (sum_e w_e x_j) / Z_i                (= sum_e alpha_e x_j, alpha row-stochastic)
//     k_i   = o_i - x_i                            (GRAND_plus.py:267)
// Z_i >= 1 in fp32, so the reference's "+1e-16" (PyG softmax) is a no-op and is dropped.
// A node without in-edges gets o_i = 0 (empty scatter row), exactly like the reference.
//
// Backward (re-derived for the folded form), with g_o = b * g_plus:
//     D_i = <g_o, o_i>;  da_e = <g_o, x_j>;  ds_e = alpha_e (da_e - D_i);  t_i = sum_e ds_e x_j
//     G_M += x_i (x) t_i;  G_u += t_i;  g_x[i] += M t_i                      (destination pass)
//     g_x[j] += alpha_e g_o[i] + ds_e p_i                                    (source pass)
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace gad {

template <int CE>
struct Row {
    float v[CE];
};

template <int CE>
__device__ __forceinline__ Row<CE> load_row(const float* __restrict__ base, int64_t i) {
    Row<CE> r;
    if constexpr (CE == 2) {
        const float2 t = *reinterpret_cast<const float2*>(base + i * 2);
        r.v[0] = t.x;
        r.v[1] = t.y;
    } else {
#pragma unroll
        for (int q = 0; q < CE / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(base + i * CE + 4 * q);
            r.v[4 * q + 0] = t.x;
            r.v[4 * q + 1] = t.y;
            r.v[4 * q + 2] = t.z;
            r.v[4 * q + 3] = t.w;
        }
    }
    return r;
}

template <int CE>
__device__ __forceinline__ void store_row(float* __restrict__ base, int64_t i, const Row<CE>& r) {
    if constexpr (CE == 2) {
        *reinterpret_cast<float2*>(base + i * 2) = make_float2(r.v[0], r.v[1]);
    } else {
#pragma unroll
        for (int q = 0; q < CE / 4; ++q)
            *reinterpret_cast<float4*>(base + i * CE + 4 * q) =
                make_float4(r.v[4 * q], r.v[4 * q + 1], r.v[4 * q + 2], r.v[4 * q + 3]);
    }
}

template <int CE>
__device__ __forceinline__ float dot(const Row<CE>& a, const Row<CE>& b) {
    float s = a.v[0] * b.v[0];
#pragma unroll
    for (int c = 1; c < CE; ++c) s = fmaf(a.v[c], b.v[c], s);
    return s;
}

// p = M^T x + u   (M row-major [a][b], Mu = {M, u})
template <int CE>
__device__ __forceinline__ Row<CE> project(const float* __restrict__ Mu, const Row<CE>& x) {
    Row<CE> p;
#pragma unroll
    for (int b = 0; b < CE; ++b) p.v[b] = Mu[CE * CE + b];
#pragma unroll
    for (int a = 0; a < CE; ++a)
#pragma unroll
        for (int b = 0; b < CE; ++b) p.v[b] = fmaf(x.v[a], Mu[a * CE + b], p.v[b]);
    return p;
}

// y = M t
template <int CE>
__device__ __forceinline__ Row<CE> apply_M(const float* __restrict__ Mu, const Row<CE>& t) {
    Row<CE> y;
#pragma unroll
    for (int a = 0; a < CE; ++a) {
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < CE; ++b) s = fmaf(Mu[a * CE + b], t.v[b], s);
        y.v[a] = s;
    }
    return y;
}

struct SoftmaxStats {
    float m;    // row max of the logits (0 for an empty row)
    float rZ;   // 1 / sum exp(s - m)    (0 for an empty row)
};

// One F-evaluation at node i.  X: state rows (global or shared), indexed by the entries of `col`
// (already relative to X).  Returns k_i = o_i - x_i; o_i and the softmax stats on request.
template <int CE, typename ColT>
__device__ __forceinline__ Row<CE> node_feval(const float* __restrict__ X, const ColT* __restrict__ col,
                                              int e_begin, int e_end, const Row<CE>& xi,
                                              const float* __restrict__ Mu, Row<CE>* o_out = nullptr,
                                              SoftmaxStats* st_out = nullptr, Row<CE>* p_out = nullptr) {
    const Row<CE> p = project<CE>(Mu, xi);
    float m = -CUDART_INF_F;
    for (int e = e_begin; e < e_end; ++e) {
        const Row<CE> xj = load_row<CE>(X, (int64_t)col[e]);
        m = fmaxf(m, dot<CE>(p, xj));
    }
    float Z = 0.f;
    Row<CE> o;
#pragma unroll
    for (int c = 0; c < CE; ++c) o.v[c] = 0.f;
    for (int e = e_begin; e < e_end; ++e) {
        const Row<CE> xj = load_row<CE>(X, (int64_t)col[e]);
        const float w = expf(dot<CE>(p, xj) - m);
        Z += w;
#pragma unroll
        for (int c = 0; c < CE; ++c) o.v[c] = fmaf(w, xj.v[c], o.v[c]);
    }
    const float rZ = (e_end > e_begin) ? 1.0f / Z : 0.f;
    Row<CE> k;
#pragma unroll
    for (int c = 0; c < CE; ++c) {
        o.v[c] *= rZ;
        k.v[c] = o.v[c] - xi.v[c];
    }
    if (o_out) *o_out = o;
    if (st_out) {
        st_out->m = (e_end > e_begin) ? m : 0.f;
        st_out->rZ = rZ;
    }
    if (p_out) *p_out = p;
    return k;
}

// Per-destination record the source pass needs: p_i, D_i and the log-sum-exp of row i.
template <int CE>
struct DstRec {
    Row<CE> p;
    float D;
    float lse;
};

// Destination pass of the backward at node i.
//   gplus = dL/dx^{l+1}_i,  a, b: x^{l+1} = a x + b (o - x)  (Euler: a = 1, b = tau)
// Returns the part of dL/dx^l_i that is local to row i:  (a - b) gplus + M t_i.
// Accumulates G_M (x) / G_u into acc[CE*CE + CE] and  <gplus, o - x>  into *gb_acc.
template <int CE, typename ColT>
__device__ __forceinline__ Row<CE> node_bwd_dst(const float* __restrict__ X, const ColT* __restrict__ col,
                                                int e_begin, int e_end, const Row<CE>& xi,
                                                const Row<CE>& gplus, float a, float b,
                                                const float* __restrict__ Mu, DstRec<CE>* rec,
                                                float* __restrict__ acc, float* gb_acc) {
    Row<CE> o, p;
    SoftmaxStats st;
    const Row<CE> k = node_feval<CE, ColT>(X, col, e_begin, e_end, xi, Mu, &o, &st, &p);
    Row<CE> go;
#pragma unroll
    for (int c = 0; c < CE; ++c) go.v[c] = b * gplus.v[c];
    const float D = dot<CE>(go, o);
    *gb_acc += dot<CE>(gplus, k);
    Row<CE> t;
#pragma unroll
    for (int c = 0; c < CE; ++c) t.v[c] = 0.f;
    for (int e = e_begin; e < e_end; ++e) {
        const Row<CE> xj = load_row<CE>(X, (int64_t)col[e]);
        const float alpha = expf(dot<CE>(p, xj) - st.m) * st.rZ;
        const float ds = alpha * (dot<CE>(go, xj) - D);
#pragma unroll
        for (int c = 0; c < CE; ++c) t.v[c] = fmaf(ds, xj.v[c], t.v[c]);
    }
#pragma unroll
    for (int aa = 0; aa < CE; ++aa)
#pragma unroll
        for (int bb = 0; bb < CE; ++bb) acc[aa * CE + bb] = fmaf(xi.v[aa], t.v[bb], acc[aa * CE + bb]);
#pragma unroll
    for (int bb = 0; bb < CE; ++bb) acc[CE * CE + bb] += t.v[bb];
    rec->p = p;
    rec->D = D;
    // alpha_e = exp(s_e - lse);  lse = m + log Z.  Empty row: never read by a source pass.
    rec->lse = (e_end > e_begin) ? st.m - logf(st.rZ) : 0.f;
    const Row<CE> Mt = apply_M<CE>(Mu, t);
    Row<CE> gs;
#pragma unroll
    for (int c = 0; c < CE; ++c) gs.v[c] = fmaf(a - b, gplus.v[c], Mt.v[c]);
    return gs;
}

// Source pass of the backward at node j: contributions of all out-edges (j -> i).
//   P: rows p_i;  DL: float2 rows (D_i, lse_i);  G: rows gplus_i;  tdst: destinations, CSC order.
template <int CE, typename ColT>
__device__ __forceinline__ Row<CE> node_bwd_src(const float* __restrict__ P, const float2* __restrict__ DL,
                                                const float* __restrict__ G, const ColT* __restrict__ tdst,
                                                int e_begin, int e_end, const Row<CE>& xj, float b) {
    Row<CE> accv;
#pragma unroll
    for (int c = 0; c < CE; ++c) accv.v[c] = 0.f;
    for (int e = e_begin; e < e_end; ++e) {
        const int64_t i = (int64_t)tdst[e];
        const Row<CE> p = load_row<CE>(P, i);
        const Row<CE> gp = load_row<CE>(G, i);
        const float2 dl = DL[i];
        const float alpha = expf(dot<CE>(p, xj) - dl.y);
        const float da = b * dot<CE>(gp, xj);
        const float ds = alpha * (da - dl.x);
        const float ab = alpha * b;
#pragma unroll
        for (int c = 0; c < CE; ++c) accv.v[c] = fmaf(ab, gp.v[c], fmaf(ds, p.v[c], accv.v[c]));
    }
    return accv;
}

// Deterministic block-wide sum of `n` per-thread accumulators into out[0..n) (thread 0..n-1
// hold the result in smem `red` on return; caller copies).  red must hold n * (blockDim/32) floats.
template <int NACC>
__device__ __forceinline__ void block_reduce(float (&acc)[NACC], float* __restrict__ red,
                                             float* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        float v = acc[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) red[k * nwarp + warp] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NACC; k += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += red[k * nwarp + w];
        out[k] = s;
    }
    __syncthreads();
}

}  // namespace gad
