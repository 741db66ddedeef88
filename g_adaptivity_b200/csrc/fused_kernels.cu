// Mesh-resident kernels: the whole multi-layer integration of one tile in ONE launch.
//
// A batch is a disjoint union of meshes (PyG Batch semantics), so a contiguous node
// range that does not split a mesh ("tile", planned on the host) is closed under the edge
// relation.  One CTA owns one tile: it stages the tile's node state and its CSR/CSC slices in
// shared memory once, runs all L layers (Euler updates or RK4 stages, src/GNN.py:273-296) with
// `__syncthreads()` between F-evaluations, and touches HBM only for the inputs, the saved layer
// states (training) and the outputs.  Every thread owns up to NPT nodes for the whole kernel, so
// per-node quantities that cross a barrier (new state, RK4 base/accumulator, the row-local part
// of the gradient) live in registers and the state buffer is updated in place -- one state
// buffer instead of a ping-pong pair.
//
// Shared memory per tile (CE = 4): forward 16 B/node + 2 B/edge + 4 B/node;
// backward 56 B/node (x, g, p, {D,lse}) + 4 B/edge + 8 B/node.  30x30 meshes: 31 KB / 78 KB,
// 50x50 meshes: 79 KB / 217 KB (B200: 227 KB per CTA).
//
// Weight gradients are accumulated per thread in registers, reduced per CTA with a fixed-order
// butterfly and written as per-tile partials; a second tiny kernel sums the partials in a fixed
// order.  No floating-point atomics anywhere: results are bit-reproducible run to run.
#include <mutex>

#include "common.cuh"
#include "node_math.cuh"

namespace gad {
namespace {

typedef uint16_t lidx_t;  // tile-local node index (tiles hold < 65536 nodes)

#ifndef GAD_FWD_MINB
#define GAD_FWD_MINB 2   // min resident CTAs per SM the kernels are compiled for (register cap)
#endif
#ifndef GAD_BWD_MINB
#define GAD_BWD_MINB 2
#endif

struct FwdSmem {
    size_t x_off, col_off, row_off, mu_off, total;
};

template <int CE>
__host__ __device__ inline FwdSmem fwd_layout(int cap_nodes, int cap_edges) {
    FwdSmem s;
    size_t o = 0;
    s.x_off = o;
    o += (size_t)cap_nodes * CE * sizeof(float);
    o = (o + 15) & ~size_t(15);
    s.row_off = o;
    o += ((size_t)cap_nodes + 1) * sizeof(int32_t);
    o = (o + 15) & ~size_t(15);
    s.mu_off = o;
    o += (CE * CE + CE) * sizeof(float);
    o = (o + 15) & ~size_t(15);
    s.col_off = o;
    o += (size_t)cap_edges * sizeof(lidx_t);
    s.total = (o + 15) & ~size_t(15);
    return s;
}

// ---- forward ------------------------------------------------------------------------------
// Largest CTA a kernel instantiated for NPT nodes per thread is ever launched with (pick_shape):
// bounds the register allocation (NPT = 4 -> 768 threads -> 85 registers per thread).
template <int NPT>
struct MaxThreads {
    static constexpr int value = NPT == 1 ? 256 : (NPT == 2 ? 512 : 768);
};
constexpr int MAX_FUSED_TILE_NODES = 4 * 768;

template <int CE, int NPT, int METHOD>
__global__ void __launch_bounds__(MaxThreads<NPT>::value, GAD_FWD_MINB) k_fused_fwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                    const int32_t* __restrict__ tile_ptr, int cap_nodes, int cap_edges,
                                                    const float* __restrict__ x0, int dim,
                                                    const float* __restrict__ Mu_g, int Lw,
                                                    const float* __restrict__ tau, int L, int64_t N,
                                                    float* __restrict__ x_phys, float* __restrict__ states) {
    extern __shared__ __align__(16) unsigned char smem[];
    const FwdSmem lay = fwd_layout<CE>(cap_nodes, cap_edges);
    float* X = reinterpret_cast<float*>(smem + lay.x_off);
    int32_t* rowS = reinterpret_cast<int32_t*>(smem + lay.row_off);
    float* Mu = reinterpret_cast<float*>(smem + lay.mu_off);
    lidx_t* colS = reinterpret_cast<lidx_t*>(smem + lay.col_off);
    constexpr int MUSZ = CE * CE + CE;

    const int tile = blockIdx.x;
    const int n0 = tile_ptr[tile], n1 = tile_ptr[tile + 1];
    const int NT = n1 - n0;
    const int e0 = rowptr[n0], ET = rowptr[n1] - e0;
    const int tid = threadIdx.x, nthr = blockDim.x;

    for (int e = tid; e < ET; e += nthr) colS[e] = (lidx_t)(col[e0 + e] - n0);
    for (int i = tid; i <= NT; i += nthr) rowS[i] = rowptr[n0 + i] - e0;
    for (int t = tid; t < MUSZ; t += nthr) Mu[t] = Mu_g[t];

    Row<CE> xi[NPT];
#pragma unroll
    for (int r = 0; r < NPT; ++r) {
        const int i = tid + r * nthr;
        if (i < NT) {
            xi[r] = load_row<CE>(x0, (int64_t)n0 + i);
            store_row<CE>(X, i, xi[r]);
        }
    }
    __syncthreads();

    const size_t state_stride = (size_t)N * CE;
    for (int l = 0; l < L; ++l) {
        if (Lw > 1 && l > 0) {
            for (int t = tid; t < MUSZ; t += nthr) Mu[t] = Mu_g[(size_t)l * MUSZ + t];
            __syncthreads();
        }
        const float h = tau[l];
        if constexpr (METHOD == GAD_METHOD_EULER) {
            Row<CE> xn[NPT];
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                const int i = tid + r * nthr;
                if (i < NT) {
                    const Row<CE> k = node_feval<CE, lidx_t>(X, colS, rowS[i], rowS[i + 1], xi[r], Mu);
#pragma unroll
                    for (int c = 0; c < CE; ++c) xn[r].v[c] = fmaf(h, k.v[c], xi[r].v[c]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                const int i = tid + r * nthr;
                if (i < NT) {
                    xi[r] = xn[r];
                    store_row<CE>(X, i, xn[r]);
                    if (states && l + 1 < L) store_row<CE>(states + (size_t)(l + 1) * state_stride, (int64_t)n0 + i, xn[r]);
                }
            }
            __syncthreads();
        } else {
            // classical RK4 on F(y) = A(y) y - y; base x and the k-accumulator stay in registers,
            // the stage input y lives in X.
            Row<CE> yb[NPT], acc[NPT];
            const float cin[4] = {0.5f * h, 0.5f * h, h, 0.f};
            const float wacc[4] = {1.f, 2.f, 2.f, 1.f};
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                yb[r] = xi[r];
#pragma unroll
                for (int c = 0; c < CE; ++c) acc[r].v[c] = 0.f;
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
#pragma unroll
                for (int r = 0; r < NPT; ++r) {
                    const int i = tid + r * nthr;
                    if (i < NT) {
                        const Row<CE> k = node_feval<CE, lidx_t>(X, colS, rowS[i], rowS[i + 1], yb[r], Mu);
#pragma unroll
                        for (int c = 0; c < CE; ++c) {
                            acc[r].v[c] = fmaf(wacc[s], k.v[c], acc[r].v[c]);
                            yb[r].v[c] = (s < 3) ? fmaf(cin[s], k.v[c], xi[r].v[c])
                                                 : fmaf(h * (1.0f / 6.0f), acc[r].v[c], xi[r].v[c]);
                        }
                    }
                }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < NPT; ++r) {
                    const int i = tid + r * nthr;
                    if (i < NT) store_row<CE>(X, i, yb[r]);
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                const int i = tid + r * nthr;
                if (i < NT) {
                    xi[r] = yb[r];
                    if (states && l + 1 < L) store_row<CE>(states + (size_t)(l + 1) * state_stride, (int64_t)n0 + i, yb[r]);
                }
            }
        }
    }
    // decoder = identity, x_phys = x[:, :dim]  (src/GNN.py:298-299)
#pragma unroll
    for (int r = 0; r < NPT; ++r) {
        const int i = tid + r * nthr;
        if (i < NT) {
            for (int d = 0; d < dim; ++d) x_phys[((int64_t)n0 + i) * dim + d] = xi[r].v[d];
        }
    }
}

// ---- backward -----------------------------------------------------------------------------
struct BwdSmem {
    size_t x_off, g_off, p_off, dl_off, row_off, trow_off, mu_off, red_off, col_off, tdst_off, total;
};

template <int CE>
__host__ __device__ inline BwdSmem bwd_layout(int cap_nodes, int cap_edges, int nwarps) {
    constexpr int NACC = CE * CE + CE + 1;
    BwdSmem s;
    size_t o = 0;
    auto bump = [&](size_t bytes) {
        size_t at = o;
        o = (o + bytes + 15) & ~size_t(15);
        return at;
    };
    s.x_off = bump((size_t)cap_nodes * CE * sizeof(float));
    s.g_off = bump((size_t)cap_nodes * CE * sizeof(float));
    s.p_off = bump((size_t)cap_nodes * CE * sizeof(float));
    s.dl_off = bump((size_t)cap_nodes * sizeof(float2));
    s.row_off = bump(((size_t)cap_nodes + 1) * sizeof(int32_t));
    s.trow_off = bump(((size_t)cap_nodes + 1) * sizeof(int32_t));
    s.mu_off = bump((CE * CE + CE) * sizeof(float));
    s.red_off = bump((size_t)NACC * nwarps * sizeof(float));
    s.col_off = bump((size_t)cap_edges * sizeof(lidx_t));
    s.tdst_off = bump((size_t)cap_edges * sizeof(lidx_t));
    s.total = o;
    return s;
}

// partials layout: [T, slots, NACC] with slots = (per_layer ? L : 1); column NACC-1 of slot l is
// sum <g^{l+1}, F(x^l)> (the gradient of a learnable step) when per_layer, else unused; when the
// weights are shared but g_tau is wanted, tau_partials [T, L] carries it.
template <int CE, int NPT>
__global__ void __launch_bounds__(MaxThreads<NPT>::value, GAD_BWD_MINB) k_fused_bwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                    const int32_t* __restrict__ t_rowptr, const int32_t* __restrict__ t_dst,
                                                    const int32_t* __restrict__ tile_ptr, int cap_nodes, int cap_edges,
                                                    const float* __restrict__ states, const float* __restrict__ g_xphys,
                                                    int dim, const float* __restrict__ Mu_g, int Lw,
                                                    const float* __restrict__ tau, int L, int64_t N,
                                                    float* __restrict__ partials, float* __restrict__ tau_partials,
                                                    float* __restrict__ g_x0) {
    constexpr int NACC = CE * CE + CE + 1;
    constexpr int MUSZ = CE * CE + CE;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const BwdSmem lay = bwd_layout<CE>(cap_nodes, cap_edges, (nthr + 31) >> 5);
    float* X = reinterpret_cast<float*>(smem + lay.x_off);
    float* G = reinterpret_cast<float*>(smem + lay.g_off);
    float* P = reinterpret_cast<float*>(smem + lay.p_off);
    float2* DL = reinterpret_cast<float2*>(smem + lay.dl_off);
    int32_t* rowS = reinterpret_cast<int32_t*>(smem + lay.row_off);
    int32_t* trowS = reinterpret_cast<int32_t*>(smem + lay.trow_off);
    float* Mu = reinterpret_cast<float*>(smem + lay.mu_off);
    float* red = reinterpret_cast<float*>(smem + lay.red_off);
    lidx_t* colS = reinterpret_cast<lidx_t*>(smem + lay.col_off);
    lidx_t* tdstS = reinterpret_cast<lidx_t*>(smem + lay.tdst_off);

    const int tile = blockIdx.x;
    const int n0 = tile_ptr[tile], n1 = tile_ptr[tile + 1];
    const int NT = n1 - n0;
    const int e0 = rowptr[n0], ET = rowptr[n1] - e0;
    const int te0 = t_rowptr[n0];  // == e0 for a closed tile; kept separate for clarity

    for (int e = tid; e < ET; e += nthr) {
        colS[e] = (lidx_t)(col[e0 + e] - n0);
        tdstS[e] = (lidx_t)(t_dst[te0 + e] - n0);
    }
    for (int i = tid; i <= NT; i += nthr) {
        rowS[i] = rowptr[n0 + i] - e0;
        trowS[i] = t_rowptr[n0 + i] - te0;
    }
    // cotangent of x_phys = x^L[:, :dim]  ->  g^L (zero in the other channels)
#pragma unroll
    for (int r = 0; r < NPT; ++r) {
        const int i = tid + r * nthr;
        if (i < NT) {
            Row<CE> g;
#pragma unroll
            for (int c = 0; c < CE; ++c) g.v[c] = (c < dim) ? g_xphys[((int64_t)n0 + i) * dim + c] : 0.f;
            store_row<CE>(G, i, g);
        }
    }
    const bool per_layer = (Lw > 1);
    const int slots = per_layer ? L : 1;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const size_t state_stride = (size_t)N * CE;

    for (int l = L - 1; l >= 0; --l) {
        if (l == L - 1 || per_layer)
            for (int t = tid; t < MUSZ; t += nthr) Mu[t] = Mu_g[(size_t)(per_layer ? l : 0) * MUSZ + t];
        const float* xl = states + (size_t)l * state_stride;
#pragma unroll
        for (int r = 0; r < NPT; ++r) {
            const int i = tid + r * nthr;
            if (i < NT) store_row<CE>(X, i, load_row<CE>(xl, (int64_t)n0 + i));
        }
        __syncthreads();  // X, G (previous layer's update), Mu visible
        const float b = tau[l];
        float gtau = 0.f;
        Row<CE> gself[NPT];
#pragma unroll
        for (int r = 0; r < NPT; ++r) {
            const int i = tid + r * nthr;
            if (i < NT) {
                const Row<CE> xi = load_row<CE>(X, i);
                const Row<CE> gp = load_row<CE>(G, i);
                DstRec<CE> rec;
                gself[r] = node_bwd_dst<CE, lidx_t>(X, colS, rowS[i], rowS[i + 1], xi, gp, 1.0f, b, Mu, &rec, acc, &gtau);
                store_row<CE>(P, i, rec.p);
                DL[i] = make_float2(rec.D, rec.lse);
            }
        }
        __syncthreads();  // P, DL visible
        const bool need_src = (l > 0) || (g_x0 != nullptr);
        if (need_src) {
#pragma unroll
            for (int r = 0; r < NPT; ++r) {
                const int i = tid + r * nthr;
                if (i < NT) {
                    const Row<CE> xj = load_row<CE>(X, i);
                    const Row<CE> sc = node_bwd_src<CE, lidx_t>(P, DL, G, tdstS, trowS[i], trowS[i + 1], xj, b);
#pragma unroll
                    for (int c = 0; c < CE; ++c) gself[r].v[c] += sc.v[c];
                }
            }
        }
        __syncthreads();  // all reads of G / X done
#pragma unroll
        for (int r = 0; r < NPT; ++r) {
            const int i = tid + r * nthr;
            if (i < NT) {
                store_row<CE>(G, i, gself[r]);
                if (l == 0 && g_x0) store_row<CE>(g_x0, (int64_t)n0 + i, gself[r]);
            }
        }
        if (per_layer) {
            acc[NACC - 1] = gtau;
            block_reduce<NACC>(acc, red, partials + ((size_t)tile * slots + l) * NACC);
#pragma unroll
            for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
        } else if (tau_partials) {
            float one[1] = {gtau};
            block_reduce<1>(one, red, tau_partials + (size_t)tile * L + l);
        }
    }
    if (!per_layer) block_reduce<NACC>(acc, red, partials + (size_t)tile * NACC);
}

// Sum the per-tile partials in a fixed order.  One warp per output column.
__global__ void k_fused_reduce(const float* __restrict__ partials, int T, int slots, int nacc, int musz,
                               float* __restrict__ gMu, const float* __restrict__ tau_partials, int L,
                               float* __restrict__ g_tau) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int ncol = slots * nacc;
    if (w < ncol) {
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += partials[(size_t)t * ncol + w];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) {
            const int l = w / nacc, a = w % nacc;
            if (a < musz) gMu[(size_t)l * musz + a] = s;
            else if (g_tau && slots > 1) g_tau[l] = s;
        }
    } else if (w < ncol + L && tau_partials && g_tau) {
        const int l = w - ncol;
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += tau_partials[(size_t)t * L + l];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) g_tau[l] = s;
    }
}

// ---- launch helpers -----------------------------------------------------------------------
struct LaunchShape {
    int npt, threads;
};

// Threads per CTA: enough that every node of the largest tile has an owner with npt <= 4 nodes per
// thread; 2 nodes per thread (ILP across the two rows) until a tile exceeds 1024 nodes.
inline LaunchShape pick_shape(int max_tile_nodes) {
    LaunchShape s;
    if (max_tile_nodes <= MaxThreads<1>::value) s.npt = 1;
    else if (max_tile_nodes <= 2 * MaxThreads<2>::value) s.npt = 2;
    else s.npt = 4;
    int th = (max_tile_nodes + s.npt - 1) / s.npt;
    th = ((th + 31) / 32) * 32;
    if (th < 64) th = 64;
    s.threads = th;
    return s;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    GAD_CHECK_ARG((int)bytes <= smem_optin_bytes(), "tile needs %zu B of shared memory (> %d): use smaller tiles", bytes,
                  smem_optin_bytes());
    GAD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // let several CTAs share an SM: without this the driver picks the smallest carve-out that fits ONE block
    GAD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return GAD_OK;
}

template <int CE, int NPT>
int launch_fwd(int method, int threads, const int32_t* rowptr, const int32_t* col, int64_t N, const int32_t* tile_ptr, int T,
               int cap_nodes, int cap_edges, const float* x0, int dim, const float* Mu, int Lw, const float* tau, int L,
               float* x_phys, float* states, cudaStream_t st) {
    const size_t bytes = fwd_layout<CE>(cap_nodes, cap_edges).total;
    int rc;
    if (method == GAD_METHOD_EULER) {
        if ((rc = set_smem(k_fused_fwd<CE, NPT, GAD_METHOD_EULER>, bytes))) return rc;
        k_fused_fwd<CE, NPT, GAD_METHOD_EULER><<<T, threads, bytes, st>>>(rowptr, col, tile_ptr, cap_nodes, cap_edges, x0, dim,
                                                                          Mu, Lw, tau, L, N, x_phys, states);
    } else {
        if ((rc = set_smem(k_fused_fwd<CE, NPT, GAD_METHOD_RK4>, bytes))) return rc;
        k_fused_fwd<CE, NPT, GAD_METHOD_RK4><<<T, threads, bytes, st>>>(rowptr, col, tile_ptr, cap_nodes, cap_edges, x0, dim,
                                                                        Mu, Lw, tau, L, N, x_phys, states);
    }
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

template <int CE, int NPT>
int launch_bwd(int threads, const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_dst,
               int64_t N, const int32_t* tile_ptr, int T, int cap_nodes, int cap_edges, const float* states,
               const float* g_xphys, int dim, const float* Mu, int Lw, const float* tau, int L, float* partials,
               float* tau_partials, float* g_x0, cudaStream_t st) {
    const size_t bytes = bwd_layout<CE>(cap_nodes, cap_edges, (threads + 31) / 32).total;
    int rc;
    if ((rc = set_smem(k_fused_bwd<CE, NPT>, bytes))) return rc;
    k_fused_bwd<CE, NPT><<<T, threads, bytes, st>>>(rowptr, col, t_rowptr, t_dst, tile_ptr, cap_nodes, cap_edges, states,
                                                    g_xphys, dim, Mu, Lw, tau, L, N, partials, tau_partials, g_x0);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

}  // namespace

#define GAD_FWD_CASE(CE_, NPT_)                                                                                          \
    return launch_fwd<CE_, NPT_>(method, sh.threads, rowptr, col, N, tile_ptr, T, max_tile_nodes, max_tile_edges, x0, dim, \
                                 Mu, Lw, tau, L, x_phys, states, st)
#define GAD_BWD_CASE(CE_, NPT_)                                                                                          \
    rc = launch_bwd<CE_, NPT_>(sh.threads, rowptr, col, t_rowptr, t_dst, N, tile_ptr, T, max_tile_nodes, max_tile_edges,  \
                               states, g_xphys, dim, Mu, Lw, tau, L, partials, tau_partials, g_x0, st)

int fused_forward(int CE, const int32_t* rowptr, const int32_t* col, int64_t N, const int32_t* tile_ptr, int T,
                  int max_tile_nodes, int max_tile_edges, const float* x0, int dim, const float* Mu, int Lw,
                  const float* tau, int L, int method, float* x_phys, float* states, cudaStream_t st) {
    GAD_CHECK_ARG(max_tile_nodes <= MAX_FUSED_TILE_NODES, "fused_forward: tile of %d nodes is too large", max_tile_nodes);
    const LaunchShape sh = pick_shape(max_tile_nodes);
    if (CE == 2) {
        if (sh.npt == 1) GAD_FWD_CASE(2, 1);
        if (sh.npt == 2) GAD_FWD_CASE(2, 2);
        GAD_FWD_CASE(2, 4);
    } else if (CE == 4) {
        if (sh.npt == 1) GAD_FWD_CASE(4, 1);
        if (sh.npt == 2) GAD_FWD_CASE(4, 2);
        GAD_FWD_CASE(4, 4);
    } else {
        if (sh.npt == 1) GAD_FWD_CASE(8, 1);
        if (sh.npt == 2) GAD_FWD_CASE(8, 2);
        GAD_FWD_CASE(8, 4);
    }
}

size_t fused_bwd_ws_floats(int CE, int T, int L) {
    const int NACC = CE * CE + CE + 1;
    return (size_t)T * L * NACC + (size_t)T * L + 64;
}

int fused_backward(int CE, const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_dst,
                   int64_t N, const int32_t* tile_ptr, int T, int max_tile_nodes, int max_tile_edges,
                   const float* states, const float* g_xphys, int dim, const float* Mu, int Lw, const float* tau, int L,
                   float* gMu, float* g_tau, float* g_x0, float* ws, size_t ws_floats, cudaStream_t st) {
    GAD_CHECK_ARG(max_tile_nodes <= MAX_FUSED_TILE_NODES, "fused_backward: tile of %d nodes is too large", max_tile_nodes);
    const int NACC = CE * CE + CE + 1, MUSZ = CE * CE + CE;
    const int slots = Lw > 1 ? L : 1;
    float* partials = ws;
    float* tau_partials = (g_tau && Lw == 1) ? ws + (size_t)T * L * NACC : nullptr;
    (void)ws_floats;
    const LaunchShape sh = pick_shape(max_tile_nodes);
    int rc = GAD_OK;
    if (CE == 2) {
        if (sh.npt == 1) GAD_BWD_CASE(2, 1);
        else if (sh.npt == 2) GAD_BWD_CASE(2, 2);
        else GAD_BWD_CASE(2, 4);
    } else if (CE == 4) {
        if (sh.npt == 1) GAD_BWD_CASE(4, 1);
        else if (sh.npt == 2) GAD_BWD_CASE(4, 2);
        else GAD_BWD_CASE(4, 4);
    } else {
        if (sh.npt == 1) GAD_BWD_CASE(8, 1);
        else if (sh.npt == 2) GAD_BWD_CASE(8, 2);
        else GAD_BWD_CASE(8, 4);
    }
    if (rc) return rc;
    const int ncol = slots * NACC + L;
    k_fused_reduce<<<(ncol + 7) / 8, 256, 0, st>>>(partials, T, slots, NACC, MUSZ, gMu, tau_partials, L, g_tau);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

}  // namespace gad
