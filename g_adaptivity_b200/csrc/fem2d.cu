// Batched differentiable 2-D P1 finite-element solve behind loss_type = 'pde_loss' on 2-D meshes (scope row f1,
// second half): torch_FEM_2D of /root/reference/firedrake_difFEM/difFEM_2d.py:345-372, which the reference runs per
// mesh in a Python loop (src/GNN.py:327-335), as one CTA per mesh over a topology shared by the batch.
//
// STATUS: first version, correctness-first.  Its arithmetic (fem2d_math.cuh) and phase order are checked on the CPU
// against the reference's fixtures (oracle/fem2d_host.cpp, tests/test_fem2d_oracle.py), and the kernels below run
// on the CPU under a thread-per-CUDA-thread emulation (oracle/fem2d_emu.cpp: FEM2D_EMULATE) with the same results;
// on the B200 they match the fixtures to 4e-6 / 1.2e-5 (forward / gradient; tests/test_fem2d_gpu.py) and take 36 ms
// for 256 meshes of 30x30, forward + backward (scripts/fem2d_check.py).  Not tuned: scalar CG rows, all-cells point
// location.  GNN.forward does not route to them yet.
//
// Phases (forward): triangle geometry -> load vector (Simpson cubature per interior node, Dirichlet values) ->
// matrix-free conjugate gradients on the interior SPD system (rows gathered through the star table, fixed-order
// block reductions, fp64) -> interpolation on the evaluation points (point location: all cells, bounding-box
// prefilter; a bin table is the next step).
// Backward: g_u by fp64 shared-memory accumulation, second CG solve for the adjoint, three gradient terms
// (matrix, load vector, interpolation) accumulated per vertex in shared memory, one store per vertex.
#ifndef FEM2D_EMULATE
#include "common.cuh"
#endif
#include "fem2d_math.cuh"

namespace gad {
#ifdef FEM2D_EMULATE          // oracle/fem2d_emu.cpp: the kernels on CPU threads; they must be visible to that file
inline namespace emu {
#else
namespace {
#endif

using namespace fem2d;

constexpr int FEM2D_THREADS = 256;
constexpr int MAX_HITS = 12;

struct F2Args {
    const int* cells;        // [T,3]
    const unsigned char* is_bc;   // [N]
    const int* star_cell;    // [N,D]
    const int* star_loc;     // [N,D]
    const float* coords;     // [B,N,2]
    const double* cen;       // [B,G,2]
    const double* sc;        // [B,G,2]
    const float* ex;         // [Q]
    const float* ey;         // [Q]
    const float* g_sol;      // [B,Q]   (backward)
    float* coeffs;           // [B,N]
    float* sol;              // [B,Q]
    double* u64;             // [B,N]   forward -> backward
    float* grad;             // [B,N,2] (backward)
    int* cg_iters;           // [B] or null
    int T, N, D, G, K, Q;
};

struct Box {
    float lx, ly, hx, hy;
};
// A point that passes the (rounded) edge tests of a cell can lie outside the cell's exact bounding box only by the
// rounding error of those tests (~1e-7 / edge length); the margin is orders of magnitude above that, so the prefilter
// never changes which cells count.
constexpr float BOX_MARGIN = 1e-3f;

struct F2Smem {
    float* xy;       // [2N]
    Tri* tri;        // [T]
    Box* box;        // [T]   bounding boxes of the cells, inflated: prefilter of the point location
    double* u;       // [N]
    double* b;       // [N]
    double* r;       // [N]
    double* p;       // [N]
    double* Ap;      // [N]
    double* acc;     // [2N]  gradient / g_u accumulators
    double* red;     // [32]
};

__host__ __device__ inline size_t f2_align(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t f2_smem_bytes(int N, int T) {
    return f2_align(2 * (size_t)N * 4) + f2_align((size_t)T * sizeof(Tri)) + f2_align((size_t)T * sizeof(Box)) +
           5 * f2_align((size_t)N * 8) + f2_align(2 * (size_t)N * 8) + f2_align(32 * 8);
}

__device__ inline F2Smem f2_carve(unsigned char* base, int N, int T) {
    F2Smem s;
    size_t o = 0;
    s.xy = reinterpret_cast<float*>(base + o), o += f2_align(2 * (size_t)N * 4);
    s.tri = reinterpret_cast<Tri*>(base + o), o += f2_align((size_t)T * sizeof(Tri));
    s.box = reinterpret_cast<Box*>(base + o), o += f2_align((size_t)T * sizeof(Box));
    s.u = reinterpret_cast<double*>(base + o), o += f2_align((size_t)N * 8);
    s.b = reinterpret_cast<double*>(base + o), o += f2_align((size_t)N * 8);
    s.r = reinterpret_cast<double*>(base + o), o += f2_align((size_t)N * 8);
    s.p = reinterpret_cast<double*>(base + o), o += f2_align((size_t)N * 8);
    s.Ap = reinterpret_cast<double*>(base + o), o += f2_align((size_t)N * 8);
    s.acc = reinterpret_cast<double*>(base + o), o += f2_align(2 * (size_t)N * 8);
    s.red = reinterpret_cast<double*>(base + o);
    return s;
}

__device__ inline P2 f2_pt(const float* xy, int i) { return P2{xy[2 * i], xy[2 * i + 1]}; }

// fixed-order block sum (every thread gets the result)
__device__ double f2_block_sum(double v, double* red) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < nw; ++k) s += red[k];
    return s;
}

__device__ void f2_geometry(const F2Args& a, const F2Smem& s, int mesh) {
    const float* c = a.coords + (size_t)mesh * a.N * 2;
    for (int i = threadIdx.x; i < 2 * a.N; i += blockDim.x) s.xy[i] = c[i];
    __syncthreads();
    for (int t = threadIdx.x; t < a.T; t += blockDim.x) {
        const P2 p0 = f2_pt(s.xy, a.cells[3 * t]), p1 = f2_pt(s.xy, a.cells[3 * t + 1]), p2 = f2_pt(s.xy, a.cells[3 * t + 2]);
        s.tri[t] = tri_geometry(p0, p1, p2);
        s.box[t] = Box{fminf(p0.x, fminf(p1.x, p2.x)) - BOX_MARGIN, fminf(p0.y, fminf(p1.y, p2.y)) - BOX_MARGIN,
                       fmaxf(p0.x, fmaxf(p1.x, p2.x)) + BOX_MARGIN, fmaxf(p0.y, fmaxf(p1.y, p2.y)) + BOX_MARGIN};
    }
    __syncthreads();
}

// y = K_II x over the nodes of this thread
__device__ void f2_spmv(const F2Args& a, const F2Smem& s, const double* x, double* y) {
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
        double acc = 0.0;
        if (!a.is_bc[i])
            for (int d = 0; d < a.D; ++d) {
                const int t = a.star_cell[i * a.D + d];
                if (t < 0) continue;
                const int k = a.star_loc[i * a.D + d];
                for (int kk = 0; kk < 3; ++kk) {
                    const int j = a.cells[3 * t + kk];
                    if (!a.is_bc[j]) acc += (double)tri_k(s.tri[t], k, kk) * x[j];
                }
            }
        y[i] = acc;
    }
    __syncthreads();
}

// conjugate gradients on K_II x = b (b in s.b, zero on Dirichlet nodes); x -> out (all threads see it after return)
__device__ int f2_cg(const F2Args& a, const F2Smem& s, double* out) {
    double rr_local = 0.0;
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
        out[i] = 0.0;
        s.r[i] = s.b[i];
        s.p[i] = s.b[i];
        rr_local += s.b[i] * s.b[i];
    }
    double rr = f2_block_sum(rr_local, s.red);
    const double bb = rr;
    int it = 0;
    for (; it < 20 * a.N && rr > 1e-28 * bb && rr > 0.0; ++it) {
        f2_spmv(a, s, s.p, s.Ap);
        double l = 0.0;
        for (int i = threadIdx.x; i < a.N; i += blockDim.x) l += s.p[i] * s.Ap[i];
        const double alpha = rr / f2_block_sum(l, s.red);
        l = 0.0;
        for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
            out[i] += alpha * s.p[i];
            s.r[i] -= alpha * s.Ap[i];
            l += s.r[i] * s.r[i];
        }
        const double rr2 = f2_block_sum(l, s.red);
        const double beta = rr2 / rr;
        for (int i = threadIdx.x; i < a.N; i += blockDim.x) s.p[i] = s.r[i] + beta * s.p[i];
        rr = rr2;
        __syncthreads();
    }
    __syncthreads();
    return it;
}

__device__ void f2_star_box(const F2Args& a, const F2Smem& s, int i, float& lx, float& ly, float& hx, float& hy) {
    lx = ly = INFINITY;
    hx = hy = -INFINITY;
    for (int d = 0; d < a.D; ++d) {
        const int t = a.star_cell[i * a.D + d];
        if (t < 0) continue;
        for (int kk = 0; kk < 3; ++kk) {
            const P2 p = f2_pt(s.xy, a.cells[3 * t + kk]);
            lx = fminf(lx, p.x), ly = fminf(ly, p.y), hx = fmaxf(hx, p.x), hy = fmaxf(hy, p.y);
        }
    }
}

// grad[v] += -coef * l_v(P) * g_c for the three vertices of cell t (c = local vertex k)
__device__ void f2_scatter(const F2Args& a, const F2Smem& s, int t, int k, P2 P, double coef) {
    const int ic = a.cells[3 * t + k], ia = a.cells[3 * t + (k + 2) % 3], ib = a.cells[3 * t + (k + 1) % 3];
    const P2 pa = f2_pt(s.xy, ia), pb = f2_pt(s.xy, ib), pc = f2_pt(s.xy, ic);
    float gx, gy, tx, ty;
    const float lc = bary(P, pa, pb, pc, &gx, &gy);
    const float la = bary(P, pb, pc, pa, &tx, &ty);
    const float lb = bary(P, pc, pa, pb, &tx, &ty);
    atomicAdd(&s.acc[2 * ia], -coef * (double)la * (double)gx);
    atomicAdd(&s.acc[2 * ia + 1], -coef * (double)la * (double)gy);
    atomicAdd(&s.acc[2 * ib], -coef * (double)lb * (double)gx);
    atomicAdd(&s.acc[2 * ib + 1], -coef * (double)lb * (double)gy);
    atomicAdd(&s.acc[2 * ic], -coef * (double)lc * (double)gx);
    atomicAdd(&s.acc[2 * ic + 1], -coef * (double)lc * (double)gy);
}

// The evaluation points of this thread: hit cells, per distinct vertex value / repeat (phim, difFEM_2d.py:28-60).
// MODE 0: sol[q].  MODE 1: g_u accumulation (s.acc[v]).  MODE 2: interpolation term of the gradient (s.acc[2v..]).
template <int MODE>
__device__ void f2_points(const F2Args& a, const F2Smem& s, int mesh) {
    for (int q = threadIdx.x; q < a.Q; q += blockDim.x) {
        const P2 P{a.ex[q], a.ey[q]};
        int ht[MAX_HITS], hm[MAX_HITS], nh = 0;
        for (int t = 0; t < a.T && nh < MAX_HITS; ++t) {
            const Box bx = s.box[t];
            if (P.x < bx.lx || P.x > bx.hx || P.y < bx.ly || P.y > bx.hy) continue;
            const int mult = inside_count(P, f2_pt(s.xy, a.cells[3 * t + 2]), f2_pt(s.xy, a.cells[3 * t + 1]), f2_pt(s.xy, a.cells[3 * t]));
            if (mult) ht[nh] = t, hm[nh] = mult, ++nh;
        }
        const double gq = (MODE == 0) ? 0.0 : (double)a.g_sol[(size_t)mesh * a.Q + q];
        double val = 0.0;
        for (int h = 0; h < nh; ++h)
            for (int k = 0; k < 3; ++k) {
                const int v = a.cells[3 * ht[h] + k];
                bool first = true;
                for (int h2 = 0; h2 < h && first; ++h2)
                    for (int k2 = 0; k2 < 3; ++k2)
                        if (a.cells[3 * ht[h2] + k2] == v) first = false;
                if (!first) continue;
                float num = 0.f, rep = 0.f;
                for (int h2 = h; h2 < nh; ++h2)
                    for (int k2 = 0; k2 < 3; ++k2) {
                        const int t2 = ht[h2];
                        if (a.cells[3 * t2 + k2] != v) continue;
                        float gx, gy;
                        const float inc = (float)hm[h2] * bary(P, f2_pt(s.xy, a.cells[3 * t2 + (k2 + 2) % 3]),
                                                               f2_pt(s.xy, a.cells[3 * t2 + (k2 + 1) % 3]), f2_pt(s.xy, v), &gx, &gy);
                        num += inc;
                        rep += (inc > 0.f) ? 1.f : 0.f;
                    }
                if (rep == 0.f) rep = 1.f;
                const float phi = num / rep;
                const double uv = (double)(float)s.u[v];
                if (MODE == 0) val += uv * (double)phi;
                if (MODE == 1) atomicAdd(&s.acc[v], gq * (double)phi);
                if (MODE == 2)
                    for (int h2 = h; h2 < nh; ++h2)
                        for (int k2 = 0; k2 < 3; ++k2)
                            if (a.cells[3 * ht[h2] + k2] == v) f2_scatter(a, s, ht[h2], k2, P, uv * gq * hm[h2] / (double)rep);
            }
        if (MODE == 0) a.sol[(size_t)mesh * a.Q + q] = (float)val;
    }
    __syncthreads();
}

// load vector (MODE 0: rhs into s.b as doubles) or its gradient term (MODE 1: needs lambda in s.r)
template <int MODE>
__device__ void f2_load(const F2Args& a, const F2Smem& s, const double* cen, const double* sc) {
    const int n = simpson_n(a.K);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = warp; i < a.N; i += nw) {
        if (a.is_bc[i]) {
            if (MODE == 0 && lane == 0) s.b[i] = u_true((double)s.xy[2 * i], (double)s.xy[2 * i + 1], cen, sc, a.G);
            continue;
        }
        float lx, ly, hx, hy;
        f2_star_box(a, s, i, lx, ly, hx, hy);
        const double hh = (double)((hx - lx) / (float)(n - 1)) * (double)((hy - ly) / (float)(n - 1)) / 9.0;
        double part = 0.0;
        for (int pq = lane; pq < n * n; pq += 32) {
            const int ia = pq / n, ib = pq % n;
            const P2 P{linspace_at(lx, hx, n, ia), linspace_at(ly, hy, n, ib)};
            float rep;
            const float phi = phi_star(P, s.xy, a.cells, a.star_cell + i * a.D, a.star_loc + i * a.D, a.D, &rep);
            const float f = (float)forcing((double)P.x, (double)P.y, cen, sc, a.G);
            const double w = (double)(simpson_w(n, ia) * simpson_w(n, ib));
            if (MODE == 0) {
                part += (double)(phi * f) * w;
            } else {
                const double coef = s.r[i] * hh * w * (double)f / (double)rep;
                for (int d = 0; d < a.D; ++d) {
                    const int t = a.star_cell[i * a.D + d];
                    if (t < 0) continue;
                    const int k = a.star_loc[i * a.D + d];
                    const int mult = inside_count(P, f2_pt(s.xy, a.cells[3 * t + (k + 2) % 3]), f2_pt(s.xy, a.cells[3 * t + (k + 1) % 3]),
                                                  f2_pt(s.xy, a.cells[3 * t + k]));
                    if (mult) f2_scatter(a, s, t, k, P, coef * mult);
                }
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            if (lane == 0) s.b[i] = (double)(float)(part * hh);
        }
    }
    __syncthreads();
}

// right-hand side of the interior system from rhs (s.b) and the Dirichlet values: b_i = -rhs_i - sum_B K_ij u_j
__device__ void f2_interior_rhs(const F2Args& a, const F2Smem& s) {
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) s.u[i] = a.is_bc[i] ? s.b[i] : 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
        double v = 0.0;
        if (!a.is_bc[i]) {
            v = -s.b[i];
            for (int d = 0; d < a.D; ++d) {
                const int t = a.star_cell[i * a.D + d];
                if (t < 0) continue;
                const int k = a.star_loc[i * a.D + d];
                for (int kk = 0; kk < 3; ++kk) {
                    const int j = a.cells[3 * t + kk];
                    if (a.is_bc[j]) v -= (double)tri_k(s.tri[t], k, kk) * s.u[j];
                }
            }
        }
        s.Ap[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) s.b[i] = s.Ap[i];
    __syncthreads();
}

__global__ void __launch_bounds__(FEM2D_THREADS) k_fem2d_fwd(F2Args a) {
    extern __shared__ __align__(16) unsigned char f2_raw[];
    const F2Smem s = f2_carve(f2_raw, a.N, a.T);
    const int mesh = blockIdx.x;
    const double* cen = a.cen + (size_t)mesh * a.G * 2;
    const double* sc = a.sc + (size_t)mesh * a.G * 2;
    f2_geometry(a, s, mesh);
    f2_load<0>(a, s, cen, sc);
    f2_interior_rhs(a, s);
    const int it = f2_cg(a, s, s.acc);              // interior solution in s.acc[0..N)
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
        if (!a.is_bc[i]) s.u[i] = s.acc[i];
        a.coeffs[(size_t)mesh * a.N + i] = (float)s.u[i];
        a.u64[(size_t)mesh * a.N + i] = s.u[i];
    }
    __syncthreads();
    f2_points<0>(a, s, mesh);
    if (a.cg_iters && threadIdx.x == 0) a.cg_iters[mesh] = it;
}

__global__ void __launch_bounds__(FEM2D_THREADS) k_fem2d_bwd(F2Args a) {
    extern __shared__ __align__(16) unsigned char f2_raw[];
    const F2Smem s = f2_carve(f2_raw, a.N, a.T);
    const int mesh = blockIdx.x;
    const double* cen = a.cen + (size_t)mesh * a.G * 2;
    const double* sc = a.sc + (size_t)mesh * a.G * 2;
    f2_geometry(a, s, mesh);
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) s.u[i] = a.u64[(size_t)mesh * a.N + i];
    for (int i = threadIdx.x; i < 2 * a.N; i += blockDim.x) s.acc[i] = 0.0;
    __syncthreads();
    f2_points<1>(a, s, mesh);                        // g_u in s.acc[0..N)
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) s.b[i] = a.is_bc[i] ? 0.0 : -s.acc[i];
    __syncthreads();
    double* lam = s.acc + a.N;                       // second half of the accumulator array, free until the scatter
    f2_cg(a, s, lam);                                // K_II lambda_I = -g_I
    // keep lambda in s.r (zero on Dirichlet nodes), then reuse s.acc as the gradient accumulator
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) s.b[i] = a.is_bc[i] ? 0.0 : lam[i];
    __syncthreads();
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) s.r[i] = s.b[i];
    for (int i = threadIdx.x; i < 2 * a.N; i += blockDim.x) s.acc[i] = 0.0;
    __syncthreads();
    f2_points<2>(a, s, mesh);                        // interpolation term
    f2_load<1>(a, s, cen, sc);                       // load-vector term
    for (int t = threadIdx.x; t < a.T; t += blockDim.x) {      // matrix term
        const Tri& g = s.tri[t];
        double Glx = 0, Gly = 0, Gux = 0, Guy = 0;
        for (int k = 0; k < 3; ++k) {
            const int v = a.cells[3 * t + k];
            Glx += s.r[v] * g.gx[k], Gly += s.r[v] * g.gy[k];
            Gux += s.u[v] * g.gx[k], Guy += s.u[v] * g.gy[k];
        }
        const double dot = Glx * Gux + Gly * Guy;
        for (int k = 0; k < 3; ++k) {
            const int v = a.cells[3 * t + k];
            const double gx = g.gx[k], gy = g.gy[k];
            const double gGu = gx * Gux + gy * Guy, gGl = gx * Glx + gy * Gly;
            atomicAdd(&s.acc[2 * v], (double)g.area * (dot * gx - gGu * Glx - gGl * Gux));
            atomicAdd(&s.acc[2 * v + 1], (double)g.area * (dot * gy - gGu * Gly - gGl * Guy));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.N; i += blockDim.x) a.grad[(size_t)mesh * a.N * 2 + i] = (float)s.acc[i];
}

#ifndef FEM2D_EMULATE
int f2_launch(const F2Args& a, int B, bool backward, cudaStream_t st) {
    const size_t bytes = f2_smem_bytes(a.N, a.T);
    GAD_CHECK_ARG((int)bytes <= smem_optin_bytes(), "fem2d: a mesh of %d nodes / %d cells needs %zu B of shared memory", a.N, a.T,
                  bytes);
    if (backward) {
        GAD_CUDA(cudaFuncSetAttribute(k_fem2d_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        k_fem2d_bwd<<<B, FEM2D_THREADS, bytes, st>>>(a);
    } else {
        GAD_CUDA(cudaFuncSetAttribute(k_fem2d_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        k_fem2d_fwd<<<B, FEM2D_THREADS, bytes, st>>>(a);
    }
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}
#endif

}  // namespace
}  // namespace gad

#ifndef FEM2D_EMULATE
using namespace gad;

extern "C" int gad_fem2d_fwd(const int32_t* cells, int32_t T, const uint8_t* is_bc, int32_t N, const int32_t* star_cell,
                             const int32_t* star_loc, int32_t D, const float* coords, const double* centers, const double* scales,
                             int32_t G, int32_t B, int32_t load_quad_points, const float* eval_x, const float* eval_y, int32_t Q,
                             float* coeffs, float* sol, double* u64, int32_t* cg_iters, void* stream) {
    GAD_CHECK_ARG(cells && is_bc && star_cell && star_loc && coords && centers && scales && eval_x && eval_y && coeffs && sol && u64,
                  "gad_fem2d_fwd: null argument");
    GAD_CHECK_ARG(T > 0 && N > 0 && D > 0 && G > 0 && B > 0 && Q > 0 && load_quad_points > 0, "gad_fem2d_fwd: bad sizes");
    F2Args a = {};
    a.cells = cells, a.is_bc = is_bc, a.star_cell = star_cell, a.star_loc = star_loc, a.coords = coords, a.cen = centers, a.sc = scales;
    a.ex = eval_x, a.ey = eval_y, a.coeffs = coeffs, a.sol = sol, a.u64 = u64, a.cg_iters = cg_iters;
    a.T = T, a.N = N, a.D = D, a.G = G, a.K = load_quad_points, a.Q = Q;
    return f2_launch(a, B, false, as_stream(stream));
}

extern "C" int gad_fem2d_bwd(const int32_t* cells, int32_t T, const uint8_t* is_bc, int32_t N, const int32_t* star_cell,
                             const int32_t* star_loc, int32_t D, const float* coords, const double* centers, const double* scales,
                             int32_t G, int32_t B, int32_t load_quad_points, const float* eval_x, const float* eval_y, int32_t Q,
                             const double* u64, const float* g_sol, float* grad, void* stream) {
    GAD_CHECK_ARG(cells && is_bc && star_cell && star_loc && coords && centers && scales && eval_x && eval_y && u64 && g_sol && grad,
                  "gad_fem2d_bwd: null argument");
    GAD_CHECK_ARG(T > 0 && N > 0 && D > 0 && G > 0 && B > 0 && Q > 0 && load_quad_points > 0, "gad_fem2d_bwd: bad sizes");
    F2Args a = {};
    a.cells = cells, a.is_bc = is_bc, a.star_cell = star_cell, a.star_loc = star_loc, a.coords = coords, a.cen = centers, a.sc = scales;
    a.ex = eval_x, a.ey = eval_y, a.g_sol = g_sol, a.u64 = const_cast<double*>(u64), a.grad = grad;
    a.T = T, a.N = N, a.D = D, a.G = G, a.K = load_quad_points, a.Q = Q;
    return f2_launch(a, B, true, as_stream(stream));
}
#endif  // FEM2D_EMULATE
