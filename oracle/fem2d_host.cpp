// Sequential host run of the arithmetic the 2-D FEM CUDA kernels use (g_adaptivity_b200/csrc/fem2d_math.cuh),
// phase by phase as a CTA will do it, so that formulas, tie handling and the matrix-free CG solve are checked on
// the CPU against oracle/fem2d_fast.py and the reference's fixtures (tests/test_fem2d_oracle.py) before any
// kernel runs.  TEST INFRASTRUCTURE ONLY -- built by oracle/build_host.py into oracle/_build/, loaded with
// ctypes by the tests; nothing in the product path links it.
//
// Follows torch_FEM_2D (/root/reference/firedrake_difFEM/difFEM_2d.py:345-372); see fem2d_math.cuh for the
// per-function citations.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../g_adaptivity_b200/csrc/fem2d_math.cuh"

using namespace fem2d;

namespace {

struct Mesh {
    const int* cells;
    int T;
    const unsigned char* is_bc;
    int N;
    const int* star_cell;
    const int* star_loc;
    int D;
    const float* coords;
};

inline P2 pt(const float* c, int i) { return P2{c[2 * i], c[2 * i + 1]}; }

// y = K_II x  (x, y over all nodes; Dirichlet entries of x are ignored, of y set to 0): gathered per node
void spmv_interior(const Mesh& m, const std::vector<Tri>& tri, const double* x, double* y) {
    for (int i = 0; i < m.N; ++i) {
        double s = 0.0;
        if (!m.is_bc[i])
            for (int d = 0; d < m.D; ++d) {
                const int t = m.star_cell[i * m.D + d];
                if (t < 0) continue;
                const int k = m.star_loc[i * m.D + d];
                for (int kk = 0; kk < 3; ++kk) {
                    const int j = m.cells[3 * t + kk];
                    if (!m.is_bc[j]) s += (double)tri_k(tri[t], k, kk) * x[j];
                }
            }
        y[i] = s;
    }
}

// conjugate gradients on K_II x = b (b = 0 on Dirichlet nodes)
int cg(const Mesh& m, const std::vector<Tri>& tri, const double* b, double* x) {
    const int N = m.N;
    std::vector<double> r(b, b + N), p(b, b + N), Ap(N);
    std::fill(x, x + N, 0.0);
    double rr = 0.0, bb = 0.0;
    for (int i = 0; i < N; ++i) rr += r[i] * r[i];
    bb = rr;
    int it = 0;
    for (; it < 20 * N && rr > 1e-28 * bb && rr > 0.0; ++it) {
        spmv_interior(m, tri, p.data(), Ap.data());
        double pAp = 0.0;
        for (int i = 0; i < N; ++i) pAp += p[i] * Ap[i];
        const double alpha = rr / pAp;
        double rr2 = 0.0;
        for (int i = 0; i < N; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * Ap[i];
            rr2 += r[i] * r[i];
        }
        const double beta = rr2 / rr;
        for (int i = 0; i < N; ++i) p[i] = r[i] + beta * p[i];
        rr = rr2;
    }
    return it;
}

// One (cell, vertex-of-interest) evaluation shared by the load-vector and interpolation gradients:
// grad[v] += -coef * mult * l_v(P) * g_c   for the three vertices v of the cell (c = the vertex whose hat it is)
void scatter_hat_grad(const Mesh& m, int t, int k, P2 P, double coef_times_mult, double* grad) {
    const int ic = m.cells[3 * t + k], ia = m.cells[3 * t + (k + 2) % 3], ib = m.cells[3 * t + (k + 1) % 3];
    const P2 a = pt(m.coords, ia), b = pt(m.coords, ib), c = pt(m.coords, ic);
    float gx, gy, tx, ty;
    const float lc = bary(P, a, b, c, &gx, &gy);
    const float la = bary(P, b, c, a, &tx, &ty);
    const float lb = bary(P, c, a, b, &tx, &ty);
    const int ids[3] = {ia, ib, ic};
    const float lam[3] = {la, lb, lc};
    for (int v = 0; v < 3; ++v) {
        grad[2 * ids[v]] -= coef_times_mult * (double)lam[v] * (double)gx;
        grad[2 * ids[v] + 1] -= coef_times_mult * (double)lam[v] * (double)gy;
    }
}

}  // namespace

extern "C" int fem2d_host(const int* cells, int T, const unsigned char* is_bc, int N, const int* star_cell, const int* star_loc,
                          int D, const float* coords, const double* cen, const double* sc, int G, int load_quad_points,
                          const float* ex, const float* ey, int Q, const float* g_sol, float* coeffs, float* sol, float* grad,
                          int* cg_iters) {
    const Mesh m{cells, T, is_bc, N, star_cell, star_loc, D, coords};
    // ---- phase 1: triangle geometry
    std::vector<Tri> tri(T);
    for (int t = 0; t < T; ++t) tri[t] = tri_geometry(pt(coords, cells[3 * t]), pt(coords, cells[3 * t + 1]), pt(coords, cells[3 * t + 2]));
    // ---- phase 2: load vector (Simpson cubature of phi_m f over the bounding box of the star) / Dirichlet values
    const int n = simpson_n(load_quad_points);
    std::vector<double> rhs(N, 0.0);
    std::vector<float> lox(N), loy(N), hix(N), hiy(N);
    for (int i = 0; i < N; ++i) {
        if (is_bc[i]) {
            rhs[i] = u_true((double)coords[2 * i], (double)coords[2 * i + 1], cen, sc, G);
            continue;
        }
        float lx = INFINITY, ly = INFINITY, hx = -INFINITY, hy = -INFINITY;
        for (int d = 0; d < D; ++d) {
            const int t = star_cell[i * D + d];
            if (t < 0) continue;
            for (int kk = 0; kk < 3; ++kk) {
                const P2 p = pt(coords, cells[3 * t + kk]);
                lx = fminf(lx, p.x), ly = fminf(ly, p.y), hx = fmaxf(hx, p.x), hy = fmaxf(hy, p.y);
            }
        }
        lox[i] = lx, loy[i] = ly, hix[i] = hx, hiy[i] = hy;
        double s = 0.0;
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                const P2 P{linspace_at(lx, hx, n, a), linspace_at(ly, hy, n, b)};
                float rep;
                const float phi = phi_star(P, coords, cells, star_cell + i * D, star_loc + i * D, D, &rep);
                const float f = (float)forcing((double)P.x, (double)P.y, cen, sc, G);
                s += (double)(phi * f) * (double)(simpson_w(n, a) * simpson_w(n, b));
            }
        const float hxs = (hx - lx) / (float)(n - 1), hys = (hy - ly) / (float)(n - 1);
        rhs[i] = (double)(float)(s * (double)hxs * (double)hys / 9.0);
    }
    // ---- phase 3: solve.  A = -K with identity rows on Dirichlet nodes:  K_II u_I = -rhs_I - K_IB u_B
    std::vector<double> u(N, 0.0), b(N, 0.0), tmp(N);
    for (int i = 0; i < N; ++i)
        if (is_bc[i]) u[i] = rhs[i];
    for (int i = 0; i < N; ++i) {
        if (is_bc[i]) continue;
        double s = -rhs[i];
        for (int d = 0; d < D; ++d) {
            const int t = star_cell[i * D + d];
            if (t < 0) continue;
            const int k = star_loc[i * D + d];
            for (int kk = 0; kk < 3; ++kk) {
                const int j = cells[3 * t + kk];
                if (is_bc[j]) s -= (double)tri_k(tri[t], k, kk) * u[j];
            }
        }
        b[i] = s;
    }
    int iters = cg(m, tri, b.data(), tmp.data());
    for (int i = 0; i < N; ++i)
        if (!is_bc[i]) u[i] = tmp[i];
    for (int i = 0; i < N; ++i) coeffs[i] = (float)u[i];
    // ---- phase 4: interpolation.  Per point: the cells that contain it; per distinct vertex value / repeat as in phim
    std::vector<double> g_u(N, 0.0), gacc(grad ? 2 * N : 0, 0.0);
    struct Hit { int t, mult; };
    std::vector<Hit> hits;
    for (int q = 0; q < Q; ++q) {
        const P2 P{ex[q], ey[q]};
        hits.clear();
        for (int t = 0; t < T; ++t) {
            const int mult = inside_count(P, pt(coords, cells[3 * t + 2]), pt(coords, cells[3 * t + 1]), pt(coords, cells[3 * t]));
            if (mult) hits.push_back(Hit{t, mult});
        }
        double val = 0.0;
        for (size_t h = 0; h < hits.size(); ++h)
            for (int k = 0; k < 3; ++k) {
                const int v = cells[3 * hits[h].t + k];
                bool first = true;     // handle vertex v once, at its first occurrence among the hit cells
                for (size_t h2 = 0; h2 < h && first; ++h2)
                    for (int k2 = 0; k2 < 3; ++k2)
                        if (cells[3 * hits[h2].t + k2] == v) first = false;
                if (!first) continue;
                float num = 0.f, rep = 0.f;
                for (size_t h2 = h; h2 < hits.size(); ++h2)
                    for (int k2 = 0; k2 < 3; ++k2) {
                        const int t2 = hits[h2].t;
                        if (cells[3 * t2 + k2] != v) continue;
                        float gx, gy;
                        const float inc = (float)hits[h2].mult * bary(P, pt(coords, cells[3 * t2 + (k2 + 2) % 3]),
                                                                      pt(coords, cells[3 * t2 + (k2 + 1) % 3]), pt(coords, v), &gx, &gy);
                        num += inc;
                        rep += (inc > 0.f) ? 1.f : 0.f;
                    }
                if (rep == 0.f) rep = 1.f;
                const float phi = num / rep;
                val += (double)coeffs[v] * (double)phi;
                if (g_sol) {
                    g_u[v] += (double)g_sol[q] * (double)phi;
                    if (grad)
                        for (size_t h2 = h; h2 < hits.size(); ++h2)
                            for (int k2 = 0; k2 < 3; ++k2)
                                if (cells[3 * hits[h2].t + k2] == v)
                                    scatter_hat_grad(m, hits[h2].t, k2, P,
                                                     (double)coeffs[v] * (double)g_sol[q] * hits[h2].mult / (double)rep, gacc.data());
                }
            }
        sol[q] = (float)val;
    }
    if (cg_iters) *cg_iters = iters;
    if (!g_sol || !grad) return 0;
    // ---- phase 5: adjoint.  lambda_I = -K_II^{-1} g_I; only interior rows of A depend on the mesh
    std::vector<double> lam(N, 0.0), gi(N, 0.0);
    for (int i = 0; i < N; ++i)
        if (!is_bc[i]) gi[i] = -g_u[i];
    iters += cg(m, tri, gi.data(), lam.data());
    if (cg_iters) *cg_iters = iters;
    // load-vector term
    for (int i = 0; i < N; ++i) {
        if (is_bc[i]) continue;
        const float lx = lox[i], ly = loy[i], hx = hix[i], hy = hiy[i];
        const double hh = (double)((hx - lx) / (float)(n - 1)) * (double)((hy - ly) / (float)(n - 1)) / 9.0;
        for (int a = 0; a < n; ++a)
            for (int bq = 0; bq < n; ++bq) {
                const P2 P{linspace_at(lx, hx, n, a), linspace_at(ly, hy, n, bq)};
                float rep;
                phi_star(P, coords, cells, star_cell + i * D, star_loc + i * D, D, &rep);
                const double coef = lam[i] * hh * (double)(simpson_w(n, a) * simpson_w(n, bq)) *
                                    (double)(float)forcing((double)P.x, (double)P.y, cen, sc, G) / (double)rep;
                for (int d = 0; d < D; ++d) {
                    const int t = star_cell[i * D + d];
                    if (t < 0) continue;
                    const int k = star_loc[i * D + d];
                    const int mult = inside_count(P, pt(coords, cells[3 * t + (k + 2) % 3]), pt(coords, cells[3 * t + (k + 1) % 3]),
                                                  pt(coords, cells[3 * t + k]));
                    if (mult) scatter_hat_grad(m, t, k, P, coef * mult, gacc.data());
                }
            }
    }
    // matrix term, per triangle
    for (int t = 0; t < T; ++t) {
        double Glx = 0, Gly = 0, Gux = 0, Guy = 0;
        for (int k = 0; k < 3; ++k) {
            const int v = cells[3 * t + k];
            const double l = is_bc[v] ? 0.0 : lam[v];
            Glx += l * tri[t].gx[k], Gly += l * tri[t].gy[k];
            Gux += u[v] * tri[t].gx[k], Guy += u[v] * tri[t].gy[k];
        }
        const double dot = Glx * Gux + Gly * Guy;
        for (int k = 0; k < 3; ++k) {
            const int v = cells[3 * t + k];
            const double gx = tri[t].gx[k], gy = tri[t].gy[k];
            const double gGu = gx * Gux + gy * Guy, gGl = gx * Glx + gy * Gly;
            gacc[2 * v] += tri[t].area * (dot * gx - gGu * Glx - gGl * Gux);
            gacc[2 * v + 1] += tri[t].area * (dot * gy - gGu * Gly - gGl * Guy);
        }
    }
    for (int i = 0; i < 2 * N; ++i) grad[i] = (float)gacc[i];
    return 0;
}
