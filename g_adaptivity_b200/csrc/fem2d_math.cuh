// Arithmetic of the 2-D differentiable P1 finite-element solve behind loss_type = 'pde_loss' on 2-D meshes
// (torch_FEM_2D, /root/reference/firedrake_difFEM/difFEM_2d.py:345-372), as host/device inline functions:
// the CUDA kernels (fem2d.cu) and the sequential host harness that checks them on the CPU against
// oracle/fem2d_fast.py (oracle/fem2d_host.cpp, tests/test_fem2d_oracle.py) share this file, so the formulas
// are verified before a kernel ever runs.
//
// Two reference behaviours are reproduced on purpose: cubature points are torch.linspace
// points bit for bit (one fused multiply-add per point, up from the lower end in the first half, down from
// the upper end in the second), and both sides of every edge test are rounded separately before they are
// compared (difFEM_2d.py:16-20), because points ON element edges are decided by those comparisons.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FEM_HD __host__ __device__ __forceinline__
#else
#define FEM_HD inline
#endif

// fp32 products and sums that must NOT be contracted into fused multiply-adds (the reference evaluates them as
// individual tensor operations, and a comparison with zero decides at points on element edges)
#if defined(__CUDA_ARCH__)
#define FEM_MUL(a, b) __fmul_rn((a), (b))
#define FEM_ADD(a, b) __fadd_rn((a), (b))
#else
#define FEM_MUL(a, b) ((a) * (b))      // host harness: built with -ffp-contract=off
#define FEM_ADD(a, b) ((a) + (b))
#endif

namespace fem2d {

struct P2 {
    float x, y;
};

// torch.linspace(lo, hi, n)[i] in fp32
FEM_HD float linspace_at(float lo, float hi, int n, int i) {
    const float step = (hi - lo) / (float)(n - 1);
    return (i < n / 2) ? fmaf(step, (float)i, lo) : fmaf(-step, (float)(n - 1 - i), hi);
}

// points per dimension of torchquad's composite Simpson rule for N points in 2-D
FEM_HD int simpson_n(int N) {
    int n = (int)(sqrt((double)N) + 1e-8);
    if (n < 3) return 3;
    return (n & 1) ? n : n - 1;
}
FEM_HD float simpson_w(int n, int i) { return (i == 0 || i == n - 1) ? 1.f : ((i & 1) ? 4.f : 2.f); }

// number of orientation tests (0, 1 or 2) under which P lies in the closed triangle (a, b, c)
FEM_HD int inside_count(P2 P, P2 a, P2 b, P2 c) {
    const float l1 = FEM_ADD(FEM_MUL(a.y - b.y, P.x), FEM_MUL(b.x - a.x, P.y));
    const float r1 = FEM_ADD(FEM_MUL(a.y - b.y, a.x), FEM_MUL(b.x - a.x, a.y));
    const float l2 = FEM_ADD(FEM_MUL(b.y - c.y, P.x), FEM_MUL(c.x - b.x, P.y));
    const float r2 = FEM_ADD(FEM_MUL(b.y - c.y, b.x), FEM_MUL(c.x - b.x, b.y));
    const float l3 = FEM_ADD(FEM_MUL(c.y - a.y, P.x), FEM_MUL(a.x - c.x, P.y));
    const float r3 = FEM_ADD(FEM_MUL(c.y - a.y, c.x), FEM_MUL(a.x - c.x, c.y));
    const int left = (l1 >= r1) && (l2 >= r2) && (l3 >= r3);
    const int right = (l1 <= r1) && (l2 <= r2) && (l3 <= r3);
    return left + right;
}

// barycentric coordinate of vertex c of triangle (a, b, c) at P and its (constant) gradient
FEM_HD float bary(P2 P, P2 a, P2 b, P2 c, float* gx, float* gy) {
    const float den = FEM_ADD(FEM_MUL(a.y - b.y, c.x - a.x), FEM_MUL(c.y - a.y, b.x - a.x));
    *gx = (a.y - b.y) / den;
    *gy = (b.x - a.x) / den;
    const float num = FEM_ADD(FEM_MUL(P.x - c.x, a.y - b.y), FEM_MUL(P.y - c.y, b.x - a.x));
    return FEM_ADD(1.f, num / den);     // its sign at points on the opposite edge enters phim's repeat count
}

// f = laplace(u_true) (difFEM_2d.py:260-266) and u_true (:268-283); centres / scales are fp64 in the reference
FEM_HD double forcing(double x, double y, const double* cen, const double* sc, int G) {
    double out = 0.0;
    for (int g = 0; g < G; ++g) {
        const double c0 = cen[2 * g], c1 = cen[2 * g + 1], s0 = sc[2 * g], s1 = sc[2 * g + 1];
        const double s04 = s0 * s0 * s0 * s0, s14 = s1 * s1 * s1 * s1;
        const double e = exp(-((c0 - x) * (c0 - x) / (s0 * s0)) - (c1 - y) * (c1 - y) / (s1 * s1));
        const double poly = 4 * c1 * c1 * s04 - 2 * s0 * s0 * s14 + 4 * s14 * (c0 - x) * (c0 - x) - 8 * c1 * s04 * y -
                            2 * s04 * (s1 * s1 - 2 * y * y);
        out = (double)(float)(out + (1.0 / (s04 * s14)) * e * poly);     // accumulated into an fp32 tensor
    }
    return out;
}
// The same forcing with the per-Gaussian constants worked out once (the kernels keep them in shared memory): the
// three fp64 divisions per Gaussian and point become multiplications by precomputed reciprocals.  Differs from
// `forcing` by fp64 rounding only (1 ulp of an fp64 intermediate, before the value is rounded to fp32).
struct GaussPre {
    double c0, c1, i0, i1, k, a0, a1, a2, a3;   // i0 = 1/s0^2, i1 = 1/s1^2, k = 1/(s0^4 s1^4); poly coefficients below
};
FEM_HD GaussPre gauss_pre(const double* cen, const double* sc, int g) {
    GaussPre q;
    const double c0 = cen[2 * g], c1 = cen[2 * g + 1], s0 = sc[2 * g], s1 = sc[2 * g + 1];
    const double s04 = s0 * s0 * s0 * s0, s14 = s1 * s1 * s1 * s1;
    q.c0 = c0, q.c1 = c1, q.i0 = 1.0 / (s0 * s0), q.i1 = 1.0 / (s1 * s1), q.k = 1.0 / (s04 * s14);
    q.a0 = 4 * c1 * c1 * s04 - 2 * s0 * s0 * s14;      // terms of `poly` that do not depend on the point
    q.a1 = 4 * s14;                                    // * (c0 - x)^2
    q.a2 = 8 * c1 * s04;                               // * y, subtracted
    q.a3 = 2 * s04;                                    // * (s1^2 - 2 y^2), subtracted
    return q;
}
FEM_HD double forcing_pre(double x, double y, const GaussPre* q, const double* sc, int G) {
    double out = 0.0;
    for (int g = 0; g < G; ++g) {
        const double dx = q[g].c0 - x, dy = q[g].c1 - y, s1 = sc[2 * g + 1];
        const double e = exp(-(dx * dx * q[g].i0) - dy * dy * q[g].i1);
        const double poly = q[g].a0 + q[g].a1 * dx * dx - q[g].a2 * y - q[g].a3 * (s1 * s1 - 2 * y * y);
        out = (double)(float)(out + q[g].k * e * poly);
    }
    return out;
}
FEM_HD double u_true(double x, double y, const double* cen, const double* sc, int G) {
    double out = 0.0;
    for (int g = 0; g < G; ++g) {
        const double c0 = cen[2 * g], c1 = cen[2 * g + 1], s0 = sc[2 * g], s1 = sc[2 * g + 1];
        out = (double)(float)(out + exp(-(x - c0) * (x - c0) / (s0 * s0) - (y - c1) * (y - c1) / (s1 * s1)));
    }
    return out;
}

// local stiffness of one triangle: gradients g[k] of the barycentric coordinates, area, K_ab = area g_a . g_b
struct Tri {
    float gx[3], gy[3], area;
};
FEM_HD Tri tri_geometry(P2 p0, P2 p1, P2 p2) {
    Tri t;
    const float twoA = (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
    const P2 p[3] = {p0, p1, p2};
    for (int k = 0; k < 3; ++k) {
        const P2 u = p[(k + 1) % 3], v = p[(k + 2) % 3];
        t.gx[k] = (u.y - v.y) / twoA;
        t.gy[k] = -(u.x - v.x) / twoA;
    }
    t.area = 0.5f * fabsf(twoA);
    return t;
}
FEM_HD float tri_k(const Tri& t, int a, int b) { return t.area * (t.gx[a] * t.gx[b] + t.gy[a] * t.gy[b]); }

// phi_m(P): sum over the star of node m (difFEM_2d.py:28-60).  `cells` [T,3]; star_cell / star_loc [D] of node m,
// -1 padded.  Returns the value; *rep_out the divisor.
FEM_HD float phi_star(P2 P, const float* coords, const int* cells, const int* star_cell, const int* star_loc, int D,
                      float* rep_out) {
    float out = 0.f, rep = 0.f;
    for (int d = 0; d < D; ++d) {
        const int t = star_cell[d];
        if (t < 0) continue;
        const int k = star_loc[d];
        const int ic = cells[3 * t + k], ia = cells[3 * t + (k + 2) % 3], ib = cells[3 * t + (k + 1) % 3];
        const P2 a = {coords[2 * ia], coords[2 * ia + 1]}, b = {coords[2 * ib], coords[2 * ib + 1]},
                 c = {coords[2 * ic], coords[2 * ic + 1]};
        const int mult = inside_count(P, a, b, c);
        if (!mult) continue;
        float gx, gy;
        const float inc = (float)mult * bary(P, a, b, c, &gx, &gy);
        out += inc;
        rep += (inc > 0.f) ? 1.f : 0.f;
    }
    if (rep == 0.f) rep = 1.f;
    *rep_out = rep;
    return out / rep;
}

}  // namespace fem2d
