"""CPU restatement of the 2-D differentiable FEM solve that follows the deformer when
`loss_type='pde_loss'` on 2-D meshes (scope row f1 of SURVEY section 8, second half):
`torch_FEM_2D` of /root/reference/firedrake_difFEM/difFEM_2d.py:345-372 with `build_mass_matrix`
(:63-117), `build_load_vector` (:159-203), `phim` / `aux` / `check*` (:16-60), `bounds_support_jr`
(:298-309), `soln` (:312-318), `f` (:260-266), `u_true_exact_2d` / `BCfn` (:268-287),
`cubature2d_v2` (:337-342).

TEST INFRASTRUCTURE ONLY (nothing in the product path imports it).  No CUDA kernel is built on it yet:
this file and its fixtures are the first step of the row (oracle before kernels).

Pinning.  The Firedrake objects the reference queries are replaced by plain arrays: `cells` [T, 3]
(= `mesh.coordinates.cell_node_map().values`) and `bc_nodes` (= `DirichletBC(V, 0, "on_boundary").nodes`).
Everything above is pinned bit for bit by tests/golden_fem2d/*.pt, minted by running the reference's own
file in place behind stand-ins for firedrake / matplotlib / torchdiffeq / torchquad
(oracle/ref_harness/make_golden_fem2d.py).  PARITY UNPINNED for one piece: the quadrature.  The reference
calls `torchquad.Simpson().integrate(fn, dim=2, N=n, integration_domain=..., backend="torch")`; torchquad
is a third-party dependency that is neither vendored in /root/reference nor pinned there (README.md:34
says `pip install torchquad`, no version) and is not installed here.  `simpson_2d` restates its published
composite Simpson rule (torchquad 0.4: N per dimension = floor(N^(1/2)), reduced by one when even, at
least 3; `linspace` grid per dimension, points ordered with dimension 0 slowest; weights
h/3 * (1, 4, 2, ..., 4, 1) applied last axis first) and the harness hands the reference this very
function, so the fixtures cannot see a difference between it and the real package."""
from __future__ import annotations

import torch


# ---- geometry of one hat function (difFEM_2d.py:16-26) --------------------------------------------
def _side(x, a, b):
    return (a[1] - b[1]) * x[0] + (b[0] - a[0]) * x[1], (a[1] - b[1]) * a[0] + (b[0] - a[0]) * a[1]


def check_left(x, a, b):
    l, r = _side(x, a, b)
    return (l >= r) * 1.0


def check_right(x, a, b):
    l, r = _side(x, a, b)
    return (l <= r) * 1.0


def hat_on_cell(x, a, b, c):
    """Value at x of the P1 function that is 1 at vertex c and 0 on edge ab, times the indicator of the
    (closed) triangle abc in either orientation (:25-26)."""
    inside = (check_left(x, a, b) * check_left(x, b, c) * check_left(x, c, a)
              + check_right(x, a, b) * check_right(x, b, c) * check_right(x, c, a))
    lin = 1 + ((x[0] - c[0]) * (a[1] - b[1]) + (x[1] - c[1]) * (b[0] - a[0])) / (
        (a[1] - b[1]) * (c[0] - a[0]) + (c[1] - a[1]) * (b[0] - a[0]))
    return inside * lin


def phim(x, n, coords, cells):
    """Basis function of node n at the points x = (x0, x1) (:28-60): sum over the cells that contain n,
    divided by the number of cells that contributed a positive value (points on shared edges / at the node
    are counted once per cell)."""
    where = torch.where(cells == n)
    out = x[0] * 0.0
    repeat = x[0] * 0.0
    for i in range(where[0].shape[0]):
        cell, k = where[0][i], where[1][i]
        c = coords[cells[cell][k]]
        a = coords[cells[cell][torch.fmod(k - 1, 3)]]     # negative remainders index from the end
        b = coords[cells[cell][torch.fmod(k - 2, 3)]]
        inc = hat_on_cell(x, a, b, c)
        out = out + inc
        repeat = repeat + (inc > 0.0) * 1.0
    return out / (repeat + 1.0 * (repeat == 0.0))


# ---- right-hand side and exact solution (difFEM_2d.py:260-287) -------------------------------------
def forcing(x, c_list, s_list):
    sol = torch.zeros(x[0].shape)
    for c, s in zip(c_list, s_list):
        sol += (1 / (s[0] ** 4 * s[1] ** 4)) * torch.exp(-((c[0] - x[0]) ** 2 / s[0] ** 2) - (c[1] - x[1]) ** 2 / s[1] ** 2) * (
            4 * c[1] ** 2 * s[0] ** 4 - 2 * s[0] ** 2 * s[1] ** 4 + 4 * s[1] ** 4 * (c[0] - x[0]) ** 2
            - 8 * c[1] * s[0] ** 4 * x[1] - 2 * s[0] ** 4 * (s[1] ** 2 - 2 * x[1] ** 2))
    return sol


def u_true(x, c_list, s_list):
    sol = torch.zeros(x[0].shape)
    for c, s in zip(c_list, s_list):
        sol += torch.exp(-(x[0] - c[0]) ** 2 / s[0] ** 2 - (x[1] - c[1]) ** 2 / s[1] ** 2)
    return sol


# ---- quadrature: restated torchquad composite Simpson (see the header: unpinned) --------------------
def simpson_points_per_dim(N: int, dim: int = 2) -> int:
    n = int(N ** (1.0 / dim) + 1e-8)
    if n < 3:
        return 3
    return n if n % 2 == 1 else n - 1


def simpson_2d(integrand, N, domain):
    """integral of `integrand` (called with points [n*n, 2], dimension 0 slowest) over the box `domain`."""
    n = simpson_points_per_dim(int(N), 2)
    lo = [torch.as_tensor(domain[d][0], dtype=torch.float32) for d in range(2)]
    hi = [torch.as_tensor(domain[d][1], dtype=torch.float32) for d in range(2)]
    g = [torch.linspace(float(lo[d]), float(hi[d]), n) for d in range(2)]
    hs = [(hi[d] - lo[d]) / (n - 1) for d in range(2)]
    X, Y = torch.meshgrid(g[0], g[1], indexing="ij")
    pts = torch.stack([X.reshape(-1), Y.reshape(-1)], dim=1)
    vals = integrand(pts).reshape(n, n)
    for d in range(2):
        vals = hs[d] / 3.0 * (vals[..., 0:-2][..., ::2] + 4 * vals[..., 1:-1][..., ::2] + vals[..., 2:][..., ::2])
        vals = torch.sum(vals, dim=2 - d - 1)
    return vals


# ---- assembly (difFEM_2d.py:63-117, 159-203, 298-309) ---------------------------------------------
def build_stiffness(cells, mesh_points, num_nodes):
    """P1 stiffness matrix from per-triangle gradients: `slopes` solves [1 x y] S = I per triangle, entries
    are area * grad(phi_p) . grad(phi_q); returned with the reference's sign (minus: integration by parts)."""
    tri = torch.stack([mesh_points[cell] for cell in cells])
    T = tri.shape[0]
    A = torch.cat((torch.ones(T, 3, 1), tri), dim=2)
    B = torch.tensor([[1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=torch.float32).repeat(T, 1, 1)
    slopes = torch.linalg.solve(A, B)
    x, y = tri[:, :, 0], tri[:, :, 1]
    area = 0.5 * torch.abs(x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
    i_idx, j_idx, k_idx = cells[:, 0], cells[:, 1], cells[:, 2]
    s_i, s_j, s_k = slopes[:, 1:, 0], slopes[:, 1:, 1], slopes[:, 1:, 2]
    w = area.unsqueeze(1)
    Mii, Mjj, Mkk = (s_i * s_i * w).sum(1), (s_j * s_j * w).sum(1), (s_k * s_k * w).sum(1)
    Mij, Mjk, Mki = (s_i * s_j * w).sum(1), (s_j * s_k * w).sum(1), (s_k * s_i * w).sum(1)
    rows = torch.cat((i_idx, j_idx, k_idx, i_idx, j_idx, k_idx, j_idx, k_idx, i_idx))
    cols = torch.cat((i_idx, j_idx, k_idx, j_idx, k_idx, i_idx, i_idx, j_idx, k_idx))
    vals = torch.cat((Mii, Mjj, Mkk, Mij, Mjk, Mki, Mij, Mjk, Mki))
    M = -torch.sparse_coo_tensor(torch.stack((rows, cols)), vals, size=(num_nodes, num_nodes))
    return M, tri, slopes


def support_box(m, coords, cells):
    idx = cells[torch.where(cells == m)[0], :].flatten()
    lo = torch.min(coords[idx].detach(), 0)[0]
    hi = torch.max(coords[idx].detach(), 0)[0]
    return [[lo[0], hi[0]], [lo[1], hi[1]]]


def build_load_vector(cells, bc_nodes, coords, num_nodes, load_quad_points, c_list, s_list):
    bc = set(int(b) for b in bc_nodes)
    rhs = torch.zeros(num_nodes, 1)
    for m in range(num_nodes):
        if m in bc:
            # torch.tensor([...]) copies the two coordinates: no gradient through the Dirichlet values (:172)
            rhs[m] = u_true(torch.tensor([coords[m, 0], coords[m, 1]]), c_list, s_list)
        else:
            def integrand(p, m=m):
                x = torch.transpose(p, 0, 1)
                return phim(x, m, coords, cells) * forcing(x, c_list, s_list)
            rhs[m] = rhs[m] + simpson_2d(integrand, load_quad_points, support_box(m, coords, cells))
    return rhs


def torch_fem_2d(cells, bc_nodes, mesh_points, quad_points, load_quad_points, c_list, s_list):
    """(coeffs [N, 1], sol on the evaluation grid) for one mesh (difFEM_2d.py:345-372): stiffness matrix
    densified, boundary rows replaced by identity rows, load vector, dense solve, interpolation."""
    cells = torch.as_tensor(cells, dtype=torch.long)
    bcn = torch.as_tensor(bc_nodes, dtype=torch.long)
    N = mesh_points.shape[0]
    A, _, _ = build_stiffness(cells, mesh_points, N)
    A = A.to_dense()
    A[bcn, :] = torch.zeros([bcn.numel(), N])
    A[bcn, bcn] = 1
    rhs = build_load_vector(cells, bcn, mesh_points, N, load_quad_points, c_list, s_list)
    coeffs = torch.linalg.solve(A, rhs)
    sol = quad_points[0] * 0.0
    for m in range(N):
        sol = sol + coeffs[m] * phim(quad_points, m, mesh_points, cells)
    return coeffs, sol
