"""Vectorised formulation of the 2-D FEM solve of oracle/fem2d_oracle.py -- the arithmetic a one-CTA-per-mesh
CUDA kernel will follow, written in torch so that it can be checked against the line-by-line restatement (and
differentiated by autograd) on the CPU before any kernel exists.  TEST INFRASTRUCTURE ONLY.

Differences from the restatement are re-associations only:
  * stiffness entries in closed form: grad(phi_k) = rot90(p_{k+1} - p_{k+2}) / (2 * signed area) instead of a
    3x3 solve per triangle (difFEM_2d.py:63-117);
  * the hat function of node m on a cell of its star is the barycentric coordinate of the vertex, the
    indicator is the sign test of the three edge functions in either orientation (:16-26), all star cells and
    all quadrature points of all nodes at once (padded star table) instead of Python loops (:28-60, 159-203);
  * Simpson weights as one outer product (torchquad composite rule, see fem2d_oracle.simpson_2d);
  * any dtype (fp64 for the conditioning checks)."""
from __future__ import annotations

import torch

from .fem2d_oracle import simpson_points_per_dim


def star_table(cells: torch.Tensor, num_nodes: int):
    """[N, D] cell ids and local vertex indices of the cells around every node, padded with -1 (D = max degree),
    in the order torch.where(cells == m) yields them (ascending cell id)."""
    T = cells.shape[0]
    deg = torch.zeros(num_nodes, dtype=torch.long)
    deg.index_add_(0, cells.reshape(-1), torch.ones(3 * T, dtype=torch.long))
    D = int(deg.max())
    cell_of = torch.full((num_nodes, D), -1, dtype=torch.long)
    loc_of = torch.zeros((num_nodes, D), dtype=torch.long)
    fill = [0] * num_nodes
    for t in range(T):
        for k in range(3):
            m = int(cells[t, k])
            cell_of[m, fill[m]], loc_of[m, fill[m]] = t, k
            fill[m] += 1
    return cell_of, loc_of


def _hat(P, a, b, c):
    """P [..., Q, 2] points, a/b/c [..., 1, 2] vertices: indicator(closed triangle) * barycentric coordinate of c."""
    inside, lin = _hat_parts(P, a, b, c)
    return inside * lin


def _hat_parts(P, a, b, c):
    # the two sides of every edge test are rounded separately and then compared, exactly as difFEM_2d.py:16-20
    # does: points ON an edge (ties) are then decided the same way as in the reference
    def sides(u, v):
        return ((u[..., 1] - v[..., 1]) * P[..., 0] + (v[..., 0] - u[..., 0]) * P[..., 1],
                (u[..., 1] - v[..., 1]) * u[..., 0] + (v[..., 0] - u[..., 0]) * u[..., 1])
    (l1, r1), (l2, r2), (l3, r3) = sides(a, b), sides(b, c), sides(c, a)
    left = (l1 >= r1).to(P.dtype) * (l2 >= r2).to(P.dtype) * (l3 >= r3).to(P.dtype)
    right = (l1 <= r1).to(P.dtype) * (l2 <= r2).to(P.dtype) * (l3 <= r3).to(P.dtype)
    inside = left + right
    lin = 1 + ((P[..., 0] - c[..., 0]) * (a[..., 1] - b[..., 1]) + (P[..., 1] - c[..., 1]) * (b[..., 0] - a[..., 0])) / (
        (a[..., 1] - b[..., 1]) * (c[..., 0] - a[..., 0]) + (c[..., 1] - a[..., 1]) * (b[..., 0] - a[..., 0]))
    return inside, lin


def basis_at(P, nodes, coords, cells, cell_of, loc_of):
    """phi_m(P[m, q]) for m in `nodes`: P [M, Q, 2] -> [M, Q]."""
    cid, k = cell_of[nodes], loc_of[nodes]                      # [M, D]
    valid = cid >= 0
    cid = cid.clamp(min=0)
    tri = cells[cid]                                             # [M, D, 3]
    pick = lambda off: coords[torch.gather(tri, 2, ((k + off) % 3).unsqueeze(-1)).squeeze(-1)].unsqueeze(2)   # [M, D, 1, 2]
    c, a, b = pick(0), pick(2), pick(1)                          # fmod(k-1,3) -> k+2, fmod(k-2,3) -> k+1 (mod 3)
    inc = _hat(P.unsqueeze(1), a, b, c) * valid.unsqueeze(-1).to(P.dtype)    # [M, D, Q]
    out = inc.sum(1)
    rep = (inc > 0).to(P.dtype).sum(1)
    return out / (rep + (rep == 0).to(P.dtype))


def forcing(P, centers, scales):
    """f = laplace(u_true) at P [..., 2]; centers / scales [G, 2]."""
    out = torch.zeros(P.shape[:-1], dtype=P.dtype)
    x, y = P[..., 0], P[..., 1]
    for c, s in zip(centers.to(P.dtype), scales.to(P.dtype)):
        out = out + (1 / (s[0] ** 4 * s[1] ** 4)) * torch.exp(-((c[0] - x) ** 2 / s[0] ** 2) - (c[1] - y) ** 2 / s[1] ** 2) * (
            4 * c[1] ** 2 * s[0] ** 4 - 2 * s[0] ** 2 * s[1] ** 4 + 4 * s[1] ** 4 * (c[0] - x) ** 2
            - 8 * c[1] * s[0] ** 4 * y - 2 * s[0] ** 4 * (s[1] ** 2 - 2 * y ** 2))
    return out


def u_true(P, centers, scales):
    out = torch.zeros(P.shape[:-1], dtype=P.dtype)
    for c, s in zip(centers.to(P.dtype), scales.to(P.dtype)):
        out = out + torch.exp(-(P[..., 0] - c[0]) ** 2 / s[0] ** 2 - (P[..., 1] - c[1]) ** 2 / s[1] ** 2)
    return out


def linspace_rows(lo, hi, n):
    """Row-wise torch.linspace(lo[r], hi[r], n), bit for bit: ATen steps up from `lo` in the first half and down
    from `hi` in the second, each point one fused multiply-add (a kernel does the same with fmaf) -- the cubature
    points must be the reference's to the last bit because points ON element edges are decided by comparisons."""
    step = (hi - lo) / (n - 1)
    i = torch.arange(n, dtype=torch.float64)
    up = (lo.double().unsqueeze(1) + step.double().unsqueeze(1) * i).to(lo.dtype)
    down = (hi.double().unsqueeze(1) - step.double().unsqueeze(1) * (n - 1 - i)).to(lo.dtype)
    return torch.where((torch.arange(n) < n // 2).unsqueeze(0), up, down)


def stiffness(cells, coords):
    """Dense -K (the reference's sign) from closed-form triangle gradients."""
    N = coords.shape[0]
    p = coords[cells]                                            # [T, 3, 2]
    d = p[:, [1, 2, 0]] - p[:, [2, 0, 1]]                        # p_{k+1} - p_{k+2}
    twoA = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1])
    g = torch.stack([d[..., 1], -d[..., 0]], dim=-1) / twoA.view(-1, 1, 1)     # grad phi_k  [T, 3, 2]
    Kloc = 0.5 * twoA.abs().view(-1, 1, 1) * (g @ g.transpose(1, 2))           # [T, 3, 3]
    A = torch.zeros(N, N, dtype=coords.dtype)
    rows = cells.unsqueeze(2).expand(-1, 3, 3).reshape(-1)
    cols = cells.unsqueeze(1).expand(-1, 3, 3).reshape(-1)
    A.index_put_((rows, cols), -Kloc.reshape(-1), accumulate=True)
    return A


def fem2d_fast(cells, bc_nodes, coords, eval_xy, load_quad_points, centers, scales):
    """Same contract as fem2d_oracle.torch_fem_2d: (coeffs [N, 1], sol on the evaluation grid eval_xy = [X, Y])."""
    cells = torch.as_tensor(cells, dtype=torch.long)
    bcn = torch.as_tensor(bc_nodes, dtype=torch.long)
    N, dt = coords.shape[0], coords.dtype
    cell_of, loc_of = star_table(cells, N)
    # ---- matrix with Dirichlet rows
    A = stiffness(cells, coords)
    is_bc = torch.zeros(N, dtype=torch.bool)
    is_bc[bcn] = True
    A = torch.where(is_bc.view(-1, 1), torch.eye(N, dtype=dt), A)
    # ---- load vector: Simpson cubature of phi_m f over the (detached) bounding box of the star of m
    inner = torch.nonzero(~is_bc).reshape(-1)
    n = simpson_points_per_dim(int(load_quad_points), 2)
    star_pts = coords.detach()[cells[cell_of[inner].clamp(min=0)]]               # [M, D, 3, 2]
    pad = (cell_of[inner] < 0).view(len(inner), -1, 1, 1)
    lo = torch.where(pad, torch.full_like(star_pts, float("inf")), star_pts).amin(dim=(1, 2))
    hi = torch.where(pad, torch.full_like(star_pts, float("-inf")), star_pts).amax(dim=(1, 2))
    gx, gy = linspace_rows(lo[:, 0], hi[:, 0], n), linspace_rows(lo[:, 1], hi[:, 1], n)   # [M, n]
    P = torch.stack([gx.unsqueeze(2).expand(-1, n, n), gy.unsqueeze(1).expand(-1, n, n)], dim=-1).reshape(len(inner), n * n, 2)
    w1 = torch.ones(n, dtype=dt)
    w1[1:-1:2], w1[2:-1:2] = 4, 2
    W = (w1.view(-1, 1) * w1.view(1, -1)).reshape(-1)
    h = (hi - lo) / (n - 1)
    vals = basis_at(P, inner, coords, cells, cell_of, loc_of) * forcing(P, centers, scales)
    rhs = torch.zeros(N, dtype=dt)
    rhs = rhs.index_put((inner,), (vals * W).sum(1) * h[:, 0] * h[:, 1] / 9.0)
    rhs = rhs.index_put((bcn,), u_true(coords.detach()[bcn], centers, scales))   # Dirichlet values: no gradient
    coeffs = torch.linalg.solve(A, rhs.unsqueeze(1))
    # ---- interpolation onto the evaluation grid
    X, Y = eval_xy
    E = torch.stack([X.reshape(-1), Y.reshape(-1)], dim=-1).to(dt)
    allnodes = torch.arange(N)
    phi = basis_at(E.unsqueeze(0).expand(N, -1, -1), allnodes, coords, cells, cell_of, loc_of)       # [N, Q]
    sol = (coeffs * phi).sum(0).reshape(X.shape)
    return coeffs, sol


# ---- hand-derived adjoint (what a backward kernel computes; checked against autograd on the CPU) ----------
# For a triangle with vertices p_k, barycentric coordinates l_k(P) and their (constant) gradients g_k:
#     d l_c(P) / d p_v = -l_v(P) g_c          d g_c / d p_v [delta] = -g_v (g_c . delta)
#     d area / d p_v   = area * g_v
# hence for K_ab = area * g_a . g_b:   sum_ab x_a y_b dK_ab/dp_v = area * ((Gx . Gy) g_v - (g_v . Gy) Gx - (g_v . Gx) Gy)
# with Gx = sum_a x_a g_a, Gy = sum_b y_b g_b.  The masks of `phim` (which cells count at a point, and the repeat
# divisor) are piecewise constant and carry no derivative; box and grid of the cubature are detached in the
# reference (bounds_support_jr, difFEM_2d.py:298-309), as are the Dirichlet values (:172).
def _bary(P, p, q, r):
    """barycentric coordinate of vertex r of triangle (p, q, r) at P, and its gradient (constant per triangle)."""
    den = (p[..., 1] - q[..., 1]) * (r[..., 0] - p[..., 0]) + (r[..., 1] - p[..., 1]) * (q[..., 0] - p[..., 0])
    g = torch.stack([(p[..., 1] - q[..., 1]), (q[..., 0] - p[..., 0])], dim=-1) / den.unsqueeze(-1)
    lam = 1 + ((P[..., 0] - r[..., 0]) * g[..., 0] + (P[..., 1] - r[..., 1]) * g[..., 1])
    return lam, g


def _star_parts(P, nodes, coords, cells, cell_of, loc_of):
    """Per (node m, star slot d, point q): counting mask / repeat divisor, the three barycentric coordinates and
    the vertex ids (c = m itself, a, b as in phim) of the slot's cell, gradient of l_c."""
    cid, k = cell_of[nodes], loc_of[nodes]
    valid = (cid >= 0)
    tri = cells[cid.clamp(min=0)]
    vid = lambda off: torch.gather(tri, 2, ((k + off) % 3).unsqueeze(-1)).squeeze(-1)        # [M, D]
    c_id, a_id, b_id = vid(0), vid(2), vid(1)
    c, a, b = (coords[i].unsqueeze(2) for i in (c_id, a_id, b_id))                           # [M, D, 1, 2]
    Pq = P.unsqueeze(1)
    inside, lin = _hat_parts(Pq, a, b, c)
    mult = inside * valid.unsqueeze(-1).to(P.dtype)        # 0, 1, or 2 when both orientation tests hold (point on a vertex line)
    rep = ((mult * lin) > 0).to(P.dtype).sum(1)            # cells that contributed a positive value (phim's divisor)
    rep = rep + (rep == 0).to(P.dtype)
    lc, gc = _bary(Pq, a, b, c)
    la, _ = _bary(Pq, b, c, a)
    lb, _ = _bary(Pq, c, a, b)
    # a point ON the edge opposite to m has value 0 (not counted in rep) but a non-zero derivative: mult keeps it
    return mult, rep, (la, lb, lc), (a_id, b_id, c_id), gc.squeeze(2)


def _scatter_basis_grad(grad, coef, parts):
    """grad[p_v] += sum_{m,d,q} coef[m,q] * d phi_m(P_q)/d p_v  for the three vertices of every star slot."""
    mult, rep, (la, lb, lc), (a_id, b_id, c_id), gc = parts
    wq = (coef / rep).unsqueeze(1) * mult                                   # [M, D, Q]
    for lam, vid_ in ((la, a_id), (lb, b_id), (lc, c_id)):
        S = -(wq * lam).sum(-1)                                             # [M, D]
        grad.index_add_(0, vid_.reshape(-1), (S.unsqueeze(-1) * gc).reshape(-1, 2))


def fem2d_forward_backward(cells, bc_nodes, coords, eval_xy, load_quad_points, centers, scales, loss_grad_fn):
    """(coeffs, sol, d loss / d coords) without autograd.  `loss_grad_fn(sol)` returns d loss / d sol."""
    with torch.no_grad():
        cells = torch.as_tensor(cells, dtype=torch.long)
        bcn = torch.as_tensor(bc_nodes, dtype=torch.long)
        N, dt = coords.shape[0], coords.dtype
        cell_of, loc_of = star_table(cells, N)
        A = stiffness(cells, coords)
        is_bc = torch.zeros(N, dtype=torch.bool)
        is_bc[bcn] = True
        A = torch.where(is_bc.view(-1, 1), torch.eye(N, dtype=dt), A)
        inner = torch.nonzero(~is_bc).reshape(-1)
        n = simpson_points_per_dim(int(load_quad_points), 2)
        star_pts = coords[cells[cell_of[inner].clamp(min=0)]]
        pad = (cell_of[inner] < 0).view(len(inner), -1, 1, 1)
        lo = torch.where(pad, torch.full_like(star_pts, float("inf")), star_pts).amin(dim=(1, 2))
        hi = torch.where(pad, torch.full_like(star_pts, float("-inf")), star_pts).amax(dim=(1, 2))
        gx, gy = linspace_rows(lo[:, 0], hi[:, 0], n), linspace_rows(lo[:, 1], hi[:, 1], n)
        P = torch.stack([gx.unsqueeze(2).expand(-1, n, n), gy.unsqueeze(1).expand(-1, n, n)], dim=-1).reshape(len(inner), n * n, 2)
        w1 = torch.ones(n, dtype=dt)
        w1[1:-1:2], w1[2:-1:2] = 4, 2
        W = (w1.view(-1, 1) * w1.view(1, -1)).reshape(-1)
        h = (hi - lo) / (n - 1)
        wq = W.unsqueeze(0) * (h[:, 0] * h[:, 1] / 9.0).unsqueeze(1) * forcing(P, centers, scales)        # [M, Q]
        load_parts = _star_parts(P, inner, coords, cells, cell_of, loc_of)
        mult, rep, (_, _, lc), _, _ = load_parts
        phi_load = (mult * lc).sum(1) / rep
        rhs = torch.zeros(N, dtype=dt)
        rhs[inner] = (phi_load * wq).sum(1)
        rhs[bcn] = u_true(coords[bcn], centers, scales)
        u = torch.linalg.solve(A, rhs.unsqueeze(1)).squeeze(1)
        X, Y = eval_xy
        E = torch.stack([X.reshape(-1), Y.reshape(-1)], dim=-1).to(dt)
        allnodes = torch.arange(N)
        ev_parts = _star_parts(E.unsqueeze(0).expand(N, -1, -1), allnodes, coords, cells, cell_of, loc_of)
        mult_e, rep_e, (_, _, lc_e), _, _ = ev_parts
        phi = (mult_e * lc_e).sum(1) / rep_e                                                                # [N, Q]
        sol = (u.unsqueeze(1) * phi).sum(0).reshape(X.shape)
        # ---- backward
        g_sol = loss_grad_fn(sol).reshape(-1).to(dt)
        g_u = phi @ g_sol
        lam = torch.linalg.solve(A.t(), g_u.unsqueeze(1)).squeeze(1)
        grad = torch.zeros(N, 2, dtype=dt)
        _scatter_basis_grad(grad, u.unsqueeze(1) * g_sol.unsqueeze(0), ev_parts)           # interpolation
        _scatter_basis_grad(grad, lam[inner].unsqueeze(1) * wq, load_parts)                # load vector
        lam_in = torch.where(is_bc, torch.zeros_like(lam), lam)                            # matrix rows of interior nodes
        p = coords[cells]
        d = p[:, [1, 2, 0]] - p[:, [2, 0, 1]]
        twoA = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1])
        g = torch.stack([d[..., 1], -d[..., 0]], dim=-1) / twoA.view(-1, 1, 1)                              # [T, 3, 2]
        area = 0.5 * twoA.abs()
        Gl = (lam_in[cells].unsqueeze(-1) * g).sum(1)                                                       # [T, 2]
        Gu = (u[cells].unsqueeze(-1) * g).sum(1)
        gv = area.view(-1, 1, 1) * ((Gl * Gu).sum(-1).view(-1, 1, 1) * g
                                    - (g * Gu.unsqueeze(1)).sum(-1, keepdim=True) * Gl.unsqueeze(1)
                                    - (g * Gl.unsqueeze(1)).sum(-1, keepdim=True) * Gu.unsqueeze(1))        # [T, 3, 2]
        grad.index_add_(0, cells.reshape(-1), gv.reshape(-1, 2))
        return u.unsqueeze(1), sol, grad
