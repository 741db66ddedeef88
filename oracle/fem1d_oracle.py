"""CPU restatement of the 1-D differentiable FEM solve that follows the deformer when
`loss_type='pde_loss'` (scope row f1 of SURVEY section 8): `torch_FEM_1D` of
/root/reference/firedrake_difFEM/difFEM_1d.py:211-238 with `build_stiffness_matrix` (:83-131,
vectorised branch), `build_load_vector` (:134-155), `soln` (:62-79), `f` (:32-36),
`u_true_exact_1d` (:38-48).  TEST INFRASTRUCTURE ONLY: imported by tests/ (and by nothing in the
product path); pinned against the reference's own file by tests/golden/fem1d_*.pt
(oracle/ref_harness/make_golden_fem1d.py runs difFEM_1d.py in place behind a matplotlib stub).

Two forms: `torch_fem_1d` is the line-by-line restatement (any dtype, differentiable through
torch autograd, dense `torch.linalg.solve` like the reference); `fem1d_adjoint` is the hand-derived
adjoint the CUDA kernels implement, in torch, so that the derivation is checked against autograd on
the CPU before any kernel runs."""
from __future__ import annotations

import torch


def f_forcing(x, c_list, s_list):          # difFEM_1d.py:32-36  (= u''_true)
    sol = torch.zeros_like(x)
    for c, s in zip(c_list, s_list):
        sol = sol + -2 * torch.exp(-(x - c) ** 2 / s ** 2) * (s ** 2 - 2 * (x - c) ** 2) / s ** 4
    return sol


def u_true(x, c_list, s_list):             # difFEM_1d.py:38-48
    sol = torch.zeros_like(x)
    for c, s in zip(c_list, s_list):
        sol = sol + torch.exp(-(x - c) ** 2 / s ** 2)
    return sol


def build_stiffness_matrix(mesh_points, stiff_quad_points=3):   # difFEM_1d.py:83-120
    k = stiff_quad_points
    mesh_diffs = torch.diff(mesh_points)
    L_start = mesh_points[:-1].view(-1, 1)
    steps = torch.arange(k + 1, dtype=mesh_points.dtype).view(1, -1).repeat(L_start.shape[0], 1)
    mesh_quad = L_start + steps * mesh_diffs.view(-1, 1) / k
    a, b = mesh_points[:-1], mesh_points[1:]
    L_dphi = (1 / (b - a)).view(-1, 1).repeat(1, k + 1)
    R_dphi = -L_dphi
    off_diags = torch.trapezoid(L_dphi * R_dphi, mesh_quad)
    internal_diag = torch.trapezoid(L_dphi[:-1] ** 2, mesh_quad[:-1]) + torch.trapezoid(R_dphi[1:] ** 2, mesh_quad[1:])
    LHS = torch.trapezoid(L_dphi[0] ** 2, mesh_quad[0])
    RHS = torch.trapezoid(R_dphi[-1] ** 2, mesh_quad[-1])
    n = mesh_points.shape[0]
    A = torch.zeros(n, n, dtype=mesh_points.dtype)
    A[1:-1, 1:-1] = torch.diag(internal_diag)
    A = A + torch.diag(off_diags, 1) + torch.diag(off_diags, -1)
    A[0, 0] = LHS
    A[-1, -1] = RHS
    return A


def build_load_vector(mesh, c_list, s_list, load_quad_points):   # difFEM_1d.py:134-155
    k = load_quad_points
    n = mesh.shape[0]
    diffs = torch.diff(mesh)
    L_start = mesh[:-1].view(-1, 1)
    ar = torch.arange(k, dtype=mesh.dtype).view(1, -1)
    phis = ar / (k - 1)
    x_vec = L_start + diffs.view(-1, 1) * ar.repeat(L_start.shape[0], 1) / (k - 1)
    f_vec = f_forcing(x_vec, c_list, s_list)
    left = torch.trapezoid(f_vec * phis, x_vec)
    right = torch.trapezoid(f_vec * torch.flip(phis, dims=[1]), x_vec)
    RHS = torch.zeros(n, dtype=mesh.dtype)
    RHS = RHS + torch.cat([torch.zeros(1, dtype=mesh.dtype), left]) + torch.cat([right, torch.zeros(1, dtype=mesh.dtype)])
    return RHS.unsqueeze(-1)


def soln(out, mesh, BC1, BC2, quad_points, num_solpoints):   # difFEM_1d.py:62-79
    out = out.squeeze()
    ext = torch.cat([BC1, out, BC2])
    gradients = (ext[1:] - ext[:-1]) / (mesh[1:] - mesh[:-1])
    idx = torch.searchsorted(mesh.detach().contiguous(), quad_points.contiguous(), right=False) - 1
    idx = torch.clamp(idx, 0, num_solpoints - 1)
    return ext[idx] + gradients[idx] * (quad_points - mesh[idx])


def torch_fem_1d(mesh_points, quad_points, c_list, s_list, load_quad_points=101, stiff_quad_points=3):
    """difFEM_1d.py:211-238 -> (coeffs [n-2, 1], sol [Q], BC1, BC2)."""
    n = mesh_points.shape[0]
    A = build_stiffness_matrix(mesh_points, stiff_quad_points)
    A_int = -A[1:-1, 1:-1]
    # the reference evaluates the Dirichlet values on `torch.tensor([mesh_points[0]])` (:221-222), a
    # fresh tensor: NO gradient flows from the boundary values to the end points (kept as is)
    BC1 = u_true(mesh_points[:1].detach(), c_list, s_list)
    BC2 = u_true(mesh_points[-1:].detach(), c_list, s_list)
    RHS = build_load_vector(mesh_points, c_list, s_list, load_quad_points)
    RHS_int = RHS[1:-1].clone()
    RHS_int[0] = RHS_int[0] + BC1 * A[0, 1]
    RHS_int[-1] = RHS_int[-1] + A[-1, -2] * BC2
    coeffs = torch.linalg.solve(A_int, RHS_int)
    sol = soln(coeffs, mesh_points, BC1, BC2, quad_points, num_solpoints=n)
    return coeffs, sol, BC1, BC2


# ------------------------------------------------------------------------------------------
# the form the CUDA kernels implement: tridiagonal (Thomas) solve + hand-derived adjoint
# ------------------------------------------------------------------------------------------
def _thomas(lower, diag, upper, rhs):
    n = diag.shape[0]
    cp, dp = torch.zeros_like(diag), torch.zeros_like(diag)
    cp[0] = upper[0] / diag[0]
    dp[0] = rhs[0] / diag[0]
    for i in range(1, n):
        den = diag[i] - lower[i] * cp[i - 1]
        cp[i] = upper[i] / den if i < n - 1 else 0.0
        dp[i] = (rhs[i] - lower[i] * dp[i - 1]) / den
    x = torch.zeros_like(diag)
    x[-1] = dp[-1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


def df_forcing(x, c_list, s_list):          # f' = u'''_true
    sol = torch.zeros_like(x)
    for c, s in zip(c_list, s_list):
        d = x - c
        sol = sol + torch.exp(-d ** 2 / s ** 2) * (12 * d / s ** 4 - 8 * d ** 3 / s ** 6)
    return sol


def du_true(x, c_list, s_list):
    sol = torch.zeros_like(x)
    for c, s in zip(c_list, s_list):
        sol = sol + torch.exp(-(x - c) ** 2 / s ** 2) * (-2 * (x - c) / s ** 2)
    return sol


def fem1d_forward_tridiag(x, quad, c_list, s_list, K):
    """Same numbers as torch_fem_1d (up to rounding) through closed-form entries: returns a dict."""
    n = x.shape[0]
    h = x[1:] - x[:-1]
    t = torch.arange(K, dtype=x.dtype) / (K - 1)
    w = torch.ones(K, dtype=x.dtype)
    w[0] = w[-1] = 0.5
    p = x[:-1, None] + h[:, None] * t[None, :]
    fv = f_forcing(p, c_list, s_list)
    left = h / (K - 1) * (w * fv * t).sum(1)
    right = h / (K - 1) * (w * fv * (1 - t)).sum(1)
    RHS = torch.zeros(n, dtype=x.dtype)
    RHS[1:] += left
    RHS[:-1] += right
    BC1, BC2 = u_true(x[:1], c_list, s_list)[0], u_true(x[-1:], c_list, s_list)[0]
    b = RHS[1:-1].clone()
    b[0] += -BC1 / h[0]
    b[-1] += -BC2 / h[-1]
    diag = -(1 / h[:-1] + 1 / h[1:])
    off = 1 / h[1:-1]                      # between internal nodes i and i + 1  (i = 1 .. n-3)
    lower = torch.cat([torch.zeros(1, dtype=x.dtype), off])
    upper = torch.cat([off, torch.zeros(1, dtype=x.dtype)])
    u_int = _thomas(lower, diag, upper, b)
    u = torch.cat([BC1.view(1), u_int, BC2.view(1)])
    idx = torch.clamp(torch.searchsorted(x.contiguous(), quad.contiguous(), right=False) - 1, 0, n - 1)
    r = (quad - x[idx]) / h[idx]
    sol = u[idx] + (u[idx + 1] - u[idx]) * r
    return dict(h=h, u=u, sol=sol, idx=idx, r=r, lower=lower, diag=diag, upper=upper, BC1=BC1, BC2=BC2, t=t, w=w, p=p)


def fem1d_adjoint(x, quad, c_list, s_list, K, g_sol):
    """dL/dx given dL/dsol (hand-derived; what csrc/fem1d.cu implements)."""
    fw = fem1d_forward_tridiag(x, quad, c_list, s_list, K)
    n = x.shape[0]
    h, u, idx, r, t, w, p = fw["h"], fw["u"], fw["idx"], fw["r"], fw["t"], fw["w"], fw["p"]
    g_u = torch.zeros_like(u)
    g_x = torch.zeros_like(x)
    du = u[idx + 1] - u[idx]
    g_u.index_add_(0, idx, g_sol * (1 - r))
    g_u.index_add_(0, idx + 1, g_sol * r)
    g_x.index_add_(0, idx, g_sol * du * (r - 1) / h[idx])
    g_x.index_add_(0, idx + 1, -g_sol * du * r / h[idx])
    lam = _thomas(fw["lower"], fw["diag"], fw["upper"], g_u[1:-1])      # A_int symmetric
    lam_full = torch.cat([torch.zeros(1, dtype=x.dtype), lam, torch.zeros(1, dtype=x.dtype)])
    ui = u.clone()
    ui[0] = 0.0
    ui[-1] = 0.0                                                          # internal coefficients, 0 at the ends
    # g_h from the matrix entries: g_A[i, j] = -lam_i u_j over the tridiagonal of the INTERNAL system
    g_h = torch.zeros_like(h)
    k = torch.arange(n - 1)
    inv2 = 1 / h ** 2
    g_h += (-lam_full[k] * ui[k]) * inv2                                  # diag of node k   (h_k as the right interval)
    g_h += (-lam_full[k + 1] * ui[k + 1]) * inv2                          # diag of node k+1 (h_k as the left interval)
    g_h += (-lam_full[k] * ui[k + 1] - lam_full[k + 1] * ui[k]) * (-inv2)  # off-diagonal pair (zero when an end is involved)
    # boundary adjustments of the right-hand side
    g_BC1 = lam_full[1] * (-1 / h[0])
    g_BC2 = lam_full[n - 2] * (-1 / h[-1])
    g_h[0] += lam_full[1] * fw["BC1"] / h[0] ** 2
    g_h[-1] += lam_full[n - 2] * fw["BC2"] / h[-1] ** 2
    # (g_BC1 + g_u[0], g_BC2 + g_u[-1] would flow into x_0 / x_{n-1} through u_true'; the reference
    # detaches the boundary values, difFEM_1d.py:221-222, so they are dropped)
    del g_BC1, g_BC2
    # load vector: left_k -> node k+1, right_k -> node k
    fv, dfv = f_forcing(p, c_list, s_list), df_forcing(p, c_list, s_list)
    gl, gr = lam_full[1:], lam_full[:-1]
    S0L, S0R = (w * fv * t).sum(1), (w * fv * (1 - t)).sum(1)
    S1La, S1Lb = (w * dfv * t * (1 - t)).sum(1), (w * dfv * t * t).sum(1)
    S1Ra, S1Rb = (w * dfv * (1 - t) * (1 - t)).sum(1), (w * dfv * (1 - t) * t).sum(1)
    s = 1.0 / (K - 1)
    dL_dxk = -s * S0L + h * s * S1La
    dL_dxk1 = s * S0L + h * s * S1Lb
    dR_dxk = -s * S0R + h * s * S1Ra
    dR_dxk1 = s * S0R + h * s * S1Rb
    g_x[:-1] += gl * dL_dxk + gr * dR_dxk
    g_x[1:] += gl * dL_dxk1 + gr * dR_dxk1
    g_x[1:] += g_h
    g_x[:-1] -= g_h
    return g_x, fw
