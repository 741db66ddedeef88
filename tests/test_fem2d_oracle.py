"""The 2-D FEM oracle (oracle/fem2d_oracle.py) against the fixtures minted from the reference's own
firedrake_difFEM/difFEM_2d.py (oracle/ref_harness/make_golden_fem2d.py).  First step of scope row f1 in
2-D: no CUDA kernel exists for it yet; the quadrature (torchquad) part of the parity is unpinned, see the
oracle's header."""
import glob
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import fem2d_oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden_fem2d", "fem2d_*.pt")))


def _run(fx, dtype=torch.float32):
    mesh = fx["mesh"].clone().to(dtype).requires_grad_(True)
    Q = int(fx["eval_points"])
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    c_list = [c.clone() for c in fx["centers"]]
    s_list = [s.clone() for s in fx["scales"]]
    coeffs, sol = O.torch_fem_2d(fx["cells"], fx["bc_nodes"], mesh, [X, Y], int(fx["load_quad_points"]), c_list, s_list)
    loss = F.mse_loss(sol, O.u_true(torch.stack([X, Y]), c_list, s_list))
    loss.backward()
    return coeffs, sol, loss, mesh.grad


def test_fixtures_present():
    assert len(GOLDEN) == 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[6:-3] for p in GOLDEN])
def test_oracle_reproduces_reference_bit_for_bit(path):
    fx = torch.load(path)
    coeffs, sol, loss, grad = _run(fx)
    assert torch.equal(coeffs.detach(), fx["coeffs"])
    assert torch.equal(sol.detach(), fx["sol"])
    assert float(loss.item()) == fx["loss"]
    assert torch.equal(grad, fx["grad_mesh"])


def test_simpson_rule_properties():
    """The restated torchquad rule: points per dimension (odd, >= 3), exactness for cubics per dimension,
    and the point order the reference's integrands rely on (dimension 0 slowest)."""
    assert [O.simpson_points_per_dim(N) for N in (1, 9, 100, 121, 441, 50000)] == [3, 3, 9, 11, 21, 223]
    seen = {}

    def fn(p):
        seen["p"] = p
        return p[:, 0] ** 3 * p[:, 1] ** 2 + 2.0

    val = O.simpson_2d(fn, 81, [[0.0, 2.0], [1.0, 3.0]])
    assert seen["p"].shape == (81, 2) and seen["p"][1, 0] == seen["p"][0, 0] and seen["p"][1, 1] > seen["p"][0, 1]
    exact = (2.0 ** 4 / 4) * ((3.0 ** 3 - 1.0) / 3) + 2.0 * 4.0
    assert abs(float(val) - exact) <= 1e-5 * exact


def test_basis_is_a_partition_of_unity_and_interpolates():
    """phim (difFEM_2d.py:28-60): the hat functions sum to one on the domain (edge / vertex repeats are
    divided out) and phi_m(x_n) = delta_mn."""
    fx = torch.load(GOLDEN[1])
    cells, mesh = fx["cells"], fx["mesh"]
    x0 = torch.linspace(0, 1, 17)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    total = sum(O.phim([X, Y], m, mesh, cells) for m in range(mesh.shape[0]))
    assert (total - 1.0).abs().max().item() <= 1e-5
    at_nodes = torch.stack([O.phim(mesh.t().contiguous(), m, mesh, cells) for m in range(mesh.shape[0])])
    assert (at_nodes - torch.eye(mesh.shape[0])).abs().max().item() <= 1e-6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[6:-3] for p in GOLDEN])
def test_vectorised_formulation_and_adjoint_match_the_reference(path):
    """oracle/fem2d_fast.py -- closed-form stiffness, padded star table, all nodes and points at once, and the
    hand-derived adjoint: the arithmetic planned for the kernels -- against the fixtures of the reference.

    The derivative of a hat function with respect to the vertices jumps across element edges, and the Simpson
    grid over a star's bounding box / the evaluation grid put points exactly ON edges, where the `>=` / `<=`
    tests of `phim` decide which cells count.  The formulation therefore reproduces the reference's decisions:
    cubature points bit for bit (linspace_rows: one fused multiply-add per point, as ATen does) and both sides of
    every edge test rounded separately before the comparison.  With that the fp32 gradient agrees with the
    reference to rounding even on the uniform mesh, where EVERY grid point lies on an edge; the fp64 evaluation
    of the same formulas breaks the ties differently and is up to 6e-3 (jittered) or O(1) (uniform) away."""
    from oracle import fem2d_fast as Fz
    fx = torch.load(path)
    Q, K = int(fx["eval_points"]), int(fx["load_quad_points"])
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    scale_c, scale_g = fx["coeffs"].abs().max().item(), fx["grad_mesh"].abs().max().item()
    res = {}
    for dt in (torch.float32, torch.float64):
        mesh = fx["mesh"].clone().to(dt).requires_grad_(True)
        tgt = Fz.u_true(torch.stack([X, Y], dim=-1).to(dt), fx["centers"], fx["scales"])
        coeffs, sol = Fz.fem2d_fast(fx["cells"], fx["bc_nodes"], mesh, [X.to(dt), Y.to(dt)], K, fx["centers"], fx["scales"])
        loss = F.mse_loss(sol, tgt)
        loss.backward()
        u2, sol2, g2 = Fz.fem2d_forward_backward(fx["cells"], fx["bc_nodes"], mesh.detach(), [X.to(dt), Y.to(dt)], K,
                                                 fx["centers"], fx["scales"], lambda s_: 2 * (s_ - tgt) / s_.numel())
        res[dt] = (coeffs.detach(), sol.detach(), float(loss), mesh.grad, u2, sol2, g2)
    c32, s32, l32, g32, u32, sol32, adj32 = res[torch.float32]
    _, _, _, g64, _, sol64b, adj64 = res[torch.float64]
    # forward, fp32, against the reference
    assert (c32 - fx["coeffs"]).abs().max().item() <= 2e-5 * scale_c
    assert (s32 - fx["sol"]).abs().max().item() <= 2e-5 * scale_c
    assert abs(l32 - fx["loss"]) <= 1e-4 * fx["loss"]
    assert (u32 - c32).abs().max().item() <= 2e-6 * scale_c and (sol32 - s32).abs().max().item() <= 2e-6 * scale_c
    # gradient, fp32, against the reference: autograd through the formulation and the hand-derived adjoint
    assert (g32 - fx["grad_mesh"]).abs().max().item() <= 5e-5 * scale_g
    assert (adj32 - fx["grad_mesh"]).abs().max().item() <= 5e-5 * scale_g
    # the adjoint formulas are exact: fp64 against autograd
    assert (adj64 - g64).abs().max().item() <= 1e-12 * scale_g


def test_linspace_rows_is_torch_linspace_bit_for_bit():
    from oracle import fem2d_fast as Fz
    torch.manual_seed(0)
    for n in (3, 9, 11, 21, 223):
        lo = torch.rand(64)
        hi = lo + torch.rand(64)
        rows = Fz.linspace_rows(lo, hi, n)
        for r in range(64):
            assert torch.equal(rows[r], torch.linspace(float(lo[r]), float(hi[r]), n))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[6:-3] for p in GOLDEN])
def test_kernel_arithmetic_on_the_host_matches_the_reference(path):
    """g_adaptivity_b200/csrc/fem2d_math.cuh -- the host/device functions the 2-D FEM kernels are written with --
    run sequentially on the CPU (oracle/fem2d_host.cpp: triangle geometry, Simpson load vector with the
    reference's tie handling, matrix-free conjugate gradients on the interior SPD system, point-wise
    interpolation, hand-derived adjoint) against the fixtures of the reference: forward 1e-5, gradient 5e-5."""
    import ctypes
    import numpy as np
    from oracle import build_host, fem2d_fast as Fz
    lib = ctypes.CDLL(build_host.build())
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    fx = torch.load(path)
    N = fx["mesh"].shape[0]
    cells = fx["cells"].numpy().astype(np.int32)
    is_bc = np.zeros(N, np.uint8)
    is_bc[fx["bc_nodes"].numpy()] = 1
    sc_, sl_ = Fz.star_table(fx["cells"], N)
    star_cell, star_loc = sc_.numpy().astype(np.int32), sl_.numpy().astype(np.int32)
    coords = fx["mesh"].numpy().astype(np.float32).copy()
    cen, scl = fx["centers"].numpy().astype(np.float64).copy(), fx["scales"].numpy().astype(np.float64).copy()
    Q = int(fx["eval_points"])
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    ex, ey = X.reshape(-1).numpy().copy(), Y.reshape(-1).numpy().copy()
    coeffs, sol, grad = np.zeros(N, np.float32), np.zeros(Q * Q, np.float32), np.zeros((N, 2), np.float32)
    iters = ctypes.c_int(0)
    args = [ptr(cells), cells.shape[0], ptr(is_bc), N, ptr(star_cell), ptr(star_loc), star_cell.shape[1], ptr(coords), ptr(cen),
            ptr(scl), cen.shape[0], int(fx["load_quad_points"]), ptr(ex), ptr(ey), Q * Q]
    assert lib.fem2d_host(*args, None, ptr(coeffs), ptr(sol), None, ctypes.byref(iters)) == 0
    tgt = Fz.u_true(torch.stack([X, Y], dim=-1), fx["centers"], fx["scales"]).reshape(-1).numpy()
    g_sol = (2 * (sol - tgt) / sol.size).astype(np.float32)             # d mse_loss / d sol
    assert lib.fem2d_host(*args, ptr(g_sol), ptr(coeffs), ptr(sol), ptr(grad), ctypes.byref(iters)) == 0
    scale_c, scale_g = fx["coeffs"].abs().max().item(), fx["grad_mesh"].abs().max().item()
    assert np.abs(coeffs - fx["coeffs"].reshape(-1).numpy()).max() <= 1e-5 * scale_c
    assert np.abs(sol - fx["sol"].reshape(-1).numpy()).max() <= 1e-5 * scale_c
    assert np.abs(grad - fx["grad_mesh"].numpy()).max() <= 5e-5 * scale_g
    assert 0 < iters.value <= 40 * N


def _emu_run(lib, cells, bc_nodes, mesh, centers, scales, K, Q, B=1):
    import ctypes
    import numpy as np
    from oracle import fem2d_fast as Fz
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    N = mesh.shape[0]
    cells32 = np.ascontiguousarray(cells, dtype=np.int32)
    is_bc = np.zeros(N, np.uint8)
    is_bc[np.asarray(bc_nodes)] = 1
    sc_, sl_ = Fz.star_table(torch.as_tensor(cells, dtype=torch.long), N)
    star_cell, star_loc = sc_.numpy().astype(np.int32), sl_.numpy().astype(np.int32)
    coords = np.stack([np.asarray(mesh, dtype=np.float32)] * B).copy()
    cen = np.stack([np.asarray(centers, dtype=np.float64)] * B).copy()
    scl = np.stack([np.asarray(scales, dtype=np.float64)] * B).copy()
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    ex, ey = X.reshape(-1).numpy().copy(), Y.reshape(-1).numpy().copy()
    coeffs, sol = np.zeros((B, N), np.float32), np.zeros((B, Q * Q), np.float32)
    grad, u64, iters = np.zeros((B, N, 2), np.float32), np.zeros((B, N), np.float64), np.zeros(B, np.int32)
    args = [ptr(cells32), cells32.shape[0], ptr(is_bc), N, ptr(star_cell), ptr(star_loc), star_cell.shape[1], ptr(coords), ptr(cen),
            ptr(scl), cen.shape[1], B, int(K), ptr(ex), ptr(ey), Q * Q]
    assert lib.fem2d_emu(*args, None, ptr(coeffs), ptr(sol), ptr(u64), None, ptr(iters)) == 0
    tgt = Fz.u_true(torch.stack([X, Y], dim=-1), torch.as_tensor(centers), torch.as_tensor(scales)).reshape(-1).numpy()
    g_sol = (2 * (sol - tgt[None]) / sol[0].size).astype(np.float32)
    assert lib.fem2d_emu(*args, ptr(g_sol), ptr(coeffs), ptr(sol), ptr(u64), ptr(grad), None) == 0
    return coeffs, sol, grad, iters, g_sol, (cells32, is_bc, star_cell, star_loc, coords, cen, scl, ex, ey)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[6:-3] for p in GOLDEN])
def test_kernels_emulated_on_cpu_threads_match_the_reference(path):
    """csrc/fem2d.cu ITSELF compiled for the CPU (oracle/fem2d_emu.cpp + oracle/cuda_emu.h: one std::thread per
    CUDA thread, pthread barriers for __syncthreads / warp shuffles, CAS for atomicAdd): k_fem2d_fwd and
    k_fem2d_bwd, a batch of two meshes, against the fixtures of the reference.  Checks the kernels' indexing,
    shared-memory carving, phase order and barriers without a GPU (the GPU run is tests/test_fem2d_gpu.py)."""
    import ctypes
    import numpy as np
    from oracle import build_host
    lib = ctypes.CDLL(build_host.build_emu())
    fx = torch.load(path)
    coeffs, sol, grad, iters, _, _ = _emu_run(lib, fx["cells"].numpy(), fx["bc_nodes"].numpy(), fx["mesh"].numpy(), fx["centers"].numpy(),
                                              fx["scales"].numpy(), fx["load_quad_points"], int(fx["eval_points"]), B=2)
    scale_c, scale_g = fx["coeffs"].abs().max().item(), fx["grad_mesh"].abs().max().item()
    for b in range(2):
        assert np.abs(coeffs[b] - fx["coeffs"].reshape(-1).numpy()).max() <= 1e-5 * scale_c
        assert np.abs(sol[b] - fx["sol"].reshape(-1).numpy()).max() <= 1e-5 * scale_c
        assert np.abs(grad[b] - fx["grad_mesh"].numpy()).max() <= 5e-5 * scale_g
    assert iters[0] == iters[1] > 0


def test_emulated_kernels_equal_the_sequential_harness_beyond_one_thread_per_node():
    """20x20 mesh (400 nodes, 722 cells > 256 threads: every strided loop wraps): emulated kernels against the
    sequential host run of the same arithmetic."""
    import ctypes
    import numpy as np
    from oracle import build_host
    from oracle.ref_harness.make_golden_fem2d import case_inputs
    emu, host = ctypes.CDLL(build_host.build_emu()), ctypes.CDLL(build_host.build())
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    n, K, Q = 20, 101, 23
    cells, bc, pts, centers, scales = case_inputs(n, 2, 0.3, 7)
    coeffs, sol, grad, iters, g_sol, (cells32, is_bc, star_cell, star_loc, coords, cen, scl, ex, ey) = _emu_run(
        emu, cells, bc, pts, centers, scales, K, Q)
    N = pts.shape[0]
    c2, s2, g2 = np.zeros(N, np.float32), np.zeros(Q * Q, np.float32), np.zeros((N, 2), np.float32)
    it2 = ctypes.c_int(0)
    assert host.fem2d_host(ptr(cells32), cells32.shape[0], ptr(is_bc), N, ptr(star_cell), ptr(star_loc), star_cell.shape[1],
                           ptr(coords[0]), ptr(cen[0]), ptr(scl[0]), cen.shape[1], K, ptr(ex), ptr(ey), Q * Q, ptr(g_sol[0]),
                           ptr(c2), ptr(s2), ptr(g2), ctypes.byref(it2)) == 0
    assert np.abs(coeffs[0] - c2).max() <= 2e-6 * np.abs(c2).max()
    assert np.abs(sol[0] - s2).max() <= 2e-6 * np.abs(c2).max()
    assert np.abs(grad[0] - g2).max() <= 1e-5 * np.abs(g2).max()
