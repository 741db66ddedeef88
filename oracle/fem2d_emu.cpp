// csrc/fem2d.cu compiled for the CPU: the kernels k_fem2d_fwd / k_fem2d_bwd run block by block with one std::thread
// per CUDA thread (oracle/cuda_emu.h).  TEST INFRASTRUCTURE ONLY: it checks the kernels' indexing, shared-memory
// carving, phase order and barriers against the reference's fixtures while no GPU is at hand.
#define FEM2D_EMULATE 1
#include "cuda_emu.h"

#include <thread>
#include <vector>

thread_local EmuDim threadIdx, blockIdx;
EmuDim blockDim;
pthread_barrier_t emu_block_barrier;
pthread_barrier_t emu_warp_barrier[32];
double emu_warp_buf[32][32];
namespace gad {
inline namespace emu {
alignas(16) unsigned char f2_raw[256 * 1024];      // the block's "shared memory" (one block at a time)
}
}  // namespace gad

#include "../g_adaptivity_b200/csrc/fem2d.cu"

namespace {
template <typename K>
void run_grid(K kernel, const gad::F2Args& a, int B) {
    const unsigned nt = gad::FEM2D_THREADS;
    blockDim = EmuDim{nt, 1, 1};
    for (int b = 0; b < B; ++b) {
        pthread_barrier_init(&emu_block_barrier, nullptr, nt);
        for (unsigned w = 0; w < nt / 32; ++w) pthread_barrier_init(&emu_warp_barrier[w], nullptr, 32);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                threadIdx = EmuDim{t, 0, 0};
                blockIdx = EmuDim{(unsigned)b, 0, 0};
                kernel(a);
            });
        for (auto& x : th) x.join();
        pthread_barrier_destroy(&emu_block_barrier);
        for (unsigned w = 0; w < nt / 32; ++w) pthread_barrier_destroy(&emu_warp_barrier[w]);
    }
}
}  // namespace

extern "C" int fem2d_emu(const int* cells, int T, const unsigned char* is_bc, int N, const int* star_cell, const int* star_loc, int D,
                         const float* coords, const double* cen, const double* sc, int G, int B, int load_quad_points, const float* ex,
                         const float* ey, int Q, const float* g_sol, float* coeffs, float* sol, double* u64, float* grad, int* cg_iters) {
    if (gad::f2_smem_bytes(N, T, D) > sizeof(gad::f2_raw)) return 1;
    gad::F2Args a = {};
    a.cells = cells, a.is_bc = is_bc, a.star_cell = star_cell, a.star_loc = star_loc, a.coords = coords, a.cen = cen, a.sc = sc;
    a.ex = ex, a.ey = ey, a.g_sol = g_sol, a.coeffs = coeffs, a.sol = sol, a.u64 = u64, a.grad = grad, a.cg_iters = cg_iters;
    a.T = T, a.N = N, a.D = D, a.G = G, a.K = load_quad_points, a.Q = Q;
    const int R = gad::f2_rows_per_thread(N);
    if (!g_sol) {
        if (R == 1) run_grid(gad::k_fem2d_fwd<1>, a, B);
        else if (R == 2) run_grid(gad::k_fem2d_fwd<2>, a, B);
        else if (R == 4) run_grid(gad::k_fem2d_fwd<4>, a, B);
        else run_grid(gad::k_fem2d_fwd<8>, a, B);
    } else {
        if (R == 1) run_grid(gad::k_fem2d_bwd<1>, a, B);
        else if (R == 2) run_grid(gad::k_fem2d_bwd<2>, a, B);
        else if (R == 4) run_grid(gad::k_fem2d_bwd<4>, a, B);
        else run_grid(gad::k_fem2d_bwd<8>, a, B);
    }
    return 0;
}
