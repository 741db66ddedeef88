// C-ABI entry points of the deformer (include/gadapt.h) and the dispatch between the
// mesh-resident kernels (fused_kernels.cu) and the streaming kernels (stream_kernels.cu).
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace gad {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int g_sm_count = 0, g_smem_optin = 0, g_l2 = 0;

static void query_device() {
    if (g_sm_count) return;
    int dev = 0;
    cudaGetDevice(&dev);
    int sm = 0, sh = 0, l2 = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&sh, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
    g_smem_optin = sh > 0 ? sh : 48 * 1024;
    g_l2 = l2;
    g_sm_count = sm > 0 ? sm : 148;
}

int sm_count() {
    query_device();
    return g_sm_count;
}
int smem_optin_bytes() {
    query_device();
    return g_smem_optin;
}

// stream_kernels.cu
template <int CE>
int stream_forward(const int32_t*, const int32_t*, int64_t, const float*, int, const float*, int, const float*, int, int,
                   float*, float*, float*, cudaStream_t);
template <int CE>
int stream_backward(const int32_t*, const int32_t*, const int32_t*, const int32_t*, int64_t, const float*, const float*,
                    int, const float*, int, const float*, int, float, float*, float*, float*, float*, cudaStream_t);
size_t stream_fwd_ws_floats(int64_t N, int CE, int method);
size_t stream_bwd_ws_floats(int64_t N, int CE);

// fused_kernels.cu
int fused_forward(int CE, const int32_t* rowptr, const int32_t* col, int64_t N, const int32_t* tile_ptr, int T,
                  int max_tile_nodes, int max_tile_edges, const float* x0, int dim, const float* Mu, int Lw,
                  const float* tau, int L, int method, float* x_phys, float* states, cudaStream_t st);
int fused_backward(int CE, const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_dst,
                   int64_t N, const int32_t* tile_ptr, int T, int max_tile_nodes, int max_tile_edges,
                   const float* states, const float* g_xphys, int dim, const float* Mu, int Lw, const float* tau, int L,
                   float* gMu, float* g_tau, float* g_x0, float* ws, size_t ws_floats, cudaStream_t st);
size_t fused_bwd_ws_floats(int CE, int T, int L);

}  // namespace gad

using namespace gad;

extern "C" int gad_version(void) { return 100; }

extern "C" const char* gad_last_error(void) { return g_err; }

extern "C" long long gad_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int gad_device_info(int* host_sm_count, int* host_smem_optin_bytes, int* host_l2_bytes) {
    int n = 0;
    GAD_CUDA(cudaGetDeviceCount(&n));
    GAD_CHECK_ARG(n > 0, "gad_device_info: no CUDA device");
    query_device();
    if (host_sm_count) *host_sm_count = g_sm_count;
    if (host_smem_optin_bytes) *host_smem_optin_bytes = g_smem_optin;
    if (host_l2_bytes) *host_l2_bytes = g_l2;
    return GAD_OK;
}

extern "C" size_t gad_deform_workspace_bytes(int64_t N, int CE, int method) {
    return stream_fwd_ws_floats(N, CE, method) * sizeof(float);
}

extern "C" size_t gad_deform_bwd_workspace_bytes(int64_t N, int CE, int T, int L) {
    const size_t a = stream_bwd_ws_floats(N, CE);
    const size_t b = fused_bwd_ws_floats(CE, T > 0 ? T : 1, L);
    return (a > b ? a : b) * sizeof(float);
}

extern "C" int gad_deform_fwd(const int32_t* rowptr, const int32_t* col, int64_t N, int64_t E,
                              const int32_t* tile_ptr, int T, int max_tile_nodes, int max_tile_edges,
                              const float* x0, int dim, int CE, const float* Mu, int Lw, const float* tau, int L,
                              int method, float* x_phys, float* states, void* workspace, size_t workspace_bytes,
                              void* stream) {
    GAD_CHECK_ARG(rowptr && col && x0 && Mu && tau && x_phys, "gad_deform_fwd: null pointer");
    GAD_CHECK_ARG(N > 0 && L > 0 && dim >= 1 && dim <= CE && (Lw == 1 || Lw == L),
                  "gad_deform_fwd: N=%lld L=%d dim=%d CE=%d Lw=%d", (long long)N, L, dim, CE, Lw);
    GAD_CHECK_ARG(CE == 2 || CE == 4 || CE == 8, "gad_deform_fwd: unsupported CE=%d", CE);
    GAD_CHECK_ARG(method == GAD_METHOD_EULER || method == GAD_METHOD_RK4, "gad_deform_fwd: unknown method %d", method);
    GAD_CHECK_ARG(!states || states == x0, "gad_deform_fwd: when states is given, x0 must alias states[0]");
    cudaStream_t st = as_stream(stream);
    (void)E;
    if (tile_ptr) {
        GAD_CHECK_ARG(T > 0 && max_tile_nodes > 0, "gad_deform_fwd: bad tiling T=%d max_tile_nodes=%d", T, max_tile_nodes);
        return fused_forward(CE, rowptr, col, N, tile_ptr, T, max_tile_nodes, max_tile_edges, x0, dim, Mu, Lw, tau, L,
                             method, x_phys, states, st);
    }
    GAD_CHECK_ARG(workspace && workspace_bytes >= stream_fwd_ws_floats(N, CE, method) * sizeof(float),
                  "gad_deform_fwd: workspace too small");
    float* ws = reinterpret_cast<float*>(workspace);
    switch (CE) {
        case 2: return stream_forward<2>(rowptr, col, N, x0, dim, Mu, Lw, tau, L, method, x_phys, states, ws, st);
        case 4: return stream_forward<4>(rowptr, col, N, x0, dim, Mu, Lw, tau, L, method, x_phys, states, ws, st);
        default: return stream_forward<8>(rowptr, col, N, x0, dim, Mu, Lw, tau, L, method, x_phys, states, ws, st);
    }
}

extern "C" int gad_deform_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr,
                              const int32_t* t_dst, int64_t N, int64_t E, const int32_t* tile_ptr, int T,
                              int max_tile_nodes, int max_tile_edges, const float* states, const float* g_xphys,
                              int dim, int CE, const float* Mu, int Lw, const float* tau, int L, float* gMu,
                              float* g_tau, float* g_x0, void* workspace, size_t workspace_bytes, void* stream) {
    GAD_CHECK_ARG(rowptr && col && t_rowptr && t_dst && states && g_xphys && Mu && tau && gMu && workspace,
                  "gad_deform_bwd: null pointer");
    GAD_CHECK_ARG(N > 0 && L > 0 && dim >= 1 && dim <= CE && (Lw == 1 || Lw == L),
                  "gad_deform_bwd: N=%lld L=%d dim=%d CE=%d Lw=%d", (long long)N, L, dim, CE, Lw);
    GAD_CHECK_ARG(CE == 2 || CE == 4 || CE == 8, "gad_deform_bwd: unsupported CE=%d", CE);
    cudaStream_t st = as_stream(stream);
    float* ws = reinterpret_cast<float*>(workspace);
    (void)E;
    if (tile_ptr) {
        GAD_CHECK_ARG(T > 0 && max_tile_nodes > 0, "gad_deform_bwd: bad tiling");
        GAD_CHECK_ARG(workspace_bytes >= fused_bwd_ws_floats(CE, T, L) * sizeof(float), "gad_deform_bwd: workspace too small");
        return fused_backward(CE, rowptr, col, t_rowptr, t_dst, N, tile_ptr, T, max_tile_nodes, max_tile_edges, states,
                              g_xphys, dim, Mu, Lw, tau, L, gMu, g_tau, g_x0, ws, workspace_bytes / sizeof(float), st);
    }
    GAD_CHECK_ARG(workspace_bytes >= stream_bwd_ws_floats(N, CE) * sizeof(float), "gad_deform_bwd: workspace too small");
    switch (CE) {
        case 2: return stream_backward<2>(rowptr, col, t_rowptr, t_dst, N, states, g_xphys, dim, Mu, Lw, tau, L, 1.0f, gMu, g_tau, g_x0, ws, st);
        case 4: return stream_backward<4>(rowptr, col, t_rowptr, t_dst, N, states, g_xphys, dim, Mu, Lw, tau, L, 1.0f, gMu, g_tau, g_x0, ws, st);
        default: return stream_backward<8>(rowptr, col, t_rowptr, t_dst, N, states, g_xphys, dim, Mu, Lw, tau, L, 1.0f, gMu, g_tau, g_x0, ws, st);
    }
}
