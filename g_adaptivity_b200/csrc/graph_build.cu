// K0 -- mesh-graph builder.
//
// Replaces the per-forward graph prologue of the reference (src/GNN.py:206-223: boolean-mask
// filter, appended corner self-loops, optional remove/add_self_loops) and the implicit
// gather/scatter index handling of PyG's propagate (src/GRAND_plus.py:233) with explicit CSR
// (by destination) and CSC (by source) structures, built once per topology and cached by the
// host.  Row order is the STABLE edge-list order, i.e. the order in which the reference's CPU
// scatter_add_ accumulates each destination row; the arrays are compared bit-for-bit with
// torch.sort(stable=True) in tests/test_graph_gpu.py.
//
// Integer work only; everything is coalesced streaming over the edge list except the final
// per-row ordering pass (rows are tiny: mesh in-degree <= 7).  Deterministic by construction:
// atomics are used only on integer counters whose final values are order-independent, and the
// arbitrary slot order they produce inside a row is removed by sorting each row on edge id.
#include "common.cuh"

namespace gad {
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---- exclusive scan (int32), three phases -------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(const int32_t* __restrict__ in,
                                                            int32_t* __restrict__ out,
                                                            int32_t* __restrict__ tile_sums, int64_t n) {
    __shared__ int32_t warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    // inclusive scan of per-thread sums across the warp, then across warps
    int32_t incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int32_t warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += warp_tot[w];
    int32_t run = warp_off + incl - sum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = run;
}

// single block: exclusive scan of the tile sums in place; total -> *total_out (may be null)
__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* __restrict__ sums, int64_t m,
                                                    int32_t* __restrict__ total_out) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t start = 0; start < m; start += 1024) {
        const int64_t i = start + threadIdx.x;
        const int32_t x = (i < m) ? sums[i] : 0;
        int32_t incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int32_t warp_off = 0;
        for (int w = 0; w < warp; ++w) warp_off += warp_tot[w];
        const int32_t carry = carry_s;
        if (i < m) sums[i] = carry + warp_off + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int32_t* __restrict__ out,
                                                           const int32_t* __restrict__ tile_sums,
                                                           int64_t n, int32_t* __restrict__ tail) {
    const int32_t off = tile_sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += off;
    (void)tail;
}

// exclusive scan of in[0..n) into out[0..n); total into *total (device).  `sums` holds
// ceil(n / SCAN_TILE) ints.
cudaError_t exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* sums, int32_t* total,
                           cudaStream_t st) {
    if (n <= 0) {
        if (total) return cudaMemsetAsync(total, 0, sizeof(int32_t), st);
        return cudaSuccess;
    }
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tile<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, sums, n);
    k_scan_sums<<<1, 1024, 0, st>>>(sums, tiles, total);
    k_scan_add<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(out, sums, n, nullptr);
    count_launch(3);
    return cudaGetLastError();
}

// ---- filtering (src/GNN.py:206-223) --------------------------------------------------------
// keep[e] for e < E0: not masked (and not a self loop when self_loops);  for E0 <= e < E0+K:
// the appended corner loops (dropped again by remove_self_loops when self_loops).
__global__ void k_keep_flags(const int64_t* __restrict__ ei, int64_t E0, const uint8_t* __restrict__ m0,
                             const uint8_t* __restrict__ m1, const uint8_t* __restrict__ m2, int64_t K,
                             int self_loops, int32_t* __restrict__ keep) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E0 + K) return;
    int32_t k;
    if (e < E0) {
        k = 1;
        if (m0 && m0[e]) k = 0;
        if (m1 && m1[e]) k = 0;
        if (m2 && m2[e]) k = 0;
        if (self_loops && ei[e] == ei[E0 + e]) k = 0;
    } else {
        k = self_loops ? 0 : 1;
    }
    keep[e] = k;
}

__global__ void k_compact(const int64_t* __restrict__ ei, int64_t E0, const int64_t* __restrict__ extra,
                          int64_t K, const int32_t* __restrict__ keep, const int32_t* __restrict__ pos,
                          int64_t Emax, int64_t* __restrict__ filt, int32_t* __restrict__ src32,
                          int32_t* __restrict__ dst32) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E0 + K || !keep[e]) return;
    int64_t s, d;
    if (e < E0) {
        s = ei[e];
        d = ei[E0 + e];
    } else {
        s = d = extra[e - E0];
    }
    const int32_t p = pos[e];
    filt[p] = s;
    filt[Emax + p] = d;
    src32[p] = (int32_t)s;
    dst32[p] = (int32_t)d;
}

// add_self_loops: arange(N) appended after the kept edges; also finalises E.
__global__ void k_append_loops(const int32_t* __restrict__ kept_total, int64_t N, int self_loops,
                               int64_t Emax, int64_t* __restrict__ filt, int32_t* __restrict__ src32,
                               int32_t* __restrict__ dst32, int32_t* __restrict__ info) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t Ef = *kept_total;
    if (i == 0) info[GAD_INFO_E] = Ef + (self_loops ? (int32_t)N : 0);
    if (!self_loops || i >= N) return;
    filt[Ef + i] = i;
    filt[Emax + Ef + i] = i;
    src32[Ef + i] = (int32_t)i;
    dst32[Ef + i] = (int32_t)i;
}

// ---- CSR / CSC ------------------------------------------------------------------------------
__global__ void k_degrees(const int32_t* __restrict__ src32, const int32_t* __restrict__ dst32,
                          const int32_t* __restrict__ info, int64_t N, int32_t* __restrict__ deg_in,
                          int32_t* __restrict__ deg_out, int32_t* __restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= info[GAD_INFO_E]) return;
    const int32_t s = src32[e], d = dst32[e];
    if (s < 0 || s >= N || d < 0 || d >= N) {
        atomicAdd(bad, 1);
        return;
    }
    atomicAdd(&deg_in[d], 1);
    atomicAdd(&deg_out[s], 1);
}

__global__ void k_fill_slots(const int32_t* __restrict__ src32, const int32_t* __restrict__ dst32,
                             const int32_t* __restrict__ info, const int32_t* __restrict__ rowptr,
                             const int32_t* __restrict__ t_rowptr, int64_t N, int32_t* __restrict__ cur_in,
                             int32_t* __restrict__ cur_out, int32_t* __restrict__ eid,
                             int32_t* __restrict__ t_eid) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= info[GAD_INFO_E]) return;
    const int32_t s = src32[e], d = dst32[e];
    if (s < 0 || s >= N || d < 0 || d >= N) return;   // reported by k_degrees; the host raises
    eid[rowptr[d] + atomicAdd(&cur_in[d], 1)] = (int32_t)e;
    t_eid[t_rowptr[s] + atomicAdd(&cur_out[s], 1)] = (int32_t)e;
}

// One thread per row: order the row's edge ids ascending (== stable sort on the row key), then
// emit the neighbour index.  Rows of a mesh graph hold <= 7 entries; insertion sort is exact and
// cheap.  `other` is src32 for CSR rows (col) and dst32 for CSC rows (t_dst).
__global__ void k_order_rows(const int32_t* __restrict__ ptr, int64_t N, int32_t* __restrict__ ids,
                             const int32_t* __restrict__ other, int32_t* __restrict__ nbr,
                             int32_t* __restrict__ ptr_last, const int32_t* __restrict__ info,
                             int32_t* __restrict__ max_deg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ptr_last[0] = info[GAD_INFO_E];  // rowptr[N] = E
    if (i >= N) return;
    const int32_t b = ptr[i];
    const int32_t e = (i + 1 < N) ? ptr[i + 1] : info[GAD_INFO_E];
    for (int32_t a = b + 1; a < e; ++a) {
        const int32_t key = ids[a];
        int32_t c = a - 1;
        while (c >= b && ids[c] > key) {
            ids[c + 1] = ids[c];
            --c;
        }
        ids[c + 1] = key;
    }
    for (int32_t a = b; a < e; ++a) nbr[a] = other[ids[a]];
    if (e - b > 0) atomicMax(max_deg, e - b);
}

__global__ void k_invert(const int32_t* __restrict__ eid, const int32_t* __restrict__ info,
                         int32_t* __restrict__ slot_of_edge) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= info[GAD_INFO_E]) return;
    slot_of_edge[eid[s]] = (int32_t)s;
}

__global__ void k_t_slot(const int32_t* __restrict__ t_eid, const int32_t* __restrict__ slot_of_edge,
                         const int32_t* __restrict__ info, int32_t* __restrict__ t_slot) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= info[GAD_INFO_E]) return;
    t_slot[s] = slot_of_edge[t_eid[s]];
}

__global__ void k_check_tiles(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const int32_t* __restrict__ tile_ptr, int T, int32_t* __restrict__ info) {
    const int t = blockIdx.x;
    if (t >= T) return;
    const int32_t n0 = tile_ptr[t], n1 = tile_ptr[t + 1];
    const int32_t e0 = rowptr[n0], e1 = rowptr[n1];
    int32_t bad = 0;
    for (int32_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int32_t j = col[e];
        bad += (j < n0 || j >= n1);
    }
    if (bad) atomicAdd(&info[GAD_INFO_CROSS_TILE], bad);
}

// Per-row ascending copy of a CSR/CSC index array.  The deformer kernels may sum a row in any
// order (the parity bar is 1e-5, not bit-exactness of the sums); ascending neighbour ids make the
// q-th gather of consecutive rows hit consecutive addresses on structured meshes, which removes
// most shared-memory bank conflicts.  The canonical arrays (stable edge order) stay untouched.
__global__ void k_sort_rows(const int32_t* __restrict__ ptr, const int32_t* __restrict__ in, int64_t N,
                            int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int32_t b = ptr[i], e = ptr[i + 1];
    for (int32_t a = b; a < e; ++a) {
        const int32_t key = in[a];
        int32_t c = a - 1;
        while (c >= b && out[c] > key) {
            out[c + 1] = out[c];
            --c;
        }
        out[c + 1] = key;
    }
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

struct GraphWs {
    int32_t *keep, *pos, *sums, *kept_total, *src32, *dst32, *deg_in, *deg_out, *cur_in, *cur_out,
        *t_eid, *slot_of_edge, *bad;
};

size_t carve_graph_ws(Carver& c, int64_t E0, int64_t K, int64_t N, int self_loops, GraphWs* w) {
    const int64_t Etot = E0 + K;
    const int64_t Emax = Etot + (self_loops ? N : 0);
    const int64_t scan_n = (Etot > N + 1 ? Etot : N + 1);
    GraphWs tmp;
    tmp.keep = c.take<int32_t>(Etot + 1);
    tmp.pos = c.take<int32_t>(Etot + 1);
    tmp.sums = c.take<int32_t>((scan_n + SCAN_TILE - 1) / SCAN_TILE + 1);
    tmp.kept_total = c.take<int32_t>(4);
    tmp.src32 = c.take<int32_t>(Emax + 1);
    tmp.dst32 = c.take<int32_t>(Emax + 1);
    tmp.deg_in = c.take<int32_t>(N + 1);
    tmp.deg_out = c.take<int32_t>(N + 1);
    tmp.cur_in = c.take<int32_t>(N + 1);
    tmp.cur_out = c.take<int32_t>(N + 1);
    tmp.t_eid = c.take<int32_t>(Emax + 1);
    tmp.slot_of_edge = c.take<int32_t>(Emax + 1);
    tmp.bad = c.take<int32_t>(4);
    if (w) *w = tmp;
    return c.off;
}

}  // namespace
}  // namespace gad

using namespace gad;

extern "C" size_t gad_graph_workspace_bytes(int64_t E0, int64_t K, int64_t N, int self_loops) {
    Carver c(nullptr, ~size_t(0));
    return carve_graph_ws(c, E0, K, N, self_loops, nullptr) + 256;
}

extern "C" int gad_graph_build(const int64_t* edge_index, int64_t E0, const uint8_t* mask_to_boundary,
                               const uint8_t* mask_to_corner, const uint8_t* mask_diff_boundary,
                               const int64_t* extra_loops, int64_t K, int self_loops, int64_t N,
                               int64_t* filt_edge_index, int32_t* rowptr, int32_t* col, int32_t* eid,
                               int32_t* t_rowptr, int32_t* t_dst, int32_t* t_slot, int32_t* info,
                               void* workspace, size_t workspace_bytes, void* stream) {
    GAD_CHECK_ARG(E0 >= 0 && K >= 0 && N > 0, "gad_graph_build: bad sizes E0=%lld K=%lld N=%lld",
                  (long long)E0, (long long)K, (long long)N);
    const int64_t Etot = E0 + K;
    const int64_t Emax = Etot + (self_loops ? N : 0);
    GAD_CHECK_ARG(Emax < (int64_t)2147483000 && N < (int64_t)2147483000,
                  "gad_graph_build: graph exceeds int32 indexing (Emax=%lld)", (long long)Emax);
    GAD_CHECK_ARG(E0 == 0 || edge_index, "gad_graph_build: edge_index is null");
    GAD_CHECK_ARG(K == 0 || extra_loops, "gad_graph_build: extra_loops is null");
    GAD_CHECK_ARG(filt_edge_index && rowptr && col && eid && t_rowptr && t_dst && t_slot && info && workspace,
                  "gad_graph_build: null output");
    Carver c(workspace, workspace_bytes);
    GraphWs w;
    carve_graph_ws(c, E0, K, N, self_loops, &w);
    GAD_CHECK_ARG(c.ok(), "gad_graph_build: workspace too small (%zu < %zu)", workspace_bytes, c.off);
    cudaStream_t st = as_stream(stream);
    const int TB = 256;

    GAD_CUDA(cudaMemsetAsync(info, 0, GAD_INFO_WORDS * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(w.deg_in, 0, (N + 1) * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(w.deg_out, 0, (N + 1) * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(w.cur_in, 0, (N + 1) * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(w.cur_out, 0, (N + 1) * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(w.bad, 0, 4 * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(w.kept_total, 0, 4 * sizeof(int32_t), st));

    if (Etot > 0) {
        k_keep_flags<<<blocks_for(Etot, TB), TB, 0, st>>>(edge_index, E0, mask_to_boundary, mask_to_corner,
                                                          mask_diff_boundary, K, self_loops, w.keep);
        GAD_LAUNCH_CHECK();
        GAD_CUDA(exclusive_scan(w.keep, w.pos, Etot, w.sums, w.kept_total, st));
        k_compact<<<blocks_for(Etot, TB), TB, 0, st>>>(edge_index, E0, extra_loops, K, w.keep, w.pos, Emax,
                                                       filt_edge_index, w.src32, w.dst32);
        GAD_LAUNCH_CHECK();
    }
    k_append_loops<<<blocks_for(self_loops ? N : 1, TB), TB, 0, st>>>(w.kept_total, N, self_loops, Emax,
                                                                      filt_edge_index, w.src32, w.dst32, info);
    GAD_LAUNCH_CHECK();
    if (Emax > 0) {
        k_degrees<<<blocks_for(Emax, TB), TB, 0, st>>>(w.src32, w.dst32, info, N, w.deg_in, w.deg_out, w.bad);
        GAD_LAUNCH_CHECK();
    }
    GAD_CUDA(exclusive_scan(w.deg_in, rowptr, N, w.sums, nullptr, st));
    GAD_CUDA(exclusive_scan(w.deg_out, t_rowptr, N, w.sums, nullptr, st));
    if (Emax > 0) {
        k_fill_slots<<<blocks_for(Emax, TB), TB, 0, st>>>(w.src32, w.dst32, info, rowptr, t_rowptr, N, w.cur_in,
                                                          w.cur_out, eid, w.t_eid);
        GAD_LAUNCH_CHECK();
    }
    k_order_rows<<<blocks_for(N, TB), TB, 0, st>>>(rowptr, N, eid, w.src32, col, rowptr + N, info,
                                                   info + GAD_INFO_MAX_IN_DEG);
    GAD_LAUNCH_CHECK();
    k_order_rows<<<blocks_for(N, TB), TB, 0, st>>>(t_rowptr, N, w.t_eid, w.dst32, t_dst, t_rowptr + N, info,
                                                   info + GAD_INFO_MAX_OUT_DEG);
    GAD_LAUNCH_CHECK();
    if (Emax > 0) {
        k_invert<<<blocks_for(Emax, TB), TB, 0, st>>>(eid, info, w.slot_of_edge);
        GAD_LAUNCH_CHECK();
        k_t_slot<<<blocks_for(Emax, TB), TB, 0, st>>>(w.t_eid, w.slot_of_edge, info, t_slot);
        GAD_LAUNCH_CHECK();
    }
    // out-of-range node ids are reported through info[7] (host raises)
    GAD_CUDA(cudaMemcpyAsync(info + 7, w.bad, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return GAD_OK;
}

extern "C" int gad_graph_sort_rows(const int32_t* ptr, const int32_t* idx, int64_t N, int32_t* idx_sorted, void* stream) {
    GAD_CHECK_ARG(ptr && idx && idx_sorted && N > 0 && idx != idx_sorted, "gad_graph_sort_rows: bad arguments");
    k_sort_rows<<<blocks_for(N, 256), 256, 0, as_stream(stream)>>>(ptr, idx, N, idx_sorted);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

extern "C" int gad_graph_check_tiles(const int32_t* rowptr, const int32_t* col, int64_t N,
                                     const int32_t* tile_ptr, int T, int32_t* info, void* stream) {
    GAD_CHECK_ARG(rowptr && col && tile_ptr && info && T > 0 && N > 0, "gad_graph_check_tiles: bad arguments");
    k_check_tiles<<<T, 128, 0, as_stream(stream)>>>(rowptr, col, tile_ptr, T, info);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// ---- content fingerprint of a topology tensor ---------------------------------------------------
// The reference's training loop hands the model a FRESH Batch object every iteration
// (src/run_GNN.py:97-105 over a DataLoader), so a cache keyed on tensor identity would rebuild the
// graph each step although the topology is the same.  The host keys its graph cache on this
// 128-bit fingerprint instead: out[0..1] += sum_i mix(word_i, i, seed), a position-sensitive sum of
// two independent 64-bit mixes (integer adds commute, so the result does not depend on the order in
// which the atomics land).  One coalesced pass at HBM speed; 16 bytes travel back to the host.
namespace gad {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_fingerprint(const unsigned char* __restrict__ data, size_t bytes,
                                                    unsigned long long seed, unsigned long long* __restrict__ out) {
    const size_t n8 = bytes >> 3;
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(data);
    unsigned long long h1 = 0, h2 = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long v = w[i];
        h1 += mix64(v + seed + i * 0x9e3779b97f4a7c15ull);
        h2 += mix64((v ^ 0xd6e8feb86659fd93ull) + (seed << 1) + i * 0xc2b2ae3d27d4eb4full);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {          // tail bytes (< 8) and the length
        unsigned long long v = 0;
        for (size_t b = n8 << 3; b < bytes; ++b) v = (v << 8) | data[b];
        h1 += mix64(v + seed + n8 * 0x9e3779b97f4a7c15ull) + mix64(bytes ^ seed);
        h2 += mix64((v ^ 0xd6e8feb86659fd93ull) + (seed << 1) + n8 * 0xc2b2ae3d27d4eb4full);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        h1 += __shfl_xor_sync(0xffffffffu, h1, d);
        h2 += __shfl_xor_sync(0xffffffffu, h2, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, h1);
        atomicAdd(out + 1, h2);
    }
}

}  // namespace
}  // namespace gad

extern "C" int gad_fingerprint(const void* data, size_t bytes, uint64_t seed, uint64_t* out, void* stream) {
    GAD_CHECK_ARG(out && (data || bytes == 0), "gad_fingerprint: bad arguments");
    GAD_CHECK_ARG((reinterpret_cast<uintptr_t>(data) & 7) == 0, "gad_fingerprint: data must be 8-byte aligned");
    const size_t n8 = bytes >> 3;
    int grid = (int)((n8 + 255) / 256);
    const int cap = 4 * gad::sm_count();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    gad::k_fingerprint<<<grid, 256, 0, gad::as_stream(stream)>>>(reinterpret_cast<const unsigned char*>(data), bytes,
                                                               (unsigned long long)seed,
                                                               reinterpret_cast<unsigned long long*>(out));
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// ---- edge masks of firedrake_mesh_to_PyG (src/data.py:465-494) on the device -----------------------
// side_bits[v]: bit k set when node v lies on boundary marker k + 1 (the per-marker DirichletBC node
// lists of :451-455, as a bit set).  Boundary = any bit; corner = more than one bit (:457-462).
//   to_boundary  = dst on the boundary and src not                               (:465)
//   to_corner    = dst is a corner                                                (:468)
//   diff_boundary= both on the boundary, different marker lists, neither a corner (:479-494)
// One thread per edge, coalesced; replaces three Python loops over the edge list.
namespace gad {
namespace {

__global__ void __launch_bounds__(256) k_edge_masks(const int64_t* __restrict__ ei, int64_t E, const uint8_t* __restrict__ side_bits,
                                                   int64_t N, uint8_t* __restrict__ to_boundary, uint8_t* __restrict__ to_corner,
                                                   uint8_t* __restrict__ diff_boundary, int32_t* __restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N) {
        atomicAdd(bad, 1);
        to_boundary[e] = to_corner[e] = diff_boundary[e] = 0;
        return;
    }
    const unsigned bs = side_bits[s], bd = side_bits[d];
    const bool s_corner = __popc(bs) > 1, d_corner = __popc(bd) > 1;
    to_boundary[e] = (bd != 0 && bs == 0) ? 1 : 0;
    to_corner[e] = d_corner ? 1 : 0;
    diff_boundary[e] = (bs != 0 && bd != 0 && bs != bd && !s_corner && !d_corner) ? 1 : 0;
}

}  // namespace
}  // namespace gad

extern "C" int gad_edge_masks(const int64_t* edge_index, int64_t E, const uint8_t* side_bits, int64_t N,
                              uint8_t* to_boundary, uint8_t* to_corner, uint8_t* diff_boundary, int32_t* info,
                              void* stream) {
    GAD_CHECK_ARG(edge_index && side_bits && to_boundary && to_corner && diff_boundary && info && E >= 0 && N > 0,
                  "gad_edge_masks: bad arguments");
    if (E == 0) return GAD_OK;
    gad::k_edge_masks<<<(unsigned)((E + 255) / 256), 256, 0, gad::as_stream(stream)>>>(edge_index, E, side_bits, N, to_boundary,
                                                                                     to_corner, diff_boundary, info + 7);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}
