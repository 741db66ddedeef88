// Peer memory for the data-parallel gradient exchange (the only collective of the path is
// the SUM all-reduce of the flat gradient, 144 + L floats per step).
//
// One process per GPU on one NVLink / NVSwitch node: every rank allocates a small receive buffer,
// publishes its CUDA IPC handle (the Python side all-gathers the handles over torch.distributed)
// and maps the buffers of its peers.  The training kernel's tail then does the all-reduce itself:
// it stores its gradient, tagged with a sequence number, straight into every peer's buffer over
// NVLink and sums what the peers stored into its own (ell_kernels.cuh: tail_finish) -- no NCCL
// launch, no extra kernel between the chain rule and Adam.
#include "common.cuh"

using namespace gad;

extern "C" size_t gad_peer_exchange_bytes(int world, int64_t n_params) {
    // [2 sequence parities][world sources][n_params] x { float value, uint32 sequence }
    if (world < 1 || n_params < 1) return 0;
    return (size_t)2 * (size_t)world * (size_t)n_params * 8;
}

extern "C" int gad_peer_alloc(size_t bytes, void** dev_ptr, void* handle_out) {
    GAD_CHECK_ARG(bytes > 0 && dev_ptr && handle_out, "gad_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == GAD_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    GAD_CUDA(cudaMalloc(&p, bytes));
    GAD_CUDA(cudaMemset(p, 0, bytes));
    GAD_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, p);
    if (err != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(err));
        return GAD_ERR_CUDA;
    }
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return GAD_OK;
}

extern "C" int gad_peer_open(const void* handle, void** dev_ptr) {
    GAD_CHECK_ARG(handle && dev_ptr, "gad_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    GAD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return GAD_OK;
}

extern "C" int gad_peer_close(void* dev_ptr) {
    if (dev_ptr) GAD_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return GAD_OK;
}

extern "C" int gad_peer_free(void* dev_ptr) {
    if (dev_ptr) GAD_CUDA(cudaFree(dev_ptr));
    return GAD_OK;
}
