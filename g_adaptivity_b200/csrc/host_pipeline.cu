// Host side of the end-to-end training loop (the reference's `for data in loader: ... optimizer.step()`,
// src/run_GNN.py:95-131, with the batches in pinned host memory): a double-buffered H2D / compute / D2H
// pipeline driven from C, so that the per-step CPU cost is a handful of CUDA API calls instead of a
// Python loop (which, with one process per GPU on a shared host, becomes the limit long before PCIe).
//
//   copy stream   : step k+1's inputs host -> device (ONE cudaMemcpyAsync: the batch is packed in the slot's
//                   input layout, DeformerTrainer.pack_host) as soon as step k+1-R has released the slot
//   compute stream: wait for the inputs, replay the slot's captured step (cudaGraphLaunch of the one-launch
//                   training kernel), copy the 4-byte loss device -> host, release the slot
//
// Slot k % n_slots receives host batch k % n_host.  The call returns when the last step has finished
// (cudaStreamSynchronize of the compute stream).  No hidden allocation besides a grow-only pool of 2 * n_slots + 1 events.
#include <vector>

#include "common.cuh"

using namespace gad;

extern "C" int gad_pipeline_run(const gad_pipeline_slot* slots, int n_slots, const void* const* host_batches,
                                int n_host, int64_t steps, float* losses_host, void* compute_stream,
                                void* copy_stream) {
    return gad_pipeline_run_relay(slots, n_slots, host_batches, n_host, steps, losses_host, compute_stream, copy_stream,
                                  nullptr);
}

// The same loop with a second path into the GPU.  On boxes whose GPUs do not have equal paths to host memory (measured:
// 20.7 against 36.2 GB/s per GPU with eight GPUs copying at once) a synchronous data-parallel step waits for the
// slowest copy.  `relay` sends the tail of every batch (bytes beyond relay->direct_bytes) through ANOTHER GPU of the
// same process' view of the box: host -> staging buffer on relay->device over that GPU's PCIe path (on relay->stream,
// a stream of that device), then staging -> this GPU's slot over NVLink (cudaMemcpyPeerAsync), and the compute stream
// waits for both parts.  Everything stays inside this process -- ordinary events order streams of different devices --
// so nothing has to be agreed with the rank that owns the other GPU; it only sees its link and copy engines shared.
extern "C" int gad_pipeline_run_relay(const gad_pipeline_slot* slots, int n_slots, const void* const* host_batches,
                                      int n_host, int64_t steps, float* losses_host, void* compute_stream,
                                      void* copy_stream, const gad_pipeline_relay* relay) {
    GAD_CHECK_ARG(slots && host_batches && losses_host && n_slots >= 2 && n_host >= 1 && steps >= 0,
                  "gad_pipeline_run: bad arguments (needs >= 2 slots for double buffering)");
    for (int s = 0; s < n_slots; ++s)
        GAD_CHECK_ARG(slots[s].dev_inputs && slots[s].graph_exec && slots[s].loss_dev && slots[s].bytes > 0,
                      "gad_pipeline_run: slot %d is incomplete", s);
    cudaStream_t cs = as_stream(copy_stream), ms = as_stream(compute_stream);
    GAD_CHECK_ARG(cs != ms, "gad_pipeline_run: the copy stream must differ from the compute stream");
    // events come from a grow-only per-thread pool (creating 2 * n_slots + 1 events costs more than a short run)
    static thread_local std::vector<cudaEvent_t> pool;
    static thread_local int pool_device = -1;
    int device = -1;
    GAD_CUDA(cudaGetDevice(&device));
    if (device != pool_device) {
        pool.clear();      // events belong to the device they were created on
        pool_device = device;
    }
    while ((int)pool.size() < 2 * n_slots + 1) {
        cudaEvent_t e;
        GAD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        pool.push_back(e);
    }
    cudaEvent_t* in_ready = pool.data();
    cudaEvent_t* slot_free = pool.data() + n_slots;
    cudaEvent_t start = pool[2 * n_slots];
    int rc = GAD_OK;
    // relay path: events of the other device (they belong to the device that was current at creation)
    static thread_local std::vector<cudaEvent_t> rpool;
    static thread_local int rpool_device = -1;
    cudaStream_t rs = nullptr;
    cudaEvent_t* relay_ready = nullptr;
    if (relay) {
        GAD_CHECK_ARG(relay->device >= 0 && relay->device != device && relay->stream && relay->staging &&
                          relay->staging_stride > 0,
                      "gad_pipeline_run_relay: incomplete relay descriptor");
        for (int s = 0; s < n_slots; ++s)
            GAD_CHECK_ARG(relay->direct_bytes < slots[s].bytes && slots[s].bytes - relay->direct_bytes <= relay->staging_stride,
                          "gad_pipeline_run_relay: slot %d: direct_bytes / staging_stride do not fit %zu bytes", s,
                          slots[s].bytes);
        rs = as_stream(relay->stream);
        GAD_CUDA(cudaSetDevice(relay->device));
        if (rpool_device != relay->device) {
            rpool.clear();
            rpool_device = relay->device;
        }
        cudaError_t e = cudaSuccess;
        while ((int)rpool.size() < n_slots && e == cudaSuccess) {
            cudaEvent_t ev;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e == cudaSuccess) rpool.push_back(ev);
        }
        cudaSetDevice(device);
        GAD_CUDA(e);
        relay_ready = rpool.data();
    }
    auto fail = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == GAD_OK) {
            set_error("gad_pipeline_run: %s failed: %s", what, cudaGetErrorString(e));
            rc = GAD_ERR_CUDA;
        }
        return e != cudaSuccess;
    };
    // the first upload must not overtake work already queued on the compute stream (it may still read the slot)
    fail(cudaEventRecord(start, ms), "cudaEventRecord");
    fail(cudaStreamWaitEvent(cs, start, 0), "cudaStreamWaitEvent");
    if (relay) fail(cudaStreamWaitEvent(rs, start, 0), "cudaStreamWaitEvent (relay)");
    auto upload = [&](int64_t k) {
        const int sid = (int)(k % n_slots);
        const size_t direct = relay ? relay->direct_bytes : slots[sid].bytes;
        if (k >= n_slots) fail(cudaStreamWaitEvent(cs, slot_free[sid], 0), "cudaStreamWaitEvent");
        fail(cudaMemcpyAsync(slots[sid].dev_inputs, host_batches[k % n_host], direct, cudaMemcpyHostToDevice, cs),
             "cudaMemcpyAsync (inputs)");
        fail(cudaEventRecord(in_ready[sid], cs), "cudaEventRecord");
        if (relay) {
            const size_t rest = slots[sid].bytes - direct;
            char* stage = static_cast<char*>(relay->staging) + (size_t)sid * relay->staging_stride;
            const char* src = static_cast<const char*>(host_batches[k % n_host]) + direct;
            if (k >= n_slots) fail(cudaStreamWaitEvent(rs, slot_free[sid], 0), "cudaStreamWaitEvent (relay)");
            fail(cudaSetDevice(relay->device), "cudaSetDevice (relay)");
            fail(cudaMemcpyAsync(stage, src, rest, cudaMemcpyHostToDevice, rs), "cudaMemcpyAsync (relay, host -> staging)");
            fail(cudaMemcpyPeerAsync(static_cast<char*>(slots[sid].dev_inputs) + direct, device, stage, relay->device, rest, rs),
                 "cudaMemcpyPeerAsync (relay, staging -> slot)");
            fail(cudaEventRecord(relay_ready[sid], rs), "cudaEventRecord (relay)");
            fail(cudaSetDevice(device), "cudaSetDevice");
        }
    };
    if (steps > 0) upload(0);
    for (int64_t k = 0; k < steps && rc == GAD_OK; ++k) {
        const int sid = (int)(k % n_slots);
        if (k + 1 < steps) upload(k + 1);
        fail(cudaStreamWaitEvent(ms, in_ready[sid], 0), "cudaStreamWaitEvent");
        if (relay) fail(cudaStreamWaitEvent(ms, relay_ready[sid], 0), "cudaStreamWaitEvent (relay)");
        fail(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(slots[sid].graph_exec), ms), "cudaGraphLaunch");
        fail(cudaMemcpyAsync(losses_host + k, slots[sid].loss_dev, sizeof(float), cudaMemcpyDeviceToHost, ms),
             "cudaMemcpyAsync (loss)");
        fail(cudaEventRecord(slot_free[sid], ms), "cudaEventRecord");
    }
    fail(cudaStreamSynchronize(ms), "cudaStreamSynchronize");
    fail(cudaStreamSynchronize(cs), "cudaStreamSynchronize");
    if (relay) fail(cudaStreamSynchronize(rs), "cudaStreamSynchronize (relay)");
    cudaSetDevice(device);
    return rc;
}

// Peer access between two devices of this process, both directions (cudaMemcpyPeerAsync then goes over NVLink
// instead of through host memory).  Already-enabled is not an error.
extern "C" int gad_enable_peer_access(int dev_a, int dev_b) {
    int cur = -1;
    GAD_CUDA(cudaGetDevice(&cur));
    int rc = GAD_OK;
    for (int dir = 0; dir < 2 && rc == GAD_OK; ++dir) {
        const int from = dir ? dev_b : dev_a, to = dir ? dev_a : dev_b;
        int can = 0;
        cudaError_t e = cudaDeviceCanAccessPeer(&can, from, to);
        if (e == cudaSuccess && !can) {
            set_error("gad_enable_peer_access: device %d cannot access device %d", from, to);
            rc = GAD_ERR_UNSUPPORTED;
            break;
        }
        if (e == cudaSuccess) e = cudaSetDevice(from);
        if (e == cudaSuccess) {
            e = cudaDeviceEnablePeerAccess(to, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                (void)cudaGetLastError();
                e = cudaSuccess;
            }
        }
        if (e != cudaSuccess) {
            set_error("gad_enable_peer_access(%d, %d) failed: %s", from, to, cudaGetErrorString(e));
            rc = GAD_ERR_CUDA;
        }
    }
    cudaSetDevice(cur);
    return rc;
}

// Pinned host staging buffers for the packed batches.  write_combined != 0: cudaHostAllocWriteCombined -- the host
// writes a batch once and never reads it back, and device reads of write-combined memory are not snooped through the
// CPU caches, which matters when eight processes pull their batches through one host memory system.
extern "C" int gad_host_alloc(size_t bytes, int write_combined, void** host_ptr) {
    GAD_CHECK_ARG(bytes > 0 && host_ptr, "gad_host_alloc: bad arguments");
    void* p = nullptr;
    GAD_CUDA(cudaHostAlloc(&p, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    *host_ptr = p;
    return GAD_OK;
}

extern "C" int gad_host_free(void* host_ptr) {
    if (host_ptr) GAD_CUDA(cudaFreeHost(host_ptr));
    return GAD_OK;
}
