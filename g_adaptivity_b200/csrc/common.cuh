// Shared host/device helpers for the gadapt sm_100a library.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gadapt.h"

namespace gad {

void set_error(const char* fmt, ...);
void count_launch(int n);   // host-side tally of kernel launches issued by this library

#define GAD_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            gad::set_error(__VA_ARGS__);    \
            return GAD_ERR_ARG;             \
        }                                   \
    } while (0)

#define GAD_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t err__ = (call);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            gad::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,         \
                           cudaGetErrorString(err__));                                   \
            return GAD_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define GAD_LAUNCH_CHECK()                                                               \
    do {                                                                                 \
        gad::count_launch(1);                                                            \
        cudaError_t err__ = cudaGetLastError();                                          \
        if (err__ != cudaSuccess) {                                                      \
            gad::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,     \
                           cudaGetErrorString(err__));                                   \
            return GAD_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// B200: 148 SMs.  Queried once per process (immutable after first use).
int sm_count();
int smem_optin_bytes();

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// bump allocator over a caller-provided workspace
struct Carver {
    char* base;
    size_t off, cap;
    Carver(void* p, size_t bytes) : base(reinterpret_cast<char*>(p)), off(0), cap(bytes) {}
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

}  // namespace gad
