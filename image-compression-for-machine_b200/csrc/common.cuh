// Shared helpers for the icm_b200 CUDA library (error plumbing, launch accounting, strided views).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#include "icm_b200.h"

namespace icm {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define ICM_CHECK_ARG(cond, ...)                    \
    do {                                            \
        if (!(cond)) {                              \
            ::icm::set_error(__VA_ARGS__);          \
            return ICM_ERR_INVALID_ARG;             \
        }                                           \
    } while (0)

#define ICM_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::icm::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return ICM_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define ICM_LAUNCH_CHECK()                                                                      \
    do {                                                                                        \
        ::icm::count_launch();                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) {                                                               \
            ::icm::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return ICM_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// device-side strided [B, C, P] view
struct View {
    char *ptr;
    long long sb, sc, sp;
};
static inline View as_view(const icm_view &v) { return View{(char *)v.ptr, v.sb, v.sc, v.sp}; }

int *tile_counter(cudaStream_t st); // csrc/conv.cu: the dynamic tile scheduler's {next, finished} pair of this stream
int sm_count();   // of the current device (cached per device)
int current_device_ordinal();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device (and per-kernel) setting, not a per-thread one: each
// call site keeps the largest size it has configured on every device ordinal.
constexpr int kMaxDeviceOrdinals = 64;
struct PerDeviceSmem {
    size_t configured[kMaxDeviceOrdinals] = {};
    bool needs(size_t bytes) const { return bytes > configured[current_device_ordinal()]; }
    void done(size_t bytes) { configured[current_device_ordinal()] = bytes; }
};
int persistent_grid_limit();          // SMs the persistent GEMM kernels may occupy (icm_set_conv_sm_limit)
void set_persistent_grid_limit(int n);

}  // namespace icm
