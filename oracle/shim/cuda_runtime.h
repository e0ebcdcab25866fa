// TEST INFRASTRUCTURE ONLY -- host emulation of the tiny slice of the CUDA runtime
// that the reference (ams3878/cpp_cuda_raytracer_dev, TEST_Dungeonrun/*.cu) uses, so
// that its own kernels can be executed on host cores as the parity oracle
// (SURVEY.md section 8(c)).  "Device" memory is plain host memory, a kernel launch is a
// loop nest over (block, thread) -- see EMU_LAUNCH.  Nothing in the product links this.
#pragma once
#ifndef RTB_EMU_CUDA_RUNTIME_H
#define RTB_EMU_CUDA_RUNTIME_H
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <typeinfo>

#define __global__
#define __device__
#define __host__

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

struct emu_dim3 { unsigned x, y, z; };
extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaDeviceReset() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

// One "launch": every block runs its threads serially; blocks are spread over host
// cores (the reference kernels only ever write per-thread-private slots).
#define EMU_LAUNCH(kern, G, B, ...)                                                   \
    do {                                                                              \
        const long emu_g_ = (long)(G);                                                \
        const unsigned emu_b_ = (unsigned)(B);                                        \
        _Pragma("omp parallel for schedule(dynamic, 2)")                              \
        for (long emu_bi_ = 0; emu_bi_ < emu_g_; ++emu_bi_) {                         \
            gridDim = emu_dim3{(unsigned)emu_g_, 1u, 1u};                             \
            blockDim = emu_dim3{emu_b_, 1u, 1u};                                      \
            blockIdx = emu_dim3{(unsigned)emu_bi_, 0u, 0u};                           \
            for (unsigned emu_t_ = 0; emu_t_ < emu_b_; ++emu_t_) {                    \
                threadIdx = emu_dim3{emu_t_, 0u, 0u};                                 \
                kern(__VA_ARGS__);                                                    \
            }                                                                         \
        }                                                                             \
    } while (0)
#endif
