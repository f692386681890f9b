// Shared helpers for libsclmd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sclmd_b200.h"

namespace sclmd {

void set_error(const char *fmt, ...);

#define SCLMD_CUDA(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            sclmd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                             __LINE__);                                                          \
            return SCLMD_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

#define SCLMD_REQUIRE(cond, ...)                                                                 \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            sclmd::set_error(__VA_ARGS__);                                                       \
            return SCLMD_ERR_ARG;                                                                \
        }                                                                                        \
    } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// device buffer with RAII, zero-initialised
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc_raw(size_t count) {      // scratch that is fully written before it is read: no zero-fill
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        n = count;
        return cudaSuccess;
    }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        n = count;
        // handles run on non-blocking streams: the zero-fill must have landed before they start
        e = cudaMemset(p, 0, count * sizeof(T));
        if (e != cudaSuccess) return e;
        return cudaDeviceSynchronize();
    }
};

int select_device(int device);
int sm_count(int device);
void release_noise_scratch();      // cached device scratch of the noise generator (noise.cu); part of sclmd_release_workspace()

// ------------------------------------------------------------------ device side
#ifdef __CUDACC__
__device__ __forceinline__ double2 ld_stream2(const double *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum; result valid in thread 0 (and all threads of warp 0). `red` >= 32 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? red[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}
#endif

}  // namespace sclmd
