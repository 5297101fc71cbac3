// Shared host/device helpers for libcorb_b200 (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/corb_b200.h"

namespace corb {

void set_error(const char* fmt, ...);
const char* get_error();

#define CORB_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            corb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return CORB_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

// The dynamic shared memory limit of a kernel is global per (function, device). It is raised ONCE per device to the opt-in
// maximum (minus the kernel's static shared memory), never to a call's exact requirement: two concurrent calls with different
// sizes could otherwise lower the limit between another call's set and its launch.
#define CORB_SMEM_OPT_IN(kernel)                                                                                       \
    do {                                                                                                               \
        static std::atomic<unsigned long long> _done{0};                                                              \
        int _dev = 0;                                                                                                  \
        CORB_CUDA(cudaGetDevice(&_dev));                                                                               \
        const unsigned long long _bit = 1ull << (_dev & 63);                                                           \
        if (!(_done.load(std::memory_order_acquire) & _bit)) {                                                         \
            int _max = 0;                                                                                              \
            cudaFuncAttributes _fa;                                                                                    \
            CORB_CUDA(cudaDeviceGetAttribute(&_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, _dev));                   \
            CORB_CUDA(cudaFuncGetAttributes(&_fa, kernel));                                                            \
            CORB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, _max - (int)_fa.sharedSizeBytes)); \
            _done.fetch_or(_bit, std::memory_order_release);                                                           \
        }                                                                                                              \
    } while (0)

#define CORB_CHECK(cond, code, ...)         \
    do {                                    \
        if (!(cond)) {                      \
            corb::set_error(__VA_ARGS__);   \
            return (code);                  \
        }                                   \
    } while (0)

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
static inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Block-wide exclusive scan of data[0..n) in shared memory, in place; returns the total to every thread.
// `warp_tmp` must hold 33 ints of shared memory. All threads of the block must call it.
__device__ __forceinline__ int block_excl_scan(int* data, int n, int* warp_tmp) {
    const int nt = blockDim.x, t = threadIdx.x;
    const int chunk = (n + nt - 1) / nt;
    const int b = t * chunk, e = min(b + chunk, n);
    int sum = 0;
    for (int i = b; i < e; i++) sum += data[i];
    // scan of per-thread sums
    const int lane = t & 31, wid = t >> 5;
    int v = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    if (lane == 31) warp_tmp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        const int nw = (nt + 31) >> 5;
        int w = lane < nw ? warp_tmp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        warp_tmp[lane] = w;  // inclusive over warps
        if (lane == 31) warp_tmp[32] = w;
    }
    __syncthreads();
    int run = v - sum + (wid ? warp_tmp[wid - 1] : 0);  // exclusive prefix of this thread's chunk
    const int total = warp_tmp[min((nt + 31) >> 5, 32) - 1];
    for (int i = b; i < e; i++) {
        int x = data[i];
        data[i] = run;
        run += x;
    }
    __syncthreads();
    return total;
}


// Vectorised variant for the latency-critical single-CTA kernels: int4 accesses (conflict-free), the per-thread values
// stay in registers, warp totals are combined by every warp itself (__reduce_add_sync), so there are two barriers and
// no serial chunk loops. `data` must be 16-byte aligned with room for n rounded up to a multiple of 4 (pad entries are
// read as zero and overwritten); n <= 16 * blockDim.x, else the generic scan above is used. `warp_tmp`: 33 ints.
__device__ __forceinline__ int block_excl_scan4(int* data, int n, int* warp_tmp) {
    const int nt = blockDim.x, t = threadIdx.x;
    const int n4 = (n + 3) >> 2;
    if (n4 > 4 * nt) return block_excl_scan(data, n, warp_tmp);
    const int per = (n4 + nt - 1) / nt;  // 1..4 consecutive int4 per thread
    const int b = t * per;
    int4 v[4];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[i] = make_int4(0, 0, 0, 0);
        if (i < per && b + i < n4) {
            v[i] = reinterpret_cast<const int4*>(data)[b + i];
            const int e0 = 4 * (b + i);
            if (e0 + 1 >= n) v[i].y = 0;
            if (e0 + 2 >= n) v[i].z = 0;
            if (e0 + 3 >= n) v[i].w = 0;
            sum += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    }
    const int lane = t & 31, wid = t >> 5, nw = (nt + 31) >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tmp[wid] = inc;
    __syncthreads();
    const int wt = lane < nw ? warp_tmp[lane] : 0;
    const int total = __reduce_add_sync(0xffffffffu, wt);
    const int base = __reduce_add_sync(0xffffffffu, lane < wid ? wt : 0);
    int run = base + inc - sum;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < per && b + i < n4) {
            int4 o4;
            o4.x = run; run += v[i].x;
            o4.y = run; run += v[i].y;
            o4.z = run; run += v[i].z;
            o4.w = run; run += v[i].w;
            reinterpret_cast<int4*>(data)[b + i] = o4;
        }
    }
    __syncthreads();
    return total;
}

// Same for 64-bit values (two packed 32-bit counters scanned at once). `warp_tmp` must hold 33 long longs.
__device__ __forceinline__ long long block_excl_scan64(long long* data, int n, long long* warp_tmp) {
    const int nt = blockDim.x, t = threadIdx.x;
    const int chunk = (n + nt - 1) / nt;
    const int b = t * chunk, e = min(b + chunk, n);
    long long sum = 0;
    for (int i = b; i < e; i++) sum += data[i];
    const int lane = t & 31, wid = t >> 5;
    long long v = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    if (lane == 31) warp_tmp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        const int nw = (nt + 31) >> 5;
        long long w = lane < nw ? warp_tmp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        warp_tmp[lane] = w;
    }
    __syncthreads();
    long long run = v - sum + (wid ? warp_tmp[wid - 1] : 0);
    const long long total = warp_tmp[((nt + 31) >> 5) - 1];
    for (int i = b; i < e; i++) {
        long long x = data[i];
        data[i] = run;
        run += x;
    }
    __syncthreads();
    return total;
}

}  // namespace corb
