// common.cuh -- shared host/device plumbing for libpansfem2_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <utility>
#include <algorithm>
#include <cmath>
#include <nvtx3/nvToolsExt.h>
#include "../../include/pansfem2_b200.h"

namespace pf2 {

void set_error(const char* fmt, ...);

#define PF2_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            pf2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return PF2_E_CUDA;                                                                    \
        }                                                                                         \
    } while (0)

#define PF2_CHECK(cond, msg)                                                   \
    do {                                                                       \
        if (!(cond)) {                                                         \
            pf2::set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);   \
            return PF2_E_INVALID;                                              \
        }                                                                      \
    } while (0)

#define PF2_TRY(call)                 \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != PF2_OK) return rc__; \
    } while (0)

#define PF2_LAUNCH_CHECK() PF2_CUDA(cudaGetLastError())

constexpr int kThreads = 256;

// NVTX range around a phase of the path (header-only NVTX 3: a no-op of a few nanoseconds unless a tool -- nsys, ncu --nvtx -- is attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
    void next(const char* name) { nvtxRangePop(); nvtxRangePushA(name); }      // consecutive phases of one function; an early return still pops
};

// element routines (element.cuh, element_generic.cuh) also compile for the host so that tests/cpp/host_elements.cu can check the
// SAME source against the reference fixtures on a machine without a GPU; the library itself only ever calls them from kernels
#define PF2_HD __host__ __device__ __forceinline__

// reduction scratch: per-block partials + a ticket counter; the last block to arrive folds the partials in a fixed
// order, so every reduction is deterministic for a given grid size.
struct ReduceScratch {
    double* partials = nullptr;   // kMaxBlocks * kMaxTerms
    unsigned int* ticket = nullptr;
};
constexpr int kMaxBlocks = 4096;
constexpr int kMaxTerms = 32;

}  // namespace pf2

struct pf2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    long long launches = 0;
    pf2::ReduceScratch red;
    double* scalars = nullptr;        // device scratch for small results (64 doubles)
    double* h_scalars = nullptr;      // pinned mirror
    double* elem_scratch = nullptr;   // per-element legacy call: coordinates in, Ke out
    void* flush_buf = nullptr;        // L2 flush scratch
    size_t flush_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    size_t l2_persist_max = 0;        // cudaDevAttrMaxPersistingL2CacheSize
    size_t l2_window_max = 0;         // cudaDevAttrMaxAccessPolicyWindowSize
    bool l2_persist_enabled = false;  // PF2_L2_PERSIST=1 enables the access-policy window around Krylov solves
    std::vector<std::pair<const void*, int>> occ_cache;
    // one full wave of a persistent kernel: resident CTAs per SM (occupancy API) x SM count
    int wave_grid(const void* kernel, int block, size_t smem = 0) {
        for (auto& kv : occ_cache) if (kv.first == kernel) return kv.second;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        int g = per_sm * sm_count;
        if (g > pf2::kMaxBlocks) g = (pf2::kMaxBlocks / sm_count) * sm_count;
        occ_cache.push_back({ kernel, g });
        return g;
    }
    // grid size for n work items with `per` items per thread, capped to a few waves of the machine
    int grid_for(long long n, int per = 1, int waves = 8) const {
        long long b = (n + (long long)pf2::kThreads * per - 1) / ((long long)pf2::kThreads * per);
        long long cap = (long long)sm_count * waves;
        if (b > cap) b = cap;
        if (b > pf2::kMaxBlocks) b = pf2::kMaxBlocks;
        if (b < 1) b = 1;
        return (int)b;
    }
};

namespace pf2 {

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of NT terms (blockDim.x == kThreads).  Result valid in thread 0.
template <int NT>
__device__ __forceinline__ void block_sum(double (&v)[NT]) {
    __shared__ double sh[NT][kThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < NT; t++) {
        v[t] = warp_sum(v[t]);
        if (lane == 0) sh[t][w] = v[t];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int t = 0; t < NT; t++) {
            double x = (lane < kThreads / 32) ? sh[t][lane] : 0.0;
            v[t] = warp_sum(x);
        }
    }
    __syncthreads();
}

// Grid-wide deterministic sum of NT terms.  Every block contributes its block sums; the LAST block to arrive folds all
// partials (fixed order) and returns true in all of its threads with the totals in v (thread 0 holds them).
// `ticket` must be zero on entry and is reset by the last block.
template <int NT>
__device__ __forceinline__ bool grid_sum_last(double (&v)[NT], double* partials, unsigned int* ticket) {
    __shared__ bool is_last;
    block_sum<NT>(v);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int t = 0; t < NT; t++) partials[(size_t)t * kMaxBlocks + blockIdx.x] = v[t];
        __threadfence();
        unsigned int prev = atomicAdd(ticket, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int t = 0; t < NT; t++) {
        double acc = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) acc += partials[(size_t)t * kMaxBlocks + b];
        v[t] = acc;
    }
    block_sum<NT>(v);
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}

// same with max
__device__ __forceinline__ bool grid_max_last(double& v, double* partials, unsigned int* ticket) {
    __shared__ bool is_last_m;
    __shared__ double shm[kThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) shm[w] = v;
    __syncthreads();
    if (w == 0) { double x = (lane < kThreads / 32) ? shm[lane] : -1.0e300; v = warp_max(x); }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        unsigned int prev = atomicAdd(ticket, 1u);
        is_last_m = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last_m) return false;
    __threadfence();
    double acc = -1.0e300;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) acc = fmax(acc, partials[b]);
    acc = warp_max(acc);
    if (lane == 0) shm[w] = acc;
    __syncthreads();
    if (w == 0) { double x = (lane < kThreads / 32) ? shm[lane] : -1.0e300; v = warp_max(x); }
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}

template <class T>
int dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    PF2_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    return PF2_OK;
}

}  // namespace pf2
