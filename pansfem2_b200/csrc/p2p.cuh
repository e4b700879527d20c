// p2p.cuh -- peer-memory collectives usable from inside any kernel (NVLink / NVSwitch, IPC-mapped arenas; see dist.cu).
#pragma once
#include "types.cuh"

namespace pf2 {

// Allreduce (sum) of <= 4 fp64 executed by ONE WARP: lane r stores this rank's values into rank r's arena, fences, raises its
// flag there; then lane r waits for rank r's flag in the LOCAL arena and lane 0 sums the slots in rank order, so the result
// is bitwise identical on every rank.  Slots and flags are double-buffered by epoch parity: a rank cannot run two epochs
// ahead of a peer because finishing an epoch needs that peer's flag.  `vals` must be visible to the whole warp (shared).
__device__ __forceinline__ void p2p_allreduce_warp(const P2PView& P, unsigned long long* epoch_ctr, double* vals, int count) {
    const int lane = threadIdx.x & 31;
    const unsigned long long epoch = *epoch_ctr + 1;
    const int par = (int)(epoch & 1ull);
    if (lane < P.world) {
        double* dst = P.slots[lane] + ((size_t)par * P.world + P.rank) * 4;
        for (int c = 0; c < count; c++) dst[c] = vals[c];
        __threadfence_system();
        *(volatile unsigned long long*)(P.flags[lane] + (size_t)par * P.world + P.rank) = epoch;
    }
    if (lane < P.world) {
        volatile unsigned long long* f = (volatile unsigned long long*)(P.flags[P.rank] + (size_t)par * P.world + lane);
        while (*f < epoch) {}
    }
    __syncwarp();
    __threadfence_system();
    if (lane == 0) {
        const volatile double* src = (const volatile double*)(P.slots[P.rank] + (size_t)par * P.world * 4);
        for (int c = 0; c < count; c++) {
            double acc = 0.0;
            for (int r = 0; r < P.world; r++) acc += src[r * 4 + c];
            vals[c] = acc;
        }
        *epoch_ctr = epoch;
    }
    __syncwarp();
}

// Epilogue of a fused SpMV + dot kernel, called by ALL threads of the last CTA (grid_sum_last returned true; thread 0 holds
// the local sum): single GPU -> store; partitioned with the peer-memory backend -> allreduce across the box first.
__device__ __forceinline__ void finish_dot(double local_sum, double* dot_out, const P2PView* p2p, unsigned long long* epoch_ctr) {
    if (p2p == nullptr) {
        if (threadIdx.x == 0) *dot_out = local_sum;
        return;
    }
    __shared__ double p2p_vals[4];
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) p2p_vals[0] = local_sum;
        __syncwarp();
        p2p_allreduce_warp(*p2p, epoch_ctr, p2p_vals, 1);
        if (threadIdx.x == 0) *dot_out = p2p_vals[0];
    }
}

}  // namespace pf2
