// spmv_tma.cuh -- CSR SpMV with a TMA (cp.async.bulk) multi-stage pipeline, the default kernel on sm_100a.
//
// The first kernel (spmv_stream_kernel) was latency-bound: 46 % of DRAM peak with 80 % of issue slots empty
// (profiles/r01_spmv_stream.txt) because every CTA alternated "load a tile" / "fold a tile".  Here the nonzero
// stream never waits for the arithmetic:
//   * a producer warp (one elected lane) walks the CTA's tiles ahead of the consumers and issues, per tile, three
//     bulk copies global -> shared (values, column indices, row pointers) that complete on an mbarrier
//     (SASS: UBLKCP + SYNCS); kStages tiles are in flight per CTA, so each SM keeps > 80 KB of HBM reads outstanding;
//   * 8 consumer warps wait on the tile's mbarrier, turn the staged (value, column) pairs into products with
//     gathered x (L2/L1 hits: x is touched ~18 times), park them in place, and fold each row with G threads;
//   * the stage is handed back through an "empty" mbarrier.
// A tile is kConsumers/G consecutive rows; its nonzeros are contiguous in CSR so every byte the TMA moves is used.
// Bulk copies need 16-byte aligned addresses and sizes, so a tile's window starts at (first nnz & ~3) and is a
// multiple of 4 entries long; the CSR arrays are allocated with 8 spare entries for the over-read.
#pragma once
#include "types.cuh"
#include "p2p.cuh"

namespace pf2 {

constexpr int kConsumers = 256;         // consumer threads per CTA (8 warps) + 1 producer warp
constexpr int kTmaThreads = kConsumers + 32;
constexpr int kMaxStages = 6;
constexpr int kTileNnz = 2592;          // largest usable nonzero count of a tile (32 rows x 81)

// Dynamic shared-memory layout (all runtime sized so that tile size / pipeline depth can be tuned per matrix):
//   stage s : val[cap] fp64 | idx[cap] int32 | rp[kConsumers+2] int64        (cap multiple of 4 -> 16-byte aligned parts)
//   tail    : full[kMaxStages] | empty[kMaxStages] mbarriers | red[8] | flag
struct TmaTail {
    unsigned long long full[kMaxStages];
    unsigned long long empty[kMaxStages];
    double red[8];
    int flag;
};
__host__ __device__ inline size_t tma_stage_bytes(int cap) { return (((size_t)cap * 12 + (kConsumers + 2) * 8) + 127) & ~(size_t)127; }
__host__ __device__ inline size_t tma_smem_bytes(int cap, int stages) { return tma_stage_bytes(cap) * stages + sizeof(TmaTail); }
// capacity needed for tiles of `rpb` rows when the longest row has `max_row` entries (+ alignment slack)
__host__ __device__ inline int tma_cap_for(int rpb, int max_row) { return ((rpb * max_row + 6) + 3) & ~3; }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk copy global -> shared::cta, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int G, bool DOT>
__global__ void __launch_bounds__(kTmaThreads, 4)
spmv_tma_kernel(int rows, const long long* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ data,
                const double* __restrict__ x, double* __restrict__ y, const CgState* __restrict__ st, double* dot_out,
                double* partials, unsigned int* ticket, int cap, int stages, int dot_lo, int dot_hi, const P2PView* p2p, unsigned long long* p2p_epoch) {
    if (DOT && st != nullptr && st->done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t stage_bytes = tma_stage_bytes(cap);
    TmaTail& sm = *reinterpret_cast<TmaTail*>(smem_raw + stage_bytes * stages);
    constexpr int RPB = kConsumers / G;
    const int ntiles = (rows + RPB - 1) / RPB;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kConsumers) {
        // ---------------- producer warp ----------------
        if (tid == kConsumers) {
            int s = 0, use = 0;      // stage index and how many times the ring has wrapped
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (use > 0) mbar_wait(&sm.empty[s], (use - 1) & 1);
                unsigned char* base = smem_raw + stage_bytes * s;
                double* sval = reinterpret_cast<double*>(base);
                int* sidx = reinterpret_cast<int*>(base + (size_t)cap * 8);
                long long* srp = reinterpret_cast<long long*>(base + (size_t)cap * 12);
                const int r0 = tile * RPB;
                const int nr = min(RPB, rows - r0);
                const long long b = indptr[r0], e = indptr[r0 + nr];
                const long long b_al = b & ~3LL;
                int cnt = (int)(e - b_al);
                cnt = max((cnt + 3) & ~3, 4);
                const unsigned rp_bytes = (unsigned)(((nr + 1) * 8 + 15) & ~15);
                mbar_expect_tx(&sm.full[s], (unsigned)cnt * 12u + rp_bytes);
                tma_load_1d(sval, data + b_al, (unsigned)cnt * 8u, &sm.full[s]);
                tma_load_1d(sidx, indices + b_al, (unsigned)cnt * 4u, &sm.full[s]);
                tma_load_1d(srp, indptr + r0, rp_bytes, &sm.full[s]);
                if (++s == stages) { s = 0; use++; }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int lr = tid / G, g = tid % G;
    double dot = 0.0;
    int s = 0, use = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&sm.full[s], use & 1);
        unsigned char* base = smem_raw + stage_bytes * s;
        double* sval = reinterpret_cast<double*>(base);
        const int* sidx = reinterpret_cast<const int*>(base + (size_t)cap * 8);
        const long long* srp = reinterpret_cast<const long long*>(base + (size_t)cap * 12);
        const int r0 = tile * RPB;
        const int nr = min(RPB, rows - r0);
        const long long b = srp[0], e = srp[nr];
        const long long b_al = b & ~3LL;
        const int lo = (int)(b - b_al), hi = (int)(e - b_al);
        // products in place (entries outside [lo, hi) are alignment padding: never dereference their columns)
#pragma unroll 4
        for (int j = lo + tid; j < hi; j += kConsumers) sval[j] = sval[j] * __ldg(x + sidx[j]);
        consumer_sync();
        double acc = 0.0;
        if (lr < nr) {
            const int rb = (int)(srp[lr] - b_al), re = (int)(srp[lr + 1] - b_al);
            for (int j = rb + g; j < re; j += G) acc += sval[j];
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, G);
        if (lr < nr && g == 0) {
            y[r0 + lr] = acc;
            if (DOT && r0 + lr >= dot_lo && r0 + lr < dot_hi) dot += acc * x[r0 + lr];
        }
        // the stage was written through the generic proxy (products); order that before the next bulk copy into it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        consumer_sync();
        if (tid == 0) mbar_arrive(&sm.empty[s]);
        if (++s == stages) { s = 0; use++; }
    }
    if (DOT) {
        // consumer-only deterministic grid reduction (the producer warp has left)
        dot = warp_sum(dot);
        if ((tid & 31) == 0) sm.red[tid >> 5] = dot;
        consumer_sync();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < 8; w++) v += sm.red[w];
            partials[blockIdx.x] = v;
            __threadfence();
            const unsigned prev = atomicAdd(ticket, 1u);
            sm.flag = (prev == gridDim.x - 1);
        }
        consumer_sync();
        if (sm.flag) {
            __threadfence();
            double acc = 0.0;
            for (unsigned b = tid; b < gridDim.x; b += kConsumers) acc += partials[b];
            acc = warp_sum(acc);
            consumer_sync();
            if ((tid & 31) == 0) sm.red[tid >> 5] = acc;
            consumer_sync();
            if (tid == 0) {
                double v = 0.0;
                for (int w = 0; w < 8; w++) v += sm.red[w];
                sm.red[0] = v;
                *ticket = 0u;
            }
            consumer_sync();
            if (p2p == nullptr) { if (tid == 0) *dot_out = sm.red[0]; }
            else if (tid < 32) { p2p_allreduce_warp(*p2p, p2p_epoch, sm.red, 1); if (tid == 0) *dot_out = sm.red[0]; }
        }
    }
}

}  // namespace pf2
