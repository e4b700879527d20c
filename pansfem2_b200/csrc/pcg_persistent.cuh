// pcg_persistent.cuh -- CG / ScalingCG (CG.h:124-154, 420-453) as ONE persistent cooperative kernel per solve.
//
// The three-kernel loop of solver.cu is at the HBM bound on one GPU once the matrix is large, but every iteration has three
// host-ordered synchronisation points: on a partitioned matrix (peers over NVLink) and on small systems those, not bytes, set the
// time per iteration.  Here one cooperative grid (every CTA resident) runs the whole solve:
//
//     init      r = b - A x0 (x0 = 0: r = b) ; z = r / D ; p = z ; b.b, z.r, r.r                         REDUCE + BARRIER
//     loop      y = A p on the SELL-32 mirror (interior slices first), p.y                               REDUCE  -> alpha
//               x += alpha p ; r -= alpha y ; z = r / D ; z.r, r.r                                       REDUCE  -> beta, ||r|| test
//               p = beta p + z  (boundary slices first: they also go into the neighbours' ghost ranges)  BARRIER
//
// Row ownership.  A warp owns the same SELL slices in every phase and lane l owns row l of each: y, z, x, r and D are only ever
// touched by their owner thread, so the two REDUCE points need NO memory ordering at all.  A REDUCE moves values, not memory: every
// CTA publishes its partial sums as self-validating 8-byte {half of the double, generation} words (the LL protocol of NCCL), CTA 0
// collects them in CTA order (deterministic), exchanges the total with the peer GPUs the same way over NVLink (rank order: bitwise
// identical on every rank) and publishes the result the same way: no atomics, no fences, one L2 round trip per hop.  Only p is read
// by other threads (the gathers of the next product): the BARRIER after the p-update is the same exchange wrapped in a release
// fence before and an acquire fence after (SASS: MEMBAR + CCTL.IVALL, which also drops the SM's L1 lines).
// Row-block partition (DIST): the owner of a boundary-plane row stores the new p straight into the neighbour's ghost range; the
// warp that completes a plane raises the neighbour's halo flag.  In the next product only the warps that reach a boundary slice
// wait for that flag, after their interior slices: the exchange overlaps the interior rows.
// alpha, beta and the convergence decision are computed redundantly by every thread from the published sums: one launch per solve.
//
// Vectors written inside the kernel are never read through the non-coherent path (no __ldg / const __restrict__).  Every spin is
// bounded (kSpinLimitNs): a peer that died turns into PF2_E_CUDA on the host instead of a hung box.
#pragma once
#include "types.cuh"
#include "p2p.cuh"
#include "spmv_sell.cuh"

namespace pf2 {

#ifndef PF2_PCG_BACKOFF
#define PF2_PCG_BACKOFF 0
#endif
// resident CTAs per SM the kernel is compiled for: 6 -> 40 registers (the product and the vector phases then run without spills inside
// their loops; at 8 -> 32 registers ncu showed as much local-memory traffic as algorithmic traffic)
#ifndef PF2_PCG_MINB
#define PF2_PCG_MINB 6
#endif
constexpr long long kSpinLimitClk = 8000000000ll;     // ~4 s of SM clocks: three orders of magnitude above any legitimate wait
// (timeouts use clock64: %globaltimer costs a long round trip and must stay off the exchange's critical path)

constexpr int kPcgMaxCtas = 2048;        // >= CTAs of any cooperative grid on this part (148 SMs x 8)
constexpr int kPcgTerms = 4;
constexpr int kPcgDbgIter = 5;

struct PcgSync {                     // device memory, zeroed before every launch
    // LL words: [parity][term][CTA][2 halves]: a CTA's partial sums of the current generation; [parity][term][2]: the result
    unsigned long long part[2][kPcgTerms][kPcgMaxCtas][2];
    // the totals come back through one mailbox per CTA (same layout): thousands of threads polling ONE line would be served serially
    // by its L2 slice (measured: 8-10 us per exchange), private lines are polled by one CTA each
    unsigned long long mbox[2][kPcgTerms][kPcgMaxCtas][2];
    unsigned int abort;              // a bounded spin timed out: every CTA leaves at its next synchronisation point
    unsigned int halo_count[2];      // boundary-plane rows pushed to the left / right neighbour in the current exchange
    unsigned int pad;
    unsigned long long dbg[6][kPcgMaxCtas];   // iteration kPcgDbgIter: %globaltimer of every CTA at the start / end of its share of the three phases
    unsigned long long t_ns[8];      // CTA 0's view: ns in the product / update / p-update phases (synchronisation included), iterations,
                                     // and the part of each phase CTA 0 spent inside the exchange (tail of the grid + latency)
};

struct PcgArgs {
    int rows, nslices, own_lo, own_hi, itrmax, warm;
    double eps;
    const long long* slice_ptr;
    const int* perm;                 // SELL-C-sigma: row of every slot (-1 = padding lane); nullptr = natural order
    const void* sell_idx;
    const double* sell_val;
    const long long* indptr;         // canonical CSR (GetDiagonal, CG.h:398-404)
    const int* diagpos;
    const double* data;
    const double* b;
    double *x, *r, *z, *p, *y, *dvec;
    CgState* st;
    PcgSync* sync;
    // row-block partition (peer-memory backend)
    const P2PView* p2p;
    unsigned long long* epoch;       // [0] allreduce epoch, [1] halo epoch (device memory, persistent across solves)
    int sendL, cntL, sendR, cntR;
};

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) { return p2p_ld_sys_u64(p); }
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// bounded wait for an LL word; false = timed out or another CTA aborted (v is then meaningless)
template <bool SYS>
__device__ __forceinline__ bool ll_wait(const unsigned long long* slot, unsigned int flag, double& v, PcgSync* sync) {
    if (ll_try_load<SYS>(slot, flag, v)) return true;
    long long t0 = 0;
    for (unsigned int n = 1;; n++) {
        if (ll_try_load<SYS>(slot, flag, v)) return true;
        if ((n & 1023u) == 0u) {
            if (t0 == 0) t0 = clock64();
            if (ld_relaxed_u32(&sync->abort)) return false;
            if (clock64() - t0 > kSpinLimitClk) { atomicExch(&sync->abort, 1u); return false; }
        }
    }
}
// the same on a plain 8-byte flag a PEER GPU raises (halo epochs)
__device__ __forceinline__ bool spin_sys_u64(const unsigned long long* flag, unsigned long long target, PcgSync* sync) {
    if (ld_relaxed_sys_u64(flag) >= target) return true;
    long long t0 = 0;
    for (unsigned int n = 1;; n++) {
        if (ld_relaxed_sys_u64(flag) >= target) return true;
        if ((n & 1023u) == 0u) {
            if (t0 == 0) t0 = clock64();
            if (ld_relaxed_u32(&sync->abort)) return false;
            if (clock64() - t0 > kSpinLimitClk) { atomicExch(&sync->abort, 1u); return false; }
        }
    }
}

// Per-thread synchronisation state: generation of the grid exchange, epoch of the cross-GPU one.
struct PcgGen {
    unsigned int gen;
    unsigned long long xepoch;
};

// Grid-wide sum of NT terms (NT = 0: nothing to sum) delivered to every thread; FENCE: also a memory barrier (everything written
// before it by any thread of the grid is visible to every thread after it).  Returns false when the solve must be abandoned.
template <int NT, bool DIST, bool FENCE>
__device__ __noinline__ bool grid_exchange(double* v, const PcgArgs& a, PcgGen& g) {
    __shared__ double s_tot[kPcgTerms];
    __shared__ int s_bad;
    PcgSync* sync = a.sync;
    if (threadIdx.x == 0) s_bad = 0;
    constexpr int NW = NT > 0 ? NT : 1;          // words on the wire (a plain barrier sends one dummy)
    const unsigned int gen = ++g.gen;
    const int par = (int)(gen & 1u);
    double w[NW];
#pragma unroll
    for (int t = 0; t < NW; t++) w[t] = (t < NT) ? v[t] : 0.0;
    block_sum<NW>(w);                            // ends with __syncthreads: every thread of the CTA has issued its stores
    if (threadIdx.x == 0) {
        if (FENCE) __threadfence();              // release: the CTA's stores (cumulative over the bar.sync) before the arrival word
#pragma unroll
        for (int t = 0; t < NW; t++) ll_store<false>(&sync->part[par][t][blockIdx.x][0], w[t], gen);
    }
    if (blockIdx.x == 0) {
        // collector: partials in CTA order -> per-thread sums in a fixed order -> block sum: deterministic for a given grid.
        // A thread's words are all requested before the first is examined (one L2 round trip when everybody has arrived).
        constexpr int kPer = kPcgMaxCtas / kThreads;
        double acc[NW], got[NW][kPer];
        unsigned int pending = 0u;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            const unsigned int c = threadIdx.x + j * kThreads;
            if (c < gridDim.x) {
#pragma unroll
                for (int t = 0; t < NW; t++) { got[t][j] = 0.0; if (!ll_try_load<false>(&sync->part[par][t][c][0], gen, got[t][j])) pending |= 1u << (j * NW + t); }
            }
        }
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            const unsigned int c = threadIdx.x + j * kThreads;
#pragma unroll
            for (int t = 0; t < NW; t++) if (pending & (1u << (j * NW + t))) { if (!ll_wait<false>(&sync->part[par][t][c][0], gen, got[t][j], sync)) s_bad = 1; }
        }
#pragma unroll
        for (int t = 0; t < NW; t++) {
            acc[t] = 0.0;
#pragma unroll
            for (int j = 0; j < kPer; j++) if (threadIdx.x + j * kThreads < gridDim.x) acc[t] += got[t][j];
        }
        block_sum<NW>(acc);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int t = 0; t < NW; t++) s_tot[t] = acc[t];
        }
        __syncthreads();
        if (DIST && NT > 0 && threadIdx.x < 32) {
            // cross-GPU: lane r sends this rank's totals to rank r and receives rank r's; lane 0 adds them in rank order
            const P2PView& P = *a.p2p;
            const unsigned long long epoch = g.xepoch + 1;
            const unsigned int flag = (unsigned int)epoch;
            const int xpar = (int)(epoch & 1ull), lane = threadIdx.x;
            double got[NW];
#pragma unroll
            for (int t = 0; t < NW; t++) got[t] = 0.0;
            if (lane < P.world) {
#pragma unroll
                for (int t = 0; t < NW; t++) ll_store<true>(P.ll[lane] + (((size_t)xpar * P.world + P.rank) * kPcgTerms + t) * 2, s_tot[t], flag);
#pragma unroll
                for (int t = 0; t < NW; t++) { if (!ll_wait<true>(P.ll[P.rank] + (((size_t)xpar * P.world + lane) * kPcgTerms + t) * 2, flag, got[t], sync)) s_bad = 1; }
            }
#pragma unroll
            for (int t = 0; t < NW; t++) {
                double sum = 0.0;
                for (int r = 0; r < P.world; r++) sum += __shfl_sync(0xffffffffu, got[t], r);
                if (lane == 0) s_tot[t] = sum;
            }
            __syncwarp();
        }
        if (DIST && NT > 0) g.xepoch++;
        __syncthreads();
        for (unsigned int c = threadIdx.x; c < gridDim.x; c += blockDim.x) {
#pragma unroll
            for (int t = 0; t < NW; t++) ll_store<false>(&sync->mbox[par][t][c][0], s_tot[t], gen);
        }
    } else if (DIST && NT > 0) g.xepoch++;
    if (blockIdx.x != 0 && threadIdx.x < NW) {
        double x = 0.0;
        if (!ll_wait<false>(&sync->mbox[par][threadIdx.x][blockIdx.x][0], gen, x, sync)) s_bad = 1;
        s_tot[threadIdx.x] = x;
    }
    if (FENCE && threadIdx.x == 0) __threadfence();      // acquire side: drops this SM's L1 lines (CCTL.IVALL)
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NT; t++) v[t] = s_tot[t];
    const bool ok = (s_bad == 0);
    __syncthreads();                                     // s_tot / s_bad are reused by the next exchange
    return ok;
}

// ---- slice ownership --------------------------------------------------------------------------------------------------------
// Slices [s_lo, s_hi) hold this rank's owned rows; those holding rows of the first / last owned node plane (the send ranges, whose
// columns reach the ghosts) are "boundary" slices.  Virtual order: interior slices first, then left boundary, then right boundary.
// Warp w owns the virtual indices w, w + nwarps, ... in EVERY phase.
struct PcgSlices {
    int s_lo, s_lb, s_rb, s_hi, n_int, n_left, n_all;
};
template <bool DIST>
__device__ __forceinline__ PcgSlices pcg_slices(const PcgArgs& a) {
    PcgSlices S;
    S.s_lo = a.own_lo / kSellC; S.s_hi = (a.own_hi + kSellC - 1) / kSellC;
    S.s_lb = S.s_lo; S.s_rb = S.s_hi;
    if (DIST) {
        if (a.cntL > 0) S.s_lb = min(S.s_hi, (a.sendL + a.cntL + kSellC - 1) / kSellC);
        if (a.cntR > 0) S.s_rb = max(S.s_lb, a.sendR / kSellC);
    }
    S.n_int = S.s_rb - S.s_lb; S.n_left = S.s_lb - S.s_lo; S.n_all = S.s_hi - S.s_lo;
    return S;
}
// slice of virtual index v; side = 0 interior, 1 left boundary, 2 right boundary
__device__ __forceinline__ int pcg_slice_of(const PcgSlices& S, int v, int& side) {
    if (v < S.n_int) { side = 0; return S.s_lb + v; }
    const int w = v - S.n_int;
    if (w < S.n_left) { side = 1; return S.s_lo + w; }
    side = 2;
    return S.s_rb + (w - S.n_left);
}
__device__ __forceinline__ int pcg_row(const PcgArgs& a, int s, int lane) {
    const int slot = s * kSellC + lane;
    const int r = a.perm ? a.perm[slot] : (slot < a.rows ? slot : -1);
    return (r >= a.own_lo && r < a.own_hi) ? r : -1;
}

// y = A x over the owned slices, interior first; returns this thread's share of x.y over its rows.  DIST: a warp that reaches a
// boundary slice first waits for that neighbour's plane of the current halo epoch.
template <class IDX, int NB, bool DIST, bool WAIT, bool CS, bool DOT>
__device__ __noinline__ double pcg_product(const PcgArgs& a, const double* x, unsigned long long halo_epoch) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const PcgSlices S = pcg_slices<DIST>(a);
    bool waitedL = false, waitedR = false;
    double dot = 0.0;
    for (int v = warp; v < S.n_all; v += nwarps) {
        int side;
        const int s = pcg_slice_of(S, v, side);
        if (DIST && WAIT && side != 0) {
            // a slab thinner than two planes has slices that touch both ghost planes: wait for both sides then
            const bool needL = (side == 1 || S.n_int == 0) && a.cntL > 0 && !waitedL;
            const bool needR = (side == 2 || S.n_int == 0) && a.cntR > 0 && !waitedR;
            if (needL || needR) {
                if (lane == 0) {
                    const unsigned long long* mine = a.p2p->halo_flags[a.p2p->rank];
                    if (needL) spin_sys_u64(mine + 0, halo_epoch, a.sync);
                    if (needR) spin_sys_u64(mine + 1, halo_epoch, a.sync);
                    __threadfence_system();          // acquire: the planes the neighbour stored before raising the flag; drops stale L1 lines
                }
                __syncwarp();
                waitedL = waitedL || needL; waitedR = waitedR || needR;
            }
        }
        const long long base = a.slice_ptr[s];
        const int width = (int)((a.slice_ptr[s + 1] - base) / kSellC);
        const int slot = s * kSellC + lane;
        const int rr = a.perm ? a.perm[slot] : (slot < a.rows ? slot : -1);
        const double acc = sell_slice_acc<IDX, NB, 6, CS, false>((const IDX*)a.sell_idx, a.sell_val, x, base, width, lane, rr);
        if (rr >= a.own_lo && rr < a.own_hi) {
            a.y[rr] = acc;
            if (DOT) dot += acc * x[rr];
        }
    }
    return dot;
}

// x += alpha p ; r -= alpha y ; z = r / D ; this thread's share of {z.r, r.r}   (CG.h:434-437, 443).  Owner rows only.
template <int MODE, bool DIST>
__device__ __noinline__ void pcg_update(const PcgArgs& a, double alpha, double* w) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const PcgSlices S = pcg_slices<DIST>(a);
    double* __restrict__ x = a.x;
    double* __restrict__ r = a.r;
    double* __restrict__ z = a.z;
    const double* p = a.p;
    const double* y = a.y;
    const double* dv = a.dvec;
    double zr = 0.0, rr = 0.0;
    for (int v = warp; v < S.n_all; v += nwarps) {
        int side;
        const int i = pcg_row(a, pcg_slice_of(S, v, side), lane);
        if (i < 0) continue;
        const double xi = x[i], pi = p[i], ri = r[i], yi = y[i];
        const double rn = ri + (-alpha) * yi;
        x[i] = xi + alpha * pi;
        r[i] = rn;
        rr += rn * rn;
        if (MODE == 0) zr += rn * rn;
        else { const double zi = rn / dv[i]; z[i] = zi; zr += zi * rn; }
    }
    w[0] = zr; w[1] = rr;
}

// p = beta p + z on the owner rows (CG.h:439), boundary slices first.  DIST: rows of a send range also go straight into the
// neighbour's ghost range; the warp that completes a plane raises that neighbour's halo flag.  INIT: p already holds its values.
template <bool DIST, bool INIT>
__device__ __noinline__ void pcg_pupdate(const PcgArgs& a, const double* zv, double beta, unsigned long long halo_epoch) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const PcgSlices S = pcg_slices<DIST>(a);
    if (S.n_all <= warp) return;
    double* __restrict__ p = a.p;
    if (DIST) {
        // boundary slices of this warp (virtual indices >= n_int), then publish
        int pushed[2] = { 0, 0 };
        const P2PView& P = *a.p2p;
        int vb = warp;
        if (vb < S.n_int) vb += ((S.n_int - vb + nwarps - 1) / nwarps) * nwarps;
        for (int v = vb; v < S.n_all; v += nwarps) {
            int side;
            const int i = pcg_row(a, pcg_slice_of(S, v, side), lane);
            if (i < 0) continue;
            double pi = p[i];
            if (!INIT) { pi = beta * pi + zv[i]; p[i] = pi; }
            if (i >= a.sendL && i < a.sendL + a.cntL) { P.left_p[P.left_recv_off + (i - a.sendL)] = pi; pushed[0]++; }
            if (i >= a.sendR && i < a.sendR + a.cntR) { P.right_p[P.right_recv_off + (i - a.sendR)] = pi; pushed[1]++; }
        }
        const int nl = __reduce_add_sync(0xffffffffu, pushed[0]), nr = __reduce_add_sync(0xffffffffu, pushed[1]);
        if (nl + nr > 0) {
            __threadfence_system();              // the pushed rows are visible at the neighbours before they are accounted for
            __syncwarp();
            if (lane == 0) {
                if (nl > 0 && atomicAdd(&a.sync->halo_count[0], (unsigned int)nl) + nl == (unsigned int)a.cntL) {
                    a.sync->halo_count[0] = 0u;
                    *(volatile unsigned long long*)(P.halo_flags[P.rank - 1] + 1) = halo_epoch;      // I am the RIGHT neighbour of rank-1
                }
                if (nr > 0 && atomicAdd(&a.sync->halo_count[1], (unsigned int)nr) + nr == (unsigned int)a.cntR) {
                    a.sync->halo_count[1] = 0u;
                    *(volatile unsigned long long*)(P.halo_flags[P.rank + 1] + 0) = halo_epoch;      // ... and the LEFT neighbour of rank+1
                }
            }
        }
    }
    if (INIT) return;
    const int n_rest = DIST ? S.n_int : S.n_all;         // not partitioned: every slice is "interior"
    for (int v = warp; v < n_rest; v += 4 * nwarps) {
        int i[4];
        double pi[4], zi[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int side;
            i[u] = (v + u * nwarps < n_rest) ? pcg_row(a, pcg_slice_of(S, v + u * nwarps, side), lane) : -1;
            const int k = max(i[u], 0);
            pi[u] = p[k]; zi[u] = zv[k];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) if (i[u] >= 0) p[i[u]] = beta * pi[u] + zi[u];
    }
}

// set-up on the owner rows: r = b - A x0 (y holds A x0 of a warm start), D, z = r / D, p = z; share of {b.b, z.r, r.r}  (CG.h:422-428)
template <int MODE, bool DIST>
__device__ __noinline__ void pcg_setup(const PcgArgs& a, double* v) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int warp = tid >> 5, nwarps = nth >> 5;
    const PcgSlices S = pcg_slices<DIST>(a);
    // ghost rows (partitioned matrix): p arrives by halo exchange, x of a cold start is defined as 0, the rest is never read
    for (int i = tid; i < a.rows; i += nth) {
        if (i >= a.own_lo && i < a.own_hi) continue;
        if (!a.warm) a.x[i] = 0.0;
        a.p[i] = 0.0;
    }
    double bb = 0.0, zr = 0.0, rr = 0.0;
    for (int vv = warp; vv < S.n_all; vv += nwarps) {
        int side;
        const int i = pcg_row(a, pcg_slice_of(S, vv, side), lane);
        if (i < 0) continue;
        const double bi = a.b[i];
        double ri = bi;
        if (a.warm) ri = bi - a.y[i]; else a.x[i] = 0.0;
        a.r[i] = ri;
        double zi = ri;
        if (MODE == 1) {
            const int dp = a.diagpos[i];
            const double d = dp >= 0 ? a.data[a.indptr[i] + dp] : 0.0;      // GetDiagonal (CG.h:398-404)
            a.dvec[i] = d;
            zi = ri / d;
            a.z[i] = zi;
        }
        a.p[i] = zi;
        bb += bi * bi; zr += zi * ri; rr += ri * ri;
    }
    v[0] = bb; v[1] = zr; v[2] = rr;
}

// MODE 0: CG (z = r), 1: ScalingCG (z = r / diag).  Launch cooperatively with gridDim.x <= min(resident CTAs, kPcgMaxCtas).
template <class IDX, int NB, int MODE, bool DIST, bool CS>
__global__ void __launch_bounds__(kThreads, PF2_PCG_MINB)
pcg_persistent_kernel(const __grid_constant__ PcgArgs a) {
    double* const zv = (MODE == 1) ? a.z : a.r;
    PcgGen g;
    g.gen = 0u;
    g.xepoch = DIST ? *(volatile unsigned long long*)(a.epoch + 0) : 0ull;
    unsigned long long halo_epoch = DIST ? *(volatile unsigned long long*)(a.epoch + 1) : 0ull;
    const bool timing = (blockIdx.x == 0 && threadIdx.x == 0);
    // phase split in SM clocks (clock64 is local and cheap), converted with the kernel's own ns / clock ratio at the end
    long long t_acc[6] = { 0, 0, 0, 0, 0, 0 }, t_prev = 0, t_x = 0, c_begin = 0;
    unsigned long long ns_begin = 0ull;
    bool ok = true;

    // ---- set-up (CG.h:422-428) ---------------------------------------------------------------------------------------
    if (a.warm) pcg_product<IDX, NB, DIST, false, CS, false>(a, a.x, 0ull);     // y = A x0 (ghost entries of x0 are valid); owner rows only
    double v[3];
    pcg_setup<MODE, DIST>(a, v);
    ok = grid_exchange<3, DIST, true>(v, a, g);
    const double bb = v[0];
    double rho = v[1], rr = v[2], beta = 0.0;
    int iter = 0;
    // a warm start may already satisfy the stopping rule; from x0 = 0 the reference always iterates (b = 0 never converges)
    bool done = a.warm && (sqrt(rr) < a.eps * sqrt(bb));
    if (DIST && ok && !done) {
        halo_epoch++;
        pcg_pupdate<true, true>(a, nullptr, 0.0, halo_epoch);            // first exchange of the boundary planes of p
    }
    if (timing) { ns_begin = global_ns(); c_begin = clock64(); t_prev = c_begin; }

    // ---- iterations (CG.h:430-449) -------------------------------------------------------------------------------------
    while (ok && !done && iter < a.itrmax) {
        double d1[1];
        const bool dbg = (iter == kPcgDbgIter && threadIdx.x == 0);
        if (dbg) a.sync->dbg[0][blockIdx.x] = global_ns();
        d1[0] = pcg_product<IDX, NB, DIST, DIST, CS, true>(a, a.p, halo_epoch);
        if (dbg) a.sync->dbg[1][blockIdx.x] = global_ns();
        if (timing) t_x = clock64();
        ok = grid_exchange<1, DIST, false>(d1, a, g);
        if (timing) t_acc[3] += clock64() - t_x;
        if (!ok) break;
        if (timing) { const long long t = clock64(); t_acc[0] += t - t_prev; t_prev = t; }
        const double alpha = rho / d1[0];
        double w[2];
        if (dbg) a.sync->dbg[2][blockIdx.x] = global_ns();
        pcg_update<MODE, DIST>(a, alpha, w);
        if (dbg) a.sync->dbg[3][blockIdx.x] = global_ns();
        if (timing) t_x = clock64();
        ok = grid_exchange<2, DIST, false>(w, a, g);
        if (timing) t_acc[4] += clock64() - t_x;
        if (!ok) break;
        if (timing) { const long long t = clock64(); t_acc[1] += t - t_prev; t_prev = t; }
        beta = w[0] / rho;
        rho = w[0];
        rr = w[1];
        iter++;
        if (sqrt(rr) < a.eps * sqrt(bb)) { done = true; break; }
        if (iter >= a.itrmax) break;
        if (DIST) halo_epoch++;
        if (dbg) a.sync->dbg[4][blockIdx.x] = global_ns();
        pcg_pupdate<DIST, false>(a, zv, beta, halo_epoch);
        if (dbg) a.sync->dbg[5][blockIdx.x] = global_ns();
        if (timing) t_x = clock64();
        ok = grid_exchange<0, DIST, true>(nullptr, a, g);
        if (timing) t_acc[5] += clock64() - t_x;
        if (timing) { const long long t = clock64(); t_acc[2] += t - t_prev; t_prev = t; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CgState* st = a.st;
        st->bb = bb; st->rr = rr; st->rho = rho; st->beta = beta; st->pAp = 0.0;
        st->iter = iter; st->done = ok ? (done ? 1 : 0) : 2; st->maxit = a.itrmax; st->eps = a.eps;
        if (DIST) { *(volatile unsigned long long*)(a.epoch + 0) = g.xepoch; *(volatile unsigned long long*)(a.epoch + 1) = halo_epoch; }
        const long long dclk = clock64() - c_begin;
        const double ns_per_clk = dclk > 0 ? (double)(global_ns() - ns_begin) / (double)dclk : 0.0;
        for (int j = 0; j < 3; j++) {
            a.sync->t_ns[j] = (unsigned long long)((double)t_acc[j] * ns_per_clk);
            a.sync->t_ns[4 + j] = (unsigned long long)((double)t_acc[3 + j] * ns_per_clk);
        }
        a.sync->t_ns[3] = (unsigned long long)iter; a.sync->t_ns[7] = 0ull;
    }
}

}  // namespace pf2
