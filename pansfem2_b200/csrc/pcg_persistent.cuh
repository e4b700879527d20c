// pcg_persistent.cuh -- CG / ScalingCG (CG.h:124-154, 420-453) as ONE persistent cooperative kernel per solve.
//
// The three-kernel loop of solver.cu pays three launch boundaries (drain + ramp) per iteration and, on a partitioned
// matrix, three host-ordered synchronisation points with the peers.  Here one cooperative grid (every CTA resident) runs the
// whole solve:
//
//     init      r = b - A x0 (x0 = 0: r = b) ; z = r / D ; p = z ; b.b, z.r, r.r                         reduce
//     loop      y = A p on the SELL-32 mirror, p.y                                                       reduce  -> alpha
//               x += alpha p ; r -= alpha y ; z = r / D ; z.r, r.r                                       reduce  -> beta, ||r|| test
//               p = beta p + z  (boundary planes first: they are stored straight into the neighbours' ghost ranges)
//                                                                                                        barrier
//
// A `reduce` is a grid barrier that carries a deterministic sum: every CTA parks its partial, the LAST CTA to arrive folds
// them in CTA order, (partitioned runs: allreduces the result with the peers over NVLink peer memory, rank order, bitwise
// identical everywhere), publishes it and releases the others.  alpha, beta and the convergence decision are computed
// redundantly by every thread from the published sums, so no state round-trips through the host: one launch per solve.
// Row-block partition (DIST): the halo of p is pushed by the CTAs that update the boundary planes, and only the warps that
// multiply the boundary slices of the NEXT product wait for the neighbours' planes (interior slices are processed first), i.e.
// the exchange overlaps the interior rows.
//
// Memory-model notes.  Vectors written inside the kernel (p, x, r, z, y, D) are never read through the non-coherent path
// (no __ldg / const __restrict__); every barrier ends with a gpu-scope fence executed after the release flag was observed
// (SASS: MEMBAR + CCTL.IVALL, which drops the SM's L1 lines), the halo wait with a system-scope one.  Every spin is bounded
// (kSpinLimitNs): a peer that died turns into PF2_E_CUDA on the host instead of a hung box.
#pragma once
#include "types.cuh"
#include "p2p.cuh"
#include "spmv_sell.cuh"

namespace pf2 {

#ifndef PF2_PCG_MINB
#define PF2_PCG_MINB 8
#endif
constexpr unsigned long long kSpinLimitNs = 4000000000ull;   // 4 s: three orders of magnitude above any legitimate wait

struct PcgSync {                     // device memory, zeroed before every launch
    unsigned int arrive;             // barrier arrivals of the current generation (reset by the last arriver)
    unsigned int abort;              // a bounded spin timed out: every CTA leaves at its next barrier
    unsigned int halo_arrive[2];     // CTAs that finished pushing the left / right boundary plane
    unsigned int pad0[28];
    unsigned int release;            // last generation released
    unsigned int pad1[31];
    double result[2][4];             // published sums, double-buffered by generation parity
    unsigned long long t_ns[4];      // CTA 0's view: ns in the product / update / p-update phases (barriers included), iterations
};

struct PcgArgs {
    int rows, nslices, own_lo, own_hi, itrmax, warm;
    double eps;
    const long long* slice_ptr;
    const void* sell_idx;
    const double* sell_val;
    const long long* indptr;         // canonical CSR (GetDiagonal, CG.h:398-404)
    const int* diagpos;
    const double* data;
    const double* b;
    double *x, *r, *z, *p, *y, *dvec;
    CgState* st;
    double* partials;
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
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= target (gpu scope); false = timed out or another CTA aborted
__device__ __forceinline__ bool spin_u32(const unsigned int* flag, unsigned int target, PcgSync* sync) {
    if (ld_relaxed_u32(flag) >= target) return true;
    const unsigned long long t0 = global_ns();
    for (unsigned int n = 1;; n++) {
        if (ld_relaxed_u32(flag) >= target) return true;
        if ((n & 255u) == 0u) {
            if (ld_relaxed_u32(&sync->abort)) return false;
            if (global_ns() - t0 > kSpinLimitNs) { atomicExch(&sync->abort, 1u); return false; }
        }
    }
}
// the same on a flag a PEER GPU writes (system scope)
__device__ __forceinline__ bool spin_sys_u64(const unsigned long long* flag, unsigned long long target, PcgSync* sync) {
    if (ld_relaxed_sys_u64(flag) >= target) return true;
    const unsigned long long t0 = global_ns();
    for (unsigned int n = 1;; n++) {
        if (ld_relaxed_sys_u64(flag) >= target) return true;
        if ((n & 255u) == 0u) {
            if (ld_relaxed_u32(&sync->abort)) return false;
            if (global_ns() - t0 > kSpinLimitNs) { atomicExch(&sync->abort, 1u); return false; }
        }
    }
}

// Allreduce (sum) of <= 4 fp64 across the box by ONE WARP, bounded spins (protocol of p2p_allreduce_warp, p2p.cuh).
__device__ __forceinline__ void pcg_allreduce_warp(const P2PView& P, unsigned long long* epoch_ctr, double* vals, int count, PcgSync* sync) {
    const int lane = threadIdx.x & 31;
    const unsigned long long epoch = *(volatile unsigned long long*)epoch_ctr + 1;
    const int par = (int)(epoch & 1ull);
    if (lane < P.world) {
        double* dst = P.slots[lane] + ((size_t)par * P.world + P.rank) * 4;
        for (int c = 0; c < count; c++) dst[c] = vals[c];
        __threadfence_system();
        *(volatile unsigned long long*)(P.flags[lane] + (size_t)par * P.world + P.rank) = epoch;
        spin_sys_u64(P.flags[P.rank] + (size_t)par * P.world + lane, epoch, sync);
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
        *(volatile unsigned long long*)epoch_ctr = epoch;
    }
    __syncwarp();
}

// Grid barrier carrying a deterministic sum of NT terms (NT = 0: plain barrier).  On return every thread of the grid holds the
// totals in v.  Returns false when the solve must be abandoned (a bounded spin timed out somewhere).
template <int NT, bool DIST>
__device__ __noinline__ bool grid_reduce_bcast(double* v, const PcgArgs& a, unsigned int& gen) {
    __shared__ int s_last;
    __shared__ double s_tot[4];
    PcgSync* sync = a.sync;
    if constexpr (NT > 0) {
        double w[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) w[t] = v[t];
        block_sum<NT>(w);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int t = 0; t < NT; t++) a.partials[(size_t)t * kMaxBlocks + blockIdx.x] = w[t];
        }
    } else {
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&sync->arrive, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        if constexpr (NT > 0) {
            __threadfence();
            double w[NT];
#pragma unroll
            for (int t = 0; t < NT; t++) {
                double acc = 0.0;
                for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) acc += __ldcg(a.partials + (size_t)t * kMaxBlocks + b);
                w[t] = acc;
            }
            block_sum<NT>(w);
            if (threadIdx.x == 0) {
#pragma unroll
                for (int t = 0; t < NT; t++) s_tot[t] = w[t];
            }
            __syncthreads();
            if (DIST && threadIdx.x < 32) pcg_allreduce_warp(*a.p2p, a.epoch, s_tot, NT, sync);
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int t = 0; t < NT; t++) sync->result[gen & 1u][t] = s_tot[t];
            sync->arrive = 0u;
            st_release_u32(&sync->release, gen + 1u);
        }
    } else if (threadIdx.x == 0) {
        spin_u32(&sync->release, gen + 1u, sync);
    }
    if (threadIdx.x == 0) __threadfence();          // acquire side: also drops this SM's L1 lines (CCTL.IVALL)
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NT; t++) v[t] = __ldcg(&sync->result[gen & 1u][t]);
    gen++;
    return ld_relaxed_u32(&sync->abort) == 0u;
}

// y = A x over the slices [s_lo, s_hi) of this rank's owned rows, interior slices first; returns this thread's share of x.y over
// the owned rows.  DIST: a warp that reaches a slice touching a boundary plane first waits for that neighbour's plane (epoch).
template <class IDX, int NB, bool DIST, bool CS, bool DOT>
__device__ __noinline__ double pcg_product(const PcgArgs& a, const double* x, unsigned long long halo_epoch) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int s_lo = a.own_lo / kSellC, s_hi = (a.own_hi + kSellC - 1) / kSellC;
    // boundary slices: those holding rows of the first / last owned node plane (the send ranges; their columns reach the ghosts)
    int s_lb = s_lo, s_rb = s_hi;
    if (DIST) {
        if (a.cntL > 0) s_lb = min(s_hi, (a.sendL + a.cntL + kSellC - 1) / kSellC);
        if (a.cntR > 0) s_rb = max(s_lb, a.sendR / kSellC);
    }
    const int n_int = s_rb - s_lb, n_left = s_lb - s_lo, n_all = s_hi - s_lo;
    bool waitedL = false, waitedR = false;
    double dot = 0.0;
    for (int v = warp; v < n_all; v += nwarps) {
        int s;
        if (v < n_int) s = s_lb + v;
        else {
            const int w = v - n_int;
            const bool left = w < n_left;
            s = left ? s_lo + w : s_rb + (w - n_left);
            if (DIST) {
                bool& waited = left ? waitedL : waitedR;
                if (!waited) {
                    const unsigned long long* mine = a.p2p->halo_flags[a.p2p->rank];
                    if (lane == 0) { spin_sys_u64(mine + (left ? 0 : 1), halo_epoch, a.sync); __threadfence_system(); }
                    __syncwarp();
                    waited = true;
                    // a slice can hold rows of both planes when the slab is a single plane thick: wait for both sides
                    if (n_int == 0) {
                        if (lane == 0) {
                            if (a.cntL > 0) spin_sys_u64(mine + 0, halo_epoch, a.sync);
                            if (a.cntR > 0) spin_sys_u64(mine + 1, halo_epoch, a.sync);
                            __threadfence_system();
                        }
                        __syncwarp();
                        waitedL = waitedR = true;
                    }
                }
            }
        }
        const long long base = a.slice_ptr[s];
        const int width = (int)((a.slice_ptr[s + 1] - base) / kSellC);
        const int r = (s * kSellC + lane < a.rows) ? s * kSellC + lane : -1;
        const double acc = sell_slice_acc<IDX, NB, 6, CS, false>((const IDX*)a.sell_idx, a.sell_val, x, base, width, lane, r);
        if (r >= a.own_lo && r < a.own_hi) {
            a.y[r] = acc;
            if (DOT) dot += acc * x[r];
        }
    }
    return dot;
}

// Push the boundary planes of p into the neighbours' ghost ranges and publish the halo epoch.  CTAs [0, nL) serve the left
// plane, [nL, nL + nR) the right one; the last CTA of a side to finish raises the flag at that neighbour.
// UPDATE: p = beta p + z on those rows first (iteration); otherwise p already holds the values (set-up).
template <bool UPDATE>
__device__ __noinline__ void pcg_push_halo(const PcgArgs& a, const double* zsrc, double beta, unsigned long long epoch) {
    const P2PView& P = *a.p2p;
    const int per = 2 * kThreads;
    int nL = a.cntL > 0 ? min((a.cntL + per - 1) / per, max(1, (int)gridDim.x / 4)) : 0;
    int nR = a.cntR > 0 ? min((a.cntR + per - 1) / per, max(1, (int)gridDim.x / 4)) : 0;
    if (nL + nR > (int)gridDim.x) { nL = a.cntL > 0 ? 1 : 0; nR = 0; }      // tiny grids: CTA 0 serves both sides in turn
    const bool tiny = (a.cntR > 0 && nR == 0);
    for (int side = 0; side < 2; side++) {
        const int cnt = side == 0 ? a.cntL : a.cntR;
        if (cnt <= 0) continue;
        const int first = side == 0 ? 0 : (tiny ? 0 : nL), ncta = side == 0 ? nL : (tiny ? 1 : nR);
        const int me = (int)blockIdx.x - first;
        if (me < 0 || me >= ncta) continue;
        const int send = side == 0 ? a.sendL : a.sendR;
        double* dst = side == 0 ? P.left_p + P.left_recv_off : P.right_p + P.right_recv_off;
        for (int j = me * kThreads + threadIdx.x; j < cnt; j += ncta * kThreads) {
            const int i = send + j;
            double v = a.p[i];
            if (UPDATE) { v = beta * v + zsrc[i]; a.p[i] = v; }
            dst[j] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (atomicAdd(&a.sync->halo_arrive[side], 1u) == (unsigned int)ncta - 1u) {
                a.sync->halo_arrive[side] = 0u;
                __threadfence_system();
                // I am the RIGHT neighbour of rank-1 (slot 1 there) and the LEFT neighbour of rank+1 (slot 0 there)
                unsigned long long* flag = side == 0 ? P.halo_flags[P.rank - 1] + 1 : P.halo_flags[P.rank + 1] + 0;
                *(volatile unsigned long long*)flag = epoch;
            }
        }
    }
}

// MODE 0: CG (z = r), 1: ScalingCG (z = r / diag).  Launch cooperatively with gridDim.x <= resident CTAs.
template <class IDX, int NB, int MODE, bool DIST, bool CS>
__global__ void __launch_bounds__(kThreads, PF2_PCG_MINB)
pcg_persistent_kernel(const __grid_constant__ PcgArgs a) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lo = a.own_lo, hi = a.own_hi;
    double* const zv = (MODE == 1) ? a.z : a.r;
    unsigned int gen = 0;
    unsigned long long halo_epoch = 0;
    if (DIST) halo_epoch = *(volatile unsigned long long*)(a.epoch + 1);
    const bool timing = (blockIdx.x == 0 && threadIdx.x == 0);
    unsigned long long t_acc[3] = { 0ull, 0ull, 0ull }, t_prev = 0ull;
    bool ok = true;

    // ---- set-up (CG.h:422-428) ---------------------------------------------------------------------------------------
    if (a.warm) {
        // r = b - A x0: the ghost entries of x0 are valid (they were exchanged with the previous solution)
        pcg_product<IDX, NB, false, CS, false>(a, a.x, 0ull);
        ok = grid_reduce_bcast<0, DIST>(nullptr, a, gen);
    }
    double v[3] = { 0.0, 0.0, 0.0 };
    for (int i = tid; i < a.rows; i += nth) {
        if (i < lo || i >= hi) {            // ghost rows: p arrives by halo exchange, the rest is never read
            if (!a.warm) a.x[i] = 0.0;
            a.p[i] = 0.0;
            continue;
        }
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
        v[0] += bi * bi; v[1] += zi * ri; v[2] += ri * ri;
    }
    ok = grid_reduce_bcast<3, DIST>(v, a, gen) && ok;
    const double bb = v[0];
    double rho = v[1], rr = v[2], beta = 0.0;
    int iter = 0;
    // a warm start may already satisfy the stopping rule; from x0 = 0 the reference always iterates (b = 0 never converges)
    bool done = a.warm && (sqrt(rr) < a.eps * sqrt(bb));
    if (DIST && ok && !done) {
        halo_epoch++;
        pcg_push_halo<false>(a, nullptr, 0.0, halo_epoch);
        ok = grid_reduce_bcast<0, DIST>(nullptr, a, gen);
    }
    if (timing) t_prev = global_ns();

    // ---- iterations (CG.h:430-449) -------------------------------------------------------------------------------------
    while (ok && !done && iter < a.itrmax) {
        double d1[1];
        d1[0] = pcg_product<IDX, NB, DIST, CS, true>(a, a.p, halo_epoch);
        ok = grid_reduce_bcast<1, DIST>(d1, a, gen);
        if (!ok) break;
        if (timing) { const unsigned long long t = global_ns(); t_acc[0] += t - t_prev; t_prev = t; }
        const double alpha = rho / d1[0];
        double w[2] = { 0.0, 0.0 };
        for (int i = lo + tid; i < hi; i += nth) {
            a.x[i] = a.x[i] + alpha * a.p[i];
            const double ri = a.r[i] + (-alpha) * a.y[i];
            a.r[i] = ri;
            w[1] += ri * ri;
            if (MODE == 0) w[0] += ri * ri;
            else { const double zi = ri / a.dvec[i]; a.z[i] = zi; w[0] += zi * ri; }
        }
        ok = grid_reduce_bcast<2, DIST>(w, a, gen);
        if (!ok) break;
        if (timing) { const unsigned long long t = global_ns(); t_acc[1] += t - t_prev; t_prev = t; }
        beta = w[0] / rho;
        rho = w[0];
        rr = w[1];
        iter++;
        if (sqrt(rr) < a.eps * sqrt(bb)) { done = true; break; }
        if (iter >= a.itrmax) break;
        if (DIST) { halo_epoch++; pcg_push_halo<true>(a, zv, beta, halo_epoch); }
        for (int i = lo + tid; i < hi; i += nth) {
            if (DIST && ((i >= a.sendL && i < a.sendL + a.cntL) || (i >= a.sendR && i < a.sendR + a.cntR))) continue;   // pushed above
            a.p[i] = beta * a.p[i] + zv[i];
        }
        ok = grid_reduce_bcast<0, DIST>(nullptr, a, gen);
        if (timing) { const unsigned long long t = global_ns(); t_acc[2] += t - t_prev; t_prev = t; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CgState* st = a.st;
        st->bb = bb; st->rr = rr; st->rho = rho; st->beta = beta; st->pAp = 0.0;
        st->iter = iter; st->done = ok ? (done ? 1 : 0) : 2; st->maxit = a.itrmax; st->eps = a.eps;
        if (DIST) *(volatile unsigned long long*)(a.epoch + 1) = halo_epoch;
        a.sync->t_ns[0] = t_acc[0]; a.sync->t_ns[1] = t_acc[1]; a.sync->t_ns[2] = t_acc[2]; a.sync->t_ns[3] = (unsigned long long)iter;
    }
}

}  // namespace pf2
