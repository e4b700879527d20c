// p2p.cuh -- peer-memory collectives usable from inside any kernel (NVLink / NVSwitch, IPC-mapped arenas; see dist.cu).
#pragma once
#include "types.cuh"

namespace pf2 {

constexpr long long kP2pSpinLimitClk = 8000000000ll;      // ~4 s of SM clocks: a peer that left the solve becomes PF2_E_CUDA, not a hung box
constexpr int kLlTerms = 4;                               // fp64 values per rank and exchange

__device__ __forceinline__ unsigned long long p2p_ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---- LL words (the low-latency protocol of NCCL): a double travels as two 8-byte words {32 bits of the value, 32-bit flag}.  An
// 8-byte store is single-copy atomic on every path (L2, NVLink), so a reader that sees the flag sees the data: no fence and no
// separate flag write, i.e. one NVLink crossing per exchange instead of store + fence + flag.
template <bool SYS>
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned int flag) {
    const unsigned long long f = (unsigned long long)flag << 32;
    const unsigned long long w0 = f | (unsigned int)__double2loint(v), w1 = f | (unsigned int)__double2hiint(v);
    if (SYS) asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
    else asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
template <bool SYS>
__device__ __forceinline__ bool ll_try_load(const unsigned long long* slot, unsigned int flag, double& v) {
    unsigned long long w0, w1;
    if (SYS) asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    if ((unsigned int)(w0 >> 32) != flag || (unsigned int)(w1 >> 32) != flag) return false;
    v = __hiloint2double((int)(unsigned int)w1, (int)(unsigned int)w0);
    return true;
}

// bounded wait on a flag word a PEER raises (halo epochs); false = timed out (the abort word of the view is set)
__device__ __forceinline__ bool p2p_wait_flag(const P2PView& P, const unsigned long long* flag, unsigned long long target) {
    if (p2p_ld_sys_u64(flag) >= target) return true;
    long long t0 = 0;
    for (unsigned int n = 1;; n++) {
        if (p2p_ld_sys_u64(flag) >= target) return true;
        if ((n & 1023u) == 0u) {
            if (t0 == 0) t0 = clock64();
            if (*(volatile unsigned int*)P.abort) return false;
            if (clock64() - t0 > kP2pSpinLimitClk) { atomicExch(P.abort, 1u); return false; }
        }
    }
}

// Allreduce (sum) of <= 4 fp64 executed by ONE WARP: lane r sends this rank's values to rank r as LL words and receives rank r's;
// lane 0 adds them in rank order, so the result is bitwise identical on every rank.  Words are double-buffered by epoch parity: a
// rank cannot run two epochs ahead of a peer because finishing an epoch needs that peer's words.  `vals` must be visible to the
// whole warp (shared memory).  Spins are bounded: on a timeout the view's abort word is set and the values are meaningless.
__device__ __forceinline__ void p2p_allreduce_warp(const P2PView& P, unsigned long long* epoch_ctr, double* vals, int count) {
    const int lane = threadIdx.x & 31;
    const unsigned long long epoch = *(volatile unsigned long long*)epoch_ctr + 1;
    const unsigned int flag = (unsigned int)epoch;
    const int par = (int)(epoch & 1ull);
    double got[kLlTerms] = { 0.0, 0.0, 0.0, 0.0 };
    if (lane < P.world) {
        for (int c = 0; c < count; c++) ll_store<true>(P.ll[lane] + (((size_t)par * P.world + P.rank) * kLlTerms + c) * 2, vals[c], flag);
        for (int c = 0; c < count; c++) {
            const unsigned long long* slot = P.ll[P.rank] + (((size_t)par * P.world + lane) * kLlTerms + c) * 2;
            long long t0 = 0;
            for (unsigned int n = 1; !ll_try_load<true>(slot, flag, got[c]); n++) {
                if ((n & 1023u) == 0u) {
                    if (t0 == 0) t0 = clock64();
                    if (*(volatile unsigned int*)P.abort) break;
                    if (clock64() - t0 > kP2pSpinLimitClk) { atomicExch(P.abort, 1u); break; }
                }
            }
        }
    }
    __syncwarp();
    for (int c = 0; c < count; c++) {
        double sum = 0.0;
        for (int r = 0; r < P.world; r++) sum += __shfl_sync(0xffffffffu, got[c], r);
        if (lane == 0) vals[c] = sum;
    }
    if (lane == 0) *(volatile unsigned long long*)epoch_ctr = epoch;
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
        CgState* st = p2p->cg1;
        if (st == nullptr) {
            if (threadIdx.x == 0) p2p_vals[0] = local_sum;
            __syncwarp();
            p2p_allreduce_warp(*p2p, epoch_ctr, p2p_vals, 1);
            if (threadIdx.x == 0) *dot_out = p2p_vals[0];
            return;
        }
        // single-reduction CG (dist.cu: solve_dist_cg1): w.u joins the {u.r, r.r} partials of the preceding update kernel ({b.b, u.r, r.r} of
        // the set-up on the first product) in ONE cross-GPU sum, and the scalar tail runs here
        const int first = st->pad;
        if (threadIdx.x == 0) { p2p_vals[0] = local_sum; p2p_vals[1] = st->red[0]; p2p_vals[2] = st->red[1]; p2p_vals[3] = st->red[2]; }
        __syncwarp();
        p2p_allreduce_warp(*p2p, epoch_ctr, p2p_vals, first ? 4 : 3);
        if (threadIdx.x == 0) {
            const double delta = p2p_vals[0];
            double gamma, rr, beta, alpha;
            if (first) {
                st->bb = p2p_vals[1]; gamma = p2p_vals[2]; rr = p2p_vals[3];
                beta = 0.0; alpha = gamma / delta;
                st->pad = 0;
            } else {
                gamma = p2p_vals[1]; rr = p2p_vals[2];
                beta = gamma / st->rho;
                alpha = gamma / (delta - beta * gamma / st->zr_new);
                st->iter = st->iter + 1;
            }
            st->rho = gamma; st->beta = beta; st->zr_new = alpha; st->rr = rr; st->pAp = delta;
            if (sqrt(rr) < st->eps * sqrt(st->bb)) st->done = 1;
        }
    }
}

}  // namespace pf2
