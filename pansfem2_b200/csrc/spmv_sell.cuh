// spmv_sell.cuh -- sliced-ELL (SELL-32) mirror of the CSR matrix and its SpMV kernel.
//
// Why: the sub-warp CSR kernel is limited by the L1/LSU, not by DRAM (ncu: l1tex 64 % busy, long-scoreboard stalls at
// full occupancy, DRAM 49 %; profiles/r01_ncu_full_cg_kernels_vector8.txt): 8 lanes share one row, so a warp's value /
// index loads hit 4 separate row segments and its x gathers ~12 sectors.  In SELL-32 a slice is 32 consecutive rows
// stored column-major: lane l owns row l of the slice, the k-th value / index loads of a warp are one contiguous
// 256 B / 128 B line, and because consecutive FEM rows have consecutive columns the k-th gather of a warp spans
// ~256 contiguous bytes of x.  No shuffles, 18 independent loads per lane (unrolled) keep > 100 KB per SM in flight.
// Padding (rows shorter than the slice's longest) is zero-valued with the row's own column, so it is harmless.
// CSR stays the canonical storage (assembly, download, ILU0): sell_val is refreshed from CSR data before a solve
// (one 16 B/nnz pass, < 0.1 % of a solve).  Algorithmic bytes stay 12*nnz + 24*rows; the stored entries are
// sell_entries >= nnz (ratio reported as `sell_fill`).
// Ragged matrices (T6 / Q8 / hex20 meshes: corner and mid-side nodes have different row lengths) use SELL-C-sigma: inside
// windows of kSellSigma rows the rows are sorted by length before being cut into slices, `sell_perm[slot]` names the row a lane
// owns (-1 = padding lane).  Column deltas stay relative to the lane's own row, so the 2-byte index stream survives.
#pragma once
#include "types.cuh"
#include "p2p.cuh"

namespace pf2 {

constexpr int kSellC = 32;
constexpr int kSellSigma = 1024;     // sorting window (rows): small enough that a slice's rows stay neighbours in the mesh

// row owned by slot `slot` (slice * 32 + lane): natural order, or the length-sorted permutation; -1 = padding lane
__device__ __forceinline__ int sell_row(const int* __restrict__ perm, int rows, int slot) {
    if (perm) return perm[slot];
    return slot < rows ? slot : -1;
}
// sort key of row r: window | (65535 - length) | position in window  -> descending length inside each window, stable
static __global__ void sell_sort_keys_kernel(int rows, const long long* __restrict__ indptr, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const unsigned long long len = (unsigned long long)min((long long)65535, indptr[r + 1] - indptr[r]);
        keys[r] = ((unsigned long long)(r / kSellSigma) << 32) | ((65535ull - len) << 12) | (unsigned long long)(r % kSellSigma);
        vals[r] = r;
    }
}
static __global__ void sell_perm_tail_kernel(int rows, int slots, int* __restrict__ perm) {
    for (int i = rows + blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += gridDim.x * blockDim.x) perm[i] = -1;
}

static __global__ void sell_slice_len_kernel(int rows, int nslices, const long long* __restrict__ indptr, const int* __restrict__ perm, long long* __restrict__ slice_ptr) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nslices; s += gridDim.x * blockDim.x) {
        int m = 0;
        for (int l = 0; l < kSellC; l++) { const int r = sell_row(perm, rows, s * kSellC + l); if (r >= 0) m = max(m, (int)(indptr[r + 1] - indptr[r])); }
        slice_ptr[s + 1] = (long long)m * kSellC;
        if (s == 0) slice_ptr[0] = 0;
    }
}

// fill indices (pattern) and the CSR->SELL position of every stored entry
static __global__ void sell_fill_kernel(int rows, int slots, const long long* __restrict__ indptr, const int* __restrict__ indices, const int* __restrict__ perm,
                                 const long long* __restrict__ slice_ptr, int* __restrict__ sell_idx, int* max_delta) {
    int md = 0;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < slots; slot += gridDim.x * blockDim.x) {
        const int r = sell_row(perm, rows, slot);
        if (r < 0) continue;
        const int s = slot / kSellC, l = slot % kSellC;
        const long long base = slice_ptr[s];
        const int width = (int)((slice_ptr[s + 1] - base) / kSellC);
        const long long b = indptr[r];
        const int len = (int)(indptr[r + 1] - b);
        for (int k = 0; k < width; k++) {
            const int c = (k < len) ? indices[b + k] : r;
            sell_idx[base + (long long)k * kSellC + l] = c;
            md = max(md, abs(c - r));
        }
    }
    for (int o = 16; o > 0; o >>= 1) md = max(md, __shfl_xor_sync(0xffffffffu, md, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_delta, md);
}
// 16-bit column deltas (col - row): FEM matrices are banded, so the index stream shrinks from 4 to 2 bytes per nonzero
static __global__ void sell_delta16_kernel(int rows, long long entries, const long long* __restrict__ slice_ptr, const int* __restrict__ perm,
                                    const int* __restrict__ sell_idx, short* __restrict__ sell_d16) {
    const int nslices = (rows + kSellC - 1) / kSellC;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nslices; s += nwarps) {
        const long long base = slice_ptr[s];
        const int width = (int)((slice_ptr[s + 1] - base) / kSellC);
        const int r = sell_row(perm, rows, s * kSellC + lane);
        for (int k = 0; k < width; k++) {
            const long long pos = base + (long long)k * kSellC + lane;
            sell_d16[pos] = (r >= 0) ? (short)(sell_idx[pos] - r) : (short)0;
        }
    }
}
// ---- block deltas: rows of an NB-dof-per-node FEM matrix hold their columns in runs of NB consecutive indices (the dofs of one
// neighbouring node), so ONE 16-bit delta per run is enough: the index stream shrinks from 2 to 2/NB bytes per nonzero.
// Requires every row to consist of complete runs (true unless Dirichlet conditions fix only some dofs of a node).
static __global__ void sell_block_check_kernel(int rows, int nb, const long long* __restrict__ indptr, const int* __restrict__ indices, int* bad) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const long long b = indptr[r];
        const int len = (int)(indptr[r + 1] - b);
        bool ok = (len % nb) == 0;
        for (int k = 0; ok && k < len; k += nb) {
            ok = (indices[b + k] % nb) == 0;                     // runs start on a node boundary: deltas are exact in node units
            for (int j = 1; j < nb; j++) ok = ok && (indices[b + k + j] == indices[b + k] + j);
        }
        if (!ok) atomicExch(bad, 1);
    }
}
// bidx[(slice_base / nb) + kb * 32 + lane] = (first column of run kb - first row of the lane's node) / nb : deltas in NODE units, so
// they fit 16 bits up to 32 767 nodes of bandwidth (hex8 meshes with planes of up to ~10 900 nodes); wider meshes store them as int32
template <class OUT>
static __global__ void sell_block_index_kernel(int rows, int nb, const long long* __restrict__ slice_ptr, const int* __restrict__ perm,
                                        const int* __restrict__ sell_idx, OUT* __restrict__ bidx) {
    const int nslices = (rows + kSellC - 1) / kSellC;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nslices; s += nwarps) {
        const long long base = slice_ptr[s];
        const int width = (int)((slice_ptr[s + 1] - base) / kSellC);
        const int r = sell_row(perm, rows, s * kSellC + lane);
        const int rbase = (r >= 0) ? r - r % nb : 0;
        for (int kb = 0; kb < width / nb; kb++) {
            // padding entries carry column r, i.e. node delta 0: a run inside the matrix with zero values
            const int d = (r >= 0) ? (sell_idx[base + (long long)(kb * nb) * kSellC + lane] - rbase) / nb : 0;
            bidx[base / nb + (long long)kb * kSellC + lane] = (OUT)d;
        }
    }
}
// padding lanes of the last slice (slots beyond `rows`; with a permutation they are the last slots too)
static __global__ void sell_pad_tail_kernel(int rows, int nslices, const long long* __restrict__ slice_ptr, int* __restrict__ sell_idx, double* __restrict__ sell_val) {
    const int s = nslices - 1;
    const long long base = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - base) / kSellC);
    for (int t = threadIdx.x; t < width * kSellC; t += blockDim.x) {
        const int l = t % kSellC;
        if (s * kSellC + l >= rows) { if (sell_idx) sell_idx[base + t] = 0; sell_val[base + t] = 0.0; }
    }
}
static __global__ void sell_values_kernel(int rows, const long long* __restrict__ indptr, const double* __restrict__ data, const int* __restrict__ perm,
                                   const long long* __restrict__ slice_ptr, double* __restrict__ sell_val) {
    // one warp per slice: lane l copies row l; reads are strided (row-contiguous), writes coalesced
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nslices = (rows + kSellC - 1) / kSellC;
    for (int s = warp; s < nslices; s += nwarps) {
        const int r = sell_row(perm, rows, s * kSellC + lane);
        const long long base = slice_ptr[s];
        const int width = (int)((slice_ptr[s + 1] - base) / kSellC);
        long long b = 0;
        int len = 0;
        if (r >= 0) { b = indptr[r]; len = (int)(indptr[r + 1] - b); }
        for (int k = 0; k < width; k++) sell_val[base + (long long)k * kSellC + lane] = (k < len) ? data[b + k] : 0.0;
    }
}

// loads of the matrix stream: evict-first (the matrix is far larger than L2) or plain (slabs small enough to stay L2-resident)
template <bool CS, class T> __device__ __forceinline__ T sell_ld(const T* p) { return CS ? __ldcs(p) : __ldg(p); }

// One lane's share of one slice: acc = sum_k val[k] * x[col[k]] for row `r` (lane `lane` of the slice starting at `base`, `width`
// stored entries per row).  IDX = int: absolute columns (4 B/nnz); IDX = short: column - row deltas (2 B/nnz); NB > 1: one delta per
// run of NB consecutive columns, in node units.  U independent value / index loads are in flight per round.
// NC = false: x is written inside the same kernel by other CTAs (persistent PCG): plain coherent loads, never the read-only path.
template <bool NC> __device__ __forceinline__ double sell_ldx(const double* x, int i) { return NC ? __ldg(x + i) : x[i]; }

// PF = true (scalar deltas only): the NEXT round's U index loads are issued before this round's gathers, so that a round pays one
// dependent memory latency (values + gathers together) instead of two (index, then gather) -- same sums in the same order.
template <class IDX, int NB, int U, bool CS, bool NC = true, bool PF = false>
__device__ __forceinline__ double sell_slice_acc(const IDX* __restrict__ sell_idx, const double* __restrict__ sell_val, const double* x,
                                                 long long base, int width, int lane, int r) {
    const double* v = sell_val + base + lane;
    const int off = (NB > 1) ? max(r, 0) - max(r, 0) % NB                // block deltas: node units relative to the first row of the lane's node
                             : ((sizeof(IDX) == 2) ? max(r, 0) : 0);     // per-entry deltas are relative to the lane's row (padding lanes: delta 0, value 0)
    double acc = 0.0;
    if constexpr (NB > 1) {
        // block deltas: one index per run of NB consecutive columns; the index stream of this slice starts at base / NB
        const IDX* cb = sell_idx + base / NB + lane;
        const int nblk = width / NB;
        constexpr int UB = (NB == 2) ? 3 : 2;              // 6 values in flight per step, like the scalar path
        int kb = 0;
        for (; kb + UB <= nblk; kb += UB) {
            double vv[UB * NB];
            int cc[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                cc[u] = off + NB * (int)sell_ld<CS>(cb + (kb + u) * kSellC);
#pragma unroll
                for (int j = 0; j < NB; j++) vv[u * NB + j] = sell_ld<CS>(v + ((kb + u) * NB + j) * kSellC);
            }
#pragma unroll
            for (int u = 0; u < UB; u++)
#pragma unroll
                for (int j = 0; j < NB; j++) acc += vv[u * NB + j] * sell_ldx<NC>(x, cc[u] + j);
        }
        for (; kb < nblk; kb++) {
            const int c0 = off + NB * (int)sell_ld<CS>(cb + kb * kSellC);
#pragma unroll
            for (int j = 0; j < NB; j++) acc += sell_ld<CS>(v + (kb * NB + j) * kSellC) * sell_ldx<NC>(x, c0 + j);
        }
    } else {
        const IDX* c = sell_idx + base + lane;
        int k = 0;
        if constexpr (PF) {
            int cn[U];
            if (U <= width) {
#pragma unroll
                for (int u = 0; u < U; u++) cn[u] = off + (int)sell_ld<CS>(c + u * kSellC);
            }
            for (; k + U <= width; k += U) {
                double vv[U];
                int cc[U];
#pragma unroll
                for (int u = 0; u < U; u++) { vv[u] = sell_ld<CS>(v + (k + u) * kSellC); cc[u] = cn[u]; }
                if (k + 2 * U <= width) {
#pragma unroll
                    for (int u = 0; u < U; u++) cn[u] = off + (int)sell_ld<CS>(c + (k + U + u) * kSellC);
                }
#pragma unroll
                for (int u = 0; u < U; u++) acc += vv[u] * sell_ldx<NC>(x, cc[u]);
            }
        }
        for (; k + U <= width; k += U) {
            double vv[U];
            int cc[U];
#pragma unroll
            for (int u = 0; u < U; u++) { vv[u] = sell_ld<CS>(v + (k + u) * kSellC); cc[u] = off + (int)sell_ld<CS>(c + (k + u) * kSellC); }
#pragma unroll
            for (int u = 0; u < U; u++) acc += vv[u] * sell_ldx<NC>(x, cc[u]);
        }
        for (; k < width; k++) acc += sell_ld<CS>(v + k * kSellC) * sell_ldx<NC>(x, off + (int)sell_ld<CS>(c + k * kSellC));
    }
    return acc;
}

// CS = false: plain (L2-allocating) loads of the matrix stream, for slabs small enough to stay L2-resident between products.
// Partitioned matrix with the peer-memory backend (p2p != nullptr): only the slices of the owned rows are multiplied, interior slices
// first; a warp that reaches a slice holding rows of a boundary plane first waits for that neighbour's plane of the current halo
// epoch (deferred there by the p-update kernel), so the exchange overlaps the interior rows.
// (ORDERED is a separate instantiation: the plain kernel must keep its 32 registers -- 8 resident CTAs per SM; with the ordering logic
// compiled in it needed 40 and the single-GPU product slowed down by 12 %.)
template <bool DOT, class IDX, bool PERM = false, int NB = 1, int U = 6, bool CS = true, bool ORDERED = false, bool PF = false>
__global__ void __launch_bounds__(kThreads)
spmv_sell_kernel(int rows, const long long* __restrict__ slice_ptr, const int* __restrict__ perm, const IDX* __restrict__ sell_idx, const double* __restrict__ sell_val,
                 const double* __restrict__ x, double* __restrict__ y, const CgState* __restrict__ st, double* dot_out,
                 double* partials, unsigned int* ticket, int dot_lo, int dot_hi, const P2PView* p2p, unsigned long long* p2p_epoch) {
    if (DOT && st != nullptr && st->done) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nslices = (rows + kSellC - 1) / kSellC;
    double dot = 0.0;
    // virtual slice order: [s_lb, s_rb) interior, then [s_lo, s_lb) left boundary, then [s_rb, s_hi) right boundary
    int s_lo = 0, s_hi = nslices, s_lb = 0, s_rb = nslices;
    constexpr bool ordered = ORDERED;
    unsigned long long halo_epoch = 0ull;
    if (ORDERED) {
        s_lo = p2p->own_lo / kSellC; s_hi = (p2p->own_hi + kSellC - 1) / kSellC;
        s_lb = s_lo; s_rb = s_hi;
        if (p2p->cntL > 0) s_lb = min(s_hi, (p2p->sendL + p2p->cntL + kSellC - 1) / kSellC);
        if (p2p->cntR > 0) s_rb = max(s_lb, p2p->sendR / kSellC);
        halo_epoch = *(volatile unsigned long long*)(p2p->epoch + 1);
    }
    const int n_int = s_rb - s_lb, n_left = s_lb - s_lo, n_all = s_hi - s_lo;
    bool waitedL = false, waitedR = false;
    for (int v = warp; v < n_all; v += nwarps) {
        int s = s_lo + v;
        if constexpr (ORDERED) {
            if (v < n_int) s = s_lb + v;
            else {
                const int w = v - n_int;
                const bool left = w < n_left;
                s = left ? s_lo + w : s_rb + (w - n_left);
                const bool needL = (left || n_int == 0) && p2p->cntL > 0 && !waitedL;
                const bool needR = (!left || n_int == 0) && p2p->cntR > 0 && !waitedR;
                if (needL || needR) {
                    if (lane == 0) {
                        const unsigned long long* mine = p2p->halo_flags[p2p->rank];
                        if (needL) p2p_wait_flag(*p2p, mine + 0, halo_epoch);
                        if (needR) p2p_wait_flag(*p2p, mine + 1, halo_epoch);
                        __threadfence_system();      // acquire: the plane the neighbour stored before raising its flag
                    }
                    __syncwarp();
                    waitedL = waitedL || needL; waitedR = waitedR || needR;
                }
            }
        }
        const long long base = slice_ptr[s];
        const int width = (int)((slice_ptr[s + 1] - base) / kSellC);
        const int r = PERM ? perm[s * kSellC + lane] : ((s * kSellC + lane < rows) ? s * kSellC + lane : -1);
        // ghost entries of x arrive from a peer GPU while this kernel runs: they must not come through the non-coherent path
        const double acc = ordered ? sell_slice_acc<IDX, NB, U, CS, false>(sell_idx, sell_val, x, base, width, lane, r)
                                   : sell_slice_acc<IDX, NB, U, CS, true, PF>(sell_idx, sell_val, x, base, width, lane, r);
        if (r >= 0) {
            y[r] = acc;
            if (DOT && r >= dot_lo && r < dot_hi) dot += acc * x[r];
        }
    }
    if (DOT) {
        double vsum[1] = { dot };
        if (grid_sum_last<1>(vsum, partials, ticket)) finish_dot(vsum[0], dot_out, p2p, p2p_epoch);
    }
}

}  // namespace pf2
