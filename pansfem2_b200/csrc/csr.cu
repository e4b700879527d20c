// csr.cu -- CSR<T> on the device: upload/download and y = A*x (replaces CSR<T>::operator*, CSR.h:109-122).
//
// HBM layout: indptr int64[rows+1], indices int32[nnz], data fp64[nnz] -- the reference's three arrays, contiguous.
// Algorithmic bytes per SpMV: 12*nnz (value + column index) + 24*rows (int64 row pointer, x read once, y written once).
//
// Two kernels, picked per matrix by plan_spmv():
//   * stream  : a CTA owns `stream_rows` consecutive rows, streams their nonzeros (values, indices) with fully
//               coalesced 128-byte loads, multiplies by the gathered x and parks the products in shared memory;
//               sub-groups of threads then fold each row from shared memory.  Row length never affects coalescing, which
//               is what short FEM rows (9/18 nnz) need to get near the HBM roofline.
//   * vector  : TPR lanes per row (2..32) with a shuffle reduction; used when a CTA's rows do not fit in shared memory.
// Both optionally fuse the dot product x.y (p.Ap of CG, CG.h:433) into the epilogue.
#include "types.cuh"
#include "p2p.cuh"
#include "spmv_tma.cuh"
#include "spmv_sell.cuh"
#include "spmv_mf.cuh"
#include <cub/cub.cuh>

namespace pf2 {

constexpr int kStreamCap = 4608;   // products per CTA in shared memory (36 KB): 256 rows x 18 nnz

__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }

// ---- streaming kernel ------------------------------------------------------------------------------------------
// G = threads cooperating on one row in the fold phase; a CTA covers kThreads/G rows per tile.
template <int G, bool DOT>
__global__ void __launch_bounds__(kThreads)
spmv_stream_kernel(int rows, const long long* __restrict__ indptr, const int* __restrict__ indices,
                   const double* __restrict__ data, const double* __restrict__ x, double* __restrict__ y,
                   const CgState* __restrict__ st, double* dot_out, double* partials, unsigned int* ticket, int dot_lo, int dot_hi, const P2PView* p2p, unsigned long long* p2p_epoch) {
    if (DOT && st != nullptr && st->done) return;
    __shared__ double prod[kStreamCap];
    __shared__ long long s_range[2];
    constexpr int RPB = kThreads / G;   // rows per CTA tile
    const int ntiles = (rows + RPB - 1) / RPB;
    double dot = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = tile * RPB;
        const int r1 = min(r0 + RPB, rows);
        if (threadIdx.x == 0) { s_range[0] = indptr[r0]; s_range[1] = indptr[r1]; }
        __syncthreads();
        const long long s = s_range[0];
        const int cnt = (int)(s_range[1] - s);
        // phase 1: coalesced stream of the tile's nonzeros
#pragma unroll 4
        for (int j = threadIdx.x; j < cnt; j += kThreads) {
            const double v = ld_stream(data + s + j);
            const int c = ld_stream(indices + s + j);
            prod[j] = v * __ldg(x + c);
        }
        __syncthreads();
        // phase 2: fold rows out of shared memory
        const int lr = threadIdx.x / G, g = threadIdx.x % G;
        const int row = r0 + lr;
        double acc = 0.0;
        if (row < r1) {
            const int b = (int)(indptr[row] - s), e = (int)(indptr[row + 1] - s);
            for (int j = b + g; j < e; j += G) acc += prod[j];
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, G);
        if (row < r1 && g == 0) {
            y[row] = acc;
            if (DOT && row >= dot_lo && row < dot_hi) dot += acc * x[row];
        }
        __syncthreads();
    }
    if (DOT) {
        double v[1] = { dot };
        if (grid_sum_last<1>(v, partials, ticket)) finish_dot(v[0], dot_out, p2p, p2p_epoch);
    }
}

// ---- vector kernel ---------------------------------------------------------------------------------------------
template <int TPR, bool DOT>
__global__ void __launch_bounds__(kThreads)
spmv_vector_kernel(int rows, const long long* __restrict__ indptr, const int* __restrict__ indices,
                   const double* __restrict__ data, const double* __restrict__ x, double* __restrict__ y,
                   const CgState* __restrict__ st, double* dot_out, double* partials, unsigned int* ticket, int dot_lo, int dot_hi, const P2PView* p2p, unsigned long long* p2p_epoch) {
    if (DOT && st != nullptr && st->done) return;
    constexpr int RPB = kThreads / TPR;
    const int lr = threadIdx.x / TPR, lane = threadIdx.x % TPR;
    double dot = 0.0;
    // the trip count is CTA-uniform so that the sub-warp shuffles below always see all 32 lanes
    for (long long base = (long long)blockIdx.x * RPB; base < rows; base += (long long)gridDim.x * RPB) {
        const long long row = base + lr;
        const bool valid = row < rows;
        double acc = 0.0;
        if (valid) {
            const long long s = indptr[row], e = indptr[row + 1];
#pragma unroll 4
            for (long long j = s + lane; j < e; j += TPR) acc += ld_stream(data + j) * __ldg(x + ld_stream(indices + j));
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, TPR);
        if (valid && lane == 0) {
            y[row] = acc;
            if (DOT && row >= dot_lo && row < dot_hi) dot += acc * x[row];
        }
    }
    if (DOT) {
        double v[1] = { dot };
        if (grid_sum_last<1>(v, partials, ticket)) finish_dot(v[0], dot_out, p2p, p2p_epoch);
    }
}

// row statistics + diagonal offsets (GetDiagonal's A.get(i,i), CG.h:398-404 / CSR.h:155-167, resolved once)
__global__ void csr_rowinfo_kernel(int rows, const long long* __restrict__ indptr, const int* __restrict__ indices,
                                   int* __restrict__ diagpos, int* maxrow) {
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        const long long s = indptr[i], e = indptr[i + 1];
        m = max(m, (int)(e - s));
        long long lo = s, hi = e - 1;
        int pos = -1;
        while (lo <= hi) {
            long long mid = (lo + hi) >> 1;
            int c = indices[mid];
            if (c == i) { pos = (int)(mid - s); break; }
            if (c < i) lo = mid + 1; else hi = mid - 1;
        }
        diagpos[i] = pos;
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxrow, m);
}

__global__ void widen_indptr_kernel(int n, const int* __restrict__ in, long long* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i];
}

int csr_finalize_structure(pf2_csr* A) {
    pf2_ctx* c = A->ctx;
    PF2_TRY(dev_alloc(&A->diagpos, (size_t)A->rows));
    int* d_max = nullptr;
    PF2_TRY(dev_alloc(&d_max, 1));
    PF2_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), c->stream));
    csr_rowinfo_kernel<<<c->grid_for(A->rows), kThreads, 0, c->stream>>>(A->rows, A->indptr, A->indices, A->diagpos, d_max);
    PF2_LAUNCH_CHECK();
    c->launches++;
    PF2_CUDA(cudaMemcpyAsync(&A->max_row, d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    PF2_CUDA(cudaFree(d_max));
    A->spmv_variant = 0;
    return PF2_OK;
}

// variant encoding: 1..5 = vector TPR 2,4,8,16,32 ; 11..15 = stream with G = 1,2,4,8,16 threads per row ;
//                   21..26 = TMA pipeline with G = 1,2,4,8,16,32
static int sell_build(pf2_csr* A);
static void plan_spmv(pf2_csr* A) {
    if (A->spmv_variant) return;
    const double mean = A->rows ? (double)A->nnz / A->rows : 1.0;
    // Measured inside the PCG loop on B200 (tools/cg_sweep.py, profiles/r01_cg_sweep_*.json): the SELL-32 thread-per-row
    // kernel wins whenever padding is small (FEM rows are near-uniform); otherwise the sub-warp CSR kernel with 2-3
    // nonzeros per lane.  The shared-memory stream and TMA-pipeline kernels stay selectable (11-15, 21-26).
    // (thread-per-row slices stop paying once rows are long: hex20 rows average 166 nonzeros and the warp-per-row CSR kernel wins)
    if (A->rows >= 64 && mean <= 96.0 && sell_build(A) == PF2_OK && (double)A->sell_entries <= 1.15 * (double)A->nnz) { A->spmv_variant = 31; return; }
    A->spmv_variant = mean <= 4 ? 1 : mean <= 10 ? 2 : mean <= 24 ? 3 : mean <= 48 ? 4 : 5;
}

template <int G, bool DOT>
static int launch_tma(pf2_csr* A, const double* x, double* y, const CgState* st, double* dot_out) {
    pf2_ctx* c = A->ctx;
    static bool configured = false;
    if (!configured) {
        PF2_CUDA(cudaFuncSetAttribute(spmv_tma_kernel<G, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    const int rpb = kConsumers / G;
    const int cap = tma_cap_for(rpb, A->max_row);
    const int stages = std::max(2, std::min(A->tma_stages, kMaxStages));
    const size_t smem = tma_smem_bytes(cap, stages);
    int per_sm = std::max(1, std::min(A->tma_ctas_per_sm, (int)((220 * 1024) / smem)));
    const int ntiles = (A->rows + rpb - 1) / rpb;
    const int grid = std::max(1, std::min(ntiles, c->sm_count * per_sm));
    spmv_tma_kernel<G, DOT><<<grid, kTmaThreads, smem, c->stream>>>(A->rows, A->indptr, A->indices, A->data, x, y, st, dot_out,
                                                                   c->red.partials, c->red.ticket, cap, stages, A->own_lo, A->own_hi, A->p2p_dev, A->p2p_epoch);
    return PF2_OK;
}

// slice pointers for the current slot->row map; returns the number of stored entries
static int sell_slices(pf2_csr* A, int nslices, long long* entries) {
    pf2_ctx* c = A->ctx;
    sell_slice_len_kernel<<<c->grid_for(nslices), kThreads, 0, c->stream>>>(A->rows, nslices, A->indptr, A->sell_perm, A->sell_ptr);
    PF2_LAUNCH_CHECK();
    void* tmp = nullptr;
    size_t bytes = 0;
    PF2_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, A->sell_ptr, A->sell_ptr, nslices + 1, c->stream));
    PF2_CUDA(cudaMalloc(&tmp, bytes ? bytes : 8));
    PF2_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, A->sell_ptr, A->sell_ptr, nslices + 1, c->stream));
    PF2_CUDA(cudaMemcpyAsync(entries, A->sell_ptr + nslices, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    PF2_CUDA(cudaFree(tmp));
    c->launches += 2;
    return PF2_OK;
}

static int sell_build(pf2_csr* A) {
    if (A->sell_ptr) return PF2_OK;
    pf2_ctx* c = A->ctx;
    const int nslices = (A->rows + kSellC - 1) / kSellC;
    const int slots = nslices * kSellC;
    PF2_TRY(dev_alloc(&A->sell_ptr, (size_t)nslices + 1));
    PF2_TRY(sell_slices(A, nslices, &A->sell_entries));
    if ((double)A->sell_entries > 1.03 * (double)A->nnz) {
        // ragged rows: sort by length inside windows of kSellSigma rows (SELL-C-sigma) and cut the slices from the sorted order
        unsigned long long *k0 = nullptr, *k1 = nullptr;
        int *v0 = nullptr;
        PF2_TRY(dev_alloc(&k0, (size_t)A->rows)); PF2_TRY(dev_alloc(&k1, (size_t)A->rows)); PF2_TRY(dev_alloc(&v0, (size_t)A->rows));
        PF2_TRY(dev_alloc(&A->sell_perm, (size_t)slots));
        sell_sort_keys_kernel<<<c->grid_for(A->rows), kThreads, 0, c->stream>>>(A->rows, A->indptr, k0, v0);
        void* tmp = nullptr;
        size_t bytes = 0;
        PF2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, v0, A->sell_perm, A->rows, 0, 64, c->stream));
        PF2_CUDA(cudaMalloc(&tmp, bytes ? bytes : 8));
        PF2_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, k0, k1, v0, A->sell_perm, A->rows, 0, 64, c->stream));
        sell_perm_tail_kernel<<<1, kThreads, 0, c->stream>>>(A->rows, slots, A->sell_perm);
        PF2_LAUNCH_CHECK();
        PF2_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(tmp); cudaFree(k0); cudaFree(k1); cudaFree(v0);
        c->launches += 3;
        long long sorted_entries = 0;
        PF2_TRY(sell_slices(A, nslices, &sorted_entries));
        A->sell_entries = sorted_entries;
    }
    PF2_TRY(dev_alloc(&A->sell_idx, (size_t)A->sell_entries));
    PF2_TRY(dev_alloc(&A->sell_val, (size_t)A->sell_entries));
    int* d_md = nullptr;
    PF2_TRY(dev_alloc(&d_md, 1));
    PF2_CUDA(cudaMemsetAsync(d_md, 0, sizeof(int), c->stream));
    sell_fill_kernel<<<c->grid_for(slots), kThreads, 0, c->stream>>>(A->rows, slots, A->indptr, A->indices, A->sell_perm, A->sell_ptr, A->sell_idx, d_md);
    PF2_LAUNCH_CHECK();
    int md = 0;
    PF2_CUDA(cudaMemcpyAsync(&md, d_md, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    PF2_CUDA(cudaFree(d_md));
    A->sell_max_delta = md;
    // node-block structure (pattern-built matrices know their dofs per node): one delta per run of NB consecutive columns.
    // Measured on B200: hex8 (runs of 3) SpMV 0.245 -> 0.203 ms at 0.69 M dof; 2-D elasticity (runs of 2) loses 5 % (the kernel is
    // bound by load latency x rounds there, and the rounds do not get shorter), so runs of 2 stay off unless PF2_SELL_BLOCK=1.
    int nb = 1;
    const char* blk = getenv("PF2_SELL_BLOCK");
    const bool want = blk ? (atoi(blk) != 0) : (A->map_ndof == 3);
    if (want && (A->map_ndof == 2 || A->map_ndof == 3) && A->rows > A->map_ndof && A->rows % A->map_ndof == 0) {
        int* d_bad = nullptr;
        PF2_TRY(dev_alloc(&d_bad, 1));
        PF2_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
        sell_block_check_kernel<<<c->grid_for(A->rows), kThreads, 0, c->stream>>>(A->rows, A->map_ndof, A->indptr, A->indices, d_bad);
        int bad = 1;
        PF2_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(d_bad);
        c->launches++;
        if (!bad) nb = A->map_ndof;
    }
    A->sell_nb = nb;
    const int gblk = std::min((nslices + 7) / 8, c->sm_count * 16);
    if (nb > 1) {
        if (md / nb + 1 <= 32767 && getenv("PF2_SELL_BLOCK32") == nullptr) {      // PF2_SELL_BLOCK32: force the wide form (tests)
            PF2_TRY(dev_alloc(&A->sell_d16, (size_t)(A->sell_entries / nb)));
            sell_block_index_kernel<short><<<gblk, kThreads, 0, c->stream>>>(A->rows, nb, A->sell_ptr, A->sell_perm, A->sell_idx, A->sell_d16);
        } else {
            PF2_TRY(dev_alloc(&A->sell_b32, (size_t)(A->sell_entries / nb)));
            sell_block_index_kernel<int><<<gblk, kThreads, 0, c->stream>>>(A->rows, nb, A->sell_ptr, A->sell_perm, A->sell_idx, A->sell_b32);
        }
        PF2_LAUNCH_CHECK();
        PF2_CUDA(cudaStreamSynchronize(c->stream));
        PF2_CUDA(cudaFree(A->sell_idx));
        A->sell_idx = nullptr;
    } else if (md <= 32767) {
        // banded matrix: keep the 2-byte delta stream and drop the 4-byte one (the absolute columns are recoverable)
        PF2_TRY(dev_alloc(&A->sell_d16, (size_t)A->sell_entries));
        sell_delta16_kernel<<<gblk, kThreads, 0, c->stream>>>(A->rows, A->sell_entries, A->sell_ptr, A->sell_perm, A->sell_idx, A->sell_d16);
        PF2_LAUNCH_CHECK();
        PF2_CUDA(cudaStreamSynchronize(c->stream));
        PF2_CUDA(cudaFree(A->sell_idx));
        A->sell_idx = nullptr;
    }
    c->launches += 2;
    A->sell_values_valid = false;
    return PF2_OK;
}

int sell_refresh(pf2_csr* A) {
    PF2_TRY(sell_build(A));
    pf2_ctx* c = A->ctx;
    const int nslices = (A->rows + kSellC - 1) / kSellC;
    sell_values_kernel<<<std::min((nslices + 7) / 8, c->sm_count * 16), kThreads, 0, c->stream>>>(A->rows, A->indptr, A->data, A->sell_perm, A->sell_ptr, A->sell_val);
    sell_pad_tail_kernel<<<1, kThreads, 0, c->stream>>>(A->rows, nslices, A->sell_ptr, A->sell_idx, A->sell_val);   // sell_idx may be null (delta form)
    PF2_LAUNCH_CHECK();
    c->launches += 2;
    A->sell_values_valid = true;
    return PF2_OK;
}

// Opt-in (PF2_SELL_L2_MB=<budget>): a slab whose SELL mirror (values + 16-bit deltas) fits that budget next to the Krylov vectors is read
// with plain loads instead of evict-first ones, to stay L2-resident from one product to the next.  Measured on 8 B200s with a 110 MB slab
// (2-D 4 M dof / 8): 0.0734 vs 0.0685 ms per PCG iteration -- the 126 MB L2 does not hold it next to the vectors, so it is off by default.
static bool sell_l2_resident(const pf2_csr* A) {
    const double mb = getenv("PF2_SELL_L2_MB") ? atof(getenv("PF2_SELL_L2_MB")) : 0.0;      // read per launch: probes switch it
    const double idx_bytes = A->sell_d16 ? 2.0 / A->sell_nb : 4.0 / A->sell_nb;
    return (double)A->sell_entries * (8.0 + idx_bytes) + 40.0 * (double)A->rows <= mb * 1.0e6;
}

static bool sell_prefetch_requested() {
    static const bool on = [] { const char* e = getenv("PF2_SELL_PREFETCH"); return e != nullptr && atoi(e) != 0; }();
    return on;
}

template <bool DOT>
static int launch_sell(pf2_csr* A, const double* x, double* y, const CgState* st, double* dot_out) {
    pf2_ctx* c = A->ctx;
    if (!A->sell_values_valid) PF2_TRY(sell_refresh(A));
    const int nslices = (A->rows + kSellC - 1) / kSellC;
    const int nb = (nslices + (kThreads / 32) - 1) / (kThreads / 32);
    if (DOT && A->p2p_dev && A->p2p_view.defer_halo_wait && A->sell_d16 && !A->sell_perm && (A->sell_nb == 1 || A->sell_nb == 3)) {
        // opt-in (PF2_HALO_DEFER): the boundary slices of this product wait for the neighbours' planes, after the interior slices
#define SELLO(NBV)                                                                                                                         \
    {                                                                                                                                      \
        const int grid = std::max(1, std::min(nb, c->wave_grid((const void*)spmv_sell_kernel<DOT, short, false, NBV, 6, true, true>, kThreads))); \
        spmv_sell_kernel<DOT, short, false, NBV, 6, true, true><<<grid, kThreads, 0, c->stream>>>(A->rows, A->sell_ptr, A->sell_perm, A->sell_d16, A->sell_val, x, y, \
                                                                                               st, dot_out, c->red.partials, c->red.ticket, A->own_lo, \
                                                                                               A->own_hi, A->p2p_dev, A->p2p_epoch);                   \
    }
        if (A->sell_nb == 3) SELLO(3) else SELLO(1)
#undef SELLO
        return PF2_OK;
    }
#define SELL(IDXT, PERMV, NBV, IDXPTR)                                                                                                     \
    {                                                                                                                                      \
        const int grid = std::max(1, std::min(nb, c->wave_grid((const void*)spmv_sell_kernel<DOT, IDXT, PERMV, NBV>, kThreads)));             \
        spmv_sell_kernel<DOT, IDXT, PERMV, NBV><<<grid, kThreads, 0, c->stream>>>(A->rows, A->sell_ptr, A->sell_perm, IDXPTR, A->sell_val, x, y, \
                                                                                  st, dot_out, c->red.partials, c->red.ticket, A->own_lo,      \
                                                                                  A->own_hi, A->p2p_dev, A->p2p_epoch);                        \
    }
    if (A->sell_b32) {         // block deltas too wide for 16 bits
        if (A->sell_nb == 3) { if (A->sell_perm) SELL(int, true, 3, A->sell_b32) else SELL(int, false, 3, A->sell_b32) }
        else { if (A->sell_perm) SELL(int, true, 2, A->sell_b32) else SELL(int, false, 2, A->sell_b32) }
    } else if (A->sell_d16) {
        if (A->sell_nb == 2) { if (A->sell_perm) SELL(short, true, 2, A->sell_d16) else SELL(short, false, 2, A->sell_d16) }
        else if (A->sell_nb == 3) {
            if (A->sell_perm) SELL(short, true, 3, A->sell_d16)
            else if (DOT && sell_l2_resident(A)) {
                const int grid = std::max(1, std::min(nb, c->wave_grid((const void*)spmv_sell_kernel<DOT, short, false, 3, 6, false>, kThreads)));
                spmv_sell_kernel<DOT, short, false, 3, 6, false><<<grid, kThreads, 0, c->stream>>>(A->rows, A->sell_ptr, A->sell_perm, A->sell_d16, A->sell_val, x, y,
                                                                                                st, dot_out, c->red.partials, c->red.ticket, A->own_lo,
                                                                                                A->own_hi, A->p2p_dev, A->p2p_epoch);
            } else SELL(short, false, 3, A->sell_d16)
        }
        else if (A->sell_perm) SELL(short, true, 1, A->sell_d16)
        else {
            // independent loads in flight per lane and round; measured at 2 M dof: 3 -> 0.0843 ms, 6 -> 0.0778, 9 -> 0.0839, 18 -> 0.0895
            // small slabs (<= 2 slices per resident warp: row-partitioned 2-D problems at 8 GPUs) are bound by the rounds of dependent loads, not by
            // occupancy: all 18 nonzeros of a Q4 row in one round measured 0.0658 vs 0.0685 ms per PCG iteration there (profiles/r02_dist_probe.md)
            static const int unroll_env = getenv("PF2_SELL_UNROLL") ? atoi(getenv("PF2_SELL_UNROLL")) : 0;
            const int unroll = unroll_env ? unroll_env : ((long long)nslices <= 2LL * c->sm_count * 8 * (kThreads / 32) ? 18 : 6);
#define SELLU(UV)                                                                                                                          \
    {                                                                                                                                      \
        const int grid = std::max(1, std::min(nb, c->wave_grid((const void*)spmv_sell_kernel<DOT, short, false, 1, UV>, kThreads)));          \
        spmv_sell_kernel<DOT, short, false, 1, UV><<<grid, kThreads, 0, c->stream>>>(A->rows, A->sell_ptr, A->sell_perm, A->sell_d16, A->sell_val, x, y, \
                                                                                     st, dot_out, c->red.partials, c->red.ticket, A->own_lo,     \
                                                                                     A->own_hi, A->p2p_dev, A->p2p_epoch);                       \
    }
            if (DOT && sell_l2_resident(A)) {
                const int grid = std::max(1, std::min(nb, c->wave_grid((const void*)spmv_sell_kernel<DOT, short, false, 1, 6, false>, kThreads)));
                spmv_sell_kernel<DOT, short, false, 1, 6, false><<<grid, kThreads, 0, c->stream>>>(A->rows, A->sell_ptr, A->sell_perm, A->sell_d16, A->sell_val, x, y,
                                                                                                st, dot_out, c->red.partials, c->red.ticket, A->own_lo,
                                                                                                A->own_hi, A->p2p_dev, A->p2p_epoch);
            } else if (unroll == 6 && sell_prefetch_requested()) {
                // opt-in (PF2_SELL_PREFETCH=1): next round's index loads ahead of this round's gathers (spmv_sell.cuh, PF)
                const int grid = std::max(1, std::min(nb, c->wave_grid((const void*)spmv_sell_kernel<DOT, short, false, 1, 6, true, false, true>, kThreads)));
                spmv_sell_kernel<DOT, short, false, 1, 6, true, false, true><<<grid, kThreads, 0, c->stream>>>(A->rows, A->sell_ptr, A->sell_perm, A->sell_d16, A->sell_val, x, y,
                                                                                                          st, dot_out, c->red.partials, c->red.ticket, A->own_lo,
                                                                                                          A->own_hi, A->p2p_dev, A->p2p_epoch);
            } else if (unroll == 9) SELLU(9) else if (unroll == 18) SELLU(18) else if (unroll == 3) SELLU(3) else SELLU(6)
#undef SELLU
        }
    } else { if (A->sell_perm) SELL(int, true, 1, A->sell_idx) else SELL(int, false, 1, A->sell_idx) }
#undef SELL
    return PF2_OK;
}

int plan_spmv_pub(pf2_csr* A) { plan_spmv(A); return PF2_OK; }

// matrix-free operator: Ke0 travels through constant memory, re-uploaded only when another matrix (or new parameters) used it last
static const pf2_csr* g_mf_owner = nullptr;
static unsigned long long g_mf_owner_version = 0;

static int mf_constant(pf2_csr* A) {
    pf2_ctx* c = A->ctx;
    PF2_CHECK(A->mf_ready && A->mf_version > 0, "matrix-free operator: assemble once after pf2_csr_matrix_free");
    if (g_mf_owner != A || g_mf_owner_version != A->mf_version) {
        const int m = (1 << A->mf_dim) * A->mf_ndof;
        PF2_CUDA(cudaMemcpyToSymbolAsync(c_mf_ke0, A->mf_ke0, sizeof(double) * (size_t)m * m, 0, cudaMemcpyHostToDevice, c->stream));
        g_mf_owner = A; g_mf_owner_version = A->mf_version;
    }
    return PF2_OK;
}

// ---- nodal-space PCG pieces (called by solve_mf_nodal, solver.cu) ----
int mf_nodal_apply(pf2_csr* A, const double* p_old, const double* z, double* p_new, double* y);

// x0 != nullptr: warm start from the reduced-numbering guess x0 (y is scratch for K x0)
int mf_nodal_begin(pf2_csr* A, int jacobi, const double* b, double* bn, double* dn, double* xn, double* r, double* z, double* p0, double* p1, int itrmax, double eps,
                   const double* x0, double* y) {
    pf2_ctx* c = A->ctx;
    PF2_TRY(mf_constant(A));
    const size_t nfull = (size_t)A->mf_n[0] * A->mf_n[1] * A->mf_n[2] * A->mf_ndof;
    const int g = c->grid_for((long long)nfull, 2);
    mf_expand_kernel<<<g, kThreads, 0, c->stream>>>(nfull, A->mf_n2g, b, A->indptr, A->diagpos, A->data, jacobi, bn, dn);
    if (x0) {
        // y = K x0 through the fused operator: with beta = 0 and done = 0 it forms p_new = 0 * p_old + x0 (into the scratch buffer p1) and y = K p_new
        mf_expand_x_kernel<<<g, kThreads, 0, c->stream>>>(nfull, A->mf_n2g, x0, xn);
        PF2_CUDA(cudaMemsetAsync(A->st, 0, sizeof(CgState), c->stream));
        PF2_TRY(mf_nodal_apply(A, xn, xn, p1, y));
        c->launches++;
    }
    mf_init_kernel<<<std::min(g, c->sm_count * 4), kThreads, 0, c->stream>>>(nfull, bn, dn, xn, r, z, p0, p1, A->st, itrmax, eps, c->red.partials, c->red.ticket,
                                                                           x0 ? y : nullptr);
    PF2_LAUNCH_CHECK();
    c->launches += 2;
    return PF2_OK;
}
int mf_nodal_apply(pf2_csr* A, const double* p_old, const double* z, double* p_new, double* y) {
    pf2_ctx* c = A->ctx;
    MfGrid G;
    G.n[0] = A->mf_n[0]; G.n[1] = A->mf_n[1]; G.n[2] = A->mf_n[2];
    G.nnode = A->mf_n[0] * A->mf_n[1] * A->mf_n[2];
#define MFN(D, N)                                                                                                                      \
    {                                                                                                                                  \
        const int nb = ((G.n[0] + MfTile<D>::TI - 1) / MfTile<D>::TI) * ((G.n[1] + MfTile<D>::TJ - 1) / MfTile<D>::TJ) *               \
                       ((D) == 3 ? (G.n[2] + MfTile<D>::TK - 1) / MfTile<D>::TK : 1);                                                  \
        const int grid = std::max(1, std::min(nb, c->wave_grid((const void*)spmv_mf_nodal_kernel<D, N>, kThreads)));                   \
        spmv_mf_nodal_kernel<D, N><<<grid, kThreads, 0, c->stream>>>(G, A->mf_n2g, A->mf_E, p_old, z, p_new, y, A->st, c->red.partials, c->red.ticket); \
    }
    if (A->mf_dim == 3) MFN(3, 3)
    else if (A->mf_ndof == 2) MFN(2, 2)
    else MFN(2, 1)
#undef MFN
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}
int mf_nodal_end(pf2_csr* A, const double* xn, double* x) {
    pf2_ctx* c = A->ctx;
    const size_t nfull = (size_t)A->mf_n[0] * A->mf_n[1] * A->mf_n[2] * A->mf_ndof;
    mf_gather_kernel<<<c->grid_for((long long)nfull, 2), kThreads, 0, c->stream>>>(nfull, A->mf_n2g, xn, x);
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

template <bool DOT>
static int launch_mf(pf2_csr* A, const double* x, double* y, const CgState* st, double* dot_out) {
    pf2_ctx* c = A->ctx;
    PF2_TRY(mf_constant(A));
    MfGrid G;
    G.n[0] = A->mf_n[0]; G.n[1] = A->mf_n[1]; G.n[2] = A->mf_n[2];
    G.nnode = A->mf_n[0] * A->mf_n[1] * A->mf_n[2];
#define MF(D, N)                                                                                                                     \
    {                                                                                                                                \
        const int nb = ((G.n[0] + MfTile<D>::TI - 1) / MfTile<D>::TI) * ((G.n[1] + MfTile<D>::TJ - 1) / MfTile<D>::TJ) *             \
                       ((D) == 3 ? (G.n[2] + MfTile<D>::TK - 1) / MfTile<D>::TK : 1);                                                \
        const int grid = std::max(1, std::min(nb, DOT ? c->wave_grid((const void*)spmv_mf_kernel<D, N, DOT>, kThreads) : c->sm_count * 16)); \
        spmv_mf_kernel<D, N, DOT><<<grid, kThreads, 0, c->stream>>>(G, A->mf_n2g, A->mf_E, x, y, st, dot_out, c->red.partials, c->red.ticket, \
                                                               A->own_lo, A->own_hi, A->p2p_dev, A->p2p_epoch);                      \
    }
    if (A->mf_dim == 3) MF(3, 3)
    else if (A->mf_ndof == 2) MF(2, 2)
    else MF(2, 1)
#undef MF
    return PF2_OK;
}

// called by assemble_device after every numeric assembly when the operator is enabled: per-element moduli and (when V / t changed) Ke0
int mf_update(pf2_csr* A, pf2_mesh* mesh, const double* modulus_dev, const double* rho_dev, const double params[5]) {
    pf2_ctx* c = A->ctx;
    const double E0 = params[0], E1 = params[1], V = params[2], p = params[3], t = params[4];
    mf_modulus_kernel<<<c->grid_for(A->mf_nelem), kThreads, 0, c->stream>>>(A->mf_nelem, modulus_dev, rho_dev, E0, E1, p, A->mf_E);
    PF2_LAUNCH_CHECK();
    c->launches++;
    if (A->mf_version == 0 || V != A->mf_V || t != A->mf_t) {
        const int npe = 1 << A->mf_dim, dim = A->mf_dim;
        int nd[8];
        double xe[24];
        PF2_CUDA(cudaMemcpyAsync(nd, mesh->conn, sizeof(int) * npe, cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaStreamSynchronize(c->stream));
        for (int a = 0; a < npe; a++) PF2_CUDA(cudaMemcpyAsync(xe + a * dim, mesh->coords + (size_t)nd[a] * dim, sizeof(double) * dim, cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaStreamSynchronize(c->stream));
        PF2_TRY(pf2_element_matrix(c, A->mf_eq, xe, 1.0, V, t, A->mf_ke0));
        A->mf_V = V; A->mf_t = t;
        A->mf_version++;
    }
    return PF2_OK;
}

template <bool DOT>
static int launch_spmv(pf2_csr* A, int variant, const double* x, double* y, const CgState* st, double* dot_out) {
    pf2_ctx* c = A->ctx;
#define ARGS A->rows, A->indptr, A->indices, A->data, x, y, st, dot_out, c->red.partials, c->red.ticket, A->own_lo, A->own_hi, A->p2p_dev, A->p2p_epoch
#define VEC(T)                                                                                               \
    {                                                                                                        \
        long long nb = ((long long)A->rows + (kThreads / T) - 1) / (kThreads / T);                              \
        int grid = (int)std::min<long long>(std::max<long long>(nb, 1), DOT ? (long long)c->wave_grid((const void*)spmv_vector_kernel<T, DOT>, kThreads) : (long long)c->sm_count * 64); \
        spmv_vector_kernel<T, DOT><<<grid, kThreads, 0, c->stream>>>(ARGS);                                   \
    }
#define STR(Gv)                                                                                              \
    {                                                                                                        \
        long long nb = ((long long)A->rows + (kThreads / Gv) - 1) / (kThreads / Gv);                            \
        int grid = (int)std::min<long long>(std::max<long long>(nb, 1), (long long)c->wave_grid((const void*)spmv_stream_kernel<Gv, DOT>, kThreads)); \
        spmv_stream_kernel<Gv, DOT><<<grid, kThreads, 0, c->stream>>>(ARGS);                                  \
    }
    switch (variant) {
        case 1: VEC(2) break;
        case 2: VEC(4) break;
        case 3: VEC(8) break;
        case 4: VEC(16) break;
        case 5: VEC(32) break;
        case 11: STR(1) break;
        case 12: STR(2) break;
        case 13: STR(4) break;
        case 14: STR(8) break;
        case 15: STR(16) break;
        case 31: PF2_TRY((launch_sell<DOT>(A, x, y, st, dot_out))); break;
        case 41: PF2_TRY((launch_mf<DOT>(A, x, y, st, dot_out))); break;
        case 21: PF2_TRY((launch_tma<1, DOT>(A, x, y, st, dot_out))); break;
        case 22: PF2_TRY((launch_tma<2, DOT>(A, x, y, st, dot_out))); break;
        case 23: PF2_TRY((launch_tma<4, DOT>(A, x, y, st, dot_out))); break;
        case 24: PF2_TRY((launch_tma<8, DOT>(A, x, y, st, dot_out))); break;
        case 25: PF2_TRY((launch_tma<16, DOT>(A, x, y, st, dot_out))); break;
        case 26: PF2_TRY((launch_tma<32, DOT>(A, x, y, st, dot_out))); break;
        default: set_error("unknown SpMV variant %d", variant); return PF2_E_INVALID;
    }
#undef ARGS
#undef VEC
#undef STR
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

static bool variant_ok(const pf2_csr* A, int variant) {
    if (variant >= 1 && variant <= 5) return true;
    if (variant >= 11 && variant <= 15) {
        int G = 1 << (variant - 11);
        return (long long)(kThreads / G) * A->max_row <= kStreamCap;
    }
    if (variant == 31) return true;
    if (variant == 41) return A->mf_ready;
    if (variant >= 21 && variant <= 26) {
        int G = 1 << (variant - 21);
        return (long long)(kConsumers / G) * A->max_row <= kTileNnz;
    }
    return false;
}

int spmv(pf2_csr* A, const double* x, double* y) {
    plan_spmv(A);
    return launch_spmv<false>(A, A->spmv_variant, x, y, nullptr, nullptr);
}
int spmv_dot(pf2_csr* A, const double* x, double* y, const CgState* st, double* dot_out) {
    plan_spmv(A);
    return launch_spmv<true>(A, A->spmv_variant, x, y, st, dot_out);
}

}  // namespace pf2

using namespace pf2;

extern "C" {

int pf2_csr_upload(pf2_ctx* ctx, int rows, const int* indptr_host, const int* indices_host, const double* data_host, pf2_csr** out) {
    PF2_CHECK(ctx && out && rows >= 0 && indptr_host, "bad arguments");
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_csr* A = new pf2_csr();
    A->ctx = ctx;
    A->rows = rows;
    A->own_lo = 0; A->own_hi = rows;
    A->nnz = indptr_host[rows];
    // spare entries: the TMA SpMV reads 16-byte aligned windows that may run a few entries past the end
    PF2_TRY(dev_alloc(&A->indptr, (size_t)rows + 1 + kCsrPad));
    PF2_TRY(dev_alloc(&A->indices, (size_t)A->nnz + kCsrPad));
    PF2_TRY(dev_alloc(&A->data, (size_t)A->nnz + kCsrPad));
    PF2_CUDA(cudaMemsetAsync(A->indptr + rows, 0, sizeof(long long) * (1 + kCsrPad), ctx->stream));
    PF2_CUDA(cudaMemsetAsync(A->indices + A->nnz, 0, sizeof(int) * kCsrPad, ctx->stream));
    PF2_CUDA(cudaMemsetAsync(A->data + A->nnz, 0, sizeof(double) * kCsrPad, ctx->stream));
    PF2_TRY(dev_alloc(&A->F, (size_t)rows));
    int* tmp = nullptr;
    PF2_TRY(dev_alloc(&tmp, (size_t)rows + 1));
    PF2_CUDA(cudaMemcpyAsync(tmp, indptr_host, sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice, ctx->stream));
    widen_indptr_kernel<<<ctx->grid_for(rows + 1), kThreads, 0, ctx->stream>>>(rows + 1, tmp, A->indptr);
    PF2_LAUNCH_CHECK();
    ctx->launches++;
    PF2_CUDA(cudaMemcpyAsync(A->indices, indices_host, sizeof(int) * (size_t)A->nnz, cudaMemcpyHostToDevice, ctx->stream));
    if (data_host) PF2_CUDA(cudaMemcpyAsync(A->data, data_host, sizeof(double) * (size_t)A->nnz, cudaMemcpyHostToDevice, ctx->stream));
    else PF2_CUDA(cudaMemsetAsync(A->data, 0, sizeof(double) * (size_t)A->nnz, ctx->stream));
    PF2_CUDA(cudaMemsetAsync(A->F, 0, sizeof(double) * (size_t)rows, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    PF2_CUDA(cudaFree(tmp));
    PF2_TRY(csr_finalize_structure(A));
    *out = A;
    return PF2_OK;
}

int pf2_csr_destroy(pf2_csr* A) {
    if (!A) return PF2_OK;
    cudaSetDevice(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    void* ptrs[] = { A->indptr, A->indices, A->data, A->F, A->diagpos, A->bmap, A->slab, A->xw, A->bw, A->st, A->sell_ptr, A->sell_perm, A->sell_idx, A->sell_d16, A->sell_b32, A->sell_val, A->p2p_dev,
                     A->ilu, A->level_rows, A->level_rows_u, A->level_ptr, A->level_ptr_u, A->ilu_ready, A->n2e_ptr, A->n2e, A->node_row0 };
    for (void* p : ptrs) if (p) cudaFree(p);
    if (A->h_st) cudaFreeHost(A->h_st);
    if (A->pcg_sync) cudaFree(A->pcg_sync);
    if (A->h_pcg_sync) cudaFreeHost(A->h_pcg_sync);
    if (A->mf_E) cudaFree(A->mf_E);
    if (A->mf_slab) cudaFree(A->mf_slab);
    if (A->cg1_s) cudaFree(A->cg1_s);
    if (g_mf_owner == A) g_mf_owner = nullptr;
    if (A->bi_slab) cudaFree(A->bi_slab);
    if (A->bi_st) cudaFree(A->bi_st);
    if (A->bi_hst) cudaFreeHost(A->bi_hst);
    for (int i = 0; i < 2; i++) if (A->bi_ev[i]) cudaEventDestroy(A->bi_ev[i]);
    for (int i = 0; i < 2; i++) if (A->ev[i]) cudaEventDestroy(A->ev[i]);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 4; j++) if (A->pev[i][j]) cudaEventDestroy(A->pev[i][j]);
    delete A;
    return PF2_OK;
}

int pf2_csr_info(pf2_csr* A, int* rows, long long* nnz) {
    if (rows) *rows = A->rows;
    if (nnz) *nnz = A->nnz;
    return PF2_OK;
}

int pf2_csr_download(pf2_csr* A, long long* indptr_host, int* indices_host, double* data_host, double* F_host) {
    cudaStream_t s = A->ctx->stream;
    if (indptr_host) PF2_CUDA(cudaMemcpyAsync(indptr_host, A->indptr, sizeof(long long) * ((size_t)A->rows + 1), cudaMemcpyDeviceToHost, s));
    if (indices_host) PF2_CUDA(cudaMemcpyAsync(indices_host, A->indices, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost, s));
    if (data_host) PF2_CUDA(cudaMemcpyAsync(data_host, A->data, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost, s));
    if (F_host) PF2_CUDA(cudaMemcpyAsync(F_host, A->F, sizeof(double) * (size_t)A->rows, cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    return PF2_OK;
}

int pf2_csr_set_values(pf2_csr* A, const double* data_host) {
    PF2_CUDA(cudaMemcpyAsync(A->data, data_host, sizeof(double) * (size_t)A->nnz, cudaMemcpyHostToDevice, A->ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(A->ctx->stream));
    A->ilu_valid = false;
    A->sell_values_valid = false;
    return PF2_OK;
}
int pf2_csr_device_F(pf2_csr* A, double** F_dev) { *F_dev = A->F; return PF2_OK; }
int pf2_csr_device_data(pf2_csr* A, double** data_dev) { *data_dev = A->data; return PF2_OK; }

int pf2_spmv(pf2_csr* A, const double* x_dev, double* y_dev) { return spmv(A, x_dev, y_dev); }

int pf2_spmv_host(pf2_csr* A, const double* x_host, double* y_host) {
    pf2_ctx* c = A->ctx;
    double *x = nullptr, *y = nullptr;
    PF2_TRY(dev_alloc(&x, (size_t)A->rows));
    PF2_TRY(dev_alloc(&y, (size_t)A->rows));
    PF2_CUDA(cudaMemcpyAsync(x, x_host, sizeof(double) * (size_t)A->rows, cudaMemcpyHostToDevice, c->stream));
    int rc = spmv(A, x, y);
    if (rc == PF2_OK) {
        PF2_CUDA(cudaMemcpyAsync(y_host, y, sizeof(double) * (size_t)A->rows, cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaStreamSynchronize(c->stream));
    }
    cudaFree(x); cudaFree(y);
    return rc;
}

int pf2_spmv_set_tma_tuning(pf2_csr* A, int stages, int ctas_per_sm) {
    A->tma_stages = stages; A->tma_ctas_per_sm = ctas_per_sm;
    return PF2_OK;
}

int pf2_csr_matrix_free(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq) {
    PF2_CHECK(A && mesh && map, "null argument");
    pf2_ctx* c = A->ctx;
    int dim = 0, npe = 0, ndof = 0;
    PF2_TRY(pf2_eq_describe(eq, &dim, &npe, &ndof));
    const int phys = eq & 0xff, shape = (eq >> 8) & 0xff;
    const bool q4 = dim == 2 && npe == 4 && (shape == 0 || shape == PF2_SHAPE_Q4), h8 = dim == 3 && npe == 8 && (shape == 0 || shape == PF2_SHAPE_HEX8);
    if (!(q4 || h8) || phys == PF2_PHYS_MASS) { set_error("matrix-free operator: Q4 / hex8 stiffness selections only"); return PF2_E_UNSUPPORTED; }
    PF2_CHECK(mesh->dim == dim && mesh->npe == npe && map->ndof == ndof && A->rows == map->kdegree, "mesh / dof map / matrix do not belong together");
    // lattice dimensions from element 0 (x-major numbering: the second local node is one x-stride away)
    int nd[8];
    PF2_CUDA(cudaMemcpyAsync(nd, mesh->conn, sizeof(int) * npe, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    int n0 = 0, n1 = 0, n2 = 1;
    if (dim == 2) { n1 = nd[1] - nd[0]; }
    else { n2 = nd[3] - nd[0]; const int s0 = nd[1] - nd[0]; n1 = (n2 > 0) ? s0 / n2 : 0; if (n2 <= 1 || n1 * n2 != s0) n1 = 0; }
    if (n1 <= 1) { set_error("matrix-free operator: mesh is not an x-major structured lattice"); return PF2_E_UNSUPPORTED; }
    n0 = mesh->nnode / (n1 * n2);
    const long long ne = (long long)(n0 - 1) * (n1 - 1) * (dim == 3 ? n2 - 1 : 1);
    if (n0 <= 1 || (long long)n0 * n1 * n2 != mesh->nnode || ne != mesh->nelem) { set_error("matrix-free operator: mesh is not an x-major structured lattice"); return PF2_E_UNSUPPORTED; }
    MfGrid G;
    G.n[0] = n0; G.n[1] = n1; G.n[2] = n2; G.nnode = mesh->nnode;
    int* bad = nullptr;
    PF2_TRY(dev_alloc(&bad, 1));
    PF2_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    mf_verify_kernel<<<c->grid_for(mesh->nelem), kThreads, 0, c->stream>>>(dim, G, mesh->nelem, mesh->conn, mesh->coords, bad);
    int hbad = 0;
    PF2_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(bad);
    c->launches++;
    if (hbad) { set_error("matrix-free operator: elements are not congruent translates on an x-major lattice"); return PF2_E_UNSUPPORTED; }
    if (!A->mf_E) PF2_TRY(dev_alloc(&A->mf_E, (size_t)mesh->nelem));
    A->mf_dim = dim; A->mf_ndof = ndof; A->mf_eq = eq; A->mf_n[0] = n0; A->mf_n[1] = n1; A->mf_n[2] = n2; A->mf_nelem = mesh->nelem;
    A->mf_n2g = map->n2g;
    A->mf_version = 0;
    A->mf_ready = true;
    // nodal-numbering PCG with the fused direction update: measured faster in 2-D (0.0618 -> 0.0585 ms per iteration at 2 M dof), a wash
    // for hex8 where the operator is DFMA-bound; PF2_MF_NODAL=0 / 1 overrides
    A->mf_nodal = getenv("PF2_MF_NODAL") ? atoi(getenv("PF2_MF_NODAL")) != 0 : (dim == 2);
    A->spmv_variant = 41;
    return PF2_OK;
}

int pf2_spmv_set_variant(pf2_csr* A, int variant) {
    if (variant == 0) { A->spmv_variant = 0; plan_spmv(A); return PF2_OK; }
    if (!variant_ok(A, variant)) { set_error("SpMV variant %d not applicable (max row %d)", variant, A->max_row); return PF2_E_UNSUPPORTED; }
    A->spmv_variant = variant;
    return PF2_OK;
}

int pf2_spmv_bench(pf2_csr* A, int variant, int reps, int flush_l2, double* ms_per_spmv) {
    pf2_ctx* c = A->ctx;
    plan_spmv(A);
    if (variant == 0) variant = A->spmv_variant;
    if (!variant_ok(A, variant)) { set_error("SpMV variant %d not applicable (max row %d)", variant, A->max_row); return PF2_E_UNSUPPORTED; }
    double *x = nullptr, *y = nullptr;
    PF2_TRY(dev_alloc(&x, (size_t)A->rows));
    PF2_TRY(dev_alloc(&y, (size_t)A->rows));
    std::vector<double> hx(A->rows);
    for (int i = 0; i < A->rows; i++) hx[i] = 1.0 + 1e-3 * (i % 97);
    PF2_CUDA(cudaMemcpyAsync(x, hx.data(), sizeof(double) * (size_t)A->rows, cudaMemcpyHostToDevice, c->stream));
    for (int i = 0; i < 3; i++) PF2_TRY(launch_spmv<false>(A, variant, x, y, nullptr, nullptr));
    double total = 0;
    if (flush_l2) {
        for (int i = 0; i < reps; i++) {
            PF2_TRY(pf2_flush_l2(c));
            PF2_TRY(pf2_timer_start(c));
            PF2_TRY(launch_spmv<false>(A, variant, x, y, nullptr, nullptr));
            double ms;
            PF2_TRY(pf2_timer_stop(c, &ms));
            total += ms;
        }
    } else {
        PF2_TRY(pf2_timer_start(c));
        for (int i = 0; i < reps; i++) PF2_TRY(launch_spmv<false>(A, variant, x, y, nullptr, nullptr));
        PF2_TRY(pf2_timer_stop(c, &total));
    }
    *ms_per_spmv = total / reps;
    cudaFree(x); cudaFree(y);
    return PF2_OK;
}

}  // extern "C"
