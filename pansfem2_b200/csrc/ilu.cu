// ilu.cu -- ILU0 / PreILU0 of the reference (CG.h:258-315) on the device.
//
// ILU(0) keeps A's pattern: unit-L strictly below the diagonal, U with the diagonal.  Row i of the factorisation and of either
// triangular solve depends on the rows k < i (k > i for the backward solve) of its own pattern, so the available parallelism is the
// LEVEL structure of that dependency graph: ~ (nx + ny) levels of a few hundred rows each on a 2-D mesh in natural ordering.
//
//   schedule   built ON THE DEVICE once per pattern: one thread per row resolves its level as soon as the rows it depends on have
//              theirs (dependencies always point to lower thread indices, CTAs are dispatched in index order, so the wait is
//              deadlock-free), a radix sort orders the rows by level, a binary search finds where each level starts.  Only the level
//              pointers (a few thousand ints) come back to the host; the column indices never leave the device (they are 4 GB at
//              configs[3]).
//   factor     one launch per level (CG.h:262-281's entry order inside a row is kept exactly: the factors are bit-identical to a
//              serial run of the same operations); done once per design iteration.
//   sweeps     PreILU0's forward / backward substitution as ONE launch each.  FEM levels in natural ordering are narrow (a few hundred
//              rows on a 2-D mesh), so one CTA of 1024 threads -- or one thread-block cluster of 8 CTAs when levels reach a few
//              thousand rows -- walks the levels in order with a CTA / cluster barrier between them: ~1 us per level instead of a
//              kernel launch per level, and no other SM is kept busy waiting.  (A first cut gave every row its own thread spinning on
//              per-row ready words, the "synchronisation-free" SpTRSV: 160 k spinning threads polling L2 made a level cost 16 us --
//              6x SLOWER than a launch per level; it is kept behind PF2_ILU_SWEEP=syncfree, the launch-per-level form behind
//              PF2_ILU_SWEEP=level, which is also what very wide levels (3-D) use.)
//
// Partitioned matrix (row block [own_lo, own_hi) of a slab): block-Jacobi ILU(0) -- the factorisation and the sweeps see only the
// owned rows and the columns inside the owned range, so no rank waits for another one.  That changes the
// preconditioner, hence the iteration count, not the converged solution.
#include "types.cuh"
#include <cub/cub.cuh>

namespace pf2 {

// ---- level schedule ------------------------------------------------------------------------------------------------------------
// level[i] = 1 + max level of the rows i depends on (0 if none); rows outside [lo, hi) get level 0 and depend on nothing.
// FORWARD: thread t <-> row t, dependencies c < i.  BACKWARD: thread t <-> row n-1-t, dependencies c > i.
// Waiting pattern (here and in the sync-free sweeps): a thread may depend on a row owned by another lane of its OWN warp, so the store
// that publishes a result sits INSIDE the retry loop -- a lane that is done publishes before it reaches the loop's reconvergence
// point, where it would otherwise wait for the very lanes that are waiting for it.
template <bool FORWARD>
__global__ void ilu_levels_kernel(int n, int lo, int hi, const long long* __restrict__ indptr, const int* __restrict__ indices,
                                  volatile int* level) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = FORWARD ? t : n - 1 - t;
    if (i < lo || i >= hi) { level[i] = 0; return; }
    const long long s = indptr[i], e = indptr[i + 1];
    long long k = FORWARD ? s : e - 1;
    int l = 0;
    bool done = false;
    while (!done) {
        bool blocked = false;
        if (FORWARD) {
            while (k < e) {
                const int c = indices[k];
                if (c >= i) { k = e; break; }
                if (c >= lo) { const int lc = level[c]; if (lc < 0) { blocked = true; break; } l = max(l, lc + 1); }
                k++;
            }
        } else {
            while (k >= s) {
                const int c = indices[k];
                if (c <= i) { k = s - 1; break; }
                if (c < hi) { const int lc = level[c]; if (lc < 0) { blocked = true; break; } l = max(l, lc + 1); }
                k--;
            }
        }
        if (!blocked) { level[i] = l; __threadfence(); done = true; }
    }
}

__global__ void ilu_iota_kernel(int n, int* v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}
// ptr[l] = first position in the level-sorted key array whose level is >= l, for l = 0 .. L
__global__ void ilu_level_ptr_kernel(int n, int L, const int* __restrict__ sorted_level, int* __restrict__ ptr) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= L; l += gridDim.x * blockDim.x) {
        int a = 0, b = n;
        while (a < b) { const int m = (a + b) >> 1; if (sorted_level[m] < l) a = m + 1; else b = m; }
        ptr[l] = a;
    }
}

static int schedule(pf2_csr* A, bool lower, std::vector<int>& h_ptr, int** d_rows, int** d_ptr_out) {
    pf2_ctx* c = A->ctx;
    const int n = A->rows;
    const int lo = A->dist ? A->own_lo : 0, hi = A->dist ? A->own_hi : n;
    int *level = nullptr, *level_sorted = nullptr, *rows_in = nullptr, *d_ptr = nullptr, *d_max = nullptr;
    PF2_TRY(dev_alloc(&level, (size_t)n)); PF2_TRY(dev_alloc(&level_sorted, (size_t)n)); PF2_TRY(dev_alloc(&rows_in, (size_t)n));
    PF2_TRY(dev_alloc(d_rows, (size_t)n)); PF2_TRY(dev_alloc(&d_max, 1));
    PF2_CUDA(cudaMemsetAsync(level, 0xff, sizeof(int) * (size_t)n, c->stream));        // -1 = not resolved yet
    const int grid = (n + kThreads - 1) / kThreads;
    if (n > 0) {
        if (lower) ilu_levels_kernel<true><<<grid, kThreads, 0, c->stream>>>(n, lo, hi, A->indptr, A->indices, level);
        else ilu_levels_kernel<false><<<grid, kThreads, 0, c->stream>>>(n, lo, hi, A->indptr, A->indices, level);
        ilu_iota_kernel<<<c->grid_for(n), kThreads, 0, c->stream>>>(n, rows_in);
    }
    PF2_LAUNCH_CHECK();
    void* tmp = nullptr;
    size_t bytes = 0, bytes2 = 0;
    PF2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, level, level_sorted, rows_in, *d_rows, n, 0, 32, c->stream));
    PF2_CUDA(cub::DeviceReduce::Max(nullptr, bytes2, level, d_max, n, c->stream));
    PF2_CUDA(cudaMalloc(&tmp, std::max(bytes, bytes2) + 8));
    PF2_CUDA(cub::DeviceReduce::Max(tmp, bytes2, level, d_max, n, c->stream));
    PF2_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, level, level_sorted, rows_in, *d_rows, n, 0, 32, c->stream));      // stable: ascending row inside a level
    int maxl = -1;
    PF2_CUDA(cudaMemcpyAsync(&maxl, d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    const int L = n ? maxl + 1 : 0;
    h_ptr.assign((size_t)L + 1, 0);
    if (L > 0) {
        PF2_TRY(dev_alloc(&d_ptr, (size_t)L + 1));
        ilu_level_ptr_kernel<<<c->grid_for(L + 1), kThreads, 0, c->stream>>>(n, L, level_sorted, d_ptr);
        PF2_LAUNCH_CHECK();
        PF2_CUDA(cudaMemcpyAsync(h_ptr.data(), d_ptr, sizeof(int) * ((size_t)L + 1), cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->launches += 5;
    cudaFree(tmp); cudaFree(level); cudaFree(level_sorted); cudaFree(rows_in); cudaFree(d_max);
    *d_ptr_out = d_ptr;
    return PF2_OK;
}

int ilu0_build_levels(pf2_csr* A) {
    if (A->level_rows) return PF2_OK;
    PF2_TRY(schedule(A, true, A->h_level_ptr, &A->level_rows, &A->level_ptr));
    PF2_TRY(schedule(A, false, A->h_level_ptr_u, &A->level_rows_u, &A->level_ptr_u));
    A->level_width = 0;
    for (size_t l = 0; l + 1 < A->h_level_ptr.size(); l++) A->level_width = std::max(A->level_width, A->h_level_ptr[l + 1] - A->h_level_ptr[l]);
    for (size_t l = 0; l + 1 < A->h_level_ptr_u.size(); l++) A->level_width = std::max(A->level_width, A->h_level_ptr_u[l + 1] - A->h_level_ptr_u[l]);
    if (!A->ilu_ready) {
        PF2_TRY(dev_alloc(&A->ilu_ready, (size_t)A->rows));
        PF2_CUDA(cudaMemsetAsync(A->ilu_ready, 0, sizeof(unsigned int) * (size_t)A->rows, A->ctx->stream));
        A->ilu_epoch = 0;
    }
    return PF2_OK;
}

// ---- factorisation: one thread factors one row of the current level following the reference's entry order (CG.h:262-281) --------
__global__ void ilu0_level_kernel(int nrows_level, const int* __restrict__ rows_of_level, int lo, int hi, const long long* __restrict__ indptr,
                                  const int* __restrict__ indices, const int* __restrict__ diagpos,
                                  const double* __restrict__ a, double* __restrict__ q) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows_level) return;
    const int i = rows_of_level[t];
    const long long s = indptr[i], e = indptr[i + 1];
    if (i < lo || i >= hi) {              // ghost row of a slab: not part of this rank's block
        for (long long n = s; n < e; n++) q[n] = a[n];
        return;
    }
    for (long long n = s; n < e; n++) {
        const int j = indices[n];
        double qij = a[n];
        if (j < lo || j >= hi) { q[n] = qij; continue; }      // coupling to a ghost column: outside the block (never read by the sweeps)
        const int lim = (i <= j) ? i : j;
        for (long long qq = s; qq < e; qq++) {
            const int k = indices[qq];
            if (k >= lim) break;
            if (k < lo) continue;
            // find j in row k
            long long l0 = indptr[k], h0 = indptr[k + 1] - 1;
            while (l0 <= h0) {
                long long mid = (l0 + h0) >> 1;
                int cc = indices[mid];
                if (cc == j) { qij -= q[qq] * q[mid]; break; }
                if (cc < j) l0 = mid + 1; else h0 = mid - 1;
            }
        }
        if (i > j) qij /= q[indptr[j] + diagpos[j]];
        q[n] = qij;
    }
}

int ilu0_factor(pf2_csr* A) {
    if (A->ilu_valid) return PF2_OK;
    pf2_ctx* c = A->ctx;
    PF2_TRY(ilu0_build_levels(A));
    if (!A->ilu) PF2_TRY(dev_alloc(&A->ilu, (size_t)A->nnz));
    const int lo = A->dist ? A->own_lo : 0, hi = A->dist ? A->own_hi : A->rows;
    const int L = (int)A->h_level_ptr.size() - 1;
    for (int l = 0; l < L; l++) {
        const int cnt = A->h_level_ptr[l + 1] - A->h_level_ptr[l];
        if (cnt <= 0) continue;
        ilu0_level_kernel<<<(cnt + 127) / 128, 128, 0, c->stream>>>(cnt, A->level_rows + A->h_level_ptr[l], lo, hi, A->indptr, A->indices,
                                                                    A->diagpos, A->data, A->ilu);
        c->launches++;
    }
    PF2_LAUNCH_CHECK();
    A->ilu_valid = true;
    return PF2_OK;
}

// ---- sweeps ------------------------------------------------------------------------------------------------------------------------
// one level of the forward (unit-L) or backward (U) substitution of PreILU0 (CG.h:289-315); one thread per row
template <bool FORWARD>
__global__ void ilu0_sweep_level_kernel(int nrows_level, const int* __restrict__ rows_of_level, int lo, int hi,
                                        const long long* __restrict__ indptr, const int* __restrict__ indices,
                                        const int* __restrict__ diagpos, const double* __restrict__ q,
                                        double* __restrict__ v, const CgState* __restrict__ st) {
    if (st != nullptr && st->done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows_level) return;
    const int i = rows_of_level[t];
    if (i < lo || i >= hi) return;
    const long long s = indptr[i], e = indptr[i + 1];
    double vi = v[i];
    if (FORWARD) {
        for (long long k = s; k < e; k++) {
            const int c = indices[k];
            if (c >= i) break;
            if (c >= lo) vi -= q[k] * v[c];
        }
    } else {
        for (long long k = e - 1; k >= s; k--) {
            const int c = indices[k];
            if (c <= i) break;
            if (c < hi) vi -= q[k] * v[c];
        }
        vi /= q[s + diagpos[i]];
    }
    v[i] = vi;
}

// The whole sweep in one launch: thread t owns the t-th row in level order; ready[c] == epoch says v[c] is final.  Every row a thread
// waits for sits earlier in level order, i.e. in a CTA that was dispatched before its own: no deadlock.  v is read and written through
// L2 (volatile) so that a value published on another SM is the one that is read.
template <bool FORWARD>
__global__ void __launch_bounds__(kThreads)
ilu0_sweep_syncfree_kernel(int n, const int* __restrict__ rows_by_level, int lo, int hi, const long long* __restrict__ indptr,
                           const int* __restrict__ indices, const int* __restrict__ diagpos, const double* __restrict__ q,
                           volatile double* v, volatile unsigned int* ready, unsigned int epoch, const CgState* __restrict__ st) {
    if (st != nullptr && st->done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = rows_by_level[t];
    if (i < lo || i >= hi) return;
    const long long s = indptr[i], e = indptr[i + 1];
    double vi = v[i];
    long long k = FORWARD ? s : e - 1;
    bool done = false;
    while (!done) {                     // see ilu_levels_kernel: the publishing store stays inside the retry loop
        bool blocked = false;
        if (FORWARD) {
            while (k < e) {
                const int c = indices[k];
                if (c >= i) { k = e; break; }
                if (c >= lo) { if (ready[c] != epoch) { blocked = true; break; } vi -= q[k] * v[c]; }
                k++;
            }
        } else {
            while (k >= s) {
                const int c = indices[k];
                if (c <= i) { k = s - 1; break; }
                if (c < hi) { if (ready[c] != epoch) { blocked = true; break; } vi -= q[k] * v[c]; }
                k--;
            }
        }
        if (!blocked) {
            if (!FORWARD) vi /= q[s + diagpos[i]];
            v[i] = vi;
            __threadfence();
            ready[i] = epoch;
            done = true;
        }
    }
}

// The whole sweep in one launch, levels walked in order by ONE CTA (CLUSTER = 1: __syncthreads between levels, v travels through
// this SM's L1) or one cluster of CLUSTER CTAs (cluster barrier with release / acquire at cluster scope; v is read through L2).
constexpr int kSweepThreads = 1024;
template <bool FORWARD, int CLUSTER>
__global__ void __launch_bounds__(kSweepThreads)
ilu0_sweep_cta_kernel(int nlevels, const int* __restrict__ level_ptr, const int* __restrict__ rows_by_level, int lo, int hi,
                      const long long* __restrict__ indptr, const int* __restrict__ indices, const int* __restrict__ diagpos,
                      const double* __restrict__ q, double* v, const CgState* __restrict__ st) {
    if (st != nullptr && st->done) return;
    const int tid = (CLUSTER > 1 ? (int)(blockIdx.x % CLUSTER) * kSweepThreads : 0) + (int)threadIdx.x;
    for (int l = FORWARD ? 1 : 0; l < nlevels; l++) {          // forward: level-0 rows have no strictly-lower entries
        const int b = level_ptr[l], cnt = level_ptr[l + 1] - b;
        for (int t = tid; t < cnt; t += CLUSTER * kSweepThreads) {
            const int i = rows_by_level[b + t];
            if (i < lo || i >= hi) continue;
            const long long s = indptr[i], e = indptr[i + 1];
            double vi = v[i];
            const int dp = diagpos[i];
            if (dp >= 0) {
                // the strictly-lower entries are [s, s + dp), the strictly-upper ones (s + dp, e): no data-dependent loop exit, so the
                // column indices, the factors and the v entries of a chunk are all requested before the first is used (three dependent
                // load rounds per row instead of one per entry); the subtractions keep the reference's order (ascending / descending k)
                constexpr int CH = 8;
                if (FORWARD) {
                    for (int k0 = 0; k0 < dp; k0 += CH) {
                        int cc[CH];
                        double qq[CH], vv[CH];
#pragma unroll
                        for (int u = 0; u < CH; u++) {
                            const bool in = k0 + u < dp;
                            cc[u] = in ? indices[s + k0 + u] : -1;
                            qq[u] = in ? q[s + k0 + u] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < CH; u++) vv[u] = (cc[u] >= lo) ? (CLUSTER > 1 ? __ldcg(v + cc[u]) : v[cc[u]]) : 0.0;
#pragma unroll
                        for (int u = 0; u < CH; u++) if (cc[u] >= lo) vi -= qq[u] * vv[u];
                    }
                } else {
                    const int nup = (int)(e - s) - dp - 1;
                    for (int k0 = 0; k0 < nup; k0 += CH) {
                        int cc[CH];
                        double qq[CH], vv[CH];
#pragma unroll
                        for (int u = 0; u < CH; u++) {
                            const bool in = k0 + u < nup;
                            const long long k = e - 1 - (k0 + u);
                            cc[u] = in ? indices[k] : -1;
                            qq[u] = in ? q[k] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < CH; u++) vv[u] = (cc[u] >= 0 && cc[u] < hi) ? (CLUSTER > 1 ? __ldcg(v + cc[u]) : v[cc[u]]) : 0.0;
#pragma unroll
                        for (int u = 0; u < CH; u++) if (cc[u] >= 0 && cc[u] < hi) vi -= qq[u] * vv[u];
                    }
                    vi /= q[s + dp];
                }
            } else if (FORWARD) {
                for (long long k = s; k < e; k++) {
                    const int c = indices[k];
                    if (c >= i) break;
                    if (c >= lo) vi -= q[k] * (CLUSTER > 1 ? __ldcg(v + c) : v[c]);
                }
            } else {
                for (long long k = e - 1; k >= s; k--) {
                    const int c = indices[k];
                    if (c <= i) break;
                    if (c < hi) vi -= q[k] * (CLUSTER > 1 ? __ldcg(v + c) : v[c]);
                }
                vi /= q[s + diagpos[i]];
            }
            v[i] = vi;
        }
        if (CLUSTER > 1) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            __syncthreads();
        }
    }
}

template <bool FORWARD>
static int launch_sweep_cta(pf2_csr* A, int cluster, int nlevels, const int* level_ptr, const int* rows, int lo, int hi, const double* q, double* v,
                            const CgState* st) {
    pf2_ctx* c = A->ctx;
    if (cluster == 1) {
        ilu0_sweep_cta_kernel<FORWARD, 1><<<1, kSweepThreads, 0, c->stream>>>(nlevels, level_ptr, rows, lo, hi, A->indptr, A->indices, A->diagpos, q, v, st);
    } else {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(8); cfg.blockDim = dim3(kSweepThreads); cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PF2_CUDA(cudaLaunchKernelEx(&cfg, ilu0_sweep_cta_kernel<FORWARD, 8>, nlevels, level_ptr, rows, lo, hi, (const long long*)A->indptr, (const int*)A->indices,
                                    (const int*)A->diagpos, q, v, st));
    }
    c->launches++;
    return PF2_OK;
}

// v = (LU)^-1 v in place; `factors` defaults to A's cached ILU(0)
int ilu0_apply(pf2_csr* A, double* v, const CgState* st, const double* factors) {
    pf2_ctx* c = A->ctx;
    const double* q = factors ? factors : A->ilu;
    const int n = A->rows;
    const int lo = A->dist ? A->own_lo : 0, hi = A->dist ? A->own_hi : n;
    // read per call (tests switch it): cta (default where levels are narrow enough) | level | syncfree
    const char* mode = getenv("PF2_ILU_SWEEP");
    const bool level_launch = (mode && !strcmp(mode, "level")) || (getenv("PF2_ILU_LEVEL_LAUNCH") != nullptr && atoi(getenv("PF2_ILU_LEVEL_LAUNCH")) != 0);
    const bool syncfree = mode && !strcmp(mode, "syncfree");
    if (!level_launch && !syncfree && n > 0 && A->level_width <= 8 * 2 * kSweepThreads) {
        const int cluster = A->level_width <= 2 * kSweepThreads ? 1 : 8;
        PF2_TRY(launch_sweep_cta<true>(A, cluster, (int)A->h_level_ptr.size() - 1, A->level_ptr, A->level_rows, lo, hi, q, v, st));
        PF2_TRY(launch_sweep_cta<false>(A, cluster, (int)A->h_level_ptr_u.size() - 1, A->level_ptr_u, A->level_rows_u, lo, hi, q, v, st));
        PF2_LAUNCH_CHECK();
        return PF2_OK;
    }
    if (syncfree && n > 0) {
        const int grid = (n + kThreads - 1) / kThreads;
        if (A->ilu_epoch >= 0xfffffff0u) {       // epochs are compared for equality: start over long before they could wrap
            PF2_CUDA(cudaMemsetAsync(A->ilu_ready, 0, sizeof(unsigned int) * (size_t)n, c->stream));
            A->ilu_epoch = 0;
        }
        ilu0_sweep_syncfree_kernel<true><<<grid, kThreads, 0, c->stream>>>(n, A->level_rows, lo, hi, A->indptr, A->indices, A->diagpos, q, v,
                                                                          A->ilu_ready, ++A->ilu_epoch, st);
        ilu0_sweep_syncfree_kernel<false><<<grid, kThreads, 0, c->stream>>>(n, A->level_rows_u, lo, hi, A->indptr, A->indices, A->diagpos, q, v,
                                                                           A->ilu_ready, ++A->ilu_epoch, st);
        c->launches += 2;
        PF2_LAUNCH_CHECK();
        return PF2_OK;
    }
    const int L = (int)A->h_level_ptr.size() - 1, Lu = (int)A->h_level_ptr_u.size() - 1;
    for (int l = 1; l < L; l++) {      // level 0 rows have no strictly-lower entries
        const int cnt = A->h_level_ptr[l + 1] - A->h_level_ptr[l];
        if (cnt <= 0) continue;
        ilu0_sweep_level_kernel<true><<<(cnt + 127) / 128, 128, 0, c->stream>>>(cnt, A->level_rows + A->h_level_ptr[l], lo, hi, A->indptr,
                                                                                A->indices, A->diagpos, q, v, st);
        c->launches++;
    }
    for (int l = 0; l < Lu; l++) {
        const int cnt = A->h_level_ptr_u[l + 1] - A->h_level_ptr_u[l];
        if (cnt <= 0) continue;
        ilu0_sweep_level_kernel<false><<<(cnt + 127) / 128, 128, 0, c->stream>>>(cnt, A->level_rows_u + A->h_level_ptr_u[l], lo, hi, A->indptr,
                                                                                 A->indices, A->diagpos, q, v, st);
        c->launches++;
    }
    PF2_LAUNCH_CHECK();
    return PF2_OK;
}

}  // namespace pf2
