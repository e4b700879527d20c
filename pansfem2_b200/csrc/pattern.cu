// pattern.cu -- mesh upload, Dirichlet numbering and the SYMBOLIC phase of assembly, all on the device.
//
// Replaces, once per mesh + boundary conditions instead of once per design iteration:
//   SetDirichlet (BoundaryCondition.h:20-25) + Renumbering (Assembling.h:175-186)  -> pf2_dofmap_create
//   the pattern that LILCSR<T>::set builds entry by entry (LILCSR.h:92-102, explicit zeros included, Assembling.h:55)
//   and the per-row sort of CSR(LILCSR&) (CSR.h:93-105)                               -> pf2_csr_pattern
// plus the precomputed scatter map used by the numeric phase (assemble.cu):
//   bmap[(e*npe+a)*npe + b] = offset, inside any row of element e's local node a, of the first free-dof column of
//   local node b.  Because numbering is node-major/dof-minor, all rows of one node share that offset, so the map is
//   npe^2 int32 per element (64 B for Q4, 256 B for hex8) instead of (npe*ndof)^2.
#include <cub/cub.cuh>
#include "types.cuh"

namespace pf2 {

int csr_finalize_structure(pf2_csr* A);

constexpr int kMaxCand = 512;   // candidate neighbour nodes of one node before sort/unique (elements/node * npe)

__global__ void mark_fixed_kernel(int nfixed, int ndof, const int* __restrict__ node, const int* __restrict__ dof,
                                  const double* __restrict__ val, int* n2g, double* ufix) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nfixed; i += gridDim.x * blockDim.x) {
        const size_t k = (size_t)node[i] * ndof + dof[i];
        n2g[k] = -1;
        ufix[k] = val[i];
    }
}
__global__ void free_flag_kernel(size_t n, const int* __restrict__ n2g, int* __restrict__ flag) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) flag[i] = (n2g[i] != -1);
}
__global__ void number_kernel(size_t n, const int* __restrict__ scan, int* __restrict__ n2g) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (n2g[i] != -1) n2g[i] = scan[i];
}

__global__ void count_n2e_kernel(size_t total, const int* __restrict__ conn, int* cnt) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) atomicAdd(&cnt[conn[i]], 1);
}
__global__ void fill_n2e_kernel(int nelem, int npe, const int* __restrict__ conn, const int* __restrict__ ptr, int* cursor, int* n2e) {
    const size_t total = (size_t)nelem * npe;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int node = conn[i];
        const int slot = atomicAdd(&cursor[node], 1);
        n2e[ptr[node] + slot] = (int)(i / npe);
    }
}

// gather plan: each node's element list in ascending order (the slots above were handed out by atomics), so that the numeric phase
// adds the element contributions of a row in the reference's own order (element loop ascending) and is bitwise reproducible
__global__ void sort_n2e_kernel(int nnode, const int* __restrict__ ptr, int* __restrict__ n2e) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnode; i += gridDim.x * blockDim.x) {
        const int b = ptr[i], e = ptr[i + 1];
        for (int a = b + 1; a < e; a++) {
            const int v = n2e[a];
            int j = a - 1;
            while (j >= b && n2e[j] > v) { n2e[j + 1] = n2e[j]; j--; }
            n2e[j + 1] = v;
        }
    }
}
__global__ void node_free_count_kernel(int nnode, int ndof, const int* __restrict__ n2g, int* __restrict__ cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nnode; i += gridDim.x * blockDim.x) {
        int c = 0;
        if (i < nnode) for (int d = 0; d < ndof; d++) c += (n2g[(size_t)i * ndof + d] != -1);
        cnt[i] = c;
    }
}
// largest number of stored entries owned by one tile of `tile` consecutive nodes
__global__ void tile_nnz_max_kernel(int nnode, int tile, const int* __restrict__ node_row0, const long long* __restrict__ indptr, unsigned long long* out) {
    const int ntiles = (nnode + tile - 1) / tile;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
        const int n0 = t * tile, n1 = min(n0 + tile, nnode);
        atomicMax(out, (unsigned long long)(indptr[node_row0[n1]] - indptr[node_row0[n0]]));
    }
}

// per node: sorted unique neighbour nodes.  PASS 0 counts, PASS 1 writes the list and the exclusive prefix of free dofs.
template <int PASS>
__global__ void node_adjacency_kernel(int nnode, int npe, int ndof, const int* __restrict__ conn, const int* __restrict__ n2e_ptr,
                                      const int* __restrict__ n2e, const int* __restrict__ n2g, int* __restrict__ adj_cnt,
                                      const int* __restrict__ adj_ptr, int* __restrict__ adj, int* __restrict__ adjfree,
                                      int* __restrict__ rowlen, int* overflow) {
    int cand[kMaxCand];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnode; i += gridDim.x * blockDim.x) {
        int k = 0;
        const int b = n2e_ptr[i], e = n2e_ptr[i + 1];
        if ((e - b) * npe > kMaxCand) { atomicExch(overflow, 1); if (PASS == 0) adj_cnt[i] = 0; continue; }
        for (int q = b; q < e; q++) {
            const int el = n2e[q];
            for (int a = 0; a < npe; a++) cand[k++] = conn[(size_t)el * npe + a];
        }
        // insertion sort (k <= 64 on structured meshes)
        for (int a = 1; a < k; a++) {
            const int v = cand[a];
            int j = a - 1;
            while (j >= 0 && cand[j] > v) { cand[j + 1] = cand[j]; j--; }
            cand[j + 1] = v;
        }
        int u = 0;
        for (int a = 0; a < k; a++) if (a == 0 || cand[a] != cand[a - 1]) cand[u++] = cand[a];
        if (PASS == 0) { adj_cnt[i] = u; }
        else {
            const int base = adj_ptr[i];
            int nfree = 0;
            for (int a = 0; a < u; a++) {
                adj[base + a] = cand[a];
                adjfree[base + a] = nfree;
                for (int d = 0; d < ndof; d++) nfree += (n2g[(size_t)cand[a] * ndof + d] != -1);
            }
            rowlen[i] = nfree;
        }
    }
}

__global__ void rowlen_kernel(int nnode, int ndof, const int* __restrict__ n2g, const int* __restrict__ rowlen, long long* __restrict__ indptr) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnode; i += gridDim.x * blockDim.x) {
        if (i == 0) indptr[0] = 0;
        for (int d = 0; d < ndof; d++) {
            const int r = n2g[(size_t)i * ndof + d];
            if (r != -1) indptr[r + 1] = rowlen[i];
        }
    }
}

__global__ void fill_indices_kernel(int nnode, int ndof, const int* __restrict__ n2g, const int* __restrict__ adj_ptr,
                                    const int* __restrict__ adj, const long long* __restrict__ indptr, int* __restrict__ indices) {
    const long long total = (long long)nnode * ndof;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / ndof);
        const int r = n2g[t];
        if (r == -1) continue;
        long long pos = indptr[r];
        for (int q = adj_ptr[i]; q < adj_ptr[i + 1]; q++) {
            const int nb = adj[q];
            for (int d = 0; d < ndof; d++) {
                const int c = n2g[(size_t)nb * ndof + d];
                if (c != -1) indices[pos++] = c;
            }
        }
    }
}

__global__ void bmap_kernel(int nelem, int npe, const int* __restrict__ conn, const int* __restrict__ adj_ptr,
                            const int* __restrict__ adj, const int* __restrict__ adjfree, int* __restrict__ bmap) {
    const long long total = (long long)nelem * npe * npe;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / (npe * npe));
        const int ab = (int)(t % (npe * npe));
        const int a = ab / npe, b = ab % npe;
        const int na = conn[(size_t)e * npe + a], nb = conn[(size_t)e * npe + b];
        int lo = adj_ptr[na], hi = adj_ptr[na + 1] - 1, pos = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int v = adj[mid];
            if (v == nb) { pos = mid; break; }
            if (v < nb) lo = mid + 1; else hi = mid - 1;
        }
        bmap[t] = pos >= 0 ? adjfree[pos] : -1;
    }
}

template <class T>
static int exclusive_scan(pf2_ctx* c, const T* in, T* out, size_t n) {
    void* tmp = nullptr;
    size_t bytes = 0;
    PF2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, c->stream));
    PF2_CUDA(cudaMalloc(&tmp, bytes ? bytes : 8));
    PF2_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    PF2_CUDA(cudaFree(tmp));
    c->launches += 2;
    return PF2_OK;
}
template <class T>
static int inclusive_scan_inplace(pf2_ctx* c, T* data, size_t n) {
    void* tmp = nullptr;
    size_t bytes = 0;
    PF2_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, data, data, n, c->stream));
    PF2_CUDA(cudaMalloc(&tmp, bytes ? bytes : 8));
    PF2_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, data, data, n, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    PF2_CUDA(cudaFree(tmp));
    c->launches += 2;
    return PF2_OK;
}

}  // namespace pf2

using namespace pf2;

extern "C" {

int pf2_mesh_create(pf2_ctx* ctx, int dim, int nnode, const double* coords_host, int npe, int nelem, const int* conn_host, pf2_mesh** out) {
    PF2_CHECK(ctx && out && coords_host && conn_host, "null argument");
    PF2_CHECK((dim == 2 && (npe == 2 || npe == 3 || npe == 4 || npe == 6 || npe == 8)) || (dim == 3 && (npe == 4 || npe == 8 || npe == 20)),
              "supported elements: T3 / Q4 / T6 / Q8 (dim 2; 3, 4, 6, 8 nodes; 2- and 3-node edges as load carriers), tet4 / hex8 / hex20 (dim 3; 4, 8, 20 nodes)");
    PF2_CHECK(nnode > 0 && nelem > 0, "empty mesh");
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_mesh* m = new pf2_mesh();
    m->ctx = ctx; m->dim = dim; m->nnode = nnode; m->npe = npe; m->nelem = nelem;
    m->own_elem_lo = 0; m->own_elem_hi = nelem;
    PF2_TRY(dev_alloc(&m->coords, (size_t)nnode * dim));
    PF2_TRY(dev_alloc(&m->conn, (size_t)nelem * npe));
    PF2_CUDA(cudaMemcpyAsync(m->coords, coords_host, sizeof(double) * (size_t)nnode * dim, cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaMemcpyAsync(m->conn, conn_host, sizeof(int) * (size_t)nelem * npe, cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = m;
    return PF2_OK;
}
// a second element list over the nodes of `base` (edges carrying surface loads, a sub-region carrying a body force): shares the
// coordinates on the device, so `base` must outlive it
int pf2_mesh_create_on_nodes(pf2_mesh* base, int npe, int nelem, const int* conn_host, pf2_mesh** out) {
    PF2_CHECK(base && out && conn_host && nelem > 0, "null argument");
    PF2_CHECK((base->dim == 2 && (npe == 2 || npe == 3 || npe == 4 || npe == 6 || npe == 8)) || (base->dim == 3 && (npe == 4 || npe == 8 || npe == 20)), "unsupported nodes per element");
    for (size_t k = 0; k < (size_t)nelem * npe; k++) PF2_CHECK(conn_host[k] >= 0 && conn_host[k] < base->nnode, "element node out of range");
    pf2_ctx* ctx = base->ctx;
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_mesh* m = new pf2_mesh();
    m->ctx = ctx; m->dim = base->dim; m->nnode = base->nnode; m->npe = npe; m->nelem = nelem;
    m->own_elem_lo = 0; m->own_elem_hi = nelem;
    m->coords = base->coords; m->shares_coords = true;
    PF2_TRY(dev_alloc(&m->conn, (size_t)nelem * npe));
    PF2_CUDA(cudaMemcpyAsync(m->conn, conn_host, sizeof(int) * (size_t)nelem * npe, cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = m;
    return PF2_OK;
}
int pf2_mesh_destroy(pf2_mesh* m) {
    if (!m) return PF2_OK;
    cudaStreamSynchronize(m->ctx->stream);
    if (!m->shares_coords) cudaFree(m->coords);
    cudaFree(m->conn);
    delete m;
    return PF2_OK;
}

int pf2_dofmap_create(pf2_ctx* ctx, int nnode, int ndof, int nfixed, const int* fix_node_host, const int* fix_dof_host,
                      const double* fix_val_host, int* kdegree_out, pf2_dofmap** out) {
    PF2_CHECK(ctx && out && nnode > 0 && ndof >= 1 && ndof <= 3 && nfixed >= 0, "bad arguments");
    for (int i = 0; i < nfixed; i++)
        PF2_CHECK(fix_node_host[i] >= 0 && fix_node_host[i] < nnode && fix_dof_host[i] >= 0 && fix_dof_host[i] < ndof, "Dirichlet entry out of range");
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_dofmap* m = new pf2_dofmap();
    m->ctx = ctx; m->nnode = nnode; m->ndof = ndof;
    const size_t n = (size_t)nnode * ndof;
    PF2_TRY(dev_alloc(&m->n2g, n));
    PF2_TRY(dev_alloc(&m->ufix, n));
    PF2_CUDA(cudaMemsetAsync(m->n2g, 0, sizeof(int) * n, ctx->stream));
    PF2_CUDA(cudaMemsetAsync(m->ufix, 0, sizeof(double) * n, ctx->stream));
    if (nfixed > 0) {
        int *dn = nullptr, *dd = nullptr;
        double* dv = nullptr;
        PF2_TRY(dev_alloc(&dn, (size_t)nfixed)); PF2_TRY(dev_alloc(&dd, (size_t)nfixed)); PF2_TRY(dev_alloc(&dv, (size_t)nfixed));
        PF2_CUDA(cudaMemcpyAsync(dn, fix_node_host, sizeof(int) * (size_t)nfixed, cudaMemcpyHostToDevice, ctx->stream));
        PF2_CUDA(cudaMemcpyAsync(dd, fix_dof_host, sizeof(int) * (size_t)nfixed, cudaMemcpyHostToDevice, ctx->stream));
        PF2_CUDA(cudaMemcpyAsync(dv, fix_val_host, sizeof(double) * (size_t)nfixed, cudaMemcpyHostToDevice, ctx->stream));
        mark_fixed_kernel<<<ctx->grid_for(nfixed), kThreads, 0, ctx->stream>>>(nfixed, ndof, dn, dd, dv, m->n2g, m->ufix);
        PF2_LAUNCH_CHECK();
        ctx->launches++;
        PF2_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(dn); cudaFree(dd); cudaFree(dv);
    }
    int *flag = nullptr, *scan = nullptr;
    PF2_TRY(dev_alloc(&flag, n)); PF2_TRY(dev_alloc(&scan, n));
    free_flag_kernel<<<ctx->grid_for((long long)n), kThreads, 0, ctx->stream>>>(n, m->n2g, flag);
    PF2_LAUNCH_CHECK();
    PF2_TRY(exclusive_scan(ctx, flag, scan, n));
    number_kernel<<<ctx->grid_for((long long)n), kThreads, 0, ctx->stream>>>(n, scan, m->n2g);
    PF2_LAUNCH_CHECK();
    ctx->launches += 2;
    int last_scan = 0, last_flag = 0;
    PF2_CUDA(cudaMemcpyAsync(&last_scan, scan + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PF2_CUDA(cudaMemcpyAsync(&last_flag, flag + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    m->kdegree = last_scan + last_flag;
    cudaFree(flag); cudaFree(scan);
    if (kdegree_out) *kdegree_out = m->kdegree;
    *out = m;
    return PF2_OK;
}
int pf2_dofmap_destroy(pf2_dofmap* m) {
    if (!m) return PF2_OK;
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->n2g); cudaFree(m->ufix);
    delete m;
    return PF2_OK;
}
int pf2_dofmap_get(pf2_dofmap* m, int* nodetoglobal_host) {
    PF2_CUDA(cudaMemcpyAsync(nodetoglobal_host, m->n2g, sizeof(int) * (size_t)m->nnode * m->ndof, cudaMemcpyDeviceToHost, m->ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return PF2_OK;
}

int pf2_csr_pattern(pf2_ctx* ctx, pf2_mesh* mesh, pf2_dofmap* map, pf2_csr** out) {
    PF2_CHECK(ctx && mesh && map && out, "null argument");
    PF2_CHECK(mesh->nnode == map->nnode, "mesh / dofmap node count mismatch");
    PF2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int nnode = mesh->nnode, npe = mesh->npe, nelem = mesh->nelem, ndof = map->ndof;
    const size_t nconn = (size_t)nelem * npe;
    // node -> elements
    int *cnt = nullptr, *n2e_ptr = nullptr, *cursor = nullptr, *n2e = nullptr;
    PF2_TRY(dev_alloc(&cnt, (size_t)nnode + 1)); PF2_TRY(dev_alloc(&n2e_ptr, (size_t)nnode + 1));
    PF2_TRY(dev_alloc(&cursor, (size_t)nnode)); PF2_TRY(dev_alloc(&n2e, nconn));
    PF2_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)nnode + 1), s));
    PF2_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)nnode, s));
    count_n2e_kernel<<<ctx->grid_for((long long)nconn), kThreads, 0, s>>>(nconn, mesh->conn, cnt);
    PF2_LAUNCH_CHECK();
    PF2_TRY(exclusive_scan(ctx, cnt, n2e_ptr, (size_t)nnode + 1));
    fill_n2e_kernel<<<ctx->grid_for((long long)nconn), kThreads, 0, s>>>(nelem, npe, mesh->conn, n2e_ptr, cursor, n2e);
    PF2_LAUNCH_CHECK();
    // node -> nodes (sorted unique), free-dof prefix
    int *adj_cnt = nullptr, *adj_ptr = nullptr, *adj = nullptr, *adjfree = nullptr, *rowlen = nullptr, *overflow = nullptr;
    PF2_TRY(dev_alloc(&adj_cnt, (size_t)nnode + 1)); PF2_TRY(dev_alloc(&adj_ptr, (size_t)nnode + 1));
    PF2_TRY(dev_alloc(&rowlen, (size_t)nnode)); PF2_TRY(dev_alloc(&overflow, 1));
    PF2_CUDA(cudaMemsetAsync(adj_cnt, 0, sizeof(int) * ((size_t)nnode + 1), s));
    PF2_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), s));
    const int gnode = ctx->grid_for(nnode);
    node_adjacency_kernel<0><<<gnode, kThreads, 0, s>>>(nnode, npe, ndof, mesh->conn, n2e_ptr, n2e, map->n2g, adj_cnt, nullptr, nullptr, nullptr, nullptr, overflow);
    PF2_LAUNCH_CHECK();
    PF2_TRY(exclusive_scan(ctx, adj_cnt, adj_ptr, (size_t)nnode + 1));
    int h_over = 0, nadj = 0;
    PF2_CUDA(cudaMemcpyAsync(&h_over, overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaMemcpyAsync(&nadj, adj_ptr + nnode, sizeof(int), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    PF2_CHECK(h_over == 0, "a node has too many adjacent elements for the symbolic phase");
    PF2_TRY(dev_alloc(&adj, (size_t)nadj)); PF2_TRY(dev_alloc(&adjfree, (size_t)nadj));
    node_adjacency_kernel<1><<<gnode, kThreads, 0, s>>>(nnode, npe, ndof, mesh->conn, n2e_ptr, n2e, map->n2g, nullptr, adj_ptr, adj, adjfree, rowlen, overflow);
    PF2_LAUNCH_CHECK();
    // CSR arrays
    pf2_csr* A = new pf2_csr();
    A->ctx = ctx;
    A->rows = map->kdegree;
    A->own_lo = 0; A->own_hi = A->rows;
    PF2_TRY(dev_alloc(&A->indptr, (size_t)A->rows + 1 + kCsrPad));
    PF2_CUDA(cudaMemsetAsync(A->indptr, 0, sizeof(long long) * ((size_t)A->rows + 1 + kCsrPad), s));
    rowlen_kernel<<<gnode, kThreads, 0, s>>>(nnode, ndof, map->n2g, rowlen, A->indptr);
    PF2_LAUNCH_CHECK();
    PF2_TRY(inclusive_scan_inplace(ctx, A->indptr, (size_t)A->rows + 1));
    PF2_CUDA(cudaMemcpyAsync(&A->nnz, A->indptr + A->rows, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    PF2_TRY(dev_alloc(&A->indices, (size_t)A->nnz + kCsrPad));
    PF2_TRY(dev_alloc(&A->data, (size_t)A->nnz + kCsrPad));
    PF2_TRY(dev_alloc(&A->F, (size_t)A->rows));
    PF2_CUDA(cudaMemsetAsync(A->indices + A->nnz, 0, sizeof(int) * kCsrPad, s));
    PF2_CUDA(cudaMemsetAsync(A->data, 0, sizeof(double) * ((size_t)A->nnz + kCsrPad), s));
    PF2_CUDA(cudaMemsetAsync(A->F, 0, sizeof(double) * (size_t)A->rows, s));
    fill_indices_kernel<<<ctx->grid_for((long long)nnode * ndof), kThreads, 0, s>>>(nnode, ndof, map->n2g, adj_ptr, adj, A->indptr, A->indices);
    PF2_LAUNCH_CHECK();
    // scatter map
    PF2_TRY(dev_alloc(&A->bmap, (size_t)nelem * npe * npe));
    A->map_npe = npe; A->map_ndof = ndof; A->map_nelem = nelem;
    bmap_kernel<<<ctx->grid_for((long long)nelem * npe * npe), kThreads, 0, s>>>(nelem, npe, mesh->conn, adj_ptr, adj, adjfree, A->bmap);
    PF2_LAUNCH_CHECK();
    ctx->launches += 7;
    // gather plan for the numeric phase (assemble_gather.cuh): kept for 2-D meshes only, where the kernel is used (a hex8 mesh would
    // carry 8 element ids per element for nothing: 0.45 GB on configs[4])
    const bool plan3d = (mesh->dim == 3 && npe == 8 && ndof == 3 && gather3d_requested());       // hex8 row gather, on request (PF2_ASSEMBLE_GATHER3D)
    if (mesh->dim == 2 || plan3d) {
        sort_n2e_kernel<<<gnode, kThreads, 0, s>>>(nnode, n2e_ptr, n2e);
        PF2_LAUNCH_CHECK();
        int* nfree = nullptr;
        unsigned long long* d_max = nullptr;
        PF2_TRY(dev_alloc(&nfree, (size_t)nnode + 1)); PF2_TRY(dev_alloc(&A->node_row0, (size_t)nnode + 1)); PF2_TRY(dev_alloc(&d_max, 1));
        node_free_count_kernel<<<gnode, kThreads, 0, s>>>(nnode, ndof, map->n2g, nfree);
        PF2_LAUNCH_CHECK();
        PF2_TRY(exclusive_scan(ctx, nfree, A->node_row0, (size_t)nnode + 1));
        PF2_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), s));
        tile_nnz_max_kernel<<<gnode, kThreads, 0, s>>>(nnode, plan3d ? kGather3Tile : kGatherTile, A->node_row0, A->indptr, d_max);
        PF2_LAUNCH_CHECK();
        unsigned long long h_max = 0;
        PF2_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, s));
        PF2_CUDA(cudaStreamSynchronize(s));
        cudaFree(nfree); cudaFree(d_max);
        A->n2e_ptr = n2e_ptr; A->n2e = n2e; A->gather_nnode = nnode;
        A->gather_smem = (size_t)h_max * sizeof(double);
        ctx->launches += 3;
        n2e_ptr = nullptr; n2e = nullptr;       // owned by the matrix from here on
    }
    PF2_CUDA(cudaStreamSynchronize(s));
    cudaFree(cnt); cudaFree(n2e_ptr); cudaFree(cursor); cudaFree(n2e);
    cudaFree(adj_cnt); cudaFree(adj_ptr); cudaFree(adj); cudaFree(adjfree); cudaFree(rowlen); cudaFree(overflow);
    PF2_TRY(csr_finalize_structure(A));
    *out = A;
    return PF2_OK;
}

}  // extern "C"
