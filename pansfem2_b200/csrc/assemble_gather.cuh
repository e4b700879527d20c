// assemble_gather.cuh -- NUMERIC phase of assembly without atomics: each stored entry of K is written exactly once.
//
// Replaces the same reference code as the scatter kernels (element routine + Assembling(K,F,u,Ke,...) Assembling.h:47-66), organised
// by ROW instead of by element: a CTA owns a tile of kGatherTile consecutive nodes.  Because free dofs are numbered node-major
// (Assembling.h:175-186), the rows of those nodes are consecutive and their stored entries are ONE contiguous range of `data`; the
// CTA accumulates that range in shared memory and writes it out with coalesced plain stores (no memset, no RED traffic).  One thread
// per node walks the node's adjacent elements in ascending element order -- the order in which the reference's element loop adds
// them -- computes the node's NDOF rows of each element matrix in registers (same routines as the scatter kernels) and adds them to
// its own rows of the tile, so no two threads ever touch the same entry: K and F come out bitwise identical from run to run.
// Used for the 2-D selections whenever the largest tile fits in shared memory; hex8 rows (3 x 81 entries per node) do not, and keep
// the scatter kernel.
#pragma once
#include "types.cuh"
#include "element.cuh"

namespace pf2 {

constexpr size_t kGatherSmemLimit = 96 * 1024;     // per CTA: keeps at least two CTAs per SM
// Tile entry k lives at shared index k + k/32: one padding word per 32 entries.  The lanes of a warp own consecutive nodes, i.e.
// offsets that differ by the node's row block (36 entries for Q4 elasticity = 288 B, which would put every 8th lane on the same
// bank); the skew spreads them over the banks while the write-out (consecutive k) stays conflict-free.
__device__ __forceinline__ int gather_slot(int k) { return k + (k >> 5); }
inline size_t gather_smem_bytes(size_t tile_entries) { return (tile_entries + (tile_entries >> 5) + 1) * sizeof(double); }

struct GatherArgs {
    int nnode;
    const double* coords; const int* conn; const int* n2g; const double* ufix;
    const int* n2e_ptr; const int* n2e; const int* node_row0; const int* bmap; const long long* indptr;
    const double* modulus; const double* rho; double E0, E1, p;
    double* data; double* F;
};

// ELEM: static DIM / NPE / NDOF and  __device__ void rows(X, a, acc) const  giving node a's rows for unit modulus
template <class ELEM>
__global__ void __launch_bounds__(kGatherTile)
assemble_gather_kernel(GatherArgs g, ELEM elem) {
    constexpr int DIM = ELEM::DIM, NPE = ELEM::NPE, NDOF = ELEM::NDOF;
    extern __shared__ double sbuf[];
    const int ntiles = (g.nnode + kGatherTile - 1) / kGatherTile;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n0 = tile * kGatherTile, n1 = min(n0 + kGatherTile, g.nnode);
        const long long base = g.indptr[g.node_row0[n0]];
        const int len = (int)(g.indptr[g.node_row0[n1]] - base);
        for (int k = threadIdx.x; k < len + (len >> 5) + 1; k += kGatherTile) sbuf[k] = 0.0;
        __syncthreads();
        const int node = n0 + threadIdx.x;
        if (node < n1) {
            int rows[NDOF], roff[NDOF];
            bool any = false;
#pragma unroll
            for (int i = 0; i < NDOF; i++) {
                rows[i] = g.n2g[(size_t)node * NDOF + i];
                roff[i] = (rows[i] != -1) ? (int)(g.indptr[rows[i]] - base) : 0;
                any |= (rows[i] != -1);
            }
            if (any) {
                double facc[NDOF];
#pragma unroll
                for (int i = 0; i < NDOF; i++) facc[i] = 0.0;
                const int qe = g.n2e_ptr[node + 1];
#pragma unroll 1
                for (int q = g.n2e_ptr[node]; q < qe; q++) {
                    const int e = g.n2e[q];
                    int nd[NPE], a = 0;
#pragma unroll
                    for (int n = 0; n < NPE; n++) { nd[n] = g.conn[(size_t)e * NPE + n]; if (nd[n] == node) a = n; }
                    double X[NPE][DIM];
#pragma unroll
                    for (int n = 0; n < NPE; n++)
#pragma unroll
                        for (int k = 0; k < DIM; k++) X[n][k] = g.coords[(size_t)nd[n] * DIM + k];
                    const double E = g.modulus ? g.modulus[e] : simp_modulus(g.rho[e], g.E0, g.E1, g.p);
                    double acc[NDOF][NPE * NDOF];
                    elem.rows(X, a, acc);
                    const int* bm = g.bmap + ((size_t)e * NPE + a) * NPE;
#pragma unroll
                    for (int b = 0; b < NPE; b++) {
                        const int off = bm[b];
                        int cfree[NDOF];
                        int rank = 0;
#pragma unroll
                        for (int j = 0; j < NDOF; j++) {
                            const int c = g.n2g[(size_t)nd[b] * NDOF + j];
                            cfree[j] = (c != -1) ? rank++ : -1;
                        }
#pragma unroll
                        for (int i = 0; i < NDOF; i++) {
                            if (rows[i] == -1) continue;
#pragma unroll
                            for (int j = 0; j < NDOF; j++) {
                                const double v = E * acc[i][b * NDOF + j];
                                if (cfree[j] >= 0) sbuf[gather_slot(roff[i] + off + cfree[j])] += v;             // Assembling.h:55
                                else {
                                    const double uf = g.ufix[(size_t)nd[b] * NDOF + j];
                                    if (uf != 0.0) facc[i] -= v * uf;                               // Assembling.h:59
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NDOF; i++) if (rows[i] != -1) g.F[rows[i]] = facc[i];
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < len; k += kGatherTile) g.data[base + k] = sbuf[gather_slot(k)];
        __syncthreads();
    }
}

// ---- hex8 (SolidLinearIsotropicElastic<ShapeFunction8Cubic, Gauss8Cubic>, Solid.h:21-64) ------------------------------------------------
// A node's three rows hold up to 3 x 81 entries, so the tile is kGather3Tile = 32 nodes (62 KB of shared memory) and every dof ROW gets its
// own thread: warp w of the 96-thread CTA owns dof row w of the tile's 32 nodes (one code path per warp), walks the node's <= 8 adjacent
// elements in ascending order, computes that one row of each element matrix in registers (solid_row<I>: bit-identical to the rows the
// scatter kernel adds) and accumulates it into its own range of the tile.  No atomics, no memset, every entry written once: K and F of a
// 3-D model are bitwise reproducible like the 2-D ones.  The price is that a node's three threads each form the Gauss-point gradients.
constexpr int kGather3Threads = 3 * kGather3Tile;

template <int I>
__device__ __forceinline__ void gather_hex8_row(const GatherArgs& g, double V, int node, int row, int roff, double* sbuf) {
    double facc = 0.0;
    const int qe = g.n2e_ptr[node + 1];
#pragma unroll 1
    for (int q = g.n2e_ptr[node]; q < qe; q++) {
        const int e = g.n2e[q];
        int nd[8], a = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) { nd[n] = g.conn[(size_t)e * 8 + n]; if (nd[n] == node) a = n; }
        double X[8][3];
#pragma unroll
        for (int n = 0; n < 8; n++)
#pragma unroll
            for (int k = 0; k < 3; k++) X[n][k] = g.coords[(size_t)nd[n] * 3 + k];
        const double E = g.modulus ? g.modulus[e] : simp_modulus(g.rho[e], g.E0, g.E1, g.p);
        double acc[24];
        solid_row<I>(X, a, V, acc);
        const int* bm = g.bmap + ((size_t)e * 8 + a) * 8;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int off = bm[b];
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int c = g.n2g[(size_t)nd[b] * 3 + j];
                const double v = E * acc[3 * b + j];
                if (c != -1) { sbuf[gather_slot(roff + off + rank)] += v; rank++; }                    // Assembling.h:55
                else {
                    const double uf = g.ufix[(size_t)nd[b] * 3 + j];
                    if (uf != 0.0) facc -= v * uf;                                                  // Assembling.h:59
                }
            }
        }
    }
    g.F[row] = facc;
}

static __global__ void __launch_bounds__(kGather3Threads)
assemble_gather_hex8_kernel(GatherArgs g, double V) {
    extern __shared__ double sbuf[];
    const int ntiles = (g.nnode + kGather3Tile - 1) / kGather3Tile;
    const int i = threadIdx.x / kGather3Tile, lane = threadIdx.x % kGather3Tile;      // dof row (warp-uniform), node of the tile
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n0 = tile * kGather3Tile, n1 = min(n0 + kGather3Tile, g.nnode);
        const long long base = g.indptr[g.node_row0[n0]];
        const int len = (int)(g.indptr[g.node_row0[n1]] - base);
        for (int k = threadIdx.x; k < len + (len >> 5) + 1; k += kGather3Threads) sbuf[k] = 0.0;
        __syncthreads();
        const int node = n0 + lane;
        if (node < n1) {
            const int row = g.n2g[(size_t)node * 3 + i];
            if (row != -1) {
                const int roff = (int)(g.indptr[row] - base);
                if (i == 0) gather_hex8_row<0>(g, V, node, row, roff, sbuf);
                else if (i == 1) gather_hex8_row<1>(g, V, node, row, roff, sbuf);
                else gather_hex8_row<2>(g, V, node, row, roff, sbuf);
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < len; k += kGather3Threads) g.data[base + k] = sbuf[gather_slot(k)];
        __syncthreads();
    }
}

inline bool gather3d_usable(const pf2_csr* A, const pf2_mesh* mesh) {
    static const bool off = getenv("PF2_ASSEMBLE_SCATTER") != nullptr;
    return !off && mesh->dim == 3 && mesh->npe == 8 && A->n2e_ptr && A->gather_nnode == mesh->nnode && A->gather_smem > 0 &&
           gather_smem_bytes(A->gather_smem / sizeof(double)) <= kGatherSmemLimit;
}
inline int assemble_gather_hex8_launch(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, double V, const double* modulus_dev, const double* rho_dev,
                                       double E0, double E1, double p) {
    pf2_ctx* c = A->ctx;
    PF2_CUDA(cudaFuncSetAttribute(assemble_gather_hex8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGatherSmemLimit));
    const GatherArgs g = { mesh->nnode, mesh->coords, mesh->conn, map->n2g, map->ufix, A->n2e_ptr, A->n2e, A->node_row0, A->bmap, A->indptr,
                           modulus_dev, rho_dev, E0, E1, p, A->data, A->F };
    const int ntiles = (mesh->nnode + kGather3Tile - 1) / kGather3Tile;
    const int grid = std::min(ntiles, c->sm_count * 32);
    assemble_gather_hex8_kernel<<<grid, kGather3Threads, gather_smem_bytes(A->gather_smem / sizeof(double)), c->stream>>>(g, V);
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

// true when the plan exists for this mesh and the largest tile fits
inline bool gather_usable(const pf2_csr* A, const pf2_mesh* mesh) {
    static const bool off = getenv("PF2_ASSEMBLE_SCATTER") != nullptr;     // tests / measurements: force the scatter kernels
    return !off && mesh->dim == 2 && A->n2e_ptr && A->gather_nnode == mesh->nnode && A->gather_smem > 0 && gather_smem_bytes(A->gather_smem / sizeof(double)) <= kGatherSmemLimit;
}

template <class ELEM>
int assemble_gather_launch(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, const ELEM& elem, const double* modulus_dev, const double* rho_dev,
                           double E0, double E1, double p) {
    pf2_ctx* c = A->ctx;
    // per function AND per device, so it is (re)stated on every launch: a host-side call of about a microsecond
    PF2_CUDA(cudaFuncSetAttribute(assemble_gather_kernel<ELEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGatherSmemLimit));
    const GatherArgs g = { mesh->nnode, mesh->coords, mesh->conn, map->n2g, map->ufix, A->n2e_ptr, A->n2e, A->node_row0, A->bmap, A->indptr,
                           modulus_dev, rho_dev, E0, E1, p, A->data, A->F };
    const int ntiles = (mesh->nnode + kGatherTile - 1) / kGatherTile;
    const int grid = std::min(ntiles, c->sm_count * 16);
    assemble_gather_kernel<ELEM><<<grid, kGatherTile, gather_smem_bytes(A->gather_smem / sizeof(double)), c->stream>>>(g, elem);
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

}  // namespace pf2
