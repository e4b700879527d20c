// spmv_mf.cuh -- matrix-free application of the assembled operator on a UNIFORM structured mesh (opt-in SpMV variant 41).
//
// On the structured benchmark meshes (SquareMesh.h:62-107 numbering, our x-major hex mesher) every element is congruent, so
// Ke = E_e * Ke0 with one shared unit-modulus matrix Ke0 and    y = K p = sum_e E_e * scatter(Ke0 * gather(p, e)).
// One thread owns one node: it gathers p on the 3^dim neighbouring nodes (through nodetoglobal, fixed dofs read as 0 -- which is
// exactly Dirichlet by elimination), walks the 2^dim adjacent elements with Ke0 in constant memory (immediate operands of the
// DFMA) and writes its NDOF rows.  Traffic per dof: p 8 B + y 8 B + nodetoglobal 4 B + E 8/ndof B  (~24 B) instead of the CSR
// stream's 10..12 B per NONZERO (~384 B/dof in 2-D elasticity, ~834 B/dof for hex8), so the operator is bound by the DFMA pipe,
// not by HBM.  The product equals the CSR product up to the order of the floating-point sums (K_ij p_j is sum_e (E_e Ke0_ij) p_j
// there, sum_e E_e (Ke0_ij p_j) here); the Krylov recurrences, preconditioner (the assembled diagonal) and stopping test are
// untouched.  CSR assembly still runs (the matrix stays available for download, ILU(0) and the default CSR kernels).
#pragma once
#include "types.cuh"
#include "p2p.cuh"

namespace pf2 {

__constant__ double c_mf_ke0[576];      // (npe*ndof)^2 <= 24^2, row-major, unit modulus

struct MfGrid {
    int n[3];       // nodes per axis (x-major: node id = (i*n[1] + j)*n[2] + k ; 2-D: n[2] = 1)
    int nnode;
};

// corner offsets of local node a: Q4 (0,0)(1,0)(1,1)(0,1) ; hex8 bottom face then top face (ShapeFunction.h:171, 299)
__device__ __forceinline__ void mf_corner(int dim, int a, int& ox, int& oy, int& oz) {
    ox = ((a + 1) & 2) ? 1 : 0;
    oy = (a & 2) ? 1 : 0;
    oz = (dim == 3 && (a & 4)) ? 1 : 0;
}

// Tile of nodes one CTA owns per step (k fastest, like the numbering) and its one-node halo.
template <int DIM> struct MfTile;
template <> struct MfTile<2> { static constexpr int TI = 4, TJ = 64, TK = 1; };
template <> struct MfTile<3> { static constexpr int TI = 4, TJ = 4, TK = 16; };     // 16 consecutive k per half-warp: conflict-free LDS.64

// One CTA = one tile of kThreads nodes.  Phase 1 stages p of the tile + halo in shared memory (ONE nodetoglobal lookup and one
// gather per staged node instead of 3^dim per node); phase 2: one thread per node walks its 2^dim elements out of shared memory.
template <int DIM, int NDOF, bool DOT>
__global__ void __launch_bounds__(kThreads)      // grid_sum_last folds kThreads / 32 warp sums
spmv_mf_kernel(MfGrid G, const int* __restrict__ n2g, const double* __restrict__ E, const double* __restrict__ x, double* __restrict__ y,
               const CgState* __restrict__ st, double* dot_out, double* partials, unsigned int* ticket, int dot_lo, int dot_hi,
               const P2PView* p2p, unsigned long long* p2p_epoch) {
    if (DOT && st != nullptr && st->done) return;
    constexpr int NC = 1 << DIM, M = NC * NDOF;
    constexpr int TI = MfTile<DIM>::TI, TJ = MfTile<DIM>::TJ, TK = MfTile<DIM>::TK;
    constexpr int HJ = TJ + 2, HK = (DIM == 3) ? TK + 2 : 1, HI = TI + 2, HALO = HI * HJ * HK;
    static_assert(TI * TJ * TK == kThreads, "one thread per tile node");
    __shared__ double sp[NDOF][HALO];      // one plane per dof: a half-warp reads 16 consecutive doubles
    __shared__ int srow[NDOF][HALO];
    const int n0 = G.n[0], n1 = G.n[1], n2 = G.n[2];
    const int e1 = n1 - 1, e2 = (DIM == 3) ? n2 - 1 : 1;
    const int tiles_i = (n0 + TI - 1) / TI, tiles_j = (n1 + TJ - 1) / TJ, tiles_k = (DIM == 3) ? (n2 + TK - 1) / TK : 1;
    const int ntiles = tiles_i * tiles_j * tiles_k;
    // this thread's node inside the tile and its centre slot in the halo box
    const int tk = threadIdx.x % TK, tj = (threadIdx.x / TK) % TJ, ti = threadIdx.x / (TK * TJ);
    const int hc = ((ti + 1) * HJ + (tj + 1)) * HK + ((DIM == 3) ? tk + 1 : 0);
    double dot = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int bk = tile % tiles_k, bj = (tile / tiles_k) % tiles_j, bi = tile / (tiles_k * tiles_j);
        const int i0 = bi * TI, j0 = bj * TJ, k0 = bk * TK;
        // phase 1: stage p (and the row numbers) of the tile + halo
        for (int h = threadIdx.x; h < HALO; h += kThreads) {
            const int hk = h % HK, hj = (h / HK) % HJ, hi = h / (HK * HJ);
            const int gi = i0 - 1 + hi, gj = j0 - 1 + hj, gk = (DIM == 3) ? k0 - 1 + hk : 0;
            const bool in = (unsigned)gi < (unsigned)n0 && (unsigned)gj < (unsigned)n1 && (unsigned)gk < (unsigned)n2;
            const size_t nid = ((size_t)gi * n1 + gj) * n2 + gk;
#pragma unroll
            for (int d = 0; d < NDOF; d++) {
                const int r = in ? n2g[nid * NDOF + d] : -1;
                srow[d][h] = r;
                sp[d][h] = (r != -1) ? x[r] : 0.0;
            }
        }
        __syncthreads();
        // phase 2
        const int i = i0 + ti, j = j0 + tj, k = k0 + tk;
        if (i < n0 && j < n1 && k < n2) {
            double acc[NDOF];
#pragma unroll
            for (int d = 0; d < NDOF; d++) acc[d] = 0.0;
#pragma unroll
            for (int a = 0; a < NC; a++) {
                int ox, oy, oz;
                mf_corner(DIM, a, ox, oy, oz);
                const int ei = i - ox, ej = j - oy, ek = k - oz;
                const bool in = (unsigned)ei < (unsigned)(n0 - 1) && (unsigned)ej < (unsigned)e1 && (DIM == 2 || (unsigned)ek < (unsigned)e2);
                if (!in) continue;
                const double Ee = __ldg(E + ((size_t)ei * e1 + ej) * e2 + ek);
                double t[NDOF];
#pragma unroll
                for (int d = 0; d < NDOF; d++) t[d] = 0.0;
#pragma unroll
                for (int b = 0; b < NC; b++) {
                    int bx, by, bz;
                    mf_corner(DIM, b, bx, by, bz);
                    const int q = hc + ((bx - ox) * HJ + (by - oy)) * HK + ((DIM == 3) ? (bz - oz) : 0);
#pragma unroll
                    for (int dj = 0; dj < NDOF; dj++) {
                        const double pv = sp[dj][q];
#pragma unroll
                        for (int di = 0; di < NDOF; di++) t[di] += c_mf_ke0[(a * NDOF + di) * M + b * NDOF + dj] * pv;
                    }
                }
#pragma unroll
                for (int d = 0; d < NDOF; d++) acc[d] += Ee * t[d];
            }
#pragma unroll
            for (int d = 0; d < NDOF; d++) {
                const int r = srow[d][hc];
                if (r == -1) continue;
                y[r] = acc[d];
                if (DOT && r >= dot_lo && r < dot_hi) dot += acc[d] * sp[d][hc];
            }
        }
        __syncthreads();
    }
    if (DOT) {
        double vsum[1] = { dot };
        if (grid_sum_last<1>(vsum, partials, ticket)) finish_dot(vsum[0], dot_out, p2p, p2p_epoch);
    }
}

// ---- nodal-space variant, fused with the direction update ---------------------------------------------------------------
// The single-GPU PCG can run in NODAL numbering (vectors of nnode*NDOF entries, fixed dofs carried as identity rows with zero
// right-hand side -- they stay 0 through every recurrence), which removes the nodetoglobal indirection from the gathers, and
// the direction update p = beta p + z (CG.h:439) can then be folded into the staging phase: the tile + halo entries of p are
// rebuilt from p_old and z while they are loaded, the tile's own entries are written to the OTHER p buffer (ping-pong, so halo
// readers of neighbouring tiles still see p_old), and the separate p-update kernel disappears.
template <int DIM, int NDOF>
__global__ void __launch_bounds__(kThreads)
spmv_mf_nodal_kernel(MfGrid G, const int* __restrict__ n2g, const double* __restrict__ E, const double* __restrict__ p_old, const double* __restrict__ z,
                     double* __restrict__ p_new, double* __restrict__ y, CgState* __restrict__ st, double* partials, unsigned int* ticket) {
    if (st->done) return;
    constexpr int NC = 1 << DIM, M = NC * NDOF;
    constexpr int TI = MfTile<DIM>::TI, TJ = MfTile<DIM>::TJ, TK = MfTile<DIM>::TK;
    constexpr int HJ = TJ + 2, HK = (DIM == 3) ? TK + 2 : 1, HI = TI + 2, HALO = HI * HJ * HK;
    __shared__ double sp[NDOF][HALO];
    const double beta = st->beta;
    const int n0 = G.n[0], n1 = G.n[1], n2 = G.n[2];
    const int e1 = n1 - 1, e2 = (DIM == 3) ? n2 - 1 : 1;
    const int tiles_i = (n0 + TI - 1) / TI, tiles_j = (n1 + TJ - 1) / TJ, tiles_k = (DIM == 3) ? (n2 + TK - 1) / TK : 1;
    const int ntiles = tiles_i * tiles_j * tiles_k;
    const int tk = threadIdx.x % TK, tj = (threadIdx.x / TK) % TJ, ti = threadIdx.x / (TK * TJ);
    const int hc = ((ti + 1) * HJ + (tj + 1)) * HK + ((DIM == 3) ? tk + 1 : 0);
    double dot = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int bk = tile % tiles_k, bj = (tile / tiles_k) % tiles_j, bi = tile / (tiles_k * tiles_j);
        const int i0 = bi * TI, j0 = bj * TJ, k0 = bk * TK;
        for (int h = threadIdx.x; h < HALO; h += kThreads) {
            const int hk = h % HK, hj = (h / HK) % HJ, hi = h / (HK * HJ);
            const int gi = i0 - 1 + hi, gj = j0 - 1 + hj, gk = (DIM == 3) ? k0 - 1 + hk : 0;
            const bool in = (unsigned)gi < (unsigned)n0 && (unsigned)gj < (unsigned)n1 && (unsigned)gk < (unsigned)n2;
            const size_t nid = ((size_t)gi * n1 + gj) * n2 + gk;
            const bool own = hi >= 1 && hi <= TI && hj >= 1 && hj <= TJ && (DIM == 2 || (hk >= 1 && hk <= TK));
#pragma unroll
            for (int d = 0; d < NDOF; d++) {
                double v = 0.0;
                if (in) {
                    v = beta * p_old[nid * NDOF + d] + z[nid * NDOF + d];          // xeaxpy (CG.h:41-49)
                    if (own) p_new[nid * NDOF + d] = v;
                }
                sp[d][h] = v;
            }
        }
        __syncthreads();
        const int i = i0 + ti, j = j0 + tj, k = k0 + tk;
        if (i < n0 && j < n1 && k < n2) {
            const size_t nid = ((size_t)i * n1 + j) * n2 + k;
            double acc[NDOF];
#pragma unroll
            for (int d = 0; d < NDOF; d++) acc[d] = 0.0;
#pragma unroll
            for (int a = 0; a < NC; a++) {
                int ox, oy, oz;
                mf_corner(DIM, a, ox, oy, oz);
                const int ei = i - ox, ej = j - oy, ek = k - oz;
                const bool in = (unsigned)ei < (unsigned)(n0 - 1) && (unsigned)ej < (unsigned)e1 && (DIM == 2 || (unsigned)ek < (unsigned)e2);
                if (!in) continue;
                const double Ee = __ldg(E + ((size_t)ei * e1 + ej) * e2 + ek);
                double t[NDOF];
#pragma unroll
                for (int d = 0; d < NDOF; d++) t[d] = 0.0;
#pragma unroll
                for (int b = 0; b < NC; b++) {
                    int bx, by, bz;
                    mf_corner(DIM, b, bx, by, bz);
                    const int q = hc + ((bx - ox) * HJ + (by - oy)) * HK + ((DIM == 3) ? (bz - oz) : 0);
#pragma unroll
                    for (int dj = 0; dj < NDOF; dj++) {
                        const double pv = sp[dj][q];
#pragma unroll
                        for (int di = 0; di < NDOF; di++) t[di] += c_mf_ke0[(a * NDOF + di) * M + b * NDOF + dj] * pv;
                    }
                }
#pragma unroll
                for (int d = 0; d < NDOF; d++) acc[d] += Ee * t[d];
            }
#pragma unroll
            for (int d = 0; d < NDOF; d++) {
                const bool fixed = n2g[nid * NDOF + d] == -1;              // identity row: y = p = 0
                const double yv = fixed ? 0.0 : acc[d];
                y[nid * NDOF + d] = yv;
                dot += yv * sp[d][hc];
            }
        }
        __syncthreads();
    }
    double vsum[1] = { dot };
    if (grid_sum_last<1>(vsum, partials, ticket) && threadIdx.x == 0) st->pAp = vsum[0];
}

// reduced <-> nodal numbering
__global__ void mf_expand_kernel(size_t nfull, const int* __restrict__ n2g, const double* __restrict__ b, const long long* __restrict__ indptr,
                                 const int* __restrict__ diagpos, const double* __restrict__ data, int jacobi, double* __restrict__ bn, double* __restrict__ dn) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nfull; i += (size_t)gridDim.x * blockDim.x) {
        const int r = n2g[i];
        bn[i] = (r != -1) ? b[r] : 0.0;
        double d = 1.0;
        if (jacobi && r != -1) { const int dp = diagpos[r]; d = dp >= 0 ? data[indptr[r] + dp] : 0.0; }      // GetDiagonal (CG.h:398-404)
        dn[i] = d;
    }
}
__global__ void mf_gather_kernel(size_t nfull, const int* __restrict__ n2g, const double* __restrict__ xn, double* __restrict__ x) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nfull; i += (size_t)gridDim.x * blockDim.x) {
        const int r = n2g[i];
        if (r != -1) x[r] = xn[i];
    }
}
// warm start: the initial guess in nodal numbering (fixed dofs are identity rows that stay 0)
__global__ void mf_expand_x_kernel(size_t nfull, const int* __restrict__ n2g, const double* __restrict__ x, double* __restrict__ xn) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nfull; i += (size_t)gridDim.x * blockDim.x) {
        const int r = n2g[i];
        xn[i] = (r != -1) ? x[r] : 0.0;
    }
}
// x = x0 (0 without y0) ; r = b - K x0 ; z = r / D ; p = z ; bb ; rho = z.r ; rr      (CG.h:422-428 in nodal numbering; both p buffers
// start as z).  y0 = K x0 of a warm start, nullptr: the reference's x0 = 0.
__global__ void __launch_bounds__(kThreads)
mf_init_kernel(size_t nfull, const double* __restrict__ bn, const double* __restrict__ dn, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
               double* __restrict__ p0, double* __restrict__ p1, CgState* st, int maxit, double eps, double* partials, unsigned int* ticket,
               const double* __restrict__ y0) {
    double v[3] = { 0.0, 0.0, 0.0 };
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nfull; i += (size_t)gridDim.x * blockDim.x) {
        const double bi = bn[i];
        double ri = bi;
        if (y0) ri = bi - y0[i]; else x[i] = 0.0;
        const double zi = ri / dn[i];
        r[i] = ri; z[i] = zi; p0[i] = zi; p1[i] = zi;
        v[0] += bi * bi; v[1] += zi * ri; v[2] += ri * ri;
    }
    if (grid_sum_last<3>(v, partials, ticket) && threadIdx.x == 0) {
        st->bb = v[0]; st->rr = v[2]; st->rho = v[1]; st->pAp = 0.0; st->beta = 0.0; st->zr_new = 0.0;
        st->iter = 0; st->maxit = maxit; st->eps = eps;
        st->done = (y0 != nullptr && sqrt(v[2]) < eps * sqrt(v[0])) ? 1 : 0;
    }
}

// lattice check: connectivity follows the x-major numbering and every element is a translate of element 0
__global__ void mf_verify_kernel(int dim, MfGrid G, int nelem, const int* __restrict__ conn, const double* __restrict__ coords, int* bad) {
    const int npe = 1 << dim;
    const int n1 = G.n[1], n2 = G.n[2], s0 = n1 * n2, s1 = n2;
    const int e1 = n1 - 1, e2 = (dim == 3) ? n2 - 1 : 1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        const int ek = e % e2, ej = (e / e2) % e1, ei = e / (e1 * e2);
        const int origin = ei * s0 + ej * s1 + ek;
        bool ok = true;
        for (int a = 0; a < npe; a++) {
            int ox, oy, oz;
            mf_corner(dim, a, ox, oy, oz);
            const int node = origin + ox * s0 + oy * s1 + oz;
            ok = ok && (conn[(size_t)e * npe + a] == node);
            if (!ok) break;
            for (int c = 0; c < dim; c++) {
                // edge vectors relative to the element's first node must equal those of element 0
                const double d0 = coords[(size_t)conn[a] * dim + c] - coords[(size_t)conn[0] * dim + c];
                const double de = coords[(size_t)node * dim + c] - coords[(size_t)origin * dim + c];
                const double scale = fabs(coords[(size_t)conn[npe == 4 ? 2 : 6] * dim + c] - coords[(size_t)conn[0] * dim + c]) + 1.0e-300;
                ok = ok && (fabs(d0 - de) <= 1.0e-9 * scale);
            }
        }
        if (!ok) atomicExch(bad, 1);
    }
}

__global__ void mf_modulus_kernel(int nelem, const double* __restrict__ modulus, const double* __restrict__ rho, double E0, double E1, double p,
                                  double* __restrict__ E) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        if (modulus) E[e] = modulus[e];
        else { const double rp = pow(rho[e], p); E[e] = E1 * rp + E0 * (1.0 - rp); }
    }
}

}  // namespace pf2
