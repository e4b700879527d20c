// assemble_generic.cu -- assembly / sensitivity / single-element kernels for the element families beyond the three
// specialised selections, and the eq-code decoder shared by every entry point.
//
//   decode_eq                 : PF2_EQ_CODE(phys, shape, quad, quad2) -> kind, shape, rules, D coefficients
//   assemble_generic_kernel   : element routine <Equation, SF, IC> + Assembling(K,F,u,Ke,...) (Assembling.h:47-66)
//   sens_generic_kernel       : reaction / compliance / sensitivity pass of the TO drivers for the same selections
//   element_generic_kernel    : one element matrix (the reference's per-element call)
// Same thread mapping as assemble.cu: one thread per (element, local node) holding that node's NDOF rows of Ke.
#include "types.cuh"
#include "element_generic.cuh"
#include "assemble_gather.cuh"

namespace pf2 {

static int shape_dim(int shape) { return (shape == PF2_SHAPE_TET4 || shape == PF2_SHAPE_HEX8 || shape == PF2_SHAPE_HEX20) ? 3 : 2; }
static int shape_npe(int shape) {
    switch (shape) {
        case PF2_SHAPE_T3: return 3; case PF2_SHAPE_T6: return 6; case PF2_SHAPE_Q4: return 4; case PF2_SHAPE_Q8: return 8;
        case PF2_SHAPE_TET4: return 4; case PF2_SHAPE_HEX8: return 8; case PF2_SHAPE_HEX20: return 20;
    }
    return 0;
}
// reference domain: 0 triangle, 1 square, 2 tetrahedron, 3 cube
static int shape_domain(int shape) {
    switch (shape) {
        case PF2_SHAPE_T3: case PF2_SHAPE_T6: return 0;
        case PF2_SHAPE_Q4: case PF2_SHAPE_Q8: return 1;
        case PF2_SHAPE_TET4: return 2;
        default: return 3;
    }
}
static int quad_domain(int quad) {
    switch (quad) {
        case PF2_QUAD_G1TRI: case PF2_QUAD_G3TRI: return 0;
        case PF2_QUAD_G1SQ: case PF2_QUAD_G4SQ: case PF2_QUAD_G9SQ: return 1;
        case PF2_QUAD_G1TET: return 2;
        case PF2_QUAD_G8CUBE: case PF2_QUAD_G27CUBE: return 3;
    }
    return -1;
}

int decode_eq(int eq, double V, EqInfo* out) {
    EqInfo q;
    q.phys = eq & 0xff; q.shape = (eq >> 8) & 0xff; q.quad = (eq >> 16) & 0xff; q.quad2 = (eq >> 24) & 0xff;
    PF2_CHECK(eq >= 0 && q.phys <= PF2_PHYS_PLANE_D_WT, "unknown equation");
    const bool gend = q.phys >= PF2_PHYS_PLANE_D;     // caller-supplied constitutive matrix
    const bool adv = q.phys == PF2_PHYS_ADVDIFF;     // quad2 carries the PF2_ADV_* mask
    PF2_CHECK(q.shape <= PF2_SHAPE_HEX20 && q.quad <= PF2_QUAD_G27CUBE && (adv || q.quad2 <= PF2_QUAD_G27CUBE), "unknown shape function / integration rule");
    PF2_CHECK(!adv || (q.quad2 >= 1 && q.quad2 <= 63), "advection-diffusion: the quad2 field must select at least one routine (PF2_ADV_*)");
    const bool solid = q.phys == PF2_PHYS_SOLID;
    if (q.shape == 0) q.shape = solid ? PF2_SHAPE_HEX8 : PF2_SHAPE_Q4;
    PF2_CHECK(shape_dim(q.shape) == (solid ? 3 : 2), "shape function does not match the equation's dimension");
    const int dom = shape_domain(q.shape);
    static const int dflt[4] = { PF2_QUAD_G1TRI, PF2_QUAD_G4SQ, PF2_QUAD_G1TET, PF2_QUAD_G8CUBE };
    static const int dflt_reduced[4] = { PF2_QUAD_G1TRI, PF2_QUAD_G1SQ, PF2_QUAD_G1TET, PF2_QUAD_G8CUBE };
    if (q.quad == 0) q.quad = dflt[dom];
    PF2_CHECK(quad_domain(q.quad) == dom, "integration rule does not belong to the shape function's reference domain");
    if (q.phys == PF2_PHYS_PLANESTRAIN_SRI || q.phys == PF2_PHYS_PLANESTRAIN_BBAR || q.phys == PF2_PHYS_PLANE_D_BBAR) {
        if (q.quad2 == 0) q.quad2 = dflt_reduced[dom];
        PF2_CHECK(quad_domain(q.quad2) == dom, "volumetric integration rule does not belong to the shape function's reference domain");
    } else if (!adv) {
        PF2_CHECK(q.quad2 == 0, "a second integration rule is only meaningful for the selective-reduced variant");
    }
    q.dim = solid ? 3 : 2;
    q.npe = shape_npe(q.shape);
    q.ndof = solid ? 3 : ((q.phys == PF2_PHYS_HEAT || q.phys == PF2_PHYS_MASS || adv) ? 1 : 2);
    q.kind = gend ? KIND_ELAST2D_D : adv ? KIND_ADVDIFF2D : solid ? KIND_SOLID3D : (q.phys == PF2_PHYS_HEAT ? KIND_HEAT2D : (q.phys == PF2_PHYS_MASS ? KIND_MASS2D : (q.phys == PF2_PHYS_MASS2 ? KIND_MASS2D_V : KIND_ELAST2D)));
    q.fast = (q.phys == PF2_PHYS_PLANESTRAIN && q.shape == PF2_SHAPE_Q4 && q.quad == PF2_QUAD_G4SQ) ||
             (q.phys == PF2_PHYS_HEAT && q.shape == PF2_SHAPE_Q4 && q.quad == PF2_QUAD_G4SQ) ||
             (solid && q.shape == PF2_SHAPE_HEX8 && q.quad == PF2_QUAD_G8CUBE);
    q.legacy = solid ? PF2_EQ_SOLID : (q.phys == PF2_PHYS_HEAT ? PF2_EQ_HEAT : PF2_EQ_PLANESTRAIN);
    // D for unit modulus
    q.npass = 1; q.cn[0] = q.cn[1] = 1.0; q.lam[0] = q.lam[1] = 0.0; q.mu[0] = q.mu[1] = 0.0;
    if (q.phys == PF2_PHYS_PLANESTRAIN_WT || q.phys == PF2_PHYS_PLANE_D_WT)
        PF2_CHECK(dom == 1 && q.quad != PF2_QUAD_G1SQ, "Wilson-Taylor: quadrilateral shapes with at least Gauss4Square (the modes vanish at the centre point)");
    if (q.phys == PF2_PHYS_PLANESTRAIN || q.phys == PF2_PHYS_PLANESTRAIN_WT || solid) {          // PlaneStrain.h:37-41, Solid.h:37-44
        const double c = 1.0 / ((1.0 + V) * (1.0 - 2.0 * V));
        q.cn[0] = (1.0 - V) * c; q.lam[0] = V * c; q.mu[0] = 0.5 * (1.0 - 2.0 * V) * c;
    } else if (q.phys == PF2_PHYS_PLANESTRESS) {            // PlaneStress.h:37-41
        const double c = 1.0 / ((1.0 - V) * (1.0 + V));
        q.cn[0] = c; q.lam[0] = V * c; q.mu[0] = 0.5 * (1.0 - V) * c;
    } else if (q.phys == PF2_PHYS_PLANESTRAIN_SRI) {        // PlaneStrain.h:79-83 (volumetric, ICV), :101-105 (deviatoric, ICD)
        q.npass = 2;
        const double k = 1.0 / (3.0 * (1.0 - 2.0 * V)), c = 1.0 / (6.0 * (1.0 + V));
        q.cn[0] = k; q.lam[0] = k; q.mu[0] = 0.0;
        q.cn[1] = 4.0 * c; q.lam[1] = -2.0 * c; q.mu[1] = 3.0 * c;
    } else if (q.phys == PF2_PHYS_PLANESTRAIN_BBAR) {       // PlaneStrain.h:143-175: B = Bvol (ICV) + Bdev (ICD), plane-strain D on both
        // Bvol^T D Bvol = (D00 + 2 D01 + D11)/4 (g_a g_b^T) ; Bdev^T D Bdev = (D00 - 2 D01 + D11)/4 [gx gx, -gx gy; -gy gx, gy gy] + D22 shear
        q.npass = 2;
        const double c = 1.0 / ((1.0 + V) * (1.0 - 2.0 * V));
        const double m = 0.5 * (1.0 - 2.0 * V) * c;
        q.cn[0] = 0.5 * c; q.lam[0] = 0.5 * c; q.mu[0] = 0.0;
        q.cn[1] = m; q.lam[1] = -m; q.mu[1] = m;
    }
    *out = q;
    return PF2_OK;
}

ElemSpec make_spec(const EqInfo& q) {
    ElemSpec sp;
    sp.npass = q.npass;
    sp.wilson_taylor = (q.phys == PF2_PHYS_PLANESTRAIN_WT) ? 1 : 0;
    if (q.npass == 2) { sp.quad[0] = q.quad2; sp.quad[1] = q.quad; }
    else { sp.quad[0] = q.quad; sp.quad[1] = q.quad; }
    for (int i = 0; i < 2; i++) { sp.cn[i] = q.cn[i]; sp.lam[i] = q.lam[i]; sp.mu[i] = q.mu[i]; }
    return sp;
}

template <int KIND, int SHAPE>
__global__ void __launch_bounds__(128)
assemble_generic_kernel(int nelem, ElemSpec sp, const double* __restrict__ coords, const int* __restrict__ conn, const int* __restrict__ n2g,
                        const double* __restrict__ ufix, const int* __restrict__ bmap, const long long* __restrict__ indptr,
                        const double* __restrict__ modulus, const double* __restrict__ rho, double E0, double E1, double p,
                        double t, double* __restrict__ data, double* __restrict__ F) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE, NDOF = KindTraits<KIND>::NDOF;
    const long long total = (long long)nelem * NPE;
    for (long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x; tid < total; tid += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(tid / NPE), a = (int)(tid % NPE);
        const int* nd = conn + (size_t)e * NPE;
        const int na = nd[a];
        int rows[NDOF];
        bool any = false;
#pragma unroll
        for (int i = 0; i < NDOF; i++) { rows[i] = n2g[(size_t)na * NDOF + i]; any |= (rows[i] != -1); }
        if (!any) continue;
        double X[NPE][DIM];
#pragma unroll
        for (int n = 0; n < NPE; n++)
#pragma unroll
            for (int k = 0; k < DIM; k++) X[n][k] = coords[(size_t)nd[n] * DIM + k];
        const double E = modulus ? modulus[e] : simp_modulus(rho[e], E0, E1, p);
        double acc[NDOF][NPE * NDOF];
        generic_rows<KIND, SHAPE>(X, a, sp, t, acc);
        const int* bm = bmap + ((size_t)e * NPE + a) * NPE;
#pragma unroll (NPE <= 8 ? NPE : 1)
        for (int b = 0; b < NPE; b++) {
            const int off = bm[b], nb = nd[b];
            int cfree[NDOF];
            int rank = 0;
#pragma unroll
            for (int j = 0; j < NDOF; j++) {
                const int c = n2g[(size_t)nb * NDOF + j];
                cfree[j] = (c != -1) ? rank++ : -1;
            }
#pragma unroll
            for (int i = 0; i < NDOF; i++) {
                if (rows[i] == -1) continue;
                const long long base = indptr[rows[i]] + off;
#pragma unroll
                for (int j = 0; j < NDOF; j++) {
                    const double v = E * acc[i][b * NDOF + j];
                    if (cfree[j] >= 0) atomicAdd(&data[base + cfree[j]], v);                  // Assembling.h:55
                    else {
                        const double uf = ufix[(size_t)nb * NDOF + j];
                        if (uf != 0.0) atomicAdd(&F[rows[i]], -(v * uf));                       // Assembling.h:59
                    }
                }
            }
        }
    }
}

// Hex20 solid (Solid.h:30-72 with ShapeFunction20Cubic, Gauss27Cubic): one WARP per element.  The thread-per-(element, node) kernel above
// carries 3 x 60 accumulators and evaluates all 27 gradient sets in every one of its 20 threads -- it spills (r01: 13.6 ms per 27.6 k
// elements).  Here lane q evaluates the gradients of integration point q ONCE (the same shape_grad, so the same bits) into shared memory,
// then lane a accumulates the three rows of node a against five column nodes at a time (45 accumulators, in registers), reading the
// gradients back as broadcasts, in the same order over the points as the reference's loop -- the entries equal the generic kernel's.
constexpr int kHex20Warps = 3, kHex20Chunk = 5;
__global__ void __launch_bounds__(kHex20Warps * 32)
assemble_hex20_warp_kernel(int nelem, ElemSpec sp, const double* __restrict__ coords, const int* __restrict__ conn, const int* __restrict__ n2g,
                           const double* __restrict__ ufix, const int* __restrict__ bmap, const long long* __restrict__ indptr,
                           const double* __restrict__ modulus, const double* __restrict__ rho, double E0, double E1, double p,
                           double* __restrict__ data, double* __restrict__ F) {
    constexpr int NPE = 20, DIM = 3, NDOF = 3, MAXQ = 27, CH = kHex20Chunk;
    __shared__ double Xs[kHex20Warps][NPE][DIM];
    __shared__ double Gs[kHex20Warps][MAXQ][DIM][NPE];
    __shared__ double Ws[kHex20Warps][MAXQ];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ng = quad_count(sp.quad[0]);
    const double cn = sp.cn[0], lam = sp.lam[0], mu = sp.mu[0];
    for (int e = blockIdx.x * kHex20Warps + wid; e < nelem; e += gridDim.x * kHex20Warps) {
        const int* nd = conn + (size_t)e * NPE;
        __syncwarp();                                   // the previous element's readers are done with this warp's tiles
        if (lane < NPE) {
            const int node = nd[lane];
#pragma unroll
            for (int k = 0; k < DIM; k++) Xs[wid][lane][k] = coords[(size_t)node * DIM + k];
        }
        __syncwarp();
        if (lane < ng) {
            double r[3], wq, det, g[DIM][NPE];
            quad_point(sp.quad[0], lane, r, wq);
            shape_grad<SH_HEX20>(Xs[wid], r, g, det);
#pragma unroll
            for (int k = 0; k < DIM; k++)
#pragma unroll
                for (int n = 0; n < NPE; n++) Gs[wid][lane][k][n] = g[k][n];
            Ws[wid][lane] = det * wq;
        }
        __syncwarp();
        if (lane >= NPE) continue;
        const int a = lane, na = nd[a];
        int rows[NDOF];
        bool any = false;
#pragma unroll
        for (int i = 0; i < NDOF; i++) { rows[i] = n2g[(size_t)na * NDOF + i]; any |= (rows[i] != -1); }
        if (!any) continue;
        const double E = modulus ? modulus[e] : simp_modulus(rho[e], E0, E1, p);
        const int* bm = bmap + ((size_t)e * NPE + a) * NPE;
#pragma unroll 1
        for (int c0 = 0; c0 < NPE; c0 += CH) {
            double acc[NDOF][CH * NDOF];
#pragma unroll
            for (int i = 0; i < NDOF; i++)
#pragma unroll
                for (int j = 0; j < CH * NDOF; j++) acc[i][j] = 0.0;
#pragma unroll 1
            for (int q = 0; q < ng; q++) {
                const double w = Ws[wid][q];
                double cg[DIM], lg[DIM], mg[DIM];
#pragma unroll
                for (int k = 0; k < DIM; k++) { const double ga = Gs[wid][q][k][a]; cg[k] = cn * ga * w; lg[k] = lam * ga * w; mg[k] = mu * ga * w; }
#pragma unroll
                for (int bb = 0; bb < CH; bb++) {
                    double g[DIM];
#pragma unroll
                    for (int k = 0; k < DIM; k++) g[k] = Gs[wid][q][k][c0 + bb];
#pragma unroll
                    for (int i = 0; i < NDOF; i++)
#pragma unroll
                        for (int j = 0; j < NDOF; j++) {
                            if (i == j) {
                                acc[i][bb * NDOF + j] += cg[i] * g[i];
#pragma unroll
                                for (int k = 0; k < DIM; k++) if (k != i) acc[i][bb * NDOF + j] += mg[k] * g[k];
                            } else { acc[i][bb * NDOF + j] += lg[i] * g[j]; acc[i][bb * NDOF + j] += mg[j] * g[i]; }
                        }
                }
            }
#pragma unroll
            for (int bb = 0; bb < CH; bb++) {
                const int b = c0 + bb, off = bm[b], nb = nd[b];
                int cfree[NDOF];
                int rank = 0;
#pragma unroll
                for (int j = 0; j < NDOF; j++) {
                    const int c = n2g[(size_t)nb * NDOF + j];
                    cfree[j] = (c != -1) ? rank++ : -1;
                }
#pragma unroll
                for (int i = 0; i < NDOF; i++) {
                    if (rows[i] == -1) continue;
                    const long long base = indptr[rows[i]] + off;
#pragma unroll
                    for (int j = 0; j < NDOF; j++) {
                        const double v = E * acc[i][bb * NDOF + j];
                        if (cfree[j] >= 0) atomicAdd(&data[base + cfree[j]], v);                  // Assembling.h:55
                        else {
                            const double uf = ufix[(size_t)nb * NDOF + j];
                            if (uf != 0.0) atomicAdd(&F[rows[i]], -(v * uf));                       // Assembling.h:59
                        }
                    }
                }
            }
        }
    }
}

template <int KIND, int SHAPE>
__global__ void __launch_bounds__(kThreads)
sens_generic_kernel(int nelem, ElemSpec sp, const double* __restrict__ coords, const int* __restrict__ conn, const double* __restrict__ u,
                    const double* __restrict__ rho, double E0, double E1, double p, double t, double scale0,
                    double* __restrict__ dfdrho, double* r_nodal, double* f_out, double* partials, unsigned int* ticket, int sum_lo, int sum_hi) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE, NDOF = KindTraits<KIND>::NDOF;
    double fsum = 0.0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        const int* nd = conn + (size_t)e * NPE;
        double X[NPE][DIM], ue[NPE][NDOF], fe[NPE][NDOF];
#pragma unroll
        for (int n = 0; n < NPE; n++) {
            const int node = nd[n];
#pragma unroll
            for (int k = 0; k < DIM; k++) X[n][k] = coords[(size_t)node * DIM + k];
#pragma unroll
            for (int k = 0; k < NDOF; k++) { ue[n][k] = u[(size_t)node * NDOF + k]; fe[n][k] = 0.0; }
        }
        const double w = r_nodal ? generic_energy<KIND, SHAPE, true>(X, ue, sp, t, fe) : generic_energy<KIND, SHAPE, false>(X, ue, sp, t, fe);
        const double rh = rho[e];
        const double E = simp_modulus(rh, E0, E1, p);
        if (e >= sum_lo && e < sum_hi) fsum += E * w;
        if (dfdrho) dfdrho[e] = -scale0 * p * (-E0 + E1) * pow(rh, p - 1.0) * w;
        if (r_nodal) {
#pragma unroll 1
            for (int n = 0; n < NPE; n++)
#pragma unroll
                for (int k = 0; k < NDOF; k++) atomicAdd(&r_nodal[(size_t)nd[n] * NDOF + k], E * fe[n][k]);
        }
    }
    double v[1] = { fsum };
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) *f_out = scale0 * v[0];
}

template <int KIND, int SHAPE>
__global__ void element_generic_kernel(ElemSpec sp, const double* __restrict__ xe, double E, double t, double* __restrict__ Ke) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE, NDOF = KindTraits<KIND>::NDOF;
    const int a = threadIdx.x;
    if (a >= NPE) return;
    double X[NPE][DIM];
    for (int n = 0; n < NPE; n++) for (int k = 0; k < DIM; k++) X[n][k] = xe[n * DIM + k];
    double acc[NDOF][NPE * NDOF];
    generic_rows<KIND, SHAPE>(X, a, sp, t, acc);
    constexpr int M = NPE * NDOF;
    for (int i = 0; i < NDOF; i++) for (int j = 0; j < M; j++) Ke[(a * NDOF + i) * M + j] = E * acc[i][j];
}

// the generic template as a row provider of the gather kernel (assemble_gather.cuh)
template <int KIND, int SHAPE>
struct ElemGeneric {
    static constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE, NDOF = KindTraits<KIND>::NDOF;
    ElemSpec sp;
    double t;
    __device__ __forceinline__ void rows(const double (&X)[NPE][DIM], int a, double (&acc)[NDOF][NPE * NDOF]) const { generic_rows<KIND, SHAPE>(X, a, sp, t, acc); }
};

// one element with a caller-supplied constitutive matrix (PlaneStiffness*, Homogenization.h:141-280)
template <int SHAPE>
__global__ void element_general_kernel(ElemSpecD sp, const double* __restrict__ xe, double t, double* __restrict__ Ke) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE, M = NPE * 2;
    const int a = threadIdx.x;
    if (a >= NPE) return;
    double X[NPE][2];
    for (int n = 0; n < NPE; n++) { X[n][0] = xe[n * 2]; X[n][1] = xe[n * 2 + 1]; }
    double acc[2][M];
    general_rows<SHAPE>(X, a, sp, t, acc);
    for (int i = 0; i < 2; i++) for (int j = 0; j < M; j++) Ke[(a * 2 + i) * M + j] = acc[i][j];
}

ElemSpecD make_spec_d(const EqInfo& q, const double D[9]) {
    ElemSpecD sp;
    sp.mode = (q.phys == PF2_PHYS_PLANE_D_BBAR) ? 1 : ((q.phys == PF2_PHYS_PLANE_D_WT) ? 2 : 0);
    if (sp.mode == 1) { sp.quad[0] = q.quad2; sp.quad[1] = q.quad; }       // volumetric rule first, as PlaneStiffnessBbar integrates
    else { sp.quad[0] = q.quad; sp.quad[1] = q.quad; }
    for (int i = 0; i < 9; i++) sp.D[i] = D[i];
    return sp;
}

int element_general_launch(pf2_ctx* ctx, const EqInfo& q, const double* xe_dev, const double D[9], double t, double* Ke_dev) {
    const ElemSpecD sp = make_spec_d(q, D);
    if (q.shape == PF2_SHAPE_T3) element_general_kernel<SH_T3><<<1, 32, 0, ctx->stream>>>(sp, xe_dev, t, Ke_dev);
    else if (q.shape == PF2_SHAPE_T6) element_general_kernel<SH_T6><<<1, 32, 0, ctx->stream>>>(sp, xe_dev, t, Ke_dev);
    else if (q.shape == PF2_SHAPE_Q4) element_general_kernel<SH_Q4><<<1, 32, 0, ctx->stream>>>(sp, xe_dev, t, Ke_dev);
    else element_general_kernel<SH_Q8><<<1, 32, 0, ctx->stream>>>(sp, xe_dev, t, Ke_dev);
    PF2_LAUNCH_CHECK();
    ctx->launches++;
    return PF2_OK;
}

// (kind, shape) -> instantiation
#define PF2_DISPATCH_SHAPE(q, CALL)                                                                  \
    do {                                                                                             \
        if ((q).kind == KIND_SOLID3D) {                                                              \
            if ((q).shape == PF2_SHAPE_TET4) { CALL(KIND_SOLID3D, SH_TET4); }                        \
            else if ((q).shape == PF2_SHAPE_HEX8) { CALL(KIND_SOLID3D, SH_HEX8); }                   \
            else { CALL(KIND_SOLID3D, SH_HEX20); }                                                   \
        } else if ((q).kind == KIND_HEAT2D) {                                                        \
            if ((q).shape == PF2_SHAPE_T3) { CALL(KIND_HEAT2D, SH_T3); }                             \
            else if ((q).shape == PF2_SHAPE_T6) { CALL(KIND_HEAT2D, SH_T6); }                        \
            else if ((q).shape == PF2_SHAPE_Q4) { CALL(KIND_HEAT2D, SH_Q4); }                        \
            else { CALL(KIND_HEAT2D, SH_Q8); }                                                       \
        } else if ((q).kind == KIND_MASS2D_V) {                                                      \
            if ((q).shape == PF2_SHAPE_T3) { CALL(KIND_MASS2D_V, SH_T3); }                           \
            else if ((q).shape == PF2_SHAPE_T6) { CALL(KIND_MASS2D_V, SH_T6); }                      \
            else if ((q).shape == PF2_SHAPE_Q4) { CALL(KIND_MASS2D_V, SH_Q4); }                      \
            else { CALL(KIND_MASS2D_V, SH_Q8); }                                                     \
        } else if ((q).kind == KIND_MASS2D) {                                                        \
            if ((q).shape == PF2_SHAPE_T3) { CALL(KIND_MASS2D, SH_T3); }                             \
            else if ((q).shape == PF2_SHAPE_T6) { CALL(KIND_MASS2D, SH_T6); }                        \
            else if ((q).shape == PF2_SHAPE_Q4) { CALL(KIND_MASS2D, SH_Q4); }                        \
            else { CALL(KIND_MASS2D, SH_Q8); }                                                       \
        } else {                                                                                     \
            if ((q).shape == PF2_SHAPE_T3) { CALL(KIND_ELAST2D, SH_T3); }                            \
            else if ((q).shape == PF2_SHAPE_T6) { CALL(KIND_ELAST2D, SH_T6); }                       \
            else if ((q).shape == PF2_SHAPE_Q4) { CALL(KIND_ELAST2D, SH_Q4); }                       \
            else { CALL(KIND_ELAST2D, SH_Q8); }                                                      \
        }                                                                                            \
    } while (0)

int assemble_generic_launch(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, const EqInfo& q, const double* modulus_dev, const double* rho_dev,
                            const double params[5]) {
    pf2_ctx* c = A->ctx;
    cudaStream_t s = c->stream;
    const ElemSpec sp = make_spec(q);
    const double E0 = params[0], E1 = params[1], p = params[3], t = params[4];
    if (q.dim == 2 && gather_usable(A, mesh)) {         // the caller skipped the memsets on the same condition
#define CALL(K, S) PF2_TRY((assemble_gather_launch(A, mesh, map, ElemGeneric<K, S>{ sp, t }, modulus_dev, rho_dev, E0, E1, p)))
        if (q.kind == KIND_HEAT2D) {
            if (q.shape == PF2_SHAPE_T3) { CALL(KIND_HEAT2D, SH_T3); } else if (q.shape == PF2_SHAPE_T6) { CALL(KIND_HEAT2D, SH_T6); }
            else if (q.shape == PF2_SHAPE_Q4) { CALL(KIND_HEAT2D, SH_Q4); } else { CALL(KIND_HEAT2D, SH_Q8); }
        } else if (q.kind == KIND_MASS2D_V) {
            if (q.shape == PF2_SHAPE_T3) { CALL(KIND_MASS2D_V, SH_T3); } else if (q.shape == PF2_SHAPE_T6) { CALL(KIND_MASS2D_V, SH_T6); }
            else if (q.shape == PF2_SHAPE_Q4) { CALL(KIND_MASS2D_V, SH_Q4); } else { CALL(KIND_MASS2D_V, SH_Q8); }
        } else if (q.kind == KIND_MASS2D) {
            if (q.shape == PF2_SHAPE_T3) { CALL(KIND_MASS2D, SH_T3); } else if (q.shape == PF2_SHAPE_T6) { CALL(KIND_MASS2D, SH_T6); }
            else if (q.shape == PF2_SHAPE_Q4) { CALL(KIND_MASS2D, SH_Q4); } else { CALL(KIND_MASS2D, SH_Q8); }
        } else {
            if (q.shape == PF2_SHAPE_T3) { CALL(KIND_ELAST2D, SH_T3); } else if (q.shape == PF2_SHAPE_T6) { CALL(KIND_ELAST2D, SH_T6); }
            else if (q.shape == PF2_SHAPE_Q4) { CALL(KIND_ELAST2D, SH_Q4); } else { CALL(KIND_ELAST2D, SH_Q8); }
        }
#undef CALL
        return PF2_OK;
    }
    if (q.kind == KIND_SOLID3D && q.shape == PF2_SHAPE_HEX20 && sp.npass == 1 && quad_count(sp.quad[0]) <= 27) {
        const char* sw = getenv("PF2_HEX20_WARP");      // 0: the thread-per-(element, node) kernel (read per call: tools/hex20_probe.py toggles it)
        const bool warp_kernel = sw == nullptr || atoi(sw) != 0;
        if (warp_kernel) {
            const int grid = std::max(1, std::min((mesh->nelem + kHex20Warps - 1) / kHex20Warps, c->sm_count * 5 * 4));
            assemble_hex20_warp_kernel<<<grid, kHex20Warps * 32, 0, s>>>(mesh->nelem, sp, mesh->coords, mesh->conn, map->n2g, map->ufix, A->bmap, A->indptr,
                                                                         modulus_dev, rho_dev, E0, E1, p, A->data, A->F);
            PF2_LAUNCH_CHECK();
            c->launches++;
            return PF2_OK;
        }
    }
    const long long work = (long long)mesh->nelem * mesh->npe;
    const int grid = (int)std::min<long long>((work + 127) / 128, (long long)c->sm_count * 32);
#define CALL(K, S) assemble_generic_kernel<K, S><<<grid, 128, 0, s>>>(mesh->nelem, sp, mesh->coords, mesh->conn, map->n2g, map->ufix, A->bmap, \
                                                                      A->indptr, modulus_dev, rho_dev, E0, E1, p, t, A->data, A->F)
    PF2_DISPATCH_SHAPE(q, CALL);
#undef CALL
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

int sens_generic_launch(pf2_mesh* mesh, const EqInfo& q, const double* u_nodal, const double* rho, const double params[6], double* f_dev,
                        double* dfdrho, double* r_nodal) {
    pf2_ctx* c = mesh->ctx;
    cudaStream_t s = c->stream;
    const ElemSpec sp = make_spec(q);
    const int grid = c->grid_for(mesh->nelem);
#define CALL(K, S) sens_generic_kernel<K, S><<<grid, kThreads, 0, s>>>(mesh->nelem, sp, mesh->coords, mesh->conn, u_nodal, rho, params[0], params[1], \
                                                                       params[3], params[4], params[5], dfdrho, r_nodal, f_dev, c->red.partials,     \
                                                                       c->red.ticket, mesh->own_elem_lo, mesh->own_elem_hi)
    PF2_DISPATCH_SHAPE(q, CALL);
#undef CALL
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

int element_generic_launch(pf2_ctx* ctx, const EqInfo& q, const double* xe_dev, double E, double t, double* Ke_dev) {
    const ElemSpec sp = make_spec(q);
#define CALL(K, S) element_generic_kernel<K, S><<<1, 32, 0, ctx->stream>>>(sp, xe_dev, E, t, Ke_dev)
    PF2_DISPATCH_SHAPE(q, CALL);
#undef CALL
    PF2_LAUNCH_CHECK();
    ctx->launches++;
    return PF2_OK;
}

}  // namespace pf2

extern "C" int pf2_eq_describe(int eq, int* dim, int* npe, int* ndof) {
    pf2::EqInfo q;
    PF2_TRY(pf2::decode_eq(eq, 0.3, &q));
    if (dim) *dim = q.dim;
    if (npe) *npe = q.npe;
    if (ndof) *ndof = q.ndof;
    return PF2_OK;
}
