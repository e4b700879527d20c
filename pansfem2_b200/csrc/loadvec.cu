// loadvec.cu -- surface / body load vectors of a whole batch of elements on the device.
//
// Replaces the per-element host routines PlaneStrainSurfaceForce / PlaneStrainBodyForce (PlaneStrain.h:421-455, 503-537),
// PlaneStressSurfaceForce / PlaneStressBodyForce (PlaneStress.h:98-167, same arithmetic) and HeatTransferSurfaceFlux
// (HeatTransfer.h:76-98) followed by Assembling(F, Fe, nodetoglobal, nodetoelement, element) (Assembling.h:132-147):
//     Fe[n][i] = sum_g N_n(r_g) f_i(x_g) * m_g * t * w_g ,   x_g = X_e^T N(r_g)
// with m_g = |dX/dr| (edge length density, one weight) on ShapeFunction2Line / 3Line and det(dX/dr) (two weights) on the area shapes.
// The reference calls a functor f(x_g); here the caller evaluates it for every integration point at once (pf2_integration_points gives
// the x_g) and hands the values over, or passes one constant vector.  One thread per element; ~1 KB of traffic per element, HBM-bound.
#include "types.cuh"
#include "element.cuh"
#include "element_generic.cuh"

namespace pf2 {

constexpr int kLoadMaxNpe = 8;

// shape functions and derivatives of any 2-D load carrier: the area shapes of element_generic.cuh and the two line shapes
struct LoadPoint {
    double N[kLoadMaxNpe];
    double d[2][kLoadMaxNpe];     // dN/dr (row 1 unused on lines)
};
__device__ __forceinline__ int load_npe(int shape) {
    switch (shape) {
        case PF2_SHAPE_LINE2: return 2;
        case PF2_SHAPE_LINE3: case PF2_SHAPE_T3: return 3;
        case PF2_SHAPE_T6: return 6;
        case PF2_SHAPE_Q4: return 4;
        default: return 8;
    }
}
__device__ __forceinline__ int load_ngauss(int quad) {
    if (quad == PF2_QUAD_G1LINE) return 1;
    if (quad == PF2_QUAD_G2LINE) return 2;
    return quad_count(quad);
}
template <int SHAPE>
__device__ __forceinline__ void load_fill(const double (&r)[3], LoadPoint& P) {
    double N[ShapeTraits<SHAPE>::NPE], d[2][ShapeTraits<SHAPE>::NPE];
    shape_n<SHAPE>(r, N);
    shape_dndr<SHAPE>(r, d);
#pragma unroll
    for (int n = 0; n < ShapeTraits<SHAPE>::NPE; n++) { P.N[n] = N[n]; P.d[0][n] = d[0][n]; P.d[1][n] = d[1][n]; }
}
// point g of the rule: shape values and the product of the weights the reference multiplies in
__device__ __forceinline__ void load_point(int shape, int quad, int g, LoadPoint& P, double& w) {
    double r[3] = { 0.0, 0.0, 0.0 };
    if (shape == PF2_SHAPE_LINE2 || shape == PF2_SHAPE_LINE3) {
        if (quad == PF2_QUAD_G2LINE) { r[0] = (g == 0 ? -1.0 : 1.0) / sqrt(3.0); w = 1.0; }      // GaussIntegration.h:49-60
        else { r[0] = 0.0; w = 2.0; }                                                             // GaussIntegration.h:27-36
        const double x = r[0];
        if (shape == PF2_SHAPE_LINE2) {                       // ShapeFunction.h:33-44
            P.N[0] = 0.5 * (1 - x); P.N[1] = 0.5 * (1 + x);
            P.d[0][0] = -0.5; P.d[0][1] = 0.5;
        } else {                                              // ShapeFunction.h:61-73
            P.N[0] = -0.5 * (1.0 - x) * x; P.N[1] = 0.5 * x * (1.0 + x); P.N[2] = (1.0 - x) * (1.0 + x);
            P.d[0][0] = -0.5 * (1.0 - 2.0 * x); P.d[0][1] = 0.5 * (1.0 + 2.0 * x); P.d[0][2] = -2.0 * x;
        }
        return;
    }
    quad_point(quad, g, r, w);
    switch (shape) {
        case PF2_SHAPE_T3: load_fill<SH_T3>(r, P); break;
        case PF2_SHAPE_T6: load_fill<SH_T6>(r, P); break;
        case PF2_SHAPE_Q4: load_fill<SH_Q4>(r, P); break;
        default: load_fill<SH_Q8>(r, P); break;
    }
}

__global__ void __launch_bounds__(128)
integration_points_kernel(int nelem, int shape, int quad, const double* __restrict__ coords, const int* __restrict__ conn, double* __restrict__ xg) {
    const int npe = load_npe(shape), ng = load_ngauss(quad);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        double X[kLoadMaxNpe][2];
        for (int n = 0; n < npe; n++) { const int nd = conn[(size_t)e * npe + n]; X[n][0] = coords[2 * (size_t)nd]; X[n][1] = coords[2 * (size_t)nd + 1]; }
        for (int g = 0; g < ng; g++) {
            LoadPoint P;
            double w;
            load_point(shape, quad, g, P, w);
            double x0 = 0.0, x1 = 0.0;
            for (int n = 0; n < npe; n++) { x0 += X[n][0] * P.N[n]; x1 += X[n][1] * P.N[n]; }      // X^T N (PlaneStrain.h:441)
            xg[((size_t)e * ng + g) * 2] = x0; xg[((size_t)e * ng + g) * 2 + 1] = x1;
        }
    }
}

__global__ void __launch_bounds__(128)
load_vector_kernel(int nelem, int shape, int quad, int ndof, const double* __restrict__ coords, const int* __restrict__ conn, const int* __restrict__ n2g,
                   double f0, double f1, const double* __restrict__ fg, double t, double* F) {
    const int npe = load_npe(shape), ng = load_ngauss(quad);
    const bool line = (shape == PF2_SHAPE_LINE2 || shape == PF2_SHAPE_LINE3);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        int nd[kLoadMaxNpe];
        double X[kLoadMaxNpe][2], fe[kLoadMaxNpe][2];
        for (int n = 0; n < npe; n++) {
            nd[n] = conn[(size_t)e * npe + n];
            X[n][0] = coords[2 * (size_t)nd[n]]; X[n][1] = coords[2 * (size_t)nd[n] + 1];
            fe[n][0] = 0.0; fe[n][1] = 0.0;
        }
        for (int g = 0; g < ng; g++) {
            LoadPoint P;
            double w;
            load_point(shape, quad, g, P, w);
            double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;                     // dXdr = dNdr * X
            for (int n = 0; n < npe; n++) {
                j00 += P.d[0][n] * X[n][0]; j01 += P.d[0][n] * X[n][1];
                if (!line) { j10 += P.d[1][n] * X[n][0]; j11 += P.d[1][n] * X[n][1]; }
            }
            const double m = line ? sqrt(j00 * j00 + j01 * j01) : (j00 * j11 - j01 * j10);      // PlaneStrain.h:443 / :524
            const double q0 = fg ? fg[((size_t)e * ng + g) * ndof] : f0;
            const double q1 = (ndof > 1) ? (fg ? fg[((size_t)e * ng + g) * ndof + 1] : f1) : 0.0;
            for (int n = 0; n < npe; n++) {
                fe[n][0] += P.N[n] * q0 * m * t * w;                                // B^T f * dl * t * w (PlaneStrain.h:451), N f dl t w (HeatTransfer.h:96)
                if (ndof > 1) fe[n][1] += P.N[n] * q1 * m * t * w;
            }
        }
        for (int n = 0; n < npe; n++)
            for (int i = 0; i < ndof; i++) {
                const int r = n2g[(size_t)nd[n] * ndof + i];
                if (r != -1) atomicAdd(&F[r], fe[n][i]);                            // Assembling.h:139-143
            }
    }
}

static int check_selection(const pf2_mesh* mesh, int shape, int quad) {
    PF2_CHECK(mesh && mesh->dim == 2, "load vectors: 2-D meshes (the reference's PlaneStrain / PlaneStress / HeatTransfer routines)");
    const bool line = (shape == PF2_SHAPE_LINE2 || shape == PF2_SHAPE_LINE3);
    const bool tri = (shape == PF2_SHAPE_T3 || shape == PF2_SHAPE_T6), sq = (shape == PF2_SHAPE_Q4 || shape == PF2_SHAPE_Q8);
    PF2_CHECK(line || tri || sq, "load vectors: shape must be LINE2, LINE3, T3, T6, Q4 or Q8");
    const int npe = shape == PF2_SHAPE_LINE2 ? 2 : (shape == PF2_SHAPE_LINE3 || shape == PF2_SHAPE_T3) ? 3 : shape == PF2_SHAPE_T6 ? 6 : shape == PF2_SHAPE_Q4 ? 4 : 8;
    PF2_CHECK(mesh->npe == npe, "load vectors: the mesh's nodes per element do not match the shape");
    if (line) PF2_CHECK(quad == PF2_QUAD_G1LINE || quad == PF2_QUAD_G2LINE, "line shapes integrate with Gauss1Line / Gauss2Line");
    if (tri) PF2_CHECK(quad == PF2_QUAD_G1TRI || quad == PF2_QUAD_G3TRI, "triangles integrate with Gauss1Triangle / Gauss3Triangle");
    if (sq) PF2_CHECK(quad == PF2_QUAD_G1SQ || quad == PF2_QUAD_G4SQ || quad == PF2_QUAD_G9SQ, "quadrilaterals integrate with Gauss1Square / 4 / 9");
    return PF2_OK;
}

}  // namespace pf2

using namespace pf2;

extern "C" {

int pf2_integration_points(pf2_mesh* mesh, int shape, int quad, double* xg_dev) {
    PF2_TRY(check_selection(mesh, shape, quad));
    PF2_CHECK(xg_dev, "null output");
    pf2_ctx* c = mesh->ctx;
    PF2_CUDA(cudaSetDevice(c->device));
    integration_points_kernel<<<c->grid_for(mesh->nelem), 128, 0, c->stream>>>(mesh->nelem, shape, quad, mesh->coords, mesh->conn, xg_dev);
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

int pf2_load_vector(pf2_mesh* mesh, pf2_dofmap* map, int shape, int quad, const double* f_const, const double* f_gauss_dev, double t, double* F_dev) {
    PF2_TRY(check_selection(mesh, shape, quad));
    PF2_CHECK(map && F_dev && (f_const || f_gauss_dev), "null argument");
    PF2_CHECK(map->ndof == 1 || map->ndof == 2, "load vectors: 1 dof (heat flux) or 2 dofs (forces) per node");
    pf2_ctx* c = mesh->ctx;
    PF2_CUDA(cudaSetDevice(c->device));
    load_vector_kernel<<<c->grid_for(mesh->nelem), 128, 0, c->stream>>>(mesh->nelem, shape, quad, map->ndof, mesh->coords, mesh->conn, map->n2g,
                                                                       f_const ? f_const[0] : 0.0, (f_const && map->ndof > 1) ? f_const[1] : 0.0,
                                                                       f_gauss_dev, t, F_dev);
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

}  // extern "C"
