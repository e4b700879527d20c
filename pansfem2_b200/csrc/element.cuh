// element.cuh -- device-side element kinematics for Q4 and hex8 with 2x2 / 2x2x2 Gauss quadrature.
//
// Restates, with register-resident 2x2 / 3x3 algebra instead of heap Matrix<T> temporaries:
//   ShapeFunction4Square::dNdr (ShapeFunction.h:186-191), ShapeFunction8Cubic::dNdr (ShapeFunction.h:318-329)
//   Gauss4Square / Gauss8Cubic points, unit weights (GaussIntegration.h:143-157, 231-253)
//   dXdr = dNdr*X ; J = det ; dNdX = dXdr^-1 * dNdr   (PlaneStrain.h:44-47, Solid.h:47-50, HeatTransfer.h:36-39)
//   with Determinant / adjugate inverse as Matrix.h:326-360.
// B^T D B is never formed as dense products: for the isotropic D / C of PlaneStrain.h:37-41 and Solid.h:37-44 the
// (node a, node b) block is  K_ab[i][j] = c_n*ga_i*gb_i + mu*sum_{k!=i} ga_k*gb_k  (i == j)
//                                     = lam*ga_i*gb_j + mu*ga_j*gb_i               (i != j)
// with g = dN/dX, c_n = (1-V)c, lam = V c, mu = (1-2V)c/2, c = E/((1+V)(1-2V)); the strain order of Solid.h:52-60
// (xx,yy,zz,xy,yz,zx) is what makes the shear terms pair up this way.
#pragma once
#include "common.cuh"

namespace pf2 {

#define PF2_INV_SQRT3 0.57735026918962584   // 1.0/sqrt(3.0) as GaussIntegration.h evaluates it

// ---- Q4 -------------------------------------------------------------------------------------------------------
// Gauss point g of Gauss4Square: (-,-),(+,-),(-,+),(+,+)
PF2_HD void q4_gauss(int g, double& r0, double& r1) {
    r0 = (g & 1) ? PF2_INV_SQRT3 : -PF2_INV_SQRT3;
    r1 = (g & 2) ? PF2_INV_SQRT3 : -PF2_INV_SQRT3;
}
// X: 4 nodes x 2.  Outputs dN/dX (gx[n], gy[n]) and det J.
PF2_HD void q4_grad(const double (&X)[4][2], double r0, double r1, double (&gx)[4], double (&gy)[4], double& det) {
    const double d0[4] = { -0.25 * (1.0 - r1), 0.25 * (1.0 - r1), 0.25 * (1.0 + r1), -0.25 * (1.0 + r1) };
    const double d1[4] = { -0.25 * (1.0 - r0), -0.25 * (1.0 + r0), 0.25 * (1.0 + r0), 0.25 * (1.0 - r0) };
    double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
#pragma unroll
    for (int n = 0; n < 4; n++) {
        J00 += d0[n] * X[n][0]; J01 += d0[n] * X[n][1];
        J10 += d1[n] * X[n][0]; J11 += d1[n] * X[n][1];
    }
    det = J00 * J11 - J01 * J10;
    const double idet = 1.0 / det;      // one reciprocal, then products (a DP division is ~10 dependent FMAs)
    const double i00 = J11 * idet, i01 = -J01 * idet, i10 = -J10 * idet, i11 = J00 * idet;
#pragma unroll
    for (int n = 0; n < 4; n++) {
        gx[n] = i00 * d0[n] + i01 * d1[n];
        gy[n] = i10 * d0[n] + i11 * d1[n];
    }
}

// ---- hex8 ------------------------------------------------------------------------------------------------------
PF2_HD double h8_sx(int n) { return ((n + 1) & 2) ? 1.0 : -1.0; }   // -,+,+,-,-,+,+,-
PF2_HD double h8_sy(int n) { return (n & 2) ? 1.0 : -1.0; }         // -,-,+,+,-,-,+,+
PF2_HD double h8_sz(int n) { return (n & 4) ? 1.0 : -1.0; }         // -,-,-,-,+,+,+,+
// Gauss8Cubic orders its points like the nodes (bottom CCW, top CCW)
PF2_HD void h8_gauss(int g, double& r0, double& r1, double& r2) {
    r0 = h8_sx(g) * PF2_INV_SQRT3; r1 = h8_sy(g) * PF2_INV_SQRT3; r2 = h8_sz(g) * PF2_INV_SQRT3;
}
PF2_HD void h8_grad(const double (&X)[8][3], double r0, double r1, double r2,
                                        double (&gx)[8], double (&gy)[8], double (&gz)[8], double& det) {
    double d0[8], d1[8], d2[8];
    double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const double sx = h8_sx(n), sy = h8_sy(n), sz = h8_sz(n);
        d0[n] = sx * 0.125 * (1.0 + sy * r1) * (1.0 + sz * r2);
        d1[n] = sy * 0.125 * (1.0 + sz * r2) * (1.0 + sx * r0);
        d2[n] = sz * 0.125 * (1.0 + sx * r0) * (1.0 + sy * r1);
#pragma unroll
        for (int k = 0; k < 3; k++) { J[0][k] += d0[n] * X[n][k]; J[1][k] += d1[n] * X[n][k]; J[2][k] += d2[n] * X[n][k]; }
    }
    det = -J[2][2] * J[0][1] * J[1][0] - J[2][1] * J[1][2] * J[0][0] - J[0][2] * J[1][1] * J[2][0]
          + J[2][0] * J[0][1] * J[1][2] + J[2][1] * J[1][0] * J[0][2] + J[0][0] * J[1][1] * J[2][2];
    // adjugate / det
    const double idet = 1.0 / det;      // one reciprocal, then products (a DP division is ~10 dependent FMAs)
    const double i00 = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * idet;
    const double i01 = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]) * idet;
    const double i02 = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet;
    const double i10 = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]) * idet;
    const double i11 = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet;
    const double i12 = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]) * idet;
    const double i20 = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * idet;
    const double i21 = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]) * idet;
    const double i22 = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
#pragma unroll
    for (int n = 0; n < 8; n++) {
        gx[n] = i00 * d0[n] + i01 * d1[n] + i02 * d2[n];
        gy[n] = i10 * d0[n] + i11 * d1[n] + i12 * d2[n];
        gz[n] = i20 * d0[n] + i21 * d1[n] + i22 * d2[n];
    }
}

// isotropic coefficients for unit modulus
struct Iso {
    double cn, lam, mu;
    PF2_HD Iso(double V) {
        const double c = 1.0 / ((1.0 + V) * (1.0 - 2.0 * V));
        cn = (1.0 - V) * c; lam = V * c; mu = 0.5 * (1.0 - 2.0 * V) * c;
    }
};

// SIMP interpolation of the drivers (sample_optimize_density_oc.cpp:123)
PF2_HD double simp_modulus(double rho, double E0, double E1, double p) {
    const double rp = pow(rho, p);
    return E1 * rp + E0 * (1.0 - rp);
}

// Rows of local node `a` of the element matrix for unit modulus (scaled by the caller).
// acc[i][b*NDOF + j], i = dof of node a.
template <int EQ> struct ElemTraits;
template <> struct ElemTraits<PF2_EQ_PLANESTRAIN> { static constexpr int DIM = 2, NPE = 4, NDOF = 2; };
template <> struct ElemTraits<PF2_EQ_HEAT> { static constexpr int DIM = 2, NPE = 4, NDOF = 1; };
template <> struct ElemTraits<PF2_EQ_SOLID> { static constexpr int DIM = 3, NPE = 8, NDOF = 3; };

PF2_HD void planestrain_rows(const double (&X)[4][2], int a, double V, double t, double (&acc)[2][8]) {
    const Iso c(V);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        double r0, r1, gx[4], gy[4], det;
        q4_gauss(g, r0, r1);
        q4_grad(X, r0, r1, gx, gy, det);
        const double w = det * t;
        double ax = gx[0], ay = gy[0];
#pragma unroll
        for (int n = 1; n < 4; n++) if (n == a) { ax = gx[n]; ay = gy[n]; }
        // node a's gradient folded with D and the weight once per point: 8 FMAs per neighbour node instead of 16 products + 16 adds
        const double cnx = c.cn * ax * w, cny = c.cn * ay * w, lmx = c.lam * ax * w, lmy = c.lam * ay * w, mux = c.mu * ax * w, muy = c.mu * ay * w;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            // two chained FMAs per entry (a sum of two products added to acc would cost a multiply, an FMA and an add)
            acc[0][2 * b]     += cnx * gx[b]; acc[0][2 * b]     += muy * gy[b];
            acc[0][2 * b + 1] += lmx * gy[b]; acc[0][2 * b + 1] += muy * gx[b];
            acc[1][2 * b]     += lmy * gx[b]; acc[1][2 * b]     += mux * gy[b];
            acc[1][2 * b + 1] += cny * gy[b]; acc[1][2 * b + 1] += mux * gx[b];
        }
    }
}

PF2_HD void heat_rows(const double (&X)[4][2], int a, double t, double (&acc)[1][4]) {
#pragma unroll
    for (int j = 0; j < 4; j++) acc[0][j] = 0.0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        double r0, r1, gx[4], gy[4], det;
        q4_gauss(g, r0, r1);
        q4_grad(X, r0, r1, gx, gy, det);
        const double w = det * t;
        double ax = gx[0], ay = gy[0];
#pragma unroll
        for (int n = 1; n < 4; n++) if (n == a) { ax = gx[n]; ay = gy[n]; }
        const double axw = ax * w, ayw = ay * w;
#pragma unroll
        for (int b = 0; b < 4; b++) { acc[0][b] += axw * gx[b]; acc[0][b] += ayw * gy[b]; }
    }
}

PF2_HD void solid_rows(const double (&X)[8][3], int a, double V, double (&acc)[3][24]) {
    const Iso c(V);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 24; j++) acc[i][j] = 0.0;
#pragma unroll 1
    for (int g = 0; g < 8; g++) {
        double r0, r1, r2, gx[8], gy[8], gz[8], det;
        h8_gauss(g, r0, r1, r2);
        h8_grad(X, r0, r1, r2, gx, gy, gz, det);
        double ga[3] = { gx[0], gy[0], gz[0] };
#pragma unroll
        for (int n = 1; n < 8; n++) if (n == a) { ga[0] = gx[n]; ga[1] = gy[n]; ga[2] = gz[n]; }
        // node a's gradient folded with C and the weight once per point
        double cg[3], lg[3], mg[3];
#pragma unroll
        for (int i = 0; i < 3; i++) { cg[i] = c.cn * ga[i] * det; lg[i] = c.lam * ga[i] * det; mg[i] = c.mu * ga[i] * det; }
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const double gb[3] = { gx[b], gy[b], gz[b] };
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    if (i == j) { acc[i][3 * b + j] += cg[i] * gb[i]; acc[i][3 * b + j] += mg[(i + 1) % 3] * gb[(i + 1) % 3]; acc[i][3 * b + j] += mg[(i + 2) % 3] * gb[(i + 2) % 3]; }
                    else { acc[i][3 * b + j] += lg[i] * gb[j]; acc[i][3 * b + j] += mg[j] * gb[i]; }
                }
        }
    }
}

// Row I of node a only (the hex8 gather assembly gives every dof row its own thread): the same operations in the same order per entry as
// solid_rows, so an element's contribution to K is bit-identical whichever kernel adds it.
template <int I>
PF2_HD void solid_row(const double (&X)[8][3], int a, double V, double (&acc)[24]) {
    const Iso c(V);
#pragma unroll
    for (int j = 0; j < 24; j++) acc[j] = 0.0;
#pragma unroll 1
    for (int g = 0; g < 8; g++) {
        double r0, r1, r2, gx[8], gy[8], gz[8], det;
        h8_gauss(g, r0, r1, r2);
        h8_grad(X, r0, r1, r2, gx, gy, gz, det);
        double ga[3] = { gx[0], gy[0], gz[0] };
#pragma unroll
        for (int n = 1; n < 8; n++) if (n == a) { ga[0] = gx[n]; ga[1] = gy[n]; ga[2] = gz[n]; }
        double mg[3];
#pragma unroll
        for (int i = 0; i < 3; i++) mg[i] = c.mu * ga[i] * det;
        const double cgi = c.cn * ga[I] * det, lgi = c.lam * ga[I] * det;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const double gb[3] = { gx[b], gy[b], gz[b] };
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (I == j) { acc[3 * b + j] += cgi * gb[I]; acc[3 * b + j] += mg[(I + 1) % 3] * gb[(I + 1) % 3]; acc[3 * b + j] += mg[(I + 2) % 3] * gb[(I + 2) % 3]; }
                else { acc[3 * b + j] += lgi * gb[j]; acc[3 * b + j] += mg[j] * gb[I]; }
            }
        }
    }
}

}  // namespace pf2
