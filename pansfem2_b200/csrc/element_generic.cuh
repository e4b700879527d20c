// element_generic.cuh -- the element families beyond Q4/hex8: one device template over
// <kind, shape> with the quadrature rule and the isotropic coefficients chosen at run time.
//
// Restates (paths relative to /root/reference/src/FEM):
//   Controller/ShapeFunction.h   dNdr of ShapeFunction3Triangle (:112-117), 6Triangle (:150-155), 4Square (:186-191),
//                                8Square (:226-246), 4Tetrahedron (:277-283), 8Cubic (:318-329), 20Cubic (:396-461)
//   Controller/GaussIntegration.h points / per-axis weights of Gauss1Triangle (:72-82), Gauss3Triangle (:94-107),
//                                Gauss1Square (:120-130), Gauss4Square (:142-157), Gauss9Square (:170-195),
//                                Gauss1Tetrahedron (:208-218), Gauss8Cubic (:230-253), Gauss27Cubic (:266-327)
//   Equation/PlaneStrain.h:21-58 (PlaneStrainStiffness), :63-125 (PlaneStrainStiffnessSRI), PlaneStress.h:21-58,
//   HeatTransfer.h:20-43, Solid.h:21-64:  Ke += B^T D B * J * t * w0 * w1 [* w2] over the rule's points.
//
// As in element.cuh, B^T D B is never formed: every D above has the shape [[cn, lam, 0], [lam, cn, 0], [0, 0, mu]]
// (resp. its 3-D analogue), so the (node a, node b) block is closed-form in the gradients.  The selective-reduced
// variant is two passes over two rules with (cn, lam, mu) = (k, k, 0), k = E/(3(1-2V)), and (4c, -2c, 3c), c = E/(6(1+V)).
#pragma once
#include "element.cuh"

namespace pf2 {

enum { SH_T3 = PF2_SHAPE_T3, SH_T6 = PF2_SHAPE_T6, SH_Q4 = PF2_SHAPE_Q4, SH_Q8 = PF2_SHAPE_Q8, SH_TET4 = PF2_SHAPE_TET4,
       SH_HEX8 = PF2_SHAPE_HEX8, SH_HEX20 = PF2_SHAPE_HEX20 };
enum { KIND_ELAST2D = 0, KIND_HEAT2D = 1, KIND_SOLID3D = 2, KIND_MASS2D = 3, KIND_MASS2D_V = 4,
       KIND_ADVDIFF2D = 5,     // lives in element_advdiff.cuh / advdiff.cu, not in the generic template
       KIND_ELAST2D_D = 6 };   // caller-supplied constitutive matrix: general_rows below, per-element entry point only

// what one launch needs to know about the element routine (filled on the host by decode_eq, passed by value)
struct ElemSpec {
    int npass;          // 1, or 2 for the selective-reduced plane-strain variant
    int quad[2];        // PF2_QUAD_* of each pass
    double cn[2], lam[2], mu[2];   // D for unit modulus (heat: unused)
    int wilson_taylor;             // 1: incompatible modes condensed out (PlaneStrainStiffnessWilsonTaylor), quads only
};

template <int SHAPE> struct ShapeTraits;
template <> struct ShapeTraits<SH_T3> { static constexpr int DIM = 2, NPE = 3; };
template <> struct ShapeTraits<SH_T6> { static constexpr int DIM = 2, NPE = 6; };
template <> struct ShapeTraits<SH_Q4> { static constexpr int DIM = 2, NPE = 4; };
template <> struct ShapeTraits<SH_Q8> { static constexpr int DIM = 2, NPE = 8; };
template <> struct ShapeTraits<SH_TET4> { static constexpr int DIM = 3, NPE = 4; };
template <> struct ShapeTraits<SH_HEX8> { static constexpr int DIM = 3, NPE = 8; };
template <> struct ShapeTraits<SH_HEX20> { static constexpr int DIM = 3, NPE = 20; };

template <int KIND> struct KindTraits;
template <> struct KindTraits<KIND_ELAST2D> { static constexpr int DIM = 2, NDOF = 2; };
template <> struct KindTraits<KIND_HEAT2D> { static constexpr int DIM = 2, NDOF = 1; };
template <> struct KindTraits<KIND_SOLID3D> { static constexpr int DIM = 3, NDOF = 3; };
template <> struct KindTraits<KIND_MASS2D> { static constexpr int DIM = 2, NDOF = 1; };     // scalar consistent mass N N^T
template <> struct KindTraits<KIND_MASS2D_V> { static constexpr int DIM = 2, NDOF = 2; };   // 2-dof consistent mass (N_a N_b) I_2

// ---- quadrature ------------------------------------------------------------------------------------------------
PF2_HD int quad_count(int quad) {
    switch (quad) {
        case PF2_QUAD_G3TRI: return 3;
        case PF2_QUAD_G4SQ: return 4;
        case PF2_QUAD_G9SQ: return 9;
        case PF2_QUAD_G8CUBE: return 8;
        case PF2_QUAD_G27CUBE: return 27;
        default: return 1;      // G1TRI, G1SQ, G1TET
    }
}
// point g of the rule and the PRODUCT of its per-axis weights (the reference multiplies Weights[g][0]*Weights[g][1][*Weights[g][2]])
PF2_HD void quad_point(int quad, int g, double (&r)[3], double& w) {
    const double s35 = sqrt(3.0 / 5.0);
    r[2] = 0.0;
    switch (quad) {
        case PF2_QUAD_G1TRI: { r[0] = 1.0 / 3.0; r[1] = 1.0 / 3.0; const double a = 1.0 / sqrt(2.0); w = a * a; break; }
        case PF2_QUAD_G3TRI: {
            r[0] = (g == 1) ? 2.0 / 3.0 : 1.0 / 6.0; r[1] = (g == 2) ? 2.0 / 3.0 : 1.0 / 6.0;
            const double a = 1.0 / sqrt(6.0); w = a * a; break;
        }
        case PF2_QUAD_G1SQ: { r[0] = 0.0; r[1] = 0.0; w = 2.0 * 2.0; break; }
        case PF2_QUAD_G4SQ: { q4_gauss(g, r[0], r[1]); w = 1.0; break; }
        case PF2_QUAD_G9SQ: {
            const int i = g % 3, j = g / 3;
            r[0] = (i - 1) * s35; r[1] = (j - 1) * s35;
            w = ((i == 1) ? 8.0 / 9.0 : 5.0 / 9.0) * ((j == 1) ? 8.0 / 9.0 : 5.0 / 9.0); break;
        }
        case PF2_QUAD_G1TET: { r[0] = 0.25; r[1] = 0.25; r[2] = 0.25; const double a = 1.0 / cbrt(6.0); w = a * a * a; break; }
        case PF2_QUAD_G8CUBE: { h8_gauss(g, r[0], r[1], r[2]); w = 1.0; break; }
        default: {  // G27CUBE
            const int i = g % 3, j = (g / 3) % 3, k = g / 9;
            r[0] = (i - 1) * s35; r[1] = (j - 1) * s35; r[2] = (k - 1) * s35;
            w = ((i == 1) ? 8.0 / 9.0 : 5.0 / 9.0) * ((j == 1) ? 8.0 / 9.0 : 5.0 / 9.0) * ((k == 1) ? 8.0 / 9.0 : 5.0 / 9.0); break;
        }
    }
}

// ---- dN/dr, d[k][n] ----------------------------------------------------------------------------------------------
template <int SHAPE>
PF2_HD void shape_dndr(const double (&r)[3], double (&d)[ShapeTraits<SHAPE>::DIM][ShapeTraits<SHAPE>::NPE]) {
    const double r0 = r[0], r1 = r[1], r2 = r[2];
    if constexpr (SHAPE == SH_T3) {
        d[0][0] = 1.0; d[0][1] = 0.0; d[0][2] = -1.0;
        d[1][0] = 0.0; d[1][1] = 1.0; d[1][2] = -1.0;
    } else if constexpr (SHAPE == SH_T6) {
        d[0][0] = 4.0 * r0 - 1.0; d[0][1] = 0.0; d[0][2] = -3.0 + 4.0 * r0 + 4.0 * r1; d[0][3] = 4.0 * r1; d[0][4] = -4.0 * r1;
        d[0][5] = 4.0 * (1.0 - 2.0 * r0 - r1);
        d[1][0] = 0.0; d[1][1] = 4.0 * r1 - 1.0; d[1][2] = -3.0 + 4.0 * r0 + 4.0 * r1; d[1][3] = 4.0 * r0;
        d[1][4] = 4.0 * (1.0 - r0 - 2.0 * r1); d[1][5] = -4.0 * r0;
    } else if constexpr (SHAPE == SH_Q4) {
        d[0][0] = -0.25 * (1.0 - r1); d[0][1] = 0.25 * (1.0 - r1); d[0][2] = 0.25 * (1.0 + r1); d[0][3] = -0.25 * (1.0 + r1);
        d[1][0] = -0.25 * (1.0 - r0); d[1][1] = -0.25 * (1.0 + r0); d[1][2] = 0.25 * (1.0 + r0); d[1][3] = 0.25 * (1.0 - r0);
    } else if constexpr (SHAPE == SH_Q8) {
        // corners n = 0..3 with signs (sx, sy): N = (1 + sx r0)(1 + sy r1)(sx r0 + sy r1 - 1)/4
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const double sx = ((n + 1) & 2) ? 1.0 : -1.0, sy = (n & 2) ? 1.0 : -1.0;
            d[0][n] = 0.25 * sx * (1.0 + sy * r1) * (2.0 * sx * r0 + sy * r1);
            d[1][n] = 0.25 * sy * (1.0 + sx * r0) * (sx * r0 + 2.0 * sy * r1);
        }
        // mid-side nodes 4:(0,-1) 5:(1,0) 6:(0,1) 7:(-1,0)
        d[0][4] = -r0 * (1.0 - r1);               d[1][4] = -0.5 * (1.0 + r0) * (1.0 - r0);
        d[0][5] = 0.5 * (1.0 + r1) * (1.0 - r1);  d[1][5] = -r1 * (1.0 + r0);
        d[0][6] = -r0 * (1.0 + r1);               d[1][6] = 0.5 * (1.0 + r0) * (1.0 - r0);
        d[0][7] = -0.5 * (1.0 + r1) * (1.0 - r1); d[1][7] = -r1 * (1.0 - r0);
    } else if constexpr (SHAPE == SH_TET4) {
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int n = 0; n < 4; n++) d[k][n] = (n == 3) ? -1.0 : ((n == k) ? 1.0 : 0.0);
    } else if constexpr (SHAPE == SH_HEX8) {
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const double sx = h8_sx(n), sy = h8_sy(n), sz = h8_sz(n);
            d[0][n] = sx * 0.125 * (1.0 + sy * r1) * (1.0 + sz * r2);
            d[1][n] = sy * 0.125 * (1.0 + sz * r2) * (1.0 + sx * r0);
            d[2][n] = sz * 0.125 * (1.0 + sx * r0) * (1.0 + sy * r1);
        }
    } else {    // SH_HEX20
        // corners: N = (1 + sx r0)(1 + sy r1)(1 + sz r2)(sx r0 + sy r1 + sz r2 - 2)/8
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const double sx = h8_sx(n), sy = h8_sy(n), sz = h8_sz(n);
            const double a = 1.0 + sx * r0, b = 1.0 + sy * r1, c = 1.0 + sz * r2;
            d[0][n] = 0.125 * sx * b * c * (2.0 * sx * r0 + sy * r1 + sz * r2 - 1.0);
            d[1][n] = 0.125 * sy * a * c * (sx * r0 + 2.0 * sy * r1 + sz * r2 - 1.0);
            d[2][n] = 0.125 * sz * a * b * (sx * r0 + sy * r1 + 2.0 * sz * r2 - 1.0);
        }
        // edge mid-points in the reference's own order (ShapeFunction.h:354-365):
        //   8,10,12,14 on r0-edges (sy,sz) = (-,-),(+,-),(-,+),(+,+) ; 9,11,13,15 on r1-edges (sx,sz) = (+,-),(-,-),(+,+),(-,+)
        //   16..19 on r2-edges (sx,sy) = (-,-),(+,-),(+,+),(-,+)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            {   // r0-edge node 8 + 2q
                const double sy = (q & 1) ? 1.0 : -1.0, sz = (q & 2) ? 1.0 : -1.0;
                const int n = 8 + 2 * q;
                d[0][n] = -0.5 * r0 * (1.0 + sy * r1) * (1.0 + sz * r2);
                d[1][n] = 0.25 * sy * (1.0 - r0 * r0) * (1.0 + sz * r2);
                d[2][n] = 0.25 * sz * (1.0 - r0 * r0) * (1.0 + sy * r1);
            }
            {   // r1-edge node 9 + 2q
                const double sx = (q & 1) ? -1.0 : 1.0, sz = (q & 2) ? 1.0 : -1.0;
                const int n = 9 + 2 * q;
                d[0][n] = 0.25 * sx * (1.0 - r1 * r1) * (1.0 + sz * r2);
                d[1][n] = -0.5 * r1 * (1.0 + sx * r0) * (1.0 + sz * r2);
                d[2][n] = 0.25 * sz * (1.0 + sx * r0) * (1.0 - r1 * r1);
            }
            {   // r2-edge node 16 + q
                const double sx = ((q + 1) & 2) ? 1.0 : -1.0, sy = (q & 2) ? 1.0 : -1.0;
                const int n = 16 + q;
                d[0][n] = 0.25 * sx * (1.0 + sy * r1) * (1.0 - r2 * r2);
                d[1][n] = 0.25 * sy * (1.0 + sx * r0) * (1.0 - r2 * r2);
                d[2][n] = -0.5 * r2 * (1.0 + sx * r0) * (1.0 + sy * r1);
            }
        }
    }
}

// N(r) of the 2-D shapes (ShapeFunction.h:102-108, 137-146, 175-182, 211-222): needed by the consistent mass matrices
template <int SHAPE>
PF2_HD void shape_n(const double (&r)[3], double (&N)[ShapeTraits<SHAPE>::NPE]) {
    const double r0 = r[0], r1 = r[1];
    if constexpr (SHAPE == SH_T3) {
        N[0] = r0; N[1] = r1; N[2] = 1.0 - r0 - r1;
    } else if constexpr (SHAPE == SH_T6) {
        const double c = 1.0 - r0 - r1;
        N[0] = r0 * (2.0 * r0 - 1.0); N[1] = r1 * (2.0 * r1 - 1.0); N[2] = c * (1.0 - 2.0 * r0 - 2.0 * r1);
        N[3] = 4.0 * r0 * r1; N[4] = 4.0 * r1 * c; N[5] = 4.0 * c * r0;
    } else if constexpr (SHAPE == SH_Q4) {
        N[0] = 0.25 * (1.0 - r0) * (1.0 - r1); N[1] = 0.25 * (1.0 + r0) * (1.0 - r1);
        N[2] = 0.25 * (1.0 + r0) * (1.0 + r1); N[3] = 0.25 * (1.0 - r0) * (1.0 + r1);
    } else if constexpr (SHAPE == SH_Q8) {
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const double sx = ((n + 1) & 2) ? 1.0 : -1.0, sy = (n & 2) ? 1.0 : -1.0;
            N[n] = 0.25 * (1.0 + sx * r0) * (1.0 + sy * r1) * (sx * r0 + sy * r1 - 1.0);
        }
        N[4] = 0.5 * (1.0 - r0) * (1.0 + r0) * (1.0 - r1); N[5] = 0.5 * (1.0 + r0) * (1.0 + r1) * (1.0 - r1);
        N[6] = 0.5 * (1.0 + r0) * (1.0 - r0) * (1.0 + r1); N[7] = 0.5 * (1.0 - r0) * (1.0 + r1) * (1.0 - r1);
    } else {
#pragma unroll
        for (int n = 0; n < ShapeTraits<SHAPE>::NPE; n++) N[n] = 0.0;      // 3-D shapes: no mass kind instantiated
    }
}

// dXdr = dNdr * X, J = det, dNdX = dXdr^-1 * dNdr   (g overwrites d in place)
template <int SHAPE>
PF2_HD void shape_grad(const double (&X)[ShapeTraits<SHAPE>::NPE][ShapeTraits<SHAPE>::DIM], const double (&r)[3],
                                           double (&g)[ShapeTraits<SHAPE>::DIM][ShapeTraits<SHAPE>::NPE], double& det) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE;
    shape_dndr<SHAPE>(r, g);
    double J[DIM][DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double v = 0.0;
#pragma unroll
            for (int n = 0; n < NPE; n++) v += g[i][n] * X[n][k];
            J[i][k] = v;
        }
    if constexpr (DIM == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double idet = 1.0 / det;      // one reciprocal, then products (a DP division is ~10 dependent FMAs)
        const double i00 = J[1][1] * idet, i01 = -J[0][1] * idet, i10 = -J[1][0] * idet, i11 = J[0][0] * idet;
#pragma unroll
        for (int n = 0; n < NPE; n++) {
            const double d0 = g[0][n], d1 = g[1][n];
            g[0][n] = i00 * d0 + i01 * d1;
            g[1][n] = i10 * d0 + i11 * d1;
        }
    } else {
        det = -J[2][2] * J[0][1] * J[1][0] - J[2][1] * J[1][2] * J[0][0] - J[0][2] * J[1][1] * J[2][0]
              + J[2][0] * J[0][1] * J[1][2] + J[2][1] * J[1][0] * J[0][2] + J[0][0] * J[1][1] * J[2][2];
        const double idet = 1.0 / det;      // one reciprocal, then products (a DP division is ~10 dependent FMAs)
        const double i00 = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * idet, i01 = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]) * idet;
        const double i02 = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet, i10 = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]) * idet;
        const double i11 = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet, i12 = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]) * idet;
        const double i20 = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * idet, i21 = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]) * idet;
        const double i22 = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
#pragma unroll
        for (int n = 0; n < NPE; n++) {
            const double d0 = g[0][n], d1 = g[1][n], d2 = g[2][n];
            g[0][n] = i00 * d0 + i01 * d1 + i02 * d2;
            g[1][n] = i10 * d0 + i11 * d1 + i12 * d2;
            g[2][n] = i20 * d0 + i21 * d1 + i22 * d2;
        }
    }
}

// ---- Wilson-Taylor incompatible modes (PlaneStrain.h:189-243) --------------------------------------------------------
// The two modes P = (1 - r0^2, 1 - r1^2) act like two extra nodes whose "gradients" are h_m = -2 r_m * column m of dXdr^-1, so the
// blocks Keaa (modes x modes), Kead (modes x nodes) come from the same closed form as the nodal blocks; then
// Ke -= Kead^T Keaa^-1 Kead.  Gradients + the two mode gradients at one integration point:
template <int SHAPE>
PF2_HD void wt_grad(const double (&X)[ShapeTraits<SHAPE>::NPE][2], const double (&r)[3], double (&g)[2][ShapeTraits<SHAPE>::NPE],
                                        double (&h)[2][2], double& det) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE;
    shape_dndr<SHAPE>(r, g);
    double J00 = 0, J01 = 0, J10 = 0, J11 = 0;
#pragma unroll
    for (int n = 0; n < NPE; n++) { J00 += g[0][n] * X[n][0]; J01 += g[0][n] * X[n][1]; J10 += g[1][n] * X[n][0]; J11 += g[1][n] * X[n][1]; }
    det = J00 * J11 - J01 * J10;
    const double idet = 1.0 / det;      // one reciprocal, then products (a DP division is ~10 dependent FMAs)
    const double i00 = J11 * idet, i01 = -J01 * idet, i10 = -J10 * idet, i11 = J00 * idet;
#pragma unroll
    for (int n = 0; n < NPE; n++) {
        const double d0 = g[0][n], d1 = g[1][n];
        g[0][n] = i00 * d0 + i01 * d1;
        g[1][n] = i10 * d0 + i11 * d1;
    }
    // dPdX = dXdr^-1 * diag(-2 r0, -2 r1): mode m has gradient (dPdX(0,m), dPdX(1,m))
    h[0][0] = i00 * (-2.0 * r[0]); h[1][0] = i10 * (-2.0 * r[0]);
    h[0][1] = i01 * (-2.0 * r[1]); h[1][1] = i11 * (-2.0 * r[1]);
}
// isotropic block K_ab[i][j] for gradients ga, gb (2-D)
PF2_HD double iso_block2(double cn, double lam, double mu, const double (&ga)[2], const double (&gb)[2], int i, int j) {
    return (i == j) ? (cn * ga[i] * gb[i] + mu * ga[1 - i] * gb[1 - i]) : (lam * ga[i] * gb[j] + mu * ga[j] * gb[i]);
}
// solve the 4x4 SPD system Kaa z = v (Gaussian elimination, no pivoting needed)
PF2_HD void solve4(double (&A)[4][4], double (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double piv = 1.0 / A[k][k];
#pragma unroll
        for (int i = k + 1; i < 4; i++) {
            const double f = A[i][k] * piv;
#pragma unroll
            for (int j = k + 1; j < 4; j++) A[i][j] -= f * A[k][j];
            v[i] -= f * v[k];
        }
    }
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        double s = v[k];
#pragma unroll
        for (int j = k + 1; j < 4; j++) s -= A[k][j] * v[j];
        v[k] = s / A[k][k];
    }
}

// rows of node a of the condensed matrix (unit modulus)
template <int SHAPE>
PF2_HD void wt_rows(const double (&X)[ShapeTraits<SHAPE>::NPE][2], int a, const ElemSpec& sp, double t,
                                        double (&acc)[2][ShapeTraits<SHAPE>::NPE * 2]) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE, M = NPE * 2;
    const double cn = sp.cn[0], lam = sp.lam[0], mu = sp.mu[0];
    double Kaa[4][4], Kad[4][M];
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) Kaa[i][j] = 0.0;
#pragma unroll
        for (int j = 0; j < M; j++) Kad[i][j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < M; j++) acc[i][j] = 0.0;
    const int ng = quad_count(sp.quad[0]);
#pragma unroll 1
    for (int q = 0; q < ng; q++) {
        double r[3], wq, det, g[2][NPE], h[2][2];
        quad_point(sp.quad[0], q, r, wq);
        wt_grad<SHAPE>(X, r, g, h, det);
        const double w = det * t * wq;
        double ga[2] = { g[0][0], g[1][0] };
#pragma unroll
        for (int n = 1; n < NPE; n++) if (n == a) { ga[0] = g[0][n]; ga[1] = g[1][n]; }
#pragma unroll
        for (int b = 0; b < NPE; b++) {
            const double gb[2] = { g[0][b], g[1][b] };
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    acc[i][b * 2 + j] += iso_block2(cn, lam, mu, ga, gb, i, j) * w;
#pragma unroll
                    for (int mm = 0; mm < 2; mm++) {
                        const double hm[2] = { h[0][mm], h[1][mm] };
                        Kad[mm * 2 + i][b * 2 + j] += iso_block2(cn, lam, mu, hm, gb, i, j) * w;
                    }
                }
        }
#pragma unroll
        for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
            for (int m2 = 0; m2 < 2; m2++) {
                const double h1[2] = { h[0][m1], h[1][m1] }, h2[2] = { h[0][m2], h[1][m2] };
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) Kaa[m1 * 2 + i][m2 * 2 + j] += iso_block2(cn, lam, mu, h1, h2, i, j) * w;
            }
    }
    // acc_i -= Kad[:, a*2+i]^T Kaa^-1 Kad
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double z[4], A[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            double v = Kad[k][0];
#pragma unroll
            for (int c = 1; c < M; c++) if (c == a * 2 + i) v = Kad[k][c];
            z[k] = v;
#pragma unroll
            for (int j = 0; j < 4; j++) A[k][j] = Kaa[k][j];
        }
        solve4(A, z);
#pragma unroll
        for (int c = 0; c < M; c++) acc[i][c] -= z[0] * Kad[0][c] + z[1] * Kad[1][c] + z[2] * Kad[2][c] + z[3] * Kad[3][c];
    }
}

// ue^T Ke ue and optionally Ke ue for the condensed element
template <int SHAPE, bool WANT_F>
PF2_HD double wt_energy(const double (&X)[ShapeTraits<SHAPE>::NPE][2], const double (&ue)[ShapeTraits<SHAPE>::NPE][2], const ElemSpec& sp,
                                            double t, double (&fe)[ShapeTraits<SHAPE>::NPE][2]) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE;
    const double cn = sp.cn[0], lam = sp.lam[0], mu = sp.mu[0];
    double Kaa[4][4], v[4] = { 0.0, 0.0, 0.0, 0.0 }, wsum = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) Kaa[i][j] = 0.0;
    const int ng = quad_count(sp.quad[0]);
#pragma unroll 1
    for (int q = 0; q < ng; q++) {
        double r[3], wq, det, g[2][NPE], h[2][2];
        quad_point(sp.quad[0], q, r, wq);
        wt_grad<SHAPE>(X, r, g, h, det);
        const double w = det * t * wq;
        double exx = 0, eyy = 0, gxy = 0;
#pragma unroll
        for (int n = 0; n < NPE; n++) { exx += g[0][n] * ue[n][0]; eyy += g[1][n] * ue[n][1]; gxy += g[1][n] * ue[n][0] + g[0][n] * ue[n][1]; }
        const double sxx = cn * exx + lam * eyy, syy = cn * eyy + lam * exx, sxy = mu * gxy;
        wsum += (sxx * exx + syy * eyy + sxy * gxy) * w;
        if constexpr (WANT_F) {
#pragma unroll
            for (int n = 0; n < NPE; n++) { fe[n][0] += (g[0][n] * sxx + g[1][n] * sxy) * w; fe[n][1] += (g[1][n] * syy + g[0][n] * sxy) * w; }
        }
#pragma unroll
        for (int mm = 0; mm < 2; mm++) {        // v = Kad ue = sum_g G^T sigma(ue)
            v[mm * 2 + 0] += (h[0][mm] * sxx + h[1][mm] * sxy) * w;
            v[mm * 2 + 1] += (h[1][mm] * syy + h[0][mm] * sxy) * w;
        }
#pragma unroll
        for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
            for (int m2 = 0; m2 < 2; m2++) {
                const double h1[2] = { h[0][m1], h[1][m1] }, h2[2] = { h[0][m2], h[1][m2] };
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) Kaa[m1 * 2 + i][m2 * 2 + j] += iso_block2(cn, lam, mu, h1, h2, i, j) * w;
            }
    }
    double z[4] = { v[0], v[1], v[2], v[3] };
    solve4(Kaa, z);                             // z = Kaa^-1 Kad ue  (the condensed mode amplitudes, sign flipped)
    wsum -= v[0] * z[0] + v[1] * z[1] + v[2] * z[2] + v[3] * z[3];
    if constexpr (WANT_F) {
        // fe -= Kad^T z = sum_g B^T sigma(G z)
#pragma unroll 1
        for (int q = 0; q < ng; q++) {
            double r[3], wq, det, g[2][NPE], h[2][2];
            quad_point(sp.quad[0], q, r, wq);
            wt_grad<SHAPE>(X, r, g, h, det);
            const double w = det * t * wq;
            const double exx = h[0][0] * z[0] + h[0][1] * z[2], eyy = h[1][0] * z[1] + h[1][1] * z[3];
            const double gxy = h[1][0] * z[0] + h[0][0] * z[1] + h[1][1] * z[2] + h[0][1] * z[3];
            const double sxx = cn * exx + lam * eyy, syy = cn * eyy + lam * exx, sxy = mu * gxy;
#pragma unroll
            for (int n = 0; n < NPE; n++) { fe[n][0] -= (g[0][n] * sxx + g[1][n] * sxy) * w; fe[n][1] -= (g[1][n] * syy + g[0][n] * sxy) * w; }
        }
    }
    return wsum;
}

// Rows of local node `a` of the element matrix for unit modulus: acc[i][b*NDOF + j], i = dof of node a.
template <int KIND, int SHAPE>
PF2_HD void generic_rows(const double (&X)[ShapeTraits<SHAPE>::NPE][ShapeTraits<SHAPE>::DIM], int a, const ElemSpec& sp,
                                             double t, double (&acc)[KindTraits<KIND>::NDOF][ShapeTraits<SHAPE>::NPE * KindTraits<KIND>::NDOF]) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE, NDOF = KindTraits<KIND>::NDOF;
    static_assert(DIM == KindTraits<KIND>::DIM, "shape / equation dimension mismatch");
    if constexpr (KIND == KIND_ELAST2D && (SHAPE == SH_Q4 || SHAPE == SH_Q8)) {
        if (sp.wilson_taylor) { wt_rows<SHAPE>(X, a, sp, t, acc); return; }
    }
#pragma unroll
    for (int i = 0; i < NDOF; i++)
#pragma unroll
        for (int j = 0; j < NPE * NDOF; j++) acc[i][j] = 0.0;
    for (int pass = 0; pass < sp.npass; pass++) {
        const double cn = sp.cn[pass], lam = sp.lam[pass], mu = sp.mu[pass];
        const int ng = quad_count(sp.quad[pass]);
#pragma unroll 1
        for (int q = 0; q < ng; q++) {
            double r[3], wq, det, g[DIM][NPE];
            quad_point(sp.quad[pass], q, r, wq);
            shape_grad<SHAPE>(X, r, g, det);
            const double w = (KIND == KIND_SOLID3D) ? det * wq : det * t * wq;
            double ga[DIM];
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                ga[k] = g[k][0];
#pragma unroll
                for (int n = 1; n < NPE; n++) if (n == a) ga[k] = g[k][n];
            }
            if constexpr (KIND == KIND_MASS2D || KIND == KIND_MASS2D_V) {
                double N[NPE];
                shape_n<SHAPE>(r, N);
                double Na = N[0];
#pragma unroll
                for (int n = 1; n < NPE; n++) if (n == a) Na = N[n];
#pragma unroll
                for (int b = 0; b < NPE; b++) {
#pragma unroll
                    for (int i = 0; i < NDOF; i++) acc[i][b * NDOF + i] += Na * N[b] * w;
                }
                continue;
            }
            // node a's gradient folded with D and the weight once per point (DIM FMAs per diagonal entry, 2 per off-diagonal one)
            double cg[DIM], lg[DIM], mg[DIM];
#pragma unroll
            for (int k = 0; k < DIM; k++) { cg[k] = cn * ga[k] * w; lg[k] = lam * ga[k] * w; mg[k] = mu * ga[k] * w; }
#pragma unroll
            for (int b = 0; b < NPE; b++) {
                if constexpr (KIND == KIND_HEAT2D) {
                    acc[0][b] += (ga[0] * w) * g[0][b]; acc[0][b] += (ga[1] * w) * g[1][b];
                } else {
#pragma unroll
                    for (int i = 0; i < NDOF; i++)
#pragma unroll
                        for (int j = 0; j < NDOF; j++) {
                            // chained FMAs into the accumulator
                            if (i == j) {
                                acc[i][b * NDOF + j] += cg[i] * g[i][b];
#pragma unroll
                                for (int k = 0; k < DIM; k++) if (k != i) acc[i][b * NDOF + j] += mg[k] * g[k][b];
                            } else { acc[i][b * NDOF + j] += lg[i] * g[j][b]; acc[i][b * NDOF + j] += mg[j] * g[i][b]; }
                        }
                }
            }
        }
    }
}

// strain energy ue^T Ke(E=1) ue of one element and, optionally, fe = Ke(E=1) ue
template <int KIND, int SHAPE, bool WANT_F>
PF2_HD double generic_energy(const double (&X)[ShapeTraits<SHAPE>::NPE][ShapeTraits<SHAPE>::DIM],
                                                 const double (&ue)[ShapeTraits<SHAPE>::NPE][KindTraits<KIND>::NDOF], const ElemSpec& sp, double t,
                                                 double (&fe)[ShapeTraits<SHAPE>::NPE][KindTraits<KIND>::NDOF]) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE;
    if constexpr (KIND == KIND_ELAST2D && (SHAPE == SH_Q4 || SHAPE == SH_Q8)) {
        if (sp.wilson_taylor) return wt_energy<SHAPE, WANT_F>(X, ue, sp, t, fe);
    }
    double wsum = 0.0;
    for (int pass = 0; pass < sp.npass; pass++) {
        const double cn = sp.cn[pass], lam = sp.lam[pass], mu = sp.mu[pass];
        const int ng = quad_count(sp.quad[pass]);
#pragma unroll 1
        for (int q = 0; q < ng; q++) {
            double r[3], wq, det, g[DIM][NPE];
            quad_point(sp.quad[pass], q, r, wq);
            shape_grad<SHAPE>(X, r, g, det);
            const double w = (KIND == KIND_SOLID3D) ? det * wq : det * t * wq;
            if constexpr (KIND == KIND_MASS2D || KIND == KIND_MASS2D_V) {
                constexpr int ND = KindTraits<KIND>::NDOF;
                double N[NPE];
                shape_n<SHAPE>(r, N);
#pragma unroll
                for (int d = 0; d < ND; d++) {
                    double ug = 0.0;
#pragma unroll
                    for (int n = 0; n < NPE; n++) ug += N[n] * ue[n][d];
                    wsum += ug * ug * w;
                    if constexpr (WANT_F) {
#pragma unroll
                        for (int n = 0; n < NPE; n++) fe[n][d] += N[n] * ug * w;
                    }
                }
            } else if constexpr (KIND == KIND_HEAT2D) {
                double qx = 0.0, qy = 0.0;
#pragma unroll
                for (int n = 0; n < NPE; n++) { qx += g[0][n] * ue[n][0]; qy += g[1][n] * ue[n][0]; }
                wsum += (qx * qx + qy * qy) * w;
                if constexpr (WANT_F) {
#pragma unroll
                    for (int n = 0; n < NPE; n++) fe[n][0] += (g[0][n] * qx + g[1][n] * qy) * w;
                }
            } else if constexpr (KIND == KIND_ELAST2D) {
                double exx = 0, eyy = 0, gxy = 0;
#pragma unroll
                for (int n = 0; n < NPE; n++) {
                    exx += g[0][n] * ue[n][0]; eyy += g[1][n] * ue[n][1]; gxy += g[1][n] * ue[n][0] + g[0][n] * ue[n][1];
                }
                const double sxx = cn * exx + lam * eyy, syy = cn * eyy + lam * exx, sxy = mu * gxy;
                wsum += (sxx * exx + syy * eyy + sxy * gxy) * w;
                if constexpr (WANT_F) {
#pragma unroll
                    for (int n = 0; n < NPE; n++) {
                        fe[n][0] += (g[0][n] * sxx + g[1][n] * sxy) * w;
                        fe[n][1] += (g[1][n] * syy + g[0][n] * sxy) * w;
                    }
                }
            } else {
                double exx = 0, eyy = 0, ezz = 0, gxy = 0, gyz = 0, gzx = 0;
#pragma unroll
                for (int n = 0; n < NPE; n++) {
                    const double ux = ue[n][0], uy = ue[n][1], uz = ue[n][2];
                    exx += g[0][n] * ux; eyy += g[1][n] * uy; ezz += g[2][n] * uz;
                    gxy += g[1][n] * ux + g[0][n] * uy; gyz += g[2][n] * uy + g[1][n] * uz; gzx += g[2][n] * ux + g[0][n] * uz;
                }
                const double sxx = cn * exx + lam * (eyy + ezz), syy = cn * eyy + lam * (exx + ezz), szz = cn * ezz + lam * (exx + eyy);
                const double sxy = mu * gxy, syz = mu * gyz, szx = mu * gzx;
                wsum += (sxx * exx + syy * eyy + szz * ezz + sxy * gxy + syz * gyz + szx * gzx) * w;
                if constexpr (WANT_F) {
#pragma unroll
                    for (int n = 0; n < NPE; n++) {
                        fe[n][0] += (g[0][n] * sxx + g[1][n] * sxy + g[2][n] * szx) * w;
                        fe[n][1] += (g[1][n] * syy + g[0][n] * sxy + g[2][n] * syz) * w;
                        fe[n][2] += (g[2][n] * szz + g[1][n] * syz + g[0][n] * szx) * w;
                    }
                }
            }
        }
    }
    return wsum;
}

// ---- arbitrary constitutive matrix (FEM/Equation/Homogenization.h:141-280) -----------------------------------------------------
// PlaneStiffness, PlaneStiffnessBbar and PlaneStiffnessWilsonTaylor take a caller-supplied 3 x 3 D (a homogenised, rotated material),
// so the isotropic closed form above does not apply: the (a, b) block is Ba^T D Bb with the 3 x 2 strain-displacement columns of the two
// nodes.  mode 0: B = [gx 0; 0 gy; gy gx];  mode 1 (B-bar): Bvol = [gx gy; gx gy; 0 0]/2 on rule quad[0] plus
// Bdev = [gx -gy; -gx gy; 2gy 2gx]/2 on rule quad[1];  mode 2: Wilson-Taylor modes condensed out as in wt_rows, Ke -= Kead^T Keaa^-1 Kead.
struct ElemSpecD {
    int mode;           // 0 plain, 1 B-bar, 2 Wilson-Taylor
    int quad[2];
    double D[9];        // row-major
};

PF2_HD void strain_columns(int part, double gx, double gy, double (&B)[3][2]) {
    if (part == 1) { B[0][0] = 0.5 * gx; B[0][1] = 0.5 * gy; B[1][0] = 0.5 * gx; B[1][1] = 0.5 * gy; B[2][0] = 0.0; B[2][1] = 0.0; }
    else if (part == 2) { B[0][0] = 0.5 * gx; B[0][1] = -0.5 * gy; B[1][0] = -0.5 * gx; B[1][1] = 0.5 * gy; B[2][0] = gy; B[2][1] = gx; }
    else { B[0][0] = gx; B[0][1] = 0.0; B[1][0] = 0.0; B[1][1] = gy; B[2][0] = gy; B[2][1] = gx; }
}
// K += (Ba^T D Bb) * w
PF2_HD void general_block(const double (&Ba)[3][2], const double (&Bb)[3][2], const double (&D)[9], double w, double (&K)[2][2]) {
    double DB[3][2];
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
        for (int j = 0; j < 2; j++) DB[p][j] = D[p * 3] * Bb[0][j] + D[p * 3 + 1] * Bb[1][j] + D[p * 3 + 2] * Bb[2][j];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) K[i][j] += (Ba[0][i] * DB[0][j] + Ba[1][i] * DB[1][j] + Ba[2][i] * DB[2][j]) * w;
}

// rows of local node a: acc[i][b*2 + j]
template <int SHAPE>
PF2_HD void general_rows(const double (&X)[ShapeTraits<SHAPE>::NPE][2], int a, const ElemSpecD& sp, double t, double (&acc)[2][ShapeTraits<SHAPE>::NPE * 2]) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE, M = NPE * 2;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < M; j++) acc[i][j] = 0.0;
    if (sp.mode != 2) {
        const int npass = (sp.mode == 1) ? 2 : 1;
        for (int pass = 0; pass < npass; pass++) {
            const int part = (sp.mode == 1) ? pass + 1 : 0, quad = sp.quad[pass];
            const int ng = quad_count(quad);
#pragma unroll 1
            for (int q = 0; q < ng; q++) {
                double r[3], wq, det, g[2][NPE];
                quad_point(quad, q, r, wq);
                shape_grad<SHAPE>(X, r, g, det);
                const double w = det * t * wq;
                double gax = g[0][0], gay = g[1][0];
#pragma unroll
                for (int n = 1; n < NPE; n++) if (n == a) { gax = g[0][n]; gay = g[1][n]; }
                double Ba[3][2];
                strain_columns(part, gax, gay, Ba);
#pragma unroll
                for (int b = 0; b < NPE; b++) {
                    double Bb[3][2], K[2][2] = { { 0.0, 0.0 }, { 0.0, 0.0 } };
                    strain_columns(part, g[0][b], g[1][b], Bb);
                    general_block(Ba, Bb, sp.D, w, K);
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) acc[i][b * 2 + j] += K[i][j];
                }
            }
        }
        return;
    }
    // Wilson-Taylor: the two incompatible modes act as extra "nodes" with gradients h_m (wt_grad)
    double Kaa[4][4], Kad[4][M];
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) Kaa[i][j] = 0.0;
#pragma unroll
        for (int j = 0; j < M; j++) Kad[i][j] = 0.0;
    }
    const int ng = quad_count(sp.quad[0]);
#pragma unroll 1
    for (int q = 0; q < ng; q++) {
        double r[3], wq, det, g[2][NPE], h[2][2];
        quad_point(sp.quad[0], q, r, wq);
        wt_grad<SHAPE>(X, r, g, h, det);
        const double w = det * t * wq;
        double gax = g[0][0], gay = g[1][0];
#pragma unroll
        for (int n = 1; n < NPE; n++) if (n == a) { gax = g[0][n]; gay = g[1][n]; }
        double Ba[3][2], G[2][3][2];
        strain_columns(0, gax, gay, Ba);
#pragma unroll
        for (int m = 0; m < 2; m++) strain_columns(0, h[0][m], h[1][m], G[m]);
#pragma unroll
        for (int b = 0; b < NPE; b++) {
            double Bb[3][2], K[2][2] = { { 0.0, 0.0 }, { 0.0, 0.0 } };
            strain_columns(0, g[0][b], g[1][b], Bb);
            general_block(Ba, Bb, sp.D, w, K);
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) acc[i][b * 2 + j] += K[i][j];
#pragma unroll
            for (int m = 0; m < 2; m++) {               // Kead += G^T D B
                double Kg[2][2] = { { 0.0, 0.0 }, { 0.0, 0.0 } };
                general_block(G[m], Bb, sp.D, w, Kg);
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) Kad[m * 2 + i][b * 2 + j] += Kg[i][j];
            }
        }
#pragma unroll
        for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
            for (int m2 = 0; m2 < 2; m2++) {            // Keaa += G^T D G
                double Kg[2][2] = { { 0.0, 0.0 }, { 0.0, 0.0 } };
                general_block(G[m1], G[m2], sp.D, w, Kg);
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) Kaa[m1 * 2 + i][m2 * 2 + j] += Kg[i][j];
            }
    }
    // row (a, i) of Kead^T Keaa^-1 Kead = (Keaa^-T v)^T Kead with v = column a*2+i of Kead: solve Keaa^T z = v
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double z[4], A[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            double v = Kad[k][0];
#pragma unroll
            for (int c = 1; c < M; c++) if (c == a * 2 + i) v = Kad[k][c];
            z[k] = v;
#pragma unroll
            for (int j = 0; j < 4; j++) A[k][j] = Kaa[j][k];
        }
        solve4(A, z);
#pragma unroll
        for (int c = 0; c < M; c++) acc[i][c] -= z[0] * Kad[0][c] + z[1] * Kad[1][c] + z[2] * Kad[2][c] + z[3] * Kad[3][c];
    }
}

}  // namespace pf2
