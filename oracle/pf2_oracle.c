/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
 *
 * oracle/pf2_oracle.c : plain-C (C99) CPU restatement of PANSFEM2's SIMP topology-optimisation hot path.
 * It exists so that the parity tests have a checker that travels to the GPU box as SOURCE (the live reference,
 * oracle/_ref/libpf2ref.so, needs /root/reference to be rebuilt).  Each function cites the reference lines it
 * restates (paths relative to /root/reference) and keeps the reference's floating-point evaluation ORDER, so on
 * identical inputs it agrees with the reference to the last bit or two (pinned in tests/test_oracle_pinned.py
 * against the reference's golden VTKs, its MMA known-answer tests and the live reference).
 *
 * Pinning status: PINNED (golden vectors tests/golden/ *.npz generated from the reference's committed outputs
 * sample/optimize/Density_{OC,MMA,CONLIN}.vtk, sample/solid/result_linear.vtk, sample/heattransfer/{static,dynamic}.vtk,
 * sample/planestrain/result.vtk, sample/advection/AdvectionSUPG{,dynamic0,dynamic99}.vtk, sample/homogenization/result_microscopic.vtk;
 * KATs from src/Optimize/Solver/test_MMA*.cpp; live-reference dumps and the stdout of the unmodified level-set and homogenisation
 * drivers).  Per row: tests/test_oracle_pinned.py, test_families_pinned.py, test_levelset_pinned.py, test_krylov_pinned.py,
 * test_advection_pinned.py, test_homogenization_pinned.py, test_homogenization_opt_pinned.py.
 *
 * Data model: flat arrays.  coords[nnode*dim], conn[nelem*npe], nodetoglobal[nnode*ndof] (-1 = Dirichlet).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { EQ_PLANESTRAIN = 0, EQ_SOLID = 1, EQ_HEAT = 2 };
enum { FILTER_DENSITY = 0, FILTER_HEAVISIDE = 1 };
enum { OPT_OC = 0, OPT_MMA = 1, OPT_CONLIN = 2 };

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* eq codes of include/pansfem2_b200.h (PF2_EQ_CODE): phys | shape << 8 | quad << 16 | quad2 << 24; a zero field is the
 * default of the physics, so the legacy values 0, 1, 2 are codes too. */
enum { PHYS_PLANESTRAIN = 0, PHYS_SOLID = 1, PHYS_HEAT = 2, PHYS_PLANESTRESS = 3, PHYS_PLANESTRAIN_SRI = 4, PHYS_MASS = 5, PHYS_PLANESTRAIN_BBAR = 6, PHYS_MASS2 = 7, PHYS_PLANESTRAIN_WT = 8 };
enum { SHAPE_T3 = 1, SHAPE_T6, SHAPE_Q4, SHAPE_Q8, SHAPE_TET4, SHAPE_HEX8, SHAPE_HEX20 };
enum { QUAD_G1TRI = 1, QUAD_G3TRI, QUAD_G1SQ, QUAD_G4SQ, QUAD_G9SQ, QUAD_G1TET, QUAD_G8CUBE, QUAD_G27CUBE };
typedef struct { int phys, shape, quad, quad2; } orc_sel;
static orc_sel decode_eq(int eq) {
    orc_sel s = { eq & 0xff, (eq >> 8) & 0xff, (eq >> 16) & 0xff, (eq >> 24) & 0xff };
    int solid = s.phys == PHYS_SOLID;
    if (!s.shape) s.shape = solid ? SHAPE_HEX8 : SHAPE_Q4;
    int tri = s.shape == SHAPE_T3 || s.shape == SHAPE_T6;
    if (!s.quad) s.quad = tri ? QUAD_G1TRI : (s.shape == SHAPE_TET4 ? QUAD_G1TET : (solid ? QUAD_G8CUBE : QUAD_G4SQ));
    if ((s.phys == PHYS_PLANESTRAIN_SRI || s.phys == PHYS_PLANESTRAIN_BBAR) && !s.quad2) s.quad2 = tri ? QUAD_G1TRI : QUAD_G1SQ;
    return s;
}
static int ndof_of(int eq) { int phys = eq & 0xff; return phys == PHYS_SOLID ? 3 : ((phys == PHYS_HEAT || phys == PHYS_MASS) ? 1 : 2); }
static int dim_of(int eq) { return (eq & 0xff) == PHYS_SOLID ? 3 : 2; }
static int npe_of(int eq) {
    static const int n[8] = { 0, 3, 6, 4, 8, 4, 8, 20 };
    return n[decode_eq(eq).shape];
}
int orc_eq_npe(int eq) { return npe_of(eq); }
#define ORC_MAX_NPE 20
#define ORC_MAX_M 60        /* hex20 x 3 dofs */

/* ------------------------------------------------------------------------------------------------------------
 * Shape functions (dN/dr, row k = d/dr_k, npe columns) -- src/FEM/Controller/ShapeFunction.h
 *   ShapeFunction3Triangle :112-117   6Triangle :150-155   4Square :186-191   8Square :226-246
 *   4Tetrahedron :277-283             8Cubic :318-329      20Cubic :396-461
 * ---------------------------------------------------------------------------------------------------------- */
static const double HX[8] = { -1, 1, 1, -1, -1, 1, 1, -1 }, HY[8] = { -1, -1, 1, 1, -1, -1, 1, 1 }, HZ[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };

static void shape_dndr(int shape, const double* r, double* d) {
    const double r0 = r[0], r1 = r[1], r2 = r[2];
    switch (shape) {
    case SHAPE_T3:
        d[0] = 1.0; d[1] = 0.0; d[2] = -1.0;
        d[3] = 0.0; d[4] = 1.0; d[5] = -1.0;
        break;
    case SHAPE_T6:
        d[0] = 4.0 * r0 - 1.0; d[1] = 0.0; d[2] = -3.0 + 4.0 * r0 + 4.0 * r1; d[3] = 4.0 * r1; d[4] = -4.0 * r1; d[5] = 4.0 * (1.0 - 2.0 * r0 - r1);
        d[6] = 0.0; d[7] = 4.0 * r1 - 1.0; d[8] = -3.0 + 4.0 * r0 + 4.0 * r1; d[9] = 4.0 * r0; d[10] = 4.0 * (1.0 - r0 - 2.0 * r1); d[11] = -4.0 * r0;
        break;
    case SHAPE_Q4:
        d[0] = -0.25 * (1.0 - r1); d[1] = 0.25 * (1.0 - r1); d[2] = 0.25 * (1.0 + r1); d[3] = -0.25 * (1.0 + r1);
        d[4] = -0.25 * (1.0 - r0); d[5] = -0.25 * (1.0 + r0); d[6] = 0.25 * (1.0 + r0); d[7] = 0.25 * (1.0 - r0);
        break;
    case SHAPE_Q8:      /* the reference differentiates each product term by term; kept in that form */
        d[0] = 0.25 * (-(1.0 - r1) * (-r0 - r1 - 1.0) - (1.0 - r0) * (1.0 - r1));
        d[1] = 0.25 * ((1.0 - r1) * (r0 - r1 - 1.0) + (1.0 + r0) * (1.0 - r1));
        d[2] = 0.25 * ((1.0 + r1) * (r0 + r1 - 1.0) + (1.0 + r0) * (1.0 + r1));
        d[3] = 0.25 * (-(1.0 + r1) * (-r0 + r1 - 1.0) - (1.0 - r0) * (1.0 + r1));
        d[4] = -r0 * (1.0 - r1);
        d[5] = 0.5 * (1.0 + r1) * (1.0 - r1);
        d[6] = -r0 * (1.0 + r1);
        d[7] = -0.5 * (1.0 + r1) * (1.0 - r1);
        d[8] = 0.25 * (-(1.0 - r0) * (-r0 - r1 - 1.0) - (1.0 - r0) * (1.0 - r1));
        d[9] = 0.25 * (-(1.0 + r0) * (r0 - r1 - 1.0) - (1.0 + r0) * (1.0 - r1));
        d[10] = 0.25 * ((1.0 + r0) * (r0 + r1 - 1.0) + (1.0 + r0) * (1.0 + r1));
        d[11] = 0.25 * ((1.0 - r0) * (-r0 + r1 - 1.0) + (1.0 - r0) * (1.0 + r1));
        d[12] = -0.5 * (1.0 + r0) * (1.0 - r0);
        d[13] = -r1 * (1.0 + r0);
        d[14] = 0.5 * (1.0 + r0) * (1.0 - r0);
        d[15] = -r1 * (1.0 - r0);
        break;
    case SHAPE_TET4:
        for (int k = 0; k < 3; k++) for (int n = 0; n < 4; n++) d[k * 4 + n] = (n == 3) ? -1.0 : (n == k ? 1.0 : 0.0);
        break;
    case SHAPE_HEX8:
        for (int n = 0; n < 8; n++) {
            /* the reference writes e.g. -0.125*(1-r1)*(1-r2); sign*0.125*(1 + s*r) evaluates to the same doubles */
            d[n]      = HX[n] * 0.125 * (1.0 + HY[n] * r1) * (1.0 + HZ[n] * r2);
            d[8 + n]  = HY[n] * 0.125 * (1.0 + HZ[n] * r2) * (1.0 + HX[n] * r0);
            d[16 + n] = HZ[n] * 0.125 * (1.0 + HX[n] * r0) * (1.0 + HY[n] * r1);
        }
        break;
    default: {          /* SHAPE_HEX20 */
        /* corners: 0.125*sx*(1 + sy r1)(1 + sz r2)(2 sx r0 + sy r1 + sz r2 - 1) is the reference's
           +-0.125*(1 -+ r1)*(1 -+ r2)*(1 +- 2 r0 +- r1 +- r2) with the signs multiplied out */
        for (int n = 0; n < 8; n++) {
            const double a = 1.0 + HX[n] * r0, b = 1.0 + HY[n] * r1, c = 1.0 + HZ[n] * r2;
            d[n]      = 0.125 * HX[n] * b * c * (2.0 * HX[n] * r0 + HY[n] * r1 + HZ[n] * r2 - 1.0);
            d[20 + n] = 0.125 * HY[n] * a * c * (HX[n] * r0 + 2.0 * HY[n] * r1 + HZ[n] * r2 - 1.0);
            d[40 + n] = 0.125 * HZ[n] * a * b * (HX[n] * r0 + HY[n] * r1 + 2.0 * HZ[n] * r2 - 1.0);
        }
        /* mid-edge nodes in the reference's order: 8,10,12,14 on r0-edges; 9,11,13,15 on r1-edges; 16..19 on r2-edges */
        static const double e0y[4] = { -1, 1, -1, 1 }, e0z[4] = { -1, -1, 1, 1 };       /* nodes 8,10,12,14 */
        static const double e1x[4] = { 1, -1, 1, -1 }, e1z[4] = { -1, -1, 1, 1 };       /* nodes 9,11,13,15 */
        static const double e2x[4] = { -1, 1, 1, -1 }, e2y[4] = { -1, -1, 1, 1 };       /* nodes 16,17,18,19 */
        for (int q = 0; q < 4; q++) {
            int n = 8 + 2 * q;
            d[n]      = -0.5 * r0 * (1.0 + e0y[q] * r1) * (1.0 + e0z[q] * r2);
            d[20 + n] = 0.25 * e0y[q] * (1.0 - r0 * r0) * (1.0 + e0z[q] * r2);
            d[40 + n] = 0.25 * e0z[q] * (1.0 - r0 * r0) * (1.0 + e0y[q] * r1);
            n = 9 + 2 * q;
            d[n]      = 0.25 * e1x[q] * (1.0 - r1 * r1) * (1.0 + e1z[q] * r2);
            d[20 + n] = -0.5 * (1.0 + e1x[q] * r0) * r1 * (1.0 + e1z[q] * r2);
            d[40 + n] = 0.25 * e1z[q] * (1.0 + e1x[q] * r0) * (1.0 - r1 * r1);
            n = 16 + q;
            d[n]      = 0.25 * e2x[q] * (1.0 + e2y[q] * r1) * (1.0 - r2 * r2);
            d[20 + n] = 0.25 * e2y[q] * (1.0 + e2x[q] * r0) * (1.0 - r2 * r2);
            d[40 + n] = -0.5 * (1.0 + e2x[q] * r0) * (1.0 + e2y[q] * r1) * r2;
        }
        break;
    }
    }
}

/* N(r) of the 2-D shapes: ShapeFunction.h:102-108 (3Triangle), :137-146 (6Triangle), :175-182 (4Square), :211-222 (8Square) */
static void shape_n2d(int shape, const double* r, double* N) {
    const double r0 = r[0], r1 = r[1];
    switch (shape) {
    case SHAPE_T3: N[0] = r0; N[1] = r1; N[2] = 1.0 - r0 - r1; break;
    case SHAPE_T6:
        N[0] = r0 * (2.0 * r0 - 1.0); N[1] = r1 * (2.0 * r1 - 1.0); N[2] = (1.0 - r0 - r1) * (1.0 - 2.0 * r0 - 2.0 * r1);
        N[3] = 4.0 * r0 * r1; N[4] = 4.0 * r1 * (1.0 - r0 - r1); N[5] = 4.0 * (1.0 - r0 - r1) * r0;
        break;
    case SHAPE_Q4:
        N[0] = 0.25 * (1.0 - r0) * (1.0 - r1); N[1] = 0.25 * (1.0 + r0) * (1.0 - r1);
        N[2] = 0.25 * (1.0 + r0) * (1.0 + r1); N[3] = 0.25 * (1.0 - r0) * (1.0 + r1);
        break;
    default:    /* SHAPE_Q8 */
        N[0] = 0.25 * (1.0 - r0) * (1.0 - r1) * (-r0 - r1 - 1.0); N[1] = 0.25 * (1.0 + r0) * (1.0 - r1) * (r0 - r1 - 1.0);
        N[2] = 0.25 * (1.0 + r0) * (1.0 + r1) * (r0 + r1 - 1.0);  N[3] = 0.25 * (1.0 - r0) * (1.0 + r1) * (-r0 + r1 - 1.0);
        N[4] = 0.5 * (1.0 - r0) * (1.0 + r0) * (1.0 - r1); N[5] = 0.5 * (1.0 + r0) * (1.0 + r1) * (1.0 - r1);
        N[6] = 0.5 * (1.0 + r0) * (1.0 - r0) * (1.0 + r1); N[7] = 0.5 * (1.0 - r0) * (1.0 + r1) * (1.0 - r1);
        break;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Integration rules: point g and per-axis weights -- src/FEM/Controller/GaussIntegration.h
 *   Gauss1Triangle :72-82  Gauss3Triangle :94-107  Gauss1Square :120-130  Gauss4Square :142-157 (order (-,-),(+,-),(-,+),(+,+))
 *   Gauss9Square :170-195  Gauss1Tetrahedron :208-218  Gauss8Cubic :230-253 (ordered like the hex8 nodes)  Gauss27Cubic :266-327
 * ---------------------------------------------------------------------------------------------------------- */
static int quad_count(int quad) {
    static const int n[9] = { 0, 1, 3, 1, 4, 9, 1, 8, 27 };
    return n[quad];
}
static void quad_point(int quad, int g, double* r, double* w) {
    const double a3 = 1.0 / sqrt(3.0), s35 = sqrt(3.0 / 5.0);
    r[0] = r[1] = r[2] = 0.0; w[0] = w[1] = w[2] = 1.0;
    switch (quad) {
    case QUAD_G1TRI: r[0] = 1.0 / 3.0; r[1] = 1.0 / 3.0; w[0] = w[1] = 1.0 / sqrt(2.0); break;
    case QUAD_G3TRI: r[0] = (g == 1) ? 2.0 / 3.0 : 1.0 / 6.0; r[1] = (g == 2) ? 2.0 / 3.0 : 1.0 / 6.0; w[0] = w[1] = 1.0 / sqrt(6.0); break;
    case QUAD_G1SQ: w[0] = w[1] = 2.0; break;
    case QUAD_G4SQ: { static const int s[4][2] = { { -1, -1 }, { 1, -1 }, { -1, 1 }, { 1, 1 } }; r[0] = s[g][0] * a3; r[1] = s[g][1] * a3; break; }
    case QUAD_G9SQ: {
        int i = g % 3, j = g / 3;
        r[0] = (i - 1) * s35; r[1] = (j - 1) * s35;
        w[0] = (i == 1) ? 8.0 / 9.0 : 5.0 / 9.0; w[1] = (j == 1) ? 8.0 / 9.0 : 5.0 / 9.0;
        break;
    }
    case QUAD_G1TET: r[0] = r[1] = r[2] = 1.0 / 4.0; w[0] = w[1] = w[2] = 1.0 / cbrt(6.0); break;
    case QUAD_G8CUBE: r[0] = HX[g] * a3; r[1] = HY[g] * a3; r[2] = HZ[g] * a3; break;
    default: {
        int i = g % 3, j = (g / 3) % 3, k = g / 9;
        r[0] = (i - 1) * s35; r[1] = (j - 1) * s35; r[2] = (k - 1) * s35;
        w[0] = (i == 1) ? 8.0 / 9.0 : 5.0 / 9.0; w[1] = (j == 1) ? 8.0 / 9.0 : 5.0 / 9.0; w[2] = (k == 1) ? 8.0 / 9.0 : 5.0 / 9.0;
        break;
    }
    }
}

/* C(m x n) = A(m x k) * B(k x n), Matrix<T>::operator*  src/LinearAlgebra/Models/Matrix.h:243-255 */
static void matmul(int m, int k, int n, const double* A, const double* B, double* Cm) {
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double v = 0.0;
            for (int l = 0; l < k; l++) v += A[i * k + l] * B[l * n + j];
            Cm[i * n + j] = v;
        }
}
/* Determinant (Matrix.h:326-342) and Inverse = adjugate/det (Matrix.h:346-360) for d = 2, 3 */
static double det_d(int d, const double* v) {
    if (d == 2) return v[0] * v[3] - v[1] * v[2];
    return -v[8] * v[1] * v[3] - v[7] * v[5] * v[0] - v[2] * v[4] * v[6] + v[6] * v[1] * v[5] + v[7] * v[3] * v[2] + v[0] * v[4] * v[8];
}
static void inv_d(int d, const double* v, double* inv) {
    double det = det_d(d, v);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            /* Cofactor(j, i): drop row j and column i */
            double minor[4];
            int c = 0;
            for (int a = 0; a < d; a++) {
                if (a == j) continue;
                for (int b = 0; b < d; b++) {
                    if (b == i) continue;
                    minor[c++] = v[a * d + b];
                }
            }
            double cd = (d == 2) ? minor[0] : (minor[0] * minor[3] - minor[1] * minor[2]);
            inv[i * d + j] = (((i + j) & 1) ? -1.0 : 1.0) * cd;
        }
    for (int i = 0; i < d * d; i++) inv[i] /= det;
}

/* Matrix<T>::Determinant for n > 3 (Laplace expansion along column 0, Matrix.h:336-341) and Inverse = adjugate / det (:346-360) */
static double det_n(int n, const double* v) {
    if (n <= 3) return (n == 1) ? v[0] : det_d(n, v);
    double value = 0.0, minor[9];
    for (int i = 0; i < n; i++) {
        int c = 0;
        for (int a = 0; a < n; a++) { if (a == i) continue; for (int b = 1; b < n; b++) minor[c++] = v[a * n + b]; }
        value += (((i & 1) ? -1.0 : 1.0)) * v[i * n] * det_n(n - 1, minor);
    }
    return value;
}
static void inv4(const double* v, double* inv) {
    double minor[9];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            int c = 0;      /* Cofactor(j, i): drop row j, column i */
            for (int a = 0; a < 4; a++) { if (a == j) continue; for (int b = 0; b < 4; b++) { if (b == i) continue; minor[c++] = v[a * 4 + b]; } }
            inv[i * 4 + j] = ((((i + j) & 1) ? -1.0 : 1.0)) * det_d(3, minor);
        }
    const double det = det_n(4, v);
    for (int i = 0; i < 16; i++) inv[i] /= det;
}

/* PlaneStrainStiffnessWilsonTaylor  src/FEM/Equation/PlaneStrain.h:189-243: incompatible modes P = (1 - r0^2, 1 - r1^2), statically condensed */
static void wilson_taylor_d(int shape, int quad, int npe, const double* xe, const double* D, double t, double* Ke);
static void wilson_taylor(int shape, int quad, int npe, const double* xe, double E, double V, double t, double* Ke) {
    double D[9] = { 1.0 - V, V, 0, V, 1.0 - V, 0, 0, 0, 0.5 * (1.0 - 2.0 * V) };
    const double f = E / ((1.0 - 2.0 * V) * (1.0 + V));
    for (int i = 0; i < 9; i++) D[i] *= f;
    wilson_taylor_d(shape, quad, npe, xe, D, t, Ke);
}
/* PlaneStrainStiffnessWilsonTaylor (PlaneStrain.h:189-243) and PlaneStiffnessWilsonTaylor (Homogenization.h:230-280) are the same
 * statements around a different D */
static void wilson_taylor_d(int shape, int quad, int npe, const double* xe, const double* D, double t, double* Ke) {
    const int m = 2 * npe, ng = quad_count(quad);
    double Keaa[16], Kead[4 * ORC_MAX_M];
    memset(Keaa, 0, sizeof Keaa); memset(Kead, 0, sizeof(double) * 4 * m); memset(Ke, 0, sizeof(double) * m * m);
    for (int g = 0; g < ng; g++) {
        double r[3], w[3], dNdr[2 * ORC_MAX_NPE], dXdr[4], inv[4], dNdX[2 * ORC_MAX_NPE];
        quad_point(quad, g, r, w);
        shape_dndr(shape, r, dNdr);
        matmul(2, npe, 2, dNdr, xe, dXdr);
        const double J = det_d(2, dXdr);
        inv_d(2, dXdr, inv);
        matmul(2, 2, npe, inv, dNdr, dNdX);
        double B[3 * ORC_MAX_M], Bt[ORC_MAX_M * 3], BtD[ORC_MAX_M * 3], BtDB[ORC_MAX_M * ORC_MAX_M];
        memset(B, 0, sizeof(double) * 3 * m);
        for (int n = 0; n < npe; n++) {
            B[0 * m + 2 * n] = dNdX[n]; B[1 * m + 2 * n + 1] = dNdX[npe + n];
            B[2 * m + 2 * n] = dNdX[npe + n]; B[2 * m + 2 * n + 1] = dNdX[n];
        }
        const double dPdr[4] = { -2.0 * r[0], 0.0, 0.0, -2.0 * r[1] };
        double dPdX[4];
        matmul(2, 2, 2, inv, dPdr, dPdX);
        const double G[12] = { dPdX[0], 0.0, dPdX[1], 0.0,
                               0.0, dPdX[2], 0.0, dPdX[3],
                               dPdX[2], dPdX[0], dPdX[3], dPdX[1] };
        double Gt[12], GtD[12], GtDG[16], GtDB[4 * ORC_MAX_M];
        for (int i = 0; i < 3; i++) for (int j = 0; j < m; j++) Bt[j * 3 + i] = B[i * m + j];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) Gt[j * 3 + i] = G[i * 4 + j];
        matmul(m, 3, 3, Bt, D, BtD); matmul(m, 3, m, BtD, B, BtDB);
        matmul(4, 3, 3, Gt, D, GtD); matmul(4, 3, 4, GtD, G, GtDG); matmul(4, 3, m, GtD, B, GtDB);
        for (int i = 0; i < m * m; i++) Ke[i] += BtDB[i] * J * t * w[0] * w[1];
        for (int i = 0; i < 16; i++) Keaa[i] += GtDG[i] * J * t * w[0] * w[1];
        for (int i = 0; i < 4 * m; i++) Kead[i] += GtDB[i] * J * t * w[0] * w[1];
    }
    double Kinv[16], KdaT[ORC_MAX_M * 4], T1[ORC_MAX_M * 4], T2[ORC_MAX_M * ORC_MAX_M];
    inv4(Keaa, Kinv);
    for (int i = 0; i < 4; i++) for (int j = 0; j < m; j++) KdaT[j * 4 + i] = Kead[i * m + j];
    matmul(m, 4, 4, KdaT, Kinv, T1);            /* (Kead^T Keaa^-1) Kead, left to right */
    matmul(m, 4, m, T1, Kead, T2);
    for (int i = 0; i < m * m; i++) Ke[i] -= T2[i];
}

/* ------------------------------------------------------------------------------------------------------------
 * Element matrices for any <Equation, SF, IC> selection.
 * PlaneStrainStiffness           src/FEM/Equation/PlaneStrain.h:21-58
 * PlaneStrainStiffnessSRI        src/FEM/Equation/PlaneStrain.h:63-125  (volumetric D with ICV, then deviatoric D with ICD)
 * PlaneStressStiffness           src/FEM/Equation/PlaneStress.h:21-58
 * SolidLinearIsotropicElastic    src/FEM/Equation/Solid.h:21-64
 * HeatTransfer                   src/FEM/Equation/HeatTransfer.h:20-43
 * xe: npe x dim coordinates of the element's nodes.  Ke: (npe*ndof)^2 row-major.
 * Accumulation order kept: Ke += ((((B^T D) B) J) t) w0 w1 [w2]   (PlaneStrain.h:56, Solid.h:62, HeatTransfer.h:41)
 * ---------------------------------------------------------------------------------------------------------- */
/* bmode: 0 = the standard B; 1 / 2 = Bvol / Bdev of PlaneStrainStiffnessBbar (PlaneStrain.h:150-156, 166-172) */
static void accumulate_rule(int phys, int shape, int quad, int dim, int npe, int ndof, const double* xe, const double* D, int ns,
                            double alpha, double t, double* Ke, int bmode) {
    const int m = npe * ndof, ng = quad_count(quad);
    for (int g = 0; g < ng; g++) {
        double r[3], w[3], dNdr[3 * ORC_MAX_NPE], dXdr[9], inv[9], dNdX[3 * ORC_MAX_NPE];
        quad_point(quad, g, r, w);
        shape_dndr(shape, r, dNdr);
        matmul(dim, npe, dim, dNdr, xe, dXdr);
        double J = det_d(dim, dXdr);
        if (phys == PHYS_MASS) {            /* ReactionDiffusionConsistentMass  ReactionDiffusion.h:40-47: N N^T J w0 w1 */
            double N[ORC_MAX_NPE];
            shape_n2d(shape, r, N);
            for (int i = 0; i < npe; i++) for (int j = 0; j < npe; j++) Ke[i * npe + j] += N[i] * N[j] * J * w[0] * w[1];
            continue;
        }
        if (phys == PHYS_MASS2) {           /* PlaneStrainMass  PlaneStrain.h:397-407: B = [N 0; 0 N], Me += B^T B J rho t w0 w1 */
            double N[ORC_MAX_NPE], Bm[2 * ORC_MAX_M], Bmt[ORC_MAX_M * 2], BtB[ORC_MAX_M * ORC_MAX_M];
            shape_n2d(shape, r, N);
            memset(Bm, 0, sizeof(double) * 2 * m);
            for (int n = 0; n < npe; n++) { Bm[0 * m + 2 * n] = N[n]; Bm[1 * m + 2 * n + 1] = N[n]; }
            for (int i = 0; i < 2; i++) for (int j = 0; j < m; j++) Bmt[j * 2 + i] = Bm[i * m + j];
            matmul(m, 2, m, Bmt, Bm, BtB);
            for (int i = 0; i < m * m; i++) Ke[i] += BtB[i] * J * alpha * t * w[0] * w[1];
            continue;
        }
        inv_d(dim, dXdr, inv);
        matmul(dim, dim, npe, inv, dNdr, dNdX);
        double B[6 * ORC_MAX_M], Bt[ORC_MAX_M * 6], BtD[ORC_MAX_M * 6], BtDB[ORC_MAX_M * ORC_MAX_M];
        memset(B, 0, sizeof(double) * (size_t)ns * m);
        if (phys == PHYS_SOLID) {
            for (int n = 0; n < npe; n++) {
                double dx = dNdX[n], dy = dNdX[npe + n], dz = dNdX[2 * npe + n];
                B[0 * m + 3 * n] = dx; B[1 * m + 3 * n + 1] = dy; B[2 * m + 3 * n + 2] = dz;
                B[3 * m + 3 * n] = dy; B[3 * m + 3 * n + 1] = dx;            /* gamma_xy */
                B[4 * m + 3 * n + 1] = dz; B[4 * m + 3 * n + 2] = dy;        /* gamma_yz */
                B[5 * m + 3 * n] = dz; B[5 * m + 3 * n + 2] = dx;            /* gamma_zx */
            }
        } else if (phys == PHYS_HEAT) {
            memcpy(B, dNdX, sizeof(double) * dim * npe);
        } else if (bmode == 1) {
            for (int n = 0; n < npe; n++) {
                B[0 * m + 2 * n] = 0.5 * dNdX[n]; B[0 * m + 2 * n + 1] = 0.5 * dNdX[npe + n];
                B[1 * m + 2 * n] = 0.5 * dNdX[n]; B[1 * m + 2 * n + 1] = 0.5 * dNdX[npe + n];
            }
        } else if (bmode == 2) {
            for (int n = 0; n < npe; n++) {
                B[0 * m + 2 * n] = 0.5 * dNdX[n];  B[0 * m + 2 * n + 1] = -0.5 * dNdX[npe + n];
                B[1 * m + 2 * n] = -0.5 * dNdX[n]; B[1 * m + 2 * n + 1] = 0.5 * dNdX[npe + n];
                B[2 * m + 2 * n] = dNdX[npe + n];  B[2 * m + 2 * n + 1] = dNdX[n];
            }
        } else {
            for (int n = 0; n < npe; n++) {
                B[0 * m + 2 * n] = dNdX[n];             /* row 0: dN/dx */
                B[1 * m + 2 * n + 1] = dNdX[npe + n];   /* row 1: dN/dy */
                B[2 * m + 2 * n] = dNdX[npe + n]; B[2 * m + 2 * n + 1] = dNdX[n];
            }
        }
        for (int i = 0; i < ns; i++) for (int j = 0; j < m; j++) Bt[j * ns + i] = B[i * m + j];
        if (phys == PHYS_HEAT) {
            matmul(m, ns, m, Bt, B, BtDB);
            for (int i = 0; i < m * m; i++) Ke[i] += BtDB[i] * J * alpha * t * w[0] * w[1];
        } else {
            matmul(m, ns, ns, Bt, D, BtD);
            matmul(m, ns, m, BtD, B, BtDB);
            if (phys == PHYS_SOLID) for (int i = 0; i < m * m; i++) Ke[i] += BtDB[i] * J * w[0] * w[1] * w[2];
            else for (int i = 0; i < m * m; i++) Ke[i] += BtDB[i] * J * t * w[0] * w[1];
        }
    }
}

void orc_element_matrix(int eq, const double* xe, double E, double V, double t, double* Ke) {
    const orc_sel s = decode_eq(eq);
    const int dim = dim_of(eq), npe = npe_of(eq), ndof = ndof_of(eq), m = npe * ndof;
    const int ns = (s.phys == PHYS_SOLID) ? 6 : ((s.phys == PHYS_HEAT || s.phys == PHYS_MASS) ? dim : 3);
    double D[36];
    memset(D, 0, sizeof D);
    memset(Ke, 0, sizeof(double) * m * m);
    if (s.phys == PHYS_PLANESTRAIN) {
        D[0] = 1.0 - V; D[1] = V; D[3] = V; D[4] = 1.0 - V; D[8] = 0.5 * (1.0 - 2.0 * V);
        double f = E / ((1.0 - 2.0 * V) * (1.0 + V));
        for (int i = 0; i < 9; i++) D[i] *= f;
    } else if (s.phys == PHYS_PLANESTRESS) {
        D[0] = 1.0; D[1] = V; D[3] = V; D[4] = 1.0; D[8] = 0.5 * (1.0 - V);
        double f = E / ((1.0 - V) * (1.0 + V));
        for (int i = 0; i < 9; i++) D[i] *= f;
    } else if (s.phys == PHYS_SOLID) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) D[i * 6 + j] = (i == j) ? 1.0 - V : V;
        for (int i = 3; i < 6; i++) D[i * 6 + i] = 0.5 * (1.0 - 2.0 * V);
        double f = E / ((1.0 + V) * (1.0 - 2.0 * V));
        for (int i = 0; i < 36; i++) D[i] *= f;
    } else if (s.phys == PHYS_PLANESTRAIN_SRI) {
        D[0] = 1.0; D[1] = 1.0; D[3] = 1.0; D[4] = 1.0;
        double f = E / (3.0 * (1.0 - 2.0 * V));
        for (int i = 0; i < 9; i++) D[i] *= f;
        accumulate_rule(s.phys, s.shape, s.quad2, dim, npe, ndof, xe, D, ns, E, t, Ke, 0);
        memset(D, 0, sizeof D);
        D[0] = 4.0; D[1] = -2.0; D[3] = -2.0; D[4] = 4.0; D[8] = 3.0;
        f = E / (6.0 * (1.0 + V));
        for (int i = 0; i < 9; i++) D[i] *= f;
    }
    if (s.phys == PHYS_PLANESTRAIN_WT) { wilson_taylor(s.shape, s.quad, npe, xe, E, V, t, Ke); return; }
    if (s.phys == PHYS_PLANESTRAIN_BBAR) {
        D[0] = 1.0 - V; D[1] = V; D[3] = V; D[4] = 1.0 - V; D[8] = 0.5 * (1.0 - 2.0 * V);
        double f = E / ((1.0 - 2.0 * V) * (1.0 + V));
        for (int i = 0; i < 9; i++) D[i] *= f;
        accumulate_rule(s.phys, s.shape, s.quad2, dim, npe, ndof, xe, D, ns, E, t, Ke, 1);
        accumulate_rule(s.phys, s.shape, s.quad, dim, npe, ndof, xe, D, ns, E, t, Ke, 2);
        return;
    }
    accumulate_rule(s.phys, s.shape, s.quad, dim, npe, ndof, xe, D, ns, E, t, Ke, 0);     /* heat / mass: E carries the coefficient */
    if (s.phys == PHYS_MASS) for (int i = 0; i < m * m; i++) Ke[i] *= E * t;            /* the reference's mass has no coefficient */
}

/* PlaneStiffness / PlaneStiffnessBbar / PlaneStiffnessWilsonTaylor   src/FEM/Equation/Homogenization.h:141-166, 170-226, 230-280:
 * the plane routines with a caller-supplied 3 x 3 constitutive matrix D (row-major).  eq: phys 10 / 11 / 12 of include/pansfem2_b200.h. */
void orc_element_matrix_d(int eq, const double* xe, const double* D, double t, double* Ke) {
    const orc_sel s0 = decode_eq(eq);
    orc_sel s = s0;
    const int npe = npe_of(eq), m = 2 * npe;
    const int tri = s.shape == SHAPE_T3 || s.shape == SHAPE_T6;
    if (s.phys == 11 && !s.quad2) s.quad2 = tri ? QUAD_G1TRI : QUAD_G1SQ;
    memset(Ke, 0, sizeof(double) * m * m);
    if (s.phys == 12) { wilson_taylor_d(s.shape, s.quad, npe, xe, D, t, Ke); return; }
    if (s.phys == 11) {
        accumulate_rule(PHYS_PLANESTRAIN_BBAR, s.shape, s.quad2, 2, npe, 2, xe, D, 3, 1.0, t, Ke, 1);
        accumulate_rule(PHYS_PLANESTRAIN_BBAR, s.shape, s.quad, 2, npe, 2, xe, D, 3, 1.0, t, Ke, 2);
        return;
    }
    accumulate_rule(PHYS_PLANESTRAIN, s.shape, s.quad, 2, npe, 2, xe, D, 3, 1.0, t, Ke, 0);
}

/* ------------------------------------------------------------------------------------------------------------
 * Boundary conditions and numbering.
 * SetDirichlet   src/FEM/Controller/BoundaryCondition.h:20-25  (mark -1, store value in u)
 * Renumbering    src/FEM/Controller/Assembling.h:175-186       (node-major, dof-minor running index)
 * ---------------------------------------------------------------------------------------------------------- */
int orc_dofmap(int nnode, int ndof, int nfixed, const int* fnode, const int* fdof, const double* fval,
               int* nodetoglobal, double* ufixed /* nnode*ndof, may be NULL */) {
    size_t n = (size_t)nnode * ndof;
    memset(nodetoglobal, 0, n * sizeof(int));
    if (ufixed) memset(ufixed, 0, n * sizeof(double));
    for (int i = 0; i < nfixed; i++) {
        nodetoglobal[(size_t)fnode[i] * ndof + fdof[i]] = -1;
        if (ufixed) ufixed[(size_t)fnode[i] * ndof + fdof[i]] = fval[i];
    }
    int k = 0;
    for (size_t i = 0; i < n; i++) if (nodetoglobal[i] != -1) nodetoglobal[i] = k++;
    return k;
}

/* ------------------------------------------------------------------------------------------------------------
 * Assembly.
 * Assembling(K,F,u,Ke,...)  src/FEM/Controller/Assembling.h:47-66  with LILCSR::get/set (LILCSR.h:92-114)
 * CSR(LILCSR&)              src/LinearAlgebra/Models/CSR.h:93-105  (per-row sort by column)
 * Assembling(F,q,...)       src/FEM/Controller/Assembling.h:152-158 (nodal loads)
 * The reference inserts every Ke entry (zeros too) so the pattern is the full element connectivity, and each stored
 * value is the left-to-right sum over elements in ascending element order starting from T() -- reproduced here by
 * building the sorted pattern first and then accumulating in element order (same adds, same order, no LIL scans).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    int n;          /* KDEGREE */
    int* indptr;    /* n+1 */
    int* indices;   /* nnz */
    double* data;   /* nnz */
    double* F;      /* n */
} orc_system;

static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

static int find_col(const orc_system* S, int row, int col) {
    int lo = S->indptr[row], hi = S->indptr[row + 1] - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (S->indices[mid] == col) return mid;
        if (S->indices[mid] < col) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

orc_system* orc_pattern(int nnode, int ndof, int npe, int nelem, const int* conn, const int* nodetoglobal, int kdegree) {
    /* node -> adjacent nodes (sorted unique), via node -> elements */
    int* cnt = (int*)calloc((size_t)nnode + 1, sizeof(int));
    for (size_t i = 0; i < (size_t)nelem * npe; i++) cnt[conn[i] + 1]++;
    for (int i = 0; i < nnode; i++) cnt[i + 1] += cnt[i];
    int* n2e = (int*)malloc(sizeof(int) * (size_t)nelem * npe);
    int* cur = (int*)malloc(sizeof(int) * (size_t)nnode);
    memcpy(cur, cnt, sizeof(int) * (size_t)nnode);
    for (int e = 0; e < nelem; e++) for (int a = 0; a < npe; a++) n2e[cur[conn[(size_t)e * npe + a]]++] = e;
    orc_system* S = (orc_system*)calloc(1, sizeof(orc_system));
    S->n = kdegree;
    S->indptr = (int*)calloc((size_t)kdegree + 1, sizeof(int));
    int maxadj = 0;
    for (int i = 0; i < nnode; i++) if (cnt[i + 1] - cnt[i] > maxadj) maxadj = cnt[i + 1] - cnt[i];
    int* tmp = (int*)malloc(sizeof(int) * (size_t)(maxadj * npe + 1));
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < nnode; i++) {
            int k = 0;
            for (int q = cnt[i]; q < cnt[i + 1]; q++) for (int a = 0; a < npe; a++) tmp[k++] = conn[(size_t)n2e[q] * npe + a];
            qsort(tmp, k, sizeof(int), cmp_int);
            int u = 0;
            for (int q = 0; q < k; q++) if (q == 0 || tmp[q] != tmp[q - 1]) tmp[u++] = tmp[q];
            int rowlen = 0;
            for (int q = 0; q < u; q++) for (int d = 0; d < ndof; d++) if (nodetoglobal[(size_t)tmp[q] * ndof + d] != -1) rowlen++;
            for (int d = 0; d < ndof; d++) {
                int r = nodetoglobal[(size_t)i * ndof + d];
                if (r == -1) continue;
                if (pass == 0) S->indptr[r + 1] = rowlen;
                else {
                    int p = S->indptr[r];
                    for (int q = 0; q < u; q++) for (int dd = 0; dd < ndof; dd++) {
                        int c = nodetoglobal[(size_t)tmp[q] * ndof + dd];
                        if (c != -1) S->indices[p++] = c;
                    }
                }
            }
        }
        if (pass == 0) {
            for (int r = 0; r < kdegree; r++) S->indptr[r + 1] += S->indptr[r];
            S->indices = (int*)malloc(sizeof(int) * (size_t)S->indptr[kdegree]);
            S->data = (double*)calloc((size_t)S->indptr[kdegree], sizeof(double));
            S->F = (double*)calloc((size_t)kdegree, sizeof(double));
        }
    }
    free(tmp); free(cur); free(n2e); free(cnt);
    return S;
}

void orc_system_free(orc_system* S) {
    if (!S) return;
    free(S->indptr); free(S->indices); free(S->data); free(S->F); free(S);
}
int orc_system_rows(const orc_system* S) { return S->n; }
long long orc_system_nnz(const orc_system* S) { return S->indptr[S->n]; }
void orc_system_get(const orc_system* S, int* indptr, int* indices, double* data, double* F) {
    if (indptr) memcpy(indptr, S->indptr, sizeof(int) * ((size_t)S->n + 1));
    if (indices) memcpy(indices, S->indices, sizeof(int) * (size_t)S->indptr[S->n]);
    if (data) memcpy(data, S->data, sizeof(double) * (size_t)S->indptr[S->n]);
    if (F) memcpy(F, S->F, sizeof(double) * (size_t)S->n);
}
orc_system* orc_system_from_csr(int n, const int* indptr, const int* indices, const double* data) {
    orc_system* S = (orc_system*)calloc(1, sizeof(orc_system));
    S->n = n;
    S->indptr = (int*)malloc(sizeof(int) * ((size_t)n + 1));
    memcpy(S->indptr, indptr, sizeof(int) * ((size_t)n + 1));
    size_t nnz = (size_t)indptr[n];
    S->indices = (int*)malloc(sizeof(int) * nnz);
    memcpy(S->indices, indices, sizeof(int) * nnz);
    S->data = (double*)malloc(sizeof(double) * nnz);
    memcpy(S->data, data, sizeof(double) * nnz);
    S->F = (double*)calloc((size_t)n, sizeof(double));
    return S;
}

/* numeric assembly into an existing pattern; Emod = per-element modulus; times[2] = {element, scatter} */
void orc_assemble_numeric(orc_system* S, int eq, const double* coords, int nelem, const int* conn,
                          const int* nodetoglobal, const double* ufixed, const double* Emod, double V, double t,
                          int nload, const int* lnode, const int* ldof, const double* lval, double* times) {
    const int dim = dim_of(eq), npe = npe_of(eq), ndof = ndof_of(eq), m = npe * ndof;
    memset(S->data, 0, sizeof(double) * (size_t)S->indptr[S->n]);
    memset(S->F, 0, sizeof(double) * (size_t)S->n);
    double Ke[ORC_MAX_M * ORC_MAX_M], xe[3 * ORC_MAX_NPE], te = 0, ts = 0;
    for (int e = 0; e < nelem; e++) {
        const int* el = conn + (size_t)e * npe;
        double t0 = now_s();
        for (int a = 0; a < npe; a++) for (int d = 0; d < dim; d++) xe[a * dim + d] = coords[(size_t)el[a] * dim + d];
        orc_element_matrix(eq, xe, Emod[e], V, t, Ke);
        double t1 = now_s();
        for (int i = 0; i < npe; i++) for (int di = 0; di < ndof; di++) {
            int r = nodetoglobal[(size_t)el[i] * ndof + di];
            if (r == -1) continue;
            for (int j = 0; j < npe; j++) for (int dj = 0; dj < ndof; dj++) {
                int c = nodetoglobal[(size_t)el[j] * ndof + dj];
                double v = Ke[(i * ndof + di) * m + (j * ndof + dj)];
                if (c != -1) S->data[find_col(S, r, c)] += v;                       /* Assembling.h:55 */
                else S->F[r] -= v * ufixed[(size_t)el[j] * ndof + dj];               /* Assembling.h:59 */
            }
        }
        double t2 = now_s();
        te += t1 - t0; ts += t2 - t1;
    }
    for (int i = 0; i < nload; i++) {                                                /* Assembling.h:152-158 */
        int r = nodetoglobal[(size_t)lnode[i] * ndof + ldof[i]];
        if (r != -1) S->F[r] += lval[i];
    }
    if (times) { times[0] = te; times[1] = ts; }
}

/* ------------------------------------------------------------------------------------------------------------
 * Advection-diffusion element family   src/FEM/Equation/Advection.h
 *   Advection :19-43   AdvectionSUPG :47-87   AdvectionShockCapturing :91-131   Diffusion :135-157   Mass :161-184   MassSUPG :188-228
 * one dof per node, 2-D shapes; every product below is evaluated in the order the reference's Matrix / Vector operators evaluate
 * the expression written there (left to right), so the result is bit-identical to the compiled reference.
 * terms: 1 Advection, 2 Diffusion, 4 AdvectionSUPG, 8 AdvectionShockCapturing, 16 Mass, 32 MassSUPG.
 * ---------------------------------------------------------------------------------------------------------- */
enum { ADV_A = 1, ADV_D = 2, ADV_S = 4, ADV_SC = 8, ADV_M = 16, ADV_MS = 32 };
static int npe_of_shape(int shape) { static const int n[8] = { 0, 3, 6, 4, 8, 4, 8, 20 }; return n[shape]; }

static void adv_one_term(int shape, int quad, int term, const double* xe, double ax, double ay, double k, double* Ke) {
    const int npe = npe_of_shape(shape), ng = quad_count(quad);
    const double a[2] = { ax, ay };
    memset(Ke, 0, sizeof(double) * npe * npe);
    for (int g = 0; g < ng; g++) {
        double r[3], w[3], N[ORC_MAX_NPE], dNdr[3 * ORC_MAX_NPE], dXdr[9], inv[9], dNdX[3 * ORC_MAX_NPE];
        quad_point(quad, g, r, w);
        shape_n2d(shape, r, N);
        shape_dndr(shape, r, dNdr);
        matmul(2, npe, 2, dNdr, xe, dXdr);
        double J = det_d(2, dXdr);
        inv_d(2, dXdr, inv);
        matmul(2, 2, npe, inv, dNdr, dNdX);
        double tau = 0.0;
        if (term == ADV_S || term == ADV_SC || term == ADV_MS) {        /* Advection.h:69-82, 113-126, 210-223 */
            double norm = 0.0;
            norm += pow(a[0], 2.0); norm += pow(a[1], 2.0);             /* Vector<T>::Norm  Vector.h:297-303 */
            norm = sqrt(norm);
            double sum = 0.0;
            for (int i = 0; i < npe; i++) {
                double v = 0.0;                                         /* dNdX.Transpose()*a */
                for (int l = 0; l < 2; l++) v += dNdX[l * npe + i] * a[l];
                sum += fabs(v) / norm;
            }
            double he = 2.0 / sum;
            double alpha = 0.5 * norm * he / k;
            tau = (term == ADV_SC) ? 0.5 * norm * he : 0.5 * he / norm;
            if (alpha <= 3.0) tau *= alpha / 3.0; else tau *= 1.0;
        }
        for (int i = 0; i < npe; i++) {
            /* the row-i factors that precede the last product */
            double m2[2] = { 0.0, 0.0 }, v1 = 0.0;
            if (term == ADV_A) { m2[0] = N[i] * a[0]; m2[1] = N[i] * a[1]; }                            /* N*c.Transpose() */
            else if (term == ADV_D) { m2[0] = dNdX[i] * k; m2[1] = dNdX[npe + i] * k; }                 /* _k*dNdX.Transpose() */
            else if (term == ADV_SC) { m2[0] = dNdX[i] * tau; m2[1] = dNdX[npe + i] * tau; }            /* tau*dNdX.Transpose() */
            else if (term == ADV_S || term == ADV_MS) {
                for (int l = 0; l < 2; l++) v1 += (dNdX[l * npe + i] * tau) * a[l];                     /* (tau*dNdX.Transpose())*a */
                m2[0] = v1 * a[0]; m2[1] = v1 * a[1];                                                   /* ...*a.Transpose() */
            }
            for (int j = 0; j < npe; j++) {
                double v;
                if (term == ADV_M) v = N[i] * N[j];                                                     /* N*N.Transpose() */
                else if (term == ADV_MS) v = v1 * N[j];                                                 /* ...*N.Transpose() */
                else { v = 0.0; for (int l = 0; l < 2; l++) v += m2[l] * dNdX[l * npe + j]; }           /* ...*dNdX */
                Ke[i * npe + j] += v * J * w[0] * w[1];
            }
        }
    }
}

/* sum of the selected terms of `group` in the samples' order; returns 0 when none is selected (Ke zeroed) */
static int adv_sum(int shape, int quad, int terms, int group, const double* xe, double ax, double ay, double k, double* Ke) {
    static const int order[6] = { ADV_A, ADV_D, ADV_S, ADV_SC, ADV_M, ADV_MS };
    const int npe = npe_of_shape(shape);
    double P[ORC_MAX_NPE * ORC_MAX_NPE];
    int any = 0;
    memset(Ke, 0, sizeof(double) * npe * npe);
    for (int q = 0; q < 6; q++) {
        if (!(terms & group & order[q])) continue;
        adv_one_term(shape, quad, order[q], xe, ax, ay, k, P);
        if (!any) memcpy(Ke, P, sizeof(double) * npe * npe);
        else for (int i = 0; i < npe * npe; i++) Ke[i] = Ke[i] + P[i];
        any = 1;
    }
    return any;
}

void orc_advdiff_element(int shape, int quad, int terms, const double* xe, double ax, double ay, double k, double* Ke) {
    adv_sum(shape, quad, terms, 63, xe, ax, ay, k, Ke);
}

/* dt == 0: sample_advectiondiffusion_static.cpp:42-55 (Ke = A + B + C [+ D]); dt > 0: one step of
 * sample_advectiondiffusion_dynamic.cpp:51-70 (Ke = (M + MS)/dt + theta*(A + D + AS), Fe = ((M + MS)/dt - (1 - theta)*(A + D + AS))*Te).
 * Tn (nnode): nodal field with the Dirichlet values already written on the fixed nodes; vel: nelem*2. */
void orc_advdiff_assemble(orc_system* S, int shape, int quad, int terms, const double* coords, int nelem, const int* conn,
                          const int* nodetoglobal, const double* vel, double k, double dt, double theta, const double* Tn) {
    const int npe = npe_of_shape(shape);
    memset(S->data, 0, sizeof(double) * (size_t)S->indptr[S->n]);
    memset(S->F, 0, sizeof(double) * (size_t)S->n);
    double KK[ORC_MAX_NPE * ORC_MAX_NPE], MM[ORC_MAX_NPE * ORC_MAX_NPE], Ke[ORC_MAX_NPE * ORC_MAX_NPE], Fe[ORC_MAX_NPE], xe[2 * ORC_MAX_NPE];
    for (int e = 0; e < nelem; e++) {
        const int* el = conn + (size_t)e * npe;
        for (int a = 0; a < npe; a++) for (int d = 0; d < 2; d++) xe[a * 2 + d] = coords[(size_t)el[a] * 2 + d];
        adv_sum(shape, quad, terms, ADV_A | ADV_D | ADV_S | ADV_SC, xe, vel[2 * e], vel[2 * e + 1], k, KK);
        if (dt == 0.0) {
            memcpy(Ke, KK, sizeof(double) * npe * npe);
            for (int i = 0; i < npe; i++) Fe[i] = 0.0;
        } else {
            adv_sum(shape, quad, terms, ADV_M | ADV_MS, xe, vel[2 * e], vel[2 * e + 1], k, MM);
            for (int i = 0; i < npe * npe; i++) Ke[i] = MM[i] / dt + KK[i] * theta;
            for (int i = 0; i < npe; i++) {
                double v = 0.0;
                for (int j = 0; j < npe; j++) v += (MM[i * npe + j] / dt - KK[i * npe + j] * (1.0 - theta)) * Tn[el[j]];
                Fe[i] = v;
            }
        }
        for (int i = 0; i < npe; i++) {
            int r = nodetoglobal[el[i]];
            if (r == -1) continue;
            for (int j = 0; j < npe; j++) {
                int c = nodetoglobal[el[j]];
                if (c != -1) S->data[find_col(S, r, c)] += Ke[i * npe + j];         /* Assembling.h:30 / :55 */
                else S->F[r] -= Ke[i * npe + j] * Tn[el[j]];                        /* Assembling.h:34 / :59 */
            }
            if (dt != 0.0) S->F[r] += Fe[i];                                        /* Assembling.h:38 */
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Load vectors: PlaneStrainSurfaceForce / PlaneStrainBodyForce (FEM/Equation/PlaneStrain.h:421-455, 503-537), PlaneStressSurfaceForce /
 * PlaneStressBodyForce (PlaneStress.h:98-167, the same arithmetic) and HeatTransferSurfaceFlux (HeatTransfer.h:76-98).
 *   per integration point:  N, dNdr ; x = X^T N ; dXdr = dNdr X ; m = sqrt((dXdr dXdr^T)(0,0)) on lines | det(dXdr) on areas
 *                           Fe += B^T f(x) * m * t * w0 [* w1]         evaluated left to right like the reference's operator chain
 * Line shapes ShapeFunction2Line / 3Line (ShapeFunction.h:20-84), rules Gauss1Line / Gauss2Line (GaussIntegration.h:18-60).
 * The force functor is replaced by its values fg[g][ndof] at the integration points (orc_integration_points gives the x_g).
 * ---------------------------------------------------------------------------------------------------------- */
enum { SHAPE_LINE2 = 8, SHAPE_LINE3 = 9, QUAD_G1LINE = 9, QUAD_G2LINE = 10 };
static int load_npe(int shape) { return shape == SHAPE_LINE2 ? 2 : shape == SHAPE_LINE3 ? 3 : npe_of_shape(shape); }
static int load_ngauss(int quad) { return quad == QUAD_G1LINE ? 1 : quad == QUAD_G2LINE ? 2 : quad_count(quad); }
static void load_point(int shape, int quad, int g, double* N, double* d /* [2][npe] */, double* w /* [2] */) {
    const int npe = load_npe(shape);
    if (shape == SHAPE_LINE2 || shape == SHAPE_LINE3) {
        double x = 0.0;
        w[0] = 2.0; w[1] = 1.0;
        if (quad == QUAD_G2LINE) { x = (g == 0 ? -1.0 : 1.0) / sqrt(3.0); w[0] = 1.0; }
        if (shape == SHAPE_LINE2) { N[0] = 0.5 * (1 - x); N[1] = 0.5 * (1 + x); d[0] = -0.5; d[1] = 0.5; }
        else {
            N[0] = -0.5 * (1.0 - x) * x; N[1] = 0.5 * x * (1.0 + x); N[2] = (1.0 - x) * (1.0 + x);
            d[0] = -0.5 * (1.0 - 2.0 * x); d[1] = 0.5 * (1.0 + 2.0 * x); d[2] = -2.0 * x;
        }
        for (int n = 0; n < npe; n++) d[npe + n] = 0.0;
        return;
    }
    double r[3], w3[3];
    quad_point(quad, g, r, w3);
    w[0] = w3[0]; w[1] = w3[1];
    shape_n2d(shape, r, N);
    shape_dndr(shape, r, d);
}
int orc_load_ngauss(int quad) { return load_ngauss(quad); }
/* xg[g][2] = X^T N(r_g) for one element */
void orc_integration_points(int shape, int quad, const double* xe, double* xg) {
    const int npe = load_npe(shape), ng = load_ngauss(quad);
    for (int g = 0; g < ng; g++) {
        double N[8], d[16], w[2];
        load_point(shape, quad, g, N, d, w);
        double x0 = 0.0, x1 = 0.0;
        for (int n = 0; n < npe; n++) { x0 += xe[2 * n] * N[n]; x1 += xe[2 * n + 1] * N[n]; }     /* X.Transpose()*N: Matrix*Vector sums over the nodes in order */
        xg[2 * g] = x0; xg[2 * g + 1] = x1;
    }
}
/* Fe[npe*ndof] of one element */
void orc_load_vector(int shape, int quad, int ndof, const double* xe, const double* fg, double t, double* Fe) {
    const int npe = load_npe(shape), ng = load_ngauss(quad);
    const int line = (shape == SHAPE_LINE2 || shape == SHAPE_LINE3);
    for (int k = 0; k < npe * ndof; k++) Fe[k] = 0.0;
    for (int g = 0; g < ng; g++) {
        double N[8], d[16], w[2];
        load_point(shape, quad, g, N, d, w);
        double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;
        for (int n = 0; n < npe; n++) { j00 += d[n] * xe[2 * n]; j01 += d[n] * xe[2 * n + 1]; j10 += d[npe + n] * xe[2 * n]; j11 += d[npe + n] * xe[2 * n + 1]; }
        const double m = line ? sqrt(j00 * j00 + j01 * j01) : (j00 * j11 - j01 * j10);
        for (int n = 0; n < npe; n++)
            for (int i = 0; i < ndof; i++) {
                double v = N[n] * fg[g * ndof + i];           /* (B^T f)(ndof*n + i) */
                v = v * m; v = v * t; v = v * w[0];
                if (!line) v = v * w[1];
                Fe[n * ndof + i] += v;
            }
    }
}


/* ------------------------------------------------------------------------------------------------------------
 * CSR<T>::operator*   src/LinearAlgebra/Models/CSR.h:109-122 (the reference's only OpenMP loop)
 * ---------------------------------------------------------------------------------------------------------- */
void orc_spmv(const orc_system* S, const double* x, double* y) {
    const int n = S->n;
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        double v = 0.0;
        for (int j = S->indptr[i], je = S->indptr[i + 1]; j < je; ++j) v += S->data[j] * x[S->indices[j]];
        y[i] = v;
    }
}
static double dot(int n, const double* a, const double* b) {   /* std::inner_product: serial left-to-right */
    double s = 0.0;
    for (int i = 0; i < n; i++) s = s + a[i] * b[i];
    return s;
}
/* CSR::get(i,i) (CSR.h:155-167) -> 0 if the diagonal is structurally absent; GetDiagonal CG.h:398-404 */
static double diag_of(const orc_system* S, int i) { int p = find_col(S, i, i); return p < 0 ? 0.0 : S->data[p]; }

/* ------------------------------------------------------------------------------------------------------------
 * ILU0     src/LinearAlgebra/Solvers/CG.h:258-284   (row-wise; unit-L strictly lower + U with diagonal in A's pattern)
 * PreILU0  src/LinearAlgebra/Solvers/CG.h:289-315
 * The reference scans k over ALL rows with two binary searches per k; only k in row i's pattern with j in row k's
 * pattern contribute, in ascending k, until k reaches min(i,j) -- so iterating row i's own column list is the same
 * sequence of subtractions.
 * ---------------------------------------------------------------------------------------------------------- */
orc_system* orc_ilu0(const orc_system* A) {
    orc_system* M = orc_system_from_csr(A->n, A->indptr, A->indices, A->data);
    for (int i = 0; i < A->n; i++) {
        for (int n = A->indptr[i]; n < A->indptr[i + 1]; n++) {
            int j = A->indices[n];
            double qij = A->data[n];
            int lim = (i <= j) ? i : j;
            for (int q = A->indptr[i]; q < A->indptr[i + 1]; q++) {
                int k = A->indices[q];
                if (k >= lim) break;
                int pkj = find_col(A, k, j);
                if (pkj >= 0) qij -= M->data[q] * M->data[pkj];
            }
            if (i > j) qij /= M->data[find_col(A, j, j)];
            M->data[n] = qij;
        }
    }
    return M;
}
void orc_preilu0(const orc_system* M, const double* b, double* v) {
    const int n = M->n;
    if (v != b) memcpy(v, b, sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++)
        for (int k = M->indptr[i]; k < M->indptr[i + 1]; k++) {
            if (M->indices[k] < i) v[i] -= M->data[k] * v[M->indices[k]]; else break;
        }
    for (int i = n - 1; i >= 0; i--) {
        for (int k = M->indptr[i + 1] - 1; k >= M->indptr[i]; k--) {
            if (M->indices[k] > i) v[i] -= M->data[k] * v[M->indices[k]]; else break;
        }
        v[i] /= diag_of(M, i);
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * CG         src/LinearAlgebra/Solvers/CG.h:124-154
 * ScalingCG  src/LinearAlgebra/Solvers/CG.h:420-453  (Jacobi: GetDiagonal :398, Scaling :409)
 * ILU0CG     src/LinearAlgebra/Solvers/CG.h:320-352
 * kind: 0 CG, 1 ScalingCG, 2 ILU0CG (M required).  x0 = 0; stop when ||r||_2 < eps*||b||_2 on the recursively
 * updated residual; returns the number of iterations performed (k+1 at convergence, itrmax on failure) and the final
 * ||r||/||b|| in *relres.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_solve(const orc_system* A, const orc_system* M, int kind, const double* b, int itrmax, double eps, double* x, double* relres) {
    const int n = A->n;
    double* r = (double*)malloc(sizeof(double) * n), *p = (double*)malloc(sizeof(double) * n);
    double* z = (double*)malloc(sizeof(double) * n), *Ap = (double*)malloc(sizeof(double) * n);
    double* D = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) { x[i] = 0.0; D[i] = (kind == 1) ? diag_of(A, i) : 1.0; }
    orc_spmv(A, x, Ap);
    for (int i = 0; i < n; i++) r[i] = b[i] - Ap[i];
    if (kind == 0) memcpy(z, r, sizeof(double) * n);
    else if (kind == 1) for (int i = 0; i < n; i++) z[i] = r[i] / D[i];
    else orc_preilu0(M, r, z);
    memcpy(p, z, sizeof(double) * n);
    double bnorm = sqrt(dot(n, b, b));
    double rho = dot(n, z, r);
    int it = itrmax;
    double rnorm = sqrt(dot(n, r, r));
    for (int k = 0; k < itrmax; ++k) {
        orc_spmv(A, p, Ap);
        double alpha = rho / dot(n, p, Ap);
        for (int i = 0; i < n; i++) x[i] = x[i] + alpha * p[i];
        for (int i = 0; i < n; i++) r[i] = r[i] + (-alpha) * Ap[i];
        double rho1;
        if (kind == 0) { rho1 = dot(n, r, r); }
        else {
            if (kind == 1) for (int i = 0; i < n; i++) z[i] = r[i] / D[i]; else orc_preilu0(M, r, z);
            rho1 = dot(n, z, r);
        }
        double beta = rho1 / rho;
        const double* zz = (kind == 0) ? r : z;
        for (int i = 0; i < n; i++) p[i] = beta * p[i] + zz[i];
        rho = rho1;
        rnorm = (kind == 0) ? sqrt(rho) : sqrt(dot(n, r, r));
        if (rnorm < eps * bnorm) { it = k + 1; break; }
    }
    if (relres) *relres = rnorm / bnorm;
    free(r); free(p); free(z); free(Ap); free(D);
    return it;
}

/* ------------------------------------------------------------------------------------------------------------
 * The non-symmetric Krylov family
 * BiCGSTAB         src/LinearAlgebra/Solvers/CG.h:159-194      kind 3
 * BiCGSTAB2        src/LinearAlgebra/Solvers/CG.h:199-253      kind 4
 * ScalingBiCGSTAB  src/LinearAlgebra/Solvers/CG.h:458-495      kind 5   (note p0 = D^-1 r0, :465)
 * ILU0BiCGSTAB     src/LinearAlgebra/Solvers/CG.h:357-393      kind 6   (M = ILU0 factors)
 * Vector updates keep the operand order of xeaxpbypcz / zeaxpby / zeawpbxmypcz / zeawpbxpcy / zeavpbwpcxpdy (CG.h:56-120).
 * Returns the number of iterations performed; *relres = final ||r||/||b|| of the recursive residual.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_solve_bicgstab(const orc_system* A, const orc_system* M, int kind, const double* b, int itrmax, double eps, double* x, double* relres) {
    const int n = A->n;
    const size_t bytes = sizeof(double) * (size_t)n;
    double *r = (double*)malloc(bytes), *rdash = (double*)malloc(bytes), *p = (double*)malloc(bytes), *Ap = (double*)malloc(bytes);
    double *s = (double*)malloc(bytes), *As = (double*)malloc(bytes), *Mp = (double*)malloc(bytes), *Ms = (double*)malloc(bytes);
    double *D = (double*)malloc(bytes), *u = (double*)calloc((size_t)n, sizeof(double)), *w = (double*)calloc((size_t)n, sizeof(double));
    double *z = (double*)calloc((size_t)n, sizeof(double)), *tkm1 = (double*)calloc((size_t)n, sizeof(double)), *y = (double*)malloc(bytes);
    for (int i = 0; i < n; i++) { x[i] = 0.0; D[i] = (kind == 5) ? diag_of(A, i) : 1.0; }
    orc_spmv(A, x, Ap);
    for (int i = 0; i < n; i++) { r[i] = b[i] - Ap[i]; rdash[i] = r[i]; }
    if (kind == 4) memset(p, 0, bytes);
    else if (kind == 5) for (int i = 0; i < n; i++) p[i] = r[i] / D[i];
    else memcpy(p, r, bytes);
    double beta = 0.0;
    double rdashr = dot(n, rdash, r);
    const double bnorm = sqrt(dot(n, b, b));
    int it = itrmax;
    double rnorm = sqrt(dot(n, r, r));
    for (int k = 0; k < itrmax; ++k) {
        if (kind != 4) {
            const double* mp = p;
            if (kind == 5) { for (int i = 0; i < n; i++) Mp[i] = p[i] / D[i]; mp = Mp; }
            else if (kind == 6) { orc_preilu0(M, p, Mp); mp = Mp; }
            orc_spmv(A, mp, Ap);
            const double alpha = rdashr / dot(n, rdash, Ap);
            for (int i = 0; i < n; i++) s[i] = 1.0 * r[i] + (-alpha) * Ap[i];
            const double* ms = s;
            if (kind == 5) { for (int i = 0; i < n; i++) Ms[i] = s[i] / D[i]; ms = Ms; }
            else if (kind == 6) { orc_preilu0(M, s, Ms); ms = Ms; }
            orc_spmv(A, ms, As);
            const double omega = dot(n, As, s) / dot(n, As, As);
            for (int i = 0; i < n; i++) x[i] = 1.0 * x[i] + alpha * mp[i] + omega * ms[i];
            for (int i = 0; i < n; i++) r[i] = 1.0 * s[i] + (-omega) * As[i];
            const double rdashr1 = dot(n, rdash, r);
            beta = alpha / omega * rdashr1 / rdashr;
            for (int i = 0; i < n; i++) p[i] = beta * p[i] + 1.0 * r[i] + (-beta * omega) * Ap[i];
            rdashr = rdashr1;
        } else {
            double* t = s; double* At = As;
            for (int i = 0; i < n; i++) p[i] = beta * p[i] + 1.0 * r[i] + (-beta) * u[i];
            orc_spmv(A, p, Ap);
            const double alpha = rdashr / dot(n, rdash, Ap);
            for (int i = 0; i < n; i++) y[i] = 1.0 * tkm1[i] + (-1.0) * r[i] + (-alpha) * w[i] + alpha * Ap[i];
            for (int i = 0; i < n; i++) t[i] = 1.0 * r[i] + (-alpha) * Ap[i];
            orc_spmv(A, t, At);
            const double Att = dot(n, At, t), AtAt = dot(n, At, At);
            double zeta, ita;
            if (k % 2 == 0) { zeta = Att / AtAt; ita = 0.0; }
            else {
                const double yy = dot(n, y, y), yt = dot(n, y, t), Aty = dot(n, At, y);
                zeta = (yy * Att - yt * Aty) / (AtAt * yy - Aty * Aty);
                ita = (AtAt * yt - Aty * Att) / (AtAt * yy - Aty * Aty);
            }
            for (int i = 0; i < n; i++) u[i] = zeta * Ap[i] + ita * (tkm1[i] - r[i] + beta * u[i]);
            for (int i = 0; i < n; i++) z[i] = ita * z[i] + zeta * r[i] + (-alpha) * u[i];
            for (int i = 0; i < n; i++) x[i] = 1.0 * x[i] + alpha * p[i] + 1.0 * z[i];
            for (int i = 0; i < n; i++) r[i] = 1.0 * t[i] + (-ita) * y[i] + (-zeta) * At[i];
            const double rdashr1 = dot(n, rdash, r);
            beta = alpha * rdashr1 / (zeta * rdashr);
            for (int i = 0; i < n; i++) w[i] = 1.0 * At[i] + beta * Ap[i];
            rdashr = rdashr1;
            memcpy(tkm1, t, bytes);
        }
        rnorm = sqrt(dot(n, r, r));
        if (rnorm < eps * bnorm) { it = k + 1; break; }
    }
    if (relres) *relres = rnorm / bnorm;
    free(r); free(rdash); free(p); free(Ap); free(s); free(As); free(Mp); free(Ms); free(D); free(u); free(w); free(z); free(tkm1); free(y);
    return it;
}

/* ------------------------------------------------------------------------------------------------------------
 * Filters.
 * HeavisideFilter::GetFilteredVariables     src/Optimize/Filter/HeavisideFilter.h:61-73
 * HeavisideFilter::GetFilteredSensitivitis  src/Optimize/Filter/HeavisideFilter.h:77-99
 * DensityFilter::GetFilteredVariables       src/Optimize/Filter/DensityFilter.h:45-56
 * DensityFilter::GetFilteredSensitivitis    src/Optimize/Filter/DensityFilter.h:60-71
 * ---------------------------------------------------------------------------------------------------------- */
void orc_filter_apply(int kind, int n, const long long* rowptr, const int* nbr, const double* w, double beta,
                      const double* s, double* rho) {
    for (int i = 0; i < n; i++) {
        double wssum = 0.0, wsum = 0.0;
        for (long long j = rowptr[i]; j < rowptr[i + 1]; j++) { wssum += w[j] * s[nbr[j]]; wsum += w[j]; }
        if (kind == FILTER_HEAVISIDE) rho[i] = 0.5 * (tanh(0.5 * beta) + tanh(beta * (wssum / wsum - 0.5))) / tanh(0.5 * beta);
        else rho[i] = wssum / wsum;
    }
}
void orc_filter_sens(int kind, int n, const long long* rowptr, const int* nbr, const double* w, double beta,
                     const double* s, const double* dfdrho, double* dfds) {
    double* dr = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) {
        if (kind == FILTER_HEAVISIDE) {
            double wssum = 0.0, wsum = 0.0;
            for (long long j = rowptr[i]; j < rowptr[i + 1]; j++) { wssum += w[j] * s[nbr[j]]; wsum += w[j]; }
            dr[i] = 0.5 * beta * (1.0 - pow(tanh(beta * (wssum / wsum - 0.5)), 2.0)) / tanh(0.5 * beta);
        } else dr[i] = 1.0;
    }
    for (int i = 0; i < n; i++) {
        double acc = 0.0, wsum = 0.0;
        for (long long j = rowptr[i]; j < rowptr[i + 1]; j++) {
            if (kind == FILTER_HEAVISIDE) acc += dfdrho[nbr[j]] * dr[nbr[j]] * w[j];
            else acc += dfdrho[nbr[j]] * w[j];
            wsum += w[j];
        }
        dfds[i] = acc / wsum;       /* normalised by the RECEIVING row's weight sum, as the reference does */
    }
    free(dr);
}

/* SensitivityFilter (Sigmund) src/Optimize/Filter/SensitivityFilter.h:44-55 ; SensitivityFilter2 (Borrvall) :88-99 */
void orc_sensitivity_filter(int kind, int n, const long long* rowptr, const int* nbr, const double* w, const double* s,
                            const double* dfds_in, double* out) {
    for (int i = 0; i < n; i++) {
        double acc = 0.0, wsum = 0.0;
        for (long long j = rowptr[i]; j < rowptr[i + 1]; j++) {
            acc += w[j] * s[nbr[j]] * dfds_in[nbr[j]];
            wsum += (kind == 2) ? w[j] : w[j] * s[nbr[j]];
        }
        out[i] = (kind == 2) ? acc / (wsum * s[i]) : acc / wsum;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * OC::UpdateVariables  src/Optimize/Solver/OC.h:78-107, constraint functor of sample_optimize_density_oc.cpp:198-207
 * (filter + volume).  xk updated in place; returns the number of bisection steps; *lambda_out = last lambda tried.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_oc_update(int n, double iota, double lmin, double lmax, double leps, double move,
                  int fkind, const long long* rowptr, const int* nbr, const double* w, double beta,
                  double weightlimit, double scale1, double* xk, const double* dfdx, const double* dgdx, double* lambda_out) {
    double lambda0 = lmin, lambda1 = lmax, lambda = 0.0;
    double* xkp1 = (double*)calloc(n, sizeof(double)), *rho = (double*)malloc(sizeof(double) * n);
    int steps = 0;
    while ((lambda1 - lambda0) / (lambda1 + lambda0) > leps) {
        lambda = 0.5 * (lambda1 + lambda0);
        for (int i = 0; i < n; i++) {
            double v = pow(-dfdx[i] / (dgdx[i] * lambda), iota) * xk[i];
            double lo = fmax(0.0, (1.0 - move) * xk[i]), hi = fmin(1.0, (1.0 + move) * xk[i]);
            if (v < lo) v = lo; else if (v > hi) v = hi;
            xkp1[i] = v;
        }
        orc_filter_apply(fkind, n, rowptr, nbr, w, beta, xkp1, rho);
        double g = 0.0;
        for (int i = 0; i < n; i++) g += scale1 * rho[i] / (weightlimit * n);
        g = g - 1.0 * scale1;
        if (g > 0.0) lambda0 = lambda; else lambda1 = lambda;
        steps++;
    }
    memcpy(xk, xkp1, sizeof(double) * n);
    if (lambda_out) *lambda_out = lambda;
    free(xkp1); free(rho);
    return steps;
}
/* OC::IsConvergence OC.h:68-73 == MMA::IsConvergence MMA.h:108-113 */
int orc_is_convergence(double f, double fprev, double eps) { return fabs(f - fprev) / (f + fprev) < eps; }

/* ------------------------------------------------------------------------------------------------------------
 * MMA  src/Optimize/Solver/MMA.h:117-419 (UpdateVariables), :423-461 (KKTNorm), :465-509 (solvels)
 * State (xkm1, xkm2, L, U, k) lives in orc_mma.  General m; the n > m branch (:260-291) and the n <= m branch
 * (:292-330) are both restated.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    int n, m, k;
    double a0, *a, *c, *d, *xmin, *xmax, *xkm1, *xkm2, *L, *U;
    double raa0, albefa, move, asyinit, asydecr, asyincr;
    int newton_steps, halvings;
    int conlin;     /* 1: CONLIN<T> (src/Optimize/Solver/CONLIN.h:89-373): p*x + q/x instead of p/(U-x) + q/(x-L), no asymptotes */
} orc_mma;

orc_mma* orc_mma_create(int n, int m, double a0, const double* a, const double* c, const double* d, const double* xmin, const double* xmax) {
    orc_mma* M = (orc_mma*)calloc(1, sizeof(orc_mma));
    M->n = n; M->m = m; M->a0 = a0;
#define DUP(dst, src, cnt) dst = (double*)malloc(sizeof(double) * (cnt)); memcpy(dst, src, sizeof(double) * (cnt))
    DUP(M->a, a, m); DUP(M->c, c, m); DUP(M->d, d, m); DUP(M->xmin, xmin, n); DUP(M->xmax, xmax, n);
#undef DUP
    M->xkm1 = (double*)calloc(n, sizeof(double)); M->xkm2 = (double*)calloc(n, sizeof(double));
    M->L = (double*)calloc(n, sizeof(double)); M->U = (double*)calloc(n, sizeof(double));
    M->raa0 = 1.0e-5; M->albefa = 0.1; M->move = 0.5; M->asyinit = 0.5; M->asydecr = 0.7; M->asyincr = 1.2;   /* MMA.h:80-85 */
    return M;
}
void orc_mma_free(orc_mma* M) {
    if (!M) return;
    free(M->a); free(M->c); free(M->d); free(M->xmin); free(M->xmax); free(M->xkm1); free(M->xkm2); free(M->L); free(M->U); free(M);
}
void orc_mma_setparameters(orc_mma* M, double raa0, double albefa, double move, double asyinit, double asydecr, double asyincr) {
    M->raa0 = raa0; M->albefa = albefa; M->move = move; M->asyinit = asyinit; M->asydecr = asydecr; M->asyincr = asyincr;
}
/* CONLIN ctor default move = 0.5 (CONLIN.h:66), SetParameters(move, epsvalue) (:71-74) */
void orc_mma_set_conlin(orc_mma* M, double move) { M->conlin = 1; M->move = move; }
void orc_mma_stats(const orc_mma* M, int* newton, int* halvings) { *newton = M->newton_steps; *halvings = M->halvings; }

static void solvels(int N, double* A, double* b, double* x) {    /* MMA.h:465-509, A row-major N x N, destroyed */
    for (int i = 0; i < N - 1; i++) {
        double pivot = fabs(A[i * N + i]);
        int pi = i;
        for (int j = i + 1; j < N; j++) if (pivot < fabs(A[j * N + i])) { pivot = fabs(A[j * N + i]); pi = j; }
        if (pi != i) {
            double tmp = b[i]; b[i] = b[pi]; b[pi] = tmp;
            for (int j = i; j < N; j++) { tmp = A[i * N + j]; A[i * N + j] = A[pi * N + j]; A[pi * N + j] = tmp; }
        }
        for (int j = i + 1; j < N; j++) {
            for (int k = i + 1; k < N; k++) A[j * N + k] -= A[i * N + k] * A[j * N + i] / A[i * N + i];
            b[j] -= b[i] * A[j * N + i] / A[i * N + i];
        }
    }
    for (int i = N - 1; i >= 0; i--) {
        x[i] = b[i];
        for (int j = N - 1; j > i; j--) x[i] -= x[j] * A[i * N + j];
        x[i] /= A[i * N + i];
    }
}

/* the separable convex approximation and its derivatives: MMA (MMA.h:211-233) | CONLIN (CONLIN.h:160-172) */
#define PHI(pv, qv, xv, j)  (M->conlin ? (pv) * (xv) + (qv) / (xv) : (pv) / (M->U[j] - (xv)) + (qv) / ((xv) - M->L[j]))
#define DPHI(pv, qv, xv, j) (M->conlin ? (pv) - (qv) / pow((xv), 2.0) : (pv) / pow(M->U[j] - (xv), 2.0) - (qv) / pow((xv) - M->L[j], 2.0))
#define D2PHI(pv, qv, xv, j) (M->conlin ? 2.0 * (qv) / pow((xv), 3.0) : 2.0 * (pv) / pow(M->U[j] - (xv), 3.0) + 2.0 * (qv) / pow((xv) - M->L[j], 3.0))

static double kktnorm(const orc_mma* M, const double* x, const double* y, double z, const double* lam, const double* gsi,
                      const double* ita, const double* mu, double zeta, const double* s, double eps,
                      const double* p, const double* q, const double* p0, const double* q0, const double* alpha,
                      const double* beta, const double* b) {
    const int n = M->n, m = M->m;
    double norm = 0.0;
    double* g = (double*)calloc(m, sizeof(double));
    double* pl = (double*)malloc(sizeof(double) * n), *ql = (double*)malloc(sizeof(double) * n);
    for (int j = 0; j < n; j++) {
        pl[j] = p0[j]; ql[j] = q0[j];
        for (int i = 0; i < m; i++) {
            pl[j] += lam[i] * p[(size_t)i * n + j];
            ql[j] += lam[i] * q[(size_t)i * n + j];
            g[i] += PHI(p[(size_t)i * n + j], q[(size_t)i * n + j], x[j], j);
        }
    }
    for (int j = 0; j < n; j++) {
        norm += pow(DPHI(pl[j], ql[j], x[j], j) - gsi[j] + ita[j], 2.0);
        norm += pow(gsi[j] * (x[j] - alpha[j]) - eps, 2.0);
        norm += pow(ita[j] * (beta[j] - x[j]) - eps, 2.0);
    }
    double la = 0.0;
    for (int i = 0; i < m; i++) {
        norm += pow(M->c[i] + M->d[i] * y[i] - lam[i] - mu[i], 2.0);
        norm += pow(g[i] - M->a[i] * z - y[i] + s[i] - b[i], 2.0);
        norm += pow(mu[i] * y[i] - eps, 2.0);
        norm += pow(lam[i] * s[i] - eps, 2.0);
        la = la + lam[i] * M->a[i];
    }
    norm += pow(M->a0 - zeta - la, 2.0);
    norm += pow(zeta * z - eps, 2.0);
    free(g); free(pl); free(ql);
    return sqrt(norm);
}

static double max2(double a, double b) { return a < b ? b : a; }     /* std::max semantics */
static double min2(double a, double b) { return b < a ? b : a; }

void orc_mma_update(orc_mma* M, double* xk, const double* dfdx, const double* gval, const double* dgdx /* m x n */) {
    const int n = M->n, m = M->m;
    double *L = M->L, *U = M->U;
    /* asymptotes MMA.h:119-142 (CONLIN has none) */
    if (M->conlin) {
        /* nothing */
    } else if (M->k < 2) {
        for (int j = 0; j < n; j++) {
            double w = M->xmax[j] - M->xmin[j];
            L[j] = xk[j] - M->asyinit * w; U[j] = xk[j] + M->asyinit * w;
            L[j] = min2(max2(xk[j] - 10.0 * w, L[j]), xk[j] - 0.01 * w);
            U[j] = min2(max2(xk[j] + 0.01 * w, U[j]), xk[j] + 10.0 * w);
        }
    } else {
        for (int j = 0; j < n; j++) {
            double w = M->xmax[j] - M->xmin[j];
            double tmp = (xk[j] - M->xkm1[j]) * (M->xkm1[j] - M->xkm2[j]);
            double fac = tmp < 0.0 ? M->asydecr : (tmp > 0.0 ? M->asyincr : 1.0);
            if (fac != 1.0) {
                L[j] = xk[j] - fac * (M->xkm1[j] - L[j]); U[j] = xk[j] + fac * (U[j] - M->xkm1[j]);
            } else {
                L[j] = xk[j] - (M->xkm1[j] - L[j]); U[j] = xk[j] + (U[j] - M->xkm1[j]);
            }
            L[j] = min2(max2(xk[j] - 10.0 * w, L[j]), xk[j] - 0.01 * w);
            U[j] = min2(max2(xk[j] + 0.01 * w, U[j]), xk[j] + 10.0 * w);
        }
    }
#define VEC(name, cnt) double* name = (double*)calloc((cnt) > 0 ? (size_t)(cnt) : 1, sizeof(double))
    VEC(alpha, n); VEC(beta, n); VEC(p0, n); VEC(q0, n); VEC(p, (size_t)m * n); VEC(q, (size_t)m * n); VEC(b, m);
    for (int j = 0; j < n; j++) {   /* MMA.h:145-160 */
        double w = M->xmax[j] - M->xmin[j];
        double dp = max2(dfdx[j], 0.0), dm = max2(-dfdx[j], 0.0);
        if (M->conlin) {        /* CONLIN.h:92-108 */
            alpha[j] = max2(M->xmin[j], xk[j] - M->move * w);
            beta[j] = min2(M->xmax[j], xk[j] + M->move * w);
            p0[j] = dp;
            q0[j] = dm * pow(xk[j], 2.0);
            continue;
        }
        alpha[j] = max2(max2(M->xmin[j], L[j] + M->albefa * (xk[j] - L[j])), xk[j] - M->move * w);
        beta[j] = min2(min2(M->xmax[j], U[j] - M->albefa * (U[j] - xk[j])), xk[j] + M->move * w);
        p0[j] = pow(U[j] - xk[j], 2.0) * (1.001 * dp + 0.001 * dm + M->raa0 / w);
        q0[j] = pow(xk[j] - L[j], 2.0) * (0.001 * dp + 1.001 * dm + M->raa0 / w);
    }
    for (int i = 0; i < m; i++) {   /* MMA.h:163-175 */
        b[i] = -gval[i];
        for (int j = 0; j < n; j++) {
            double w = M->xmax[j] - M->xmin[j];
            double dp = max2(dgdx[(size_t)i * n + j], 0.0), dm = max2(-dgdx[(size_t)i * n + j], 0.0);
            if (M->conlin) {    /* CONLIN.h:114-122 */
                p[(size_t)i * n + j] = dp;
                q[(size_t)i * n + j] = dm * pow(xk[j], 2.0);
                b[i] += p[(size_t)i * n + j] * xk[j] + q[(size_t)i * n + j] / xk[j];
                continue;
            }
            p[(size_t)i * n + j] = pow(U[j] - xk[j], 2.0) * (1.001 * dp + 0.001 * dm + M->raa0 / w);
            q[(size_t)i * n + j] = pow(xk[j] - L[j], 2.0) * (0.001 * dp + 1.001 * dm + M->raa0 / w);
            b[i] += p[(size_t)i * n + j] / (U[j] - xk[j]) + q[(size_t)i * n + j] / (xk[j] - L[j]);
        }
    }
    /* initial point MMA.h:178-197 */
    double eps = 1.0, z = 1.0, zeta = 1.0;
    VEC(x, n); VEC(y, m); VEC(lam, m); VEC(s, m); VEC(gsi, n); VEC(ita, n); VEC(mu, m);
    for (int i = 0; i < m; i++) { y[i] = 1.0; lam[i] = 1.0; s[i] = 1.0; mu[i] = max2(1.0, 0.5 * M->c[i]); }
    for (int j = 0; j < n; j++) {
        x[j] = 0.5 * (alpha[j] + beta[j]);
        gsi[j] = max2(1.0, 1.0 / (x[j] - alpha[j]));
        ita[j] = max2(1.0, 1.0 / (beta[j] - x[j]));
    }
    VEC(pl, n); VEC(ql, n); VEC(G, (size_t)m * n); VEC(Dx, n); VEC(dtx, n);
    VEC(Dy, m); VEC(Dlam, m); VEC(dty, m); VEC(dtlam, m); VEC(Dlamy, m); VEC(dtlamy, m);
    VEC(dx, n); VEC(dy, m); VEC(dlam, m); VEC(dgsi, n); VEC(dita, n); VEC(dmu, m); VEC(ds, m);
    VEC(xn, n); VEC(yn, m); VEC(lamn, m); VEC(gsin, n); VEC(itan, n); VEC(mun, m); VEC(sn, m);
    const int NS = (n > m ? m : n) + 1;
    VEC(A, (size_t)NS * NS); VEC(B, NS); VEC(sol, NS);
    M->newton_steps = 0; M->halvings = 0;
    for (int l = 0; eps > 1.0e-7; l++) {    /* MMA.h:199 */
        for (int j = 0; j < n; j++) {
            pl[j] = p0[j]; ql[j] = q0[j];
            for (int i = 0; i < m; i++) { pl[j] += lam[i] * p[(size_t)i * n + j]; ql[j] += lam[i] * q[(size_t)i * n + j]; }
        }
        for (int i = 0; i < m; i++) for (int j = 0; j < n; j++)
            G[(size_t)i * n + j] = DPHI(p[(size_t)i * n + j], q[(size_t)i * n + j], x[j], j);
        for (int j = 0; j < n; j++) {
            Dx[j] = D2PHI(pl[j], ql[j], x[j], j) + gsi[j] / (x[j] - alpha[j]) + ita[j] / (beta[j] - x[j]);
            dtx[j] = DPHI(pl[j], ql[j], x[j], j) - eps / (x[j] - alpha[j]) + eps / (beta[j] - x[j]);
        }
        double la = 0.0;
        for (int i = 0; i < m; i++) {
            Dy[i] = M->d[i] + mu[i] / y[i];
            Dlam[i] = s[i] / lam[i];
            dty[i] = M->c[i] + M->d[i] * y[i] - lam[i] - eps / y[i];
            la = la + lam[i] * M->a[i];
        }
        double dtz = M->a0 - eps / z - la;
        for (int i = 0; i < m; i++) {
            dtlam[i] = -M->a[i] * z - y[i] - b[i] + eps / lam[i];
            for (int j = 0; j < n; j++) dtlam[i] += PHI(p[(size_t)i * n + j], q[(size_t)i * n + j], x[j], j);
            Dlamy[i] = Dlam[i] + 1.0 / Dy[i];
            dtlamy[i] = dtlam[i] + dty[i] / Dy[i];
        }
        double dz;
        memset(A, 0, sizeof(double) * (size_t)NS * NS);
        if (n > m) {   /* MMA.h:260-291 */
            const int N = m + 1;
            for (int ii = 0; ii < m; ii++) {
                for (int jj = 0; jj < m; jj++) for (int kk = 0; kk < n; kk++) A[ii * N + jj] += G[(size_t)ii * n + kk] * G[(size_t)jj * n + kk] / Dx[kk];
                A[ii * N + ii] += Dlamy[ii];
                A[ii * N + m] = M->a[ii];
                A[m * N + ii] = M->a[ii];
            }
            A[m * N + m] = -zeta / z;
            for (int ii = 0; ii < m; ii++) {
                B[ii] = dtlamy[ii];
                for (int jj = 0; jj < n; jj++) B[ii] -= G[(size_t)ii * n + jj] * dtx[jj] / Dx[jj];
            }
            B[m] = dtz;
            solvels(N, A, B, sol);
            for (int i = 0; i < m; i++) dlam[i] = sol[i];
            dz = sol[m];
            for (int j = 0; j < n; j++) {
                dx[j] = -dtx[j] / Dx[j];
                for (int i = 0; i < m; i++) dx[j] -= G[(size_t)i * n + j] * dlam[i] / Dx[j];
            }
        } else {       /* MMA.h:292-330 */
            const int N = n + 1;
            for (int ii = 0; ii < n; ii++) {
                for (int jj = 0; jj < n; jj++) for (int kk = 0; kk < m; kk++) A[ii * N + jj] += G[(size_t)kk * n + ii] * G[(size_t)kk * n + jj] / Dlamy[kk];
                A[ii * N + ii] += Dx[ii];
                for (int jj = 0; jj < m; jj++) {
                    A[ii * N + n] -= G[(size_t)jj * n + ii] * M->a[jj] / Dlamy[jj];
                    A[n * N + ii] -= G[(size_t)jj * n + ii] * M->a[jj] / Dlamy[jj];
                    A[n * N + n] += M->a[jj] * M->a[jj] / Dlamy[jj];
                }
            }
            A[n * N + n] += zeta / z;
            for (int ii = 0; ii < n; ii++) {
                B[ii] = -dtx[ii];
                for (int jj = 0; jj < m; jj++) B[ii] -= G[(size_t)jj * n + ii] * dtlamy[jj] / Dlamy[jj];
            }
            B[n] = -dtz;
            for (int jj = 0; jj < m; jj++) B[n] += M->a[jj] * dtlamy[jj] / Dlamy[jj];
            solvels(N, A, B, sol);
            for (int j = 0; j < n; j++) dx[j] = sol[j];
            dz = sol[n];
            for (int i = 0; i < m; i++) {
                dlam[i] = -M->a[i] * dz / Dlamy[i] + dtlamy[i] / Dlamy[i];
                for (int j = 0; j < n; j++) dlam[i] += G[(size_t)i * n + j] * dx[j] / Dlamy[i];
            }
        }
        for (int i = 0; i < m; i++) {   /* MMA.h:332-336 */
            dy[i] = dlam[i] / Dy[i] - dty[i] / Dy[i];
            dmu[i] = -mu[i] * dy[i] / y[i] - mu[i] + eps / y[i];
            ds[i] = -s[i] * dlam[i] / lam[i] - s[i] + eps / lam[i];
        }
        for (int j = 0; j < n; j++) {
            dgsi[j] = -gsi[j] * dx[j] / (x[j] - alpha[j]) - gsi[j] + eps / (x[j] - alpha[j]);
            dita[j] = ita[j] * dx[j] / (beta[j] - x[j]) - ita[j] + eps / (beta[j] - x[j]);
        }
        double dzeta = -zeta * dz / z - zeta + eps / z;
        double txmax = 0.0;             /* MMA.h:346-360 */
        for (int j = 0; j < n; j++) {
            double t = max2(max2(-1.01 * dx[j] / (x[j] - alpha[j]), 1.01 * dx[j] / (beta[j] - x[j])), max2(-1.01 * dgsi[j] / gsi[j], -1.01 * dita[j] / ita[j]));
            if (txmax < t) txmax = t;
        }
        double tymax = 0.0;
        for (int i = 0; i < m; i++) {
            double t = max2(max2(-1.01 * dy[i] / y[i], -1.01 * dlam[i] / lam[i]), max2(-1.01 * dmu[i] / mu[i], -1.01 * ds[i] / s[i]));
            if (tymax < t) tymax = t;
        }
        double tau = 1.0 / max2(max2(max2(1.0, txmax), max2(tymax, -1.01 * dz / z)), -1.01 * dzeta / zeta);
        double dwl = kktnorm(M, x, y, z, lam, gsi, ita, mu, zeta, s, eps, p, q, p0, q0, alpha, beta, b);
        double zn = z, zetan = zeta, dwl1 = 0.0;
        for (int ll = 0; ll < 50; ll++) {   /* MMA.h:371-391 */
            for (int j = 0; j < n; j++) { xn[j] = x[j] + tau * dx[j]; gsin[j] = gsi[j] + tau * dgsi[j]; itan[j] = ita[j] + tau * dita[j]; }
            for (int i = 0; i < m; i++) { yn[i] = y[i] + tau * dy[i]; lamn[i] = lam[i] + tau * dlam[i]; mun[i] = mu[i] + tau * dmu[i]; sn[i] = s[i] + tau * ds[i]; }
            zn = z + tau * dz; zetan = zeta + tau * dzeta;
            dwl1 = kktnorm(M, xn, yn, zn, lamn, gsin, itan, mun, zetan, sn, eps, p, q, p0, q0, alpha, beta, b);
            if (dwl1 < dwl) break;
            tau *= 0.5;
            M->halvings++;
        }
        memcpy(x, xn, sizeof(double) * n); memcpy(gsi, gsin, sizeof(double) * n); memcpy(ita, itan, sizeof(double) * n);
        memcpy(y, yn, sizeof(double) * m); memcpy(lam, lamn, sizeof(double) * m); memcpy(mu, mun, sizeof(double) * m); memcpy(s, sn, sizeof(double) * m);
        z = zn; zeta = zetan;
        M->newton_steps++;
        /* MMA.h:405-410: KKTNorm(accepted point) is the value just computed for the accepted trial */
        if (dwl1 < 0.9 * eps) eps *= 0.1;
    }
    M->k++;                                     /* MMA.h:414-418 */
    memcpy(M->xkm2, M->xkm1, sizeof(double) * n);
    memcpy(M->xkm1, xk, sizeof(double) * n);
    memcpy(xk, x, sizeof(double) * n);
    free(alpha); free(beta); free(p0); free(q0); free(p); free(q); free(b);
    free(x); free(y); free(lam); free(s); free(gsi); free(ita); free(mu);
    free(pl); free(ql); free(G); free(Dx); free(dtx); free(Dy); free(Dlam); free(dty); free(dtlam); free(Dlamy); free(dtlamy);
    free(dx); free(dy); free(dlam); free(dgsi); free(dita); free(dmu); free(ds);
    free(xn); free(yn); free(lamn); free(gsin); free(itan); free(mun); free(sn); free(A); free(B); free(sol);
#undef VEC
}

/* ------------------------------------------------------------------------------------------------------------
 * Reaction / compliance / sensitivity passes of the drivers:
 *   sample/optimize/sample_optimize_density_oc.cpp:136-162  (General.h:82-96 ElementVector, Assembling.h:119,163)
 * r = K_full(rho) u by element scatter in element order; f = scale0 * sum_nodes u_n . r_n (inner_product of
 * Vector<T>, i.e. per-node dot then serial sum); dfdrho_i = -scale0 p (E1-E0) rho_i^(p-1) ue^T Ke(E=1) ue.
 * ---------------------------------------------------------------------------------------------------------- */
double orc_compliance_sens(int eq, int nnode, const double* coords, int nelem, const int* conn, const double* u /* nnode*ndof */,
                           const double* rho, double E0, double E1, double V, double t, double p, double scale0,
                           double* r_out /* nnode*ndof */, double* dfdrho) {
    const int dim = dim_of(eq), npe = npe_of(eq), ndof = ndof_of(eq), m = npe * ndof;
    double Ke[ORC_MAX_M * ORC_MAX_M], xe[3 * ORC_MAX_NPE], ue[ORC_MAX_M], Keue[ORC_MAX_M];
    memset(r_out, 0, sizeof(double) * (size_t)nnode * ndof);
    for (int e = 0; e < nelem; e++) {
        const int* el = conn + (size_t)e * npe;
        for (int a = 0; a < npe; a++) for (int d = 0; d < dim; d++) xe[a * dim + d] = coords[(size_t)el[a] * dim + d];
        for (int a = 0; a < npe; a++) for (int d = 0; d < ndof; d++) ue[a * ndof + d] = u[(size_t)el[a] * ndof + d];
        double E = E1 * pow(rho[e], p) + E0 * (1.0 - pow(rho[e], p));
        orc_element_matrix(eq, xe, E, V, t, Ke);
        for (int i = 0; i < m; i++) { double v = 0.0; for (int j = 0; j < m; j++) v += Ke[i * m + j] * ue[j]; Keue[i] = v; }
        for (int a = 0; a < npe; a++) for (int d = 0; d < ndof; d++) r_out[(size_t)el[a] * ndof + d] += Keue[a * ndof + d];
        if (dfdrho) {
            orc_element_matrix(eq, xe, 1.0, V, t, Ke);
            for (int i = 0; i < m; i++) { double v = 0.0; for (int j = 0; j < m; j++) v += Ke[i * m + j] * ue[j]; Keue[i] = v; }
            double w = 0.0;
            for (int i = 0; i < m; i++) w += ue[i] * Keue[i];
            dfdrho[e] = -scale0 * p * (-E0 + E1) * pow(rho[e], p - 1.0) * w;
        }
    }
    double f = 0.0;
    for (int i = 0; i < nnode; i++) {
        double d = 0.0;
        for (int k = 0; k < ndof; k++) d += u[(size_t)i * ndof + k] * r_out[(size_t)i * ndof + k];
        f = f + d;
    }
    return scale0 * f;
}

/* ------------------------------------------------------------------------------------------------------------
 * The SIMP design loop  sample/optimize/sample_optimize_density_oc.cpp:83-208 / ..._mma.cpp:83-204
 * params / optp / outputs exactly as ref_simp_run in oracle/ref_shim.cpp; hist[5*k] = {f, g, seconds, converged, cg_iters}.
 * phase[8] = {filter, assembly, (unused), solve, reaction+sensitivity, (unused), filter-sens, update}.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_simp_run(int eq, int nnode, const double* coords, int nelem, const int* conn,
                 int nfixed, const int* fnode, const int* fdof, const double* fval,
                 int nload, const int* lnode, const int* ldof, const double* lval,
                 int fkind, const long long* rowptr, const int* nbr, const double* w,
                 int opt_kind, const double* optp, const double* params, int niter, int check_convergence,
                 double* s, double* rho, double* u_out, double* r_out, double* hist, double* phase) {
    const int npe = npe_of(eq), ndof = ndof_of(eq);
    const double E0 = params[0], E1 = params[1], V = params[2], p = params[3], weightlimit = params[4];
    const double scale0 = params[5], scale1 = params[6], thick = params[7];
    double beta = params[8];
    const int beta_period = (int)params[9], itrmax = (int)params[10];
    const double cgeps = params[11];
    const size_t nd = (size_t)nnode * ndof;
    int* n2g = (int*)malloc(sizeof(int) * nd);
    double* ufix = (double*)malloc(sizeof(double) * nd);
    int kdeg = orc_dofmap(nnode, ndof, nfixed, fnode, fdof, fval, n2g, ufix);
    orc_system* S = orc_pattern(nnode, ndof, npe, nelem, conn, n2g, kdeg);
    double* Emod = (double*)malloc(sizeof(double) * nelem), *sol = (double*)malloc(sizeof(double) * kdeg);
    double* dfdrho = (double*)malloc(sizeof(double) * nelem), *dgdrho = (double*)malloc(sizeof(double) * nelem);
    double* dfds = (double*)malloc(sizeof(double) * nelem), *dgds = (double*)malloc(sizeof(double) * nelem);
    double* u = (double*)malloc(sizeof(double) * nd);
    orc_mma* mma = NULL;
    double fprev = 0.0, epsvalue = 1.0e-5;   /* OC.h:49-50 */
    if (opt_kind == OPT_MMA) {
        double* xmin = (double*)malloc(sizeof(double) * nelem), *xmax = (double*)malloc(sizeof(double) * nelem);
        for (int i = 0; i < nelem; i++) { xmin[i] = optp[11]; xmax[i] = optp[12]; }
        mma = orc_mma_create(nelem, 1, optp[7], &optp[8], &optp[9], &optp[10], xmin, xmax);
        orc_mma_setparameters(mma, optp[0], optp[1], optp[2], optp[3], optp[4], optp[5]);
        epsvalue = optp[6];
        free(xmin); free(xmax);
    } else if (opt_kind == OPT_CONLIN) {      /* sample_optimize_density_CONLIN.cpp:80-85 */
        double* xmin = (double*)malloc(sizeof(double) * nelem), *xmax = (double*)malloc(sizeof(double) * nelem);
        for (int i = 0; i < nelem; i++) { xmin[i] = optp[6]; xmax[i] = optp[7]; }
        mma = orc_mma_create(nelem, 1, optp[2], &optp[3], &optp[4], &optp[5], xmin, xmax);
        orc_mma_set_conlin(mma, optp[0]);
        epsvalue = optp[1];
        free(xmin); free(xmax);
    }
    if (phase) memset(phase, 0, sizeof(double) * 8);
    int k = 0;
    for (; k < niter; k++) {
        double tstart = now_s(), t0 = tstart, t1;
        if (beta_period > 0 && k % beta_period == 0) beta *= 2.0;
        orc_filter_apply(fkind, nelem, rowptr, nbr, w, beta, s, rho);
        double g = 0.0;
        for (int i = 0; i < nelem; i++) { g += scale1 * rho[i] / (weightlimit * nelem); dgdrho[i] = scale1 / (weightlimit * nelem); }
        g -= 1.0 * scale1;
        t1 = now_s(); if (phase) phase[0] += t1 - t0; t0 = t1;
        for (int i = 0; i < nelem; i++) Emod[i] = E1 * pow(rho[i], p) + E0 * (1.0 - pow(rho[i], p));
        orc_assemble_numeric(S, eq, coords, nelem, conn, n2g, ufix, Emod, V, thick, nload, lnode, ldof, lval, NULL);
        t1 = now_s(); if (phase) phase[1] += t1 - t0; t0 = t1;
        double relres;
        int its = orc_solve(S, NULL, 1, S->F, itrmax, cgeps, sol, &relres);
        for (size_t i = 0; i < nd; i++) u[i] = (n2g[i] != -1) ? sol[n2g[i]] : ufix[i];   /* Disassembling Assembling.h:163 */
        t1 = now_s(); if (phase) phase[3] += t1 - t0; t0 = t1;
        double f = orc_compliance_sens(eq, nnode, coords, nelem, conn, u, rho, E0, E1, V, thick, p, scale0, r_out, dfdrho);
        t1 = now_s(); if (phase) phase[4] += t1 - t0; t0 = t1;
        orc_filter_sens(fkind, nelem, rowptr, nbr, w, beta, s, dfdrho, dfds);
        orc_filter_sens(fkind, nelem, rowptr, nbr, w, beta, s, dgdrho, dgds);
        t1 = now_s(); if (phase) phase[6] += t1 - t0; t0 = t1;
        hist[5 * k + 0] = f; hist[5 * k + 1] = g; hist[5 * k + 3] = 0; hist[5 * k + 4] = its;
        if (check_convergence && orc_is_convergence(f, fprev, epsvalue)) {
            hist[5 * k + 2] = now_s() - tstart; hist[5 * k + 3] = 1;
            k++;
            break;
        }
        if (opt_kind == OPT_OC) {
            orc_oc_update(nelem, optp[0], optp[1], optp[2], optp[3], optp[4], fkind, rowptr, nbr, w, beta, weightlimit, scale1, s, dfds, dgds, NULL);
        } else {
            orc_mma_update(mma, s, dfds, &g, dgds);
        }
        fprev = f;
        t1 = now_s(); if (phase) phase[7] += t1 - t0;
        hist[5 * k + 2] = now_s() - tstart;
    }
    if (u_out) memcpy(u_out, u, sizeof(double) * nd);
    orc_system_free(S); orc_mma_free(mma);
    free(n2g); free(ufix); free(Emod); free(sol); free(dfdrho); free(dgdrho); free(dfds); free(dgds); free(u);
    return k;
}

/* ------------------------------------------------------------------------------------------------------------
 * Level-set topology optimisation  sample/optimize/sample_optimize_levelset.cpp:75-192   
 *   PlaneStressStiffness<Q4, Gauss4Square>                     src/FEM/Equation/PlaneStress.h:21-58
 *   ReactionDiffusionConsistentMass / Stiffness / Reaction     src/FEM/Equation/ReactionDiffusion.h:21-48, 82-107, 111-145
 *   InterpolateNodalFromElemental / ElementalFromNodal         src/FEM/Equation/General.h:193-207, 225-235
 * prm = { Vmax, tau, E0, Emin, nu, nvol, dt, d, p };  hist[3*t] = { objective[t], vol, lambda };  same contract as
 * ref_levelset_run (oracle/ref_shim.cpp).
 * ---------------------------------------------------------------------------------------------------------- */
static void shape_n_q4(const double* r, double* N) {          /* ShapeFunction4Square::N  ShapeFunction.h:175-182 */
    N[0] = 0.25 * (1.0 - r[0]) * (1.0 - r[1]); N[1] = 0.25 * (1.0 + r[0]) * (1.0 - r[1]);
    N[2] = 0.25 * (1.0 + r[0]) * (1.0 + r[1]); N[3] = 0.25 * (1.0 - r[0]) * (1.0 + r[1]);
}

/* Me (4x4), Ke = D dNdX^T dNdX (4x4), Fe = int N C (u - lambda) with u interpolated from the nodal field un */
static void rd_element_q4(const double* xe, double Dcoef, const double* un, double Ccoef, double lambda, double* Me, double* Ke, double* Fe) {
    memset(Me, 0, sizeof(double) * 16); memset(Ke, 0, sizeof(double) * 16); memset(Fe, 0, sizeof(double) * 4);
    for (int g = 0; g < 4; g++) {
        double r[3], w[3], N[4], dNdr[8], dXdr[4], inv[4], dNdX[8];
        quad_point(QUAD_G4SQ, g, r, w);
        shape_n_q4(r, N);
        shape_dndr(SHAPE_Q4, r, dNdr);
        matmul(2, 4, 2, dNdr, xe, dXdr);
        double J = det_d(2, dXdr);
        inv_d(2, dXdr, inv);
        matmul(2, 2, 4, inv, dNdr, dNdX);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Me[i * 4 + j] += N[i] * N[j] * J * w[0] * w[1];
        double DBt[8], BtB[16];                                  /* (_D*dNdX.Transpose())*dNdX*J*w0*w1 */
        for (int i = 0; i < 4; i++) for (int k = 0; k < 2; k++) DBt[i * 2 + k] = dNdX[k * 4 + i] * Dcoef;
        matmul(4, 2, 4, DBt, dNdX, BtB);
        for (int i = 0; i < 16; i++) Ke[i] += BtB[i] * J * w[0] * w[1];
        double u = 0.0;
        for (int i = 0; i < 4; i++) u += N[i] * un[i];
        double f = Ccoef * (u - lambda);
        for (int i = 0; i < 4; i++) Fe[i] += N[i] * f * J * w[0] * w[1];
    }
}

int orc_levelset_run(int nnode, const double* coords, int nelem, const int* conn,
                     int nfixed, const int* fnode, const int* fdof, const double* fval,
                     int nload, const int* lnode, const int* ldof, const double* lval,
                     int nphi, const int* pnode, const double* prm, int tmax,
                     double* phi, double* str, double* u_out, double* hist, int* converged) {
    const double Vmax = prm[0], tau = prm[1], E0 = prm[2], Emin = prm[3], nu = prm[4], nvol = prm[5], dt = prm[6], d = prm[7], p = prm[8];
    const int eq = PHYS_PLANESTRESS | (SHAPE_Q4 << 8) | (QUAD_G4SQ << 16);
    const double A1 = -1.5 * (1.0 - nu) * (1.0 - 14.0 * nu + 15.0 * pow(nu, 2.0)) * E0 / ((1.0 + nu) * (7.0 - 5.0 * nu) * pow(1.0 - 2.0 * nu, 2.0));
    const double A2 = 7.5 * (1.0 - nu) * E0 / ((1.0 + nu) * (7.0 - 5.0 * nu));
    const double c = A1 / (A1 + 2.0 * A2);
    /* displacement system */
    int* n2g = (int*)malloc(sizeof(int) * (size_t)nnode * 2);
    double* ufix = (double*)malloc(sizeof(double) * (size_t)nnode * 2);
    int kdeg = orc_dofmap(nnode, 2, nfixed, fnode, fdof, fval, n2g, ufix);
    orc_system* K = orc_pattern(nnode, 2, 4, nelem, conn, n2g, kdeg);
    /* level-set system: phi = 0 on the listed nodes */
    int* n2g2 = (int*)malloc(sizeof(int) * (size_t)nnode);
    int* pdof = (int*)calloc((size_t)nphi, sizeof(int));
    double* pval = (double*)calloc((size_t)nphi, sizeof(double));
    int tdeg = orc_dofmap(nnode, 1, nphi, pnode, pdof, pval, n2g2, NULL);
    orc_system* T = orc_pattern(nnode, 1, 4, nelem, conn, n2g2, tdeg);
    double* Emod = (double*)malloc(sizeof(double) * (size_t)nelem), *x = (double*)malloc(sizeof(double) * (size_t)kdeg);
    double* u = (double*)malloc(sizeof(double) * (size_t)nnode * 2), *TD = (double*)malloc(sizeof(double) * (size_t)nelem);
    double* TDN = (double*)malloc(sizeof(double) * (size_t)nnode), *objective = (double*)calloc((size_t)tmax, sizeof(double));
    int* count = (int*)malloc(sizeof(int) * (size_t)nnode);
    double* y = (double*)malloc(sizeof(double) * (size_t)tdeg);
    double volInit = 0.0;
    for (int i = 0; i < nelem; i++) volInit += str[i];
    volInit /= (double)nelem;
    *converged = 0;
    int t = 0;
    for (; t < tmax; t++) {
        for (int i = 0; i < nelem; i++) Emod[i] = Emin + str[i] * (E0 - Emin);
        orc_assemble_numeric(K, eq, coords, nelem, conn, n2g, ufix, Emod, nu, 1.0, nload, lnode, ldof, lval, NULL);
        double relres;
        orc_solve(K, NULL, 1, K->F, 100000, 1.0e-10, x, &relres);
        for (size_t i = 0; i < (size_t)nnode * 2; i++) u[i] = (n2g[i] != -1) ? x[n2g[i]] : ufix[i];
        memcpy(u_out, u, sizeof(double) * (size_t)nnode * 2);
        /* objective and topological derivative (driver :112-122) */
        for (int e = 0; e < nelem; e++) {
            const int* el = conn + (size_t)e * 4;
            double xe[8], ue[8], Ke[64], Keue[8];
            for (int a = 0; a < 4; a++) for (int k = 0; k < 2; k++) { xe[a * 2 + k] = coords[(size_t)el[a] * 2 + k]; ue[a * 2 + k] = u[(size_t)el[a] * 2 + k]; }
            orc_element_matrix(eq, xe, Emod[e], nu, 1.0, Ke);
            for (int i = 0; i < 8; i++) { double v = 0.0; for (int j = 0; j < 8; j++) v += Ke[i * 8 + j] * ue[j]; Keue[i] = v; }
            double w = 0.0;
            for (int i = 0; i < 8; i++) w += ue[i] * Keue[i];
            objective[t] += w;
            orc_element_matrix(eq, xe, (A1 + 2.0 * A2) * (1.0 - pow(c, 2.0)), c, 1.0, Ke);
            for (int i = 0; i < 8; i++) { double v = 0.0; for (int j = 0; j < 8; j++) v += Ke[i * 8 + j] * ue[j]; Keue[i] = v; }
            const double a = 1.0e-4 + str[e] * (1.0 - 1.0e-4);
            w = 0.0;
            for (int i = 0; i < 8; i++) w += (ue[i] * a) * Keue[i];          /* (a*ue)*(Ke*ue): scalar*Vector, then the dot */
            TD[e] = w;
        }
        for (int i = 0; i < nnode; i++) { TDN[i] = 0.0; count[i] = 0; }
        for (int e = 0; e < nelem; e++) for (int a = 0; a < 4; a++) { TDN[conn[(size_t)e * 4 + a]] += TD[e]; count[conn[(size_t)e * 4 + a]]++; }
        for (int i = 0; i < nnode; i++) TDN[i] /= (double)count[i];
        double vol = 0.0;
        for (int i = 0; i < nelem; i++) vol += str[i];
        vol /= (double)nelem;
        const double frac = 1.0 - (t + 1) / (double)nvol;
        const double ex = Vmax + (volInit - Vmax) * (frac > 0.0 ? frac : 0.0);
        double tsum = 0.0;
        for (int i = 0; i < nnode; i++) tsum += TDN[i];
        const double lambda = tsum / (double)nnode * exp(p * ((vol - ex) / ex + d));
        hist[3 * t] = objective[t]; hist[3 * t + 1] = vol; hist[3 * t + 2] = lambda;
        if (t > nvol && fabs(vol - Vmax) < 0.005) {
            int ok = 1;
            for (int k = 1; k <= 5; k++) ok = ok && (fabs(objective[t] - objective[t - k]) < 0.01 * fabs(objective[t]));
            if (ok) { *converged = 1; t++; break; }
        }
        double Cc = 0.0;
        for (int i = 0; i < nnode; i++) Cc += fabs(TDN[i]);
        Cc = nelem / Cc;
        /* reaction-diffusion step (driver :147-176); phi is held at 0 on the listed nodes (SetDirichlet) */
        for (int i = 0; i < nphi; i++) phi[pnode[i]] = 0.0;
        memset(T->data, 0, sizeof(double) * (size_t)T->indptr[T->n]);
        memset(T->F, 0, sizeof(double) * (size_t)T->n);
        for (int e = 0; e < nelem; e++) {
            const int* el = conn + (size_t)e * 4;
            double xe[8], un[4], phie[4], Me[16], Ke[16], Fe[4], Te[16], Ye[4];
            for (int a = 0; a < 4; a++) { xe[a * 2] = coords[(size_t)el[a] * 2]; xe[a * 2 + 1] = coords[(size_t)el[a] * 2 + 1]; un[a] = TDN[el[a]]; phie[a] = phi[el[a]]; }
            rd_element_q4(xe, tau * nelem, un, Cc, lambda, Me, Ke, Fe);
            for (int i = 0; i < 16; i++) Te[i] = Me[i] / dt + Ke[i];
            for (int i = 0; i < 4; i++) { double v = 0.0; for (int j = 0; j < 4; j++) v += (Me[i * 4 + j] / dt) * phie[j]; Ye[i] = v + Fe[i]; }
            for (int i = 0; i < 4; i++) {
                int r = n2g2[el[i]];
                if (r == -1) continue;
                for (int j = 0; j < 4; j++) {
                    int cc = n2g2[el[j]];
                    if (cc != -1) T->data[find_col(T, r, cc)] += Te[i * 4 + j];
                    else T->F[r] -= Te[i * 4 + j] * phi[el[j]];
                }
            }
            for (int i = 0; i < 4; i++) { int r = n2g2[el[i]]; if (r != -1) T->F[r] += Ye[i]; }
        }
        orc_solve(T, NULL, 1, T->F, 100000, 1.0e-10, y, &relres);
        for (int i = 0; i < nnode; i++) {
            if (n2g2[i] != -1) phi[i] = y[n2g2[i]];
            phi[i] = fmax(fmin(1.0, phi[i]), -1.0);
        }
        for (int e = 0; e < nelem; e++) {
            double v = 0.0;
            for (int a = 0; a < 4; a++) v += phi[conn[(size_t)e * 4 + a]];
            v /= 4.0;
            str[e] = (v < 0.0) ? 0.0 : 1.0;
        }
    }
    free(n2g); free(ufix); free(n2g2); free(pdof); free(pval); free(Emod); free(x); free(u); free(TD); free(TDN); free(objective); free(count); free(y);
    orc_system_free(K); orc_system_free(T);
    return t;
}
