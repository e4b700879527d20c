// tests/cpp/host_elements.cu -- TEST INFRASTRUCTURE.  The element routines of the library (pansfem2_b200/csrc/element.cuh,
// element_generic.cuh, element_advdiff.cuh) are __host__ __device__; this file instantiates them ON THE HOST, with the library's own
// eq-code decoder (decode_eq / make_spec from libpansfem2_b200.so), so that the CPU test suite can compare the very source the kernels
// run against the oracle on a machine without a GPU (tests/test_host_elements.py).  It is never linked into the product and the
// library never calls the element routines from host code.
#include "types.cuh"
#include "element_advdiff.cuh"

namespace pf2 { ElemSpec make_spec(const EqInfo& q); ElemSpecD make_spec_d(const EqInfo& q, const double D[9]); }
using namespace pf2;

template <int KIND, int SHAPE>
static void rows_generic(const ElemSpec& sp, const double* xe, double E, double t, double* Ke) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE, NDOF = KindTraits<KIND>::NDOF, M = NPE * NDOF;
    double X[NPE][DIM];
    for (int n = 0; n < NPE; n++) for (int k = 0; k < DIM; k++) X[n][k] = xe[n * DIM + k];
    for (int a = 0; a < NPE; a++) {         // what thread `a` of element_generic_kernel does
        double acc[NDOF][M];
        generic_rows<KIND, SHAPE>(X, a, sp, t, acc);
        for (int i = 0; i < NDOF; i++) for (int j = 0; j < M; j++) Ke[(a * NDOF + i) * M + j] = E * acc[i][j];
    }
}

template <int KIND, int SHAPE>
static double energy_generic(const ElemSpec& sp, const double* xe, const double* ue_in, double t, double* fe_out) {
    constexpr int DIM = ShapeTraits<SHAPE>::DIM, NPE = ShapeTraits<SHAPE>::NPE, NDOF = KindTraits<KIND>::NDOF;
    double X[NPE][DIM], ue[NPE][NDOF], fe[NPE][NDOF];
    for (int n = 0; n < NPE; n++) {
        for (int k = 0; k < DIM; k++) X[n][k] = xe[n * DIM + k];
        for (int k = 0; k < NDOF; k++) { ue[n][k] = ue_in[n * NDOF + k]; fe[n][k] = 0.0; }
    }
    const double w = generic_energy<KIND, SHAPE, true>(X, ue, sp, t, fe);
    for (int n = 0; n < NPE; n++) for (int k = 0; k < NDOF; k++) fe_out[n * NDOF + k] = fe[n][k];
    return w;
}

template <int SHAPE>
static void rows_advdiff(const AdvSpec& sp, const double* xe, double* KK, double* MM) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE;
    double X[NPE][2];
    for (int n = 0; n < NPE; n++) { X[n][0] = xe[n * 2]; X[n][1] = xe[n * 2 + 1]; }
    for (int a = 0; a < NPE; a++) {
        double accK[NPE], accM[NPE];
        advdiff_rows<SHAPE>(X, a, sp, accK, accM);
        for (int b = 0; b < NPE; b++) { KK[a * NPE + b] = accK[b]; MM[a * NPE + b] = accM[b]; }
    }
}

// Ke as pf2_element_matrix_d returns it (general_rows: PlaneStiffness / Bbar / WilsonTaylor with a caller-supplied D)
template <int SHAPE>
static void rows_general(const ElemSpecD& sp, const double* xe, double t, double* Ke) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE, M = NPE * 2;
    double X[NPE][2];
    for (int n = 0; n < NPE; n++) { X[n][0] = xe[n * 2]; X[n][1] = xe[n * 2 + 1]; }
    for (int a = 0; a < NPE; a++) {
        double acc[2][M];
        general_rows<SHAPE>(X, a, sp, t, acc);
        for (int i = 0; i < 2; i++) for (int j = 0; j < M; j++) Ke[(a * 2 + i) * M + j] = acc[i][j];
    }
}

#define DISPATCH(q, CALL)                                                                                               \
    do {                                                                                                                \
        if ((q).kind == KIND_SOLID3D) {                                                                                 \
            if ((q).shape == PF2_SHAPE_TET4) { CALL(KIND_SOLID3D, SH_TET4); }                                           \
            else if ((q).shape == PF2_SHAPE_HEX8) { CALL(KIND_SOLID3D, SH_HEX8); }                                      \
            else { CALL(KIND_SOLID3D, SH_HEX20); }                                                                      \
        } else {                                                                                                        \
            const int k__ = (q).kind;                                                                                   \
            if ((q).shape == PF2_SHAPE_T3) { DISPATCH_KIND(k__, SH_T3, CALL); }                                         \
            else if ((q).shape == PF2_SHAPE_T6) { DISPATCH_KIND(k__, SH_T6, CALL); }                                    \
            else if ((q).shape == PF2_SHAPE_Q4) { DISPATCH_KIND(k__, SH_Q4, CALL); }                                    \
            else { DISPATCH_KIND(k__, SH_Q8, CALL); }                                                                   \
        }                                                                                                               \
    } while (0)
#define DISPATCH_KIND(k, S, CALL)                              \
    do {                                                       \
        if ((k) == KIND_HEAT2D) { CALL(KIND_HEAT2D, S); }      \
        else if ((k) == KIND_MASS2D) { CALL(KIND_MASS2D, S); } \
        else if ((k) == KIND_MASS2D_V) { CALL(KIND_MASS2D_V, S); } \
        else { CALL(KIND_ELAST2D, S); }                        \
    } while (0)

extern "C" {

// Ke as pf2_element_matrix returns it; specialised = 1 routes the three fast selections through element.cuh like the library does
int pf2host_element_matrix(int eq, const double* xe, double E, double V, double t, int specialised, double* Ke) {
    EqInfo q;
    PF2_TRY(decode_eq(eq, V, &q));
    if (q.kind == KIND_ADVDIFF2D) {         // (E, V, t) = (ax, ay, k)
        const AdvSpec sp = { q.quad, q.quad2, E, V, t };
        double KK[64], MM[64];
        if (q.shape == PF2_SHAPE_T3) rows_advdiff<SH_T3>(sp, xe, KK, MM);
        else if (q.shape == PF2_SHAPE_T6) rows_advdiff<SH_T6>(sp, xe, KK, MM);
        else if (q.shape == PF2_SHAPE_Q4) rows_advdiff<SH_Q4>(sp, xe, KK, MM);
        else rows_advdiff<SH_Q8>(sp, xe, KK, MM);
        for (int i = 0; i < q.npe * q.npe; i++) Ke[i] = KK[i] + MM[i];
        return PF2_OK;
    }
    if (q.fast && specialised) {
        if (q.legacy == PF2_EQ_PLANESTRAIN) {
            double X[4][2], acc[2][8];
            for (int n = 0; n < 4; n++) for (int k = 0; k < 2; k++) X[n][k] = xe[n * 2 + k];
            for (int a = 0; a < 4; a++) { planestrain_rows(X, a, V, t, acc); for (int i = 0; i < 2; i++) for (int j = 0; j < 8; j++) Ke[(a * 2 + i) * 8 + j] = E * acc[i][j]; }
        } else if (q.legacy == PF2_EQ_HEAT) {
            double X[4][2], acc[1][4];
            for (int n = 0; n < 4; n++) for (int k = 0; k < 2; k++) X[n][k] = xe[n * 2 + k];
            for (int a = 0; a < 4; a++) { heat_rows(X, a, t, acc); for (int j = 0; j < 4; j++) Ke[a * 4 + j] = E * acc[0][j]; }
        } else {
            double X[8][3], acc[3][24];
            for (int n = 0; n < 8; n++) for (int k = 0; k < 3; k++) X[n][k] = xe[n * 3 + k];
            for (int a = 0; a < 8; a++) { solid_rows(X, a, V, acc); for (int i = 0; i < 3; i++) for (int j = 0; j < 24; j++) Ke[(a * 3 + i) * 24 + j] = E * acc[i][j]; }
        }
        return PF2_OK;
    }
    const ElemSpec sp = make_spec(q);
#define CALL(K, S) rows_generic<K, S>(sp, xe, E, t, Ke)
    DISPATCH(q, CALL);
#undef CALL
    return PF2_OK;
}

// ue^T Ke(E = 1) ue and fe = Ke(E = 1) ue through generic_energy (what sens_generic_kernel evaluates per element)
int pf2host_element_energy(int eq, const double* xe, const double* ue, double V, double t, double* w_out, double* fe) {
    EqInfo q;
    PF2_TRY(decode_eq(eq, V, &q));
    PF2_CHECK(q.kind != KIND_ADVDIFF2D, "no energy form for the advection-diffusion operator");
    const ElemSpec sp = make_spec(q);
#define CALL(K, S) *w_out = energy_generic<K, S>(sp, xe, ue, t, fe)
    DISPATCH(q, CALL);
#undef CALL
    return PF2_OK;
}

int pf2host_element_matrix_d(int eq, const double* xe, const double* D, double t, double* Ke) {
    EqInfo q;
    PF2_TRY(decode_eq(eq, 0.0, &q));
    PF2_CHECK(q.kind == KIND_ELAST2D_D, "not a PF2_PHYS_PLANE_D* selection");
    const ElemSpecD sp = make_spec_d(q, D);
    if (q.shape == PF2_SHAPE_T3) rows_general<SH_T3>(sp, xe, t, Ke);
    else if (q.shape == PF2_SHAPE_T6) rows_general<SH_T6>(sp, xe, t, Ke);
    else if (q.shape == PF2_SHAPE_Q4) rows_general<SH_Q4>(sp, xe, t, Ke);
    else rows_general<SH_Q8>(sp, xe, t, Ke);
    return PF2_OK;
}

// the two group matrices of advdiff_rows separately (K = A + D + AS + SC, M = M + MS)
int pf2host_advdiff_groups(int eq, const double* xe, double ax, double ay, double k, double* KK, double* MM) {
    EqInfo q;
    PF2_TRY(decode_eq(eq, 0.0, &q));
    PF2_CHECK(q.kind == KIND_ADVDIFF2D, "not an advection-diffusion selection");
    const AdvSpec sp = { q.quad, q.quad2, ax, ay, k };
    if (q.shape == PF2_SHAPE_T3) rows_advdiff<SH_T3>(sp, xe, KK, MM);
    else if (q.shape == PF2_SHAPE_T6) rows_advdiff<SH_T6>(sp, xe, KK, MM);
    else if (q.shape == PF2_SHAPE_Q4) rows_advdiff<SH_Q4>(sp, xe, KK, MM);
    else rows_advdiff<SH_Q8>(sp, xe, KK, MM);
    return PF2_OK;
}

}  // extern "C"
