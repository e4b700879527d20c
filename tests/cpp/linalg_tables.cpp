// tests/cpp/linalg_tables.cpp -- the boundary types of the path: every public operation of
// PANSFEM2::Vector<T> and PANSFEM2::Matrix<T> (LinearAlgebra/Models/Vector.h, Matrix.h) plus LILCSR<T> / CSR<T> construction, get / set and
// the per-row sort of CSR(LILCSR&), on awkward numbers, printed with full precision.  Built against the reference's headers for the golden
// (tests/golden/make_golden.py linalg -> tests/golden/linalg_tables.txt) and against the header mirror in tests/test_linalg_tables.py
// (-DMIRROR_HOST_ONLY there: the mirror's CSR::operator* and solvers run on the device and are covered by the GPU suite).
#include <cstdio>
#include <iostream>
#include <sstream>
#include <vector>
#include "LinearAlgebra/Models/Vector.h"
#include "LinearAlgebra/Models/Matrix.h"
#include "LinearAlgebra/Models/LILCSR.h"
#include "LinearAlgebra/Models/CSR.h"

using namespace PANSFEM2;

static void vec(const char* name, Vector<double> v) { std::printf("%s (%d)", name, v.SIZE()); for (int i = 0; i < v.SIZE(); i++) std::printf(" %.17g", v(i)); std::printf("\n"); }
static void mat(const char* name, Matrix<double> m) {
    std::printf("%s %d x %d\n", name, m.ROW(), m.COL());
    for (int i = 0; i < m.ROW(); i++) { for (int j = 0; j < m.COL(); j++) std::printf(" %.17g", m(i, j)); std::printf("\n"); }
}

int main() {
    Vector<double> a({ 1.0/3.0, -2.5, 1.0e-7 }), b({ 4.0, 0.125, -6.02e3 }), z(3);
    vec("a", a); vec("b", b); vec("zero", z);
    vec("a+b", a + b); vec("a-b", a - b); vec("-a", -a); vec("a*2.5", a*2.5); vec("2.5*a", 2.5*a); vec("a/7", a/7.0);
    std::printf("a.b %.17g  |a| %.17g  |b| %.17g\n", a*b, a.Norm(), b.Norm());
    vec("a normal", a.Normal()); vec("a x b", VectorProduct(a, b)); vec("a vstack b", a.Vstack(b)); vec("segment", a.Vstack(b).Segment(1, 4));
    Vector<double> c = a; c += b; vec("+=", c); c -= a; vec("-=", c); c *= 3.0; vec("*=", c); c /= -0.7; vec("/=", c);
    vec("from std::vector", Vector<double>(std::vector<double>({ 9.5, -1.25 })));
    mat("a transpose", a.Transpose()); mat("a * b^T", a*b.Transpose());
    mat("signed zeros of an outer product", Vector<double>({ 0.0, 2.0, -0.0 })*Vector<double>({ -1.0, 0.0, 3.0 }).Transpose()); mat("diagonal(a)", Diagonal(a)); mat("identity", Identity<double>(3));
    std::ostringstream os; os << a; std::printf("vector stream [%s]\n", os.str().c_str());

    Matrix<double> M(3, 3), N(3, 2);
    const double mv[9] = { 2.0, -1.0, 0.5, 1.0/3.0, 4.0, -2.0, 0.25, 1.0e-3, 5.0 }, nv[6] = { 1.0, 2.0, -3.0, 0.5, 7.0, -1.0/9.0 };
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M(i, j) = mv[3*i + j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 2; j++) N(i, j) = nv[2*i + j];
    mat("M", M); mat("N", N); mat("M+M^T", M + M.Transpose()); mat("M-M^T", M - M.Transpose()); mat("-M", -M); mat("M*N", M*N); vec("M*a", M*a);
    mat("M*1.5", M*1.5); mat("1.5*M", 1.5*M); mat("M/3", M/3.0);
    std::printf("det M %.17g\n", M.Determinant()); mat("inverse M", M.Inverse()); mat("M * inverse", M*M.Inverse()); mat("cofactor(1,2)", M.Cofactor(1, 2));
    Matrix<double> M2(2, 2); M2(0, 0) = 3.0; M2(0, 1) = -1.0/7.0; M2(1, 0) = 2.0; M2(1, 1) = 0.3;
    std::printf("det 2x2 %.17g\n", M2.Determinant()); mat("inverse 2x2", M2.Inverse());
    Matrix<double> M4(4, 4);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M4(i, j) = 1.0/(1.0 + i + 2.0*j) + (i == j ? 2.0 : 0.0);
    std::printf("det 4x4 %.17g\n", M4.Determinant()); mat("inverse 4x4", M4.Inverse());
    mat("vstack", M.Vstack(N.Transpose())); mat("hstack", M.Hstack(N)); mat("block", M.Hstack(N).Block(1, 2, 2, 3));
    Matrix<double> P = M; P += M; mat("+=", P); P -= M.Transpose(); mat("-=", P); P *= 0.5; mat("*=", P); P /= 4.0; mat("/=", P);
    mat("from vector", Matrix<double>(a)); vec("to vector", Vector<double>(Matrix<double>(a)));
    std::ostringstream om; om << N; std::printf("matrix stream [%s]\n", om.str().c_str());

    //----------LILCSR insertion order, explicit zeros, get of an absent entry; CSR(LILCSR&) sorts each row----------
    LILCSR<double> L(4, 4);
    L.set(2, 3, 1.5); L.set(2, 0, -2.0); L.set(2, 2, 0.0); L.set(0, 0, 4.0); L.set(3, 1, 7.0); L.set(2, 0, L.get(2, 0) + 0.25); L.set(1, 1, 1.0);
    std::printf("lil %d x %d  get(2,0) %.17g  get(2,1) %.17g  get(2,2) %.17g\n", L.ROWS, L.COLS, L.get(2, 0), L.get(2, 1), L.get(2, 2));
    CSR<double> A(L);
    std::printf("csr %d x %d  get(2,3) %.17g  get(2,1) %.17g  get(3,1) %.17g\n", A.ROWS, A.COLS, A.get(2, 3), A.get(2, 1), A.get(3, 1));
    A.set(2, 3, -9.0);
    std::printf("csr after set  get(2,3) %.17g\n", A.get(2, 3));
    return 0;
}
