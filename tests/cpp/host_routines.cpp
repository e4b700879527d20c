// tests/cpp/host_routines.cpp -- the host-side helpers of the element / controller headers that take no device path: Area, Volume,
// CenterOfGravity, ElementVector (both overloads), WeakSpring, LagrangeInterpolation(+Derivative) (FEM/Equation/General.h),
// HeatTransferSurfaceFlux (HeatTransfer.h), ShapeFunction3Line (ShapeFunction.h), SetDirichlet / SetPeriodic / Renumbering
// (BoundaryCondition.h, Assembling.h), and the host-side integrals of Homogenization.h (unit-strain load columns, homogenised constitutive
// matrix, check integral).  Built against the reference's headers for the golden (tests/golden/make_golden.py routines ->
// tests/golden/host_routines.txt) and against the header mirror in tests/test_host_routines.py.
#include <cstdio>
#include <cmath>
#include <vector>
#include "LinearAlgebra/Models/Vector.h"
#include "LinearAlgebra/Models/Matrix.h"
#include "LinearAlgebra/Models/LILCSR.h"
#include "LinearAlgebra/Models/CSR.h"
#include "FEM/Controller/ShapeFunction.h"
#include "FEM/Controller/GaussIntegration.h"
#include "FEM/Controller/BoundaryCondition.h"
#include "FEM/Controller/Assembling.h"
#include "FEM/Equation/General.h"
#include "FEM/Equation/HeatTransfer.h"
#include "FEM/Equation/Homogenization.h"

using namespace PANSFEM2;

static void vec(const char* name, Vector<double> v) { std::printf("%s", name); for (int i = 0; i < v.SIZE(); i++) std::printf(" %.17g", v(i)); std::printf("\n"); }
static void stdvec(const char* name, const std::vector<double>& v) { std::printf("%s", name); for (double x : v) std::printf(" %.17g", x); std::printf("\n"); }
static void mat(const char* name, Matrix<double> m) {
    std::printf("%s %d x %d\n", name, m.ROW(), m.COL());
    for (int i = 0; i < m.ROW(); i++) { for (int j = 0; j < m.COL(); j++) std::printf(" %.17g", m(i, j)); std::printf("\n"); }
}
static void n2e(const char* name, const std::vector<std::vector<std::pair<int, int> > >& v) {
    std::printf("%s", name);
    for (const auto& node : v) { std::printf(" ["); for (const auto& d : node) std::printf("(%d,%d)", d.first, d.second); std::printf("]"); }
    std::printf("\n");
}
static void numbering(const char* name, const std::vector<std::vector<int> >& v) {
    std::printf("%s", name);
    for (const auto& node : v) { std::printf(" ["); for (int d : node) std::printf("%d ", d); std::printf("]"); }
    std::printf("\n");
}

int main() {
    //----------measures and centres----------
    std::vector<Vector<double> > x2 = { Vector<double>({ 0.1, -0.2 }), Vector<double>({ 2.3, 0.1 }), Vector<double>({ 2.0, 1.9 }), Vector<double>({ -0.3, 1.4 }),
                                       Vector<double>({ 1.2, -0.1 }), Vector<double>({ 2.2, 1.0 }), Vector<double>({ 0.9, 1.7 }), Vector<double>({ -0.1, 0.6 }) };
    std::vector<int> q4 = { 0, 1, 2, 3 }, q8 = { 0, 1, 2, 3, 4, 5, 6, 7 }, t3 = { 0, 1, 2 };
    std::printf("area Q4 G4 %.17g\n", Area<double, ShapeFunction4Square, Gauss4Square>(x2, q4));
    std::printf("area Q4 G1 %.17g\n", Area<double, ShapeFunction4Square, Gauss1Square>(x2, q4));
    std::printf("area Q8 G9 %.17g\n", Area<double, ShapeFunction8Square, Gauss9Square>(x2, q8));
    std::printf("area T3 G1 %.17g\n", Area<double, ShapeFunction3Triangle, Gauss1Triangle>(x2, t3));
    std::printf("area T3 G3 %.17g\n", Area<double, ShapeFunction3Triangle, Gauss3Triangle>(x2, t3));
    std::vector<Vector<double> > x3 = { Vector<double>({ 0, 0, 0 }), Vector<double>({ 1.1, 0.1, 0 }), Vector<double>({ 1.0, 1.2, 0.1 }), Vector<double>({ -0.1, 1.0, 0 }),
                                       Vector<double>({ 0.1, 0, 0.9 }), Vector<double>({ 1.2, 0, 1.0 }), Vector<double>({ 1.0, 1.1, 1.2 }), Vector<double>({ 0, 0.9, 1.1 }) };
    std::vector<int> h8 = { 0, 1, 2, 3, 4, 5, 6, 7 }, t4 = { 0, 1, 3, 4 };
    std::printf("volume H8 G8 %.17g\n", Volume<double, ShapeFunction8Cubic, Gauss8Cubic>(x3, h8));
    std::printf("volume H8 G27 %.17g\n", Volume<double, ShapeFunction8Cubic, Gauss27Cubic>(x3, h8));
    std::printf("volume T4 G1 %.17g\n", Volume<double, ShapeFunction4Tetrahedron, Gauss1Tetrahedron>(x3, t4));
    vec("centre Q4", CenterOfGravity(x2, q4)); vec("centre T3", CenterOfGravity(x2, t3)); vec("centre H8", CenterOfGravity(x3, h8));

    //----------element vectors, weak spring, Lagrange basis----------
    std::vector<Vector<double> > u(8, Vector<double>(3));
    for (int i = 0; i < 8; i++) for (int d = 0; d < 3; d++) u[i](d) = 10.0*i + d + 0.25;
    std::vector<std::vector<std::pair<int, int> > > map2 = { { { 0, 0 }, { 1, 1 } }, { { 0, 2 }, { 1, 3 } }, { { 0, 4 }, { 1, 5 } } };
    vec("element vector", ElementVector(u, map2, { 5, 2, 7 }));
    std::vector<std::vector<std::vector<std::pair<int, int> > > > maps = { { { { 0, 0 }, { 1, 1 } }, { { 0, 2 }, { 1, 3 } } }, { { { 2, 4 } }, { { 2, 5 } }, { { 2, 6 } } } };
    vec("element vector (groups)", ElementVector(u, maps, { { 1, 6 }, { 0, 3, 4 } }));
    Matrix<double> Ks;
    std::vector<std::vector<std::pair<int, int> > > ns;
    WeakSpring<double>(Ks, ns, { 4, 1, 6 }, { 0, 1 }, x2, 1.0e-9);
    mat("weak spring", Ks); n2e("weak spring map", ns);
    std::vector<double> xs = { 1.0e-3, 0.2, 0.4, 0.6, 0.8, 0.999 };
    for (double at : { 0.5, 1.0e-3, 0.73 }) { stdvec("lagrange", LagrangeInterpolation(xs, at)); stdvec("lagrange'", LagrangeInterpolationDerivative(xs, at)); }

    //----------surface flux on 2- and 3-node edges----------
    auto flux = [](Vector<double> p) { return 3.0 + 2.0*p(0) - p(1)*p(1); };
    Vector<double> Fe;
    std::vector<std::vector<std::pair<int, int> > > nf;
    HeatTransferSurfaceFlux<double, ShapeFunction2Line, Gauss1Line>(Fe, nf, { 1, 2 }, { 0 }, x2, flux, 0.7); vec("flux 2Line G1", Fe); n2e("flux map", nf);
    HeatTransferSurfaceFlux<double, ShapeFunction2Line, Gauss2Line>(Fe, nf, { 1, 2 }, { 0 }, x2, flux, 0.7); vec("flux 2Line G2", Fe);
    HeatTransferSurfaceFlux<double, ShapeFunction3Line, Gauss2Line>(Fe, nf, { 1, 2, 5 }, { 0 }, x2, flux, 0.7); vec("flux 3Line G2", Fe);
    for (double r : { -1.0, -0.3, 0.0, 0.8 }) {
        vec("3Line N", ShapeFunction3Line<double>::N(Vector<double>({ r })));
        mat("3Line dNdr", ShapeFunction3Line<double>::dNdr(Vector<double>({ r })));
    }
    std::printf("3Line points"); for (auto p : ShapeFunction3Line<double>::Points) std::printf(" %g", p(0)); std::printf(" d %d n %d\n", ShapeFunction3Line<double>::d, ShapeFunction3Line<double>::n);

    //----------homogenisation integrals on a distorted Q4 and a T6----------
    {
        std::vector<Vector<double> > chi0(8, Vector<double>(2)), chi1(8, Vector<double>(2)), chi2(8, Vector<double>(2));
        for (int i = 0; i < 8; i++) for (int d = 0; d < 2; d++) { chi0[i](d) = 0.01*std::sin(1.0 + 3*i + d); chi1[i](d) = -0.02*std::cos(0.5*i - d); chi2[i](d) = 0.005*(i - 3.5)*(d + 1); }
        Matrix<double> Fes;
        std::vector<std::vector<std::pair<int, int> > > nh;
        HomogenizePlaneStrainBodyForce<double, ShapeFunction4Square, Gauss4Square>(Fes, nh, { 0, 1, 2, 3 }, { 0, 1 }, x2, 210000.0, 0.3, 0.8);
        mat("homogenize body force Q4", Fes); n2e("homogenize map", nh);
        mat("homogenize constitutive Q4", HomogenizePlaneStrainConstitutive<double, ShapeFunction4Square, Gauss4Square>(x2, q4, chi0, chi1, chi2, 210000.0, 0.3, 0.8));
        mat("homogenize check Q4", HomogenizePlaneStrainCheck<double, ShapeFunction4Square, Gauss4Square>(x2, q4, chi0, chi1, chi2, 0.8));
        std::vector<int> t6 = { 0, 1, 2, 4, 5, 6 };
        HomogenizePlaneStrainBodyForce<double, ShapeFunction6Triangle, Gauss3Triangle>(Fes, nh, t6, { 0, 1 }, x2, 1.0, 0.25, 1.0);
        mat("homogenize body force T6", Fes);
        mat("homogenize constitutive T6", HomogenizePlaneStrainConstitutive<double, ShapeFunction6Triangle, Gauss3Triangle>(x2, t6, chi0, chi1, chi2, 1.0, 0.25, 1.0));
    }

    //----------every Assembling overload, Disassembling: a 2-element patch with a fixed node and prescribed values----------
    {
        std::vector<std::vector<int> > n2g = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
        std::vector<Vector<double> > uu(5, Vector<double>(2));
        std::vector<std::pair<std::pair<int, int>, double> > fix = { { { 1, 0 }, 0.25 }, { { 1, 1 }, -0.5 }, { { 4, 1 }, 2.0 } };
        SetDirichlet(uu, n2g, fix);
        const int KD = Renumbering(n2g);
        LILCSR<double> K(KD, KD);
        std::vector<double> F(KD, 0.0);
        const std::vector<std::vector<int> > patch = { { 0, 1, 2 }, { 2, 1, 4, 3 } };
        for (const auto& el : patch) {
            const int m = 2*(int)el.size();
            Matrix<double> Ke(m, m);
            Vector<double> Fe(m);
            std::vector<std::vector<std::pair<int, int> > > map(el.size(), std::vector<std::pair<int, int> >(2));
            for (int i = 0; i < (int)el.size(); i++) { map[i][0] = std::make_pair(0, 2*i); map[i][1] = std::make_pair(1, 2*i + 1); }
            for (int i = 0; i < m; i++) { Fe(i) = 0.1*(i + 1) - 0.3; for (int j = 0; j < m; j++) Ke(i, j) = 1.0/(1.0 + i + 2.0*j) + (i == j ? 3.0 : 0.0) + 0.01*el[0]; }
            Assembling(K, F, uu, Ke, Fe, n2g, map, el);          // K + F + Fe
            Assembling(K, F, uu, Ke, n2g, map, el);              // K + Dirichlet lift
            Assembling(K, Ke, n2g, map, el);                     // K only
            Assembling(F, Fe, n2g, map, el);                     // F += Fe
            Assembling(F, uu, Ke, n2g, map, el);                 // lift only
        }
        {   // one matrix over two node groups: 2-dof nodes {0, 2} and a 1-dof group {3} (dof 1 of node 3)
            Matrix<double> Ke(5, 5);
            for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) Ke(i, j) = 0.5*i - 0.25*j + (i == j ? 2.0 : 0.0);
            std::vector<std::vector<std::vector<std::pair<int, int> > > > maps = { { { { 0, 0 }, { 1, 1 } }, { { 0, 2 }, { 1, 3 } } }, { { { 1, 4 } } } };
            Assembling(K, F, uu, Ke, n2g, maps, { { 0, 1 }, { 3 } });
        }
        Assembling(F, std::vector<std::pair<std::pair<int, int>, double> >({ { { 0, 1 }, -100.0 }, { { 1, 0 }, 7.0 }, { { 3, 0 }, 1.5 } }), n2g);
        numbering("patch numbering", n2g);
        std::printf("KDEGREE %d\n", KD);
        CSR<double> A(K);
        for (int i = 0; i < KD; i++) { std::printf("row %d:", i); for (int j = 0; j < KD; j++) std::printf(" %.17g", A.get(i, j)); std::printf("\n"); }
        stdvec("patch F", F);
        std::vector<double> sol(KD);
        for (int i = 0; i < KD; i++) sol[i] = 10.0 + i;
        Disassembling(uu, sol, n2g);
        for (auto& v : uu) vec("u", v);
    }

    //----------numbering: Dirichlet marks, periodic pairs----------
    std::vector<std::vector<int> > n2g(7, std::vector<int>(2, 0));
    std::vector<std::pair<std::pair<int, int>, double> > fixed = { { { 0, 0 }, 0.0 }, { { 0, 1 }, 0.0 }, { { 3, 1 }, 0.5 } };
    SetDirichlet(n2g, fixed);
    numbering("marked", n2g);
    std::printf("KDEGREE %d\n", Renumbering(n2g)); numbering("renumbered", n2g);
    std::vector<std::vector<int> > n2p(7, std::vector<int>(2, 0));
    std::printf("KDEGREE periodic %d\n", SetPeriodic(n2p, { { 0, 5 }, { 1, 6 }, { 2, 4 } })); numbering("periodic", n2p);
    RemoveBoundaryConditions(n2p); numbering("removed", n2p);
    return 0;
}
