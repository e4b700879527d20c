//  sample/planestrain/sample_planestrain.cpp of the reference with the batched, device-resident API of pansfem2_b200:
//  stiffness of all elements, the body force of all elements and the traction of all loaded edges are each ONE call
//  (B200::AssembleBatched / B200::AssembleLoadVector) instead of a per-element loop; the functors are the sample's own.
//  Prints the displacements node by node (the sample's result.vtk carries the same numbers).
#include <iostream>
#include <iomanip>
#include <vector>

#include "../../src/LinearAlgebra/Models/Vector.h"
#include "../../src/FEM/Controller/ShapeFunction.h"
#include "../../src/FEM/Controller/GaussIntegration.h"
#include "../../src/FEM/Controller/Assembling.h"
#include "../../src/B200/Batched.h"

using namespace PANSFEM2;

int main() {
    //  model of sample_planestrain.cpp:21-28
    std::vector<Vector<double> > x = { { 0, 0 }, { 1, 0 }, { 2, 0 }, { 2, 1 }, { 1, 1 }, { 0, 1 } };
    std::vector<std::vector<int> > elements = { { 0, 1, 4 }, { 1, 2, 3 }, { 3, 4, 1 }, { 4, 5, 0 } };
    std::vector<std::vector<int> > edges = { { 3, 4 }, { 4, 5 } };
    B200::BcList ufixed = { { { 0, 0 }, 0 }, { { 0, 1 }, 0 }, { { 5, 0 }, 0 } };
    B200::BcList qfixed = { { { 3, 1 }, -100.0 } };

    B200::Model model(x, elements, 2, ufixed);
    std::vector<double> F;
    pf2_csr* K = B200::AssembleBatched<B200::PlaneStrainStiffnessTag<ShapeFunction3Triangle, Gauss1Triangle> >(
        model, std::vector<double>(elements.size(), 210000.0), 0.3, 1.0, qfixed, F);
    B200::AssembleLoadVector<ShapeFunction3Triangle, Gauss1Triangle>(model, elements, [](Vector<double> _x) {
        Vector<double> f = { 0.0, -300.0 };
        return f;
    }, 1.0, F);
    B200::AssembleLoadVector<ShapeFunction2Line, Gauss1Line>(model, edges, [](Vector<double> _x) {
        Vector<double> f = { 0.0, -200.0 };
        return f;
    }, 1.0, F);
    std::vector<double> result = B200::SolveResident(K, PF2_SOLVER_CG, F, 100000, 1.0e-10);

    std::vector<Vector<double> > u(x.size(), Vector<double>(2));
    std::vector<std::vector<int> > nodetoglobal = model.NodeToGlobal();
    for (size_t i = 0; i < x.size(); i++) for (int d = 0; d < 2; d++) u[i](d) = nodetoglobal[i][d] >= 0 ? result[nodetoglobal[i][d]] : 0.0;
    std::cout << std::setprecision(12);
    for (size_t i = 0; i < x.size(); i++) std::cout << "u " << i << " " << u[i](0) << " " << u[i](1) << std::endl;
    return 0;
}
