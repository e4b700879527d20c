//  pansfem2_b200/sample/optimize/sample_optimize_density_families.cpp
//  The SIMP cantilever of sample/optimize/sample_optimize_density_oc.cpp on OTHER element selections of the reference, driven
//  through the batched API (B200/Batched.h): the same template tags the reference's element routines take select the kernel.
//      usage: sample_optimize_density_families [t3|q8sri] [nx ny] [iterations]
//  t3    : PlaneStressStiffness<ShapeFunction3Triangle, Gauss1Triangle>, every SquareMesh cell cut in two
//  q8sri : PlaneStrainStiffnessSRI<ShapeFunction8Square, Gauss4Square (volumetric), Gauss9Square (deviatoric)>
//  Prints one line per design iteration (objective, weight); tests/test_gpu_cpp_dropin.py replays the run in the oracle.
#include <iostream>
#include <iomanip>
#include <string>
#include <vector>
#include <map>
#include <cmath>

#include "../../src/LinearAlgebra/Models/Vector.h"
#include "../../src/FEM/Controller/ShapeFunction.h"
#include "../../src/FEM/Controller/GaussIntegration.h"
#include "../../src/PrePost/Mesher/SquareMesh.h"
#include "../../src/FEM/Equation/General.h"
#include "../../src/Optimize/Filter/DensityFilter.h"
#include "../../src/B200/Batched.h"

using namespace PANSFEM2;

template<class TAG>
int Run(std::vector<Vector<double> >& x, std::vector<std::vector<int> >& elements, int nx, int ny, int iterations) {
    B200::BcList ufixed, qfixed;
    for (int i = 0; i < (int)x.size(); i++) {
        if (std::fabs(x[i](0)) < 1.0e-9) { ufixed.push_back({ { i, 0 }, 0.0 }); ufixed.push_back({ { i, 1 }, 0.0 }); }
        if (std::fabs(x[i](0) - nx) < 1.0e-9 && std::fabs(x[i](1) - 0.5*ny) < 1.0e-9) qfixed.push_back({ { i, 1 }, -1.0 });
    }
    //----------Filter lists: all pairs within R of the centres of gravity (sample_optimize_density_oc.cpp:44-60)----------
    std::vector<Vector<double> > cg(elements.size());
    for (size_t i = 0; i < elements.size(); i++) cg[i] = CenterOfGravity(x, elements[i]);
    const double R = 1.5;
    std::vector<std::vector<int> > neighbors(elements.size());
    std::vector<std::vector<double> > w(elements.size());
    for (size_t i = 0; i < elements.size(); i++) for (size_t j = 0; j < elements.size(); j++) {
        const double d = (cg[i] - cg[j]).Norm();
        if (d <= R) { neighbors[i].push_back((int)j); w[i].push_back((R - d)/R); }
    }
    DensityFilter<double> filter(elements.size(), neighbors, w);

    B200::Model model(x, elements, 2, ufixed);
    B200::SimpParameters prm;
    prm.beta_period = 0;
    B200::DesignLoop<TAG> loop(model, filter, PF2_OPT_OC, { 0.5, 0.0, 1.0e4, 1.0e-3, 0.15 }, prm, qfixed, std::vector<double>(elements.size(), 0.5));
    std::cout << std::setprecision(15);
    for (int k = 0; k < iterations; k++) {
        B200::IterationReport it = loop.Iterate(false);
        std::cout << "k = " << k << "\tObjective:\t" << it.f << "\tWeight:\t" << it.g << "\tCG:\t" << it.cg_iterations << std::endl;
    }
    return 0;
}

int main(int argc, char** argv) {
    const std::string family = argc > 1 ? argv[1] : "t3";
    const int nx = argc > 3 ? std::stoi(argv[2]) : 30, ny = argc > 3 ? std::stoi(argv[3]) : 20;
    const int iterations = argc > 4 ? std::stoi(argv[4]) : 5;
    SquareMesh<double> mesh(nx, ny, nx, ny);
    std::vector<Vector<double> > x = mesh.GenerateNodes();
    std::vector<std::vector<int> > quads = mesh.GenerateElements();
    if (family == "t3") {
        std::vector<std::vector<int> > elements;
        for (const auto& q : quads) { elements.push_back({ q[1], q[2], q[0] }); elements.push_back({ q[2], q[3], q[0] }); }
        return Run<B200::PlaneStressStiffnessTag<ShapeFunction3Triangle, Gauss1Triangle> >(x, elements, nx, ny, iterations);
    }
    //  q8: add one node per cell edge, shared between the two cells on it
    std::map<std::pair<int, int>, int> mid;
    std::vector<std::vector<int> > elements;
    for (const auto& q : quads) {
        std::vector<int> e(q);
        for (int a = 0; a < 4; a++) {
            const int n0 = q[a], n1 = q[(a + 1)%4];
            const std::pair<int, int> key(std::min(n0, n1), std::max(n0, n1));
            auto it = mid.find(key);
            if (it == mid.end()) {
                it = mid.insert({ key, (int)x.size() }).first;
                x.push_back((x[n0] + x[n1])/2.0);
            }
            e.push_back(it->second);
        }
        elements.push_back(e);
    }
    return Run<B200::PlaneStrainStiffnessSRITag<ShapeFunction8Square, Gauss4Square, Gauss9Square> >(x, elements, nx, ny, iterations);
}
