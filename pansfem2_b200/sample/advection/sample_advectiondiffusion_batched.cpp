//  pansfem2_b200/sample/advection/sample_advectiondiffusion_batched.cpp
//  The two advection-diffusion problems of the reference's sample/advection (sample_advectiondiffusion_static.cpp: steady SUPG,
//  a = 1 at 60 degrees, k = 1e-6; sample_advectiondiffusion_dynamic.cpp: rotating cone, Crank-Nicolson + SUPG, 100 steps of pi/50)
//  driven through the batched, device-resident API (B200::AdvectionDiffusion): same inputs, same parameters, same VTK fields.
//      usage: sample_advectiondiffusion_batched static|dynamic <model directory> <output.vtk> [steps]
#include <iostream>
#include <fstream>
#include <string>
#include <vector>
#include <cmath>

#include "../../src/LinearAlgebra/Models/Vector.h"
#include "../../src/PrePost/Import/ImportFromCSV.h"
#include "../../src/FEM/Controller/ShapeFunction.h"
#include "../../src/FEM/Controller/GaussIntegration.h"
#include "../../src/FEM/Equation/General.h"
#include "../../src/PrePost/Export/ExportToVTK.h"
#include "../../src/B200/Batched.h"

using namespace PANSFEM2;

int main(int argc, char** argv) {
    if (argc < 4) { std::cerr << "usage: " << argv[0] << " static|dynamic <model directory> <output.vtk> [steps]" << std::endl; return 2; }
    const bool dynamic = std::string(argv[1]) == "dynamic";
    const std::string model_path = std::string(argv[2]) + "/";
    const int steps = argc > 4 ? std::stoi(argv[4]) : 100;

    std::vector<Vector<double> > x;
    ImportNodesFromCSV(x, model_path + "Node.csv");
    std::vector<std::vector<int> > elements;
    ImportElementsFromCSV(elements, model_path + "Element.csv");
    std::vector<std::pair<std::pair<int, int>, double> > ufixed;
    ImportDirichletFromCSV(ufixed, model_path + (dynamic ? "DirichletD.csv" : "Dirichlet.csv"));

    std::vector<Vector<double> > T(x.size(), Vector<double>(1));
    B200::Model model(x, elements, 1, ufixed);

    if (!dynamic) {
        //----------sample_advectiondiffusion_static.cpp:29-58----------
        const double a = 1.0, theta = 60.0, k = 1.0e-6;
        typedef B200::AdvectionDiffusionTag<ShapeFunction3Triangle, Gauss1Triangle, PF2_ADV_ADVECTION | PF2_ADV_DIFFUSION | PF2_ADV_SUPG> Eq;
        B200::AdvectionDiffusion<Eq> problem(model, { Vector<double>({ a*cos(theta*M_PI/180.0), a*sin(theta*M_PI/180.0) }) }, k, T);
        const int iters = problem.Solve();
        std::cout << "BiCGSTAB iterations = " << iters << std::endl;
        problem.Get(T);
    } else {
        //----------sample_advectiondiffusion_dynamic.cpp:27-75----------
        Vector<double> O = Vector<double>({ 0.5, 0.75 });
        for (size_t i = 0; i < x.size(); i++) {
            const double r = (x[i] - O).Norm();
            if (r <= 0.25) T[i](0) = 0.5*(cos(4.0*M_PI*r) + 1.0);
        }
        const double dt = M_PI/50.0, theta = 0.5, k = 0.0;
        std::vector<Vector<double> > velocity;
        for (auto element : elements) {
            Vector<double> ge = CenterOfGravity(x, element);
            velocity.push_back(Vector<double>({ -(ge(1) - 0.5), ge(0) - 0.5 }));
        }
        typedef B200::AdvectionDiffusionTag<ShapeFunction3Triangle, Gauss1Triangle,
                                            PF2_ADV_MASS | PF2_ADV_MASS_SUPG | PF2_ADV_ADVECTION | PF2_ADV_DIFFUSION | PF2_ADV_SUPG> Eq;
        B200::AdvectionDiffusion<Eq> problem(model, velocity, k, T);
        for (int t = 0; t < steps; t++) {
            const int iters = problem.Step(dt, theta);
            std::cout << "t = " << t << "\tBiCGSTAB iterations = " << iters << std::endl;
        }
        problem.Get(T);
    }

    std::ofstream fout(argv[3]);
    MakeHeadderToVTK(fout);
    AddPointsToVTK(x, fout);
    AddElementToVTK(elements, fout);
    AddElementTypes(std::vector<int>(elements.size(), 5), fout);
    AddPointScalers(T, "T", fout, true);
    fout.close();
    return 0;
}
