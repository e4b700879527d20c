//  pansfem2_b200/sample/optimize/sample_optimize_levelset_batched.cpp
//  The level-set cantilever of the reference's sample/optimize/sample_optimize_levelset.cpp driven through the batched,
//  device-resident API (B200::LevelSetLoop): same problem, same parameters, same console lines, same VTK fields.
//      usage: sample_optimize_levelset_batched [nx ny] [output.vtk]
#include <iostream>
#include <fstream>
#include <string>
#include <vector>
#include <cmath>

#include "../../src/LinearAlgebra/Models/Vector.h"
#include "../../src/FEM/Controller/ShapeFunction.h"
#include "../../src/FEM/Controller/GaussIntegration.h"
#include "../../src/PrePost/Export/ExportToVTK.h"
#include "../../src/PrePost/Mesher/SquareMesh.h"
#include "../../src/B200/Batched.h"

using namespace PANSFEM2;

int main(int argc, char** argv) {
    const int nx = argc > 2 ? std::stoi(argv[1]) : 60, ny = argc > 2 ? std::stoi(argv[2]) : 40;
    const std::string out = argc > 3 ? argv[3] : (argc == 2 ? argv[1] : "");

    //----------Design region, loads and the phi boundary (sample_optimize_levelset.cpp:41-68)----------
    SquareMesh<double> mesh(nx, ny, nx, ny);
    std::vector<Vector<double> > x = mesh.GenerateNodes();
    std::vector<std::vector<int> > elements = mesh.GenerateElements();
    auto ufixed = mesh.GenerateFixedlist({ 0, 1 }, [](Vector<double> _x) { return std::fabs(_x(0)) < 1.0e-5; });
    auto qfixed = mesh.GenerateFixedlist({ 1 }, [&](Vector<double> _x) { return std::fabs(_x(0) - nx) < 1.0e-5 && std::fabs(_x(1) - 0.5*ny) < 1.0 + 1.0e-5; });
    for (auto& q : qfixed) q.second = -1.0;
    auto phifixed = mesh.GenerateFixedlist({ 0 }, [&](Vector<double> _x) {
        return std::fabs(_x(0)) < 1.0e-5 || std::fabs(_x(0) - nx) < 1.0e-5 || std::fabs(_x(1)) < 1.0e-5 || std::fabs(_x(1) - ny) < 1.0e-5;
    });

    B200::Model model(x, elements, 2, ufixed);
    B200::LevelSetParameters prm;
    B200::LevelSetLoop loop(model, qfixed, phifixed, prm);
    for (int t = 0; t < prm.tmax; t++) {
        B200::LevelSetReport it = loop.Iterate();
        if (it.converged) { std::cout << "----------Convergence----------" << std::endl; break; }
        std::cout << "t = " << t << "\tCompliance = " << it.objective/(double)elements.size() << "\tVolume = " << it.volume << "\tLambda = " << it.lambda << std::endl;
    }

    if (!out.empty()) {
        std::vector<Vector<double> > phi, u;
        std::vector<double> str;
        loop.Get(phi, str, u);
        std::ofstream fout(out);
        MakeHeadderToVTK(fout);
        AddPointsToVTK(x, fout);
        AddElementToVTK(elements, fout);
        AddElementTypes(std::vector<int>(elements.size(), 9), fout);
        AddPointVectors(u, "u", fout, true);
        AddPointScalers(phi, "phi", fout, false);
        AddElementScalers(str, "str", fout, true);
    }
    return 0;
}
