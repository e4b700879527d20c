//  pansfem2_b200/sample/optimize/sample_optimize_density_batched.cpp
//  The SIMP cantilever of the reference's sample/optimize/sample_optimize_density_{oc,mma,CONLIN}.cpp driven through the batched,
//  device-resident API (B200/Batched.h).  Same problem, same parameters, same VTK output; the design never leaves the GPU.
//      usage: sample_optimize_density_batched [oc|mma|conlin] [nx ny] [output.vtk]
#include <cstdlib>
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
#include "../../src/FEM/Equation/General.h"
#include "../../src/Optimize/Filter/HeavisideFilter.h"
#include "../../src/B200/Batched.h"

using namespace PANSFEM2;

int main(int argc, char** argv) {
    const std::string optimizer = argc > 1 ? argv[1] : "oc";
    const int nx = argc > 3 ? std::stoi(argv[2]) : 60, ny = argc > 3 ? std::stoi(argv[3]) : 40;
    const std::string out = argc > 4 ? argv[4] : (argc == 3 ? argv[2] : "");

    //----------Generate design region (sample_optimize_density_oc.cpp:24-41)----------
    SquareMesh<double> mesh(nx, ny, nx, ny);
    std::vector<Vector<double> > x = mesh.GenerateNodes();
    std::vector<std::vector<int> > elements = mesh.GenerateElements();
    auto ufixed = mesh.GenerateFixedlist({ 0, 1 }, [](Vector<double> _x) { return std::fabs(_x(0)) < 1.0e-5; });
    auto qfixed = mesh.GenerateFixedlist({ 1 }, [&](Vector<double> _x) { return std::fabs(_x(0) - nx) < 1.0e-5 && std::fabs(_x(1) - 0.5*ny) < 1.0e-5; });
    for (auto& q : qfixed) q.second = -1.0;

    //----------Neighbour lists, R = 1.5 (structured equivalent of the all-pairs search, :50-60)----------
    const double R = 1.5;
    std::vector<std::vector<int> > neighbors(elements.size());
    std::vector<std::vector<double> > w(elements.size());
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
        const int ii = i + di, jj = j + dj;
        if (ii < 0 || ii >= nx || jj < 0 || jj >= ny) continue;
        const double d = std::sqrt((double)(di*di + dj*dj));
        if (d <= R) { neighbors[ny*i + j].push_back(ny*ii + jj); w[ny*i + j].push_back((R - d)/R); }
    }
    HeavisideFilter<double> filter(elements.size(), neighbors, w);

    //----------Device model + design loop----------
    typedef B200::PlaneStrainStiffnessTag<ShapeFunction4Square, Gauss4Square> Equation;
    B200::Model model(x, elements, 2, ufixed);
    B200::SimpParameters prm;
    std::vector<double> optp = optimizer == "oc" ? std::vector<double>{ 0.5, 0.0, 1.0e4, 1.0e-3, 0.15 }
                             : optimizer == "conlin" ? std::vector<double>{ 0.2, 1.0e-6, 1.0, 0.0, 10000.0, 0.0, 0.01, 1.0 }
                                                 : std::vector<double>{ 1.0e-5, 0.1, 0.2, 0.5, 0.7, 1.2, 1.0e-6, 1.0, 0.0, 10000.0, 0.0, 0.01, 1.0 };
    const int optkind = optimizer == "oc" ? PF2_OPT_OC : (optimizer == "conlin" ? PF2_OPT_CONLIN : PF2_OPT_MMA);
    B200::DesignLoop<Equation> loop(model, filter, optkind, optp, prm, qfixed, std::vector<double>(elements.size(), 0.5));
    if (std::getenv("PF2_SAMPLE_WARM_START")) {         //  opt-in: every solve starts from the previous displacements (the reference starts from 0)
        loop.Reset(std::vector<double>(elements.size(), 0.5));
        loop.SetWarmStart(true);
    }

    int k = 0;
    for (; k < 500; k++) {
        B200::IterationReport it = loop.Iterate();
        std::cout << "k = " << k << "\tObjective:\t" << it.f/prm.scale0 << "\tWeight:\t" << it.g/prm.scale1 << "\tCG:\t" << it.cg_iterations << std::endl;
        if (it.converged) { std::cout << "--------------------Optimized--------------------" << std::endl; break; }
    }

    if (!out.empty()) {
        std::vector<double> s, rho;
        std::vector<Vector<double> > u, r;
        loop.Get(s, rho, u, r);
        std::ofstream fout(out);
        MakeHeadderToVTK(fout);
        AddPointsToVTK(x, fout);
        AddElementToVTK(elements, fout);
        AddElementTypes(std::vector<int>(elements.size(), 9), fout);
        AddPointVectors(u, "u", fout, true);
        AddPointVectors(r, "r", fout, false);
        AddElementScalers(rho, "s", fout, true);
        fout.close();
        //  the same dump written by the library from the device-resident fields (pf2_simp_export_vtk).  Without the reactions it is
        //  byte-identical to the host writers' file; with them only up to the summation order of the reaction pass (reactions at free
        //  nodes are round-off noise, and two passes add the element contributions in different orders).
        std::ofstream fnor(out + ".nor.vtk");
        MakeHeadderToVTK(fnor);
        AddPointsToVTK(x, fnor);
        AddElementToVTK(elements, fnor);
        AddElementTypes(std::vector<int>(elements.size(), 9), fnor);
        AddPointVectors(u, "u", fnor, true);
        AddElementScalers(rho, "s", fnor, true);
        fnor.close();
        loop.ExportVTK(out + ".device.vtk", 9, false);
        loop.ExportVTK(out + ".device_r.vtk", 9, true);
    }
    return 0;
}
