// tests/cpp/io_formats.cpp -- the data formats either side of the path (VTK out, CSV / VTK in), exercised through the PrePost headers.
// Compiled once against the reference's headers (tests/golden/make_golden.py io -> tests/golden/io_formats_*.txt) and, in the CPU test
// suite, against the header mirror pansfem2_b200/src: both builds must write the same bytes and parse the same values.
//     usage: io_formats <scratch directory>        (-DIO_VTK2 selects ImportFromVTK2.h, whose class has the same name)
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <cmath>

#include "LinearAlgebra/Models/Vector.h"
#include "PrePost/Export/ExportToVTK.h"
#include "PrePost/Import/ImportFromCSV.h"
#ifdef IO_VTK2
#include "PrePost/Import/ImportFromVTK2.h"
#else
#include "PrePost/Import/ImportFromVTK.h"
#endif

using namespace PANSFEM2;

static void dump(const char* name, std::vector<Vector<double> >& v) {
    std::printf("%s %zu\n", name, v.size());
    for (auto& x : v) { for (int i = 0; i < x.SIZE(); i++) std::printf(" %.17g", x(i)); std::printf("\n"); }
}
static void dump(const char* name, const std::vector<double>& v) {
    std::printf("%s %zu\n", name, v.size());
    for (double x : v) std::printf(" %.17g\n", x);
}
static void dump(const char* name, const std::vector<std::vector<int> >& v) {
    std::printf("%s %zu\n", name, v.size());
    for (const auto& e : v) { for (int n : e) std::printf(" %d", n); std::printf("\n"); }
}
static void dump(const char* name, const std::vector<std::pair<std::pair<int, int>, double> >& v) {
    std::printf("%s %zu\n", name, v.size());
    for (const auto& b : v) std::printf(" %d %d %.17g\n", b.first.first, b.first.second, b.second);
}

int main(int argc, char** argv) {
    const std::string dir = std::string(argc > 1 ? argv[1] : ".") + "/";

    //----------a small mixed mesh with awkward numbers----------
    std::vector<Vector<double> > x = { Vector<double>({ 0.0, 0.0 }), Vector<double>({ 1.0/3.0, 1.0e-7 }), Vector<double>({ 123456.789, -2.5 }),
                                      Vector<double>({ -0.125, 1.0e+20 }), Vector<double>({ 2.0, 3.0 }), Vector<double>({ 6.02214076e23, -1.0/7.0 }) };
    std::vector<std::vector<int> > elements = { { 0, 1, 4, 3 }, { 1, 2, 4 }, { 2, 5, 4 } };
    std::vector<Vector<double> > u(x.size(), Vector<double>(2)), w3(x.size(), Vector<double>(3)), phi(x.size(), Vector<double>(1));
    std::vector<double> T(x.size()), s(elements.size());
    for (size_t i = 0; i < x.size(); i++) {
        u[i](0) = std::sin(1.0 + i)*1.0e-3; u[i](1) = -std::cos(2.0*i)*12345.678;
        w3[i](0) = 1.0*i; w3[i](1) = 0.5*i*i; w3[i](2) = -1.0/(1.0 + i);
        phi[i](0) = std::tanh(0.3*i - 1.0);
        T[i] = 300.0 - 17.25*i;
    }
    for (size_t e = 0; e < elements.size(); e++) s[e] = 0.001 + 0.4995*e;

    //----------every writer of ExportToVTK.h----------
    {
        std::ofstream fout(dir + "out.vtk");
        MakeHeadderToVTK(fout);
        AddPointsToVTK(x, fout);
        AddElementToVTK(elements, fout);
        AddElementTypes(std::vector<int>({ 9, 5, 5 }), fout);
        AddPointVectors(u, "u", fout, true);
        AddPointVectors(w3, "w", fout, false);
        AddPointScalers(T, "T", fout, false);
        AddPointScalers(phi, "phi", fout, false);
        AddElementScalers(s, "s", fout, true);
        fout.close();
    }

    //----------CSV readers: the drivers' input files----------
    {
        std::ofstream(dir + "Node.csv") << "ID,x0,x1,x2\n0,0.0,0.05,1e-3\n1,0.15000000000000002,-7,2.5E+2\n2,1,2,3\n";
        std::ofstream(dir + "Element.csv") << "Cell ID,Point Index 0,Point Index 1,Point Index 2\n0,23,3,2\n1,25,5,4\n\n";
        std::ofstream(dir + "Dirichlet.csv") << "idn,u,v,w\n0,0,free,0.5\n7,free,free,-1e-2\n9,1,2,3\n";
        std::ofstream(dir + "Neumann.csv") << "idn,fx,fy\n3,free,-100\n4,2.5,free\n";
        std::ofstream(dir + "Initial.csv") << "idn,u,v\n1,0.25,free\n2,free,-4\n";
        std::ofstream(dir + "Periodic.csv") << "master,slave\n0,10\n1,11\n5,15\n";
        std::vector<Vector<double> > nodes;
        std::vector<std::vector<int> > cells;
        std::vector<std::pair<std::pair<int, int>, double> > ufixed, qfixed;
        std::vector<std::pair<int, int> > periodic;
        std::vector<Vector<double> > init(4, Vector<double>(2));
        std::printf("ok %d %d %d %d %d %d\n", (int)ImportNodesFromCSV(nodes, dir + "Node.csv"), (int)ImportElementsFromCSV(cells, dir + "Element.csv"),
                    (int)ImportDirichletFromCSV(ufixed, dir + "Dirichlet.csv"), (int)ImportNeumannFromCSV(qfixed, dir + "Neumann.csv"),
                    (int)ImportInitialFromCSV(init, dir + "Initial.csv"), (int)ImportPeriodicFromCSV(periodic, dir + "Periodic.csv"));
        dump("nodes", nodes); dump("cells", cells); dump("dirichlet", ufixed); dump("neumann", qfixed); dump("initial", init);
        std::printf("periodic %zu\n", periodic.size());
        for (const auto& p : periodic) std::printf(" %d %d\n", p.first, p.second);
        std::vector<Vector<double> > none;
        std::printf("missing %d\n", (int)ImportNodesFromCSV(none, dir + "no_such_file.csv"));
    }

    //----------VTK readers on the file written above----------
#ifdef IO_VTK2
    {
        ImportModelFromVTK<double> model(dir + "out.vtk");
        std::vector<Vector<double> > nodes = model.GenerateNodes();
        std::vector<std::vector<int> > cells = model.GenerateElements();
        dump("vtk2 nodes", nodes); dump("vtk2 cells", cells);
        nodes = model.GenerateNodes();
        dump("vtk2 nodes again", nodes);
    }
#else
    {
        ImportModelFromVTK<double> model(dir + "out.vtk", 2);
        std::vector<Vector<double> > nodes = model.ImportPOINTS();
        std::vector<std::vector<int> > cells = model.ImportCELLS();
        std::vector<Vector<double> > uu = model.ImportPOINTVECTORS("u");
        std::vector<double> TT = model.ImportPOINTSCALARS("T");
        std::vector<double> ss = model.ImportCELLSCALARS("s");
        dump("vtk nodes", nodes); dump("vtk cells", cells); dump("vtk u", uu); dump("vtk T", TT); dump("vtk s", ss);
        //  forward-scanner behaviour: "w" lies before the current position -> rewind + empty; the next call then finds it from the top;
        //  a cell-data name asked of the point section is never found
        std::vector<Vector<double> > w1 = model.ImportPOINTVECTORS("w"), w2 = model.ImportPOINTVECTORS("w");
        std::vector<double> nope = model.ImportPOINTSCALARS("s"), phi1 = model.ImportPOINTSCALARS("phi"), late = model.ImportCELLSCALARS("s");
        std::vector<Vector<double> > cv = model.ImportCELLVECTORS("u");
        dump("vtk w first", w1); dump("vtk w second", w2); dump("vtk s as point scalars", nope); dump("vtk phi", phi1); dump("vtk s again", late);
        dump("vtk u as cell vectors", cv);
        ImportModelFromVTK<double> model3(dir + "out.vtk", 3);
        std::vector<Vector<double> > n3 = model3.ImportPOINTS(), w = model3.ImportPOINTVECTORS("w");
        dump("vtk nodes 3d", n3); dump("vtk w 3d", w);
    }
#endif
    return 0;
}
