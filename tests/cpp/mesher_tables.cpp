// tests/cpp/mesher_tables.cpp -- every query of SquareMesh<T> and SquareMesh2<T> (PrePost/Mesher/SquareMesh.h) on a few grids, printed
// with full precision.  Built against the reference's headers for the golden (tests/golden/make_golden.py meshers ->
// tests/golden/mesher_tables.txt) and against the header mirror in tests/test_mesher_tables.py: same numbers, same order.
#include <cstdio>
#include <cmath>
#include <vector>
#include "LinearAlgebra/Models/Vector.h"
#include "PrePost/Mesher/SquareMesh.h"
#include "PrePost/Mesher/AnnulusMesh.h"
#include "PrePost/Mesher/SquareAnnulusMesh.h"
#include "PrePost/Mesher/SquareCircleAnnulusMesh.h"

using namespace PANSFEM2;

static void nodes(const char* name, std::vector<Vector<double> > v) {
    std::printf("%s %zu\n", name, v.size());
    for (auto& p : v) std::printf(" %.17g %.17g\n", p(0), p(1));
}
static void lists(const char* name, const std::vector<std::vector<int> >& v) {
    std::printf("%s %zu\n", name, v.size());
    for (const auto& e : v) { for (int n : e) std::printf(" %d", n); std::printf("\n"); }
}
static void ids(const char* name, const std::vector<int>& v) {
    std::printf("%s %zu:", name, v.size());
    for (int n : v) std::printf(" %d", n);
    std::printf("\n");
}
static void fixed(const char* name, const std::vector<std::pair<std::pair<int, int>, double> >& v) {
    std::printf("%s %zu:", name, v.size());
    for (const auto& b : v) std::printf(" (%d,%d,%g)", b.first.first, b.first.second, b.second);
    std::printf("\n");
}

int main() {
    const int grids[3][2] = { { 4, 3 }, { 1, 1 }, { 5, 2 } };
    for (const auto& g : grids) {
        const int nx = g[0], ny = g[1];
        const double lx = 2.5*nx, ly = 0.7*ny;
        std::printf("== SquareMesh %d x %d\n", nx, ny);
        SquareMesh<double> mesh(lx, ly, nx, ny);
        nodes("nodes", mesh.GenerateNodes());
        nodes("nodes2", mesh.GenerateNodes2());
        lists("elements", mesh.GenerateElements());
        lists("elements2", mesh.GenerateElements2());
        lists("edges", mesh.GenerateEdges());
        lists("edges2", mesh.GenerateEdges2());
        ids("elements right half", mesh.GenerateElementIdsSelected([&](Vector<double> p) { return p(0) > 0.5*lx - 1.0e-9; }));
        ids("edges on top", mesh.GenerateEdgeIdsSelected([&](Vector<double> p) { return std::fabs(p(1) - ly) < 1.0e-9; }));
        ids("edges on the left or bottom", mesh.GenerateEdgeIdsSelected([&](Vector<double> p) { return p(0) < 1.0e-9 || p(1) < 1.0e-9; }));
        fixed("clamp x = 0", mesh.GenerateFixedlist({ 0, 1 }, [](Vector<double> p) { return std::fabs(p(0)) < 1.0e-9; }));
        fixed("clamp2 x = 0", mesh.GenerateFixedlist2({ 0, 1 }, [](Vector<double> p) { return std::fabs(p(0)) < 1.0e-9; }));
        fixed("roller2 upper half", mesh.GenerateFixedlist2({ 1 }, [&](Vector<double> p) { return p(1) > 0.5*ly; }));
    }
    const double ratios[2][2] = { { 1.3, 1.1 }, { 0.8, 2.0 } };
    const int grids2[3][2] = { { 6, 4 }, { 5, 3 }, { 2, 2 } };
    for (const auto& r : ratios) for (const auto& g : grids2) {
        std::printf("== SquareMesh2 %d x %d ratios %g %g\n", g[0], g[1], r[0], r[1]);
        SquareMesh2<double> mesh(3.0, 2.0, g[0], g[1], r[0], r[1]);
        nodes("nodes", mesh.GenerateNodes());
        lists("elements", mesh.GenerateElements());
        lists("edges", mesh.GenerateEdges());
        fixed("right side", mesh.GenerateFixedlist({ 1 }, [](Vector<double> p) { return std::fabs(p(0) - 3.0) < 1.0e-9; }));
        fixed("centre lines", mesh.GenerateFixedlist({ 0 }, [](Vector<double> p) { return std::fabs(p(0) - 1.5) < 1.0e-12 || std::fabs(p(1) - 1.0) < 1.0e-12; }));
    }
    //----------ring-shaped meshers----------
    {
        std::printf("== AnnulusMesh\n");
        AnnulusMesh<double> mesh(0.4, 1.7, 3, 7);
        nodes("nodes", mesh.GenerateNodes()); lists("elements", mesh.GenerateElements()); lists("edges", mesh.GenerateEdges());
        fixed("outer circle", mesh.GenerateFixedlist({ 0, 1 }, [](Vector<double> p) { return std::fabs(p.Norm() - 1.7) < 1.0e-9; }));
        fixed("upper half", mesh.GenerateFixedlist({ 1 }, [](Vector<double> p) { return p(1) > 1.0e-9; }));
    }
    {
        std::printf("== SquareAnnulusMesh\n");
        SquareAnnulusMesh<double> mesh(1.0, 1.4, 0.35, 0.6, 4, 3, 2);
        nodes("nodes", mesh.GenerateNodes()); lists("elements", mesh.GenerateElements()); lists("edges", mesh.GenerateEdges());
        fixed("outer frame by the fixed-list coordinates", mesh.GenerateFixedlist({ 0, 1 }, [](Vector<double> p) { return std::fabs(std::fabs(p(0)) - 0.5) < 1.0e-9 || std::fabs(std::fabs(p(1)) - 0.7) < 1.0e-9; }));
        fixed("right of centre", mesh.GenerateFixedlist({ 0 }, [](Vector<double> p) { return p(0) > 0.2; }));
        SquareAnnulusMesh<double> same(1.0, 1.0, 1.0 - 0.2, 1.0 - 0.6, 10, 10, 10);       // sample_optimize_homogenization.cpp:48 with a = 0.2, b = 0.6
        std::vector<Vector<double> > all = same.GenerateNodes();
        std::printf("homogenization cell %zu nodes, node 217: %.17g %.17g\n", all.size(), all[217](0), all[217](1));
    }
    for (int variant = 0; variant < 2; variant++) {
        std::printf("== SquareAnnulusMesh2 %d\n", variant);
        SquareAnnulusMesh2<double> mesh = variant == 0 ? SquareAnnulusMesh2<double>(1.5, 1.0, 6, 5, 2, 3) : SquareAnnulusMesh2<double>(1.0, 1.0, 4, 4, 2, 2);
        nodes("nodes", mesh.GenerateNodes()); lists("elements", mesh.GenerateElements()); lists("edges", mesh.GenerateEdges());
        fixed("top", mesh.GenerateFixedlist({ 1 }, [](Vector<double> p) { return std::fabs(p(1) - 1.0) < 1.0e-9; }));
        fixed("right of the hole", mesh.GenerateFixedlist({ 0, 1 }, [](Vector<double> p) { return p(0) > 0.8; }));
    }
    {
        SquareAnnulusMesh2<double> cell(1.0, 1.0, 20, 20, 10, 10);                       // sample/homogenization/sample_homogenization.cpp:31
        std::vector<Vector<double> > all = cell.GenerateNodes();
        std::vector<std::vector<int> > el = cell.GenerateElements();
        std::printf("homogenization cell 2: %zu nodes %zu elements, node 200: %.17g %.17g, element 150: %d %d %d %d\n", all.size(), el.size(), all[200](0), all[200](1), el[150][0], el[150][1], el[150][2], el[150][3]);
    }
    {
        std::printf("== SquareCircleAnnulusMesh\n");
        SquareCircleAnnulusMesh<double> mesh(2.0, 1.0, 0.3, 1.5, 3, 2, 3);
        nodes("nodes", mesh.GenerateNodes()); lists("elements", mesh.GenerateElements()); lists("edges", mesh.GenerateEdges());
        fixed("hole", mesh.GenerateFixedlist({ 0, 1 }, [](Vector<double> p) { return std::fabs(p.Norm() - 0.3) < 1.0e-9; }));
        fixed("left side", mesh.GenerateFixedlist({ 0 }, [](Vector<double> p) { return std::fabs(p(0) + 1.0) < 1.0e-9; }));
    }
    return 0;
}
