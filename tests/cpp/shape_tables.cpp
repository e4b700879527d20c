//  Prints N(r), dNdr(r) of every shape-function class at a few points and the Points / Weights of every integration rule.
//  Compiled once against the reference's src/ (tests/golden/make_golden.py -> tests/golden/shape_tables.txt) and once against
//  the header mirror pansfem2_b200/src (tests/test_mirror_host.py); the two outputs must agree.
#include <iostream>
#include <iomanip>
#include <vector>
#include <cmath>
#include <cassert>
#include <numeric>
#include <algorithm>
#include "LinearAlgebra/Models/Vector.h"
#include "LinearAlgebra/Models/Matrix.h"
#include "FEM/Controller/ShapeFunction.h"
#include "FEM/Controller/GaussIntegration.h"

using namespace PANSFEM2;

template<template<class>class SF>
void shape(const char* name, const std::vector<Vector<double> >& pts) {
    for (auto r : pts) {
        Vector<double> N = SF<double>::N(r);
        Matrix<double> d = SF<double>::dNdr(r);
        std::cout << name << " N";
        for (int i = 0; i < SF<double>::n; i++) std::cout << " " << N(i);
        std::cout << "\n" << name << " dNdr";
        for (int k = 0; k < SF<double>::d; k++) for (int i = 0; i < SF<double>::n; i++) std::cout << " " << d(k, i);
        std::cout << "\n";
    }
}

template<template<class>class IC>
void rule(const char* name, int dim) {
    for (int g = 0; g < IC<double>::N; g++) {
        std::cout << name << " " << g;
        Vector<double> pt = IC<double>::Points[g];          //  the reference's Vector::operator() is not const
        for (int k = 0; k < dim; k++) std::cout << " " << pt(k);
        for (int k = 0; k < dim; k++) std::cout << " " << IC<double>::Weights[g][k];
        std::cout << "\n";
    }
}

int main() {
    std::cout << std::setprecision(17);
    std::vector<Vector<double> > p1 = { { 0.3 }, { -0.7 } };
    std::vector<Vector<double> > p2 = { { 0.21, 0.37 }, { -0.4, 0.8 }, { 0.0, 0.0 } };
    std::vector<Vector<double> > p3 = { { 0.21, 0.37, -0.55 }, { -0.4, 0.8, 0.1 }, { 0.0, 0.0, 0.0 } };
    shape<ShapeFunction2Line>("2Line", p1);
    shape<ShapeFunction3Triangle>("3Triangle", p2);
    shape<ShapeFunction6Triangle>("6Triangle", p2);
    shape<ShapeFunction4Square>("4Square", p2);
    shape<ShapeFunction8Square>("8Square", p2);
    shape<ShapeFunction4Tetrahedron>("4Tetrahedron", p3);
    shape<ShapeFunction8Cubic>("8Cubic", p3);
    shape<ShapeFunction20Cubic>("20Cubic", p3);
    rule<Gauss1Line>("Gauss1Line", 1);
    rule<Gauss2Line>("Gauss2Line", 1);
    rule<Gauss1Triangle>("Gauss1Triangle", 2);
    rule<Gauss3Triangle>("Gauss3Triangle", 2);
    rule<Gauss1Square>("Gauss1Square", 2);
    rule<Gauss4Square>("Gauss4Square", 2);
    rule<Gauss9Square>("Gauss9Square", 2);
    rule<Gauss1Tetrahedron>("Gauss1Tetrahedron", 3);
    rule<Gauss8Cubic>("Gauss8Cubic", 3);
    rule<Gauss27Cubic>("Gauss27Cubic", 3);
    return 0;
}
