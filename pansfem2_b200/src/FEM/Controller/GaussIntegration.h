//  pansfem2_b200/src/FEM/Controller/GaussIntegration.h
//  Quadrature policy classes mirroring src/FEM/Controller/GaussIntegration.h:134-157 (Gauss4Square) and :222-253
//  (Gauss8Cubic): static N, Points, Weights (per-axis weights, multiplied by the caller).  Point order is the
//  reference's: (-,-),(+,-),(-,+),(+,+) in 2-D; bottom face CCW then top face CCW in 3-D.
#pragma once
#include <vector>
#include <cmath>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    template<class T>
    class Gauss4Square {
public:
        static const int N = 4;
        static const std::vector<Vector<T> > Points;
        static const std::vector<std::vector<T> > Weights;
    };
    template<class T>
    const std::vector<Vector<T> > Gauss4Square<T>::Points = { { -1.0/sqrt(3.0), -1.0/sqrt(3.0) }, { 1.0/sqrt(3.0), -1.0/sqrt(3.0) },
                                                             { -1.0/sqrt(3.0), 1.0/sqrt(3.0) }, { 1.0/sqrt(3.0), 1.0/sqrt(3.0) } };
    template<class T>
    const std::vector<std::vector<T> > Gauss4Square<T>::Weights = std::vector<std::vector<T> >(4, std::vector<T>(2, 1.0));

    template<class T>
    class Gauss8Cubic {
public:
        static const int N = 8;
        static const std::vector<Vector<T> > Points;
        static const std::vector<std::vector<T> > Weights;
    };
    template<class T>
    const std::vector<Vector<T> > Gauss8Cubic<T>::Points = { { -1.0/sqrt(3.0), -1.0/sqrt(3.0), -1.0/sqrt(3.0) }, { 1.0/sqrt(3.0), -1.0/sqrt(3.0), -1.0/sqrt(3.0) },
                                                            { 1.0/sqrt(3.0), 1.0/sqrt(3.0), -1.0/sqrt(3.0) }, { -1.0/sqrt(3.0), 1.0/sqrt(3.0), -1.0/sqrt(3.0) },
                                                            { -1.0/sqrt(3.0), -1.0/sqrt(3.0), 1.0/sqrt(3.0) }, { 1.0/sqrt(3.0), -1.0/sqrt(3.0), 1.0/sqrt(3.0) },
                                                            { 1.0/sqrt(3.0), 1.0/sqrt(3.0), 1.0/sqrt(3.0) }, { -1.0/sqrt(3.0), 1.0/sqrt(3.0), 1.0/sqrt(3.0) } };
    template<class T>
    const std::vector<std::vector<T> > Gauss8Cubic<T>::Weights = std::vector<std::vector<T> >(8, std::vector<T>(3, 1.0));
}
