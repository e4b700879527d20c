//  pansfem2_b200/src/FEM/Controller/GaussIntegration.h
//  Quadrature policy classes mirroring src/FEM/Controller/GaussIntegration.h: Gauss1Line (:18-36), Gauss2Line (:40-60),
//  Gauss1Triangle (:64-82), Gauss3Triangle (:86-107), Gauss1Square (:112-130), Gauss4Square (:134-157), Gauss9Square
//  (:162-195), Gauss1Tetrahedron (:200-218), Gauss8Cubic (:222-253), Gauss27Cubic (:258-327): static N, Points, Weights
//  (per-axis weights, multiplied by the caller).  Point order is the reference's: Gauss4Square (-,-),(+,-),(-,+),(+,+);
//  Gauss8Cubic bottom face CCW then top face CCW; the 9 / 27 point rules run r0 fastest.
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

    namespace B200 {
        //  tensor-product 3-point rule: index (i, j, k) with r0 fastest
        template<class T>
        inline std::vector<Vector<T> > Tensor3Points(int _dim) {
            const T p[3] = { -sqrt(3.0/5.0), 0.0, sqrt(3.0/5.0) };
            std::vector<Vector<T> > pts;
            const int n = (_dim == 2) ? 9 : 27;
            for (int g = 0; g < n; g++) {
                if (_dim == 2) pts.push_back(Vector<T>({ p[g%3], p[g/3] }));
                else pts.push_back(Vector<T>({ p[g%3], p[(g/3)%3], p[g/9] }));
            }
            return pts;
        }
        template<class T>
        inline std::vector<std::vector<T> > Tensor3Weights(int _dim) {
            const T w[3] = { 5.0/9.0, 8.0/9.0, 5.0/9.0 };
            std::vector<std::vector<T> > ws;
            const int n = (_dim == 2) ? 9 : 27;
            for (int g = 0; g < n; g++) {
                if (_dim == 2) ws.push_back({ w[g%3], w[g/3] });
                else ws.push_back({ w[g%3], w[(g/3)%3], w[g/9] });
            }
            return ws;
        }
    }

#define PF2_GAUSS_CLASS(NAME, COUNT)                                        \
    template<class T>                                                       \
    class NAME {                                                            \
public:                                                                     \
        static const int N = COUNT;                                         \
        static const std::vector<Vector<T> > Points;                        \
        static const std::vector<std::vector<T> > Weights;                  \
    };

    PF2_GAUSS_CLASS(Gauss1Line, 1)
    template<class T> const std::vector<Vector<T> > Gauss1Line<T>::Points = { { T() } };
    template<class T> const std::vector<std::vector<T> > Gauss1Line<T>::Weights = { { 2.0 } };

    PF2_GAUSS_CLASS(Gauss2Line, 2)
    template<class T> const std::vector<Vector<T> > Gauss2Line<T>::Points = { { -1.0/sqrt(3.0) }, { 1.0/sqrt(3.0) } };
    template<class T> const std::vector<std::vector<T> > Gauss2Line<T>::Weights = { { 1.0 }, { 1.0 } };

    PF2_GAUSS_CLASS(Gauss1Triangle, 1)
    template<class T> const std::vector<Vector<T> > Gauss1Triangle<T>::Points = { { 1.0/3.0, 1.0/3.0 } };
    template<class T> const std::vector<std::vector<T> > Gauss1Triangle<T>::Weights = { { 1.0/sqrt(2.0), 1.0/sqrt(2.0) } };

    PF2_GAUSS_CLASS(Gauss3Triangle, 3)
    template<class T> const std::vector<Vector<T> > Gauss3Triangle<T>::Points = { { 1.0/6.0, 1.0/6.0 }, { 2.0/3.0, 1.0/6.0 }, { 1.0/6.0, 2.0/3.0 } };
    template<class T> const std::vector<std::vector<T> > Gauss3Triangle<T>::Weights = std::vector<std::vector<T> >(3, std::vector<T>(2, 1.0/sqrt(6.0)));

    PF2_GAUSS_CLASS(Gauss1Square, 1)
    template<class T> const std::vector<Vector<T> > Gauss1Square<T>::Points = { { T(), T() } };
    template<class T> const std::vector<std::vector<T> > Gauss1Square<T>::Weights = { { 2.0, 2.0 } };

    PF2_GAUSS_CLASS(Gauss9Square, 9)
    template<class T> const std::vector<Vector<T> > Gauss9Square<T>::Points = B200::Tensor3Points<T>(2);
    template<class T> const std::vector<std::vector<T> > Gauss9Square<T>::Weights = B200::Tensor3Weights<T>(2);

    PF2_GAUSS_CLASS(Gauss1Tetrahedron, 1)
    template<class T> const std::vector<Vector<T> > Gauss1Tetrahedron<T>::Points = { { 1.0/4.0, 1.0/4.0, 1.0/4.0 } };
    template<class T> const std::vector<std::vector<T> > Gauss1Tetrahedron<T>::Weights = { { 1.0/cbrt(6.0), 1.0/cbrt(6.0), 1.0/cbrt(6.0) } };

    PF2_GAUSS_CLASS(Gauss27Cubic, 27)
    template<class T> const std::vector<Vector<T> > Gauss27Cubic<T>::Points = B200::Tensor3Points<T>(3);
    template<class T> const std::vector<std::vector<T> > Gauss27Cubic<T>::Weights = B200::Tensor3Weights<T>(3);
#undef PF2_GAUSS_CLASS
}
