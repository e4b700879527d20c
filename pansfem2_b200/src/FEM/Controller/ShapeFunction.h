//  pansfem2_b200/src/FEM/Controller/ShapeFunction.h
//  Shape-function policy classes used as template arguments of the element routines, mirroring
//  src/FEM/Controller/ShapeFunction.h:160-191 (ShapeFunction4Square) and :288-329 (ShapeFunction8Cubic):
//  static d, n, Points, N(r), dNdr(r).  On the hot path they are compile-time TAGS that select a CUDA kernel
//  instantiation; N / dNdr stay callable for user code.
#pragma once
#include <vector>
#include "../../LinearAlgebra/Models/Vector.h"
#include "../../LinearAlgebra/Models/Matrix.h"

namespace PANSFEM2 {
    //********************4NodesSquare********************
    template<class T>
    class ShapeFunction4Square {
public:
        static const int d = 2;
        static const int n = 4;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) {
            Vector<T> v(n);
            for (int i = 0; i < n; i++) v(i) = 0.25*(1.0 + Points[i](0)*_r(0))*(1.0 + Points[i](1)*_r(1));
            return v;
        }
        static Matrix<T> dNdr(Vector<T> _r) {
            Matrix<T> m(d, n);
            for (int i = 0; i < n; i++) {
                m(0, i) = 0.25*Points[i](0)*(1.0 + Points[i](1)*_r(1));
                m(1, i) = 0.25*Points[i](1)*(1.0 + Points[i](0)*_r(0));
            }
            return m;
        }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction4Square<T>::Points = { { -1.0, -1.0 }, { 1.0, -1.0 }, { 1.0, 1.0 }, { -1.0, 1.0 } };

    //********************8NodesCubic********************
    template<class T>
    class ShapeFunction8Cubic {
public:
        static const int d = 3;
        static const int n = 8;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) {
            Vector<T> v(n);
            for (int i = 0; i < n; i++) v(i) = 0.125*(1.0 + Points[i](0)*_r(0))*(1.0 + Points[i](1)*_r(1))*(1.0 + Points[i](2)*_r(2));
            return v;
        }
        static Matrix<T> dNdr(Vector<T> _r) {
            Matrix<T> m(d, n);
            for (int i = 0; i < n; i++) {
                const T a = 1.0 + Points[i](0)*_r(0), b = 1.0 + Points[i](1)*_r(1), c = 1.0 + Points[i](2)*_r(2);
                m(0, i) = 0.125*Points[i](0)*b*c;
                m(1, i) = 0.125*Points[i](1)*c*a;
                m(2, i) = 0.125*Points[i](2)*a*b;
            }
            return m;
        }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction8Cubic<T>::Points = { { -1.0, -1.0, -1.0 }, { 1.0, -1.0, -1.0 }, { 1.0, 1.0, -1.0 }, { -1.0, 1.0, -1.0 },
                                                                    { -1.0, -1.0, 1.0 }, { 1.0, -1.0, 1.0 }, { 1.0, 1.0, 1.0 }, { -1.0, 1.0, 1.0 } };
}
