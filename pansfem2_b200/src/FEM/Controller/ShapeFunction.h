//  pansfem2_b200/src/FEM/Controller/ShapeFunction.h
//  Shape-function policy classes used as template arguments of the element routines, mirroring
//  src/FEM/Controller/ShapeFunction.h: ShapeFunction2Line (:20-48), 3Line (:52-82), 3Triangle (:87-118), 6Triangle (:122-156),
//  4Square (:160-191), 8Square (:196-247), 4Tetrahedron (:251-284), 8Cubic (:288-329), 20Cubic (:334-461):
//  static d, n, Points, N(r), dNdr(r).  On the hot path they are compile-time TAGS that select a CUDA kernel
//  instantiation (B200/ElementSelect.h); N / dNdr stay callable for user code and the host-side load vectors.
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

    //********************2NodesLine (edges of the surface-force routines)********************
    template<class T>
    class ShapeFunction2Line {
public:
        static const int d = 1;
        static const int n = 2;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) { Vector<T> v(n); v(0) = 0.5*(1 - _r(0)); v(1) = 0.5*(1 + _r(0)); return v; }
        static Matrix<T> dNdr(Vector<T> _r) { Matrix<T> m(d, n); m(0, 0) = -0.5; m(0, 1) = 0.5; return m; }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction2Line<T>::Points = { { -1.0 }, { 1.0 } };

    //********************3NodesLine: end nodes first, then the mid node********************
    template<class T>
    class ShapeFunction3Line {
public:
        static const int d = 1;
        static const int n = 3;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) { const T r = _r(0); Vector<T> v(n); v(0) = -0.5*(1.0 - r)*r; v(1) = 0.5*r*(1.0 + r); v(2) = (1.0 - r)*(1.0 + r); return v; }
        static Matrix<T> dNdr(Vector<T> _r) { const T r = _r(0); Matrix<T> m(d, n); m(0, 0) = -0.5*(1.0 - 2.0*r); m(0, 1) = 0.5*(1.0 + 2.0*r); m(0, 2) = -2.0*r; return m; }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction3Line<T>::Points = { { -1.0 }, { 1.0 }, { T() } };

    //********************3NodesTriangle: area coordinates (r0, r1, 1 - r0 - r1)********************
    template<class T>
    class ShapeFunction3Triangle {
public:
        static const int d = 2;
        static const int n = 3;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) { Vector<T> v(n); v(0) = _r(0); v(1) = _r(1); v(2) = 1.0 - _r(0) - _r(1); return v; }
        static Matrix<T> dNdr(Vector<T> _r) {
            Matrix<T> m(d, n);
            m(0, 0) = 1.0; m(0, 1) = 0.0; m(0, 2) = -1.0;
            m(1, 0) = 0.0; m(1, 1) = 1.0; m(1, 2) = -1.0;
            return m;
        }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction3Triangle<T>::Points = { { 1.0, T() }, { T(), 1.0 }, { T(), T() } };

    //********************6NodesTriangle********************
    template<class T>
    class ShapeFunction6Triangle {
public:
        static const int d = 2;
        static const int n = 6;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) {
            const T a = _r(0), b = _r(1), c = 1.0 - _r(0) - _r(1);
            Vector<T> v(n);
            v(0) = a*(2.0*a - 1.0); v(1) = b*(2.0*b - 1.0); v(2) = c*(1.0 - 2.0*a - 2.0*b);
            v(3) = 4.0*a*b; v(4) = 4.0*b*c; v(5) = 4.0*c*a;
            return v;
        }
        static Matrix<T> dNdr(Vector<T> _r) {
            const T a = _r(0), b = _r(1);
            Matrix<T> m(d, n);
            m(0, 0) = 4.0*a - 1.0; m(0, 1) = 0.0; m(0, 2) = -3.0 + 4.0*a + 4.0*b; m(0, 3) = 4.0*b; m(0, 4) = -4.0*b; m(0, 5) = 4.0*(1.0 - 2.0*a - b);
            m(1, 0) = 0.0; m(1, 1) = 4.0*b - 1.0; m(1, 2) = -3.0 + 4.0*a + 4.0*b; m(1, 3) = 4.0*a; m(1, 4) = 4.0*(1.0 - a - 2.0*b); m(1, 5) = -4.0*a;
            return m;
        }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction6Triangle<T>::Points = { { 1.0, T() }, { T(), 1.0 }, { T(), T() }, { 0.5, 0.5 }, { T(), 0.5 }, { 0.5, T() } };

    //********************8NodesSquare (serendipity): corners then mid-sides (0,-1) (1,0) (0,1) (-1,0)********************
    template<class T>
    class ShapeFunction8Square {
public:
        static const int d = 2;
        static const int n = 8;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) {
            Vector<T> v(n);
            for (int i = 0; i < 4; i++) {
                const T sx = Points[i](0), sy = Points[i](1);
                v(i) = 0.25*(1.0 + sx*_r(0))*(1.0 + sy*_r(1))*(sx*_r(0) + sy*_r(1) - 1.0);
            }
            v(4) = 0.5*(1.0 - _r(0)*_r(0))*(1.0 - _r(1)); v(5) = 0.5*(1.0 + _r(0))*(1.0 - _r(1)*_r(1));
            v(6) = 0.5*(1.0 - _r(0)*_r(0))*(1.0 + _r(1)); v(7) = 0.5*(1.0 - _r(0))*(1.0 - _r(1)*_r(1));
            return v;
        }
        static Matrix<T> dNdr(Vector<T> _r) {
            Matrix<T> m(d, n);
            for (int i = 0; i < 4; i++) {
                const T sx = Points[i](0), sy = Points[i](1);
                m(0, i) = 0.25*sx*(1.0 + sy*_r(1))*(2.0*sx*_r(0) + sy*_r(1));
                m(1, i) = 0.25*sy*(1.0 + sx*_r(0))*(sx*_r(0) + 2.0*sy*_r(1));
            }
            m(0, 4) = -_r(0)*(1.0 - _r(1));           m(1, 4) = -0.5*(1.0 - _r(0)*_r(0));
            m(0, 5) = 0.5*(1.0 - _r(1)*_r(1));        m(1, 5) = -_r(1)*(1.0 + _r(0));
            m(0, 6) = -_r(0)*(1.0 + _r(1));           m(1, 6) = 0.5*(1.0 - _r(0)*_r(0));
            m(0, 7) = -0.5*(1.0 - _r(1)*_r(1));       m(1, 7) = -_r(1)*(1.0 - _r(0));
            return m;
        }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction8Square<T>::Points = { { -1.0, -1.0 }, { 1.0, -1.0 }, { 1.0, 1.0 }, { -1.0, 1.0 }, { T(), -1.0 }, { 1.0, T() }, { T(), 1.0 }, { -1.0, T() } };

    //********************4NodesTetrahedron: volume coordinates (r0, r1, r2, 1 - r0 - r1 - r2)********************
    template<class T>
    class ShapeFunction4Tetrahedron {
public:
        static const int d = 3;
        static const int n = 4;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) { Vector<T> v(n); v(0) = _r(0); v(1) = _r(1); v(2) = _r(2); v(3) = 1.0 - _r(0) - _r(1) - _r(2); return v; }
        static Matrix<T> dNdr(Vector<T> _r) {
            Matrix<T> m(d, n);
            for (int k = 0; k < d; k++) for (int i = 0; i < n; i++) m(k, i) = (i == 3) ? -1.0 : (i == k ? 1.0 : 0.0);
            return m;
        }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction4Tetrahedron<T>::Points = { { 1.0, T(), T() }, { T(), 1.0, T() }, { T(), T(), 1.0 }, { T(), T(), T() } };

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

    //********************20NodesCubic (serendipity)********************
    //  corners as 8Cubic, then mid-edges: 8..11 bottom face, 12..15 top face, 16..19 vertical edges - the nodes
    //  ShapeFunction20Cubic::N / dNdr of the reference interpolate (ShapeFunction.h:368-461).
    template<class T>
    class ShapeFunction20Cubic {
public:
        static const int d = 3;
        static const int n = 20;
        static const std::vector<Vector<T> > Points;
        static Vector<T> N(Vector<T> _r) {
            Vector<T> v(n);
            for (int i = 0; i < n; i++) {
                const T sx = Points[i](0), sy = Points[i](1), sz = Points[i](2);
                const T a = 1.0 + sx*_r(0), b = 1.0 + sy*_r(1), c = 1.0 + sz*_r(2);
                if (i < 8) v(i) = 0.125*a*b*c*(sx*_r(0) + sy*_r(1) + sz*_r(2) - 2.0);
                else if (sx == T()) v(i) = 0.25*(1.0 - _r(0)*_r(0))*b*c;
                else if (sy == T()) v(i) = 0.25*a*(1.0 - _r(1)*_r(1))*c;
                else v(i) = 0.25*a*b*(1.0 - _r(2)*_r(2));
            }
            return v;
        }
        static Matrix<T> dNdr(Vector<T> _r) {
            Matrix<T> m(d, n);
            for (int i = 0; i < n; i++) {
                const T sx = Points[i](0), sy = Points[i](1), sz = Points[i](2);
                const T a = 1.0 + sx*_r(0), b = 1.0 + sy*_r(1), c = 1.0 + sz*_r(2);
                if (i < 8) {
                    m(0, i) = 0.125*sx*b*c*(2.0*sx*_r(0) + sy*_r(1) + sz*_r(2) - 1.0);
                    m(1, i) = 0.125*sy*a*c*(sx*_r(0) + 2.0*sy*_r(1) + sz*_r(2) - 1.0);
                    m(2, i) = 0.125*sz*a*b*(sx*_r(0) + sy*_r(1) + 2.0*sz*_r(2) - 1.0);
                } else if (sx == T()) {
                    m(0, i) = -0.5*_r(0)*b*c; m(1, i) = 0.25*sy*(1.0 - _r(0)*_r(0))*c; m(2, i) = 0.25*sz*(1.0 - _r(0)*_r(0))*b;
                } else if (sy == T()) {
                    m(0, i) = 0.25*sx*(1.0 - _r(1)*_r(1))*c; m(1, i) = -0.5*_r(1)*a*c; m(2, i) = 0.25*sz*a*(1.0 - _r(1)*_r(1));
                } else {
                    m(0, i) = 0.25*sx*b*(1.0 - _r(2)*_r(2)); m(1, i) = 0.25*sy*a*(1.0 - _r(2)*_r(2)); m(2, i) = -0.5*_r(2)*a*b;
                }
            }
            return m;
        }
    };
    template<class T>
    const std::vector<Vector<T> > ShapeFunction20Cubic<T>::Points = { { -1.0, -1.0, -1.0 }, { 1.0, -1.0, -1.0 }, { 1.0, 1.0, -1.0 }, { -1.0, 1.0, -1.0 },
                                                                     { -1.0, -1.0, 1.0 }, { 1.0, -1.0, 1.0 }, { 1.0, 1.0, 1.0 }, { -1.0, 1.0, 1.0 },
                                                                     { T(), -1.0, -1.0 }, { 1.0, T(), -1.0 }, { T(), 1.0, -1.0 }, { -1.0, T(), -1.0 },
                                                                     { T(), -1.0, 1.0 }, { 1.0, T(), 1.0 }, { T(), 1.0, 1.0 }, { -1.0, T(), 1.0 },
                                                                     { -1.0, -1.0, T() }, { 1.0, -1.0, T() }, { 1.0, 1.0, T() }, { -1.0, 1.0, T() } };
}
