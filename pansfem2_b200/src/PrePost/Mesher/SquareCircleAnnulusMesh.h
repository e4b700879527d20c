//  pansfem2_b200/src/PrePost/Mesher/SquareCircleAnnulusMesh.h
//  SquareCircleAnnulusMesh<T>(a, b, r, p, nx, ny, nr) of src/PrePost/Mesher/SquareCircleAnnulusMesh.h:20-146: a rectangle a x b with a
//  circular hole of radius r, both centred at the origin; layer i blends the circle (weight 1 - t) and the rectangle's boundary
//  (weight t) with t = (i / nr)^p, the loop running like SquareAnnulusMesh's.  Topology and queries: B200/RingMesh.h.
#pragma once
#include <cmath>
#include "../../B200/RingMesh.h"

namespace PANSFEM2 {
    template<class T>
    class SquareCircleAnnulusMesh : public B200::RingMesh<T, SquareCircleAnnulusMesh<T> > {
public:
        SquareCircleAnnulusMesh(T _a, T _b, T _r, T _p, int _nx, int _ny, int _nr)
            : B200::RingMesh<T, SquareCircleAnnulusMesh<T> >(2*(_nx + _ny), _nr), a(_a), b(_b), r(_r), p(_p), nx(_nx), ny(_ny) {}
        ~SquareCircleAnnulusMesh() {}
        Vector<T> Position(int _layer, int _position) {
            const T t = pow(_layer/(T)this->layers, p);
            const int n = this->around;
            if (_position < ny) {
                const int j = _position; const T theta = 2*M_PI*(j - 0.5*ny)/(T)n;
                return Vector<T>({ (1 - t)*r*cos(theta) + t*0.5*a, (1 - t)*r*sin(theta) + t*b*(j/(T)ny - 0.5) });
            }
            if (_position < ny + nx) {
                const int j = _position - ny; const T theta = 2*M_PI*(j + 0.5*ny)/(T)n;
                return Vector<T>({ (1 - t)*r*cos(theta) + t*a*(0.5 - j/(T)nx), (1 - t)*r*sin(theta) + t*0.5*b });
            }
            if (_position < 2*ny + nx) {
                const int j = _position - ny - nx; const T theta = 2*M_PI*(j + 0.5*ny + nx)/(T)n;
                return Vector<T>({ (1 - t)*r*cos(theta) - t*0.5*a, (1 - t)*r*sin(theta) + t*b*(0.5 - j/(T)ny) });
            }
            const int j = _position - 2*ny - nx; const T theta = 2*M_PI*(j + 1.5*ny + nx)/(T)n;
            return Vector<T>({ (1 - t)*r*cos(theta) + t*a*(j/(T)nx - 0.5), (1 - t)*r*sin(theta) - t*0.5*b });
        }
private:
        T a, b, r, p;
        int nx, ny;
    };
}
