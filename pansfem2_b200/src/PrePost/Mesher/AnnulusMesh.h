//  pansfem2_b200/src/PrePost/Mesher/AnnulusMesh.h
//  AnnulusMesh<T>(r0, r1, nr, nt) of src/PrePost/Mesher/AnnulusMesh.h:20-106: nr layers of nt Q4 elements between the circles of
//  radius r0 and r1, nodes at equal angles starting on the +x axis.  Topology and queries: B200/RingMesh.h.
#pragma once
#include <cmath>
#include "../../B200/RingMesh.h"

namespace PANSFEM2 {
    template<class T>
    class AnnulusMesh : public B200::RingMesh<T, AnnulusMesh<T> > {
public:
        AnnulusMesh(T _r0, T _r1, int _nr, int _nt) : B200::RingMesh<T, AnnulusMesh<T> >(_nt, _nr), r0(_r0), r1(_r1) {}
        ~AnnulusMesh() {}
        Vector<T> Position(int _layer, int _position) {
            const T r = (r1 - r0)*_layer/(T)this->layers + r0, theta = 2.0*M_PI*_position/(T)this->around;
            return Vector<T>({ r*cos(theta), r*sin(theta) });
        }
private:
        T r0, r1;
    };
}
