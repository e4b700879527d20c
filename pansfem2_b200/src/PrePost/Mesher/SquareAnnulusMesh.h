//  pansfem2_b200/src/PrePost/Mesher/SquareAnnulusMesh.h
//  SquareAnnulusMesh<T>(a, b, c, d, nx, ny, nt) of src/PrePost/Mesher/SquareAnnulusMesh.h:20-137: the frame between the rectangles
//  c x d (inner, layer 0) and a x b (outer, layer nt), both centred at the origin; a loop runs up the right side (ny nodes), along the
//  top to the left (nx), down the left side (ny) and back along the bottom (nx).  Topology and queries: B200/RingMesh.h.
//  Reference quirk kept: GenerateFixedlist evaluates its predicate with the two rectangles' roles SWAPPED (:101-131 interpolate
//  (1 - t) a + t c where GenerateNodes uses t a + (1 - t) c), so for a != c or b != d the tested coordinates are those of the
//  mirrored layer.  SquareAnnulusMesh2 (:140-358, a rectangle with a rectangular block of cells removed) is not mirrored.
#pragma once
#include "../../B200/RingMesh.h"

namespace PANSFEM2 {
    template<class T>
    class SquareAnnulusMesh : public B200::RingMesh<T, SquareAnnulusMesh<T> > {
public:
        SquareAnnulusMesh(T _a, T _b, T _c, T _d, int _nx, int _ny, int _nt)
            : B200::RingMesh<T, SquareAnnulusMesh<T> >(2*(_nx + _ny), _nt), a(_a), b(_b), c(_c), d(_d), nx(_nx), ny(_ny) {}
        ~SquareAnnulusMesh() {}
        Vector<T> Position(int _layer, int _position) { const T t = _layer/(T)this->layers; return OnLoop(_position, t, (1 - t)); }
        Vector<T> FixedPosition(int _layer, int _position) { const T t = _layer/(T)this->layers; return OnLoop(_position, (1 - t), t); }
private:
        //  point of the loop that blends the outer rectangle with weight wo and the inner one with weight wi
        Vector<T> OnLoop(int _p, T wo, T wi) const {
            if (_p < ny) { const T s = _p/(T)ny - 0.5; return Vector<T>({ wo*0.5*a + wi*0.5*c, wo*b*s + wi*d*s }); }
            if (_p < ny + nx) { const T s = 0.5 - (_p - ny)/(T)nx; return Vector<T>({ wo*a*s + wi*c*s, wo*0.5*b + wi*0.5*d }); }
            if (_p < 2*ny + nx) { const T s = 0.5 - (_p - ny - nx)/(T)ny; return Vector<T>({ -wo*0.5*a - wi*0.5*c, wo*b*s + wi*d*s }); }
            const T s = (_p - 2*ny - nx)/(T)nx - 0.5;
            return Vector<T>({ wo*a*s + wi*c*s, -wo*0.5*b - wi*0.5*d });
        }
        T a, b, c, d;
        int nx, ny;
    };
}
