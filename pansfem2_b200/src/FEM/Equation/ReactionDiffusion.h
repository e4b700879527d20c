//  pansfem2_b200/src/FEM/Equation/ReactionDiffusion.h
//  ReactionDiffusionConsistentMass (src/FEM/Equation/ReactionDiffusion.h:20-21), ReactionDiffusionLumpedMass (:51-52),
//  ReactionDiffusionStiffness (:81-82) and ReactionDiffusionReaction (:110-111) with the reference's signatures.
//  The two matrices run on the B200 (PF2_PHYS_MASS, and PF2_PHYS_HEAT with alpha = D, t = 1); the lumped mass is a scaled identity and
//  the reaction vector takes a C++ functor, so both stay host loops.  The batched level-set loop (B200::LevelSetLoop) needs none
//  of them per element: it assembles T = Me/dt + Ke once and the right-hand side on the device.
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void ReactionDiffusionConsistentMass(Matrix<T>& _Me, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x) {
        assert(_doulist.size() == 1);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_MASS, SF, IC>::value, 1, _Me, _nodetoelement, _element, _doulist, _x, T(1), T(0), T(1));
    }

    template<class T, template<class>class SF, template<class>class IC>
    void ReactionDiffusionStiffness(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _D) {
        assert(_doulist.size() == 1);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_HEAT, SF, IC>::value, 1, _Ke, _nodetoelement, _element, _doulist, _x, _D, T(0), T(1));
    }

    namespace B200 {
        //  geometry of one integration point of a 2-D element: N, dNdX, J
        template<class T, template<class>class SF, template<class>class IC>
        inline T PointGeometry(int _g, const std::vector<int>& _element, std::vector<Vector<T> >& _x, Vector<T>& _N, Matrix<T>& _dNdX) {
            const int n = (int)_element.size();
            Matrix<T> X(n, 2);
            for (int i = 0; i < n; i++) { X(i, 0) = _x[_element[i]](0); X(i, 1) = _x[_element[i]](1); }
            _N = SF<T>::N(IC<T>::Points[_g]);
            Matrix<T> dNdr = SF<T>::dNdr(IC<T>::Points[_g]);
            Matrix<T> dXdr = dNdr*X;
            _dNdX = dXdr.Inverse()*dNdr;
            return dXdr.Determinant();
        }
    }

    template<class T, template<class>class SF, template<class>class IC>
    void ReactionDiffusionLumpedMass(Matrix<T>& _Me, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x) {
        assert(_doulist.size() == 1);
        const int n = (int)_element.size();
        _nodetoelement = std::vector<std::vector<std::pair<int, int> > >(n, std::vector<std::pair<int, int> >(1));
        for (int i = 0; i < n; i++) _nodetoelement[i][0] = std::make_pair(_doulist[0], i);
        T Area = T();
        for (int g = 0; g < IC<T>::N; g++) {
            Vector<T> N; Matrix<T> dNdX;
            Area += B200::PointGeometry<T, SF, IC>(g, _element, _x, N, dNdX)*IC<T>::Weights[g][0]*IC<T>::Weights[g][1];
        }
        _Me = Identity<T>(n);
        _Me *= Area/(T)n;
    }

    template<class T, template<class>class SF, template<class>class IC, class F>
    void ReactionDiffusionReaction(Vector<T>& _Fe, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, std::vector<Vector<T> >& _u, F _f) {
        assert(_doulist.size() == 1);
        const int n = (int)_element.size();
        _Fe = Vector<T>(n);
        _nodetoelement = std::vector<std::vector<std::pair<int, int> > >(n, std::vector<std::pair<int, int> >(1));
        for (int i = 0; i < n; i++) _nodetoelement[i][0] = std::make_pair(_doulist[0], i);
        Vector<T> U(n);
        for (int i = 0; i < n; i++) U(i) = _u[_element[i]](0);
        for (int g = 0; g < IC<T>::N; g++) {
            Vector<T> N; Matrix<T> dNdX;
            const T J = B200::PointGeometry<T, SF, IC>(g, _element, _x, N, dNdX);
            const T u = N*U;
            Vector<T> dudX = dNdX*U;
            _Fe += N*_f(u, dudX)*J*IC<T>::Weights[g][0]*IC<T>::Weights[g][1];
        }
    }
}
