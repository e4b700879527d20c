//  pansfem2_b200/src/FEM/Equation/PlaneStrain.h
//  PlaneStrainStiffness<T, SF, IC> (src/FEM/Equation/PlaneStrain.h:20-21), PlaneStrainStiffnessSRI<T, SF, ICV, ICD> (:62-63),
//  PlaneStrainStiffnessBbar (:128-129), PlaneStrainStiffnessWilsonTaylor (:188-189), PlaneStrainMass (:385-386), PlaneStrainSurfaceForce (:420-421) and PlaneStrainBodyForce
//  (:502-503) with the reference's signatures.
//  Stiffness: any of T3 / T6 / Q4 / Q8 with a rule of its reference domain, computed on the B200.  The two load vectors take a
//  C++ functor and stay on the host (a few flops per edge / element).
#pragma once
#include <vector>
#include <cassert>
#include <cmath>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStrainStiffness(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        assert(_doulist.size() == 2);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_PLANESTRAIN, SF, IC>::value, 2, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, _t);
    }

    template<class T, template<class>class SF, template<class>class ICV, template<class>class ICD>
    void PlaneStrainStiffnessSRI(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        assert(_doulist.size() == 2);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCodeSRI<SF, ICV, ICD>::value, 2, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, _t);
    }

    //  B-bar as the reference implements it (PlaneStrain.h:128-185): volumetric part of B integrated with ICV, deviatoric part with ICD
    template<class T, template<class>class SF, template<class>class ICV, template<class>class ICD>
    void PlaneStrainStiffnessBbar(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        assert(_doulist.size() == 2);
        assert((int)_element.size() == SF<T>::n);
        const int eq = (B200::EqCodeSRI<SF, ICV, ICD>::value & ~0xff) | PF2_PHYS_PLANESTRAIN_BBAR;
        B200::ElementMatrix<T>(eq, 2, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, _t);
    }

    //  Wilson-Taylor incompatible modes, statically condensed (PlaneStrain.h:188-243); quadrilaterals with Gauss4Square / Gauss9Square
    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStrainStiffnessWilsonTaylor(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        assert(_doulist.size() == 2);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_PLANESTRAIN_WT, SF, IC>::value, 2, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, _t);
    }

    //  consistent mass rho * N^T N * t (PlaneStrain.h:385-408)
    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStrainMass(Matrix<T>& _Me, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _rho, T _t) {
        assert(_doulist.size() == 2);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_MASS2, SF, IC>::value, 2, _Me, _nodetoelement, _element, _doulist, _x, _rho, T(0), _t);
    }

    namespace B200 {
        //  Fe = sum_g N^T f(x_g) * measure_g * t * w_g for a 2-dof field; BODY: measure = det(dXdr), weights w0*w1; else edge length, weight w0
        template<class T, template<class>class SF, template<class>class IC, class F, bool BODY>
        void LoadVector2D(Vector<T>& _Fe, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, F _f, T _t) {
            const int n = (int)_element.size();
            _Fe = Vector<T>(2*n);
            _nodetoelement = std::vector<std::vector<std::pair<int, int> > >(n, std::vector<std::pair<int, int> >(2));
            for (int i = 0; i < n; i++) { _nodetoelement[i][0] = std::make_pair(_doulist[0], 2*i); _nodetoelement[i][1] = std::make_pair(_doulist[1], 2*i + 1); }
            for (int g = 0; g < IC<T>::N; g++) {
                Vector<T> N = SF<T>::N(IC<T>::Points[g]);
                Matrix<T> dNdr = SF<T>::dNdr(IC<T>::Points[g]);
                Vector<T> xg(2);
                Matrix<T> dXdr(SF<T>::d, 2);
                for (int i = 0; i < n; i++) for (int k = 0; k < 2; k++) {
                    xg(k) += N(i)*_x[_element[i]](k);
                    for (int a = 0; a < SF<T>::d; a++) dXdr(a, k) += dNdr(a, i)*_x[_element[i]](k);
                }
                T measure, w = IC<T>::Weights[g][0];
                if (BODY) { measure = dXdr.Determinant(); w *= IC<T>::Weights[g][1]; }
                else measure = sqrt(dXdr(0, 0)*dXdr(0, 0) + dXdr(0, 1)*dXdr(0, 1));
                Vector<T> f = _f(xg);
                for (int i = 0; i < n; i++) for (int k = 0; k < 2; k++) _Fe(2*i + k) += N(i)*f(k)*measure*_t*w;
            }
        }
    }

    template<class T, template<class>class SF, template<class>class IC, class F>
    void PlaneStrainSurfaceForce(Vector<T>& _Fe, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, F _f, T _t) {
        assert(_doulist.size() == 2);
        B200::LoadVector2D<T, SF, IC, F, false>(_Fe, _nodetoelement, _element, _doulist, _x, _f, _t);
    }

    template<class T, template<class>class SF, template<class>class IC, class F>
    void PlaneStrainBodyForce(Vector<T>& _Fe, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, F _f, T _t) {
        assert(_doulist.size() == 2);
        B200::LoadVector2D<T, SF, IC, F, true>(_Fe, _nodetoelement, _element, _doulist, _x, _f, _t);
    }
}
