//  pansfem2_b200/src/FEM/Equation/PlaneStress.h
//  PlaneStressStiffness<T, SF, IC> (src/FEM/Equation/PlaneStress.h:20-21), PlaneStressSurfaceForce (:97-98) and
//  PlaneStressBodyForce (:133-134) with the reference's signatures; stiffness on the B200, load vectors on the host
//  (they are the plane-strain ones: the integrand does not involve D).
#pragma once
#include "PlaneStrain.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStressStiffness(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        assert(_doulist.size() == 2);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_PLANESTRESS, SF, IC>::value, 2, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, _t);
    }

    //  PlaneStressMass (PlaneStress.h:62-63): the same rho * N^T N * t
    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStressMass(Matrix<T>& _Me, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _rho, T _t) {
        PlaneStrainMass<T, SF, IC>(_Me, _nodetoelement, _element, _doulist, _x, _rho, _t);
    }

    template<class T, template<class>class SF, template<class>class IC, class F>
    void PlaneStressSurfaceForce(Vector<T>& _Fe, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, F _f, T _t) {
        assert(_doulist.size() == 2);
        B200::LoadVector2D<T, SF, IC, F, false>(_Fe, _nodetoelement, _element, _doulist, _x, _f, _t);
    }

    template<class T, template<class>class SF, template<class>class IC, class F>
    void PlaneStressBodyForce(Vector<T>& _Fe, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, F _f, T _t) {
        assert(_doulist.size() == 2);
        B200::LoadVector2D<T, SF, IC, F, true>(_Fe, _nodetoelement, _element, _doulist, _x, _f, _t);
    }
}
