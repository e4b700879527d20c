//  pansfem2_b200/src/FEM/Equation/Solid.h
//  SolidLinearIsotropicElastic<T, SF, IC> with the reference's signature (src/FEM/Equation/Solid.h:20-21).
//  Tet4 + Gauss1Tetrahedron, Hex8 / Hex20 + Gauss8Cubic / Gauss27Cubic (B200/ElementSelect.h); computed on the B200.
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void SolidLinearIsotropicElastic(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V) {
        assert(_doulist.size() == 3);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_SOLID, SF, IC>::value, 3, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, T(1));
    }
}
