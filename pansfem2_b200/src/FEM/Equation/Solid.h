//  pansfem2_b200/src/FEM/Equation/Solid.h
//  SolidLinearIsotropicElastic<T, SF, IC> with the reference's signature (src/FEM/Equation/Solid.h:20-21).
//  Supported selection: <double, ShapeFunction8Cubic, Gauss8Cubic>.
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void SolidLinearIsotropicElastic(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V) {
        static_assert(B200::IsH8Gauss8<SF, IC>::value, "pansfem2_b200: SolidLinearIsotropicElastic is built for ShapeFunction8Cubic + Gauss8Cubic");
        assert(_doulist.size() == 3);
        assert(_element.size() == 8);
        B200::ElementMatrix<T>(PF2_EQ_SOLID, 3, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, T(1));
    }
}
