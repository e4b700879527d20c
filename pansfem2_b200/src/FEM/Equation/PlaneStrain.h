//  pansfem2_b200/src/FEM/Equation/PlaneStrain.h
//  PlaneStrainStiffness<T, SF, IC> with the reference's signature (src/FEM/Equation/PlaneStrain.h:20-21).
//  Supported selection: <double, ShapeFunction4Square, Gauss4Square> (the one every TO sample uses).
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStrainStiffness(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        static_assert(B200::IsQ4Gauss4<SF, IC>::value, "pansfem2_b200: PlaneStrainStiffness is built for ShapeFunction4Square + Gauss4Square");
        assert(_doulist.size() == 2);
        assert(_element.size() == 4);
        B200::ElementMatrix<T>(PF2_EQ_PLANESTRAIN, 2, _Ke, _nodetoelement, _element, _doulist, _x, _E, _V, _t);
    }
}
