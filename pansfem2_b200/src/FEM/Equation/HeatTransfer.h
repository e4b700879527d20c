//  pansfem2_b200/src/FEM/Equation/HeatTransfer.h
//  HeatTransfer<T, SF, IC> with the reference's signature (src/FEM/Equation/HeatTransfer.h:19-20); 2-D only, as there.
//  Supported selection: <double, ShapeFunction4Square, Gauss4Square>.
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void HeatTransfer(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _alpha, T _t) {
        static_assert(B200::IsQ4Gauss4<SF, IC>::value, "pansfem2_b200: HeatTransfer is built for ShapeFunction4Square + Gauss4Square");
        assert(_doulist.size() == 1);
        assert(_element.size() == 4);
        B200::ElementMatrix<T>(PF2_EQ_HEAT, 1, _Ke, _nodetoelement, _element, _doulist, _x, _alpha, T(0), _t);
    }
}
