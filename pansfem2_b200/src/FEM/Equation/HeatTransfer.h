//  pansfem2_b200/src/FEM/Equation/HeatTransfer.h
//  HeatTransfer<T, SF, IC> with the reference's signature (src/FEM/Equation/HeatTransfer.h:19-20); 2-D only, as there.
//  Any of T3 / T6 / Q4 / Q8 with a rule of its reference domain (B200/ElementSelect.h); computed on the B200.
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    template<class T, template<class>class SF, template<class>class IC>
    void HeatTransfer(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _alpha, T _t) {
        assert(_doulist.size() == 1);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_HEAT, SF, IC>::value, 1, _Ke, _nodetoelement, _element, _doulist, _x, _alpha, T(0), _t);
    }

    //  HeatCapacity (HeatTransfer.h:47-71): rho * c * N N^T * t, the scalar consistent mass on the device
    template<class T, template<class>class SF, template<class>class IC>
    void HeatCapacity(Matrix<T>& _Ce, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _rho, T _c, T _t) {
        assert(_doulist.size() == 1);
        assert((int)_element.size() == SF<T>::n);
        B200::ElementMatrix<T>(B200::EqCode<PF2_PHYS_MASS, SF, IC>::value, 1, _Ce, _nodetoelement, _element, _doulist, _x, _rho*_c, T(0), _t);
    }
}
