//  pansfem2_b200/src/FEM/Equation/HeatTransfer.h
//  HeatTransfer<T, SF, IC> with the reference's signature (src/FEM/Equation/HeatTransfer.h:19-20); 2-D only, as there.
//  Any of T3 / T6 / Q4 / Q8 with a rule of its reference domain (B200/ElementSelect.h); computed on the B200.
#pragma once
#include <vector>
#include <cassert>
#include <cmath>
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

    //  HeatTransferSurfaceFlux (HeatTransfer.h:73-98): Fe = sum_g N f(x_g) |dX/dr| t w_g over an edge element; takes a C++ functor, host loop
    template<class T, template<class>class SF, template<class>class IC, class F>
    void HeatTransferSurfaceFlux(Vector<T>& _Fe, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, F _f, T _t) {
        assert(_doulist.size() == 1);
        const int n = (int)_element.size();
        _Fe = Vector<T>(n);
        _nodetoelement.assign(n, std::vector<std::pair<int, int> >(1));
        for (int i = 0; i < n; i++) _nodetoelement[i][0] = std::make_pair(_doulist[0], i);
        for (int g = 0; g < IC<T>::N; g++) {
            Vector<T> N = SF<T>::N(IC<T>::Points[g]);
            Matrix<T> dNdr = SF<T>::dNdr(IC<T>::Points[g]);
            Vector<T> xg(2);
            T tx = T(), ty = T();
            for (int i = 0; i < n; i++) {
                xg(0) += _x[_element[i]](0)*N(i); xg(1) += _x[_element[i]](1)*N(i);
                tx += dNdr(0, i)*_x[_element[i]](0); ty += dNdr(0, i)*_x[_element[i]](1);
            }
            const T dl = sqrt(tx*tx + ty*ty);
            _Fe += N*_f(xg)*dl*_t*IC<T>::Weights[g][0];
        }
    }
}
