//  pansfem2_b200/src/B200/ElementSelect.h
//  Maps the reference's compile-time selection <Equation, ShapeFunction, Integration> onto the kernel instantiations of
//  libpansfem2_b200.so.  Unsupported combinations fail at COMPILE time (there is no CPU fallback to fall into).
#pragma once
#include <type_traits>
#include "Device.h"
#include "../FEM/Controller/ShapeFunction.h"
#include "../FEM/Controller/GaussIntegration.h"

namespace PANSFEM2 { namespace B200 {
    template<template<class>class SF, template<class>class IC>
    struct IsQ4Gauss4 : std::integral_constant<bool, std::is_same<SF<double>, ShapeFunction4Square<double> >::value && std::is_same<IC<double>, Gauss4Square<double> >::value> {};
    template<template<class>class SF, template<class>class IC>
    struct IsH8Gauss8 : std::integral_constant<bool, std::is_same<SF<double>, ShapeFunction8Cubic<double> >::value && std::is_same<IC<double>, Gauss8Cubic<double> >::value> {};

    //  one element matrix through the device (the reference's per-element call, kept for compatibility and parity tests)
    template<class T>
    inline void ElementMatrix(int _eq, int _ndof, Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement,
                              const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        static_assert(std::is_same<T, double>::value, "the B200 path is instantiated for T = double");
        const int npe = (int)_element.size(), dim = (_eq == PF2_EQ_SOLID) ? 3 : 2, m = npe*_ndof;
        _nodetoelement = std::vector<std::vector<std::pair<int, int> > >(npe, std::vector<std::pair<int, int> >(_ndof));
        for (int i = 0; i < npe; i++) for (int d = 0; d < _ndof; d++) _nodetoelement[i][d] = std::make_pair(_doulist[d], _ndof*i + d);
        std::vector<double> xe((size_t)npe*dim);
        for (int i = 0; i < npe; i++) for (int d = 0; d < dim; d++) xe[(size_t)i*dim + d] = _x[_element[i]](d);
        _Ke = Matrix<T>(m, m);
        Check(pf2_element_matrix(Device::Context(), _eq, xe.data(), _E, _V, _t, _Ke.Values().data()), "pf2_element_matrix");
    }
} }
