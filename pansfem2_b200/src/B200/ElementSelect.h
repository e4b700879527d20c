//  pansfem2_b200/src/B200/ElementSelect.h
//  Maps the reference's compile-time selection <Equation, ShapeFunction, Integration> onto the kernel instantiations of
//  libpansfem2_b200.so.  Unsupported combinations fail at COMPILE time (there is no CPU fallback to fall into).
#pragma once
#include <type_traits>
#include "Device.h"
#include "../FEM/Controller/ShapeFunction.h"
#include "../FEM/Controller/GaussIntegration.h"

namespace PANSFEM2 { namespace B200 {
    //  ShapeFunction / Gauss policy class -> PF2_SHAPE_* / PF2_QUAD_* and the reference domain (0 triangle, 1 square, 2 tetrahedron, 3 cube)
    template<class SF> struct ShapeCode { static const int value = -1, domain = -1; };
    template<> struct ShapeCode<ShapeFunction3Triangle<double> > { static const int value = PF2_SHAPE_T3, domain = 0; };
    template<> struct ShapeCode<ShapeFunction6Triangle<double> > { static const int value = PF2_SHAPE_T6, domain = 0; };
    template<> struct ShapeCode<ShapeFunction4Square<double> > { static const int value = PF2_SHAPE_Q4, domain = 1; };
    template<> struct ShapeCode<ShapeFunction8Square<double> > { static const int value = PF2_SHAPE_Q8, domain = 1; };
    template<> struct ShapeCode<ShapeFunction4Tetrahedron<double> > { static const int value = PF2_SHAPE_TET4, domain = 2; };
    template<> struct ShapeCode<ShapeFunction8Cubic<double> > { static const int value = PF2_SHAPE_HEX8, domain = 3; };
    template<> struct ShapeCode<ShapeFunction20Cubic<double> > { static const int value = PF2_SHAPE_HEX20, domain = 3; };
    //  the two line shapes carry surface loads (domain 4: the line [-1, 1])
    template<> struct ShapeCode<ShapeFunction2Line<double> > { static const int value = PF2_SHAPE_LINE2, domain = 4; };
    template<> struct ShapeCode<ShapeFunction3Line<double> > { static const int value = PF2_SHAPE_LINE3, domain = 4; };
    template<class IC> struct QuadCode { static const int value = -1, domain = -1; };
    template<> struct QuadCode<Gauss1Line<double> > { static const int value = PF2_QUAD_G1LINE, domain = 4; };
    template<> struct QuadCode<Gauss2Line<double> > { static const int value = PF2_QUAD_G2LINE, domain = 4; };
    template<> struct QuadCode<Gauss1Triangle<double> > { static const int value = PF2_QUAD_G1TRI, domain = 0; };
    template<> struct QuadCode<Gauss3Triangle<double> > { static const int value = PF2_QUAD_G3TRI, domain = 0; };
    template<> struct QuadCode<Gauss1Square<double> > { static const int value = PF2_QUAD_G1SQ, domain = 1; };
    template<> struct QuadCode<Gauss4Square<double> > { static const int value = PF2_QUAD_G4SQ, domain = 1; };
    template<> struct QuadCode<Gauss9Square<double> > { static const int value = PF2_QUAD_G9SQ, domain = 1; };
    template<> struct QuadCode<Gauss1Tetrahedron<double> > { static const int value = PF2_QUAD_G1TET, domain = 2; };
    template<> struct QuadCode<Gauss8Cubic<double> > { static const int value = PF2_QUAD_G8CUBE, domain = 3; };
    template<> struct QuadCode<Gauss27Cubic<double> > { static const int value = PF2_QUAD_G27CUBE, domain = 3; };

    //  <physics, SF, IC> -> eq code of include/pansfem2_b200.h; anything the library has no kernel for fails here, at compile time
    template<int PHYS, template<class>class SF, template<class>class IC>
    struct EqCode {
        static_assert(ShapeCode<SF<double> >::value > 0, "pansfem2_b200: no kernel for this shape function");
        static_assert(QuadCode<IC<double> >::value > 0, "pansfem2_b200: no kernel for this integration rule");
        static_assert(ShapeCode<SF<double> >::domain == QuadCode<IC<double> >::domain, "pansfem2_b200: integration rule does not belong to the shape function's reference domain");
        static_assert((PHYS == PF2_PHYS_SOLID) == (ShapeCode<SF<double> >::domain >= 2), "pansfem2_b200: shape function does not match the equation's dimension");
        static const int value = PF2_EQ_CODE(PHYS, ShapeCode<SF<double> >::value, QuadCode<IC<double> >::value, 0);
    };
    template<template<class>class SF, template<class>class ICV, template<class>class ICD>
    struct EqCodeSRI {
        static_assert(QuadCode<ICV<double> >::value > 0 && QuadCode<ICV<double> >::domain == ShapeCode<SF<double> >::domain, "pansfem2_b200: volumetric rule does not fit the shape function");
        static const int value = EqCode<PF2_PHYS_PLANESTRAIN_SRI, SF, ICD>::value | (QuadCode<ICV<double> >::value << 24);
    };

    //  one element matrix through the device (the reference's per-element call, kept for compatibility and parity tests)
    template<class T>
    inline void ElementMatrix(int _eq, int _ndof, Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement,
                              const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        static_assert(std::is_same<T, double>::value, "the B200 path is instantiated for T = double");
        const int npe = (int)_element.size(), dim = ((_eq & 0xff) == PF2_PHYS_SOLID) ? 3 : 2, m = npe*_ndof;
        _nodetoelement = std::vector<std::vector<std::pair<int, int> > >(npe, std::vector<std::pair<int, int> >(_ndof));
        for (int i = 0; i < npe; i++) for (int d = 0; d < _ndof; d++) _nodetoelement[i][d] = std::make_pair(_doulist[d], _ndof*i + d);
        std::vector<double> xe((size_t)npe*dim);
        for (int i = 0; i < npe; i++) for (int d = 0; d < dim; d++) xe[(size_t)i*dim + d] = _x[_element[i]](d);
        _Ke = Matrix<T>(m, m);
        Check(pf2_element_matrix(Device::Context(), _eq, xe.data(), _E, _V, _t, _Ke.Values().data()), "pf2_element_matrix");
    }
} }
