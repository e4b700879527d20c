//  pansfem2_b200/src/FEM/Equation/Homogenization.h
//  The routines of src/FEM/Equation/Homogenization.h with the reference's signatures:
//      PlaneStiffness (:141-166), PlaneStiffnessBbar (:170-226), PlaneStiffnessWilsonTaylor (:230-280) - plane elements with a
//      caller-supplied 3 x 3 constitutive matrix: computed on the B200 (PF2_PHYS_PLANE_D*, csrc/element_generic.cuh general_rows)
//      through pf2_element_matrix_d;
//      HomogenizePlaneStrainBodyForce (:20-57), HomogenizePlaneStrainConstitutive (:61-101), HomogenizePlaneStrainCheck (:105-137) -
//      the three unit-strain load columns of a cell element and the two post-processing integrals over the characteristic
//      displacements: a few dozen flops per element on host containers, kept as host loops like the other load vectors.
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    namespace B200 {
        //  <SF, IC[, ICV]> -> eq code of the general-D selections
        template<int PHYS, template<class>class SF, template<class>class IC>
        struct EqCodePlaneD { static const int value = (EqCode<PF2_PHYS_PLANESTRAIN, SF, IC>::value & ~0xff) | PHYS; };

        template<class T>
        inline void ElementMatrixD(int _eq, Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element,
                                   const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, Matrix<T>& _D, T _t) {
            static_assert(std::is_same<T, double>::value, "the B200 path is instantiated for T = double");
            assert(_doulist.size() == 2 && _D.ROW() == 3 && _D.COL() == 3);
            const int n = (int)_element.size();
            _nodetoelement.assign(n, std::vector<std::pair<int, int> >(2));
            for (int i = 0; i < n; i++) { _nodetoelement[i][0] = std::make_pair(_doulist[0], 2*i); _nodetoelement[i][1] = std::make_pair(_doulist[1], 2*i + 1); }
            std::vector<double> xe((size_t)2*n);
            for (int i = 0; i < n; i++) { xe[2*(size_t)i] = _x[_element[i]](0); xe[2*(size_t)i + 1] = _x[_element[i]](1); }
            double D9[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) D9[3*i + j] = _D(i, j);
            _Ke = Matrix<T>(2*n, 2*n);
            Check(pf2_element_matrix_d(Device::Context(), _eq, xe.data(), D9, _t, _Ke.Values().data()), "pf2_element_matrix_d");
        }

        //  strain-displacement matrix B (3 x 2n) and J = det(dX/dr) at integration point g
        template<class T, template<class>class SF, template<class>class IC>
        inline T StrainMatrix(int _g, const std::vector<int>& _element, std::vector<Vector<T> >& _x, Matrix<T>& _B) {
            const int n = (int)_element.size();
            Matrix<T> X(n, 2);
            for (int i = 0; i < n; i++) { X(i, 0) = _x[_element[i]](0); X(i, 1) = _x[_element[i]](1); }
            Matrix<T> dNdr = SF<T>::dNdr(IC<T>::Points[_g]);
            Matrix<T> dXdr = dNdr*X;
            Matrix<T> dNdX = dXdr.Inverse()*dNdr;
            _B = Matrix<T>(3, 2*n);
            for (int i = 0; i < n; i++) { _B(0, 2*i) = dNdX(0, i); _B(1, 2*i + 1) = dNdX(1, i); _B(2, 2*i) = dNdX(1, i); _B(2, 2*i + 1) = dNdX(0, i); }
            return dXdr.Determinant();
        }
        template<class T>
        inline Matrix<T> PlaneStrainD(T _E, T _V) {
            Matrix<T> D(3, 3);
            D(0, 0) = 1.0 - _V; D(0, 1) = _V; D(1, 0) = _V; D(1, 1) = 1.0 - _V; D(2, 2) = 0.5*(1.0 - 2.0*_V);
            D *= _E/((1.0 - 2.0*_V)*(1.0 + _V));
            return D;
        }
        //  characteristic displacements of an element as a 2n x 3 matrix (one column per unit strain)
        template<class T>
        inline Matrix<T> Characteristic(const std::vector<int>& _element, std::vector<Vector<T> >& _chi0, std::vector<Vector<T> >& _chi1, std::vector<Vector<T> >& _chi2) {
            const int n = (int)_element.size();
            Matrix<T> CHI(2*n, 3);
            for (int i = 0; i < n; i++) for (int d = 0; d < 2; d++) {
                CHI(2*i + d, 0) = _chi0[_element[i]](d); CHI(2*i + d, 1) = _chi1[_element[i]](d); CHI(2*i + d, 2) = _chi2[_element[i]](d);
            }
            return CHI;
        }
    }

    //  Fes (2n x 3) = sum_g B^T D J t w: the nodal loads of the three unit macroscopic strains
    template<class T, template<class>class SF, template<class>class IC>
    void HomogenizePlaneStrainBodyForce(Matrix<T>& _Fes, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _E, T _V, T _t) {
        assert(_doulist.size() == 2);
        const int n = (int)_element.size();
        _Fes = Matrix<T>(2*n, 3);
        _nodetoelement.assign(n, std::vector<std::pair<int, int> >(2));
        for (int i = 0; i < n; i++) { _nodetoelement[i][0] = std::make_pair(_doulist[0], 2*i); _nodetoelement[i][1] = std::make_pair(_doulist[1], 2*i + 1); }
        Matrix<T> D = B200::PlaneStrainD(_E, _V);
        for (int g = 0; g < IC<T>::N; g++) {
            Matrix<T> B;
            const T J = B200::StrainMatrix<T, SF, IC>(g, _element, _x, B);
            _Fes += B.Transpose()*D*J*_t*IC<T>::Weights[g][0]*IC<T>::Weights[g][1];
        }
    }

    //  C (3 x 3) = sum_g D (I - B chi) J t w
    template<class T, template<class>class SF, template<class>class IC>
    Matrix<T> HomogenizePlaneStrainConstitutive(std::vector<Vector<T> >& _x, std::vector<int>& _element, std::vector<Vector<T> >& _chi0, std::vector<Vector<T> >& _chi1, std::vector<Vector<T> >& _chi2, T _E, T _V, T _t) {
        Matrix<T> C(3, 3), CHI = B200::Characteristic(_element, _chi0, _chi1, _chi2), D = B200::PlaneStrainD(_E, _V), I = Identity<T>(3);
        for (int g = 0; g < IC<T>::N; g++) {
            Matrix<T> B;
            const T J = B200::StrainMatrix<T, SF, IC>(g, _element, _x, B);
            C += D*(I - B*CHI)*J*_t*IC<T>::Weights[g][0]*IC<T>::Weights[g][1];
        }
        return C;
    }

    //  C (3 x 3) = sum_g -B chi J t w
    template<class T, template<class>class SF, template<class>class IC>
    Matrix<T> HomogenizePlaneStrainCheck(std::vector<Vector<T> >& _x, std::vector<int>& _element, std::vector<Vector<T> >& _chi0, std::vector<Vector<T> >& _chi1, std::vector<Vector<T> >& _chi2, T _t) {
        Matrix<T> C(3, 3), CHI = B200::Characteristic(_element, _chi0, _chi1, _chi2);
        for (int g = 0; g < IC<T>::N; g++) {
            Matrix<T> B;
            const T J = B200::StrainMatrix<T, SF, IC>(g, _element, _x, B);
            C += -B*CHI*J*_t*IC<T>::Weights[g][0]*IC<T>::Weights[g][1];
        }
        return C;
    }

    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStiffness(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, Matrix<T> _D, T _t) {
        B200::ElementMatrixD<T>(B200::EqCodePlaneD<PF2_PHYS_PLANE_D, SF, IC>::value, _Ke, _nodetoelement, _element, _doulist, _x, _D, _t);
    }
    //  volumetric part of B integrated with ICV, deviatoric part with ICD
    template<class T, template<class>class SF, template<class>class ICV, template<class>class ICD>
    void PlaneStiffnessBbar(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, Matrix<T> _D, T _t) {
        static_assert(B200::QuadCode<ICV<double> >::value > 0 && B200::QuadCode<ICV<double> >::domain == B200::ShapeCode<SF<double> >::domain, "pansfem2_b200: volumetric rule does not fit the shape function");
        const int eq = B200::EqCodePlaneD<PF2_PHYS_PLANE_D_BBAR, SF, ICD>::value | (B200::QuadCode<ICV<double> >::value << 24);
        B200::ElementMatrixD<T>(eq, _Ke, _nodetoelement, _element, _doulist, _x, _D, _t);
    }
    template<class T, template<class>class SF, template<class>class IC>
    void PlaneStiffnessWilsonTaylor(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, Matrix<T> _D, T _t) {
        B200::ElementMatrixD<T>(B200::EqCodePlaneD<PF2_PHYS_PLANE_D_WT, SF, IC>::value, _Ke, _nodetoelement, _element, _doulist, _x, _D, _t);
    }
}
