//  pansfem2_b200/src/FEM/Equation/Advection.h
//  The advection-diffusion element routines of the reference with its signatures (src/FEM/Equation/Advection.h):
//      Advection :19-20   AdvectionSUPG :47-48   AdvectionShockCapturing :91-92   Diffusion :135-136   Mass :161-162   MassSUPG :188-189
//  One dof per node; any 2-D <ShapeFunction, Integration>.  Each call is one PF2_PHYS_ADVDIFF selection whose routine mask has a
//  single bit (include/pansfem2_b200.h) and runs on the B200 (csrc/element_advdiff.cuh).  A driver that sums several routines per
//  element and assembles them - the two samples under sample/advection - does all of that in one launch through the batched
//  B200::AssembleAdvectionDiffusion (B200/Batched.h -> pf2_advdiff_assemble).
#pragma once
#include <vector>
#include <cassert>
#include "../../B200/ElementSelect.h"

namespace PANSFEM2 {
    namespace B200 {
        //  <SF, IC> + routine mask -> eq code; (ax, ay, k) travel in the (E, V, t) slots of pf2_element_matrix
        template<class T, template<class>class SF, template<class>class IC>
        inline void AdvectionDiffusionMatrix(int _terms, Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element,
                                             const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _ax, T _ay, T _k) {
            assert(_doulist.size() == 1);
            assert((int)_element.size() == SF<T>::n);
            ElementMatrix<T>(EqCode<PF2_PHYS_ADVDIFF, SF, IC>::value | (_terms << 24), 1, _Ke, _nodetoelement, _element, _doulist, _x, _ax, _ay, _k);
        }
    }

    template<class T, template<class>class SF, template<class>class IC>
    void Advection(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _cx, T _cy) {
        B200::AdvectionDiffusionMatrix<T, SF, IC>(PF2_ADV_ADVECTION, _Ke, _nodetoelement, _element, _doulist, _x, _cx, _cy, T(0));
    }

    template<class T, template<class>class SF, template<class>class IC>
    void AdvectionSUPG(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _ax, T _ay, T _k) {
        B200::AdvectionDiffusionMatrix<T, SF, IC>(PF2_ADV_SUPG, _Ke, _nodetoelement, _element, _doulist, _x, _ax, _ay, _k);
    }

    template<class T, template<class>class SF, template<class>class IC>
    void AdvectionShockCapturing(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _ax, T _ay, T _k) {
        B200::AdvectionDiffusionMatrix<T, SF, IC>(PF2_ADV_SHOCK, _Ke, _nodetoelement, _element, _doulist, _x, _ax, _ay, _k);
    }

    template<class T, template<class>class SF, template<class>class IC>
    void Diffusion(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _k) {
        B200::AdvectionDiffusionMatrix<T, SF, IC>(PF2_ADV_DIFFUSION, _Ke, _nodetoelement, _element, _doulist, _x, T(0), T(0), _k);
    }

    template<class T, template<class>class SF, template<class>class IC>
    void Mass(Matrix<T>& _Ce, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x) {
        B200::AdvectionDiffusionMatrix<T, SF, IC>(PF2_ADV_MASS, _Ce, _nodetoelement, _element, _doulist, _x, T(0), T(0), T(0));
    }

    template<class T, template<class>class SF, template<class>class IC>
    void MassSUPG(Matrix<T>& _Ce, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _ax, T _ay, T _k) {
        B200::AdvectionDiffusionMatrix<T, SF, IC>(PF2_ADV_MASS_SUPG, _Ce, _nodetoelement, _element, _doulist, _x, _ax, _ay, _k);
    }
}
