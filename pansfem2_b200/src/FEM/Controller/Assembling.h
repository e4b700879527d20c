//  pansfem2_b200/src/FEM/Controller/Assembling.h
//  The per-element (legacy) assembly interface of src/FEM/Controller/Assembling.h on host containers:
//      K+F+Fe :22, K+F (Dirichlet lift) :47, K+F over several node groups (mixed interpolations) :71, K only :99, F += Fe :119,
//      lift only :132, nodal Neumann :152,
//      Disassembling :163, Renumbering :175
//  Semantics kept: Dirichlet rows skipped, fixed columns lifted into F (F -= Ke*u_fixed), every Ke entry inserted
//  (explicit zeros included).  These overloads only move numbers between host containers; the hot path is the batched
//  device assembly (B200/Batched.h -> pf2_csr_pattern + pf2_assemble).
#pragma once
#include <vector>
#include <utility>
#include "../../LinearAlgebra/Models/LILCSR.h"
#include "../../LinearAlgebra/Models/Matrix.h"
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    namespace B200 {
        //  visits every (row dof, column dof) pair of an element whose row is free
        template<class F>
        inline void ForEachFreeRowPair(const std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::vector<std::pair<int, int> > >& _nodetoelement,
                                       const std::vector<int>& _element, F _visit) {
            for (size_t i = 0; i < _element.size(); i++) for (const auto& di : _nodetoelement[i]) {
                const int row = _nodetoglobal[_element[i]][di.first];
                if (row == -1) continue;
                for (size_t j = 0; j < _element.size(); j++) for (const auto& dj : _nodetoelement[j])
                    _visit(row, di.second, _nodetoglobal[_element[j]][dj.first], dj.second, _element[j], dj.first);
            }
        }
    }

    template<class T>
    void Assembling(LILCSR<T>& _K, std::vector<T>& _F, std::vector<Vector<T> >& _u, Matrix<T>& _Ke, const std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element) {
        B200::ForEachFreeRowPair(_nodetoglobal, _nodetoelement, _element, [&](int row, int lr, int col, int lc, int node, int dof) {
            if (col != -1) _K.set(row, col, _K.get(row, col) + _Ke(lr, lc));
            else _F[row] -= _Ke(lr, lc)*_u[node](dof);
        });
    }
    template<class T>
    void Assembling(LILCSR<T>& _K, std::vector<T>& _F, std::vector<Vector<T> >& _u, Matrix<T>& _Ke, Vector<T>& _Fe, const std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element) {
        Assembling(_K, _F, _u, _Ke, _nodetoglobal, _nodetoelement, _element);
        Assembling(_F, _Fe, _nodetoglobal, _nodetoelement, _element);
    }
    //  one element matrix over several node groups (e.g. velocity nodes + pressure nodes): rows and columns run over every dof of every group
    template<class T>
    void Assembling(LILCSR<T>& _K, std::vector<T>& _F, std::vector<Vector<T> >& _u, Matrix<T>& _Ke, const std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::vector<std::vector<std::pair<int, int> > > >& _nodetoelements, const std::vector<std::vector<int> >& _elements) {
        struct Dof { int global, local, node, dof; };
        std::vector<Dof> dofs;
        for (size_t g = 0; g < _elements.size(); g++) for (size_t i = 0; i < _elements[g].size(); i++) for (const auto& d : _nodetoelements[g][i])
            dofs.push_back(Dof{ _nodetoglobal[_elements[g][i]][d.first], d.second, _elements[g][i], d.first });
        for (const Dof& r : dofs) {
            if (r.global == -1) continue;
            for (const Dof& c : dofs) {
                if (c.global != -1) _K.set(r.global, c.global, _K.get(r.global, c.global) + _Ke(r.local, c.local));
                else _F[r.global] -= _Ke(r.local, c.local)*_u[c.node](c.dof);
            }
        }
    }
    template<class T>
    void Assembling(LILCSR<T>& _K, Matrix<T>& _Ke, const std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element) {
        B200::ForEachFreeRowPair(_nodetoglobal, _nodetoelement, _element, [&](int row, int lr, int col, int lc, int, int) {
            if (col != -1) _K.set(row, col, _K.get(row, col) + _Ke(lr, lc));
        });
    }
    template<class T>
    void Assembling(std::vector<T>& _F, Vector<T>& _Fe, const std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element) {
        for (size_t i = 0; i < _element.size(); i++) for (const auto& di : _nodetoelement[i]) {
            const int row = _nodetoglobal[_element[i]][di.first];
            if (row != -1) _F[row] += _Fe(di.second);
        }
    }
    template<class T>
    void Assembling(std::vector<T>& _F, std::vector<Vector<T> >& _u, Matrix<T>& _Ke, const std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element) {
        B200::ForEachFreeRowPair(_nodetoglobal, _nodetoelement, _element, [&](int row, int lr, int col, int lc, int node, int dof) {
            if (col == -1) _F[row] -= _Ke(lr, lc)*_u[node](dof);
        });
    }
    template<class T>
    void Assembling(std::vector<T>& _F, const std::vector<std::pair<std::pair<int, int>, T> >& _f, const std::vector<std::vector<int> >& _nodetoglobal) {
        for (const auto& load : _f) {
            const int row = _nodetoglobal[load.first.first][load.first.second];
            if (row != -1) _F[row] += load.second;
        }
    }
    template<class T>
    void Disassembling(std::vector<Vector<T> >& _u, const std::vector<T>& _result, const std::vector<std::vector<int> >& _nodetoglobal) {
        for (size_t i = 0; i < _nodetoglobal.size(); i++) for (size_t j = 0; j < _nodetoglobal[i].size(); j++)
            if (_nodetoglobal[i][j] != -1) _u[i]((int)j) = _result[_nodetoglobal[i][j]];
    }
    //  node-major, dof-minor running index over the free dofs; returns KDEGREE
    inline int Renumbering(std::vector<std::vector<int> >& _nodetoglobal) {
        int next = 0;
        for (auto& node : _nodetoglobal) for (auto& dof : node) if (dof != -1) dof = next++;
        return next;
    }
}
