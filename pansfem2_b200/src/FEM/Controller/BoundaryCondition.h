//  pansfem2_b200/src/FEM/Controller/BoundaryCondition.h
//  SetDirichlet (src/FEM/Controller/BoundaryCondition.h:20-34), SetPeriodic (:38-62) and RemoveBoundaryConditions (:66-72) on the
//  caller's host containers (Dirichlet by elimination: fixed dofs are marked -1).  The batched path builds the same map on the device
//  with pf2_dofmap_create.  `inline` added: the reference defines non-template free functions in a header (:66).
#pragma once
#include <vector>
#include <utility>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    template<class T>
    void SetDirichlet(std::vector<Vector<T> >& _u, std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::pair<std::pair<int, int>, T> >& _ufixed) {
        for (const auto& bc : _ufixed) {
            _u[bc.first.first](bc.first.second) = bc.second;
            _nodetoglobal[bc.first.first][bc.first.second] = -1;
        }
    }
    template<class T>
    void SetDirichlet(std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::pair<std::pair<int, int>, T> >& _ufixed) {
        for (const auto& bc : _ufixed) _nodetoglobal[bc.first.first][bc.first.second] = -1;
    }
    //  slave nodes share the master's equation numbers: mark the slaves, number what is left node-major, then copy; returns KDEGREE
    inline int SetPeriodic(std::vector<std::vector<int> >& _nodetoglobal, const std::vector<std::pair<int, int> >& _ufixed) {
        for (const auto& pair : _ufixed) for (auto& dof : _nodetoglobal[pair.second]) dof = -1;
        int next = 0;
        for (auto& node : _nodetoglobal) for (auto& dof : node) if (dof != -1) dof = next++;
        for (const auto& pair : _ufixed) _nodetoglobal[pair.second] = std::vector<int>(_nodetoglobal[pair.first].begin(), _nodetoglobal[pair.first].begin() + _nodetoglobal[pair.second].size());
        return next;
    }
    inline void RemoveBoundaryConditions(std::vector<std::vector<int> >& _nodetoglobal) {
        for (auto& node : _nodetoglobal) for (auto& dof : node) dof = 0;
    }
}
