//  pansfem2_b200/src/FEM/Controller/BoundaryCondition.h
//  SetDirichlet (src/FEM/Controller/BoundaryCondition.h:20-34) and RemoveBoundaryConditions (:66-72) on the caller's host
//  containers (Dirichlet by elimination: fixed dofs are marked -1).  The batched path builds the same map on the device
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
    inline void RemoveBoundaryConditions(std::vector<std::vector<int> >& _nodetoglobal) {
        for (auto& node : _nodetoglobal) for (auto& dof : node) dof = 0;
    }
}
