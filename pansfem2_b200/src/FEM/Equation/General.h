//  pansfem2_b200/src/FEM/Equation/General.h
//  The two helpers of src/FEM/Equation/General.h the TO drivers use: CenterOfGravity (:71-78) and ElementVector (:82-96).
#pragma once
#include <vector>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    //**********Get element's center of gravity**********
    template<class T>
    Vector<T> CenterOfGravity(std::vector<Vector<T> >& _x, std::vector<int>& _element) {
        Vector<T> center(_x[0].SIZE());
        for (int node : _element) center += _x[node];
        return center/(T)_element.size();
    }

    //**********Get element vector**********
    template<class T>
    Vector<T> ElementVector(std::vector<Vector<T> >& _u, const std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element) {
        int size = 0;
        for (const auto& dofs : _nodetoelement) size += (int)dofs.size();
        Vector<T> ue(size);
        for (size_t i = 0; i < _nodetoelement.size(); i++) for (const auto& dof : _nodetoelement[i]) ue(dof.second) = _u[_element[i]](dof.first);
        return ue;
    }
}
