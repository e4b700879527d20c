//  pansfem2_b200/src/FEM/Equation/General.h
//  The helpers of src/FEM/Equation/General.h the TO drivers use: CenterOfGravity (:71-78), ElementVector (:82-96) and the
//  nodal <-> elemental averaging of the level-set driver, InterpolateNodalFromElemental (:179-207) / InterpolateElementalFromNodal
//  (:211-235).  Host container operations; the batched level-set loop does the same averaging on the device (csrc/levelset.cu).
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

    //**********Nodal value = mean of the adjacent elements' values**********
    namespace B200 {
        template<class V, class S>
        std::vector<V> NodalFromElemental(int _nodesize, V _un0, const std::vector<V>& _ue, const std::vector<std::vector<int> >& _elements) {
            std::vector<V> un(_nodesize, _un0);
            std::vector<int> count(_nodesize, 0);
            for (size_t i = 0; i < _elements.size(); i++) for (int node : _elements[i]) { un[node] += _ue[i]; count[node]++; }
            for (int i = 0; i < _nodesize; i++) un[i] /= (S)count[i];
            return un;
        }
        template<class V, class S>
        std::vector<V> ElementalFromNodal(V _ue0, const std::vector<V>& _un, const std::vector<std::vector<int> >& _elements) {
            std::vector<V> ue(_elements.size(), _ue0);
            for (size_t i = 0; i < _elements.size(); i++) {
                for (int node : _elements[i]) ue[i] += _un[node];
                ue[i] /= (S)_elements[i].size();
            }
            return ue;
        }
    }
    template<class T>
    std::vector<T> InterpolateNodalFromElemental(int _nodesize, T _un0, std::vector<T> _ue, std::vector<std::vector<int> > _elements) {
        return B200::NodalFromElemental<T, T>(_nodesize, _un0, _ue, _elements);
    }
    template<class T, template<class>class U>
    std::vector<U<T> > InterpolateNodalFromElemental(int _nodesize, U<T> _un0, std::vector<U<T> > _ue, std::vector<std::vector<int> > _elements) {
        return B200::NodalFromElemental<U<T>, T>(_nodesize, _un0, _ue, _elements);
    }
    template<class T>
    std::vector<T> InterpolateElementalFromNodal(T _ue0, std::vector<T> _un, std::vector<std::vector<int> > _elements) {
        return B200::ElementalFromNodal<T, T>(_ue0, _un, _elements);
    }
    template<class T, template<class>class U>
    std::vector<U<T> > InterpolateElementalFromNodal(U<T> _ue0, std::vector<U<T> > _un, std::vector<std::vector<int> > _elements) {
        return B200::ElementalFromNodal<U<T>, T>(_ue0, _un, _elements);
    }
}
