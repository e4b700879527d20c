//  pansfem2_b200/src/FEM/Equation/General.h
//  The helpers of src/FEM/Equation/General.h: Area (:20-41), Volume (:45-66), CenterOfGravity (:71-78), ElementVector (:82-96 and the
//  multi-element overload :100-119), WeakSpring (:123-134), LagrangeInterpolation / ...Derivative (:138-173) and the
//  nodal <-> elemental averaging of the level-set driver, InterpolateNodalFromElemental (:179-207) / InterpolateElementalFromNodal
//  (:211-235).  Host container operations; the batched level-set loop does the same averaging on the device (csrc/levelset.cu).
#pragma once
#include <vector>
#include <utility>
#include "../../LinearAlgebra/Models/Vector.h"
#include "../../LinearAlgebra/Models/Matrix.h"

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

    //**********Element vector over several node groups (mixed interpolations)**********
    template<class T>
    Vector<T> ElementVector(std::vector<Vector<T> >& _u, const std::vector<std::vector<std::vector<std::pair<int, int> > > >& _nodetoelements, const std::vector<std::vector<int> >& _elements) {
        int size = 0;
        for (const auto& group : _nodetoelements) for (const auto& dofs : group) size += (int)dofs.size();
        Vector<T> ue(size);
        for (size_t g = 0; g < _nodetoelements.size(); g++) for (size_t i = 0; i < _nodetoelements[g].size(); i++)
            for (const auto& dof : _nodetoelements[g][i]) ue(dof.second) = _u[_elements[g][i]](dof.first);
        return ue;
    }

    //**********Measure of an element: sum of det(dX/dr) * weights over the rule**********
    namespace B200 {
        template<class T, template<class>class SF, template<class>class IC, int DIM>
        T Measure(std::vector<Vector<T> >& _x, std::vector<int>& _element) {
            const int n = (int)_element.size();
            Matrix<T> X(n, DIM);
            for (int i = 0; i < n; i++) for (int k = 0; k < DIM; k++) X(i, k) = _x[_element[i]](k);
            T measure = T();
            for (int g = 0; g < IC<T>::N; g++) {
                Matrix<T> dXdr = SF<T>::dNdr(IC<T>::Points[g])*X;
                T term = dXdr.Determinant();
                for (int k = 0; k < DIM; k++) term *= IC<T>::Weights[g][k];
                measure += term;
            }
            return measure;
        }
    }
    template<class T, template<class>class SF, template<class>class IC>
    T Area(std::vector<Vector<T> >& _x, std::vector<int>& _element) { return B200::Measure<T, SF, IC, 2>(_x, _element); }
    template<class T, template<class>class SF, template<class>class IC>
    T Volume(std::vector<Vector<T> >& _x, std::vector<int>& _element) { return B200::Measure<T, SF, IC, 3>(_x, _element); }

    //**********Weak spring against rigid-body modes: alpha * I.  (The reference maps EVERY dof of node i to local column n*i,
    //          General.h:129 - kept, callers' assembled matrices depend on it.)**********
    template<class T>
    void WeakSpring(Matrix<T>& _Ke, std::vector<std::vector<std::pair<int, int> > >& _nodetoelement, const std::vector<int>& _element, const std::vector<int>& _doulist, std::vector<Vector<T> >& _x, T _alpha) {
        const int n = (int)_doulist.size();
        _nodetoelement.assign(_element.size(), std::vector<std::pair<int, int> >(n));
        for (size_t i = 0; i < _element.size(); i++) for (int j = 0; j < n; j++) _nodetoelement[i][j] = std::make_pair(_doulist[j], n*(int)i);
        _Ke = _alpha*Identity<T>(n*(int)_element.size());
    }

    //**********Lagrange basis on the abscissae _xs and its derivative, evaluated at _x**********
    template<class T>
    std::vector<T> LagrangeInterpolation(std::vector<T> _xs, T _x) {
        std::vector<T> N(_xs.size(), 1.0);
        for (size_t i = 0; i < _xs.size(); i++) for (size_t j = 0; j < _xs.size(); j++) if (i != j) N[i] *= (_x - _xs[j])/(T)(_xs[i] - _xs[j]);
        return N;
    }
    template<class T>
    std::vector<T> LagrangeInterpolationDerivative(std::vector<T> _xs, T _x) {
        std::vector<T> dN(_xs.size(), T());
        for (size_t i = 0; i < _xs.size(); i++) for (size_t j = 0; j < _xs.size(); j++) {
            if (j == i) continue;
            T term = 1.0;                           //  1/(x_i - x_j) * prod_{k != i, j} (x - x_k)/(x_i - x_k), factors in ascending k
            for (size_t k = 0; k < _xs.size(); k++) {
                if (k == i) continue;
                if (k != j) term *= (_x - _xs[k])/(T)(_xs[i] - _xs[k]);
                else term *= 1.0/(T)(_xs[i] - _xs[k]);
            }
            dN[i] += term;
        }
        return dN;
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
