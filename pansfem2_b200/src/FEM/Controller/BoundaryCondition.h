//  pansfem2_b200/src/FEM/Controller/BoundaryCondition.h
//  Boundary-condition bookkeeping on the caller's host containers, with the reference's names and argument order
//  (src/FEM/Controller/BoundaryCondition.h): SetDirichlet with and without the field (:20-34), SetPeriodic (:38-62),
//  RemoveBoundaryConditions (:66-72).
//
//  Convention (Dirichlet by elimination): `numbering[node][dof]` holds -1 for a constrained dof and, after Renumbering
//  (Assembling.h), the equation number of a free one.  The batched path builds the same table on the device with
//  pf2_dofmap_create; these overloads serve the per-element path and user code.  The non-template functions are `inline`
//  because this is a header (the reference defines them without it, :38 and :66).
#pragma once
#include <algorithm>
#include <utility>
#include <vector>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    namespace B200 {
        typedef std::vector<std::vector<int> > Numbering;
        inline void MarkConstrained(Numbering& numbering, int node, int dof) { numbering[node][dof] = -1; }
    }

    //  prescribe values: write them into the field and take the dofs out of the system
    template<class T>
    void SetDirichlet(std::vector<Vector<T> >& field, B200::Numbering& numbering, const std::vector<std::pair<std::pair<int, int>, T> >& prescribed) {
        for (const auto& entry : prescribed) {
            const int node = entry.first.first, dof = entry.first.second;
            field[node](dof) = entry.second;
            B200::MarkConstrained(numbering, node, dof);
        }
    }

    //  take the listed dofs out of the system without touching a field
    template<class T>
    void SetDirichlet(B200::Numbering& numbering, const std::vector<std::pair<std::pair<int, int>, T> >& prescribed) {
        for (const auto& entry : prescribed) B200::MarkConstrained(numbering, entry.first.first, entry.first.second);
    }

    //  (master, slave) node pairs: the slaves are taken out, what is left is numbered node-major / dof-minor, and every slave then
    //  shares its master's equation numbers.  Returns the number of equations.
    inline int SetPeriodic(B200::Numbering& numbering, const std::vector<std::pair<int, int> >& pairs) {
        for (const auto& p : pairs) std::fill(numbering[p.second].begin(), numbering[p.second].end(), -1);
        int equations = 0;
        for (auto& node : numbering) for (int& dof : node) if (dof != -1) dof = equations++;
        for (const auto& p : pairs) std::copy_n(numbering[p.first].begin(), numbering[p.second].size(), numbering[p.second].begin());
        return equations;
    }

    //  forget every constraint (all dofs free, unnumbered)
    inline void RemoveBoundaryConditions(B200::Numbering& numbering) {
        for (auto& node : numbering) std::fill(node.begin(), node.end(), 0);
    }
}
