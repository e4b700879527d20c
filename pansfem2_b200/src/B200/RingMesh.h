//  pansfem2_b200/src/B200/RingMesh.h
//  Topology shared by the reference's three ring-shaped Q4 meshers (AnnulusMesh.h, SquareAnnulusMesh.h, SquareCircleAnnulusMesh.h):
//  `layers + 1` closed loops of `around` nodes, node id = around*layer + position; element (layer, position) joins two consecutive
//  positions of two consecutive loops, counter-clockwise seen from outside the hole; the boundary edges are the inner loop run
//  backwards followed by the outer loop.  A mesher derives from RingMesh<T, Itself> and supplies Position(layer, position) for the
//  nodes and - where the reference evaluates the fixed-list predicate at different coordinates - FixedPosition(layer, position).
#pragma once
#include <algorithm>
#include <cassert>
#include <utility>
#include <vector>
#include "../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 { namespace B200 {
    template<class T, class MESHER>
    class RingMesh {
public:
        std::vector<Vector<T> > GenerateNodes() {
            std::vector<Vector<T> > nodes((size_t)around*(layers + 1));
            for (int i = 0; i <= layers; i++) for (int j = 0; j < around; j++) nodes[(size_t)around*i + j] = Self().Position(i, j);
            return nodes;
        }
        std::vector<std::vector<int> > GenerateElements() {
            std::vector<std::vector<int> > elements((size_t)around*layers);
            for (int i = 0; i < layers; i++) for (int j = 0; j < around; j++) {
                const int next = (j + 1)%around;
                elements[(size_t)around*i + j] = { around*i + j, around*(i + 1) + j, around*(i + 1) + next, around*i + next };
            }
            return elements;
        }
        std::vector<std::vector<int> > GenerateEdges() {
            std::vector<std::vector<int> > edges((size_t)2*around);
            for (int j = 0; j < around; j++) {
                const int next = (j + 1)%around;
                edges[around - j - 1] = { next, j };
                edges[j + around] = { around*layers + j, around*layers + next };
            }
            return edges;
        }
        template<class F>
        std::vector<std::pair<std::pair<int, int>, T> > GenerateFixedlist(std::vector<int> _ulist, F _iscorrespond) {
            assert(0 <= *std::min_element(_ulist.begin(), _ulist.end()));
            std::vector<std::pair<std::pair<int, int>, T> > ufixed;
            for (int i = 0; i <= layers; i++) for (int j = 0; j < around; j++)
                if (_iscorrespond(Self().FixedPosition(i, j))) for (int dof : _ulist) ufixed.push_back({ { around*i + j, dof }, T() });
            return ufixed;
        }
        Vector<T> FixedPosition(int _layer, int _position) { return Self().Position(_layer, _position); }
protected:
        RingMesh(int _around, int _layers) : around(_around), layers(_layers) {}
        int around, layers;
private:
        MESHER& Self() { return static_cast<MESHER&>(*this); }
    };
} }
