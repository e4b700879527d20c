//  pansfem2_b200/src/PrePost/Mesher/SquareAnnulusMesh.h
//  SquareAnnulusMesh<T>(a, b, c, d, nx, ny, nt) of src/PrePost/Mesher/SquareAnnulusMesh.h:20-137: the frame between the rectangles
//  c x d (inner, layer 0) and a x b (outer, layer nt), both centred at the origin; a loop runs up the right side (ny nodes), along the
//  top to the left (nx), down the left side (ny) and back along the bottom (nx).  Topology and queries: B200/RingMesh.h.
//  Reference quirk kept: GenerateFixedlist evaluates its predicate with the two rectangles' roles SWAPPED (:101-131 interpolate
//  (1 - t) a + t c where GenerateNodes uses t a + (1 - t) c), so for a != c or b != d the tested coordinates are those of the
//  mirrored layer.
//  SquareAnnulusMesh2<T>(a, b, nx, ny, nv, nw) (:140-358): the nx x ny grid of the rectangle a x b with the central block of nv x nw cells
//  removed; nodes are numbered column by column skipping those strictly inside the hole, elements cell column by cell column (below
//  the hole, then above it), the boundary edges are those of the OUTER rectangle, counter-clockwise from the origin.
#pragma once
#include <algorithm>
#include <cassert>
#include <utility>
#include <vector>
#include "../../B200/RingMesh.h"

namespace PANSFEM2 {
    template<class T>
    class SquareAnnulusMesh : public B200::RingMesh<T, SquareAnnulusMesh<T> > {
public:
        SquareAnnulusMesh(T _a, T _b, T _c, T _d, int _nx, int _ny, int _nt)
            : B200::RingMesh<T, SquareAnnulusMesh<T> >(2*(_nx + _ny), _nt), a(_a), b(_b), c(_c), d(_d), nx(_nx), ny(_ny) {}
        ~SquareAnnulusMesh() {}
        Vector<T> Position(int _layer, int _position) { const T t = _layer/(T)this->layers; return OnLoop(_position, t, (1 - t)); }
        Vector<T> FixedPosition(int _layer, int _position) { const T t = _layer/(T)this->layers; return OnLoop(_position, (1 - t), t); }
private:
        //  point of the loop that blends the outer rectangle with weight wo and the inner one with weight wi
        Vector<T> OnLoop(int _p, T wo, T wi) const {
            if (_p < ny) { const T s = _p/(T)ny - 0.5; return Vector<T>({ wo*0.5*a + wi*0.5*c, wo*b*s + wi*d*s }); }
            if (_p < ny + nx) { const T s = 0.5 - (_p - ny)/(T)nx; return Vector<T>({ wo*a*s + wi*c*s, wo*0.5*b + wi*0.5*d }); }
            if (_p < 2*ny + nx) { const T s = 0.5 - (_p - ny - nx)/(T)ny; return Vector<T>({ -wo*0.5*a - wi*0.5*c, wo*b*s + wi*d*s }); }
            const T s = (_p - 2*ny - nx)/(T)nx - 0.5;
            return Vector<T>({ wo*a*s + wi*c*s, -wo*0.5*b - wi*0.5*d });
        }
        T a, b, c, d;
        int nx, ny;
    };

    template<class T>
    class SquareAnnulusMesh2 {
public:
        SquareAnnulusMesh2(T _a, T _b, int _nx, int _ny, int _nv, int _nw) : a(_a), b(_b), nx(_nx), ny(_ny), nv(_nv), nw(_nw), nxv((_nx - _nv)/2), nyw((_ny - _nw)/2) {
            assert((_nx - _nv)%2 == 0 && (_ny - _nw)%2 == 0);
        }
        ~SquareAnnulusMesh2() {}

        std::vector<Vector<T> > GenerateNodes() {
            std::vector<Vector<T> > nodes;
            ForEachNode([&](int, int i, int j) { nodes.push_back(Position(i, j)); });
            return nodes;
        }
        std::vector<std::vector<int> > GenerateElements() {
            const std::vector<int> id = NodeIds();
            std::vector<std::vector<int> > elements;
            for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
                if (nxv <= i && i < nxv + nv && nyw <= j && j < nyw + nw) continue;         //  a removed cell
                elements.push_back({ id[At(i, j)], id[At(i + 1, j)], id[At(i + 1, j + 1)], id[At(i, j + 1)] });
            }
            return elements;
        }
        std::vector<std::vector<int> > GenerateEdges() {
            const std::vector<int> id = NodeIds();
            std::vector<std::vector<int> > edges;
            for (int i = 0; i < nx; i++) edges.push_back({ id[At(i, 0)], id[At(i + 1, 0)] });
            for (int j = 0; j < ny; j++) edges.push_back({ id[At(nx, j)], id[At(nx, j + 1)] });
            for (int i = nx; i > 0; i--) edges.push_back({ id[At(i, ny)], id[At(i - 1, ny)] });
            for (int j = ny; j > 0; j--) edges.push_back({ id[At(0, j)], id[At(0, j - 1)] });
            return edges;
        }
        template<class F>
        std::vector<std::pair<std::pair<int, int>, T> > GenerateFixedlist(std::vector<int> _ulist, F _iscorrespond) {
            assert(0 <= *std::min_element(_ulist.begin(), _ulist.end()));
            std::vector<std::pair<std::pair<int, int>, T> > ufixed;
            ForEachNode([&](int id, int i, int j) { if (_iscorrespond(Position(i, j))) for (int dof : _ulist) ufixed.push_back({ { id, dof }, T() }); });
            return ufixed;
        }
private:
        int At(int _i, int _j) const { return (ny + 1)*_i + _j; }
        Vector<T> Position(int _i, int _j) const { return Vector<T>({ a*_i/(T)nx, b*_j/(T)ny }); }
        //  _visit(node id, grid column, grid row) over the grid points that carry a node, column by column
        template<class V>
        void ForEachNode(V _visit) const {
            int id = 0;
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++) {
                if (nxv < i && i < nxv + nv && nyw < j && j < nyw + nw) continue;           //  strictly inside the hole
                _visit(id++, i, j);
            }
        }
        std::vector<int> NodeIds() const {
            std::vector<int> id((size_t)(nx + 1)*(ny + 1), -1);
            ForEachNode([&](int n, int i, int j) { id[At(i, j)] = n; });
            return id;
        }
        T a, b;
        int nx, ny, nv, nw, nxv, nyw;
    };
}
