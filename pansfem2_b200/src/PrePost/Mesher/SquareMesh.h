//  pansfem2_b200/src/PrePost/Mesher/SquareMesh.h
//  Structured rectangle meshers with the numbering of src/PrePost/Mesher/SquareMesh.h: SquareMesh<T> (:20-245; Q4: node id (ny+1)*i + j,
//  element id ny*i + j with counter-clockwise nodes; the "2" variants are the 8-node serendipity mesh - corner nodes first, then the
//  mid-points of the vertical edges, then those of the horizontal edges -; boundary edges counter-clockwise from the origin; fixed lists
//  node-major / dof-minor) and SquareMesh2<T> (:248-420; geometric grading towards the four sides).  Checked against the reference's
//  headers in tests/test_mesher_tables.py.  Plus BoxMesh<T>, OUR x-major hex8 mesher (the reference has none): node id
//  ((ny+1)*i + j)*(nz+1) + k, bottom face CCW then top face CCW.
#pragma once
#include <vector>
#include <utility>
#include <algorithm>
#include <cassert>
#include <cmath>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    template<class T>
    class SquareMesh {
public:
        SquareMesh(T _x, T _y, int _nx, int _ny) : x(_x), y(_y), nx(_nx), ny(_ny) {}
        ~SquareMesh() {}

        std::vector<Vector<T> > GenerateNodes() {
            std::vector<Vector<T> > nodes((size_t)(nx + 1)*(ny + 1));
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++) nodes[(size_t)(ny + 1)*i + j] = Position(i, j);
            return nodes;
        }
        std::vector<std::vector<int> > GenerateElements() {
            std::vector<std::vector<int> > elements((size_t)nx*ny);
            for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
                const int n0 = (ny + 1)*i + j;
                elements[(size_t)ny*i + j] = { n0, n0 + (ny + 1), n0 + (ny + 1) + 1, n0 + 1 };
            }
            return elements;
        }
        template<class F>
        std::vector<int> GenerateElementIdsSelected(F _iscorrespond) {
            std::vector<int> ids;
            for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++)
                if (_iscorrespond(Position(i, j)) && _iscorrespond(Position(i + 1, j)) && _iscorrespond(Position(i + 1, j + 1)) && _iscorrespond(Position(i, j + 1))) ids.push_back(ny*i + j);
            return ids;
        }
        template<class F>
        std::vector<std::pair<std::pair<int, int>, T> > GenerateFixedlist(std::vector<int> _ulist, F _iscorrespond) {
            assert(0 <= *std::min_element(_ulist.begin(), _ulist.end()));
            std::vector<std::pair<std::pair<int, int>, T> > ufixed;
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++)
                if (_iscorrespond(Position(i, j))) for (int dof : _ulist) ufixed.push_back({ { (ny + 1)*i + j, dof }, T() });
            return ufixed;
        }
        //----------8-node serendipity variant----------
        std::vector<Vector<T> > GenerateNodes2() {
            std::vector<Vector<T> > nodes((size_t)(2*nx + 1)*(2*ny + 1) - (size_t)nx*ny);
            ForEachNode2([&](int id, const Vector<T>& p) { nodes[id] = p; });
            return nodes;
        }
        std::vector<std::vector<int> > GenerateElements2() {
            std::vector<std::vector<int> > elements((size_t)nx*ny);
            for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
                const int n0 = (ny + 1)*i + j;
                elements[(size_t)ny*i + j] = { n0, n0 + (ny + 1), n0 + (ny + 1) + 1, n0 + 1, MidH(i, j), MidV(i + 1, j), MidH(i, j + 1), MidV(i, j) };
            }
            return elements;
        }
        //----------boundary edges, counter-clockwise: bottom, right, top, left----------
        std::vector<std::vector<int> > GenerateEdges() {
            std::vector<std::vector<int> > edges((size_t)2*(nx + ny));
            ForEachEdge([&](int id, int a, int b, int) { edges[id] = { a, b }; });
            return edges;
        }
        std::vector<std::vector<int> > GenerateEdges2() {
            std::vector<std::vector<int> > edges((size_t)2*(nx + ny));
            ForEachEdge([&](int id, int a, int b, int mid) { edges[id] = { a, b, mid }; });
            return edges;
        }
        //  edges whose two end nodes satisfy the predicate, in the reference's order (bottom/top interleaved along x, then right/left along y)
        template<class F>
        std::vector<int> GenerateEdgeIdsSelected(F _iscorrespond) {
            std::vector<Vector<T> > nodes = GenerateNodes();
            std::vector<int> ids;
            ForEachEdge([&](int id, int a, int b, int) { if (_iscorrespond(nodes[a]) && _iscorrespond(nodes[b])) ids.push_back(id); });
            return ids;
        }
        template<class F>
        std::vector<std::pair<std::pair<int, int>, T> > GenerateFixedlist2(std::vector<int> _ulist, F _iscorrespond) {
            assert(0 <= *std::min_element(_ulist.begin(), _ulist.end()));
            std::vector<std::pair<std::pair<int, int>, T> > ufixed;
            ForEachNode2([&](int id, const Vector<T>& p) { if (_iscorrespond(p)) for (int dof : _ulist) ufixed.push_back({ { id, dof }, T() }); });
            return ufixed;
        }
private:
        Vector<T> Position(int _i, int _j) const { return Vector<T>({ x*(_i/(T)nx), y*(_j/(T)ny) }); }
        //  mid-point of the vertical edge above corner (i, j) / of the horizontal edge right of corner (i, j)
        int MidV(int _i, int _j) const { return ny*_i + _j + (nx + 1)*(ny + 1); }
        int MidH(int _i, int _j) const { return (ny + 1)*_i + _j + (nx + 1)*(2*ny + 1); }
        //  corners, then vertical-edge mid-points, then horizontal-edge mid-points: ascending ids
        template<class V>
        void ForEachNode2(V _visit) const {
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++) _visit((ny + 1)*i + j, Position(i, j));
            for (int i = 0; i <= nx; i++) for (int j = 0; j < ny; j++) _visit(MidV(i, j), Vector<T>({ x*(i/(T)nx), y*((j + 0.5)/(T)ny) }));
            for (int i = 0; i < nx; i++) for (int j = 0; j <= ny; j++) _visit(MidH(i, j), Vector<T>({ x*((i + 0.5)/(T)nx), y*(j/(T)ny) }));
        }
        //  _visit(edge id, first node, second node, mid node of the 8-node mesh)
        template<class V>
        void ForEachEdge(V _visit) const {
            for (int i = 0; i < nx; i++) {
                _visit(i, (ny + 1)*i, (ny + 1)*(i + 1), MidH(i, 0));
                _visit(2*nx + ny - i - 1, (ny + 1)*(i + 1) + ny, (ny + 1)*i + ny, MidH(i, ny));
            }
            for (int j = 0; j < ny; j++) {
                _visit(j + nx, (ny + 1)*nx + j, (ny + 1)*nx + j + 1, MidV(nx, j));
                _visit(2*(nx + ny) - j - 1, j + 1, j, MidV(0, j));
            }
        }
        T x, y;
        int nx, ny;
    };

    //  rectangle graded geometrically towards its sides (ratios rx, ry per cell, symmetric about the centre lines)
    template<class T>
    class SquareMesh2 {
public:
        SquareMesh2(T _x, T _y, int _nx, int _ny, T _rx, T _ry) : x(_x), y(_y), rx(_rx), ry(_ry), nx(_nx), ny(_ny) {}
        ~SquareMesh2() {}

        std::vector<Vector<T> > GenerateNodes() {
            std::vector<Vector<T> > nodes((size_t)(nx + 1)*(ny + 1));
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++) nodes[(size_t)(ny + 1)*i + j] = Position(i, j);
            return nodes;
        }
        std::vector<std::vector<int> > GenerateElements() { return SquareMesh<T>(x, y, nx, ny).GenerateElements(); }
        std::vector<std::vector<int> > GenerateEdges() { return SquareMesh<T>(x, y, nx, ny).GenerateEdges(); }
        template<class F>
        std::vector<std::pair<std::pair<int, int>, T> > GenerateFixedlist(std::vector<int> _ulist, F _iscorrespond) {
            assert(0 <= *std::min_element(_ulist.begin(), _ulist.end()));
            std::vector<std::pair<std::pair<int, int>, T> > ufixed;
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++)
                if (_iscorrespond(Position(i, j))) for (int dof : _ulist) ufixed.push_back({ { (ny + 1)*i + j, dof }, T() });
            return ufixed;
        }
private:
        //  distance from the nearer side grows like r0 (r^k - 1), k cells away from it; the centre line sits at exactly one half
        static T Graded(T _length, T _ratio, int _n, int _k) {
            const T r0 = 0.5*_length/(pow(_ratio, _n/2.0) - 1.0);
            if (_k < _n/2.0) return r0*(pow(_ratio, _k) - 1.0);
            if (_n/2.0 < _k) return _length - r0*(pow(_ratio, _n - _k) - 1.0);
            return 0.5*_length;
        }
        Vector<T> Position(int _i, int _j) const { return Vector<T>({ Graded(x, rx, nx, _i), Graded(y, ry, ny, _j) }); }
        T x, y, rx, ry;
        int nx, ny;
    };

    template<class T>
    class BoxMesh {
public:
        BoxMesh(T _x, T _y, T _z, int _nx, int _ny, int _nz) : x(_x), y(_y), z(_z), nx(_nx), ny(_ny), nz(_nz) {}
        std::vector<Vector<T> > GenerateNodes() {
            std::vector<Vector<T> > nodes((size_t)(nx + 1)*(ny + 1)*(nz + 1));
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++) for (int k = 0; k <= nz; k++) nodes[Id(i, j, k)] = Position(i, j, k);
            return nodes;
        }
        std::vector<std::vector<int> > GenerateElements() {
            std::vector<std::vector<int> > elements((size_t)nx*ny*nz);
            for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++)
                elements[((size_t)ny*i + j)*nz + k] = { (int)Id(i, j, k), (int)Id(i + 1, j, k), (int)Id(i + 1, j + 1, k), (int)Id(i, j + 1, k),
                                                        (int)Id(i, j, k + 1), (int)Id(i + 1, j, k + 1), (int)Id(i + 1, j + 1, k + 1), (int)Id(i, j + 1, k + 1) };
            return elements;
        }
        template<class F>
        std::vector<std::pair<std::pair<int, int>, T> > GenerateFixedlist(std::vector<int> _ulist, F _iscorrespond) {
            std::vector<std::pair<std::pair<int, int>, T> > ufixed;
            for (int i = 0; i <= nx; i++) for (int j = 0; j <= ny; j++) for (int k = 0; k <= nz; k++)
                if (_iscorrespond(Position(i, j, k))) for (int dof : _ulist) ufixed.push_back({ { (int)Id(i, j, k), dof }, T() });
            return ufixed;
        }
private:
        size_t Id(int _i, int _j, int _k) const { return ((size_t)(ny + 1)*_i + _j)*(nz + 1) + _k; }
        Vector<T> Position(int _i, int _j, int _k) const { return Vector<T>({ x*(_i/(T)nx), y*(_j/(T)ny), z*(_k/(T)nz) }); }
        T x, y, z;
        int nx, ny, nz;
    };
}
