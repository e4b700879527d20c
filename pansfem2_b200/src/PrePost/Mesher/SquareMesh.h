//  pansfem2_b200/src/PrePost/Mesher/SquareMesh.h
//  Structured Q4 rectangle mesher with the numbering of src/PrePost/Mesher/SquareMesh.h:62-107,194-207: node id (ny+1)*i + j,
//  element id ny*i + j with counter-clockwise nodes, fixed lists node-major / dof-minor.  Plus BoxMesh<T>, OUR x-major hex8
//  mesher (the reference has none): node id ((ny+1)*i + j)*(nz+1) + k, bottom face CCW then top face CCW.
#pragma once
#include <vector>
#include <utility>
#include <algorithm>
#include <cassert>
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
private:
        Vector<T> Position(int _i, int _j) const { return Vector<T>({ x*(_i/(T)nx), y*(_j/(T)ny) }); }
        T x, y;
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
