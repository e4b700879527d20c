//  pansfem2_b200/src/PrePost/Export/ExportToVTK.h
//  Legacy-ASCII VTK writers with the reference's names and on-disk format (src/PrePost/Export/ExportToVTK.h:19-137): default
//  ostream precision (6 significant digits), tab separators, vectors padded to three components.  It is the format of every
//  golden file; off the timed path.  `inline` added to the non-template functions (the reference defines them in a header).
#pragma once
#include <fstream>
#include <string>
#include <vector>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    namespace B200 {
        template<class T>
        inline void WritePadded3(const Vector<T>& _v, std::ofstream& _fout) {
            for (int i = 0; i < 3; i++) _fout << (i < _v.SIZE() ? _v(i) : T()) << "\t";
            _fout << std::endl;
        }
        inline void SectionHeader(const char* _kind, size_t _count, bool _isheader, std::ofstream& _fout) { if (_isheader) _fout << "\n" << _kind << "\t" << _count << "\n"; }
    }
    inline void MakeHeadderToVTK(std::ofstream& _fout) { _fout << "# vtk DataFile Version 4.1\nvtk output\nASCII\nDATASET UNSTRUCTURED_GRID\n"; }
    template<class T>
    void AddPointsToVTK(std::vector<Vector<T> > _nodes, std::ofstream& _fout) {
        _fout << "\nPOINTS\t" << _nodes.size() << "\tfloat\n";
        for (const auto& node : _nodes) B200::WritePadded3(node, _fout);
    }
    inline void AddElementToVTK(std::vector<std::vector<int> > _elements, std::ofstream& _fout) {
        size_t total = 0;
        for (const auto& e : _elements) total += e.size() + 1;
        _fout << "\nCELLS " << _elements.size() << "\t" << total << "\n";
        for (const auto& e : _elements) { _fout << e.size() << "\t"; for (int node : e) _fout << node << "\t"; _fout << std::endl; }
    }
    inline void AddElementTypes(std::vector<int> _elementtypes, std::ofstream& _fout) {
        _fout << "\nCELL_TYPES\t" << _elementtypes.size() << "\n";
        for (int t : _elementtypes) _fout << t << "\n";
    }
    template<class T>
    void AddPointScalers(std::vector<T> _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::SectionHeader("POINT_DATA", _values.size(), _isheader, _fout);
        _fout << "SCALARS " << _symbol << " float\nLOOKUP_TABLE default\n";
        for (const auto& v : _values) _fout << v << std::endl;
    }
    //  one-component nodal fields held as Vector<T> (the level-set driver's phi): Vector's operator<< ends each component with a newline
    template<class T>
    void AddPointScalers(std::vector<Vector<T> > _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::SectionHeader("POINT_DATA", _values.size(), _isheader, _fout);
        _fout << "SCALARS " << _symbol << " float\nLOOKUP_TABLE default\n";
        for (const auto& v : _values) _fout << v;
    }
    template<class T>
    void AddPointVectors(std::vector<Vector<T> > _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::SectionHeader("POINT_DATA", _values.size(), _isheader, _fout);
        _fout << "VECTORS " << _symbol << " float\n";
        for (const auto& v : _values) B200::WritePadded3(v, _fout);
    }
    template<class T>
    void AddElementScalers(std::vector<T> _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::SectionHeader("CELL_DATA", _values.size(), _isheader, _fout);
        _fout << "SCALARS " << _symbol << " float\nLOOKUP_TABLE default\n";
        for (const auto& v : _values) _fout << v << std::endl;
    }
}
