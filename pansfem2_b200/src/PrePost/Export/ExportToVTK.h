//  pansfem2_b200/src/PrePost/Export/ExportToVTK.h
//  Legacy-ASCII VTK writers with the reference's names and on-disk format (src/PrePost/Export/ExportToVTK.h:19-137): default
//  ostream precision (6 significant digits), tab separators, vectors padded to three components.  It is the format of every
//  golden file; off the timed path.  All eight entry points go through one array emitter (B200::VtkBlock); the bytes are checked
//  against the reference's writers in tests/test_io_formats.py.  `inline` on the non-template functions: this is a header.
#pragma once
#include <fstream>
#include <string>
#include <vector>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    namespace B200 {
        //  "<section>\t<count>" (optional) + "<kind> <symbol> float" (+ lookup table for scalars), then one record per entry
        struct VtkBlock {
            std::ofstream& out;
            VtkBlock(std::ofstream& _out, const char* _section, size_t _count, bool _withsection, const char* _kind, const std::string& _symbol) : out(_out) {
                if (_withsection) out << "\n" << _section << "\t" << _count << "\n";
                out << _kind << " " << _symbol << " float\n";
                if (_kind[0] == 'S') out << "LOOKUP_TABLE default\n";
            }
        };
        //  vectors are padded with zeros to three components
        template<class T>
        inline void Triple(const Vector<T>& _v, std::ofstream& _out) {
            for (int i = 0; i < 3; i++) { if (i < _v.SIZE()) _out << _v(i); else _out << T(); _out << "\t"; }
            _out << std::endl;
        }
    }

    inline void MakeHeadderToVTK(std::ofstream& _fout) {
        static const char* lines[4] = { "# vtk DataFile Version 4.1", "vtk output", "ASCII", "DATASET UNSTRUCTURED_GRID" };
        for (const char* line : lines) _fout << line << "\n";
    }
    template<class T>
    void AddPointsToVTK(std::vector<Vector<T> > _nodes, std::ofstream& _fout) {
        _fout << "\nPOINTS\t" << _nodes.size() << "\tfloat\n";
        for (const auto& node : _nodes) B200::Triple(node, _fout);
    }
    inline void AddElementToVTK(std::vector<std::vector<int> > _elements, std::ofstream& _fout) {
        size_t entries = _elements.size();
        for (const auto& e : _elements) entries += e.size();
        _fout << "\nCELLS " << _elements.size() << "\t" << entries << "\n";
        for (const auto& e : _elements) {
            _fout << e.size() << "\t";
            for (int node : e) _fout << node << "\t";
            _fout << std::endl;
        }
    }
    inline void AddElementTypes(std::vector<int> _elementtypes, std::ofstream& _fout) {
        _fout << "\nCELL_TYPES\t" << _elementtypes.size() << "\n";
        for (int type : _elementtypes) _fout << type << "\n";
    }
    template<class T>
    void AddPointScalers(std::vector<T> _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::VtkBlock block(_fout, "POINT_DATA", _values.size(), _isheader, "SCALARS", _symbol);
        for (const auto& v : _values) _fout << v << std::endl;
    }
    //  one-component nodal fields held as Vector<T> (the level-set driver's phi): Vector's operator<< ends each component with a newline
    template<class T>
    void AddPointScalers(std::vector<Vector<T> > _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::VtkBlock block(_fout, "POINT_DATA", _values.size(), _isheader, "SCALARS", _symbol);
        for (const auto& v : _values) _fout << v;
    }
    template<class T>
    void AddPointVectors(std::vector<Vector<T> > _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::VtkBlock block(_fout, "POINT_DATA", _values.size(), _isheader, "VECTORS", _symbol);
        for (const auto& v : _values) B200::Triple(v, _fout);
    }
    template<class T>
    void AddElementScalers(std::vector<T> _values, std::string _symbol, std::ofstream& _fout, bool _isheader) {
        B200::VtkBlock block(_fout, "CELL_DATA", _values.size(), _isheader, "SCALARS", _symbol);
        for (const auto& v : _values) _fout << v << std::endl;
    }
}
