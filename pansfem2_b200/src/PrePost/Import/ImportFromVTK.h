//  pansfem2_b200/src/PrePost/Import/ImportFromVTK.h
//  Reader for the legacy-ASCII VTK files the drivers write (src/PrePost/Import/ImportFromVTK.h:23-268): ImportModelFromVTK<T>(file,
//  dimension) with ImportPOINTS / ImportCELLS / ImportPOINTVECTORS / ImportPOINTSCALARS / ImportCELLVECTORS / ImportCELLSCALARS.
//  The reference's reader is a forward scanner and callers depend on that: every call continues from where the previous one stopped,
//  POINT_DATA / CELL_DATA lines met on the way switch the section (and carry the array length), and a call that reaches the end of
//  the file without finding its array rewinds the stream and returns an EMPTY container.  One scanner implements all six calls here.
//  Host code on host containers; the file format is the one ExportToVTK.h writes.
#pragma once
#include <cassert>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    template<class T>
    class ImportModelFromVTK {
public:
        ImportModelFromVTK(std::string _fname, int _dimension) : dimension(_dimension), section(0), npoints(0), ncells(0) {
            assert(0 < _dimension && _dimension < 4);
            ifs.open(_fname);
            if (!ifs.is_open()) std::cout << "VTK file " << _fname << " open error!" << std::endl;
        }
        ~ImportModelFromVTK() {}

        std::vector<Vector<T> > ImportPOINTS() {
            std::vector<Vector<T> > nodes;
            if (Seek("POINTS", 0, "")) { nodes.assign(npoints, Vector<T>(dimension)); for (auto& node : nodes) ReadComponents(node); }
            return nodes;
        }
        std::vector<std::vector<int> > ImportCELLS() {
            std::vector<std::vector<int> > elements;
            if (Seek("CELLS", 0, "")) {
                elements.resize(ncells);
                for (auto& element : elements) {
                    std::string line;
                    std::getline(ifs, line);
                    std::stringstream values(line);
                    int n = 0;
                    values >> n;
                    element.assign(n, 0);
                    for (auto& node : element) values >> node;
                }
            }
            return elements;
        }
        std::vector<Vector<T> > ImportPOINTVECTORS(std::string _keyword) { return Vectors(1, _keyword); }
        std::vector<T> ImportPOINTSCALARS(std::string _keyword) { return Scalars(1, _keyword); }
        std::vector<Vector<T> > ImportCELLVECTORS(std::string _keyword) { return Vectors(2, _keyword); }
        std::vector<T> ImportCELLSCALARS(std::string _keyword) { return Scalars(2, _keyword); }

private:
        //  Scan forward for a line containing _tag (inside section _where when _where != 0; with second token == _name when a name is
        //  given).  POINT_DATA / CELL_DATA lines are consumed first, as in the reference, so "POINTS" / "CELLS" never see them.
        //  Not found: rewind, report false.
        bool Seek(const char* _tag, int _where, const std::string& _name) {
            std::string line;
            while (std::getline(ifs, line)) {
                if (line.find("POINT_DATA") != std::string::npos) { Count(line, npoints); section = 1; continue; }
                if (line.find("CELL_DATA") != std::string::npos) { Count(line, ncells); section = 2; continue; }
                if (_where != 0 && section != _where) continue;
                if (line.find(_tag) == std::string::npos) continue;
                if (_where == 0) { Count(line, _tag[0] == 'P' ? npoints : ncells); return true; }
                std::stringstream header(line);
                std::string kind, name;
                header >> kind >> name;
                if (name == _name) return true;
            }
            ifs.clear();
            ifs.seekg(0, std::ios_base::beg);
            return false;
        }
        static void Count(const std::string& _line, int& _count) { std::stringstream s(_line); std::string word; s >> word >> _count; }
        void ReadComponents(Vector<T>& _v) {
            std::string line;
            std::getline(ifs, line);
            std::stringstream values(line);
            for (int i = 0; i < dimension; i++) values >> _v(i);
        }
        std::vector<Vector<T> > Vectors(int _where, const std::string& _keyword) {
            std::vector<Vector<T> > out;
            if (Seek("VECTORS", _where, _keyword)) { out.assign(_where == 1 ? npoints : ncells, Vector<T>(dimension)); for (auto& v : out) ReadComponents(v); }
            return out;
        }
        std::vector<T> Scalars(int _where, const std::string& _keyword) {
            std::vector<T> out;
            if (Seek("SCALARS", _where, _keyword)) {
                out.assign(_where == 1 ? npoints : ncells, T());
                std::string line;
                std::getline(ifs, line);                //  LOOKUP_TABLE
                for (auto& v : out) { std::getline(ifs, line); std::stringstream value(line); value >> v; }
            }
            return out;
        }
        std::ifstream ifs;
        int dimension, section, npoints, ncells;
    };
}
