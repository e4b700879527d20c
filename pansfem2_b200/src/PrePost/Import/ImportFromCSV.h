//  pansfem2_b200/src/PrePost/Import/ImportFromCSV.h
//  CSV readers with the reference's names, return convention (bool) and file format (src/PrePost/Import/ImportFromCSV.h:23-178):
//  one header line, then "id,v0,v1,..." rows; boundary-condition files use the token "free" for unconstrained components.
#pragma once
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <utility>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    namespace B200 {
        //  calls _row(id, tokens after the id) for every non-empty data line
        template<class F>
        inline bool ReadCsvRows(const std::string& _fname, const char* _what, F _row) {
            std::ifstream ifs(_fname);
            if (!ifs.is_open()) { std::cout << _what << " file " << _fname << " open error!" << std::endl; return false; }
            std::string line;
            std::getline(ifs, line);                        //  header
            while (std::getline(ifs, line)) {
                while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
                if (line.empty()) continue;
                std::istringstream ss(line);
                std::string tok;
                std::vector<std::string> toks;
                while (std::getline(ss, tok, ',')) toks.push_back(tok);
                if (toks.empty() || toks[0].empty()) continue;
                _row(std::stoi(toks[0]), std::vector<std::string>(toks.begin() + 1, toks.end()));
            }
            return true;
        }
    }
    template<class T>
    bool ImportNodesFromCSV(std::vector<Vector<T> >& _nodes, std::string _fname) {
        return B200::ReadCsvRows(_fname, "Node", [&](int, const std::vector<std::string>& t) {
            std::vector<T> x;
            for (const auto& s : t) x.push_back(std::stod(s));
            _nodes.push_back(Vector<T>(x));
        });
    }
    inline bool ImportElementsFromCSV(std::vector<std::vector<int> >& _elements, std::string _fname) {
        return B200::ReadCsvRows(_fname, "Element", [&](int, const std::vector<std::string>& t) {
            std::vector<int> e;
            for (const auto& s : t) e.push_back(std::stoi(s));
            _elements.push_back(e);
        });
    }
    template<class T>
    bool ImportDirichletFromCSV(std::vector<std::pair<std::pair<int, int>, T> >& _ufixed, std::string _fname) {
        return B200::ReadCsvRows(_fname, "Dirichlet Condition", [&](int id, const std::vector<std::string>& t) {
            for (size_t i = 0; i < t.size(); i++) if (t[i] != "free") _ufixed.push_back(std::make_pair(std::make_pair(id, (int)i), (T)std::stod(t[i])));
        });
    }
    template<class T>
    bool ImportNeumannFromCSV(std::vector<std::pair<std::pair<int, int>, T> >& _qfixed, std::string _fname) {
        return B200::ReadCsvRows(_fname, "Neumann Condition", [&](int id, const std::vector<std::string>& t) {
            for (size_t i = 0; i < t.size(); i++) if (t[i] != "free") _qfixed.push_back(std::make_pair(std::make_pair(id, (int)i), (T)std::stod(t[i])));
        });
    }
    //  initial values: row "id,v0,v1,..." overwrites the components of _u[id] that are not "free" (ImportFromCSV.h:181-217)
    template<class T>
    bool ImportInitialFromCSV(std::vector<Vector<T> >& _u, std::string _fname) {
        return B200::ReadCsvRows(_fname, "Initial Condition", [&](int id, const std::vector<std::string>& t) {
            for (int i = 0; i < _u[id].SIZE() && i < (int)t.size(); i++) if (t[i] != "free") _u[id](i) = (T)std::stod(t[i]);
        });
    }
    //  periodic pairs: rows "master,slave" (ImportFromCSV.h:221-250)
    inline bool ImportPeriodicFromCSV(std::vector<std::pair<int, int> >& _ufixed, std::string _fname) {
        return B200::ReadCsvRows(_fname, "Periodic Boundary Condition", [&](int master, const std::vector<std::string>& t) {
            _ufixed.push_back(std::make_pair(master, std::stoi(t.at(0))));
        });
    }
}
