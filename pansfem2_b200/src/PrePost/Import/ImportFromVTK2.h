//  pansfem2_b200/src/PrePost/Import/ImportFromVTK2.h
//  The token-based mesh reader of the reference (src/PrePost/Import/ImportFromVTK2.h:16-78): ImportModelFromVTK<T>(file) with
//  GenerateNodes() - three coordinates per point - and GenerateElements(); each call looks at the whole file.  (Same class name as
//  in ImportFromVTK.h, as in the reference: a translation unit includes one of the two.)
//  The file is tokenised once, on first use; both calls then walk the token list for their keyword, as the reference walks the stream.
#pragma once
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>
#include "../../LinearAlgebra/Models/Vector.h"

namespace PANSFEM2 {
    template<class T>
    class ImportModelFromVTK {
public:
        ImportModelFromVTK() = delete;
        ImportModelFromVTK(std::string _fname) : source(_fname), loaded(false) {}
        ~ImportModelFromVTK() {}

        std::vector<Vector<T> > GenerateNodes() {
            std::vector<Vector<T> > nodes;
            Walk("POINTS", [&](size_t& at, int count) {
                at++;                                                   //  the data type word ("float")
                for (int i = 0; i < count && at + 2 < tokens.size(); i++, at += 3)
                    nodes.push_back(Vector<T>({ (T)std::atof(tokens[at].c_str()), (T)std::atof(tokens[at + 1].c_str()), (T)std::atof(tokens[at + 2].c_str()) }));
            });
            return nodes;
        }
        std::vector<std::vector<int> > GenerateElements() {
            std::vector<std::vector<int> > elements;
            Walk("CELLS", [&](size_t& at, int count) {
                at++;                                                   //  the total number of integers in the block
                for (int i = 0; i < count && at < tokens.size(); i++) {
                    const int n = std::atoi(tokens[at++].c_str());
                    std::vector<int> element;
                    for (int j = 0; j < n && at < tokens.size(); j++) element.push_back(std::atoi(tokens[at++].c_str()));
                    elements.push_back(element);
                }
            });
            return elements;
        }
private:
        //  calls _block(position after the count, count) for every occurrence of the keyword
        template<class F>
        void Walk(const char* _keyword, F _block) {
            if (!loaded) {
                std::ifstream ifs(source);
                tokens.assign(std::istream_iterator<std::string>(ifs), std::istream_iterator<std::string>());
                loaded = true;
            }
            for (size_t at = 0; at < tokens.size(); ) {
                if (tokens[at++] != _keyword || at >= tokens.size()) continue;
                const int count = std::atoi(tokens[at++].c_str());
                _block(at, count);
            }
        }
        std::string source;
        bool loaded;
        std::vector<std::string> tokens;
    };
}
