//  pansfem2_b200/src/LinearAlgebra/Models/LILCSR.h
//  Mirror of the reference's list-of-lists sparse matrix (src/LinearAlgebra/Models/LILCSR.h:24-63, global namespace):
//  rows of (column, value) pairs in insertion order, set() = linear scan then append (:92-102), get() = linear scan
//  (:106-114).  It is the host-side container of the per-element legacy assembly path; the batched path never
//  builds it (pf2_csr_pattern + pf2_assemble).
#pragma once
#include <vector>
#include <utility>
#include <cassert>
#include <iostream>

template<class T> class CSR;

template<class T>
class LILCSR {
public:
    LILCSR() : ROWS(0), COLS(0) {}
    ~LILCSR() {}
    LILCSR(int _rows, int _cols) : ROWS(_rows), COLS(_cols), data(_rows) {}
    LILCSR(CSR<T> _matrix);

    const int ROWS;
    const int COLS;

    bool set(int _row, int _col, T _data) {
        for (auto& entry : data[_row]) if (entry.first == _col) { entry.second = _data; return true; }
        data[_row].push_back(std::make_pair(_col, _data));
        return false;
    }
    T get(int _row, int _col) const {
        for (const auto& entry : data[_row]) if (entry.first == _col) return entry.second;
        return T();
    }

    template<class F> friend class CSR;
    template<class T1, class T2> friend const std::vector<T1> operator*(const LILCSR<T1>& _m, const std::vector<T2>& _vec);
    template<class T1, class T2> friend const LILCSR<T1> operator+(const LILCSR<T1>& _m1, const LILCSR<T2>& _m2);
    template<class T1, class T2> friend const LILCSR<T1> operator-(const LILCSR<T1>& _m1, const LILCSR<T2>& _m2);
    template<class T1, class T2> friend const LILCSR<T1> operator*(const LILCSR<T1>& _m, T2 _a);
    template<class T1, class T2> friend const LILCSR<T1> operator/(const LILCSR<T1>& _m, T2 _a);

private:
    std::vector<std::vector<std::pair<int, T> > > data;
};

template<class T1, class T2>
inline const std::vector<T1> operator*(const LILCSR<T1>& _m, const std::vector<T2>& _vec) {
    assert(_m.COLS == (int)_vec.size());
    std::vector<T1> v(_m.ROWS, T1());
    for (int i = 0; i < _m.ROWS; i++) for (const auto& e : _m.data[i]) v[i] += e.second*_vec[e.first];
    return v;
}
template<class T1, class T2>
inline const LILCSR<T1> operator+(const LILCSR<T1>& _m1, const LILCSR<T2>& _m2) {
    assert(_m1.ROWS == _m2.ROWS && _m1.COLS == _m2.COLS);
    LILCSR<T1> m(_m1);
    for (int i = 0; i < _m2.ROWS; i++) for (const auto& e : _m2.data[i]) m.set(i, e.first, m.get(i, e.first) + e.second);
    return m;
}
template<class T1, class T2>
inline const LILCSR<T1> operator-(const LILCSR<T1>& _m1, const LILCSR<T2>& _m2) {
    assert(_m1.ROWS == _m2.ROWS && _m1.COLS == _m2.COLS);
    LILCSR<T1> m(_m1);
    for (int i = 0; i < _m2.ROWS; i++) for (const auto& e : _m2.data[i]) m.set(i, e.first, m.get(i, e.first) - e.second);
    return m;
}
template<class T1, class T2>
inline const LILCSR<T1> operator*(const LILCSR<T1>& _m, T2 _a) {
    LILCSR<T1> m(_m);
    for (auto& row : m.data) for (auto& e : row) e.second *= _a;
    return m;
}
template<class T1, class T2>
inline const LILCSR<T2> operator*(T1 _a, const LILCSR<T2>& _m) { return _m*_a; }
template<class T1, class T2>
inline const LILCSR<T1> operator/(const LILCSR<T1>& _m, T2 _a) {
    LILCSR<T1> m(_m);
    for (auto& row : m.data) for (auto& e : row) e.second /= _a;
    return m;
}
template<class F>
inline std::ostream& operator<<(std::ostream& _out, const LILCSR<F>& _mat) {
    for (int i = 0; i < _mat.ROWS; i++) { for (int j = 0; j < _mat.COLS; j++) _out << _mat.get(i, j) << "\t"; _out << std::endl; }
    return _out;
}
#include "CSR.h"
template<class T>
inline LILCSR<T>::LILCSR(CSR<T> _matrix) : ROWS(_matrix.ROWS), COLS(_matrix.COLS), data(_matrix.ROWS) {
    for (int i = 0; i < ROWS; i++) for (int k = _matrix.indptr[i]; k < _matrix.indptr[i + 1]; k++) data[i].push_back(std::make_pair(_matrix.indices[k], _matrix.data[k]));
}
