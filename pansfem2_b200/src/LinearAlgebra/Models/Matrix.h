//  pansfem2_b200/src/LinearAlgebra/Models/Matrix.h
//  Mirror of the reference's dense small matrix (src/LinearAlgebra/Models/Matrix.h:24-89): row-major, value
//  semantics, same member names.  Determinant / Inverse keep the reference's closed forms for <= 3x3 and the
//  cofactor recursion above (Matrix.h:326-360).  Boundary type only.
#pragma once
#include <vector>
#include <cmath>
#include <cassert>
#include <iostream>
#include "Vector.h"

namespace PANSFEM2 {
    template<class T>
    class Matrix {
public:
        Matrix() : row(0), col(0) {}
        virtual ~Matrix() {}
        Matrix(int _row, int _col) : row(_row), col(_col), values((size_t)_row*_col, T()) {}
        Matrix(const Vector<T>& _vec) : row(_vec.SIZE()), col(1), values(_vec.values) {}

        int ROW() const { return row; }
        int COL() const { return col; }
        T& operator()(int _i, int _j) { assert(0 <= _i && _i < row && 0 <= _j && _j < col); return values[(size_t)_i*col + _j]; }
        const T& operator()(int _i, int _j) const { assert(0 <= _i && _i < row && 0 <= _j && _j < col); return values[(size_t)_i*col + _j]; }

        Matrix<T>& operator+=(const Matrix<T>& _mat) { assert(row == _mat.row && col == _mat.col); for (size_t i = 0; i < values.size(); i++) values[i] += _mat.values[i]; return *this; }
        Matrix<T>& operator-=(const Matrix<T>& _mat) { assert(row == _mat.row && col == _mat.col); for (size_t i = 0; i < values.size(); i++) values[i] -= _mat.values[i]; return *this; }
        Matrix<T>& operator*=(T _a) { for (auto& v : values) v *= _a; return *this; }
        Matrix<T>& operator/=(T _a) { for (auto& v : values) v /= _a; return *this; }

        Matrix<T> operator+(const Matrix<T>& _mat) const { Matrix<T> r(*this); r += _mat; return r; }
        Matrix<T> operator-(const Matrix<T>& _mat) const { Matrix<T> r(*this); r -= _mat; return r; }
        Matrix<T> operator-() const { Matrix<T> r(*this); for (auto& v : r.values) v = -v; return r; }
        Matrix<T> operator*(const Matrix<T>& _mat) const {
            assert(col == _mat.row);
            Matrix<T> r(row, _mat.col);
            for (int i = 0; i < row; i++) for (int j = 0; j < _mat.col; j++) {
                T s = T();
                for (int k = 0; k < col; k++) s += (*this)(i, k)*_mat(k, j);
                r(i, j) = s;
            }
            return r;
        }
        Vector<T> operator*(const Vector<T>& _vec) const {
            assert(col == _vec.SIZE());
            Vector<T> r(row);
            for (int i = 0; i < row; i++) { T s = T(); for (int k = 0; k < col; k++) s += (*this)(i, k)*_vec(k); r(i) = s; }
            return r;
        }
        Matrix<T> operator*(T _a) const { Matrix<T> r(*this); r *= _a; return r; }
        Matrix<T> operator/(T _a) const { Matrix<T> r(*this); r /= _a; return r; }

        Matrix<T> Transpose() const { Matrix<T> r(col, row); for (int i = 0; i < row; i++) for (int j = 0; j < col; j++) r(j, i) = (*this)(i, j); return r; }
        T Determinant() const {
            assert(row == col && row != 0);
            const std::vector<T>& v = values;
            if (row == 1) return v[0];
            if (row == 2) return v[0]*v[3] - v[1]*v[2];
            if (row == 3) return -v[8]*v[1]*v[3] - v[7]*v[5]*v[0] - v[2]*v[4]*v[6] + v[6]*v[1]*v[5] + v[7]*v[3]*v[2] + v[0]*v[4]*v[8];
            T s = T();
            for (int i = 0; i < row; i++) s += ((i & 1) ? -1.0 : 1.0)*(*this)(i, 0)*Cofactor(i, 0).Determinant();
            return s;
        }
        Matrix<T> Inverse() const {
            assert(row == col);
            Matrix<T> r(row, col);
            if (row == 1) { r(0, 0) = 1.0/values[0]; return r; }
            for (int i = 0; i < row; i++) for (int j = 0; j < col; j++) r(i, j) = (((i + j) & 1) ? -1.0 : 1.0)*Cofactor(j, i).Determinant();
            return r/Determinant();
        }
        Matrix<T> Cofactor(int _i, int _j) const {
            assert(0 <= _i && _i < row && 0 <= _j && _j < col);
            Matrix<T> r(row - 1, col - 1);
            for (int i = 0, a = 0; i < row; i++) {
                if (i == _i) continue;
                for (int j = 0, b = 0; j < col; j++) { if (j == _j) continue; r(a, b++) = (*this)(i, j); }
                a++;
            }
            return r;
        }
        Matrix<T> Vstack(const Matrix<T>& _mat) const { assert(col == _mat.col); Matrix<T> r(row + _mat.row, col); r.values = values; r.values.insert(r.values.end(), _mat.values.begin(), _mat.values.end()); return r; }
        Matrix<T> Hstack(const Matrix<T>& _mat) const {
            assert(row == _mat.row);
            Matrix<T> r(row, col + _mat.col);
            for (int i = 0; i < row; i++) { for (int j = 0; j < col; j++) r(i, j) = (*this)(i, j); for (int j = 0; j < _mat.col; j++) r(i, col + j) = _mat(i, j); }
            return r;
        }
        Matrix<T> Block(int _row, int _col, int _h, int _w) const {
            assert(0 <= _row && _row + _h <= row && 0 <= _col && _col + _w <= col);
            Matrix<T> r(_h, _w);
            for (int i = 0; i < _h; i++) for (int j = 0; j < _w; j++) r(i, j) = (*this)(_row + i, _col + j);
            return r;
        }

        std::vector<T>& Values() { return values; }                 //  flat row-major view for the device boundary (not in the reference)
        const std::vector<T>& Values() const { return values; }

        template<class F> friend class Vector;
protected:
        int row, col;
        std::vector<T> values;
    };

    template<class T> inline Vector<T>::Vector(const Matrix<T>& _mat) : values(_mat.values) { assert(_mat.col == 1); }
    template<class T> inline Matrix<T> Vector<T>::Transpose() const { Matrix<T> r(1, SIZE()); r.values = values; return r; }
    //  column vector times a 1 x n matrix: the outer product, each entry ONE product (no accumulation from zero: 0 * -1 stays -0, Vector.h:253-263)
    template<class T> inline Matrix<T> Vector<T>::operator*(const Matrix<T>& _mat) const {
        assert(_mat.row == 1);
        Matrix<T> r(SIZE(), _mat.col);
        for (int i = 0; i < SIZE(); i++) for (int j = 0; j < _mat.col; j++) r.values[(size_t)i*_mat.col + j] = values[i]*_mat.values[j];
        return r;
    }

    template<class U>
    inline std::ostream& operator<<(std::ostream& _out, const Matrix<U>& _mat) {
        for (int i = 0; i < _mat.ROW(); i++) { for (int j = 0; j < _mat.COL(); j++) _out << _mat(i, j) << "\t"; _out << std::endl; }
        return _out;
    }
    template<class U> inline Matrix<U> operator*(U _a, const Matrix<U>& _mat) { return _mat*_a; }
    template<class U> inline Matrix<U> Identity(int _row) { Matrix<U> r(_row, _row); for (int i = 0; i < _row; i++) r(i, i) = 1.0; return r; }
    template<class U> inline Matrix<U> Diagonal(const Vector<U>& _vec) { Matrix<U> r(_vec.SIZE(), _vec.SIZE()); for (int i = 0; i < _vec.SIZE(); i++) r(i, i) = _vec(i); return r; }
}
