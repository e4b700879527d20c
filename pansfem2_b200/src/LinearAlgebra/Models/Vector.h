//  pansfem2_b200/src/LinearAlgebra/Models/Vector.h
//  Mirror of the reference's dense small vector (src/LinearAlgebra/Models/Vector.h:25-80): same class name,
//  namespace and member signatures, storage re-done on std::vector (value semantics come for free).
//  It is a boundary type only: the device side works on flat SoA arrays.
#pragma once
#include <vector>
#include <cmath>
#include <cassert>
#include <iostream>
#include <initializer_list>

namespace PANSFEM2 {
    template<class T> class Matrix;

    template<class T>
    class Vector {
public:
        Vector() {}
        virtual ~Vector() {}
        Vector(int _size) : values(_size, T()) {}
        Vector(const std::initializer_list<T>& _vec) : values(_vec) {}
        Vector(const std::vector<T>& _vec) : values(_vec) {}
        Vector(const Matrix<T>& _mat);

        int SIZE() const { return (int)values.size(); }
        T& operator()(int _i) { assert(0 <= _i && _i < SIZE()); return values[_i]; }
        const T& operator()(int _i) const { assert(0 <= _i && _i < SIZE()); return values[_i]; }

        Vector<T>& operator+=(const Vector<T>& _vec) { assert(SIZE() == _vec.SIZE()); for (int i = 0; i < SIZE(); i++) values[i] += _vec.values[i]; return *this; }
        Vector<T>& operator-=(const Vector<T>& _vec) { assert(SIZE() == _vec.SIZE()); for (int i = 0; i < SIZE(); i++) values[i] -= _vec.values[i]; return *this; }
        Vector<T>& operator*=(T _a) { for (auto& v : values) v *= _a; return *this; }
        Vector<T>& operator/=(T _a) { for (auto& v : values) v /= _a; return *this; }

        Vector<T> operator+(const Vector<T>& _vec) const { Vector<T> r(*this); r += _vec; return r; }
        Vector<T> operator-(const Vector<T>& _vec) const { Vector<T> r(*this); r -= _vec; return r; }
        Vector<T> operator-() const { Vector<T> r(*this); for (auto& v : r.values) v = -v; return r; }
        //  inner product, accumulated left to right like the reference (Vector.h:244)
        T operator*(const Vector<T>& _vec) const { assert(SIZE() == _vec.SIZE()); T s = T(); for (int i = 0; i < SIZE(); i++) s += values[i]*_vec.values[i]; return s; }
        Matrix<T> operator*(const Matrix<T>& _mat) const;
        Vector<T> operator*(T _a) const { Vector<T> r(*this); r *= _a; return r; }
        Vector<T> operator/(T _a) const { Vector<T> r(*this); r /= _a; return r; }

        T Norm() const { T s = T(); for (auto v : values) s += v*v; return sqrt(s); }
        Matrix<T> Transpose() const;
        Vector<T> Vstack(const Vector<T>& _vec) const { Vector<T> r(*this); r.values.insert(r.values.end(), _vec.values.begin(), _vec.values.end()); return r; }
        Vector<T> Segment(int _head, int _tail) const { assert(0 <= _head && _head <= _tail && _tail <= SIZE()); return Vector<T>(std::vector<T>(values.begin() + _head, values.begin() + _tail)); }
        Vector<T> Normal() const { return (*this)/Norm(); }

        const std::vector<T>& Values() const { return values; }     //  flat view for the device boundary (not in the reference)

        template<class F> friend class Matrix;
protected:
        std::vector<T> values;
    };

    template<class U>
    inline std::ostream& operator<<(std::ostream& _out, const Vector<U>& _vec) {
        for (int i = 0; i < _vec.SIZE(); i++) _out << _vec(i) << std::endl;
        return _out;
    }
    template<class U>
    inline Vector<U> operator*(U _a, const Vector<U>& _vec) { return _vec*_a; }
    //  std::inner_product(u.begin(), u.end(), r.begin(), 0.0) on std::vector<Vector<T>> needs T + Vector*Vector: provided by operator* above.
    template<class U>
    inline Vector<U> VectorProduct(Vector<U> _a, Vector<U> _b) {
        assert(_a.SIZE() == 3 && _b.SIZE() == 3);
        return Vector<U>({ _a(1)*_b(2) - _a(2)*_b(1), _a(2)*_b(0) - _a(0)*_b(2), _a(0)*_b(1) - _a(1)*_b(0) });
    }
}
#include "Matrix.h"
