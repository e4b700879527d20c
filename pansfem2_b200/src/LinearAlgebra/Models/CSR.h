//  pansfem2_b200/src/LinearAlgebra/Models/CSR.h
//  Mirror of the reference's CSR matrix (src/LinearAlgebra/Models/CSR.h:25-75, global namespace).
//  The three host arrays keep the reference's layout and semantics (built from LILCSR by per-row sort, :93-105;
//  set/get :126-167).  For T = double the product y = A*x (:109-122) and everything in Solvers/CG.h run on the B200
//  through the C ABI: the matrix is mirrored to the device lazily and re-uploaded only after a host-side mutation.
#pragma once
#include <vector>
#include <algorithm>
#include <cassert>
#include <iostream>
#include <memory>
#include <type_traits>
#include "../../B200/Device.h"

template<class T> class LILCSR;

namespace PANSFEM2 { namespace B200 {
    //  shared device mirror of one CSR<double>; copies of a CSR share it until one of them is mutated
    struct CsrDevice {
        pf2_csr* handle;
        bool owned;             //  false: adopted from a B200::Model, which destroys it itself
        CsrDevice() : handle(nullptr), owned(true) {}
        ~CsrDevice() { if (handle && owned) pf2_csr_destroy(handle); }
    };
} }

template<class T>
class CSR {
public:
    CSR() : ROWS(0), COLS(0) {}
    ~CSR() {}
    CSR(int _rows, int _cols) : ROWS(_rows), COLS(_cols), indptr(_rows + 1, 0) {}
    CSR(LILCSR<T>& _matrix) : ROWS(_matrix.ROWS), COLS(_matrix.COLS), indptr(_matrix.ROWS + 1, 0) {
        for (int i = 0; i < ROWS; i++) {
            auto& row = _matrix.data[i];
            std::sort(row.begin(), row.end());       //  the LILCSR argument is sorted in place, as the reference does (:99)
            indptr[i + 1] = indptr[i] + (int)row.size();
            for (const auto& e : row) { indices.push_back(e.first); data.push_back(e.second); }
        }
    }
    //  view of an already assembled device matrix (batched path; the handle stays owned by whoever built it, e.g. B200::Model):
    //  host arrays are filled on first host access
    explicit CSR(pf2_csr* _device) : ROWS(DeviceRows(_device)), COLS(DeviceRows(_device)) {
        device = std::make_shared<PANSFEM2::B200::CsrDevice>();
        device->handle = _device;
        device->owned = false;
        host_stale = true;
    }

    const int ROWS;
    const int COLS;

    const std::vector<T> operator*(const std::vector<T>& _vec) {
        assert((int)_vec.size() == COLS);
        static_assert(std::is_same<T, double>::value, "CSR<T>::operator* runs on the B200 and is instantiated for T = double only (no CPU fallback)");
        std::vector<T> v(ROWS, T());
        PANSFEM2::B200::Check(pf2_spmv_host(Device(), _vec.data(), v.data()), "pf2_spmv_host");
        return v;
    }

    bool set(int _row, int _col, T _data) {
        SyncHost();
        auto first = indices.begin() + indptr[_row], last = indices.begin() + indptr[_row + 1];
        auto pos = std::lower_bound(first, last, _col);
        const size_t k = pos - indices.begin();
        device.reset();
        if (pos != last && *pos == _col) { data[k] = _data; return true; }
        indices.insert(pos, _col);
        data.insert(data.begin() + k, _data);
        for (int i = _row + 1; i <= ROWS; i++) indptr[i] += 1;
        return false;
    }
    T get(int _row, int _col) const {
        const_cast<CSR<T>*>(this)->SyncHost();
        auto first = indices.begin() + indptr[_row], last = indices.begin() + indptr[_row + 1];
        auto pos = std::lower_bound(first, last, _col);
        return (pos != last && *pos == _col) ? data[pos - indices.begin()] : T();
    }

    //  device mirror (T = double): uploaded on first use
    pf2_csr* Device() {
        static_assert(std::is_same<T, double>::value, "the B200 path is instantiated for T = double");
        if (!device || !device->handle) {
            device = std::make_shared<PANSFEM2::B200::CsrDevice>();
            PANSFEM2::B200::Check(pf2_csr_upload(PANSFEM2::B200::Device::Context(), ROWS, indptr.data(), indices.data(), data.data(), &device->handle), "pf2_csr_upload");
        }
        return device->handle;
    }

    template<class F> friend CSR<F> ILU0(CSR<F>& _A);
    template<class F> friend std::vector<F> PreILU0(CSR<F>& _A, std::vector<F>& _b);
    template<class F> friend class LILCSR;
    template<class T1, class T2> friend const CSR<T1> operator+(const CSR<T1>& _m1, const CSR<T2>& _m2);
    template<class T1, class T2> friend const CSR<T1> operator-(const CSR<T1>& _m1, const CSR<T2>& _m2);
    template<class T1, class T2> friend const CSR<T1> operator*(const CSR<T1>& _m, T2 _a);
    template<class T1, class T2> friend const CSR<T1> operator/(const CSR<T1>& _m, T2 _a);

private:
    static int DeviceRows(pf2_csr* _d) { int r = 0; pf2_csr_info(_d, &r, nullptr); return r; }
    void SyncHost() {
        if (!host_stale) return;
        if constexpr (std::is_same<T, double>::value) {
            long long nnz = 0;
            int rows = 0;
            pf2_csr_info(device->handle, &rows, &nnz);
            std::vector<long long> ip(rows + 1);
            indices.resize(nnz); data.resize(nnz);
            PANSFEM2::B200::Check(pf2_csr_download(device->handle, ip.data(), indices.data(), data.data(), nullptr), "pf2_csr_download");
            indptr.assign(ip.begin(), ip.end());
        }
        host_stale = false;
    }
    std::vector<int> indptr;
    std::vector<int> indices;
    std::vector<T> data;
    std::shared_ptr<PANSFEM2::B200::CsrDevice> device;
    bool host_stale = false;
};

template<class T1, class T2>
inline const CSR<T1> operator+(const CSR<T1>& _m1, const CSR<T2>& _m2) {
    assert(_m1.ROWS == _m2.ROWS && _m1.COLS == _m2.COLS);
    CSR<T1> m(_m1);
    if (_m2.ROWS > 0) _m2.get(0, 0);      //  a device-adopted operand fills its host arrays first
    for (int i = 0; i < _m2.ROWS; i++) for (int k = _m2.indptr[i]; k < _m2.indptr[i + 1]; k++) m.set(i, _m2.indices[k], m.get(i, _m2.indices[k]) + _m2.data[k]);
    return m;
}
template<class T1, class T2>
inline const CSR<T1> operator-(const CSR<T1>& _m1, const CSR<T2>& _m2) {
    assert(_m1.ROWS == _m2.ROWS && _m1.COLS == _m2.COLS);
    CSR<T1> m(_m1);
    if (_m2.ROWS > 0) _m2.get(0, 0);      //  a device-adopted operand fills its host arrays first
    for (int i = 0; i < _m2.ROWS; i++) for (int k = _m2.indptr[i]; k < _m2.indptr[i + 1]; k++) m.set(i, _m2.indices[k], m.get(i, _m2.indices[k]) - _m2.data[k]);
    return m;
}
template<class T1, class T2>
inline const CSR<T1> operator*(const CSR<T1>& _m, T2 _a) {
    CSR<T1> m(_m);
    const_cast<CSR<T1>&>(m).get(0, 0);
    for (auto& v : m.data) v *= _a;
    m.device.reset();
    return m;
}
template<class T1, class T2>
inline const CSR<T2> operator*(T1 _a, const CSR<T2>& _m) { return _m*_a; }
template<class T1, class T2>
inline const CSR<T1> operator/(const CSR<T1>& _m, T2 _a) {
    CSR<T1> m(_m);
    const_cast<CSR<T1>&>(m).get(0, 0);
    for (auto& v : m.data) v /= _a;
    m.device.reset();
    return m;
}
template<class F>
inline std::ostream& operator<<(std::ostream& _out, const CSR<F>& _mat) {
    for (int i = 0; i < _mat.ROWS; i++) { for (int j = 0; j < _mat.COLS; j++) _out << _mat.get(i, j) << "\t"; _out << std::endl; }
    return _out;
}
#include "LILCSR.h"
