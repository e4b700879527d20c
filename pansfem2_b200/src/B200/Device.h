//  pansfem2_b200/src/B200/Device.h
//  Bridge between the header mirror of PANSFEM2's template API and the C ABI of libpansfem2_b200.so
//  (include/pansfem2_b200.h).  Not part of the reference: everything here lives in PANSFEM2::B200.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../../include/pansfem2_b200.h"

namespace PANSFEM2 {
namespace B200 {
    //  The reference signals misuse with assert() (abort); a failing C-ABI call is treated the same way,
    //  except PF2_E_NOCONV which the solvers report like the reference ("Convergence:faild") and carry on.
    inline void Check(int _rc, const char* _what) {
        if (_rc != PF2_OK && _rc != PF2_E_NOCONV) {
            std::fprintf(stderr, "pansfem2_b200: %s failed (%d): %s\n", _what, _rc, pf2_last_error());
            std::abort();
        }
    }

    //  One context per process on device $PF2_DEVICE (default 0); created on first use.
    class Device {
public:
        static pf2_ctx* Context() {
            static Device instance;
            return instance.ctx;
        }
private:
        Device() : ctx(nullptr) {
            const char* env = std::getenv("PF2_DEVICE");
            Check(pf2_ctx_create(env ? std::atoi(env) : 0, nullptr, &ctx), "pf2_ctx_create");
        }
        ~Device() { pf2_ctx_destroy(ctx); }
        pf2_ctx* ctx;
    };

    //  RAII device buffer of doubles
    class Buffer {
public:
        Buffer() : ptr(nullptr), count(0) {}
        explicit Buffer(size_t _count) : ptr(nullptr), count(0) { Resize(_count); }
        Buffer(const Buffer&) = delete;
        Buffer& operator=(const Buffer&) = delete;
        ~Buffer() { Release(); }
        void Resize(size_t _count) {
            if (_count == count) return;
            Release();
            void* p = nullptr;
            Check(pf2_malloc(Device::Context(), _count * sizeof(double), &p), "pf2_malloc");
            ptr = static_cast<double*>(p);
            count = _count;
        }
        void Upload(const std::vector<double>& _v) {
            Resize(_v.size());
            Check(pf2_memcpy_h2d(Device::Context(), ptr, _v.data(), _v.size() * sizeof(double)), "pf2_memcpy_h2d");
        }
        std::vector<double> Download() const {
            std::vector<double> v(count);
            Check(pf2_memcpy_d2h(Device::Context(), v.data(), ptr, count * sizeof(double)), "pf2_memcpy_d2h");
            return v;
        }
        double* Get() const { return ptr; }
        size_t Size() const { return count; }
private:
        void Release() { if (ptr) { pf2_free(Device::Context(), ptr); ptr = nullptr; count = 0; } }
        double* ptr;
        size_t count;
    };
}
}
