//  pansfem2_b200/src/Optimize/Filter/HeavisideFilter.h
//  HeavisideFilter<T> with the reference's interface (src/Optimize/Filter/HeavisideFilter.h:18-23): density filter followed by
//  the tanh projection with sharpness beta (default 1, :37,:46), sensitivities by the chain rule.  Runs on the B200.
#pragma once
#include <vector>
#include <memory>
#include "DensityFilter.h"

namespace PANSFEM2 {
    template<class T>
    class HeavisideFilter {
public:
        HeavisideFilter() : n(0), beta(1.0) {}
        ~HeavisideFilter() {}
        HeavisideFilter(int _n, std::vector<std::vector<int> > _neighbors, std::vector<std::vector<T> > _w) : n(_n), beta(1.0), device(std::make_shared<B200::FilterDevice>(PF2_FILTER_HEAVISIDE, _n, _neighbors, _w)) {}

        void UpdateBeta(T _beta) { beta = _beta; }
        std::vector<T> GetFilteredVariables(std::vector<T> _s) {
            std::vector<T> rho(n);
            B200::Check(pf2_filter_set_beta(device->handle, beta), "pf2_filter_set_beta");
            B200::Check(pf2_filter_apply_host(device->handle, _s.data(), rho.data()), "pf2_filter_apply_host");
            return rho;
        }
        std::vector<T> GetFilteredSensitivitis(std::vector<T> _s, std::vector<T> _dfdrho) {
            std::vector<T> dfds(n);
            B200::Check(pf2_filter_set_beta(device->handle, beta), "pf2_filter_set_beta");
            B200::Check(pf2_filter_sens_host(device->handle, _s.data(), _dfdrho.data(), dfds.data()), "pf2_filter_sens_host");
            return dfds;
        }
        pf2_filter* Device() const { B200::Check(pf2_filter_set_beta(device->handle, beta), "pf2_filter_set_beta"); return device->handle; }
private:
        const int n;
        T beta;
        std::shared_ptr<B200::FilterDevice> device;
    };
}
