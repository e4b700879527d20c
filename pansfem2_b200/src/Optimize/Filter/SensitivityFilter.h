//  pansfem2_b200/src/Optimize/Filter/SensitivityFilter.h
//  SensitivityFilter<T> (Sigmund) and SensitivityFilter2<T> (Borrvall) with the reference's interface
//  (src/Optimize/Filter/SensitivityFilter.h:16-23, 60-67): one operation, GetFilteredSensitivitis(s, dfds), which runs
//  on the B200 over the flattened neighbour lists (pf2_filter_sens with PF2_FILTER_SENS_SIGMUND / _BORRVALL).
#pragma once
#include <vector>
#include <memory>
#include "DensityFilter.h"

namespace PANSFEM2 {
    template<class T>
    class SensitivityFilter {
public:
        SensitivityFilter(int _n, std::vector<std::vector<int> > _neighbors, std::vector<std::vector<T> > _w) : n(_n), device(std::make_shared<B200::FilterDevice>(PF2_FILTER_SENS_SIGMUND, _n, _neighbors, _w)) {}
        ~SensitivityFilter() {}

        std::vector<T> GetFilteredSensitivitis(std::vector<T> _s, std::vector<T> _dfds) {
            std::vector<T> out(n);
            B200::Check(pf2_filter_sens_host(device->handle, _s.data(), _dfds.data(), out.data()), "pf2_filter_sens_host");
            return out;
        }
private:
        const int n;
        std::shared_ptr<B200::FilterDevice> device;
    };

    template<class T>
    class SensitivityFilter2 {
public:
        SensitivityFilter2(int _n, std::vector<std::vector<int> > _neighbors, std::vector<std::vector<T> > _w) : n(_n), device(std::make_shared<B200::FilterDevice>(PF2_FILTER_SENS_BORRVALL, _n, _neighbors, _w)) {}
        ~SensitivityFilter2() {}

        std::vector<T> GetFilteredSensitivitis(std::vector<T> _s, std::vector<T> _dfds) {
            std::vector<T> out(n);
            B200::Check(pf2_filter_sens_host(device->handle, _s.data(), _dfds.data(), out.data()), "pf2_filter_sens_host");
            return out;
        }
private:
        const int n;
        std::shared_ptr<B200::FilterDevice> device;
    };
}
