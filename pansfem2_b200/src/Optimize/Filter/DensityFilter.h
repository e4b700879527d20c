//  pansfem2_b200/src/Optimize/Filter/DensityFilter.h
//  DensityFilter<T> with the reference's interface (src/Optimize/Filter/DensityFilter.h:18-23); the ragged neighbour
//  lists are flattened once and both operations run on the B200 (pf2_filter_apply / pf2_filter_sens).
#pragma once
#include <vector>
#include <memory>
#include "../../B200/Device.h"

namespace PANSFEM2 {
    namespace B200 {
        struct FilterDevice {
            pf2_filter* handle;
            FilterDevice(int _kind, int _n, const std::vector<std::vector<int> >& _neighbors, const std::vector<std::vector<double> >& _w) : handle(nullptr) {
                std::vector<long long> rowptr(_n + 1, 0);
                for (int i = 0; i < _n; i++) rowptr[i + 1] = rowptr[i] + (long long)_neighbors[i].size();
                std::vector<int> nbr; std::vector<double> w;
                nbr.reserve(rowptr[_n]); w.reserve(rowptr[_n]);
                for (int i = 0; i < _n; i++) { nbr.insert(nbr.end(), _neighbors[i].begin(), _neighbors[i].end()); w.insert(w.end(), _w[i].begin(), _w[i].end()); }
                Check(pf2_filter_create(Device::Context(), _kind, _n, rowptr.data(), nbr.data(), w.data(), &handle), "pf2_filter_create");
            }
            ~FilterDevice() { if (handle) pf2_filter_destroy(handle); }
        };
    }

    template<class T>
    class DensityFilter {
public:
        DensityFilter() : n(0) {}
        ~DensityFilter() {}
        DensityFilter(int _n, std::vector<std::vector<int> > _neighbors, std::vector<std::vector<T> > _w) : n(_n), device(std::make_shared<B200::FilterDevice>(PF2_FILTER_DENSITY, _n, _neighbors, _w)) {}

        std::vector<T> GetFilteredVariables(std::vector<T> _s) {
            std::vector<T> rho(n);
            B200::Check(pf2_filter_apply_host(device->handle, _s.data(), rho.data()), "pf2_filter_apply_host");
            return rho;
        }
        std::vector<T> GetFilteredSensitivitis(std::vector<T> _s, std::vector<T> _dfdrho) {
            std::vector<T> dfds(n);
            B200::Check(pf2_filter_sens_host(device->handle, _s.data(), _dfdrho.data(), dfds.data()), "pf2_filter_sens_host");
            return dfds;
        }
        pf2_filter* Device() const { return device->handle; }      //  for the batched path
private:
        const int n;
        std::shared_ptr<B200::FilterDevice> device;
    };
}
