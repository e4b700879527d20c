//  pansfem2_b200/src/Optimize/Solver/OC.h
//  OC<T> with the reference's interface (src/Optimize/Solver/OC.h:16-107).
//  UpdateVariables<F>(..., F gkp1) keeps the opaque constraint functor: candidates are evaluated on the device, the
//  functor is called on the host once per bisection step exactly as OC.h:82-99 does.  The overload taking a
//  B200::FilteredVolumeConstraint runs the WHOLE bisection on the device (what the drivers' lambda computes,
//  sample_optimize_density_oc.cpp:198-207), which is the hot path.
#pragma once
#include <vector>
#include <memory>
#include <iostream>
#include "../../B200/Device.h"

namespace PANSFEM2 {
    namespace B200 {
        //  g(x) = scale*sum_i filter(x)_i/(limit*n) - scale, with `filter` any of the mirrored filter classes
        struct FilteredVolumeConstraint {
            pf2_filter* filter;
            double limit, scale;
            template<class FILTER>
            FilteredVolumeConstraint(const FILTER& _filter, double _limit, double _scale = 1.0) : filter(_filter.Device()), limit(_limit), scale(_scale) {}
        };
        struct OcDevice {
            pf2_oc* handle;
            OcDevice(int n, double iota, double lmin, double lmax, double leps, double move) : handle(nullptr) { Check(pf2_oc_create(Device::Context(), n, iota, lmin, lmax, leps, move, &handle), "pf2_oc_create"); }
            ~OcDevice() { if (handle) pf2_oc_destroy(handle); }
        };
    }

    template<class T>
    class OC {
public:
        OC(int _n, T _iota, T _lambdamin, T _lambdamax, T _lambdaeps, T _movelimit, const std::vector<T>& _xmin, const std::vector<T>& _xmax)
            : n(_n), xmin(_xmin), xmax(_xmax), lambdamin(_lambdamin), lambdamax(_lambdamax), lambdaeps(_lambdaeps),
              device(std::make_shared<B200::OcDevice>(_n, _iota, _lambdamin, _lambdamax, _lambdaeps, _movelimit)) {}
        ~OC() {}

        bool IsConvergence(T _currentf0) {
            int converged = 0;
            B200::Check(pf2_oc_is_convergence(device->handle, _currentf0, &converged), "pf2_oc_is_convergence");
            return converged != 0;
        }

        template<class F>
        void UpdateVariables(std::vector<T>& _xk, T _f, std::vector<T> _dfdx, T _g, std::vector<T> _dgdx, F _gkp1) {
            (void)_g;
            T lambda0 = lambdamin, lambda1 = lambdamax, lambda = T();
            std::vector<T> xkp1(n, T());
            while ((lambda1 - lambda0)/(lambda1 + lambda0) > lambdaeps) {
                lambda = 0.5*(lambda1 + lambda0);
                B200::Check(pf2_oc_candidate_host(device->handle, _xk.data(), _dfdx.data(), _dgdx.data(), lambda, xkp1.data()), "pf2_oc_candidate_host");
                if (_gkp1(xkp1) > T()) lambda0 = lambda; else lambda1 = lambda;
            }
            std::cout << lambda;
            B200::Check(pf2_oc_commit(device->handle, _f), "pf2_oc_commit");
            _xk = xkp1;
        }

        //  device-resident bisection (hot path)
        void UpdateVariables(std::vector<T>& _xk, T _f, std::vector<T> _dfdx, T _g, std::vector<T> _dgdx, const B200::FilteredVolumeConstraint& _gkp1) {
            (void)_g;
            B200::Buffer x, df, dg;
            x.Upload(_xk); df.Upload(_dfdx); dg.Upload(_dgdx);
            double lambda = 0.0;
            int steps = 0;
            B200::Check(pf2_oc_update(device->handle, _gkp1.filter, _gkp1.limit, _gkp1.scale, x.Get(), _f, df.Get(), dg.Get(), &steps, &lambda), "pf2_oc_update");
            std::cout << lambda;
            _xk = x.Download();
        }
private:
        const int n;
        std::vector<T> xmin, xmax;      //  stored but unused, as in the reference (OC.h:86-91 clamps to literal 0 and 1)
        T lambdamin, lambdamax, lambdaeps;
        std::shared_ptr<B200::OcDevice> device;
    };
}
