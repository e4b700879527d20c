//  pansfem2_b200/src/Optimize/Solver/CONLIN.h
//  CONLIN<T> with the reference's interface (src/Optimize/Solver/CONLIN.h:18-26): ctor, SetParameters(move, epsvalue),
//  IsConvergence, UpdateVariables(xk, f, dfdx, g, dgdx).  The convex-linearised subproblem (p*x + q/x per variable) is
//  solved by the same device-resident primal-dual interior-point driver as MMA (pf2_mma_update on a handle made by
//  pf2_conlin_create); only the approximation terms differ, so the handle type is shared.
#pragma once
#include <vector>
#include <memory>
#include <cassert>
#include "MMA.h"

namespace PANSFEM2 {
    template<class T>
    class CONLIN {
public:
        CONLIN(int _n, int _m, T _a0, std::vector<T> _a, std::vector<T> _c, std::vector<T> _d, const std::vector<T>& _xmin, const std::vector<T>& _xmax)
            : n(_n), m(_m), device(std::make_shared<B200::MmaDevice>()) {
            assert((int)_a.size() == _m && (int)_c.size() == _m && (int)_d.size() == _m && (int)_xmin.size() == _n && (int)_xmax.size() == _n);
            B200::Check(pf2_conlin_create(B200::Device::Context(), _n, _m, _a0, _a.data(), _c.data(), _d.data(), _xmin.data(), _xmax.data(), &device->handle), "pf2_conlin_create");
        }
        ~CONLIN() {}

        void SetParameters(T _move, T _epsvalue) {
            B200::Check(pf2_conlin_set_parameters(device->handle, _move, _epsvalue), "pf2_conlin_set_parameters");
        }
        bool IsConvergence(T _currentf0) {
            int converged = 0;
            B200::Check(pf2_mma_is_convergence(device->handle, _currentf0, &converged), "pf2_mma_is_convergence");
            return converged != 0;
        }
        void UpdateVariables(std::vector<T>& _xk, T _f, std::vector<T> _dfdx, std::vector<T> _g, std::vector<std::vector<T> > _dgdx) {
            assert((int)_g.size() == m && (int)_dgdx.size() == m);
            std::vector<T> flat;
            flat.reserve((size_t)m*n);
            for (const auto& row : _dgdx) flat.insert(flat.end(), row.begin(), row.end());
            B200::Buffer x, df, dg;
            x.Upload(_xk); df.Upload(_dfdx); dg.Upload(flat);
            int steps = 0;
            B200::Check(pf2_mma_update(device->handle, x.Get(), _f, df.Get(), _g.data(), dg.Get(), &steps), "pf2_mma_update");
            _xk = x.Download();
        }
private:
        const int n, m;
        std::shared_ptr<B200::MmaDevice> device;
    };
}
