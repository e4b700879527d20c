//  pansfem2_b200/src/LinearAlgebra/Solvers/CG.h
//  Mirror of the Krylov solver family of the reference (src/LinearAlgebra/Solvers/CG.h, global namespace):
//      CG :124-154, ILU0 :258-284, PreILU0 :289-315, ILU0CG :320-352, GetDiagonal :398, Scaling :409, ScalingCG :420-453,
//      BiCGSTAB :159-194, BiCGSTAB2 :199-253, ILU0BiCGSTAB :357-393, ScalingBiCGSTAB :458-495
//  Same signatures and the same observable behaviour (x0 = 0, stop on ||r|| < eps*||b|| of the recursive residual,
//  "Convergence:faild" on stdout at itrmax, last iterate returned); for T = double the iterations run on the B200.
#pragma once
#include <cmath>
#include <numeric>
#include <vector>
#include <iostream>
#include "../Models/CSR.h"

namespace PANSFEM2 { namespace B200 {
    inline std::vector<double> Solve(CSR<double>& _A, int _solver, const std::vector<double>& _b, int _itrmax, double _eps, bool _report) {
        assert((int)_b.size() == _A.ROWS);
        std::vector<double> x(_b.size(), 0.0);
        int iters = 0;
        double relres = 0.0;
        const int rc = pf2_solve_host(_A.Device(), _solver, _b.data(), x.data(), _itrmax, _eps, &iters, &relres);
        Check(rc, "pf2_solve_host");
        if (rc == PF2_E_NOCONV) std::cout << "\nConvergence:faild" << std::endl;
        else if (_report) std::cout << "\tConvergence:" << iters - 1 << std::endl;      //  the reference prints the 0-based index
        return x;
    }
} }

//********************{a}-{b}********************
template<class T>
inline std::vector<T> subtract(std::vector<T> _a, std::vector<T> _b) {
    std::vector<T> v(_b.size());
    for (size_t i = 0; i < v.size(); i++) v[i] = _a[i] - _b[i];
    return v;
}
//********************{x}={x}+a{y}********************
template<class T>
inline void xexpay(std::vector<T>& _x, T _a, const std::vector<T>& _y) { for (size_t i = 0; i < _x.size(); i++) _x[i] = _x[i] + _a*_y[i]; }
//********************{x}=a{x}+{y}********************
template<class T>
inline void xeaxpy(T _a, std::vector<T>& _x, const std::vector<T>& _y) { for (size_t i = 0; i < _x.size(); i++) _x[i] = _a*_x[i] + _y[i]; }

//********************CG method********************
template<class T>
std::vector<T> CG(CSR<T>& _A, const std::vector<T>& _b, int _itrmax, T _eps) {
    return PANSFEM2::B200::Solve(_A, PF2_SOLVER_CG, _b, _itrmax, _eps, true);
}

//********************Incomplete LU(0) decomposition********************
template<class T>
CSR<T> ILU0(CSR<T>& _A) {
    CSR<T> M(_A);                       //  same pattern; values replaced by the factors computed on the device
    M.get(0, 0);
    M.device.reset();
    PANSFEM2::B200::Check(pf2_ilu0_factor(_A.Device()), "pf2_ilu0_factor");
    PANSFEM2::B200::Check(pf2_ilu0_download(_A.Device(), M.data.data()), "pf2_ilu0_download");
    return M;
}

//********************Solve with ILU(0)*******************
template<class T>
std::vector<T> PreILU0(CSR<T>& _A, std::vector<T>& _b) {
    //  _A holds the factors (unit-L strictly lower + U with diagonal, one CSR); level-scheduled sweeps on the device
    std::vector<T> v(_b.size());
    PANSFEM2::B200::Check(pf2_preilu0_host(_A.Device(), _b.data(), v.data()), "pf2_preilu0_host");
    return v;
}

//*******************ILU(0) preconditioning CG method********************
template<class T>
std::vector<T> ILU0CG(CSR<T>& _A, CSR<T>& _M, const std::vector<T>& _b, int _itrmax, T _eps) {
    (void)_M;                           //  the factors are recomputed (and cached) next to _A on the device
    return PANSFEM2::B200::Solve(_A, PF2_SOLVER_ILU0CG, _b, _itrmax, _eps, true);
}

//********************Get diagonal vector of matrix _A********************
template<class T>
std::vector<T> GetDiagonal(CSR<T>& _A) {
    std::vector<T> v(_A.ROWS);
    for (int i = 0; i < _A.ROWS; i++) v[i] = _A.get(i, i);
    return v;
}
//********************Scaling matrix********************
template<class T>
std::vector<T> Scaling(std::vector<T>& _D, std::vector<T>& _b) {
    std::vector<T> v(_D.size());
    for (size_t i = 0; i < _D.size(); i++) v[i] = _b[i]/_D[i];
    return v;
}

//********************Scaling preconditioning CG method********************
template<class T>
std::vector<T> ScalingCG(CSR<T>& _A, const std::vector<T>& _b, int _itrmax, T _eps) {
    return PANSFEM2::B200::Solve(_A, PF2_SOLVER_SCALINGCG, _b, _itrmax, _eps, false);
}

//********************BiCGSTAB method********************
template<class T>
std::vector<T> BiCGSTAB(CSR<T>& _A, std::vector<T>& _b, int _itrmax, T _eps) {
    return PANSFEM2::B200::Solve(_A, PF2_SOLVER_BICGSTAB, _b, _itrmax, _eps, false);
}

//********************BiCGSTAB2 method********************
template<class T>
std::vector<T> BiCGSTAB2(CSR<T>& _A, std::vector<T>& _b, int _itrmax, T _eps) {
    return PANSFEM2::B200::Solve(_A, PF2_SOLVER_BICGSTAB2, _b, _itrmax, _eps, false);
}

//*******************ILU(0) preconditioning BiCGSTAB method*******************
template<class T>
std::vector<T> ILU0BiCGSTAB(CSR<T>& _A, CSR<T>& _M, std::vector<T>& _b, int _itrmax, T _eps) {
    (void)_M;                           //  the factors are recomputed (and cached) next to _A on the device
    return PANSFEM2::B200::Solve(_A, PF2_SOLVER_ILU0BICGSTAB, _b, _itrmax, _eps, true);
}

//********************Scaling preconditioning BiCGSTAB method********************
template<class T>
std::vector<T> ScalingBiCGSTAB(CSR<T>& _A, std::vector<T>& _b, int _itrmax, T _eps) {
    return PANSFEM2::B200::Solve(_A, PF2_SOLVER_SCALINGBICGSTAB, _b, _itrmax, _eps, false);
}
