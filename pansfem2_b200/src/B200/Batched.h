//  pansfem2_b200/src/B200/Batched.h
//  The batched, device-resident entry points of the hot path, carrying the reference's template selection as tags.
//  A driver that used to loop `for element: Eq<T,SF,IC>(Ke,...); Assembling(K,F,u,Ke,...)` calls AssembleBatched<Eq<SF,IC>> once;
//  a driver that used to run the whole design iteration on the host (sample_optimize_density_oc.cpp:83-208) steps a DesignLoop.
#pragma once
#include <vector>
#include <cassert>
#include <iostream>
#include <utility>
#include <memory>
#include "ElementSelect.h"
#include "../LinearAlgebra/Models/CSR.h"

namespace PANSFEM2 { namespace B200 {
    //  equation tags == the reference's element routines with their template arguments
    template<template<class>class SF, template<class>class IC>
    struct PlaneStrainStiffnessTag { static const int eq = EqCode<PF2_PHYS_PLANESTRAIN, SF, IC>::value; static const int ndof = 2; };
    template<template<class>class SF, template<class>class IC>
    struct PlaneStressStiffnessTag { static const int eq = EqCode<PF2_PHYS_PLANESTRESS, SF, IC>::value; static const int ndof = 2; };
    template<template<class>class SF, template<class>class ICV, template<class>class ICD>
    struct PlaneStrainStiffnessSRITag { static const int eq = EqCodeSRI<SF, ICV, ICD>::value; static const int ndof = 2; };
    template<template<class>class SF, template<class>class IC>
    struct SolidLinearIsotropicElasticTag { static const int eq = EqCode<PF2_PHYS_SOLID, SF, IC>::value; static const int ndof = 3; };
    template<template<class>class SF, template<class>class IC>
    struct HeatTransferTag { static const int eq = EqCode<PF2_PHYS_HEAT, SF, IC>::value; static const int ndof = 1; };
    //  the Advection.h routines a driver sums per element (TERMS = PF2_ADV_ADVECTION | PF2_ADV_DIFFUSION | ...)
    template<template<class>class SF, template<class>class IC, int TERMS>
    struct AdvectionDiffusionTag { static const int eq = EqCode<PF2_PHYS_ADVDIFF, SF, IC>::value | (TERMS << 24); static const int ndof = 1; };

    typedef std::vector<std::pair<std::pair<int, int>, double> > BcList;
    inline void SplitBc(const BcList& _bc, std::vector<int>& _node, std::vector<int>& _dof, std::vector<double>& _val) {
        for (const auto& b : _bc) { _node.push_back(b.first.first); _dof.push_back(b.first.second); _val.push_back(b.second); }
    }

    //  what a load functor returns: Vector<double> (forces) or a scalar (HeatTransferSurfaceFlux's flux)
    inline Vector<double> LoadValue(const Vector<double>& _v, int) { return _v; }
    inline Vector<double> LoadValue(double _v, int) { Vector<double> f = { _v }; return f; }

    //  mesh + Dirichlet numbering + symbolic CSR pattern on the device: built once, reused every design iteration
    class Model {
public:
        Model(std::vector<Vector<double> >& _x, const std::vector<std::vector<int> >& _elements, int _ndof, const BcList& _ufixed)
            : mesh(nullptr), dofmap(nullptr), pattern(nullptr), nnode((int)_x.size()), nelem((int)_elements.size()), ndof(_ndof), KDEGREE(0) {
            const int dim = _x[0].SIZE(), npe = (int)_elements[0].size();
            std::vector<double> coords((size_t)nnode*dim);
            for (int i = 0; i < nnode; i++) for (int d = 0; d < dim; d++) coords[(size_t)i*dim + d] = _x[i](d);
            std::vector<int> conn((size_t)nelem*npe);
            for (int e = 0; e < nelem; e++) for (int a = 0; a < npe; a++) conn[(size_t)e*npe + a] = _elements[e][a];
            Check(pf2_mesh_create(Device::Context(), dim, nnode, coords.data(), npe, nelem, conn.data(), &mesh), "pf2_mesh_create");
            std::vector<int> fn, fd; std::vector<double> fv;
            SplitBc(_ufixed, fn, fd, fv);
            Check(pf2_dofmap_create(Device::Context(), nnode, _ndof, (int)fn.size(), fn.data(), fd.data(), fv.data(), &KDEGREE, &dofmap), "pf2_dofmap_create");
            Check(pf2_csr_pattern(Device::Context(), mesh, dofmap, &pattern), "pf2_csr_pattern");
        }
        ~Model() { pf2_csr_destroy(pattern); pf2_dofmap_destroy(dofmap); pf2_mesh_destroy(mesh); }
        Model(const Model&) = delete;
        Model& operator=(const Model&) = delete;

        //  nodetoglobal exactly as SetDirichlet + Renumbering produce it
        std::vector<std::vector<int> > NodeToGlobal() const {
            std::vector<int> flat((size_t)nnode*ndof);
            Check(pf2_dofmap_get(dofmap, flat.data()), "pf2_dofmap_get");
            std::vector<std::vector<int> > n2g(nnode, std::vector<int>(ndof));
            for (int i = 0; i < nnode; i++) for (int d = 0; d < ndof; d++) n2g[i][d] = flat[(size_t)i*ndof + d];
            return n2g;
        }
        pf2_mesh* mesh;
        pf2_dofmap* dofmap;
        pf2_csr* pattern;
        const int nnode, nelem, ndof;
        int KDEGREE;
    };

    //  element loop + Assembling(K,F,u,Ke,...) + Assembling(F,q,...) of the drivers in one call; returns K's device handle
    //  (owned by the model) and fills F.  _modulus[e] is the per-element E (or conductivity); _V, _t as the element routine takes.
    template<class EQTAG>
    inline pf2_csr* AssembleBatched(Model& _model, const std::vector<double>& _modulus, double _V, double _t, const BcList& _qfixed, std::vector<double>& _F) {
        Buffer E;
        E.Upload(_modulus);
        std::vector<int> ln, ld; std::vector<double> lv;
        SplitBc(_qfixed, ln, ld, lv);
        const double params[5] = { 0.0, 0.0, _V, 1.0, _t };
        Check(pf2_assemble(_model.pattern, _model.mesh, _model.dofmap, EQTAG::eq, E.Get(), nullptr, params, (int)ln.size(), ln.data(), ld.data(), lv.data()), "pf2_assemble");
        _F.resize(_model.KDEGREE);
        Check(pf2_csr_download(_model.pattern, nullptr, nullptr, nullptr, _F.data()), "pf2_csr_download");
        return _model.pattern;
    }

    //  The load-vector loops of the drivers (sample_planestrain.cpp:41-61) in one call each: PlaneStrainBodyForce / PlaneStressBodyForce
    //  over area elements, PlaneStrainSurfaceForce / PlaneStressSurfaceForce / HeatTransferSurfaceFlux over edges, each followed by
    //  Assembling(F, Fe, ...).  _f is the reference's functor f(x) -> Vector<double> (size = dofs per node); it is evaluated on the host at
    //  every integration point of the batch (the device hands the points out), integration and assembly run on the device, and the result
    //  is ADDED to _F (size KDEGREE of the model) like the reference's Assembling does.
    template<template<class>class SF, template<class>class IC, class F>
    inline void AssembleLoadVector(Model& _model, const std::vector<std::vector<int> >& _elements, F _f, double _t, std::vector<double>& _F) {
        static_assert(ShapeCode<SF<double> >::value > 0 && QuadCode<IC<double> >::value > 0, "pansfem2_b200: no load-vector kernel for this shape / rule");
        static_assert(ShapeCode<SF<double> >::domain == QuadCode<IC<double> >::domain, "pansfem2_b200: integration rule does not belong to the shape function's reference domain");
        static_assert(ShapeCode<SF<double> >::domain == 0 || ShapeCode<SF<double> >::domain == 1 || ShapeCode<SF<double> >::domain == 4, "pansfem2_b200: load vectors are 2-D");
        if (_elements.empty()) return;
        const int nel = (int)_elements.size(), npe = (int)_elements[0].size(), ng = IC<double>::N, ndof = _model.ndof;
        //  a second element list over the model's nodes: the loaded edges / areas
        std::vector<int> conn((size_t)nel*npe);
        for (int e = 0; e < nel; e++) for (int a = 0; a < npe; a++) conn[(size_t)e*npe + a] = _elements[e][a];
        pf2_mesh* carrier = nullptr;
        Check(pf2_mesh_create_on_nodes(_model.mesh, npe, nel, conn.data(), &carrier), "pf2_mesh_create_on_nodes");
        Buffer xg((size_t)nel*ng*2), fg, Fd;
        Check(pf2_integration_points(carrier, ShapeCode<SF<double> >::value, QuadCode<IC<double> >::value, xg.Get()), "pf2_integration_points");
        const std::vector<double> x = xg.Download();
        std::vector<double> f((size_t)nel*ng*ndof);
        for (size_t q = 0; q < (size_t)nel*ng; q++) {
            Vector<double> xq = { x[2*q], x[2*q + 1] };
            const Vector<double> fq = LoadValue(_f(xq), ndof);
            for (int i = 0; i < ndof; i++) f[q*ndof + i] = fq(i);
        }
        fg.Upload(f);
        _F.resize(_model.KDEGREE, 0.0);
        Fd.Upload(_F);
        Check(pf2_load_vector(carrier, _model.dofmap, ShapeCode<SF<double> >::value, QuadCode<IC<double> >::value, nullptr, fg.Get(), _t, Fd.Get()), "pf2_load_vector");
        _F = Fd.Download();
        pf2_mesh_destroy(carrier);
    }

    //  solve K u = F on the device with K still resident; kind = PF2_SOLVER_*
    inline std::vector<double> SolveResident(pf2_csr* _K, int _solver, const std::vector<double>& _F, int _itrmax, double _eps, int* _iters = nullptr) {
        std::vector<double> x(_F.size());
        double relres = 0.0;
        int iters = 0;
        const int rc = pf2_solve_host(_K, _solver, _F.data(), x.data(), _itrmax, _eps, &iters, &relres);
        Check(rc, "pf2_solve_host");
        if (rc == PF2_E_NOCONV) std::cout << "\nConvergence:faild" << std::endl;
        if (_iters) *_iters = iters;
        return x;
    }

    //  the SIMP design loop of sample/optimize/sample_optimize_density_{oc,mma}.cpp with every field resident on the device
    struct SimpParameters {
        double E0 = 0.0001, E1 = 210000.0, Poisson = 0.3, p = 3.0, weightlimit = 0.5, scale0 = 1.0e5, scale1 = 1.0, thickness = 1.0;
        double beta0 = 0.5; int beta_period = 40; int cg_itrmax = 100000; double cg_eps = 1.0e-10;
    };
    struct IterationReport { double f, g; bool converged; int cg_iterations; double cg_relres; int optimizer_steps; double beta; int k; };

    template<class EQTAG>
    class DesignLoop {
public:
        //  _filter: a mirrored DensityFilter<double> / HeavisideFilter<double>; _optimizer: PF2_OPT_OC with {iota,lmin,lmax,leps,move}
        //  or PF2_OPT_MMA with {raa0,albefa,move,asyinit,asydecr,asyincr,epsvalue,a0,a,c,d,xmin,xmax}
        //  or PF2_OPT_CONLIN with {move,epsvalue,a0,a,c,d,xmin,xmax}
        template<class FILTER>
        DesignLoop(Model& _model, const FILTER& _filter, int _optimizer, const std::vector<double>& _optp, const SimpParameters& _prm, const BcList& _qfixed, const std::vector<double>& _s0)
            : model(_model), handle(nullptr) {
            std::vector<int> ln, ld; std::vector<double> lv;
            SplitBc(_qfixed, ln, ld, lv);
            const double params[12] = { _prm.E0, _prm.E1, _prm.Poisson, _prm.p, _prm.weightlimit, _prm.scale0, _prm.scale1, _prm.thickness,
                                        _prm.beta0, (double)_prm.beta_period, (double)_prm.cg_itrmax, _prm.cg_eps };
            Check(pf2_simp_create(Device::Context(), model.mesh, model.dofmap, model.pattern, _filter.Device(), EQTAG::eq, _optimizer, _optp.data(), params,
                                  (int)ln.size(), ln.data(), ld.data(), lv.data(), &handle), "pf2_simp_create");
            Check(pf2_simp_set_design(handle, _s0.data()), "pf2_simp_set_design");
        }
        ~DesignLoop() { pf2_simp_destroy(handle); }
        DesignLoop(const DesignLoop&) = delete;

        IterationReport Iterate(bool _checkconvergence = true) {
            double st[8];
            Check(pf2_simp_iterate(handle, _checkconvergence ? 1 : 0, st), "pf2_simp_iterate");
            return IterationReport{ st[0], st[1], st[2] != 0.0, (int)st[3], st[4], (int)st[5], st[6], (int)st[7] };
        }
        //  fields of the last iteration in the reference's containers (for the VTK dump)
        void Get(std::vector<double>& _s, std::vector<double>& _rho, std::vector<Vector<double> >& _u, std::vector<Vector<double> >& _r) {
            _s.resize(model.nelem); _rho.resize(model.nelem);
            std::vector<double> u((size_t)model.nnode*model.ndof), r((size_t)model.nnode*model.ndof);
            Check(pf2_simp_get(handle, _s.data(), _rho.data(), u.data(), r.data()), "pf2_simp_get");
            _u.assign(model.nnode, Vector<double>(model.ndof)); _r.assign(model.nnode, Vector<double>(model.ndof));
            for (int i = 0; i < model.nnode; i++) for (int d = 0; d < model.ndof; d++) { _u[i](d) = u[(size_t)i*model.ndof + d]; _r[i](d) = r[(size_t)i*model.ndof + d]; }
        }
        //  the drivers' per-iteration VTK file (sample_optimize_density_oc.cpp:175-184) straight from the device-resident state
        void ExportVTK(const std::string& _path, int _celltype, bool _withreactions = true) {
            Check(pf2_simp_export_vtk(handle, _path.c_str(), _celltype, _withreactions ? 1 : 0), "pf2_simp_export_vtk");
        }
        //  back to iteration 0 with the design _s0 (optimiser history, Heaviside beta and the warm-start state included)
        void Reset(const std::vector<double>& _s0) { Check(pf2_simp_reset(handle, _s0.data()), "pf2_simp_reset"); }
        //  every displacement solve starts from the previous iteration's u instead of 0 (the reference always starts from 0, CG.h:423): same
        //  stopping rule, same converged solution to the solver tolerance, fewer ScalingCG iterations
        void SetWarmStart(bool _on) { Check(pf2_simp_set_warm_start(handle, _on ? 1 : 0), "pf2_simp_set_warm_start"); }
private:
        Model& model;
        pf2_simp* handle;
    };

    //  the level-set design loop of sample/optimize/sample_optimize_levelset.cpp with every field resident on the device
    struct LevelSetParameters {
        double Vmax = 0.5, tau = 2.0e-4, E0 = 1.0, Emin = 1.0e-4, nu = 0.3, nvol = 100, dt = 0.1, d = -0.02, p = 4.0;
        int tmax = 200;
    };
    struct LevelSetReport { double objective, volume, lambda; bool converged; int cg_iterations; double cg_relres; int cg_iterations_phi; int t; };

    class LevelSetLoop {
public:
        //  _model: Q4 mesh with the 2-dof displacement numbering; _phifixed: nodes where phi is held at 0 (GenerateFixedlist({ 0 }, ...))
        LevelSetLoop(Model& _model, const BcList& _qfixed, const BcList& _phifixed, const LevelSetParameters& _prm) : model(_model), handle(nullptr) {
            std::vector<int> ln, ld, pn, pd; std::vector<double> lv, pv;
            SplitBc(_qfixed, ln, ld, lv);
            SplitBc(_phifixed, pn, pd, pv);
            const double prm[9] = { _prm.Vmax, _prm.tau, _prm.E0, _prm.Emin, _prm.nu, _prm.nvol, _prm.dt, _prm.d, _prm.p };
            Check(pf2_levelset_create(Device::Context(), model.mesh, model.dofmap, model.pattern, (int)pn.size(), pn.data(), prm, _prm.tmax,
                                      (int)ln.size(), ln.data(), ld.data(), lv.data(), &handle), "pf2_levelset_create");
        }
        ~LevelSetLoop() { pf2_levelset_destroy(handle); }
        LevelSetLoop(const LevelSetLoop&) = delete;

        LevelSetReport Iterate(bool _checkconvergence = true) {
            double st[8];
            Check(pf2_levelset_iterate(handle, _checkconvergence ? 1 : 0, st), "pf2_levelset_iterate");
            return LevelSetReport{ st[0], st[1], st[2], st[3] != 0.0, (int)st[4], st[5], (int)st[6], (int)st[7] };
        }
        void Get(std::vector<Vector<double> >& _phi, std::vector<double>& _str, std::vector<Vector<double> >& _u) {
            std::vector<double> phi(model.nnode), u((size_t)model.nnode*2);
            _str.resize(model.nelem);
            Check(pf2_levelset_get(handle, phi.data(), _str.data(), u.data()), "pf2_levelset_get");
            _phi.assign(model.nnode, Vector<double>(1)); _u.assign(model.nnode, Vector<double>(2));
            for (int i = 0; i < model.nnode; i++) { _phi[i](0) = phi[i]; _u[i](0) = u[2*(size_t)i]; _u[i](1) = u[2*(size_t)i + 1]; }
        }
private:
        Model& model;
        pf2_levelset* handle;
    };

    //  The scalar advection-diffusion problem of sample/advection with the field resident on the device.
    //      steady (sample_advectiondiffusion_static.cpp:42-58):   Solve() assembles K = A + D + AS [+ SC], F = -K T_fixed and solves
    //      transient (sample_advectiondiffusion_dynamic.cpp:44-75): Step(dt, theta) assembles K = (M + MS)/dt + theta (A + D + AS),
    //          F = ((M + MS)/dt - (1 - theta)(A + D + AS)) T - K T_fixed, solves and writes the new field back (Disassembling)
    //  one assembly launch + one BiCGSTAB solve per call; nothing returns to the host until Get().
    template<class EQTAG>
    class AdvectionDiffusion {
public:
        //  _model: 1-dof numbering; _velocity: one Vector per element, or a single Vector for a uniform field; _T0: initial nodal field
        AdvectionDiffusion(Model& _model, const std::vector<Vector<double> >& _velocity, double _k, std::vector<Vector<double> >& _T0)
            : model(_model), k(_k), ax(0.0), ay(0.0), uniform(_velocity.size() == 1), T((size_t)_model.nnode), x((size_t)_model.KDEGREE) {
            static_assert(EQTAG::ndof == 1, "advection-diffusion is a scalar problem");
            assert(model.ndof == 1);
            assert(uniform || (int)_velocity.size() == model.nelem);
            std::vector<Vector<double> >& vel = const_cast<std::vector<Vector<double> >&>(_velocity);     //  Vector<T>::operator() is non-const in the reference
            if (uniform) { ax = vel[0](0); ay = vel[0](1); }
            else {
                std::vector<double> v((size_t)model.nelem*2);
                for (int e = 0; e < model.nelem; e++) { v[2*(size_t)e] = vel[e](0); v[2*(size_t)e + 1] = vel[e](1); }
                velocity.Upload(v);
            }
            std::vector<double> t0((size_t)model.nnode);
            for (int i = 0; i < model.nnode; i++) t0[i] = _T0[i](0);
            T.Upload(t0);
        }
        AdvectionDiffusion(const AdvectionDiffusion&) = delete;

        //  returns the Krylov iteration count; _solver = PF2_SOLVER_BICGSTAB / _BICGSTAB2 / _SCALINGBICGSTAB / _ILU0BICGSTAB
        int Step(double _dt, double _theta, int _solver = PF2_SOLVER_BICGSTAB, int _itrmax = 100000, double _eps = 1.0e-10) {
            const double prm[6] = { ax, ay, k, 1.0/_dt, _theta, 1.0 - _theta };
            return Advance(prm, true, _solver, _itrmax, _eps);
        }
        int Solve(int _solver = PF2_SOLVER_BICGSTAB, int _itrmax = 100000, double _eps = 1.0e-10) {
            const double prm[6] = { ax, ay, k, 0.0, 1.0, 0.0 };
            return Advance(prm, false, _solver, _itrmax, _eps);
        }
        void Get(std::vector<Vector<double> >& _T) {
            std::vector<double> t = T.Download();
            _T.assign(model.nnode, Vector<double>(1));
            for (int i = 0; i < model.nnode; i++) _T[i](0) = t[i];
        }
private:
        int Advance(const double _prm[6], bool _transient, int _solver, int _itrmax, double _eps) {
            Check(pf2_advdiff_assemble(model.pattern, model.mesh, model.dofmap, EQTAG::eq, uniform ? nullptr : velocity.Get(), _prm, _transient ? T.Get() : nullptr), "pf2_advdiff_assemble");
            double* F = nullptr;
            Check(pf2_csr_device_F(model.pattern, &F), "pf2_csr_device_F");
            int iters = 0;
            double relres = 0.0;
            const int rc = pf2_solve(model.pattern, _solver, F, x.Get(), _itrmax, _eps, &iters, &relres);
            Check(rc, "pf2_solve");
            if (rc == PF2_E_NOCONV) std::cout << "\nConvergence:faild" << std::endl;
            Check(pf2_disassemble(model.dofmap, x.Get(), T.Get()), "pf2_disassemble");
            return iters;
        }
        Model& model;
        double k, ax, ay;
        bool uniform;
        Buffer velocity, T, x;
    };
} }
