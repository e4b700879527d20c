// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// oracle/ref_shim.cpp : a thin extern "C" shim around the UNMODIFIED reference headers.
// It #includes the headers where they lie under /root/reference/src (path given with
// -I at build time, see oracle/Makefile); no reference source is copied into this repo.
// The shared object it builds (oracle/_ref/libpf2ref.so) is used
//   * by tests/ as the live oracle for the CUDA path and for the C restatement
//     (oracle/pf2_oracle.c),
//   * by bench.py's cpu_baseline / --impl reference legs as the timed CPU reference.
//
// Every entry point names the reference routine it drives (file:line relative to
// /root/reference).

// CSR<T>/LILCSR<T> keep indptr/indices/data private (src/LinearAlgebra/Models/CSR.h:72-74);
// the shim must read them to hand fixtures to the tests.
#include <vector>
#include <utility>
#include <algorithm>
#include <numeric>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <chrono>
#include <iostream>
#include <fstream>
#include <sstream>
#include <string>
#include <cassert>
#include <omp.h>
#define private public
#include "LinearAlgebra/Models/Vector.h"
#include "LinearAlgebra/Models/Matrix.h"
#include "LinearAlgebra/Models/LILCSR.h"
#include "LinearAlgebra/Models/CSR.h"
#undef private
#include "LinearAlgebra/Solvers/CG.h"
#include "FEM/Equation/PlaneStrain.h"
#include "FEM/Equation/PlaneStress.h"
#include "FEM/Equation/Solid.h"
#include "FEM/Equation/HeatTransfer.h"
#include "FEM/Equation/Advection.h"
#include "FEM/Equation/Homogenization.h"
#include "FEM/Equation/ReactionDiffusion.h"
#include "FEM/Equation/General.h"
#include "FEM/Controller/ShapeFunction.h"
#include "FEM/Controller/GaussIntegration.h"
#include "FEM/Controller/BoundaryCondition.h"
#include "FEM/Controller/Assembling.h"
#include "Optimize/Solver/OC.h"
#include "Optimize/Solver/MMA.h"
#include "Optimize/Solver/CONLIN.h"
#include "Optimize/Filter/SensitivityFilter.h"
#include "Optimize/Filter/HeavisideFilter.h"
#include "Optimize/Filter/DensityFilter.h"
#include "PrePost/Mesher/SquareMesh.h"

using namespace PANSFEM2;

namespace {

enum { EQ_PLANESTRAIN = 0, EQ_SOLID = 1, EQ_HEAT = 2 };
enum { FILTER_DENSITY = 0, FILTER_HEAVISIDE = 1 };
enum { OPT_OC = 0, OPT_MMA = 1, OPT_CONLIN = 2 };

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// silence the reference's std::cout chatter (OC.h:101 prints lambda, CG.h:147 prints iterations)
struct Quiet {
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};

// eq codes of include/pansfem2_b200.h (PF2_EQ_CODE): phys | shape << 8 | quad << 16 | quad2 << 24, 0 = the physics' default
enum { PHYS_PLANESTRAIN = 0, PHYS_SOLID = 1, PHYS_HEAT = 2, PHYS_PLANESTRESS = 3, PHYS_PLANESTRAIN_SRI = 4, PHYS_MASS = 5, PHYS_PLANESTRAIN_BBAR = 6, PHYS_MASS2 = 7, PHYS_PLANESTRAIN_WT = 8 };
enum { SHAPE_T3 = 1, SHAPE_T6, SHAPE_Q4, SHAPE_Q8, SHAPE_TET4, SHAPE_HEX8, SHAPE_HEX20 };
enum { QUAD_G1TRI = 1, QUAD_G3TRI, QUAD_G1SQ, QUAD_G4SQ, QUAD_G9SQ, QUAD_G1TET, QUAD_G8CUBE, QUAD_G27CUBE };
struct Sel { int phys, shape, quad, quad2; };
Sel decode(int eq) {
    Sel s = { eq & 0xff, (eq >> 8) & 0xff, (eq >> 16) & 0xff, (eq >> 24) & 0xff };
    bool solid = s.phys == PHYS_SOLID;
    if (!s.shape) s.shape = solid ? SHAPE_HEX8 : SHAPE_Q4;
    bool tri = s.shape == SHAPE_T3 || s.shape == SHAPE_T6;
    if (!s.quad) s.quad = tri ? QUAD_G1TRI : (s.shape == SHAPE_TET4 ? QUAD_G1TET : (solid ? QUAD_G8CUBE : QUAD_G4SQ));
    if ((s.phys == PHYS_PLANESTRAIN_SRI || s.phys == PHYS_PLANESTRAIN_BBAR) && !s.quad2) s.quad2 = tri ? QUAD_G1TRI : QUAD_G1SQ;
    return s;
}
int ndof_of(int eq) { int phys = eq & 0xff; return phys == PHYS_SOLID ? 3 : ((phys == PHYS_HEAT || phys == PHYS_MASS) ? 1 : 2); }

typedef std::vector<std::vector<std::pair<int, int> > > N2E;

// 2-D selections: the element routine <SF, IC> of the chosen physics, SRI with <SF, ICV, ICD> (PlaneStrain.h:63)
template<template<class>class SF, template<class>class IC, template<class>class ICV>
void em2d(const Sel& s, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& element, std::vector<Vector<double> >& x, double E, double V, double t) {
    switch (s.phys) {
        case PHYS_PLANESTRAIN: PlaneStrainStiffness<double, SF, IC>(Ke, n2e, element, { 0, 1 }, x, E, V, t); break;
        case PHYS_PLANESTRESS: PlaneStressStiffness<double, SF, IC>(Ke, n2e, element, { 0, 1 }, x, E, V, t); break;
        case PHYS_PLANESTRAIN_SRI: PlaneStrainStiffnessSRI<double, SF, ICV, IC>(Ke, n2e, element, { 0, 1 }, x, E, V, t); break;
        case PHYS_MASS: ReactionDiffusionConsistentMass<double, SF, IC>(Ke, n2e, element, { 0 }, x); Ke *= E*t; break;    // ReactionDiffusion.h:21 (E = t = 1)
        case PHYS_PLANESTRAIN_BBAR: PlaneStrainStiffnessBbar<double, SF, ICV, IC>(Ke, n2e, element, { 0, 1 }, x, E, V, t); break;   // PlaneStrain.h:129
        case PHYS_PLANESTRAIN_WT: PlaneStrainStiffnessWilsonTaylor<double, SF, IC>(Ke, n2e, element, { 0, 1 }, x, E, V, t); break;          // PlaneStrain.h:189
        case PHYS_MASS2: PlaneStrainMass<double, SF, IC>(Ke, n2e, element, { 0, 1 }, x, E, t); break;                                // PlaneStrain.h:386 (E = rho)
        default: HeatTransfer<double, SF, IC>(Ke, n2e, element, { 0 }, x, E, t); break;
    }
}
template<template<class>class SF>
void em_tri(const Sel& s, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, double E, double V, double t) {
    if (s.quad == QUAD_G1TRI) { if (s.quad2 == QUAD_G3TRI) em2d<SF, Gauss1Triangle, Gauss3Triangle>(s, Ke, n2e, el, x, E, V, t); else em2d<SF, Gauss1Triangle, Gauss1Triangle>(s, Ke, n2e, el, x, E, V, t); }
    else { if (s.quad2 == QUAD_G3TRI) em2d<SF, Gauss3Triangle, Gauss3Triangle>(s, Ke, n2e, el, x, E, V, t); else em2d<SF, Gauss3Triangle, Gauss1Triangle>(s, Ke, n2e, el, x, E, V, t); }
}
template<template<class>class SF, template<class>class IC>
void em_sq2(const Sel& s, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, double E, double V, double t) {
    if (s.quad2 == QUAD_G4SQ) em2d<SF, IC, Gauss4Square>(s, Ke, n2e, el, x, E, V, t);
    else if (s.quad2 == QUAD_G9SQ) em2d<SF, IC, Gauss9Square>(s, Ke, n2e, el, x, E, V, t);
    else em2d<SF, IC, Gauss1Square>(s, Ke, n2e, el, x, E, V, t);
}
template<template<class>class SF>
void em_sq(const Sel& s, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, double E, double V, double t) {
    if (s.quad == QUAD_G1SQ) em_sq2<SF, Gauss1Square>(s, Ke, n2e, el, x, E, V, t);
    else if (s.quad == QUAD_G9SQ) em_sq2<SF, Gauss9Square>(s, Ke, n2e, el, x, E, V, t);
    else em_sq2<SF, Gauss4Square>(s, Ke, n2e, el, x, E, V, t);
}
template<template<class>class SF>
void em_cube(const Sel& s, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, double E, double V) {
    if (s.quad == QUAD_G27CUBE) SolidLinearIsotropicElastic<double, SF, Gauss27Cubic>(Ke, n2e, el, { 0, 1, 2 }, x, E, V);
    else SolidLinearIsotropicElastic<double, SF, Gauss8Cubic>(Ke, n2e, el, { 0, 1, 2 }, x, E, V);
}

// One element matrix through the reference's own template selection.
void element_matrix(int eq, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& element, std::vector<Vector<double> >& x, double E, double V, double t) {
    Sel s = decode(eq);
    switch (s.shape) {
        case SHAPE_T3: em_tri<ShapeFunction3Triangle>(s, Ke, n2e, element, x, E, V, t); break;
        case SHAPE_T6: em_tri<ShapeFunction6Triangle>(s, Ke, n2e, element, x, E, V, t); break;
        case SHAPE_Q4: em_sq<ShapeFunction4Square>(s, Ke, n2e, element, x, E, V, t); break;
        case SHAPE_Q8: em_sq<ShapeFunction8Square>(s, Ke, n2e, element, x, E, V, t); break;
        case SHAPE_TET4: SolidLinearIsotropicElastic<double, ShapeFunction4Tetrahedron, Gauss1Tetrahedron>(Ke, n2e, element, { 0, 1, 2 }, x, E, V); break;
        case SHAPE_HEX8: em_cube<ShapeFunction8Cubic>(s, Ke, n2e, element, x, E, V); break;
        default: em_cube<ShapeFunction20Cubic>(s, Ke, n2e, element, x, E, V); break;
    }
}

std::vector<Vector<double> > make_nodes(int dim, int nnode, const double* coords) {
    std::vector<Vector<double> > x(nnode);
    for (int i = 0; i < nnode; i++) {
        std::vector<double> c(coords + (size_t)dim * i, coords + (size_t)dim * (i + 1));
        x[i] = Vector<double>(c);
    }
    return x;
}

std::vector<std::vector<int> > make_elements(int npe, int nelem, const int* conn) {
    std::vector<std::vector<int> > e(nelem);
    for (int i = 0; i < nelem; i++) e[i] = std::vector<int>(conn + (size_t)npe * i, conn + (size_t)npe * (i + 1));
    return e;
}

typedef std::vector<std::pair<std::pair<int, int>, double> > BCList;
BCList make_bc(int n, const int* node, const int* dof, const double* val) {
    BCList l(n);
    for (int i = 0; i < n; i++) l[i] = { { node[i], dof[i] }, val[i] };
    return l;
}

struct RefSystem {
    CSR<double>* K = nullptr;
    std::vector<double> F;
    std::vector<std::vector<int> > nodetoglobal;
    ~RefSystem() { delete K; }
};

struct RefFilter {
    int kind;
    int n;
    HeavisideFilter<double>* h = nullptr;
    DensityFilter<double>* d = nullptr;
    ~RefFilter() { delete h; delete d; }
    std::vector<double> apply(double beta, const std::vector<double>& s) {
        if (kind == FILTER_HEAVISIDE) { h->UpdateBeta(beta); return h->GetFilteredVariables(s); }
        return d->GetFilteredVariables(s);
    }
    std::vector<double> sens(double beta, const std::vector<double>& s, const std::vector<double>& dfdrho) {
        if (kind == FILTER_HEAVISIDE) { h->UpdateBeta(beta); return h->GetFilteredSensitivitis(s, dfdrho); }
        return d->GetFilteredSensitivitis(s, dfdrho);
    }
};

RefFilter* make_filter(int kind, int n, const long long* rowptr, const int* nbr, const double* w) {
    std::vector<std::vector<int> > neighbors(n);
    std::vector<std::vector<double> > ww(n);
    for (int i = 0; i < n; i++) {
        neighbors[i].assign(nbr + rowptr[i], nbr + rowptr[i + 1]);
        ww[i].assign(w + rowptr[i], w + rowptr[i + 1]);
    }
    RefFilter* f = new RefFilter();
    f->kind = kind; f->n = n;
    if (kind == FILTER_HEAVISIDE) f->h = new HeavisideFilter<double>(n, neighbors, ww);
    else f->d = new DensityFilter<double>(n, neighbors, ww);
    return f;
}

}  // namespace

extern "C" {

int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

// ---- element routines: PlaneStrain.h:21-58 / Solid.h:21-64 / HeatTransfer.h:20-43 ----
// xe: npe*dim coordinates of the element's nodes; Ke_out: (npe*ndof)^2 row-major.
int ref_element_matrix(int eq, int dim, int npe, const double* xe, double E, double V, double t, double* Ke_out) {
    std::vector<Vector<double> > x = make_nodes(dim, npe, xe);
    std::vector<int> element(npe);
    std::iota(element.begin(), element.end(), 0);
    Matrix<double> Ke;
    std::vector<std::vector<std::pair<int, int> > > n2e;
    element_matrix(eq, Ke, n2e, element, x, E, V, t);
    int m = npe * ndof_of(eq);
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Ke_out[i * m + j] = Ke(i, j);
    return m;
}

// ---- SetDirichlet (BoundaryCondition.h:20) + Renumbering (Assembling.h:175) + element loop with
//      Assembling (Assembling.h:47-66) + nodal loads (Assembling.h:152) + CSR(LILCSR&) (CSR.h:93-105) ----
// Emod: per-element modulus (already SIMP-interpolated by the caller); times[3] = {element, assembling, tocsr}.
void* ref_assemble(int eq, int dim, int nnode, const double* coords, int npe, int nelem, const int* conn,
                   int nfixed, const int* fnode, const int* fdof, const double* fval,
                   int nload, const int* lnode, const int* ldof, const double* lval,
                   const double* Emod, double V, double t, double* times) {
    int ndof = ndof_of(eq);
    std::vector<Vector<double> > x = make_nodes(dim, nnode, coords);
    std::vector<std::vector<int> > elements = make_elements(npe, nelem, conn);
    BCList ufixed = make_bc(nfixed, fnode, fdof, fval), qfixed = make_bc(nload, lnode, ldof, lval);

    RefSystem* sys = new RefSystem();
    std::vector<Vector<double> > u(nnode, Vector<double>(ndof));
    sys->nodetoglobal = std::vector<std::vector<int> >(nnode, std::vector<int>(ndof, 0));
    SetDirichlet(u, sys->nodetoglobal, ufixed);
    int KDEGREE = Renumbering(sys->nodetoglobal);
    LILCSR<double> K(KDEGREE, KDEGREE);
    sys->F = std::vector<double>(KDEGREE, 0.0);
    double te = 0, ta = 0;
    for (int i = 0; i < nelem; i++) {
        std::vector<std::vector<std::pair<int, int> > > n2e;
        Matrix<double> Ke;
        double t0 = now();
        element_matrix(eq, Ke, n2e, elements[i], x, Emod[i], V, t);
        double t1 = now();
        Assembling(K, sys->F, u, Ke, sys->nodetoglobal, n2e, elements[i]);
        double t2 = now();
        te += t1 - t0; ta += t2 - t1;
    }
    Assembling(sys->F, qfixed, sys->nodetoglobal);
    double t3 = now();
    sys->K = new CSR<double>(K);
    double t4 = now();
    if (times) { times[0] = te; times[1] = ta; times[2] = t4 - t3; }
    return sys;
}

void* ref_system_from_csr(int n, const int* indptr, const int* indices, const double* data) {
    RefSystem* sys = new RefSystem();
    sys->K = new CSR<double>(n, n);
    sys->K->indptr.assign(indptr, indptr + n + 1);
    sys->K->indices.assign(indices, indices + indptr[n]);
    sys->K->data.assign(data, data + indptr[n]);
    sys->F.assign(n, 0.0);
    return sys;
}
void ref_system_free(void* h) { delete (RefSystem*)h; }
int ref_system_rows(void* h) { return ((RefSystem*)h)->K->ROWS; }
long long ref_system_nnz(void* h) { return (long long)((RefSystem*)h)->K->data.size(); }
void ref_system_get(void* h, int* indptr, int* indices, double* data, double* F) {
    RefSystem* s = (RefSystem*)h;
    if (indptr) std::copy(s->K->indptr.begin(), s->K->indptr.end(), indptr);
    if (indices) std::copy(s->K->indices.begin(), s->K->indices.end(), indices);
    if (data) std::copy(s->K->data.begin(), s->K->data.end(), data);
    if (F) std::copy(s->F.begin(), s->F.end(), F);
}
void ref_system_nodetoglobal(void* h, int* out) {
    RefSystem* s = (RefSystem*)h;
    size_t k = 0;
    for (auto& n : s->nodetoglobal) for (int d : n) out[k++] = d;
}

// ---- CSR<T>::operator* (CSR.h:109-122) ----
void ref_spmv(void* h, const double* x, double* y, int repeat, double* seconds) {
    RefSystem* s = (RefSystem*)h;
    std::vector<double> xv(x, x + s->K->COLS), yv;
    double t0 = now();
    for (int r = 0; r < (repeat > 0 ? repeat : 1); r++) yv = (*s->K) * xv;
    double t1 = now();
    if (seconds) *seconds = (t1 - t0) / (repeat > 0 ? repeat : 1);
    std::copy(yv.begin(), yv.end(), y);
}

// ---- CG (CG.h:124-154), ScalingCG (CG.h:420-453), ILU0 (CG.h:258-284), PreILU0 (:289-315), ILU0CG (:320-352) ----
// kind: 0 = CG, 1 = ScalingCG, 2 = ILU0CG (factor computed here, timed separately in seconds[1]).
void ref_solve(void* h, int kind, const double* b, int itrmax, double eps, double* x, double* seconds) {
    RefSystem* s = (RefSystem*)h;
    Quiet q;
    std::vector<double> bv(b, b + s->K->ROWS), xv;
    double t0 = now(), tf = 0;
    if (kind == 0) xv = CG(*s->K, bv, itrmax, eps);
    else if (kind == 1) xv = ScalingCG(*s->K, bv, itrmax, eps);
    else if (kind == 3) xv = BiCGSTAB(*s->K, bv, itrmax, eps);               // CG.h:159
    else if (kind == 4) xv = BiCGSTAB2(*s->K, bv, itrmax, eps);              // CG.h:199
    else if (kind == 5) xv = ScalingBiCGSTAB(*s->K, bv, itrmax, eps);        // CG.h:458
    else if (kind == 6) {                                                    // CG.h:357
        CSR<double> M = ILU0(*s->K);
        tf = now() - t0;
        t0 = now();
        xv = ILU0BiCGSTAB(*s->K, M, bv, itrmax, eps);
    } else {
        CSR<double> M = ILU0(*s->K);
        tf = now() - t0;
        t0 = now();
        xv = ILU0CG(*s->K, M, bv, itrmax, eps);
    }
    double t1 = now();
    if (seconds) { seconds[0] = t1 - t0; seconds[1] = tf; }
    std::copy(xv.begin(), xv.end(), x);
}
void* ref_ilu0(void* h) {
    RefSystem* s = (RefSystem*)h;
    RefSystem* m = new RefSystem();
    m->K = new CSR<double>(ILU0(*s->K));
    m->F.assign(s->K->ROWS, 0.0);
    return m;
}
void ref_preilu0(void* hM, const double* b, double* x) {
    RefSystem* m = (RefSystem*)hM;
    std::vector<double> bv(b, b + m->K->ROWS);
    std::vector<double> xv = PreILU0(*m->K, bv);
    std::copy(xv.begin(), xv.end(), x);
}

// ---- filters: HeavisideFilter.h:61-99, DensityFilter.h:45-71 ----
void* ref_filter_create(int kind, int n, const long long* rowptr, const int* nbr, const double* w) {
    return make_filter(kind, n, rowptr, nbr, w);
}
void ref_filter_free(void* h) { delete (RefFilter*)h; }
void ref_filter_apply(void* h, double beta, const double* s, double* rho) {
    RefFilter* f = (RefFilter*)h;
    std::vector<double> r = f->apply(beta, std::vector<double>(s, s + f->n));
    std::copy(r.begin(), r.end(), rho);
}
void ref_filter_sens(void* h, double beta, const double* s, const double* dfdrho, double* dfds) {
    RefFilter* f = (RefFilter*)h;
    std::vector<double> r = f->sens(beta, std::vector<double>(s, s + f->n), std::vector<double>(dfdrho, dfdrho + f->n));
    std::copy(r.begin(), r.end(), dfds);
}

// ---- OC (OC.h:46-107) with the sample's constraint functor (sample_optimize_density_oc.cpp:198-207) ----
void* ref_oc_create(int n, double iota, double lmin, double lmax, double leps, double move) {
    return new OC<double>(n, iota, lmin, lmax, leps, move, std::vector<double>(n, 0.01), std::vector<double>(n, 1.0));
}
void ref_oc_free(void* h) { delete (OC<double>*)h; }
int ref_oc_isconvergence(void* h, double f) { return ((OC<double>*)h)->IsConvergence(f) ? 1 : 0; }
void ref_oc_update(void* h, void* hfilter, double beta, double weightlimit, double scale1, int n, double* s,
                   double f, const double* dfds, double g, const double* dgds) {
    Quiet q;
    OC<double>* oc = (OC<double>*)h;
    RefFilter* filter = (RefFilter*)hfilter;
    std::vector<double> sv(s, s + n);
    oc->UpdateVariables(sv, f, std::vector<double>(dfds, dfds + n), g, std::vector<double>(dgds, dgds + n),
        [&](std::vector<double> _xkp1) {
            double gg = 0.0;
            std::vector<double> rho = filter->apply(beta, _xkp1);
            for (int i = 0; i < n; i++) gg += scale1 * rho[i] / (weightlimit * n);
            return gg - 1.0 * scale1;
        });
    std::copy(sv.begin(), sv.end(), s);
}

// ---- MMA (MMA.h:64-419) ----
void* ref_mma_create(int n, int m, double a0, const double* a, const double* c, const double* d,
                     const double* xmin, const double* xmax) {
    return new MMA<double>(n, m, a0, std::vector<double>(a, a + m), std::vector<double>(c, c + m),
                           std::vector<double>(d, d + m), std::vector<double>(xmin, xmin + n),
                           std::vector<double>(xmax, xmax + n));
}
void ref_mma_free(void* h) { delete (MMA<double>*)h; }
void ref_mma_setparameters(void* h, double raa0, double albefa, double move, double asyinit, double asydecr,
                           double asyincr, double epsvalue) {
    ((MMA<double>*)h)->SetParameters(raa0, albefa, move, asyinit, asydecr, asyincr, epsvalue);
}
int ref_mma_isconvergence(void* h, double f) { return ((MMA<double>*)h)->IsConvergence(f) ? 1 : 0; }
void ref_mma_update(void* h, int n, int m, double* x, double f, const double* dfdx, const double* g, const double* dgdx) {
    std::vector<double> xv(x, x + n);
    std::vector<std::vector<double> > dg(m);
    for (int i = 0; i < m; i++) dg[i].assign(dgdx + (size_t)i * n, dgdx + (size_t)(i + 1) * n);
    ((MMA<double>*)h)->UpdateVariables(xv, f, std::vector<double>(dfdx, dfdx + n), std::vector<double>(g, g + m), dg);
    std::copy(xv.begin(), xv.end(), x);
}

// ---- CONLIN (CONLIN.h:59-373) ----
void* ref_conlin_create(int n, int m, double a0, const double* a, const double* c, const double* d, const double* xmin, const double* xmax) {
    return new CONLIN<double>(n, m, a0, std::vector<double>(a, a + m), std::vector<double>(c, c + m), std::vector<double>(d, d + m),
                              std::vector<double>(xmin, xmin + n), std::vector<double>(xmax, xmax + n));
}
void ref_conlin_free(void* h) { delete (CONLIN<double>*)h; }
void ref_conlin_setparameters(void* h, double move, double epsvalue) { ((CONLIN<double>*)h)->SetParameters(move, epsvalue); }
int ref_conlin_isconvergence(void* h, double f) { return ((CONLIN<double>*)h)->IsConvergence(f) ? 1 : 0; }
void ref_conlin_update(void* h, int n, int m, double* x, double f, const double* dfdx, const double* g, const double* dgdx) {
    std::vector<double> xv(x, x + n);
    std::vector<std::vector<double> > dg(m);
    for (int i = 0; i < m; i++) dg[i].assign(dgdx + (size_t)i * n, dgdx + (size_t)(i + 1) * n);
    ((CONLIN<double>*)h)->UpdateVariables(xv, f, std::vector<double>(dfdx, dfdx + n), std::vector<double>(g, g + m), dg);
    std::copy(xv.begin(), xv.end(), x);
}

// ---- SensitivityFilter / SensitivityFilter2 (SensitivityFilter.h:44-55, 88-99); kind 2 = Sigmund, 3 = Borrvall ----
void ref_sensitivity_filter(int kind, int n, const long long* rowptr, const int* nbr, const double* w, const double* s, const double* dfds, double* out) {
    std::vector<std::vector<int> > neighbors(n);
    std::vector<std::vector<double> > ww(n);
    for (int i = 0; i < n; i++) { neighbors[i].assign(nbr + rowptr[i], nbr + rowptr[i + 1]); ww[i].assign(w + rowptr[i], w + rowptr[i + 1]); }
    std::vector<double> r;
    if (kind == 2) r = SensitivityFilter<double>(n, neighbors, ww).GetFilteredSensitivitis(std::vector<double>(s, s + n), std::vector<double>(dfds, dfds + n));
    else r = SensitivityFilter2<double>(n, neighbors, ww).GetFilteredSensitivitis(std::vector<double>(s, s + n), std::vector<double>(dfds, dfds + n));
    std::copy(r.begin(), r.end(), out);
}

// ---- the non-symmetric system of sample/advection/sample_advectiondiffusion_static.cpp:37-58: T3 elements,
//      Ke = Advection + Diffusion + AdvectionSUPG (Advection.h), assembled with the sample's calls.  Used as a fixture for
//      the BiCGSTAB family (the sample solves it with BiCGSTAB and its output is the committed AdvectionSUPG.vtk). ----
void* ref_advection_system(int nnode, const double* coords, int nelem, const int* conn, int nfixed, const int* fnode, const double* fval,
                           double a, double theta_deg, double k) {
    std::vector<Vector<double> > x = make_nodes(2, nnode, coords);
    std::vector<std::vector<int> > elements = make_elements(3, nelem, conn);
    BCList ufixed(nfixed);
    for (int i = 0; i < nfixed; i++) ufixed[i] = { { fnode[i], 0 }, fval[i] };
    RefSystem* sys = new RefSystem();
    std::vector<Vector<double> > T(x.size(), Vector<double>(1));
    sys->nodetoglobal = std::vector<std::vector<int> >(x.size(), std::vector<int>(1, 0));
    SetDirichlet(T, sys->nodetoglobal, ufixed);
    int KDEGREE = Renumbering(sys->nodetoglobal);
    LILCSR<double> K(KDEGREE, KDEGREE);
    sys->F.assign(KDEGREE, 0.0);
    const double ax = a*cos(theta_deg*M_PI/180.0), ay = a*sin(theta_deg*M_PI/180.0);
    for (auto element : elements) {
        N2E nodetoelement;
        Matrix<double> A, B, C;
        Advection<double, ShapeFunction3Triangle, Gauss1Triangle>(A, nodetoelement, element, { 0 }, x, ax, ay);
        Diffusion<double, ShapeFunction3Triangle, Gauss1Triangle>(B, nodetoelement, element, { 0 }, x, k);
        AdvectionSUPG<double, ShapeFunction3Triangle, Gauss1Triangle>(C, nodetoelement, element, { 0 }, x, ax, ay, k);
        Matrix<double> Ke = A + B + C;
        Assembling(K, sys->F, T, Ke, sys->nodetoglobal, nodetoelement, element);
    }
    sys->K = new CSR<double>(K);
    return sys;
}

// ---- PlaneStiffness / PlaneStiffnessBbar / PlaneStiffnessWilsonTaylor (Homogenization.h:141-280) for any 2-D <SF, IC[, ICV]>;
//      mode 0 / 1 / 2; quad2 = ICV of the B-bar variant ----
}   // extern "C"
namespace {
template<template<class>class SF, template<class>class IC, template<class>class ICV>
void plane_d(int mode, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, Matrix<double> D, double t) {
    if (mode == 1) PlaneStiffnessBbar<double, SF, ICV, IC>(Ke, n2e, el, { 0, 1 }, x, D, t);
    else if (mode == 2) PlaneStiffnessWilsonTaylor<double, SF, IC>(Ke, n2e, el, { 0, 1 }, x, D, t);
    else PlaneStiffness<double, SF, IC>(Ke, n2e, el, { 0, 1 }, x, D, t);
}
template<template<class>class SF>
void plane_d_tri(int mode, int quad, int quad2, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, Matrix<double> D, double t) {
    if (quad == QUAD_G3TRI) { if (quad2 == QUAD_G3TRI) plane_d<SF, Gauss3Triangle, Gauss3Triangle>(mode, Ke, n2e, el, x, D, t); else plane_d<SF, Gauss3Triangle, Gauss1Triangle>(mode, Ke, n2e, el, x, D, t); }
    else { if (quad2 == QUAD_G3TRI) plane_d<SF, Gauss1Triangle, Gauss3Triangle>(mode, Ke, n2e, el, x, D, t); else plane_d<SF, Gauss1Triangle, Gauss1Triangle>(mode, Ke, n2e, el, x, D, t); }
}
template<template<class>class SF, template<class>class IC>
void plane_d_sq2(int mode, int quad2, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, Matrix<double> D, double t) {
    if (quad2 == QUAD_G4SQ) plane_d<SF, IC, Gauss4Square>(mode, Ke, n2e, el, x, D, t);
    else if (quad2 == QUAD_G9SQ) plane_d<SF, IC, Gauss9Square>(mode, Ke, n2e, el, x, D, t);
    else plane_d<SF, IC, Gauss1Square>(mode, Ke, n2e, el, x, D, t);
}
template<template<class>class SF>
void plane_d_sq(int mode, int quad, int quad2, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, Matrix<double> D, double t) {
    if (quad == QUAD_G1SQ) plane_d_sq2<SF, Gauss1Square>(mode, quad2, Ke, n2e, el, x, D, t);
    else if (quad == QUAD_G9SQ) plane_d_sq2<SF, Gauss9Square>(mode, quad2, Ke, n2e, el, x, D, t);
    else plane_d_sq2<SF, Gauss4Square>(mode, quad2, Ke, n2e, el, x, D, t);
}
}   // namespace
extern "C" {
int ref_plane_d_element(int shape, int quad, int quad2, int mode, int npe, const double* xe, const double* D9, double t, double* Ke_out) {
    std::vector<Vector<double> > x = make_nodes(2, npe, xe);
    std::vector<int> element(npe);
    std::iota(element.begin(), element.end(), 0);
    Matrix<double> D(3, 3), Ke;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) D(i, j) = D9[3 * i + j];
    N2E n2e;
    switch (shape) {
        case SHAPE_T3: plane_d_tri<ShapeFunction3Triangle>(mode, quad, quad2, Ke, n2e, element, x, D, t); break;
        case SHAPE_T6: plane_d_tri<ShapeFunction6Triangle>(mode, quad, quad2, Ke, n2e, element, x, D, t); break;
        case SHAPE_Q8: plane_d_sq<ShapeFunction8Square>(mode, quad, quad2, Ke, n2e, element, x, D, t); break;
        default: plane_d_sq<ShapeFunction4Square>(mode, quad, quad2, Ke, n2e, element, x, D, t); break;
    }
    const int m = 2 * npe;
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Ke_out[i * m + j] = Ke(i, j);
    return m;
}

// ---- the advection-diffusion element family (Advection.h:19-229) for any 2-D <SF, IC>, and the systems the two advection samples
//      build from it.  terms: 1 Advection, 2 Diffusion, 4 AdvectionSUPG, 8 AdvectionShockCapturing, 16 Mass, 32 MassSUPG. ----
}   // extern "C"
namespace {
enum { ADV_A = 1, ADV_D = 2, ADV_S = 4, ADV_SC = 8, ADV_M = 16, ADV_MS = 32 };
template<template<class>class SF, template<class>class IC>
void adv_term(int term, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, double ax, double ay, double k) {
    switch (term) {
        case ADV_A: Advection<double, SF, IC>(Ke, n2e, el, { 0 }, x, ax, ay); break;
        case ADV_D: Diffusion<double, SF, IC>(Ke, n2e, el, { 0 }, x, k); break;
        case ADV_S: AdvectionSUPG<double, SF, IC>(Ke, n2e, el, { 0 }, x, ax, ay, k); break;
        case ADV_SC: AdvectionShockCapturing<double, SF, IC>(Ke, n2e, el, { 0 }, x, ax, ay, k); break;
        case ADV_M: Mass<double, SF, IC>(Ke, n2e, el, { 0 }, x); break;
        default: MassSUPG<double, SF, IC>(Ke, n2e, el, { 0 }, x, ax, ay, k); break;
    }
}
void adv_term_sel(int shape, int quad, int term, Matrix<double>& Ke, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, double ax, double ay, double k) {
#define ADV_TRI(SF) do { if (quad == QUAD_G3TRI) adv_term<SF, Gauss3Triangle>(term, Ke, n2e, el, x, ax, ay, k); else adv_term<SF, Gauss1Triangle>(term, Ke, n2e, el, x, ax, ay, k); } while (0)
#define ADV_SQ(SF) do { if (quad == QUAD_G1SQ) adv_term<SF, Gauss1Square>(term, Ke, n2e, el, x, ax, ay, k); else if (quad == QUAD_G9SQ) adv_term<SF, Gauss9Square>(term, Ke, n2e, el, x, ax, ay, k); \
                        else adv_term<SF, Gauss4Square>(term, Ke, n2e, el, x, ax, ay, k); } while (0)
    switch (shape) {
        case SHAPE_T3: ADV_TRI(ShapeFunction3Triangle); break;
        case SHAPE_T6: ADV_TRI(ShapeFunction6Triangle); break;
        case SHAPE_Q8: ADV_SQ(ShapeFunction8Square); break;
        default: ADV_SQ(ShapeFunction4Square); break;
    }
#undef ADV_TRI
#undef ADV_SQ
}
// sum of the selected terms among `group`, in the samples' order (A + D + AS [+ SC]; M + MS)
bool adv_sum(int shape, int quad, int terms, int group, Matrix<double>& S, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, double ax, double ay, double k) {
    bool any = false;
    static const int order[6] = { ADV_A, ADV_D, ADV_S, ADV_SC, ADV_M, ADV_MS };
    for (int t : order) {
        if (!(terms & group & t)) continue;
        Matrix<double> P;
        adv_term_sel(shape, quad, t, P, n2e, el, x, ax, ay, k);
        if (!any) S = P; else S = S + P;
        any = true;
    }
    if (!any) {
        S = Matrix<double>(el.size(), el.size());
        n2e = N2E(el.size(), std::vector<std::pair<int, int> >(1));
        for (size_t i = 0; i < el.size(); i++) n2e[i][0] = std::make_pair(0, (int)i);
    }
    return any;
}
}   // namespace
extern "C" {

// one term (or the sum of several) on one element
int ref_advdiff_element(int shape, int quad, int terms, int npe, const double* xe, double ax, double ay, double k, double* Ke_out) {
    std::vector<Vector<double> > x = make_nodes(2, npe, xe);
    std::vector<int> element(npe);
    std::iota(element.begin(), element.end(), 0);
    Matrix<double> Ke;
    N2E n2e;
    adv_sum(shape, quad, terms, 63, Ke, n2e, element, x, ax, ay, k);
    for (int i = 0; i < npe; i++) for (int j = 0; j < npe; j++) Ke_out[i * npe + j] = Ke(i, j);
    return npe;
}

// dt == 0: the static sample's system (sample_advectiondiffusion_static.cpp:42-55), Ke = A + B + C [+ D];
// dt > 0: one step of the dynamic sample (sample_advectiondiffusion_dynamic.cpp:51-70), Ke = (M + MS)/dt + theta*(A + D + AS),
// Fe = ((M + MS)/dt - (1 - theta)*(A + D + AS))*Te.  vel: per-element velocity (nelem*2); Tn: nodal field (fixed nodes are overwritten
// with their Dirichlet values, as SetDirichlet does).
void* ref_advdiff_system(int shape, int quad, int terms, int nnode, const double* coords, int npe, int nelem, const int* conn, int nfixed,
                         const int* fnode, const double* fval, const double* vel, double k, double dt, double theta, const double* Tn) {
    std::vector<Vector<double> > x = make_nodes(2, nnode, coords);
    std::vector<std::vector<int> > elements = make_elements(npe, nelem, conn);
    BCList ufixed(nfixed);
    for (int i = 0; i < nfixed; i++) ufixed[i] = { { fnode[i], 0 }, fval[i] };
    RefSystem* sys = new RefSystem();
    std::vector<Vector<double> > T(x.size(), Vector<double>(1));
    if (Tn) for (int i = 0; i < nnode; i++) T[i](0) = Tn[i];
    sys->nodetoglobal = std::vector<std::vector<int> >(x.size(), std::vector<int>(1, 0));
    SetDirichlet(T, sys->nodetoglobal, ufixed);
    int KDEGREE = Renumbering(sys->nodetoglobal);
    LILCSR<double> K(KDEGREE, KDEGREE);
    sys->F.assign(KDEGREE, 0.0);
    for (int e = 0; e < nelem; e++) {
        std::vector<int> element = elements[e];
        const double ax = vel[2 * e], ay = vel[2 * e + 1];
        N2E nodetoelement;
        Matrix<double> KK;     // (a default-constructed Matrix must be assigned before it dies: Matrix.h:91-99 leaves `values` dangling)
        adv_sum(shape, quad, terms, ADV_A | ADV_D | ADV_S | ADV_SC, KK, nodetoelement, element, x, ax, ay, k);
        if (dt == 0.0) {
            Assembling(K, sys->F, T, KK, sys->nodetoglobal, nodetoelement, element);
        } else {
            Matrix<double> MM;
            adv_sum(shape, quad, terms, ADV_M | ADV_MS, MM, nodetoelement, element, x, ax, ay, k);
            Vector<double> Te = ElementVector(T, nodetoelement, element);
            Matrix<double> Ke = MM/dt + theta*KK;
            Vector<double> Fe = (MM/dt - (1.0 - theta)*KK)*Te;
            Assembling(K, sys->F, T, Ke, Fe, sys->nodetoglobal, nodetoelement, element);
        }
    }
    sys->K = new CSR<double>(K);
    return sys;
}

// ---- the level-set design loop of sample/optimize/sample_optimize_levelset.cpp:75-192, element routines and helpers
//      called exactly as there (PlaneStressStiffness, ReactionDiffusion{ConsistentMass,Stiffness,Reaction},
//      InterpolateNodalFromElemental / InterpolateElementalFromNodal, ScalingCG) ----
// prm = { Vmax, tau, E0, Emin, nu, nvol, dt, d, p };  phi_io (nnode), str_io (nelem) in/out;  u_out (nnode*2);
// hist[3*t] = { objective[t] (the sample prints objective/nelem), vol, lambda };  *converged = 1 when the sample's test fired.
// Returns the number of iterations whose history was recorded.
int ref_levelset_run(int nnode, const double* coords, int nelem, const int* conn,
                     int nfixed, const int* fnode, const int* fdof, const double* fval,
                     int nload, const int* lnode, const int* ldof, const double* lval,
                     int nphi, const int* pnode, const double* prm, int tmax,
                     double* phi_io, double* str_io, double* u_out, double* hist, int* converged) {
    Quiet q;
    const double Vmax = prm[0], tau = prm[1], E0 = prm[2], Emin = prm[3], nu = prm[4], nvol = prm[5], dt = prm[6], d = prm[7], p = prm[8];
    std::vector<Vector<double> > x = make_nodes(2, nnode, coords);
    std::vector<std::vector<int> > elements = make_elements(4, nelem, conn);
    BCList ufixed = make_bc(nfixed, fnode, fdof, fval), qfixed = make_bc(nload, lnode, ldof, lval);
    BCList phifixed(nphi);
    for (int i = 0; i < nphi; i++) phifixed[i] = { { pnode[i], 0 }, 0.0 };

    double A1 = -1.5*(1.0 - nu)*(1.0 - 14.0*nu + 15.0*pow(nu, 2.0))*E0/((1.0 + nu)*(7.0 - 5.0*nu)*pow(1.0 - 2.0*nu, 2.0));
    double A2 = 7.5*(1.0 - nu)*E0/((1.0 + nu)*(7.0 - 5.0*nu));
    double c = A1/(A1 + 2.0*A2);

    std::vector<Vector<double> > phi(nnode, Vector<double>(1));
    for (int i = 0; i < nnode; i++) phi[i](0) = phi_io[i];
    std::vector<double> str(str_io, str_io + nelem);
    double volInit = std::accumulate(str.begin(), str.end(), 0.0)/(double)elements.size();
    std::vector<double> objective(tmax);
    *converged = 0;
    int t = 0;
    for (; t < tmax; t++) {
        std::vector<Vector<double> > u(x.size(), Vector<double>(2));
        std::vector<std::vector<int> > nodetoglobal(x.size(), std::vector<int>(2, 0));
        SetDirichlet(u, nodetoglobal, ufixed);
        int KDEGREE = Renumbering(nodetoglobal);
        LILCSR<double> K(KDEGREE, KDEGREE);
        std::vector<double> F(KDEGREE, 0.0);
        for (size_t i = 0; i < elements.size(); i++) {
            N2E nodetoelement;
            Matrix<double> Ke;
            PlaneStressStiffness<double, ShapeFunction4Square, Gauss4Square>(Ke, nodetoelement, elements[i], { 0, 1 }, x, Emin + str[i]*(E0 - Emin), nu, 1.0);
            Assembling(K, F, u, Ke, nodetoglobal, nodetoelement, elements[i]);
        }
        Assembling(F, qfixed, nodetoglobal);
        CSR<double> Kmod(K);
        std::vector<double> result = ScalingCG(Kmod, F, 100000, 1.0e-10);
        Disassembling(u, result, nodetoglobal);
        for (int i = 0; i < nnode; i++) { u_out[2*i] = u[i](0); u_out[2*i + 1] = u[i](1); }

        std::vector<Vector<double> > TD(elements.size(), Vector<double>(1));
        for (size_t i = 0; i < elements.size(); i++) {
            N2E nodetoelement;
            Matrix<double> Ke;
            PlaneStressStiffness<double, ShapeFunction4Square, Gauss4Square>(Ke, nodetoelement, elements[i], { 0, 1 }, x, Emin + str[i]*(E0 - Emin), nu, 1.0);
            Vector<double> ue = ElementVector(u, nodetoelement, elements[i]);
            objective[t] += ue*(Ke*ue);
            PlaneStressStiffness<double, ShapeFunction4Square, Gauss4Square>(Ke, nodetoelement, elements[i], { 0, 1 }, x, (A1 + 2.0*A2)*(1.0 - pow(c, 2.0)), c, 1.0);
            TD[i](0) = (1.0e-4 + str[i]*(1.0 - 1.0e-4))*ue*(Ke*ue);
        }
        std::vector<Vector<double> > TDN = InterpolateNodalFromElemental<double, Vector>(x.size(), Vector<double>(1), TD, elements);

        double vol = std::accumulate(str.begin(), str.end(), 0.0)/(double)elements.size();
        double ex = Vmax + (volInit - Vmax)*std::max(0.0, 1.0 - (t + 1)/(double)nvol);
        double lambda = std::accumulate(TDN.begin(), TDN.end(), Vector<double>(1))(0)/(double)x.size()*exp(p*((vol - ex)/ex + d));
        hist[3*t] = objective[t]; hist[3*t + 1] = vol; hist[3*t + 2] = lambda;

        if (t > nvol && fabs(vol - Vmax) < 0.005 &&
            fabs(objective[t] - objective[t - 5]) < 0.01*fabs(objective[t]) &&
            fabs(objective[t] - objective[t - 4]) < 0.01*fabs(objective[t]) &&
            fabs(objective[t] - objective[t - 3]) < 0.01*fabs(objective[t]) &&
            fabs(objective[t] - objective[t - 2]) < 0.01*fabs(objective[t]) &&
            fabs(objective[t] - objective[t - 1]) < 0.01*fabs(objective[t])) {
            *converged = 1;
            t++;
            break;
        }

        double C = 0.0;
        for (size_t i = 0; i < x.size(); i++) C += fabs(TDN[i](0));
        C = elements.size()/C;

        std::vector<std::vector<int> > nodetoglobal2(x.size(), std::vector<int>(1, 0));
        SetDirichlet(phi, nodetoglobal2, phifixed);
        int TDEGREE = Renumbering(nodetoglobal2);
        LILCSR<double> T(TDEGREE, TDEGREE);
        std::vector<double> Y(TDEGREE, 0.0);
        for (size_t i = 0; i < elements.size(); i++) {
            N2E nodetoelement;
            Matrix<double> Me, Ke;
            Vector<double> Fe;
            ReactionDiffusionConsistentMass<double, ShapeFunction4Square, Gauss4Square>(Me, nodetoelement, elements[i], { 0 }, x);
            ReactionDiffusionStiffness<double, ShapeFunction4Square, Gauss4Square>(Ke, nodetoelement, elements[i], { 0 }, x, tau*elements.size());
            ReactionDiffusionReaction<double, ShapeFunction4Square, Gauss4Square>(Fe, nodetoelement, elements[i], { 0 }, x, TDN, [&](double _u, Vector<double> _dudX) {
                return C*(_u - lambda);
            });
            Vector<double> phie = ElementVector(phi, nodetoelement, elements[i]);
            Matrix<double> Te = Me/dt + Ke;
            Vector<double> Ye = Me/dt*phie + Fe;
            Assembling(T, Y, phi, Te, nodetoglobal2, nodetoelement, elements[i]);
            Assembling(Y, Ye, nodetoglobal2, nodetoelement, elements[i]);
        }
        CSR<double> Tmod(T);
        std::vector<double> result2 = ScalingCG(Tmod, Y, 100000, 1.0e-10);
        Disassembling(phi, result2, nodetoglobal2);
        for (size_t i = 0; i < x.size(); i++) phi[i](0) = std::max(std::min(1.0, phi[i](0)), -1.0);
        std::vector<Vector<double> > phie = InterpolateElementalFromNodal<double, Vector>(Vector<double>(1), phi, elements);
        for (size_t i = 0; i < elements.size(); i++) str[i] = (phie[i](0) < 0.0) ? 0.0 : 1.0;
    }
    for (int i = 0; i < nnode; i++) phi_io[i] = phi[i](0);
    std::copy(str.begin(), str.end(), str_io);
    return t;
}

// ---- the SIMP design loop of sample/optimize/sample_optimize_density_{oc,mma}.cpp:83-208 on a caller-supplied
//      mesh (so the same driver serves plane strain Q4, heat Q4 and solid hex8).  The loop body follows the
//      sample line by line; only VTK output is dropped and timings added.
// params: [0]=E0 [1]=E1 [2]=V [3]=p [4]=weightlimit [5]=scale0 [6]=scale1 [7]=thickness [8]=beta0
//         [9]=beta_period (design iterations between beta doublings; <=0: never) [10]=cg itrmax [11]=cg eps
// oc: [iota,lmin,lmax,leps,move]; mma: [raa0,albefa,move,asyinit,asydecr,asyincr,epsvalue,a0,a,c,d,xmin,xmax]
// outputs: s (in/out, nelem), rho (nelem), u (nnode*ndof), r (nnode*ndof), hist[4*niter] = {f, g, seconds, converged}
//          phase[8] accumulates {filter, element+assembling, tocsr, solve, reaction, sensitivity, filter-sens, update}
// returns the number of design iterations performed (the converged iteration counts, as in the sample).
int ref_simp_run(int eq, int dim, int nnode, const double* coords, int npe, int nelem, const int* conn,
                 int nfixed, const int* fnode, const int* fdof, const double* fval,
                 int nload, const int* lnode, const int* ldof, const double* lval,
                 int filter_kind, const long long* rowptr, const int* nbr, const double* w,
                 int opt_kind, const double* optp, const double* params, int niter, int check_convergence,
                 double* s_io, double* rho_out, double* u_out, double* r_out, double* hist, double* phase) {
    Quiet q;
    int ndof = ndof_of(eq);
    std::vector<Vector<double> > x = make_nodes(dim, nnode, coords);
    std::vector<std::vector<int> > elements = make_elements(npe, nelem, conn);
    BCList ufixed = make_bc(nfixed, fnode, fdof, fval), qfixed = make_bc(nload, lnode, ldof, lval);
    RefFilter* filter = make_filter(filter_kind, nelem, rowptr, nbr, w);
    std::vector<double> s(s_io, s_io + nelem);

    double E0 = params[0], E1 = params[1], Poisson = params[2], p = params[3], weightlimit = params[4];
    double scale0 = params[5], scale1 = params[6], thick = params[7], beta = params[8];
    int beta_period = (int)params[9];
    int itrmax = (int)params[10];
    double cgeps = params[11];

    OC<double>* oc = nullptr;
    MMA<double>* mma = nullptr;
    CONLIN<double>* conlin = nullptr;
    if (opt_kind == OPT_CONLIN) {
        conlin = new CONLIN<double>(nelem, 1, optp[2], std::vector<double>(1, optp[3]), std::vector<double>(1, optp[4]),
                                    std::vector<double>(1, optp[5]), std::vector<double>(nelem, optp[6]), std::vector<double>(nelem, optp[7]));
        conlin->SetParameters(optp[0], optp[1]);
    } else if (opt_kind == OPT_OC) {
        oc = new OC<double>(nelem, optp[0], optp[1], optp[2], optp[3], optp[4], std::vector<double>(nelem, 0.01), std::vector<double>(nelem, 1.0));
    } else {
        mma = new MMA<double>(nelem, 1, optp[7], std::vector<double>(1, optp[8]), std::vector<double>(1, optp[9]),
                              std::vector<double>(1, optp[10]), std::vector<double>(nelem, optp[11]), std::vector<double>(nelem, optp[12]));
        mma->SetParameters(optp[0], optp[1], optp[2], optp[3], optp[4], optp[5], optp[6]);
    }
    if (phase) for (int i = 0; i < 8; i++) phase[i] = 0;

    int k = 0;
    std::vector<double> rho;
    std::vector<Vector<double> > u, r;
    for (; k < niter; k++) {
        double tstart = now(), t0 = tstart, t1;
        if (beta_period > 0 && k % beta_period == 0) beta *= 2.0;
        rho = filter->apply(beta, s);
        double g = 0.0;
        std::vector<double> dgdrho(nelem, 0.0);
        for (int i = 0; i < nelem; i++) {
            g += scale1 * rho[i] / (weightlimit * nelem);
            dgdrho[i] = scale1 / (weightlimit * nelem);
        }
        g -= 1.0 * scale1;
        t1 = now(); if (phase) phase[0] += t1 - t0; t0 = t1;

        u = std::vector<Vector<double> >(nnode, Vector<double>(ndof));
        std::vector<std::vector<int> > nodetoglobal(nnode, std::vector<int>(ndof, 0));
        SetDirichlet(u, nodetoglobal, ufixed);
        int KDEGREE = Renumbering(nodetoglobal);
        LILCSR<double> K(KDEGREE, KDEGREE);
        std::vector<double> F(KDEGREE, 0.0);
        for (int i = 0; i < nelem; i++) {
            double E = E1 * pow(rho[i], p) + E0 * (1.0 - pow(rho[i], p));
            std::vector<std::vector<std::pair<int, int> > > n2e;
            Matrix<double> Ke;
            element_matrix(eq, Ke, n2e, elements[i], x, E, Poisson, thick);
            Assembling(K, F, u, Ke, nodetoglobal, n2e, elements[i]);
        }
        Assembling(F, qfixed, nodetoglobal);
        t1 = now(); if (phase) phase[1] += t1 - t0; t0 = t1;
        CSR<double> Kmod(K);
        t1 = now(); if (phase) phase[2] += t1 - t0; t0 = t1;
        std::vector<double> result = ScalingCG(Kmod, F, itrmax, cgeps);
        Disassembling(u, result, nodetoglobal);
        t1 = now(); if (phase) phase[3] += t1 - t0; t0 = t1;

        RemoveBoundaryConditions(nodetoglobal);
        KDEGREE = Renumbering(nodetoglobal);
        std::vector<double> RF(KDEGREE, 0.0);
        r = std::vector<Vector<double> >(nnode, Vector<double>(ndof));
        for (int i = 0; i < nelem; i++) {
            double E = E1 * pow(rho[i], p) + E0 * (1.0 - pow(rho[i], p));
            std::vector<std::vector<std::pair<int, int> > > n2e;
            Matrix<double> Ke;
            element_matrix(eq, Ke, n2e, elements[i], x, E, Poisson, thick);
            Vector<double> Keue = Ke * ElementVector(u, n2e, elements[i]);
            Assembling(RF, Keue, nodetoglobal, n2e, elements[i]);
        }
        Disassembling(r, RF, nodetoglobal);
        double f = scale0 * std::inner_product(u.begin(), u.end(), r.begin(), 0.0);
        t1 = now(); if (phase) phase[4] += t1 - t0; t0 = t1;

        std::vector<double> dfdrho(nelem, 0.0);
        for (int i = 0; i < nelem; i++) {
            std::vector<std::vector<std::pair<int, int> > > n2e;
            Matrix<double> Ke;
            element_matrix(eq, Ke, n2e, elements[i], x, 1.0, Poisson, thick);
            Vector<double> ue = ElementVector(u, n2e, elements[i]);
            dfdrho[i] = -scale0 * p * (-E0 + E1) * pow(rho[i], p - 1.0) * (ue * (Ke * ue));
        }
        t1 = now(); if (phase) phase[5] += t1 - t0; t0 = t1;
        std::vector<double> dfds = filter->sens(beta, s, dfdrho);
        std::vector<double> dgds = filter->sens(beta, s, dgdrho);
        t1 = now(); if (phase) phase[6] += t1 - t0; t0 = t1;

        hist[4 * k + 0] = f; hist[4 * k + 1] = g; hist[4 * k + 3] = 0;
        bool conv = oc ? oc->IsConvergence(f) : (mma ? mma->IsConvergence(f) : conlin->IsConvergence(f));
        if (check_convergence && conv) {
            hist[4 * k + 2] = now() - tstart; hist[4 * k + 3] = 1;
            k++;
            break;
        }
        if (oc) {
            oc->UpdateVariables(s, f, dfds, g, dgds, [&](std::vector<double> _xkp1) {
                double gg = 0.0;
                std::vector<double> rr = filter->apply(beta, _xkp1);
                for (int i = 0; i < nelem; i++) gg += scale1 * rr[i] / (weightlimit * nelem);
                return gg - 1.0 * scale1;
            });
        } else if (mma) {
            mma->UpdateVariables(s, f, dfds, { g }, { dgds });
        } else {
            conlin->UpdateVariables(s, f, dfds, { g }, { dgds });
        }
        t1 = now(); if (phase) phase[7] += t1 - t0;
        hist[4 * k + 2] = now() - tstart;
    }
    std::copy(s.begin(), s.end(), s_io);
    if (rho_out) std::copy(rho.begin(), rho.end(), rho_out);
    for (int i = 0; i < nnode; i++) for (int d = 0; d < ndof; d++) {
        if (u_out) u_out[(size_t)i * ndof + d] = u[i](d);
        if (r_out) r_out[(size_t)i * ndof + d] = r[i](d);
    }
    delete filter; delete oc; delete mma; delete conlin;
    return k;
}

// ---- SquareMesh<T> (SquareMesh.h:62-207): numbering fixture for the product mesher ----
void ref_squaremesh(double lx, double ly, int nx, int ny, double* coords, int* conn) {
    SquareMesh<double> mesh(lx, ly, nx, ny);
    std::vector<Vector<double> > x = mesh.GenerateNodes();
    std::vector<std::vector<int> > e = mesh.GenerateElements();
    for (size_t i = 0; i < x.size(); i++) { coords[2 * i] = x[i](0); coords[2 * i + 1] = x[i](1); }
    for (size_t i = 0; i < e.size(); i++) for (int j = 0; j < 4; j++) conn[4 * i + j] = e[i][j];
}

}  // extern "C"

// ---- load vectors with the reference's own functor interface: PlaneStrainSurfaceForce / BodyForce (PlaneStrain.h:421, 503),
//      PlaneStressSurfaceForce / BodyForce (PlaneStress.h:98, 134), HeatTransferSurfaceFlux (HeatTransfer.h:76).
//      kind: 0 plane-strain, 1 plane-stress, 2 heat flux; the force density is the affine field f_i(x) = c[3i] + c[3i+1] x + c[3i+2] y ----
namespace {
struct AffineForce {
    const double* c; int ndof;
    Vector<double> operator()(Vector<double> x) const {
        Vector<double> f(ndof);
        for (int i = 0; i < ndof; i++) f(i) = c[3 * i] + c[3 * i + 1] * x(0) + c[3 * i + 2] * x(1);
        return f;
    }
};
struct AffineFlux {
    const double* c;
    double operator()(Vector<double> x) const { return c[0] + c[1] * x(0) + c[2] * x(1); }
};
template<template<class>class SF, template<class>class IC, bool BODY>
void load_vec(int kind, Vector<double>& Fe, N2E& n2e, const std::vector<int>& el, std::vector<Vector<double> >& x, const double* c, double t) {
    if (kind == 2) { HeatTransferSurfaceFlux<double, SF, IC>(Fe, n2e, el, { 0 }, x, AffineFlux{ c }, t); return; }
    AffineForce f{ c, 2 };
    if (BODY) { if (kind == 0) PlaneStrainBodyForce<double, SF, IC>(Fe, n2e, el, { 0, 1 }, x, f, t); else PlaneStressBodyForce<double, SF, IC>(Fe, n2e, el, { 0, 1 }, x, f, t); }
    else { if (kind == 0) PlaneStrainSurfaceForce<double, SF, IC>(Fe, n2e, el, { 0, 1 }, x, f, t); else PlaneStressSurfaceForce<double, SF, IC>(Fe, n2e, el, { 0, 1 }, x, f, t); }
}
}   // namespace
extern "C" {
// shape: 8 = 2Line, 9 = 3Line (surface), 1 T3, 2 T6, 3 Q4, 4 Q8 (body); quad: 9 Gauss1Line, 10 Gauss2Line, else the area rules
int ref_load_vector(int kind, int shape, int quad, int npe, const double* xe, const double* coef, double t, double* Fe_out) {
    std::vector<Vector<double> > x = make_nodes(2, npe, xe);
    std::vector<int> el(npe);
    std::iota(el.begin(), el.end(), 0);
    Vector<double> Fe;
    N2E n2e;
    switch (shape) {
        case 8: if (quad == 10) load_vec<ShapeFunction2Line, Gauss2Line, false>(kind, Fe, n2e, el, x, coef, t); else load_vec<ShapeFunction2Line, Gauss1Line, false>(kind, Fe, n2e, el, x, coef, t); break;
        case 9: if (quad == 10) load_vec<ShapeFunction3Line, Gauss2Line, false>(kind, Fe, n2e, el, x, coef, t); else load_vec<ShapeFunction3Line, Gauss1Line, false>(kind, Fe, n2e, el, x, coef, t); break;
        case SHAPE_T3: if (quad == QUAD_G3TRI) load_vec<ShapeFunction3Triangle, Gauss3Triangle, true>(kind, Fe, n2e, el, x, coef, t); else load_vec<ShapeFunction3Triangle, Gauss1Triangle, true>(kind, Fe, n2e, el, x, coef, t); break;
        case SHAPE_T6: if (quad == QUAD_G3TRI) load_vec<ShapeFunction6Triangle, Gauss3Triangle, true>(kind, Fe, n2e, el, x, coef, t); else load_vec<ShapeFunction6Triangle, Gauss1Triangle, true>(kind, Fe, n2e, el, x, coef, t); break;
        case SHAPE_Q8:
            if (quad == QUAD_G1SQ) load_vec<ShapeFunction8Square, Gauss1Square, true>(kind, Fe, n2e, el, x, coef, t);
            else if (quad == QUAD_G9SQ) load_vec<ShapeFunction8Square, Gauss9Square, true>(kind, Fe, n2e, el, x, coef, t);
            else load_vec<ShapeFunction8Square, Gauss4Square, true>(kind, Fe, n2e, el, x, coef, t);
            break;
        default:
            if (quad == QUAD_G1SQ) load_vec<ShapeFunction4Square, Gauss1Square, true>(kind, Fe, n2e, el, x, coef, t);
            else if (quad == QUAD_G9SQ) load_vec<ShapeFunction4Square, Gauss9Square, true>(kind, Fe, n2e, el, x, coef, t);
            else load_vec<ShapeFunction4Square, Gauss4Square, true>(kind, Fe, n2e, el, x, coef, t);
            break;
    }
    const int m = Fe.SIZE();
    for (int i = 0; i < m; i++) Fe_out[i] = Fe(i);
    return m;
}
}   // extern "C"
