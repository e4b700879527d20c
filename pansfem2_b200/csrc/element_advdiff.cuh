// element_advdiff.cuh -- the advection-diffusion element family: one scalar dof per node, 2-D shapes.
//
// Restates /root/reference/src/FEM/Equation/Advection.h:
//   Advection :19-43                  Ke_ab += N_a (c . grad N_b) J w
//   Diffusion :135-157                Ke_ab += k (grad N_a . grad N_b) J w
//   AdvectionSUPG :47-87              Ke_ab += tau (a . grad N_a)(a . grad N_b) J w
//   AdvectionShockCapturing :91-131   Ke_ab += tau_sc (grad N_a . grad N_b) J w
//   Mass :161-184                     Ce_ab += N_a N_b J w
//   MassSUPG :188-228                 Ce_ab += tau (a . grad N_a) N_b J w
// with, per integration point (:69-82), he = 2 / sum_i |a . grad N_i| / |a|, alpha = |a| he / (2k),
// tau = he / (2|a|) * min(alpha/3, 1), tau_sc = |a| he / 2 * min(alpha/3, 1)   (k = 0 gives alpha = inf, factor 1, as there).
// The reference forms each term as a chain of Matrix<T> products; here the (a, b) entry is written out, the terms a caller sums
// (sample_advectiondiffusion_static.cpp:42-51, ..._dynamic.cpp:57-69) are selected by a bit mask and accumulated in one pass into
// two rows: the "stiffness" group K = A + D + AS + SC and the "mass" group M = M + MS, which the time discretisation weights apart.
#pragma once
#include "element_generic.cuh"

namespace pf2 {

enum { ADV_A = PF2_ADV_ADVECTION, ADV_D = PF2_ADV_DIFFUSION, ADV_S = PF2_ADV_SUPG, ADV_SC = PF2_ADV_SHOCK, ADV_M = PF2_ADV_MASS,
       ADV_MS = PF2_ADV_MASS_SUPG };

struct AdvSpec {
    int quad, terms;        // PF2_QUAD_* and the PF2_ADV_* mask
    double ax, ay, k;       // advection velocity, diffusion coefficient
};

// rows of local node `a`: accK[b] (stiffness group), accM[b] (mass group)
template <int SHAPE>
PF2_HD void advdiff_rows(const double (&X)[ShapeTraits<SHAPE>::NPE][2], int a, const AdvSpec& sp, double (&accK)[ShapeTraits<SHAPE>::NPE],
                         double (&accM)[ShapeTraits<SHAPE>::NPE]) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE;
    static_assert(ShapeTraits<SHAPE>::DIM == 2, "the advection-diffusion routines are 2-D");
#pragma unroll
    for (int b = 0; b < NPE; b++) { accK[b] = 0.0; accM[b] = 0.0; }
    const int terms = sp.terms;
    const double ax = sp.ax, ay = sp.ay, k = sp.k;
    const int ng = quad_count(sp.quad);
#pragma unroll 1
    for (int q = 0; q < ng; q++) {
        double r[3], wq, det, g[2][NPE], N[NPE];
        quad_point(sp.quad, q, r, wq);
        shape_grad<SHAPE>(X, r, g, det);
        shape_n<SHAPE>(r, N);
        const double w = det * wq;
        double tau = 0.0, tausc = 0.0;
        if (terms & (ADV_S | ADV_SC | ADV_MS)) {
            const double norm = sqrt(ax * ax + ay * ay);
            double sum = 0.0;
#pragma unroll
            for (int n = 0; n < NPE; n++) sum += fabs(ax * g[0][n] + ay * g[1][n]);
            const double he = 2.0 * norm / sum, alpha = 0.5 * norm * he / k;       // he = 2 / sum_i (|a . grad N_i| / |a|)
            const double f = (alpha <= 3.0) ? alpha / 3.0 : 1.0;
            tau = 0.5 * he / norm * f;
            tausc = 0.5 * norm * he * f;
        }
        double Na = N[0], gax = g[0][0], gay = g[1][0];
#pragma unroll
        for (int n = 1; n < NPE; n++) if (n == a) { Na = N[n]; gax = g[0][n]; gay = g[1][n]; }
        const double ca = ax * gax + ay * gay;
#pragma unroll
        for (int b = 0; b < NPE; b++) {
            const double cb = ax * g[0][b] + ay * g[1][b], dab = gax * g[0][b] + gay * g[1][b];
            double kk = 0.0, mm = 0.0;
            if (terms & ADV_A) kk += Na * cb;
            if (terms & ADV_D) kk += k * dab;
            if (terms & ADV_S) kk += tau * ca * cb;
            if (terms & ADV_SC) kk += tausc * dab;
            if (terms & ADV_M) mm += Na * N[b];
            if (terms & ADV_MS) mm += tau * ca * N[b];
            accK[b] += kk * w;
            accM[b] += mm * w;
        }
    }
}

}  // namespace pf2
