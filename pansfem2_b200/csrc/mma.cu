// mma.cu -- Svanberg's MMA with the primal-dual interior-point subproblem solver, on the device.
//
// Replaces MMA<T> (Optimize/Solver/MMA.h): ctor :64-88, SetParameters :92-100, IsConvergence :108-113,
// UpdateVariables :117-419, KKTNorm :423-461, solvels :465-509.
//
// Everything of length n lives in HBM and is touched only by grid-wide kernels; the (m+1)x(m+1) reduced system
// (MMA.h:260-291, the n > m branch) is formed from device-side reductions and solved by the last CTA of the kernel
// that produced the sums (dense elimination with partial pivoting as solvels).  One Newton step = three passes:
//   newton1 : Dx, delta~x, G and the reduced-system sums + KKT norm of the current point        ~ (12+2m)*8 B/variable
//   newton2 : dx, dxi, deta and the maximal step (max-reduction)                               ~ (11+2m)*8 B/variable
//   trial   : the trial point and its KKT norm (line search, MMA.h:371-391); repeated only on rejection
// The host reads back ~100 bytes per Newton step (accept flag, eps) to drive the data-dependent loops.
// m <= 4 constraints (the SIMP drivers use m = 1); the tiny-problem branch n <= m (MMA.h:292-330) is handled by a
// single-thread kernel that restates the reference loop literally.
#include "types.cuh"

namespace pf2 {

int dist_allreduce(pf2_dist* d, double* dev, int count);
int dist_allreduce_max(pf2_dist* d, double* dev, int count);

__device__ __forceinline__ double sq(double x) { return x * x; }

// dense elimination with partial pivoting, as MMA<T>::solvels (MMA.h:465-509); N <= kMaxM+1
__device__ void solvels_dev(int N, double (*A)[kMaxM + 1], double* b, double* x) {
    for (int i = 0; i < N - 1; i++) {
        double pivot = fabs(A[i][i]);
        int pi = i;
        for (int j = i + 1; j < N; j++) if (pivot < fabs(A[j][i])) { pivot = fabs(A[j][i]); pi = j; }
        if (pi != i) {
            double tmp = b[i]; b[i] = b[pi]; b[pi] = tmp;
            for (int j = i; j < N; j++) { tmp = A[i][j]; A[i][j] = A[pi][j]; A[pi][j] = tmp; }
        }
        for (int j = i + 1; j < N; j++) {
            for (int k = i + 1; k < N; k++) A[j][k] -= A[i][k] * A[j][i] / A[i][i];
            b[j] -= b[i] * A[j][i] / A[i][i];
        }
    }
    for (int i = N - 1; i >= 0; i--) {
        x[i] = b[i];
        for (int j = N - 1; j > i; j--) x[i] -= x[j] * A[i][j];
        x[i] /= A[i][i];
    }
}

// the m-sized part of KKTNorm (MMA.h:449-458) given g_i = sum_j p_ij/(U-x) + q_ij/(x-L)
template <int M>
__device__ double kkt_small(const MmaSmall* S, const double* y, const double* lam, const double* s, const double* mu,
                            double z, double zeta, const double* gs, double eps) {
    double norm = 0.0, la = 0.0;
#pragma unroll
    for (int i = 0; i < M; i++) {
        norm += sq(S->c[i] + S->d[i] * y[i] - lam[i] - mu[i]);
        norm += sq(gs[i] - S->a[i] * z - y[i] + s[i] - S->b[i]);
        norm += sq(mu[i] * y[i] - eps);
        norm += sq(lam[i] * s[i] - eps);
        la += lam[i] * S->a[i];
    }
    norm += sq(S->a0 - zeta - la);
    norm += sq(zeta * z - eps);
    return norm;
}

// ---- set-up pass: asymptotes, move limits, p0 q0 p q b, starting point (MMA.h:119-197) -----------------------------
// CONLIN = true turns the same machinery into CONLIN<T> (Optimize/Solver/CONLIN.h:89-373): the convex approximation
// p*x + q/x replaces p/(U-x) + q/(x-L) (no asymptotes), with p = max(df,0), q = max(-df,0)*xk^2 (:100-122) and the
// move limits alpha = max(xmin, xk - move*w), beta = min(xmax, xk + move*w) (:92-96).
template <int M, bool CONLIN>
__global__ void __launch_bounds__(kThreads)
mma_setup_kernel(int lo, int hi, int n, int k, MmaParams P, const double* __restrict__ xk, const double* __restrict__ xkm1,
                 const double* __restrict__ xkm2, const double* __restrict__ xmin, const double* __restrict__ xmax,
                 const double* __restrict__ dfdx, const double* __restrict__ dgdx, double* __restrict__ L, double* __restrict__ U,
                 double* __restrict__ alpha, double* __restrict__ beta, double* __restrict__ p0, double* __restrict__ q0,
                 double* __restrict__ p, double* __restrict__ q, double* __restrict__ x, double* __restrict__ gsi,
                 double* __restrict__ ita, MmaSmall* S, const double* __restrict__ gval, double* partials, unsigned int* ticket) {
    double bs[M];
#pragma unroll
    for (int i = 0; i < M; i++) bs[i] = 0.0;
    for (int j = lo + blockIdx.x * blockDim.x + threadIdx.x; j < hi; j += gridDim.x * blockDim.x) {
        const double xj = xk[j], w = xmax[j] - xmin[j];
        if (CONLIN) {
            const double al = fmax(xmin[j], xj - P.move * w), be = fmin(xmax[j], xj + P.move * w);
            alpha[j] = al; beta[j] = be;
            const double df = dfdx[j];
            p0[j] = fmax(df, 0.0);
            q0[j] = fmax(-df, 0.0) * sq(xj);
#pragma unroll
            for (int i = 0; i < M; i++) {
                const double dg = dgdx[(size_t)i * n + j];
                const double pij = fmax(dg, 0.0), qij = fmax(-dg, 0.0) * sq(xj);
                p[(size_t)i * n + j] = pij; q[(size_t)i * n + j] = qij;
                bs[i] += pij * xj + qij / xj;
            }
            const double x0 = 0.5 * (al + be);
            x[j] = x0;
            gsi[j] = fmax(1.0, 1.0 / (x0 - al));
            ita[j] = fmax(1.0, 1.0 / (be - x0));
            continue;
        }
        double Lj, Uj;
        if (k < 2) {
            Lj = xj - P.asyinit * w; Uj = xj + P.asyinit * w;
        } else {
            const double x1 = xkm1[j];
            const double tmp = (xj - x1) * (x1 - xkm2[j]);
            if (tmp < 0.0) { Lj = xj - P.asydecr * (x1 - L[j]); Uj = xj + P.asydecr * (U[j] - x1); }
            else if (tmp > 0.0) { Lj = xj - P.asyincr * (x1 - L[j]); Uj = xj + P.asyincr * (U[j] - x1); }
            else { Lj = xj - (x1 - L[j]); Uj = xj + (U[j] - x1); }
        }
        Lj = fmin(fmax(xj - 10.0 * w, Lj), xj - 0.01 * w);
        Uj = fmin(fmax(xj + 0.01 * w, Uj), xj + 10.0 * w);
        L[j] = Lj; U[j] = Uj;
        const double al = fmax(fmax(xmin[j], Lj + P.albefa * (xj - Lj)), xj - P.move * w);
        const double be = fmin(fmin(xmax[j], Uj - P.albefa * (Uj - xj)), xj + P.move * w);
        alpha[j] = al; beta[j] = be;
        const double ux = Uj - xj, xl = xj - Lj, r = P.raa0 / w;
        const double df = dfdx[j];
        const double dfp = fmax(df, 0.0), dfm = fmax(-df, 0.0);
        p0[j] = sq(ux) * (1.001 * dfp + 0.001 * dfm + r);
        q0[j] = sq(xl) * (0.001 * dfp + 1.001 * dfm + r);
#pragma unroll
        for (int i = 0; i < M; i++) {
            const double dg = dgdx[(size_t)i * n + j];
            const double dgp = fmax(dg, 0.0), dgm = fmax(-dg, 0.0);
            const double pij = sq(ux) * (1.001 * dgp + 0.001 * dgm + r);
            const double qij = sq(xl) * (0.001 * dgp + 1.001 * dgm + r);
            p[(size_t)i * n + j] = pij; q[(size_t)i * n + j] = qij;
            bs[i] += pij / ux + qij / xl;
        }
        const double x0 = 0.5 * (al + be);
        x[j] = x0;
        gsi[j] = fmax(1.0, 1.0 / (x0 - al));
        ita[j] = fmax(1.0, 1.0 / (be - x0));
    }
    if (grid_sum_last<M>(bs, partials, ticket) && threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < M; i++) S->red[i] = bs[i];
    }
}
// the m-sized tail of the set-up pass; in a partitioned run the sums in S->red have been allreduced in between
template <int M>
__global__ void mma_setup_small_kernel(MmaSmall* S, const double* __restrict__ gval) {
    for (int i = 0; i < M; i++) {
        S->b[i] = -gval[i] + S->red[i];
        S->y[i] = 1.0; S->lam[i] = 1.0; S->s[i] = 1.0; S->mu[i] = fmax(1.0, 0.5 * S->c[i]);
    }
    S->z = 1.0; S->zeta = 1.0; S->eps = 1.0; S->newton = 0; S->halvings = 0; S->accept = 0; S->ll = 0;
}

// ---- Newton pass 1 (MMA.h:201-291 + the m-sized updates :332-343,:353-359 + KKTNorm of the current point :361) ------
template <int M, bool CONLIN>
__global__ void __launch_bounds__(kThreads)
mma_newton1_kernel(int lo, int hi, int n, const double* __restrict__ x, const double* __restrict__ L, const double* __restrict__ U,
                   const double* __restrict__ alpha, const double* __restrict__ beta, const double* __restrict__ p0,
                   const double* __restrict__ q0, const double* __restrict__ p, const double* __restrict__ q,
                   const double* __restrict__ gsi, const double* __restrict__ ita, double* __restrict__ Dx,
                   double* __restrict__ dtx, MmaSmall* S, double* partials, unsigned int* ticket) {
    constexpr int NT = M * M + 2 * M + 1;
    double v[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) v[t] = 0.0;
    double lam[M];
#pragma unroll
    for (int i = 0; i < M; i++) lam[i] = S->lam[i];
    const double eps = S->eps;
    for (int j = lo + blockIdx.x * blockDim.x + threadIdx.x; j < hi; j += gridDim.x * blockDim.x) {
        const double xj = x[j], ux = CONLIN ? 1.0 : U[j] - xj, xl = CONLIN ? xj : xj - L[j], xa = xj - alpha[j], bx = beta[j] - xj;
        double pl = p0[j], ql = q0[j], G[M];
        const double iux2 = 1.0 / sq(ux), ixl2 = 1.0 / sq(xl);
#pragma unroll
        for (int i = 0; i < M; i++) {
            const double pij = p[(size_t)i * n + j], qij = q[(size_t)i * n + j];
            pl += lam[i] * pij; ql += lam[i] * qij;
            G[i] = CONLIN ? pij - qij / sq(xj) : pij / sq(ux) - qij / sq(xl);
            v[M * M + M + i] += CONLIN ? pij * xj + qij / xj : pij / ux + qij / xl;
        }
        const double gs = gsi[j], it = ita[j];
        const double dxj = (CONLIN ? 2.0 * ql / (sq(xj) * xj) : 2.0 * pl / (sq(ux) * ux) + 2.0 * ql / (sq(xl) * xl)) + gs / xa + it / bx;
        const double grad = CONLIN ? pl - ql / sq(xj) : pl * iux2 - ql * ixl2;
        const double dt = grad - eps / xa + eps / bx;
        Dx[j] = dxj; dtx[j] = dt;
#pragma unroll
        for (int i = 0; i < M; i++) {
#pragma unroll
            for (int l = 0; l < M; l++) v[i * M + l] += G[i] * G[l] / dxj;
            v[M * M + i] += G[i] * dt / dxj;
        }
        v[NT - 1] += sq(grad - gs + it) + sq(gs * xa - eps) + sq(it * bx - eps);
    }
    if (grid_sum_last<NT>(v, partials, ticket) && threadIdx.x == 0) {
#pragma unroll
        for (int t = 0; t < NT; t++) S->red[t] = v[t];
    }
}
// reduced (m+1)x(m+1) system and the m-sized Newton updates from the (all)reduced sums in S->red
template <int M>
__global__ void mma_newton1_small_kernel(MmaSmall* S) {
    constexpr int NT = M * M + 2 * M + 1;
    double v[NT];
    for (int t = 0; t < NT; t++) v[t] = S->red[t];
    const double eps = S->eps;
    {
        double Dy[M], Dlam[M], dty[M], dtlam[M], Dlamy[M], dtlamy[M], gsum[M];
        double la = 0.0;
#pragma unroll
        for (int i = 0; i < M; i++) {
            gsum[i] = v[M * M + M + i];
            Dy[i] = S->d[i] + S->mu[i] / S->y[i];
            Dlam[i] = S->s[i] / S->lam[i];
            dty[i] = S->c[i] + S->d[i] * S->y[i] - S->lam[i] - eps / S->y[i];
            la += S->lam[i] * S->a[i];
        }
        const double dtz = S->a0 - eps / S->z - la;
#pragma unroll
        for (int i = 0; i < M; i++) {
            dtlam[i] = -S->a[i] * S->z - S->y[i] - S->b[i] + eps / S->lam[i] + gsum[i];
            Dlamy[i] = Dlam[i] + 1.0 / Dy[i];
            dtlamy[i] = dtlam[i] + dty[i] / Dy[i];
        }
        double A[kMaxM + 1][kMaxM + 1], B[kMaxM + 1], sol[kMaxM + 1];
        for (int i = 0; i <= M; i++) for (int l = 0; l <= M; l++) A[i][l] = 0.0;
        for (int i = 0; i < M; i++) {
            for (int l = 0; l < M; l++) A[i][l] = v[i * M + l];
            A[i][i] += Dlamy[i];
            A[i][M] = S->a[i];
            A[M][i] = S->a[i];
            B[i] = dtlamy[i] - v[M * M + i];
        }
        A[M][M] = -S->zeta / S->z;
        B[M] = dtz;
        solvels_dev(M + 1, A, B, sol);
        double tymax = 0.0;
        for (int i = 0; i < M; i++) {
            const double dl = sol[i];
            S->dlam[i] = dl;
            const double dyi = dl / Dy[i] - dty[i] / Dy[i];
            S->dy[i] = dyi;
            S->dmu[i] = -S->mu[i] * dyi / S->y[i] - S->mu[i] + eps / S->y[i];
            S->ds[i] = -S->s[i] * dl / S->lam[i] - S->s[i] + eps / S->lam[i];
            const double t = fmax(fmax(-1.01 * dyi / S->y[i], -1.01 * dl / S->lam[i]), fmax(-1.01 * S->dmu[i] / S->mu[i], -1.01 * S->ds[i] / S->s[i]));
            if (tymax < t) tymax = t;
        }
        S->dz = sol[M];
        S->dzeta = -S->zeta * S->dz / S->z - S->zeta + eps / S->z;
        S->tymax = tymax;
        S->dwl = sqrt(v[NT - 1] + kkt_small<M>(S, S->y, S->lam, S->s, S->mu, S->z, S->zeta, gsum, eps));
    }
}

// ---- Newton pass 2: dx, dxi, deta, maximal step (MMA.h:282-287, 338-360) ----------------------------------------------
template <int M, bool CONLIN>
__global__ void __launch_bounds__(kThreads)
mma_newton2_kernel(int lo, int hi, int n, const double* __restrict__ x, const double* __restrict__ L, const double* __restrict__ U,
                   const double* __restrict__ alpha, const double* __restrict__ beta, const double* __restrict__ p,
                   const double* __restrict__ q, const double* __restrict__ gsi, const double* __restrict__ ita,
                   const double* __restrict__ Dx, const double* __restrict__ dtx, double* __restrict__ dx,
                   double* __restrict__ dgsi, double* __restrict__ dita, MmaSmall* S, double* partials, unsigned int* ticket) {
    double dlam[M];
#pragma unroll
    for (int i = 0; i < M; i++) dlam[i] = S->dlam[i];
    const double eps = S->eps;
    double txmax = 0.0;
    for (int j = lo + blockIdx.x * blockDim.x + threadIdx.x; j < hi; j += gridDim.x * blockDim.x) {
        const double xj = x[j], ux = CONLIN ? 1.0 : U[j] - xj, xl = CONLIN ? xj : xj - L[j], xa = xj - alpha[j], bx = beta[j] - xj, dxx = Dx[j];
        double d = -dtx[j] / dxx;
#pragma unroll
        for (int i = 0; i < M; i++) {
            const double Gi = CONLIN ? p[(size_t)i * n + j] - q[(size_t)i * n + j] / sq(xj) : p[(size_t)i * n + j] / sq(ux) - q[(size_t)i * n + j] / sq(xl);
            d -= Gi * dlam[i] / dxx;
        }
        const double gs = gsi[j], it = ita[j];
        const double dg = -gs * d / xa - gs + eps / xa;
        const double di = it * d / bx - it + eps / bx;
        dx[j] = d; dgsi[j] = dg; dita[j] = di;
        const double t = fmax(fmax(-1.01 * d / xa, 1.01 * d / bx), fmax(-1.01 * dg / gs, -1.01 * di / it));
        if (txmax < t) txmax = t;
    }
    if (grid_max_last(txmax, partials, ticket) && threadIdx.x == 0) S->red[0] = txmax;
}
__global__ void mma_newton2_small_kernel(MmaSmall* S) {
    const double txmax = S->red[0];
    const double m1 = fmax(fmax(1.0, fmax(txmax, 0.0)), fmax(S->tymax, -1.01 * S->dz / S->z));
    S->tau = 1.0 / fmax(m1, -1.01 * S->dzeta / S->zeta);
    S->ll = 0; S->accept = 0;
}

// ---- line-search trial (MMA.h:371-410) ---------------------------------------------------------------------------------
template <int M, bool CONLIN>
__global__ void __launch_bounds__(kThreads)
mma_trial_kernel(int lo, int hi, int n, const double* __restrict__ x, const double* __restrict__ L, const double* __restrict__ U,
                 const double* __restrict__ alpha, const double* __restrict__ beta, const double* __restrict__ p0,
                 const double* __restrict__ q0, const double* __restrict__ p, const double* __restrict__ q,
                 const double* __restrict__ gsi, const double* __restrict__ ita, const double* __restrict__ dx,
                 const double* __restrict__ dgsi, const double* __restrict__ dita, double* __restrict__ xn,
                 double* __restrict__ gsin, double* __restrict__ itan, MmaSmall* S, double* partials, unsigned int* ticket) {
    if (S->accept) return;
    constexpr int NT = M + 1;
    double v[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) v[t] = 0.0;
    const double tau = S->tau, eps = S->eps;
    double lamn[M];
#pragma unroll
    for (int i = 0; i < M; i++) lamn[i] = S->lam[i] + tau * S->dlam[i];
    for (int j = lo + blockIdx.x * blockDim.x + threadIdx.x; j < hi; j += gridDim.x * blockDim.x) {
        const double xj = x[j] + tau * dx[j], gs = gsi[j] + tau * dgsi[j], it = ita[j] + tau * dita[j];
        xn[j] = xj; gsin[j] = gs; itan[j] = it;
        const double ux = CONLIN ? 1.0 : U[j] - xj, xl = CONLIN ? xj : xj - L[j];
        double pl = p0[j], ql = q0[j];
#pragma unroll
        for (int i = 0; i < M; i++) {
            const double pij = p[(size_t)i * n + j], qij = q[(size_t)i * n + j];
            pl += lamn[i] * pij; ql += lamn[i] * qij;
            v[i] += CONLIN ? pij * xj + qij / xj : pij / ux + qij / xl;
        }
        v[M] += sq((CONLIN ? pl - ql / sq(xj) : pl / sq(ux) - ql / sq(xl)) - gs + it) + sq(gs * (xj - alpha[j]) - eps) + sq(it * (beta[j] - xj) - eps);
    }
    if (grid_sum_last<NT>(v, partials, ticket) && threadIdx.x == 0) {
#pragma unroll
        for (int t = 0; t < NT; t++) S->red[t] = v[t];
    }
}
template <int M>
__global__ void mma_trial_small_kernel(MmaSmall* S) {
    if (S->accept) return;
    constexpr int NT = M + 1;
    double v[NT];
    for (int t = 0; t < NT; t++) v[t] = S->red[t];
    const double tau = S->tau, eps = S->eps;
    double lamn[M];
    for (int i = 0; i < M; i++) lamn[i] = S->lam[i] + tau * S->dlam[i];
    {
        double yn[M], sn[M], mun[M], gsum[M];
#pragma unroll
        for (int i = 0; i < M; i++) {
            yn[i] = S->y[i] + tau * S->dy[i]; mun[i] = S->mu[i] + tau * S->dmu[i]; sn[i] = S->s[i] + tau * S->ds[i];
            gsum[i] = v[i];
        }
        const double zn = S->z + tau * S->dz, zetan = S->zeta + tau * S->dzeta;
        const double dwl1 = sqrt(v[M] + kkt_small<M>(S, yn, lamn, sn, mun, zn, zetan, gsum, eps));
        S->dwl1 = dwl1;
        bool accept = dwl1 < S->dwl;
        if (!accept) {
            if (S->ll >= 49) accept = true;           // the 50th trial is kept whatever its norm (MMA.h:371-391)
            else { S->ll = S->ll + 1; S->tau = tau * 0.5; S->halvings = S->halvings + 1; }
        }
        if (accept) {
#pragma unroll
            for (int i = 0; i < M; i++) { S->y[i] = yn[i]; S->lam[i] = lamn[i]; S->mu[i] = mun[i]; S->s[i] = sn[i]; }
            S->z = zn; S->zeta = zetan;
            S->newton = S->newton + 1;
            if (dwl1 < 0.9 * eps) S->eps = eps * 0.1;      // MMA.h:405-410
            S->accept = 1;
        }
    }
}

// ---- tiny problems, n <= m (MMA.h:292-330): one thread restates the reference loop ---------------------------------------
constexpr int kTinyN = 4;
__global__ void mma_tiny_kernel(int n, int m, int k, MmaParams P, double* xk, double* xkm1, double* xkm2, const double* xmin,
                                const double* xmax, const double* dfdx, const double* dgdx, const double* gval, double* Lg,
                                double* Ug, MmaSmall* S) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double L[kTinyN], U[kTinyN], alpha[kTinyN], beta[kTinyN], p0[kTinyN], q0[kTinyN], p[kMaxM][kTinyN], q[kMaxM][kTinyN], b[kMaxM];
    for (int j = 0; j < n; j++) {
        const double w = xmax[j] - xmin[j];
        if (k < 2) { L[j] = xk[j] - P.asyinit * w; U[j] = xk[j] + P.asyinit * w; }
        else {
            const double tmp = (xk[j] - xkm1[j]) * (xkm1[j] - xkm2[j]);
            const double fac = tmp < 0.0 ? P.asydecr : (tmp > 0.0 ? P.asyincr : 1.0);
            L[j] = xk[j] - fac * (xkm1[j] - Lg[j]); U[j] = xk[j] + fac * (Ug[j] - xkm1[j]);
        }
        L[j] = fmin(fmax(xk[j] - 10.0 * w, L[j]), xk[j] - 0.01 * w);
        U[j] = fmin(fmax(xk[j] + 0.01 * w, U[j]), xk[j] + 10.0 * w);
        Lg[j] = L[j]; Ug[j] = U[j];
        alpha[j] = fmax(fmax(xmin[j], L[j] + P.albefa * (xk[j] - L[j])), xk[j] - P.move * w);
        beta[j] = fmin(fmin(xmax[j], U[j] - P.albefa * (U[j] - xk[j])), xk[j] + P.move * w);
        const double dp = fmax(dfdx[j], 0.0), dm = fmax(-dfdx[j], 0.0);
        p0[j] = sq(U[j] - xk[j]) * (1.001 * dp + 0.001 * dm + P.raa0 / w);
        q0[j] = sq(xk[j] - L[j]) * (0.001 * dp + 1.001 * dm + P.raa0 / w);
    }
    for (int i = 0; i < m; i++) {
        b[i] = -gval[i];
        for (int j = 0; j < n; j++) {
            const double w = xmax[j] - xmin[j];
            const double dp = fmax(dgdx[i * n + j], 0.0), dm = fmax(-dgdx[i * n + j], 0.0);
            p[i][j] = sq(U[j] - xk[j]) * (1.001 * dp + 0.001 * dm + P.raa0 / w);
            q[i][j] = sq(xk[j] - L[j]) * (0.001 * dp + 1.001 * dm + P.raa0 / w);
            b[i] += p[i][j] / (U[j] - xk[j]) + q[i][j] / (xk[j] - L[j]);
        }
        S->b[i] = b[i];
    }
    double eps = 1.0, z = 1.0, zeta = 1.0;
    double x[kTinyN], gsi[kTinyN], ita[kTinyN], y[kMaxM], lam[kMaxM], s[kMaxM], mu[kMaxM];
    for (int i = 0; i < m; i++) { y[i] = 1.0; lam[i] = 1.0; s[i] = 1.0; mu[i] = fmax(1.0, 0.5 * S->c[i]); }
    for (int j = 0; j < n; j++) { x[j] = 0.5 * (alpha[j] + beta[j]); gsi[j] = fmax(1.0, 1.0 / (x[j] - alpha[j])); ita[j] = fmax(1.0, 1.0 / (beta[j] - x[j])); }
    auto kkt = [&](const double* X, const double* Y, double Z, const double* LAM, const double* GSI, const double* ITA, const double* MU,
                   double ZETA, const double* SS) {
        double norm = 0.0, g[kMaxM], la = 0.0;
        for (int i = 0; i < m; i++) g[i] = 0.0;
        for (int j = 0; j < n; j++) {
            double pl = p0[j], ql = q0[j];
            for (int i = 0; i < m; i++) { pl += LAM[i] * p[i][j]; ql += LAM[i] * q[i][j]; g[i] += p[i][j] / (U[j] - X[j]) + q[i][j] / (X[j] - L[j]); }
            norm += sq(pl / sq(U[j] - X[j]) - ql / sq(X[j] - L[j]) - GSI[j] + ITA[j]);
            norm += sq(GSI[j] * (X[j] - alpha[j]) - eps);
            norm += sq(ITA[j] * (beta[j] - X[j]) - eps);
        }
        for (int i = 0; i < m; i++) {
            norm += sq(S->c[i] + S->d[i] * Y[i] - LAM[i] - MU[i]);
            norm += sq(g[i] - S->a[i] * Z - Y[i] + SS[i] - b[i]);
            norm += sq(MU[i] * Y[i] - eps);
            norm += sq(LAM[i] * SS[i] - eps);
            la += LAM[i] * S->a[i];
        }
        norm += sq(S->a0 - ZETA - la);
        norm += sq(ZETA * Z - eps);
        return sqrt(norm);
    };
    int newton = 0, halvings = 0;
    for (int l = 0; eps > 1.0e-7 && l < 100000; l++) {
        double pl[kTinyN], ql[kTinyN], G[kMaxM][kTinyN], Dx[kTinyN], dtx[kTinyN];
        for (int j = 0; j < n; j++) {
            pl[j] = p0[j]; ql[j] = q0[j];
            for (int i = 0; i < m; i++) { pl[j] += lam[i] * p[i][j]; ql[j] += lam[i] * q[i][j]; G[i][j] = p[i][j] / sq(U[j] - x[j]) - q[i][j] / sq(x[j] - L[j]); }
            Dx[j] = 2.0 * pl[j] / (sq(U[j] - x[j]) * (U[j] - x[j])) + 2.0 * ql[j] / (sq(x[j] - L[j]) * (x[j] - L[j])) + gsi[j] / (x[j] - alpha[j]) + ita[j] / (beta[j] - x[j]);
            dtx[j] = pl[j] / sq(U[j] - x[j]) - ql[j] / sq(x[j] - L[j]) - eps / (x[j] - alpha[j]) + eps / (beta[j] - x[j]);
        }
        double Dy[kMaxM], Dlam[kMaxM], dty[kMaxM], dtlam[kMaxM], Dlamy[kMaxM], dtlamy[kMaxM], la = 0.0;
        for (int i = 0; i < m; i++) {
            Dy[i] = S->d[i] + mu[i] / y[i]; Dlam[i] = s[i] / lam[i];
            dty[i] = S->c[i] + S->d[i] * y[i] - lam[i] - eps / y[i];
            la += lam[i] * S->a[i];
        }
        const double dtz = S->a0 - eps / z - la;
        for (int i = 0; i < m; i++) {
            dtlam[i] = -S->a[i] * z - y[i] - b[i] + eps / lam[i];
            for (int j = 0; j < n; j++) dtlam[i] += p[i][j] / (U[j] - x[j]) + q[i][j] / (x[j] - L[j]);
            Dlamy[i] = Dlam[i] + 1.0 / Dy[i];
            dtlamy[i] = dtlam[i] + dty[i] / Dy[i];
        }
        double A[kMaxM + 1][kMaxM + 1], B[kMaxM + 1], sol[kMaxM + 1], dx[kTinyN], dlam[kMaxM], dz;
        for (int i = 0; i <= kMaxM; i++) for (int j = 0; j <= kMaxM; j++) A[i][j] = 0.0;
        if (n > m) {
            for (int ii = 0; ii < m; ii++) {
                for (int jj = 0; jj < m; jj++) for (int kk = 0; kk < n; kk++) A[ii][jj] += G[ii][kk] * G[jj][kk] / Dx[kk];
                A[ii][ii] += Dlamy[ii]; A[ii][m] = S->a[ii]; A[m][ii] = S->a[ii];
            }
            A[m][m] = -zeta / z;
            for (int ii = 0; ii < m; ii++) { B[ii] = dtlamy[ii]; for (int jj = 0; jj < n; jj++) B[ii] -= G[ii][jj] * dtx[jj] / Dx[jj]; }
            B[m] = dtz;
            solvels_dev(m + 1, A, B, sol);
            for (int i = 0; i < m; i++) dlam[i] = sol[i];
            dz = sol[m];
            for (int j = 0; j < n; j++) { dx[j] = -dtx[j] / Dx[j]; for (int i = 0; i < m; i++) dx[j] -= G[i][j] * dlam[i] / Dx[j]; }
        } else {
            for (int ii = 0; ii < n; ii++) {
                for (int jj = 0; jj < n; jj++) for (int kk = 0; kk < m; kk++) A[ii][jj] += G[kk][ii] * G[kk][jj] / Dlamy[kk];
                A[ii][ii] += Dx[ii];
                for (int jj = 0; jj < m; jj++) {
                    A[ii][n] -= G[jj][ii] * S->a[jj] / Dlamy[jj];
                    A[n][ii] -= G[jj][ii] * S->a[jj] / Dlamy[jj];
                    A[n][n] += S->a[jj] * S->a[jj] / Dlamy[jj];
                }
            }
            A[n][n] += zeta / z;
            for (int ii = 0; ii < n; ii++) { B[ii] = -dtx[ii]; for (int jj = 0; jj < m; jj++) B[ii] -= G[jj][ii] * dtlamy[jj] / Dlamy[jj]; }
            B[n] = -dtz;
            for (int jj = 0; jj < m; jj++) B[n] += S->a[jj] * dtlamy[jj] / Dlamy[jj];
            solvels_dev(n + 1, A, B, sol);
            for (int j = 0; j < n; j++) dx[j] = sol[j];
            dz = sol[n];
            for (int i = 0; i < m; i++) {
                dlam[i] = -S->a[i] * dz / Dlamy[i] + dtlamy[i] / Dlamy[i];
                for (int j = 0; j < n; j++) dlam[i] += G[i][j] * dx[j] / Dlamy[i];
            }
        }
        double dy[kMaxM], dmu[kMaxM], ds[kMaxM], dgsi[kTinyN], dita[kTinyN];
        for (int i = 0; i < m; i++) {
            dy[i] = dlam[i] / Dy[i] - dty[i] / Dy[i];
            dmu[i] = -mu[i] * dy[i] / y[i] - mu[i] + eps / y[i];
            ds[i] = -s[i] * dlam[i] / lam[i] - s[i] + eps / lam[i];
        }
        for (int j = 0; j < n; j++) {
            dgsi[j] = -gsi[j] * dx[j] / (x[j] - alpha[j]) - gsi[j] + eps / (x[j] - alpha[j]);
            dita[j] = ita[j] * dx[j] / (beta[j] - x[j]) - ita[j] + eps / (beta[j] - x[j]);
        }
        const double dzeta = -zeta * dz / z - zeta + eps / z;
        double txmax = 0.0, tymax = 0.0;
        for (int j = 0; j < n; j++) {
            const double t = fmax(fmax(-1.01 * dx[j] / (x[j] - alpha[j]), 1.01 * dx[j] / (beta[j] - x[j])), fmax(-1.01 * dgsi[j] / gsi[j], -1.01 * dita[j] / ita[j]));
            if (txmax < t) txmax = t;
        }
        for (int i = 0; i < m; i++) {
            const double t = fmax(fmax(-1.01 * dy[i] / y[i], -1.01 * dlam[i] / lam[i]), fmax(-1.01 * dmu[i] / mu[i], -1.01 * ds[i] / s[i]));
            if (tymax < t) tymax = t;
        }
        double tau = 1.0 / fmax(fmax(fmax(1.0, txmax), fmax(tymax, -1.01 * dz / z)), -1.01 * dzeta / zeta);
        const double dwl = kkt(x, y, z, lam, gsi, ita, mu, zeta, s);
        double xn[kTinyN], gsin[kTinyN], itan[kTinyN], yn[kMaxM], lamn[kMaxM], mun[kMaxM], sn[kMaxM], zn = z, zetan = zeta, dwl1 = 0.0;
        for (int ll = 0; ll < 50; ll++) {
            for (int j = 0; j < n; j++) { xn[j] = x[j] + tau * dx[j]; gsin[j] = gsi[j] + tau * dgsi[j]; itan[j] = ita[j] + tau * dita[j]; }
            for (int i = 0; i < m; i++) { yn[i] = y[i] + tau * dy[i]; lamn[i] = lam[i] + tau * dlam[i]; mun[i] = mu[i] + tau * dmu[i]; sn[i] = s[i] + tau * ds[i]; }
            zn = z + tau * dz; zetan = zeta + tau * dzeta;
            dwl1 = kkt(xn, yn, zn, lamn, gsin, itan, mun, zetan, sn);
            if (dwl1 < dwl) break;
            tau *= 0.5;
            halvings++;
        }
        for (int j = 0; j < n; j++) { x[j] = xn[j]; gsi[j] = gsin[j]; ita[j] = itan[j]; }
        for (int i = 0; i < m; i++) { y[i] = yn[i]; lam[i] = lamn[i]; mu[i] = mun[i]; s[i] = sn[i]; }
        z = zn; zeta = zetan;
        newton++;
        if (dwl1 < 0.9 * eps) eps *= 0.1;
    }
    for (int j = 0; j < n; j++) { xkm2[j] = xkm1[j]; xkm1[j] = xk[j]; xk[j] = x[j]; }
    S->newton = newton; S->halvings = halvings; S->eps = eps;
}

}  // namespace pf2

using namespace pf2;

namespace pf2 {

template <int M, bool CONLIN>
static int mma_update_impl(pf2_mma* mm, double* xk, const double* dfdx, const double* dgdx, int* newton_out) {
    pf2_ctx* c = mm->ctx;
    cudaStream_t s = c->stream;
    pf2_dist* d = mm->dist;
    const int n = mm->n, lo = mm->lo, hi = mm->hi;
    const int grid = c->grid_for(hi - lo);
    constexpr int NT1 = M * M + 2 * M + 1;
    mma_setup_kernel<M, CONLIN><<<grid, kThreads, 0, s>>>(lo, hi, n, mm->k, mm->P, xk, mm->xkm1, mm->xkm2, mm->xmin, mm->xmax, dfdx, dgdx, mm->L, mm->U,
                                                 mm->alpha, mm->beta, mm->p0, mm->q0, mm->p, mm->q, mm->x, mm->gsi, mm->ita, mm->S,
                                                 mm->gval, c->red.partials, c->red.ticket);
    if (d) PF2_TRY(dist_allreduce(d, mm->S->red, M));
    mma_setup_small_kernel<M><<<1, 1, 0, s>>>(mm->S, mm->gval);
    PF2_LAUNCH_CHECK();
    c->launches += 2;
    double eps = 1.0;
    int guard = 0;
    while (eps > 1.0e-7) {      // MMA.h:199
        mma_newton1_kernel<M, CONLIN><<<grid, kThreads, 0, s>>>(lo, hi, n, mm->x, mm->L, mm->U, mm->alpha, mm->beta, mm->p0, mm->q0, mm->p, mm->q, mm->gsi,
                                                       mm->ita, mm->Dx, mm->dtx, mm->S, c->red.partials, c->red.ticket);
        if (d) PF2_TRY(dist_allreduce(d, mm->S->red, NT1));
        mma_newton1_small_kernel<M><<<1, 1, 0, s>>>(mm->S);
        mma_newton2_kernel<M, CONLIN><<<grid, kThreads, 0, s>>>(lo, hi, n, mm->x, mm->L, mm->U, mm->alpha, mm->beta, mm->p, mm->q, mm->gsi, mm->ita, mm->Dx,
                                                       mm->dtx, mm->dx, mm->dgsi, mm->dita, mm->S, c->red.partials, c->red.ticket);
        if (d) PF2_TRY(dist_allreduce_max(d, mm->S->red, 1));
        mma_newton2_small_kernel<<<1, 1, 0, s>>>(mm->S);
        c->launches += 4;
        bool accepted = false;
        while (!accepted) {
            mma_trial_kernel<M, CONLIN><<<grid, kThreads, 0, s>>>(lo, hi, n, mm->x, mm->L, mm->U, mm->alpha, mm->beta, mm->p0, mm->q0, mm->p, mm->q, mm->gsi,
                                                         mm->ita, mm->dx, mm->dgsi, mm->dita, mm->xn, mm->gsin, mm->itan, mm->S,
                                                         c->red.partials, c->red.ticket);
            if (d) PF2_TRY(dist_allreduce(d, mm->S->red, M + 1));
            mma_trial_small_kernel<M><<<1, 1, 0, s>>>(mm->S);
            PF2_LAUNCH_CHECK();
            c->launches += 2;
            PF2_CUDA(cudaMemcpyAsync(mm->h_S, mm->S, sizeof(MmaSmall), cudaMemcpyDeviceToHost, s));
            PF2_CUDA(cudaStreamSynchronize(s));
            accepted = mm->h_S->accept != 0;
        }
        std::swap(mm->x, mm->xn); std::swap(mm->gsi, mm->gsin); std::swap(mm->ita, mm->itan);
        eps = mm->h_S->eps;
        if (!(eps == eps) || ++guard > 10000) { set_error("MMA subproblem diverged (eps=%g after %d Newton steps)", eps, guard); return PF2_E_NOCONV; }
    }
    // MMA.h:414-418
    PF2_CUDA(cudaMemcpyAsync(mm->xkm2, mm->xkm1, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    PF2_CUDA(cudaMemcpyAsync(mm->xkm1, xk, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    PF2_CUDA(cudaMemcpyAsync(xk + lo, mm->x + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyDeviceToDevice, s));
    if (newton_out) *newton_out = mm->h_S->newton;
    return PF2_OK;
}

int mma_update(pf2_mma* mm, double* xk, double f, const double* dfdx, const double* g_host, const double* dgdx, int* newton_out) {
    pf2_ctx* c = mm->ctx;
    cudaStream_t s = c->stream;
    PF2_CUDA(cudaSetDevice(c->device));
    PF2_CUDA(cudaMemcpyAsync(mm->gval, g_host, sizeof(double) * mm->m, cudaMemcpyHostToDevice, s));
    int rc;
    if (mm->n <= mm->m) {
        PF2_CHECK(mm->n <= kTinyN, "n <= m is supported for n <= 4 only");
        PF2_CHECK(!mm->conlin, "CONLIN with n <= m is not built");
        mma_tiny_kernel<<<1, 32, 0, s>>>(mm->n, mm->m, mm->k, mm->P, xk, mm->xkm1, mm->xkm2, mm->xmin, mm->xmax, dfdx, dgdx, mm->gval, mm->L, mm->U, mm->S);
        PF2_LAUNCH_CHECK();
        c->launches++;
        PF2_CUDA(cudaMemcpyAsync(mm->h_S, mm->S, sizeof(MmaSmall), cudaMemcpyDeviceToHost, s));
        PF2_CUDA(cudaStreamSynchronize(s));
        if (newton_out) *newton_out = mm->h_S->newton;
        rc = PF2_OK;
    } else {
        if (mm->conlin) {
            switch (mm->m) {
                case 1: rc = mma_update_impl<1, true>(mm, xk, dfdx, dgdx, newton_out); break;
                case 2: rc = mma_update_impl<2, true>(mm, xk, dfdx, dgdx, newton_out); break;
                case 3: rc = mma_update_impl<3, true>(mm, xk, dfdx, dgdx, newton_out); break;
                default: rc = mma_update_impl<4, true>(mm, xk, dfdx, dgdx, newton_out); break;
            }
        } else {
            switch (mm->m) {
                case 1: rc = mma_update_impl<1, false>(mm, xk, dfdx, dgdx, newton_out); break;
                case 2: rc = mma_update_impl<2, false>(mm, xk, dfdx, dgdx, newton_out); break;
                case 3: rc = mma_update_impl<3, false>(mm, xk, dfdx, dgdx, newton_out); break;
                default: rc = mma_update_impl<4, false>(mm, xk, dfdx, dgdx, newton_out); break;
            }
        }
    }
    if (rc != PF2_OK) return rc;
    mm->previousvalue = f;   // MMA.h:414-415
    mm->k++;
    return PF2_OK;
}

}  // namespace pf2

extern "C" {

int pf2_mma_create(pf2_ctx* ctx, int n, int m, double a0, const double* a_host, const double* c_host, const double* d_host,
                   const double* xmin_host, const double* xmax_host, pf2_mma** out) {
    PF2_CHECK(ctx && out && n > 0 && m >= 1, "bad arguments");
    PF2_CHECK(m <= kMaxM, "at most 4 constraints are supported");
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_mma* mm = new pf2_mma();
    mm->ctx = ctx; mm->n = n; mm->m = m;
    mm->lo = 0; mm->hi = n;
    const size_t N = (size_t)n;
    double** vecs[] = { &mm->xmin, &mm->xmax, &mm->xkm1, &mm->xkm2, &mm->L, &mm->U, &mm->alpha, &mm->beta, &mm->p0, &mm->q0, &mm->x,
                        &mm->gsi, &mm->ita, &mm->xn, &mm->gsin, &mm->itan, &mm->dx, &mm->dgsi, &mm->dita, &mm->Dx, &mm->dtx };
    for (double** v : vecs) { PF2_TRY(dev_alloc(v, N)); PF2_CUDA(cudaMemsetAsync(*v, 0, sizeof(double) * N, ctx->stream)); }
    PF2_TRY(dev_alloc(&mm->p, N * m)); PF2_TRY(dev_alloc(&mm->q, N * m));
    PF2_TRY(dev_alloc(&mm->gval, (size_t)kMaxM));
    PF2_TRY(dev_alloc(&mm->S, 1));
    PF2_CUDA(cudaHostAlloc((void**)&mm->h_S, sizeof(MmaSmall), cudaHostAllocDefault));
    memset(mm->h_S, 0, sizeof(MmaSmall));
    mm->h_S->a0 = a0;
    for (int i = 0; i < m; i++) { mm->h_S->a[i] = a_host[i]; mm->h_S->c[i] = c_host[i]; mm->h_S->d[i] = d_host[i]; }
    PF2_CUDA(cudaMemcpyAsync(mm->S, mm->h_S, sizeof(MmaSmall), cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaMemcpyAsync(mm->xmin, xmin_host, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaMemcpyAsync(mm->xmax, xmax_host, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = mm;
    return PF2_OK;
}
/* CONLIN<T> (Optimize/Solver/CONLIN.h:18-26): same handle type and update / convergence entry points as MMA */
int pf2_conlin_create(pf2_ctx* ctx, int n, int m, double a0, const double* a_host, const double* c_host, const double* d_host,
                      const double* xmin_host, const double* xmax_host, pf2_mma** out) {
    PF2_TRY(pf2_mma_create(ctx, n, m, a0, a_host, c_host, d_host, xmin_host, xmax_host, out));
    (*out)->conlin = true;
    (*out)->P.move = 0.5;                 // CONLIN.h:66
    return PF2_OK;
}
int pf2_conlin_set_parameters(pf2_mma* mm, double move, double epsvalue) {     // CONLIN.h:71-74
    PF2_CHECK(mm && mm->conlin, "not a CONLIN optimiser");
    mm->P.move = move;
    mm->epsvalue = epsvalue;
    return PF2_OK;
}

int pf2_mma_destroy(pf2_mma* mm) {
    if (!mm) return PF2_OK;
    cudaStreamSynchronize(mm->ctx->stream);
    double* vecs[] = { mm->xmin, mm->xmax, mm->xkm1, mm->xkm2, mm->L, mm->U, mm->alpha, mm->beta, mm->p0, mm->q0, mm->x, mm->gsi, mm->ita,
                       mm->xn, mm->gsin, mm->itan, mm->dx, mm->dgsi, mm->dita, mm->Dx, mm->dtx, mm->p, mm->q, mm->gval };
    for (double* v : vecs) if (v) cudaFree(v);
    cudaFree(mm->S); cudaFreeHost(mm->h_S);
    delete mm;
    return PF2_OK;
}
int pf2_mma_set_parameters(pf2_mma* mm, double raa0, double albefa, double move, double asyinit, double asydecr, double asyincr, double epsvalue) {
    mm->P = { raa0, albefa, move, asyinit, asydecr, asyincr };
    mm->epsvalue = epsvalue;
    return PF2_OK;
}
int pf2_mma_is_convergence(pf2_mma* mm, double f, int* converged) {
    *converged = fabs(f - mm->previousvalue) / (f + mm->previousvalue) < mm->epsvalue;   // MMA.h:108-113
    return PF2_OK;
}
int pf2_mma_update(pf2_mma* mm, double* x_dev, double f, const double* dfdx_dev, const double* g_host, const double* dgdx_dev, int* newton_steps_out) {
    return mma_update(mm, x_dev, f, dfdx_dev, g_host, dgdx_dev, newton_steps_out);
}

}  // extern "C"
