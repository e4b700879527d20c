// bicgstab.cu -- the non-symmetric Krylov family on the device:
//     BiCGSTAB (CG.h:159-194), BiCGSTAB2 (CG.h:199-253), ILU0BiCGSTAB (CG.h:357-393), ScalingBiCGSTAB (CG.h:458-495)
// Same recurrences, same operand order inside every vector update (xeaxpbypcz / zeaxpby / zeawpbxmypcz / zeawpbxpcy /
// zeavpbwpcxpdy, CG.h:56-120), x0 = 0, stop on ||r|| < eps*||b|| of the recursive residual.
// As in solver.cu every scalar (alpha, omega, beta, zeta, eta, the dots, the iteration counter, `done`) lives in device
// memory; the host enqueues chunks of iterations and polls `done` one chunk behind; after convergence every kernel
// early-exits so x is the iterate the reference returns.  Reductions are deterministic (last-CTA fold).
#include "types.cuh"

namespace pf2 {

int spmv(pf2_csr* A, const double* x, double* y);
int ilu0_factor(pf2_csr* A);
int ilu0_apply(pf2_csr* A, double* v, const CgState* st, const double* factors = nullptr);

struct BiState {
    double rdashr, alpha, omega, beta, zeta, eta;
    double bb, rr;
    int iter, done, maxit, pad;
    double eps;
};

enum { PRE_NONE = 0, PRE_JACOBI = 1, PRE_ILU = 2 };

// x = 0 ; r = b ; rdash = r ; p = r (Jacobi: p = r / D, CG.h:465) ; rdashr = rdash.r ; bb = b.b
template <int PRE>
__global__ void __launch_bounds__(kThreads)
bi_init_kernel(int n, const double* __restrict__ b, const long long* __restrict__ indptr, const int* __restrict__ diagpos,
               const double* __restrict__ data, double* __restrict__ dvec, double* __restrict__ x, double* __restrict__ r,
               double* __restrict__ rdash, double* __restrict__ p, BiState* st, int maxit, double eps, double* partials, unsigned int* ticket) {
    double v[1] = { 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double bi = b[i];
        x[i] = 0.0; r[i] = bi; rdash[i] = bi;
        double pi = bi;
        if (PRE == PRE_JACOBI) {
            const int dp = diagpos[i];
            const double d = dp >= 0 ? data[indptr[i] + dp] : 0.0;
            dvec[i] = d;
            pi = bi / d;
        }
        p[i] = pi;
        v[0] += bi * bi;
    }
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) {
        st->rdashr = v[0]; st->bb = v[0]; st->rr = v[0];
        st->alpha = st->omega = st->beta = st->zeta = st->eta = 0.0;
        st->iter = 0; st->done = 0; st->maxit = maxit; st->eps = eps;
    }
}

// out = in / D  (Scaling, CG.h:409-415)
__global__ void bi_scale_kernel(int n, const double* __restrict__ in, const double* __restrict__ dvec, double* __restrict__ out, const BiState* st) {
    if (st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] / dvec[i];
}

// alpha = rdashr / (rdash . Ap)
__global__ void __launch_bounds__(kThreads)
bi_alpha_kernel(int n, const double* __restrict__ rdash, const double* __restrict__ Ap, BiState* st, double* partials, unsigned int* ticket) {
    if (st->done) return;
    double v[1] = { 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[0] += rdash[i] * Ap[i];
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) st->alpha = st->rdashr / v[0];
}

// s = 1.0*r + (-alpha)*Ap
__global__ void bi_s_kernel(int n, const double* __restrict__ r, const double* __restrict__ Ap, double* __restrict__ s, const BiState* st) {
    if (st->done) return;
    const double alpha = st->alpha;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s[i] = 1.0 * r[i] + (-alpha) * Ap[i];
}

// omega = (As . s) / (As . As)
__global__ void __launch_bounds__(kThreads)
bi_omega_kernel(int n, const double* __restrict__ As, const double* __restrict__ s, BiState* st, double* partials, unsigned int* ticket) {
    if (st->done) return;
    double v[2] = { 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const double a = As[i]; v[0] += a * s[i]; v[1] += a * a; }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) st->omega = v[0] / v[1];
}

// x = 1.0*x + alpha*Mp + omega*Ms ; r = 1.0*s + (-omega)*As ; rdashr' = rdash.r ; beta ; convergence
__global__ void __launch_bounds__(kThreads)
bi_update_kernel(int n, const double* __restrict__ Mp, const double* __restrict__ Ms, const double* __restrict__ s, const double* __restrict__ As,
                 const double* __restrict__ rdash, double* __restrict__ x, double* __restrict__ r, BiState* st, double* partials, unsigned int* ticket) {
    if (st->done) return;
    const double alpha = st->alpha, omega = st->omega;
    double v[2] = { 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] = 1.0 * x[i] + alpha * Mp[i] + omega * Ms[i];
        const double ri = 1.0 * s[i] + (-omega) * As[i];
        r[i] = ri;
        v[0] += rdash[i] * ri;
        v[1] += ri * ri;
    }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) {
        st->beta = alpha / omega * v[0] / st->rdashr;
        st->rdashr = v[0];
        st->rr = v[1];
        st->iter = st->iter + 1;
        if (sqrt(v[1]) < st->eps * sqrt(st->bb)) st->done = 1;
    }
}

// p = beta*p + 1.0*r + (-beta*omega)*Ap     (the reference also does this on the converging iteration; x is frozen, so skip)
__global__ void bi_p_kernel(int n, const double* __restrict__ r, const double* __restrict__ Ap, double* __restrict__ p, const BiState* st) {
    if (st->done) return;
    const double beta = st->beta, c = -st->beta * st->omega;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = beta * p[i] + 1.0 * r[i] + c * Ap[i];
}

// ---- BiCGSTAB2 ---------------------------------------------------------------------------------------------------
// p = beta*p + 1.0*r + (-beta)*u
__global__ void bi2_p_kernel(int n, const double* __restrict__ r, const double* __restrict__ u, double* __restrict__ p, const BiState* st) {
    if (st->done) return;
    const double beta = st->beta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = beta * p[i] + 1.0 * r[i] + (-beta) * u[i];
}
// y = 1.0*tkm1 + (-1.0)*r + (-alpha)*w + alpha*Ap ; t = 1.0*r + (-alpha)*Ap
__global__ void bi2_yt_kernel(int n, const double* __restrict__ tkm1, const double* __restrict__ r, const double* __restrict__ w,
                              const double* __restrict__ Ap, double* __restrict__ y, double* __restrict__ t, const BiState* st) {
    if (st->done) return;
    const double alpha = st->alpha;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        y[i] = 1.0 * tkm1[i] + (-1.0) * r[i] + (-alpha) * w[i] + alpha * Ap[i];
        t[i] = 1.0 * r[i] + (-alpha) * Ap[i];
    }
}
// zeta, eta (CG.h:224-234): even iterations use the one-dimensional minimiser, odd ones the two-dimensional one
__global__ void __launch_bounds__(kThreads)
bi2_zeta_kernel(int n, const double* __restrict__ At, const double* __restrict__ t, const double* __restrict__ y, BiState* st, double* partials,
                unsigned int* ticket) {
    if (st->done) return;
    double v[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double a = At[i], ti = t[i], yi = y[i];
        v[0] += a * ti; v[1] += a * a; v[2] += yi * yi; v[3] += yi * ti; v[4] += a * yi;
    }
    if (grid_sum_last<5>(v, partials, ticket) && threadIdx.x == 0) {
        const double Att = v[0], AtAt = v[1], yy = v[2], yt = v[3], Aty = v[4];
        if ((st->iter & 1) == 0) { st->zeta = Att / AtAt; st->eta = 0.0; }
        else {
            st->zeta = (yy * Att - yt * Aty) / (AtAt * yy - Aty * Aty);
            st->eta = (AtAt * yt - Aty * Att) / (AtAt * yy - Aty * Aty);
        }
    }
}
// u = zeta*Ap + eta*(tkm1 - r + beta*u) ; z = eta*z + zeta*r + (-alpha)*u ; x = 1.0*x + alpha*p + 1.0*z ;
// r = 1.0*t + (-eta)*y + (-zeta)*At ; rdashr' ; beta ; convergence            (CG.h:235-243; u, z, x use the OLD r)
__global__ void __launch_bounds__(kThreads)
bi2_update_kernel(int n, const double* __restrict__ Ap, const double* __restrict__ tkm1, const double* __restrict__ p, const double* __restrict__ t,
                  const double* __restrict__ y, const double* __restrict__ At, const double* __restrict__ rdash, double* __restrict__ u,
                  double* __restrict__ z, double* __restrict__ x, double* __restrict__ r, BiState* st, double* partials, unsigned int* ticket) {
    if (st->done) return;
    const double alpha = st->alpha, beta = st->beta, zeta = st->zeta, eta = st->eta;
    double v[2] = { 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double ro = r[i];
        const double ui = zeta * Ap[i] + eta * (tkm1[i] - ro + beta * u[i]);
        u[i] = ui;
        const double zi = eta * z[i] + zeta * ro + (-alpha) * ui;
        z[i] = zi;
        x[i] = 1.0 * x[i] + alpha * p[i] + 1.0 * zi;
        const double rn = 1.0 * t[i] + (-eta) * y[i] + (-zeta) * At[i];
        r[i] = rn;
        v[0] += rdash[i] * rn;
        v[1] += rn * rn;
    }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) {
        st->beta = alpha * v[0] / (zeta * st->rdashr);
        st->rdashr = v[0];
        st->rr = v[1];
        st->iter = st->iter + 1;
        if (sqrt(v[1]) < st->eps * sqrt(st->bb)) st->done = 1;
    }
}
// w = 1.0*At + beta*Ap ; tkm1 = t
__global__ void bi2_w_kernel(int n, const double* __restrict__ At, const double* __restrict__ Ap, const double* __restrict__ t, double* __restrict__ w,
                             double* __restrict__ tkm1, const BiState* st) {
    if (st->done) return;
    const double beta = st->beta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { w[i] = 1.0 * At[i] + beta * Ap[i]; tkm1[i] = t[i]; }
}

static int ensure_bi_workspace(pf2_csr* A) {
    if (A->bi_slab) return PF2_OK;
    const size_t np = (((size_t)A->rows) + 31) & ~(size_t)31;
    PF2_TRY(dev_alloc(&A->bi_slab, 12 * np));
    PF2_TRY(dev_alloc((BiState**)&A->bi_st, 1));
    PF2_CUDA(cudaMemset(A->bi_st, 0, sizeof(BiState)));
    PF2_CUDA(cudaHostAlloc((void**)&A->bi_hst, 2 * sizeof(BiState), cudaHostAllocDefault));
    if (!A->bi_ev[0]) {
        PF2_CUDA(cudaEventCreateWithFlags(&A->bi_ev[0], cudaEventDisableTiming));
        PF2_CUDA(cudaEventCreateWithFlags(&A->bi_ev[1], cudaEventDisableTiming));
    }
    return PF2_OK;
}

int solve_bicgstab(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    PF2_CHECK(!A->dist, "the BiCGSTAB family is not partitioned across GPUs");
    PF2_CHECK(itrmax >= 0, "itrmax");
    PF2_CUDA(cudaSetDevice(c->device));
    PF2_TRY(ensure_bi_workspace(A));
    const int n = A->rows;
    const size_t np = (((size_t)n) + 31) & ~(size_t)31;
    double* W = A->bi_slab;
    double *r = W, *rdash = W + np, *p = W + 2 * np, *Ap = W + 3 * np, *s = W + 4 * np, *As = W + 5 * np, *Mp = W + 6 * np, *Ms = W + 7 * np;
    double *dvec = W + 8 * np, *u = W + 9 * np, *w = W + 10 * np, *z = W + 11 * np;       // BiCGSTAB2: y = Mp, tkm1 = Ms
    BiState* st = (BiState*)A->bi_st;
    BiState* hst = (BiState*)A->bi_hst;
    cudaStream_t sm = c->stream;
    const int g = std::min(c->grid_for(n, 2), c->sm_count * 8);
    const int pre = solver == PF2_SOLVER_SCALINGBICGSTAB ? PRE_JACOBI : (solver == PF2_SOLVER_ILU0BICGSTAB ? PRE_ILU : PRE_NONE);
    const bool two = solver == PF2_SOLVER_BICGSTAB2;
    if (pre == PRE_ILU) PF2_TRY(ilu0_factor(A));
    if (pre == PRE_JACOBI) bi_init_kernel<PRE_JACOBI><<<g, kThreads, 0, sm>>>(n, b, A->indptr, A->diagpos, A->data, dvec, x, r, rdash, p, st, itrmax, eps, c->red.partials, c->red.ticket);
    else bi_init_kernel<PRE_NONE><<<g, kThreads, 0, sm>>>(n, b, A->indptr, A->diagpos, A->data, dvec, x, r, rdash, p, st, itrmax, eps, c->red.partials, c->red.ticket);
    c->launches++;
    if (two) {
        // p, u, tkm1, w, z start at zero (CG.h:205-209); the first p-update (beta = 0) then gives p = r
        PF2_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * np, sm));
        PF2_CUDA(cudaMemsetAsync(Ms, 0, sizeof(double) * np, sm));
        PF2_CUDA(cudaMemsetAsync(u, 0, sizeof(double) * 3 * np, sm));
    }
    PF2_LAUNCH_CHECK();

    auto iteration = [&]() -> int {
        if (!two) {
            const double* mp = p;
            if (pre == PRE_JACOBI) { bi_scale_kernel<<<g, kThreads, 0, sm>>>(n, p, dvec, Mp, st); c->launches++; mp = Mp; }
            else if (pre == PRE_ILU) { PF2_CUDA(cudaMemcpyAsync(Mp, p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, sm)); PF2_TRY(ilu0_apply(A, Mp, nullptr)); mp = Mp; }
            PF2_TRY(spmv(A, mp, Ap));
            bi_alpha_kernel<<<g, kThreads, 0, sm>>>(n, rdash, Ap, st, c->red.partials, c->red.ticket);
            bi_s_kernel<<<g, kThreads, 0, sm>>>(n, r, Ap, s, st);
            const double* ms = s;
            if (pre == PRE_JACOBI) { bi_scale_kernel<<<g, kThreads, 0, sm>>>(n, s, dvec, Ms, st); c->launches++; ms = Ms; }
            else if (pre == PRE_ILU) { PF2_CUDA(cudaMemcpyAsync(Ms, s, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, sm)); PF2_TRY(ilu0_apply(A, Ms, nullptr)); ms = Ms; }
            PF2_TRY(spmv(A, ms, As));
            bi_omega_kernel<<<g, kThreads, 0, sm>>>(n, As, s, st, c->red.partials, c->red.ticket);
            bi_update_kernel<<<g, kThreads, 0, sm>>>(n, mp, ms, s, As, rdash, x, r, st, c->red.partials, c->red.ticket);
            bi_p_kernel<<<g, kThreads, 0, sm>>>(n, r, Ap, p, st);
            c->launches += 5;
        } else {
            double *y = Mp, *tkm1 = Ms, *t = s, *At = As;
            bi2_p_kernel<<<g, kThreads, 0, sm>>>(n, r, u, p, st);
            PF2_TRY(spmv(A, p, Ap));
            bi_alpha_kernel<<<g, kThreads, 0, sm>>>(n, rdash, Ap, st, c->red.partials, c->red.ticket);
            bi2_yt_kernel<<<g, kThreads, 0, sm>>>(n, tkm1, r, w, Ap, y, t, st);
            PF2_TRY(spmv(A, t, At));
            bi2_zeta_kernel<<<g, kThreads, 0, sm>>>(n, At, t, y, st, c->red.partials, c->red.ticket);
            bi2_update_kernel<<<g, kThreads, 0, sm>>>(n, Ap, tkm1, p, t, y, At, rdash, u, z, x, r, st, c->red.partials, c->red.ticket);
            bi2_w_kernel<<<g, kThreads, 0, sm>>>(n, At, Ap, t, w, tkm1, st);
            c->launches += 6;
        }
        PF2_LAUNCH_CHECK();
        return PF2_OK;
    };

    const int chunk = (pre == PRE_ILU) ? 2 : 16;
    int enq = 0, slot = 0;
    bool have_prev = false, finished = false;
    while (!finished) {
        const int todo = std::min(chunk, itrmax - enq);
        for (int k = 0; k < todo; k++) PF2_TRY(iteration());
        enq += todo;
        PF2_CUDA(cudaMemcpyAsync(&hst[slot], st, sizeof(BiState), cudaMemcpyDeviceToHost, sm));
        PF2_CUDA(cudaEventRecord(A->bi_ev[slot], sm));
        if (have_prev) {
            PF2_CUDA(cudaEventSynchronize(A->bi_ev[slot ^ 1]));
            if (hst[slot ^ 1].done) finished = true;
        }
        if (!finished && (enq >= itrmax || todo == 0)) finished = true;
        have_prev = true;
        slot ^= 1;
    }
    PF2_CUDA(cudaMemcpyAsync(&hst[0], st, sizeof(BiState), cudaMemcpyDeviceToHost, sm));
    PF2_CUDA(cudaStreamSynchronize(sm));
    const BiState last = hst[0];
    A->total_iters += last.iter;
    if (iters_out) *iters_out = last.iter;
    if (relres_out) *relres_out = sqrt(last.rr) / sqrt(last.bb);
    if (!last.done) {
        set_error("Convergence:faild after %d iterations (relres %.3e)", last.iter, sqrt(last.rr) / sqrt(last.bb));
        return PF2_E_NOCONV;
    }
    return PF2_OK;
}

}  // namespace pf2
