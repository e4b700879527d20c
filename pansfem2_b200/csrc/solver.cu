// solver.cu -- CG / ScalingCG / ILU0CG on the device (replaces CG.h:124-154, 420-453, 320-352) and
//              ILU0 / PreILU0 (CG.h:258-315) with level scheduling.
//
// One iteration of the Jacobi-preconditioned solver is three memory-bound kernels:
//   K1  y = A p, fused p.y                                   (spmv_dot, csr.cu)        12*nnz + 24*n bytes
//   K2  alpha = rho/p.y ; x += alpha p ; r -= alpha y ; z = r/D ; z.r, r.r ; last CTA: beta, convergence, counter
//                                                                                       read x,p,r,y,D write x,r,z = 64*n
//   K3  p = beta p + z                                                                  24*n
// alpha, beta, the residual norms, the iteration counter and the `done` flag live in device memory (CgState); the
// host only polls `done` once per chunk of iterations, one chunk behind the GPU, so the queue never drains.
// After convergence every kernel early-exits, so x is exactly the iterate the reference returns (CG.h:443-448).
#include "types.cuh"

namespace pf2 {

int spmv(pf2_csr* A, const double* x, double* y);
int spmv_dot(pf2_csr* A, const double* x, double* y, const CgState* st, double* dot_out);

// x0 = 0 ; r = b - A*x0 = b ; z = M^-1 r ; p = z ; bb = b.b ; rho = z.r ; rr = r.r      (CG.h:422-428)
// MODE 0: no preconditioner, 1: Jacobi (z = r / diag), 2: ILU (z filled in later)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
cg_init_kernel(int n, const double* __restrict__ b, const long long* __restrict__ indptr, const int* __restrict__ diagpos,
               const double* __restrict__ data, double* __restrict__ dvec, double* __restrict__ x, double* __restrict__ r,
               double* __restrict__ z, double* __restrict__ p, CgState* st, int maxit, double eps, double* partials,
               unsigned int* ticket, const double* __restrict__ y0) {
    // y0 = A x0 of a warm start (x keeps x0, r = b - y0); nullptr: the reference's x0 = 0
    double v[3] = { 0.0, 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double bi = b[i];
        double ri = bi;
        if (y0) ri = bi - y0[i]; else x[i] = 0.0;
        r[i] = ri;
        double zi = ri;
        if (MODE == 1) {
            const int dp = diagpos[i];
            const double d = dp >= 0 ? data[indptr[i] + dp] : 0.0;      // GetDiagonal (CG.h:398-404), gathered once per solve
            dvec[i] = d;
            zi = ri / d;
        }
        if (MODE != 2) { z[i] = zi; p[i] = zi; v[1] += zi * ri; }
        v[0] += bi * bi;
        v[2] += ri * ri;
    }
    if (grid_sum_last<3>(v, partials, ticket) && threadIdx.x == 0) {
        st->bb = v[0]; st->rr = v[2]; st->rho = v[1]; st->pAp = 0.0; st->beta = 0.0; st->zr_new = 0.0;
        st->iter = 0; st->maxit = maxit; st->eps = eps;
        st->done = (y0 != nullptr && sqrt(v[2]) < eps * sqrt(v[0])) ? 1 : 0;      // only a warm start can begin converged
    }
}

// the scalar tail of one iteration: beta = rho'/rho ; rho = rho' ; convergence test ||r|| < eps*||b||   (CG.h:437-448)
__device__ __forceinline__ void cg_finalize(CgState* st, double zr, double rr) {
    st->beta = zr / st->rho;
    st->rho = zr;
    st->rr = rr;
    st->iter = st->iter + 1;
    if (sqrt(rr) < st->eps * sqrt(st->bb)) st->done = 1;
}

// K2.  MODE 0/1 finish the iteration here; MODE 2 (ILU) only updates x, r and r.r -- z comes from the triangular solves.
// Two elements per thread with 128-bit loads/stores (all vectors are 16-byte aligned; an odd tail is handled scalar).
template <int MODE>
__global__ void __launch_bounds__(kThreads)
cg_update_kernel(int n, const double* __restrict__ p, const double* __restrict__ y, const double* __restrict__ dvec,
                 double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, CgState* st, double* partials,
                 unsigned int* ticket) {
    if (st->done) return;
    const double alpha = st->rho / st->pAp;
    double v[2] = { 0.0, 0.0 };
    const int n2 = n >> 1;
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* y2 = reinterpret_cast<const double2*>(y);
    const double2* d2 = reinterpret_cast<const double2*>(dvec);
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    double2* z2 = reinterpret_cast<double2*>(z);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const double2 pp = p2[i], yy = y2[i];
        double2 xx = x2[i], rr = r2[i];
        xx.x = xx.x + alpha * pp.x; xx.y = xx.y + alpha * pp.y;
        rr.x = rr.x + (-alpha) * yy.x; rr.y = rr.y + (-alpha) * yy.y;
        x2[i] = xx; r2[i] = rr;
        v[1] += rr.x * rr.x + rr.y * rr.y;
        if (MODE == 0) { v[0] += rr.x * rr.x + rr.y * rr.y; }
        else if (MODE == 1) {
            const double2 dd = d2[i];
            double2 zz;
            zz.x = rr.x / dd.x; zz.y = rr.y / dd.y;
            z2[i] = zz;
            v[0] += zz.x * rr.x + zz.y * rr.y;
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int i = n - 1;
        x[i] = x[i] + alpha * p[i];
        const double ri = r[i] + (-alpha) * y[i];
        r[i] = ri;
        v[1] += ri * ri;
        if (MODE == 0) v[0] += ri * ri;
        else if (MODE == 1) { const double zi = ri / dvec[i]; z[i] = zi; v[0] += zi * ri; }
    }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) {
        if (MODE == 2) st->rr = v[1];
        else cg_finalize(st, v[0], v[1]);
    }
}

// ILU path: rho' = z.r after the triangular solves, then the scalar tail
__global__ void __launch_bounds__(kThreads)
cg_dot_finalize_kernel(int n, const double* __restrict__ z, const double* __restrict__ r, CgState* st, int init,
                       const double* __restrict__ zsrc, double* __restrict__ p, double* partials, unsigned int* ticket) {
    if (!init && st->done) return;
    double v[1] = { 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        v[0] += z[i] * r[i];
        if (init) p[i] = zsrc[i];
    }
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) {
        if (init) st->rho = v[0];
        else cg_finalize(st, v[0], st->rr);
    }
}

// K3: p = beta p + z   (xeaxpy, CG.h:41-49); MODE 0 uses r as z
__global__ void __launch_bounds__(kThreads)
cg_pupdate_kernel(int n, const double* __restrict__ z, double* __restrict__ p, const CgState* __restrict__ st) {
    // the iteration that set `done` still updated p in the reference; x is what matters and it is frozen, so skip
    if (st->done) return;
    const double beta = st->beta;
    const int n2 = n >> 1;
    const double2* z2 = reinterpret_cast<const double2*>(z);
    double2* p2 = reinterpret_cast<double2*>(p);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const double2 zz = z2[i];
        double2 pp = p2[i];
        pp.x = beta * pp.x + zz.x; pp.y = beta * pp.y + zz.y;
        p2[i] = pp;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = beta * p[n - 1] + z[n - 1];
}

static int ensure_workspace(pf2_csr* A) {
    if (A->r) return PF2_OK;
    const size_t n = (size_t)A->rows;
    const size_t np = (n + 31) & ~(size_t)31;     // keep every vector 256-byte aligned inside the slab
    PF2_TRY(dev_alloc(&A->slab, 5 * np));
    A->r = A->slab; A->p = A->slab + np; A->z = A->slab + 2 * np; A->y = A->slab + 3 * np; A->dvec = A->slab + 4 * np;
    PF2_TRY(dev_alloc(&A->st, 1));
    PF2_CUDA(cudaMemset(A->st, 0, sizeof(CgState)));      // red[] / pad are only written by the partitioned path; keep the D2H poll copies clean
    for (int i = 0; i < 2; i++) for (int j = 0; j < 4; j++) PF2_CUDA(cudaEventCreate(&A->pev[i][j]));
    PF2_CUDA(cudaHostAlloc((void**)&A->h_st, 2 * sizeof(CgState), cudaHostAllocDefault));
    PF2_CUDA(cudaEventCreateWithFlags(&A->ev[0], cudaEventDisableTiming));
    PF2_CUDA(cudaEventCreateWithFlags(&A->ev[1], cudaEventDisableTiming));
    return PF2_OK;
}

int ilu0_factor(pf2_csr* A);
int ilu0_apply(pf2_csr* A, double* v, const CgState* st, const double* factors = nullptr);
int ilu0_build_levels(pf2_csr* A);

// enqueue one iteration
// keep the Krylov vectors resident in L2 while the matrix streams through (evict-first loads): the vectors are a third
// of the bytes of a Jacobi-PCG iteration on 2-D problems
static void l2_window(pf2_csr* A, bool on) {
    pf2_ctx* c = A->ctx;
    if (!c->l2_persist_enabled || c->l2_persist_max == 0 || c->l2_window_max == 0 || !A->slab) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (on) {
        const size_t np = (((size_t)A->rows) + 31) & ~(size_t)31;
        size_t bytes = 5 * np * sizeof(double);
        if (bytes > c->l2_window_max) bytes = c->l2_window_max;
        attr.accessPolicyWindow.base_ptr = A->slab;
        attr.accessPolicyWindow.num_bytes = bytes;
        double ratio = (double)c->l2_persist_max / (double)bytes;
        attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
    }
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    if (!on) { if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError(); }
}

// enqueue one iteration; `pev` (4 events) brackets the three kernels when this iteration is a timing sample
static int enqueue_iteration(pf2_csr* A, int solver, double* x, cudaEvent_t* pev) {
    pf2_ctx* c = A->ctx;
    const int n = A->rows;
    if (pev) PF2_CUDA(cudaEventRecord(pev[0], c->stream));
    PF2_TRY(spmv_dot(A, A->p, A->y, A->st, &A->st->pAp));
    if (pev) PF2_CUDA(cudaEventRecord(pev[1], c->stream));
#define UPD(M) cg_update_kernel<M><<<std::min(c->grid_for(n, 4), c->wave_grid((const void*)cg_update_kernel<M>, kThreads)), kThreads, 0, c->stream>>>(n, A->p, A->y, A->dvec, x, A->r, A->z, A->st, c->red.partials, c->red.ticket)
    if (solver == PF2_SOLVER_CG) { UPD(0); }
    else if (solver == PF2_SOLVER_SCALINGCG) { UPD(1); }
    else {
        UPD(2);
        c->launches++;
        PF2_CUDA(cudaMemcpyAsync(A->z, A->r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        PF2_TRY(ilu0_apply(A, A->z, A->st));
        cg_dot_finalize_kernel<<<std::min(c->grid_for(n, 2), c->wave_grid((const void*)cg_dot_finalize_kernel, kThreads)), kThreads, 0, c->stream>>>(n, A->z, A->r, A->st, 0, nullptr, nullptr, c->red.partials, c->red.ticket);
    }
#undef UPD
    c->launches++;
    if (pev) PF2_CUDA(cudaEventRecord(pev[2], c->stream));
    cg_pupdate_kernel<<<std::min(c->grid_for(n, 4), c->wave_grid((const void*)cg_pupdate_kernel, kThreads)), kThreads, 0, c->stream>>>(n, solver == PF2_SOLVER_CG ? A->r : A->z, A->p, A->st);
    c->launches++;
    if (pev) PF2_CUDA(cudaEventRecord(pev[3], c->stream));
    PF2_LAUNCH_CHECK();
    return PF2_OK;
}

static void harvest_profile(pf2_csr* A, int slot) {
    if (!A->pev_armed[slot]) return;
    A->pev_armed[slot] = false;
    float ms[3];
    for (int j = 0; j < 3; j++) if (cudaEventElapsedTime(&ms[j], A->pev[slot][j], A->pev[slot][j + 1]) != cudaSuccess) { cudaGetLastError(); return; }
    for (int j = 0; j < 3; j++) A->prof_ms[j] += ms[j];
    A->prof_samples++;
}

int ensure_workspace_pub(pf2_csr* A) { return ensure_workspace(A); }
int solve_dist(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out);
int solve_bicgstab(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int* iters_out, double* relres_out);

// ---- PCG in nodal numbering with the matrix-free operator (single GPU; csr.cu / spmv_mf.cuh) -----------------------------
int mf_nodal_begin(pf2_csr* A, int jacobi, const double* b, double* bn, double* dn, double* xn, double* r, double* z, double* p0, double* p1, int itrmax, double eps,
                   const double* x0, double* y);
int mf_nodal_apply(pf2_csr* A, const double* p_old, const double* z, double* p_new, double* y);
int mf_nodal_end(pf2_csr* A, const double* xn, double* x);

static int solve_mf_nodal(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    PF2_CUDA(cudaSetDevice(c->device));
    PF2_TRY(ensure_workspace(A));          // device state, events, pinned mirror
    const size_t nfull = (size_t)A->mf_n[0] * A->mf_n[1] * A->mf_n[2] * A->mf_ndof;
    const size_t np = (nfull + 31) & ~(size_t)31;
    if (!A->mf_slab) PF2_TRY(dev_alloc(&A->mf_slab, 8 * np));
    double *bn = A->mf_slab, *dn = bn + np, *xn = bn + 2 * np, *r = bn + 3 * np, *z = bn + 4 * np, *y = bn + 5 * np;
    double* P[2] = { bn + 6 * np, bn + 7 * np };
    const int n = (int)nfull;
    PF2_TRY(mf_nodal_begin(A, solver == PF2_SOLVER_SCALINGCG ? 1 : 0, b, bn, dn, xn, r, z, P[0], P[1], itrmax, eps, warm ? x : nullptr, y));
    const int gu = std::min(c->grid_for(n, 4), c->wave_grid((const void*)cg_update_kernel<1>, kThreads));
    const int chunk = 32;
    int enq = 0, slot = 0;
    bool have_prev = false, finished = false;
    CgState last;
    memset(&last, 0, sizeof last);
    while (!finished) {
        const int todo = std::min(chunk, itrmax - enq);
        for (int k = 0; k < todo; k++) {
            const bool sample = (k == todo / 2) && enq > 0;
            cudaEvent_t* pev = sample ? A->pev[slot] : nullptr;
            const int it = enq + k;
            if (pev) PF2_CUDA(cudaEventRecord(pev[0], c->stream));
            PF2_TRY(mf_nodal_apply(A, P[it & 1], z, P[(it + 1) & 1], y));                       // p = beta p + z ; y = K p ; p.y
            if (pev) PF2_CUDA(cudaEventRecord(pev[1], c->stream));
            cg_update_kernel<1><<<gu, kThreads, 0, c->stream>>>(n, P[(it + 1) & 1], y, dn, xn, r, z, A->st, c->red.partials, c->red.ticket);
            c->launches++;
            if (pev) { PF2_CUDA(cudaEventRecord(pev[2], c->stream)); PF2_CUDA(cudaEventRecord(pev[3], c->stream)); A->pev_armed[slot] = true; }
        }
        PF2_LAUNCH_CHECK();
        enq += todo;
        PF2_CUDA(cudaMemcpyAsync(&A->h_st[slot], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaEventRecord(A->ev[slot], c->stream));
        if (have_prev) {
            PF2_CUDA(cudaEventSynchronize(A->ev[slot ^ 1]));
            last = A->h_st[slot ^ 1];
            if (!last.done) harvest_profile(A, slot ^ 1); else A->pev_armed[slot ^ 1] = false;
            if (last.done) finished = true;
        }
        if (!finished && (enq >= itrmax || todo == 0)) finished = true;
        have_prev = true;
        slot ^= 1;
    }
    PF2_TRY(mf_nodal_end(A, xn, x));
    PF2_CUDA(cudaMemcpyAsync(&A->h_st[0], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    A->pev_armed[0] = A->pev_armed[1] = false;
    last = A->h_st[0];
    A->total_iters += last.iter;
    if (iters_out) *iters_out = last.iter;
    if (relres_out) *relres_out = sqrt(last.rr) / sqrt(last.bb);
    if (!last.done) {
        set_error("Convergence:faild after %d iterations (relres %.3e)", last.iter, sqrt(last.rr) / sqrt(last.bb));
        return PF2_E_NOCONV;
    }
    return PF2_OK;
}

int pcg_persistent_solve(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out);

// warm = 1: x holds the initial guess x0 (the reference always starts from 0, CG.h:423; SURVEY.md section 7 hard part 1 sanctions the
// overload: same recurrences and stopping rule ||r|| < eps ||b||, parity is on the converged solution)
int solve_x0(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    NvtxRange nv("pf2_solve (Krylov loop)");
    PF2_CHECK(solver >= 0 && solver <= PF2_SOLVER_ILU0BICGSTAB, "unknown solver");
    PF2_CHECK(itrmax >= 0, "itrmax");
    if (solver >= PF2_SOLVER_BICGSTAB) return solve_bicgstab(A, solver, b, x, itrmax, eps, iters_out, relres_out);
    PF2_CUDA(cudaSetDevice(c->device));
    if (solver != PF2_SOLVER_ILU0CG && !(A->spmv_variant == 41)) {
        const int rc = pcg_persistent_solve(A, solver, b, x, itrmax, eps, warm, iters_out, relres_out);
        if (rc != PF2_E_UNSUPPORTED) return rc;
    }
    if (A->dist) return solve_dist(A, solver, b, x, itrmax, eps, warm, iters_out, relres_out);
    if (A->spmv_variant == 41 && A->mf_nodal && A->mf_version > 0 && solver != PF2_SOLVER_ILU0CG)
        return solve_mf_nodal(A, solver, b, x, itrmax, eps, warm, iters_out, relres_out);
    PF2_TRY(ensure_workspace(A));
    const int n = A->rows;
    if (solver == PF2_SOLVER_ILU0CG) warm = 0;
    if (warm) PF2_TRY(spmv(A, x, A->y));
    PF2_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 7) == 0, "x must be 16-byte aligned (pf2_malloc gives 256)");
    const int grid = std::min(c->grid_for(n, 2), c->sm_count * 4);
    if (solver == PF2_SOLVER_ILU0CG) PF2_TRY(ilu0_factor(A));
    l2_window(A, true);
#define INIT(M) cg_init_kernel<M><<<grid, kThreads, 0, c->stream>>>(n, b, A->indptr, A->diagpos, A->data, A->dvec, x, A->r, A->z, A->p, A->st, itrmax, eps, c->red.partials, c->red.ticket, warm ? A->y : nullptr)
    if (solver == PF2_SOLVER_CG) { INIT(0); }
    else if (solver == PF2_SOLVER_SCALINGCG) { INIT(1); }
    else {
        INIT(2);
        c->launches++;
        PF2_CUDA(cudaMemcpyAsync(A->z, A->r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        PF2_TRY(ilu0_apply(A, A->z, nullptr));
        cg_dot_finalize_kernel<<<grid, kThreads, 0, c->stream>>>(n, A->z, A->r, A->st, 1, A->z, A->p, c->red.partials, c->red.ticket);
    }
#undef INIT
    c->launches++;
    PF2_LAUNCH_CHECK();

    // chunked enqueue; poll `done` one chunk behind
    const int chunk = (solver == PF2_SOLVER_ILU0CG) ? 4 : 32;
    int enq = 0, slot = 0;
    bool have_prev = false;
    CgState last;
    memset(&last, 0, sizeof last);
    bool finished = false;
    while (!finished) {
        const int todo = std::min(chunk, itrmax - enq);
        for (int k = 0; k < todo; k++) {
            const bool sample = (k == todo / 2) && enq > 0;
            PF2_TRY(enqueue_iteration(A, solver, x, sample ? A->pev[slot] : nullptr));
            if (sample) A->pev_armed[slot] = true;
        }
        enq += todo;
        PF2_CUDA(cudaMemcpyAsync(&A->h_st[slot], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaEventRecord(A->ev[slot], c->stream));
        if (have_prev) {
            PF2_CUDA(cudaEventSynchronize(A->ev[slot ^ 1]));
            last = A->h_st[slot ^ 1];
            if (!last.done) harvest_profile(A, slot ^ 1); else A->pev_armed[slot ^ 1] = false;
            if (last.done) finished = true;
        }
        if (!finished && (enq >= itrmax || todo == 0)) {
            PF2_CUDA(cudaEventSynchronize(A->ev[slot]));
            last = A->h_st[slot];
            finished = true;
        }
        have_prev = true;
        slot ^= 1;
    }
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    A->pev_armed[0] = A->pev_armed[1] = false;
    l2_window(A, false);
    // the freshest state (the chunk in flight may have converged)
    PF2_CUDA(cudaMemcpyAsync(&A->h_st[0], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    last = A->h_st[0];
    A->total_iters += last.iter;
    if (iters_out) *iters_out = last.iter;
    if (relres_out) *relres_out = sqrt(last.rr) / sqrt(last.bb);
    if (!last.done) {
        set_error("Convergence:faild after %d iterations (relres %.3e)", last.iter, sqrt(last.rr) / sqrt(last.bb));
        return PF2_E_NOCONV;
    }
    return PF2_OK;
}

int solve(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int* iters_out, double* relres_out) {
    return solve_x0(A, solver, b, x, itrmax, eps, 0, iters_out, relres_out);
}

}  // namespace pf2

using namespace pf2;

extern "C" {

int pf2_solve(pf2_csr* A, int solver, const double* b_dev, double* x_dev, int itrmax, double eps, int* iters_out, double* relres_out) {
    return solve(A, solver, b_dev, x_dev, itrmax, eps, iters_out, relres_out);
}

int pf2_solve_x0(pf2_csr* A, int solver, const double* b_dev, double* x_dev, int itrmax, double eps, int* iters_out, double* relres_out) {
    return solve_x0(A, solver, b_dev, x_dev, itrmax, eps, 1, iters_out, relres_out);
}

int pf2_csr_set_pcg_mode(pf2_csr* A, int mode) {
    PF2_CHECK(A && mode >= -1 && mode <= 1, "mode: -1 environment default, 0 three kernels per iteration, 1 persistent kernel");
    A->pcg_mode = mode;
    return PF2_OK;
}

int pf2_csr_set_cg_variant(pf2_csr* A, int variant) {
    PF2_CHECK(A && variant >= -1 && variant <= 1, "variant: -1 environment default, 0 the reference's recurrences, 1 single-reduction");
    A->cg_variant = variant;
    return PF2_OK;
}

int pf2_solve_host(pf2_csr* A, int solver, const double* b_host, double* x_host, int itrmax, double eps, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    const size_t n = (size_t)A->rows;
    if (!A->xw) { PF2_TRY(dev_alloc(&A->xw, n)); PF2_TRY(dev_alloc(&A->bw, n)); }
    PF2_CUDA(cudaMemcpyAsync(A->bw, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    int rc = solve(A, solver, A->bw, A->xw, itrmax, eps, iters_out, relres_out);
    if (rc != PF2_OK && rc != PF2_E_NOCONV) return rc;
    PF2_CUDA(cudaMemcpyAsync(x_host, A->xw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}

int pf2_csr_solver_stats(pf2_csr* A, double out[8]) {
    const double k = A->prof_samples ? 1.0 / (double)A->prof_samples : 0.0;
    out[0] = A->prof_ms[0] * k; out[1] = A->prof_ms[1] * k; out[2] = A->prof_ms[2] * k;
    out[3] = (double)A->prof_samples; out[4] = (double)A->total_iters; out[5] = (double)A->spmv_variant;
    out[6] = (double)A->rows; out[7] = (double)A->nnz;
    if (A->prof_samples == 0 && A->pcg_iters > 0) {
        // persistent kernel: per-iteration phase times from its in-kernel %globaltimer stamps (barriers included)
        const double ki = 1.0e-6 / (double)A->pcg_iters;
        for (int j = 0; j < 3; j++) out[j] = A->pcg_phase_ns[j] * ki;
        out[3] = (double)A->pcg_iters;
    }
    return PF2_OK;
}
int pf2_csr_solver_stats_reset(pf2_csr* A) {
    A->prof_ms[0] = A->prof_ms[1] = A->prof_ms[2] = 0.0; A->prof_samples = 0; A->total_iters = 0;
    A->pcg_kernel_ms = 0.0; A->pcg_iters = 0; A->pcg_solves = 0; A->cg1_solves = 0;
    A->pcg_phase_ns[0] = A->pcg_phase_ns[1] = A->pcg_phase_ns[2] = 0.0;
    A->pcg_wait_ns[0] = A->pcg_wait_ns[1] = A->pcg_wait_ns[2] = 0.0;
    return PF2_OK;
}
int pf2_csr_pcg_stats(pf2_csr* A, double out[12]) {
    out[0] = A->pcg_kernel_ms; out[1] = (double)A->pcg_iters; out[2] = (double)A->pcg_solves; out[3] = (double)A->pcg_grid;
    const double ki = A->pcg_iters ? 1.0e-6 / (double)A->pcg_iters : 0.0;
    for (int j = 0; j < 3; j++) { out[4 + j] = A->pcg_phase_ns[j] * ki; out[8 + j] = A->pcg_wait_ns[j] * ki; }
    out[7] = (double)A->sell_entries;
    out[11] = (double)A->cg1_solves;
    return PF2_OK;
}

int pf2_ilu0_factor(pf2_csr* A) { return ilu0_factor(A); }

int pf2_ilu0_download(pf2_csr* A, double* data_host) {
    PF2_CHECK(A->ilu_valid, "call pf2_ilu0_factor first");
    PF2_CUDA(cudaMemcpyAsync(data_host, A->ilu, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost, A->ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(A->ctx->stream));
    return PF2_OK;
}

int pf2_preilu0_host(pf2_csr* M, const double* b_host, double* x_host) {
    pf2_ctx* c = M->ctx;
    PF2_TRY(ilu0_build_levels(M));
    const size_t n = (size_t)M->rows;
    if (!M->xw) { PF2_TRY(dev_alloc(&M->xw, n)); PF2_TRY(dev_alloc(&M->bw, n)); }
    PF2_CUDA(cudaMemcpyAsync(M->xw, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    PF2_TRY(ilu0_apply(M, M->xw, nullptr, M->data));
    PF2_CUDA(cudaMemcpyAsync(x_host, M->xw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    return PF2_OK;
}

int pf2_ilu0_solve_host(pf2_csr* A, const double* b_host, double* x_host) {
    pf2_ctx* c = A->ctx;
    PF2_TRY(ilu0_factor(A));
    const size_t n = (size_t)A->rows;
    if (!A->xw) { PF2_TRY(dev_alloc(&A->xw, n)); PF2_TRY(dev_alloc(&A->bw, n)); }
    PF2_CUDA(cudaMemcpyAsync(A->xw, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    PF2_TRY(ilu0_apply(A, A->xw, nullptr));
    PF2_CUDA(cudaMemcpyAsync(x_host, A->xw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    return PF2_OK;
}

}  // extern "C"
