// filter.cu -- density / Heaviside filters and the OC design update on the device.
//
//   DensityFilter<T>::GetFilteredVariables / GetFilteredSensitivitis      DensityFilter.h:45-71
//   HeavisideFilter<T>::GetFilteredVariables / GetFilteredSensitivitis    HeavisideFilter.h:61-99
//   OC<T>::IsConvergence / UpdateVariables                                OC.h:68-107
//
// The neighbour lists stay in the caller's ragged (CSR) form and are summed in list order, as the reference does.
// One thread per element; all kernels are HBM-bound: 16 B/element + 12 B per neighbour entry.
// The OC bisection runs entirely on the device: per step one candidate kernel and one filter+volume kernel whose last
// CTA takes the bisection decision (lambda0/lambda1, termination) in device memory; the host only polls `done`.
#include "types.cuh"

namespace pf2 {

int dist_halo(pf2_dist* d, double* vec, const int halo[6]);
int dist_allreduce(pf2_dist* d, double* dev, int count);

__device__ __forceinline__ double heaviside(double st, double beta) {
    return 0.5 * (tanh(0.5 * beta) + tanh(beta * (st - 0.5))) / tanh(0.5 * beta);
}

// OC.h:94-98 then the loop condition OC.h:82
__device__ __forceinline__ void oc_decide(OcState* oc, double rho_sum) {
    const double g = oc->volscale * rho_sum - oc->volshift;
    if (g > 0.0) oc->l0 = oc->lambda; else oc->l1 = oc->lambda;
    oc->steps = oc->steps + 1;
    if (!((oc->l1 - oc->l0) / (oc->l1 + oc->l0) > oc->eps)) oc->done = 1;
}
__global__ void oc_decide_kernel(OcState* oc) { if (!oc->done) oc_decide(oc, oc->partial); }

// rho = filter(s); optionally the grid-wide sum of rho (volume constraint) with an OC bisection decision.
template <int KIND, bool SUM>
__global__ void __launch_bounds__(kThreads)
filter_apply_kernel(int n, const long long* __restrict__ rowptr, const int* __restrict__ nbr, const double* __restrict__ w,
                    double beta, const double* __restrict__ s, double* __restrict__ rho, double* sum_out, OcState* oc,
                    double* partials, unsigned int* ticket, int sum_lo, int sum_hi) {
    if (oc != nullptr && oc->done) return;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double wssum = 0.0, wsum = 0.0;
        for (long long j = rowptr[i], je = rowptr[i + 1]; j < je; j++) {
            const double wj = w[j];
            wssum += wj * s[nbr[j]];
            wsum += wj;
        }
        const double st = wssum / wsum;
        const double r = (KIND == PF2_FILTER_HEAVISIDE) ? heaviside(st, beta) : st;
        rho[i] = r;
        if (i >= sum_lo && i < sum_hi) acc += r;
    }
    if (SUM) {
        double v[1] = { acc };
        if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) {
            if (sum_out) *sum_out = v[0];
            if (oc) {
                if (oc->defer) oc->partial = v[0];
                else oc_decide(oc, v[0]);
            }
        }
    }
}

// d rho / d s~ of the Heaviside projection (HeavisideFilter.h:79-87)
__global__ void __launch_bounds__(kThreads)
heaviside_slope_kernel(int n, const long long* __restrict__ rowptr, const int* __restrict__ nbr, const double* __restrict__ w,
                       double beta, const double* __restrict__ s, double* __restrict__ dr) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double wssum = 0.0, wsum = 0.0;
        for (long long j = rowptr[i], je = rowptr[i + 1]; j < je; j++) {
            const double wj = w[j];
            wssum += wj * s[nbr[j]];
            wsum += wj;
        }
        const double th = tanh(beta * (wssum / wsum - 0.5));
        dr[i] = 0.5 * beta * (1.0 - th * th) / tanh(0.5 * beta);
    }
}

// dfds_i = sum_j dfdrho[n_ij] * dr[n_ij] * w_ij / sum_j w_ij   (normalised by the RECEIVING row, HeavisideFilter.h:89-97,
// DensityFilter.h:62-69).  A second field may ride along: either an array g2 or the constant c2 (the volume constraint's
// dg/drho is constant, driver :104).
template <bool CHAIN>
__global__ void __launch_bounds__(kThreads)
filter_sens_kernel(int n, const long long* __restrict__ rowptr, const int* __restrict__ nbr, const double* __restrict__ w,
                   const double* __restrict__ dr, const double* __restrict__ g1, double* __restrict__ out1,
                   const double* __restrict__ g2, double c2, double* __restrict__ out2) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double a1 = 0.0, a2 = 0.0, wsum = 0.0;
        for (long long j = rowptr[i], je = rowptr[i + 1]; j < je; j++) {
            const int nb = nbr[j];
            const double wj = w[j];
            const double d = CHAIN ? dr[nb] : 1.0;
            if (CHAIN) a1 += g1[nb] * d * wj; else a1 += g1[nb] * wj;
            if (out2) { const double v2 = g2 ? g2[nb] : c2; if (CHAIN) a2 += v2 * d * wj; else a2 += v2 * wj; }
            wsum += wj;
        }
        out1[i] = a1 / wsum;
        if (out2) out2[i] = a2 / wsum;
    }
}

// SensitivityFilter (Sigmund, SensitivityFilter.h:44-55):   dfds_i = sum_j w_ij s_j dfds_j / (s_i sum_j w_ij)
// SensitivityFilter2 (Borrvall, SensitivityFilter.h:88-99): dfds_i = sum_j w_ij s_j dfds_j / sum_j w_ij s_j
template <int KIND>
__global__ void __launch_bounds__(kThreads)
sensitivity_filter_kernel(int n, const long long* __restrict__ rowptr, const int* __restrict__ nbr, const double* __restrict__ w,
                          const double* __restrict__ s, const double* __restrict__ g, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double acc = 0.0, wsum = 0.0;
        for (long long j = rowptr[i], je = rowptr[i + 1]; j < je; j++) {
            const int nb = nbr[j];
            const double wj = w[j], sj = s[nb];
            acc += wj * sj * g[nb];
            wsum += (KIND == PF2_FILTER_SENS_SIGMUND) ? wj : wj * sj;
        }
        out[i] = (KIND == PF2_FILTER_SENS_SIGMUND) ? acc / (wsum * s[i]) : acc / wsum;
    }
}

// OC candidate (OC.h:85-92): x+ = clamp((-dfdx/(dgdx*lambda))^iota * x, max(0,(1-move)x), min(1,(1+move)x))
__global__ void __launch_bounds__(kThreads)
oc_candidate_kernel(int n, const double* __restrict__ xk, const double* __restrict__ dfdx, const double* __restrict__ dgdx,
                    double iota, double move, OcState* oc, double* __restrict__ xnew) {
    if (oc->done) return;
    const double lambda = 0.5 * (oc->l1 + oc->l0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = xk[i];
        double v = pow(-dfdx[i] / (dgdx[i] * lambda), iota) * x;
        const double lo = fmax(0.0, (1.0 - move) * x), hi = fmin(1.0, (1.0 + move) * x);
        if (v < lo) v = lo; else if (v > hi) v = hi;
        xnew[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) oc->lambda = lambda;
}

int filter_apply(pf2_filter* f, const double* s, double* rho, double* sum_out, OcState* oc) {
    pf2_ctx* c = f->ctx;
    if (f->kind >= PF2_FILTER_SENS_SIGMUND) { set_error("SensitivityFilter has no GetFilteredVariables (SensitivityFilter.h:18-23)"); return PF2_E_UNSUPPORTED; }
    const int grid = c->grid_for(f->n);
    const bool sum = sum_out != nullptr || oc != nullptr;
#define FA(K, S) filter_apply_kernel<K, S><<<grid, kThreads, 0, c->stream>>>(f->n, f->rowptr, f->nbr, f->w, f->beta, s, rho, sum_out, oc, c->red.partials, c->red.ticket, f->sum_lo, f->sum_hi)
    if (f->kind == PF2_FILTER_HEAVISIDE) { if (sum) FA(PF2_FILTER_HEAVISIDE, true); else FA(PF2_FILTER_HEAVISIDE, false); }
    else { if (sum) FA(PF2_FILTER_DENSITY, true); else FA(PF2_FILTER_DENSITY, false); }
#undef FA
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

// dfds from dfdrho; optionally a second sensitivity (array g2 or constant c2) in the same pass
int filter_sens(pf2_filter* f, const double* s, const double* g1, double* out1, const double* g2, double c2, double* out2) {
    pf2_ctx* c = f->ctx;
    const int grid = c->grid_for(f->n);
    if (f->kind >= PF2_FILTER_SENS_SIGMUND) {
        if (f->kind == PF2_FILTER_SENS_SIGMUND) sensitivity_filter_kernel<PF2_FILTER_SENS_SIGMUND><<<grid, kThreads, 0, c->stream>>>(f->n, f->rowptr, f->nbr, f->w, s, g1, out1);
        else sensitivity_filter_kernel<PF2_FILTER_SENS_BORRVALL><<<grid, kThreads, 0, c->stream>>>(f->n, f->rowptr, f->nbr, f->w, s, g1, out1);
        PF2_LAUNCH_CHECK();
        c->launches++;
        PF2_CHECK(out2 == nullptr, "sensitivity filters take one field at a time");
        return PF2_OK;
    }
    if (f->kind == PF2_FILTER_HEAVISIDE) {
        if (!f->dr) PF2_TRY(dev_alloc(&f->dr, (size_t)f->n));
        heaviside_slope_kernel<<<grid, kThreads, 0, c->stream>>>(f->n, f->rowptr, f->nbr, f->w, f->beta, s, f->dr);
        c->launches++;
        if (f->dist) PF2_TRY(dist_halo(f->dist, f->dr, f->ehalo));      // ghost planes' slopes come from their owners
        filter_sens_kernel<true><<<grid, kThreads, 0, c->stream>>>(f->n, f->rowptr, f->nbr, f->w, f->dr, g1, out1, g2, c2, out2);
    } else {
        filter_sens_kernel<false><<<grid, kThreads, 0, c->stream>>>(f->n, f->rowptr, f->nbr, f->w, nullptr, g1, out1, g2, c2, out2);
    }
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

}  // namespace pf2

using namespace pf2;

namespace pf2 {
int oc_update(pf2_oc* oc, pf2_filter* filter, double weightlimit, double scale1, double* x, double f, const double* dfdx,
              const double* dgdx, int* steps_out, double* lambda_out) {
    pf2_ctx* c = oc->ctx;
    PF2_CHECK(filter->n == oc->n, "filter / optimiser size mismatch");
    const int n = oc->n;
    OcState init;
    init.l0 = oc->lmin; init.l1 = oc->lmax; init.lambda = 0.0; init.eps = oc->leps;
    // g(x) = sum_i scale1*rho_i/(weightlimit*n) - 1.0*scale1   (driver :199-206)
    init.volscale = scale1 / (weightlimit * (double)(oc->n_global ? oc->n_global : n)); init.volshift = 1.0 * scale1;
    init.steps = 0;
    init.partial = 0.0; init.pad = 0;
    init.defer = filter->dist ? 1 : 0;
    init.done = !((init.l1 - init.l0) / (init.l1 + init.l0) > init.eps);
    oc->h_st[0] = init;
    PF2_CUDA(cudaMemcpyAsync(oc->st, &oc->h_st[0], sizeof(OcState), cudaMemcpyHostToDevice, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));   // h_st[0] is reused as a polling slot below
    const int grid = c->grid_for(n);
    const int chunk = 8;
    int slot = 0;
    bool have_prev = false, finished = init.done != 0;
    int enq = 0;
    while (!finished) {
        for (int k = 0; k < chunk; k++) {
            oc_candidate_kernel<<<grid, kThreads, 0, c->stream>>>(n, x, dfdx, dgdx, oc->iota, oc->move, oc->st, oc->xnew);
            c->launches++;
            PF2_TRY(filter_apply(filter, oc->xnew, oc->rho, nullptr, oc->st));
            if (filter->dist) {
                PF2_TRY(dist_allreduce(filter->dist, &oc->st->partial, 1));
                oc_decide_kernel<<<1, 1, 0, c->stream>>>(oc->st);
                c->launches++;
            }
        }
        PF2_LAUNCH_CHECK();
        enq += chunk;
        PF2_CUDA(cudaMemcpyAsync(&oc->h_st[slot], oc->st, sizeof(OcState), cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaEventRecord(oc->ev[slot], c->stream));
        if (have_prev) {
            PF2_CUDA(cudaEventSynchronize(oc->ev[slot ^ 1]));
            if (oc->h_st[slot ^ 1].done) finished = true;
        }
        if (!finished && enq >= 4096) { set_error("OC bisection did not terminate in %d steps", enq); return PF2_E_NOCONV; }
        have_prev = true;
        slot ^= 1;
    }
    // x <- last candidate (OC.h:106); state back for the caller
    if (init.done) {
        // the reference leaves xkp1 = zeros when the loop body never runs (OC.h:81,106)
        PF2_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, c->stream));
    } else {
        PF2_CUDA(cudaMemcpyAsync(x, oc->xnew, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    }
    PF2_CUDA(cudaMemcpyAsync(&oc->h_st[0], oc->st, sizeof(OcState), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    if (steps_out) *steps_out = oc->h_st[0].steps;
    if (lambda_out) *lambda_out = oc->h_st[0].lambda;
    oc->previousvalue = f;   // OC.h:104-105
    oc->k++;
    return PF2_OK;
}
}  // namespace pf2

extern "C" {

int pf2_filter_create(pf2_ctx* ctx, int kind, int n, const long long* rowptr_host, const int* nbr_host, const double* w_host, pf2_filter** out) {
    PF2_CHECK(ctx && out && rowptr_host && nbr_host && w_host && n > 0, "bad arguments");
    PF2_CHECK(kind >= PF2_FILTER_DENSITY && kind <= PF2_FILTER_SENS_BORRVALL, "unknown filter kind");
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_filter* f = new pf2_filter();
    f->ctx = ctx; f->kind = kind; f->n = n; f->nnb = rowptr_host[n];
    f->sum_lo = 0; f->sum_hi = n;
    PF2_TRY(dev_alloc(&f->rowptr, (size_t)n + 1));
    PF2_TRY(dev_alloc(&f->nbr, (size_t)f->nnb));
    PF2_TRY(dev_alloc(&f->w, (size_t)f->nnb));
    PF2_CUDA(cudaMemcpyAsync(f->rowptr, rowptr_host, sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaMemcpyAsync(f->nbr, nbr_host, sizeof(int) * (size_t)f->nnb, cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaMemcpyAsync(f->w, w_host, sizeof(double) * (size_t)f->nnb, cudaMemcpyHostToDevice, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = f;
    return PF2_OK;
}
int pf2_filter_destroy(pf2_filter* f) {
    if (!f) return PF2_OK;
    cudaStreamSynchronize(f->ctx->stream);
    void* ptrs[] = { f->rowptr, f->nbr, f->w, f->dr, f->hs, f->hr, f->hd };
    for (void* p : ptrs) if (p) cudaFree(p);
    delete f;
    return PF2_OK;
}
int pf2_filter_set_beta(pf2_filter* f, double beta) { f->beta = beta; return PF2_OK; }
int pf2_filter_apply(pf2_filter* f, const double* s_dev, double* rho_dev) { return filter_apply(f, s_dev, rho_dev, nullptr, nullptr); }
int pf2_filter_sens(pf2_filter* f, const double* s_dev, const double* dfdrho_dev, double* dfds_dev) {
    return filter_sens(f, s_dev, dfdrho_dev, dfds_dev, nullptr, 0.0, nullptr);
}
static int filter_stage(pf2_filter* f) {
    if (f->hs) return PF2_OK;
    PF2_TRY(dev_alloc(&f->hs, (size_t)f->n)); PF2_TRY(dev_alloc(&f->hr, (size_t)f->n)); PF2_TRY(dev_alloc(&f->hd, (size_t)f->n));
    return PF2_OK;
}
int pf2_filter_apply_host(pf2_filter* f, const double* s_host, double* rho_host) {
    pf2_ctx* c = f->ctx;
    PF2_TRY(filter_stage(f));
    PF2_CUDA(cudaMemcpyAsync(f->hs, s_host, sizeof(double) * (size_t)f->n, cudaMemcpyHostToDevice, c->stream));
    PF2_TRY(filter_apply(f, f->hs, f->hr, nullptr, nullptr));
    PF2_CUDA(cudaMemcpyAsync(rho_host, f->hr, sizeof(double) * (size_t)f->n, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    return PF2_OK;
}
int pf2_filter_sens_host(pf2_filter* f, const double* s_host, const double* dfdrho_host, double* dfds_host) {
    pf2_ctx* c = f->ctx;
    PF2_TRY(filter_stage(f));
    PF2_CUDA(cudaMemcpyAsync(f->hs, s_host, sizeof(double) * (size_t)f->n, cudaMemcpyHostToDevice, c->stream));
    PF2_CUDA(cudaMemcpyAsync(f->hd, dfdrho_host, sizeof(double) * (size_t)f->n, cudaMemcpyHostToDevice, c->stream));
    PF2_TRY(filter_sens(f, f->hs, f->hd, f->hr, nullptr, 0.0, nullptr));
    PF2_CUDA(cudaMemcpyAsync(dfds_host, f->hr, sizeof(double) * (size_t)f->n, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    return PF2_OK;
}

int pf2_oc_create(pf2_ctx* ctx, int n, double iota, double lambdamin, double lambdamax, double lambdaeps, double movelimit, pf2_oc** out) {
    PF2_CHECK(ctx && out && n > 0, "bad arguments");
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_oc* oc = new pf2_oc();
    oc->ctx = ctx; oc->n = n; oc->iota = iota; oc->lmin = lambdamin; oc->lmax = lambdamax; oc->leps = lambdaeps; oc->move = movelimit;
    PF2_TRY(dev_alloc(&oc->xnew, (size_t)n)); PF2_TRY(dev_alloc(&oc->rho, (size_t)n)); PF2_TRY(dev_alloc(&oc->st, 1));
    PF2_CUDA(cudaHostAlloc((void**)&oc->h_st, 2 * sizeof(OcState), cudaHostAllocDefault));
    PF2_CUDA(cudaEventCreateWithFlags(&oc->ev[0], cudaEventDisableTiming));
    PF2_CUDA(cudaEventCreateWithFlags(&oc->ev[1], cudaEventDisableTiming));
    *out = oc;
    return PF2_OK;
}
int pf2_oc_destroy(pf2_oc* oc) {
    if (!oc) return PF2_OK;
    cudaStreamSynchronize(oc->ctx->stream);
    cudaFree(oc->xnew); cudaFree(oc->rho); cudaFree(oc->st); cudaFreeHost(oc->h_st);
    cudaEventDestroy(oc->ev[0]); cudaEventDestroy(oc->ev[1]);
    delete oc;
    return PF2_OK;
}
int pf2_oc_is_convergence(pf2_oc* oc, double f, int* converged) {
    *converged = fabs(f - oc->previousvalue) / (f + oc->previousvalue) < oc->epsvalue;   // OC.h:68-73
    return PF2_OK;
}
int pf2_oc_candidate_host(pf2_oc* oc, const double* x_host, const double* dfdx_host, const double* dgdx_host, double lambda, double* xnew_host) {
    pf2_ctx* c = oc->ctx;
    const size_t n = (size_t)oc->n, nb = sizeof(double) * n;
    double* buf = nullptr;
    PF2_TRY(dev_alloc(&buf, 3 * n));
    PF2_CUDA(cudaMemcpyAsync(buf, x_host, nb, cudaMemcpyHostToDevice, c->stream));
    PF2_CUDA(cudaMemcpyAsync(buf + n, dfdx_host, nb, cudaMemcpyHostToDevice, c->stream));
    PF2_CUDA(cudaMemcpyAsync(buf + 2 * n, dgdx_host, nb, cudaMemcpyHostToDevice, c->stream));
    // a bisection state whose midpoint is exactly `lambda`
    OcState stt;
    memset(&stt, 0, sizeof stt);
    stt.l0 = lambda; stt.l1 = lambda;
    PF2_CUDA(cudaMemcpyAsync(oc->st, &stt, sizeof(OcState), cudaMemcpyHostToDevice, c->stream));
    oc_candidate_kernel<<<c->grid_for(oc->n), kThreads, 0, c->stream>>>(oc->n, buf, buf + n, buf + 2 * n, oc->iota, oc->move, oc->st, oc->xnew);
    PF2_LAUNCH_CHECK();
    c->launches++;
    PF2_CUDA(cudaMemcpyAsync(xnew_host, oc->xnew, nb, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(buf);
    return PF2_OK;
}
int pf2_oc_commit(pf2_oc* oc, double f) { oc->previousvalue = f; oc->k++; return PF2_OK; }

int pf2_oc_update(pf2_oc* oc, pf2_filter* filter, double weightlimit, double scale1, double* x_dev, double f,
                  const double* dfdx_dev, const double* dgdx_dev, int* steps_out, double* lambda_out) {
    return oc_update(oc, filter, weightlimit, scale1, x_dev, f, dfdx_dev, dgdx_dev, steps_out, lambda_out);
}

}  // extern "C"
