// dist.cu -- row-block (x-slab) partitioned solves across the GPUs of one box.
//
// The reference has no distributed code; this is the B200-native analogue of its single OpenMP loop.  Each rank owns a
// contiguous range of rows [own_lo, own_hi) of its LOCAL matrix (local mesh = owned element planes + one ghost element
// plane per side, assembled locally: ghost elements are recomputed on both sides, so assembly needs no communication).
// Per PCG iteration:
//     halo exchange of p (one node plane per side, contiguous ranges)        ncclSend/ncclRecv, grouped
//     y = A p on all local rows, p.y over OWNED rows                          spmv_dot (csr.cu) + allreduce of 1 fp64
//     x, r, z update on owned rows, z.r and r.r                               dcg_update_kernel + allreduce of 2 fp64
//     beta, convergence flag                                                 dcg_scalars_kernel (1 thread, device memory)
//     p = beta p + z on owned rows                                            dcg_pupdate_kernel
// All scalars stay in device memory; every rank takes the same decisions because the allreduced values are bitwise
// identical on all ranks, so the NCCL call sequences match without any host-side agreement.
#include <nccl.h>
#include <dlfcn.h>
#include "types.cuh"

// NCCL is bound lazily with dlopen so that (a) single-GPU users never load it and (b) inside a Python process the copy
// torch already loaded (same soname, possibly newer than the system one) is reused instead of a second, older library.
namespace pf2 { namespace nccl {
    static void* lib = nullptr;
    static ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    static ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    static ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    static ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    static ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    static ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    static ncclResult_t (*GroupStart)() = nullptr;
    static ncclResult_t (*GroupEnd)() = nullptr;
    static const char* (*GetErrorString)(ncclResult_t) = nullptr;
    static int load() {
        if (lib) return PF2_OK;
        const char* env = getenv("PF2_NCCL_LIB");
        const char* names[] = { env, "libnccl.so.2", "libnccl.so" };
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // already in the process (e.g. torch's bundled copy)?
        for (int i = 0; !h && i < 3; i++) if (names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!h) { set_error("cannot load NCCL (libnccl.so.2): %s", dlerror()); return PF2_E_UNSUPPORTED; }
#define SYM(name) *(void**)(&name) = dlsym(h, "nccl" #name); if (!name) { set_error("NCCL symbol nccl" #name " missing"); return PF2_E_UNSUPPORTED; }
        SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(Send) SYM(Recv) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
        lib = h;
        return PF2_OK;
    }
} }

struct pf2_dist {
    pf2_ctx* ctx = nullptr;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
};

namespace pf2 {

int spmv_dot(pf2_csr* A, const double* x, double* y, const CgState* st, double* dot_out);
int ensure_workspace_pub(pf2_csr* A);

#define PF2_NCCL(call)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                                        \
            pf2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, nccl::GetErrorString(r__));        \
            return PF2_E_CUDA;                                                                           \
        }                                                                                                \
    } while (0)

// MODE 0: CG, 1: ScalingCG
template <int MODE>
__global__ void __launch_bounds__(kThreads)
dcg_init_kernel(int lo, int hi, int n, const double* __restrict__ b, const long long* __restrict__ indptr, const int* __restrict__ diagpos,
                const double* __restrict__ data, double* __restrict__ dvec, double* __restrict__ x, double* __restrict__ r,
                double* __restrict__ z, double* __restrict__ p, CgState* st, double* partials, unsigned int* ticket) {
    double v[2] = { 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i < lo || i >= hi) { x[i] = 0.0; r[i] = 0.0; z[i] = 0.0; p[i] = 0.0; dvec[i] = 1.0; continue; }
        const double bi = b[i];
        x[i] = 0.0;
        r[i] = bi;
        double zi = bi;
        if (MODE == 1) {
            const int dp = diagpos[i];
            const double d = dp >= 0 ? data[indptr[i] + dp] : 0.0;
            dvec[i] = d;
            zi = bi / d;
        }
        z[i] = zi; p[i] = zi;
        v[0] += bi * bi;
        v[1] += zi * bi;
    }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) { st->red[0] = v[0]; st->red[1] = v[1]; }
}

// phase 0: after the allreduce of {b.b, z.r} ; phase 1: after the allreduce of {z.r, r.r}
__global__ void dcg_scalars_kernel(CgState* st, int phase, int maxit, double eps) {
    if (phase == 0) {
        st->bb = st->red[0]; st->rr = st->red[0]; st->rho = st->red[1]; st->pAp = 0.0; st->beta = 0.0;
        st->iter = 0; st->done = 0; st->maxit = maxit; st->eps = eps;
    } else {
        if (st->done) return;
        const double zr = st->red[0], rr = st->red[1];
        st->beta = zr / st->rho;
        st->rho = zr;
        st->rr = rr;
        st->iter = st->iter + 1;
        if (sqrt(rr) < st->eps * sqrt(st->bb)) st->done = 1;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
dcg_update_kernel(int lo, int hi, const double* __restrict__ p, const double* __restrict__ y, const double* __restrict__ dvec,
                  double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, CgState* st, double* partials,
                  unsigned int* ticket) {
    if (st->done) return;
    const double alpha = st->rho / st->pAp;
    double v[2] = { 0.0, 0.0 };
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        x[i] = x[i] + alpha * p[i];
        const double ri = r[i] + (-alpha) * y[i];
        r[i] = ri;
        v[1] += ri * ri;
        if (MODE == 0) { v[0] += ri * ri; }
        else { const double zi = ri / dvec[i]; z[i] = zi; v[0] += zi * ri; }
    }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) { st->red[0] = v[0]; st->red[1] = v[1]; }
}

__global__ void __launch_bounds__(kThreads)
dcg_pupdate_kernel(int lo, int hi, const double* __restrict__ z, double* __restrict__ p, const CgState* __restrict__ st) {
    if (st->done) return;
    const double beta = st->beta;
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) p[i] = beta * p[i] + z[i];
}

// one node plane per side; ranges are contiguous in the local numbering
int dist_halo(pf2_dist* d, double* vec, const int halo[6]) {
    const int sendL = halo[0], recvL = halo[1], cntL = halo[2], sendR = halo[3], recvR = halo[4], cntR = halo[5];
    cudaStream_t s = d->ctx->stream;
    PF2_NCCL(nccl::GroupStart());
    if (d->rank > 0 && cntL > 0) {
        PF2_NCCL(nccl::Send(vec + sendL, cntL, ncclDouble, d->rank - 1, d->comm, s));
        PF2_NCCL(nccl::Recv(vec + recvL, cntL, ncclDouble, d->rank - 1, d->comm, s));
    }
    if (d->rank < d->nranks - 1 && cntR > 0) {
        PF2_NCCL(nccl::Send(vec + sendR, cntR, ncclDouble, d->rank + 1, d->comm, s));
        PF2_NCCL(nccl::Recv(vec + recvR, cntR, ncclDouble, d->rank + 1, d->comm, s));
    }
    PF2_NCCL(nccl::GroupEnd());
    d->ctx->launches++;
    return PF2_OK;
}

int dist_allreduce(pf2_dist* d, double* dev, int count) {
    PF2_NCCL(nccl::AllReduce(dev, dev, count, ncclDouble, ncclSum, d->comm, d->ctx->stream));
    d->ctx->launches++;
    return PF2_OK;
}

int dist_allreduce_max(pf2_dist* d, double* dev, int count) {
    PF2_NCCL(nccl::AllReduce(dev, dev, count, ncclDouble, ncclMax, d->comm, d->ctx->stream));
    d->ctx->launches++;
    return PF2_OK;
}

int solve_dist(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    pf2_dist* d = A->dist;
    if (solver == PF2_SOLVER_ILU0CG) { set_error("ILU0CG is not available on a partitioned matrix yet"); return PF2_E_UNSUPPORTED; }
    PF2_TRY(ensure_workspace_pub(A));
    const int n = A->rows, lo = A->own_lo, hi = A->own_hi;
    const int grid = std::min(c->grid_for(n, 2), c->sm_count * 4);
    cudaStream_t s = c->stream;
    if (solver == PF2_SOLVER_CG) dcg_init_kernel<0><<<grid, kThreads, 0, s>>>(lo, hi, n, b, A->indptr, A->diagpos, A->data, A->dvec, x, A->r, A->z, A->p, A->st, c->red.partials, c->red.ticket);
    else dcg_init_kernel<1><<<grid, kThreads, 0, s>>>(lo, hi, n, b, A->indptr, A->diagpos, A->data, A->dvec, x, A->r, A->z, A->p, A->st, c->red.partials, c->red.ticket);
    PF2_LAUNCH_CHECK();
    PF2_TRY(dist_allreduce(d, A->st->red, 2));
    dcg_scalars_kernel<<<1, 1, 0, s>>>(A->st, 0, itrmax, eps);
    PF2_TRY(dist_halo(d, A->p, A->halo));
    c->launches += 2;

    const int chunk = 32;
    const int ugrid = std::min(c->grid_for(hi - lo, 2), c->sm_count * 4);
    int enq = 0, slot = 0;
    bool have_prev = false, finished = false;
    CgState last;
    memset(&last, 0, sizeof last);
    while (!finished) {
        const int todo = std::min(chunk, itrmax - enq);
        for (int k = 0; k < todo; k++) {
            const bool sample = (k == todo / 2) && enq > 0;      // one SpMV per chunk is bracketed by events (roofline)
            if (sample) PF2_CUDA(cudaEventRecord(A->pev[slot][0], s));
            PF2_TRY(spmv_dot(A, A->p, A->y, A->st, &A->st->pAp));
            if (sample) { PF2_CUDA(cudaEventRecord(A->pev[slot][1], s)); A->pev_armed[slot] = true; }
            PF2_TRY(dist_allreduce(d, &A->st->pAp, 1));
            if (solver == PF2_SOLVER_CG) dcg_update_kernel<0><<<ugrid, kThreads, 0, s>>>(lo, hi, A->p, A->y, A->dvec, x, A->r, A->z, A->st, c->red.partials, c->red.ticket);
            else dcg_update_kernel<1><<<ugrid, kThreads, 0, s>>>(lo, hi, A->p, A->y, A->dvec, x, A->r, A->z, A->st, c->red.partials, c->red.ticket);
            PF2_TRY(dist_allreduce(d, A->st->red, 2));
            dcg_scalars_kernel<<<1, 1, 0, s>>>(A->st, 1, itrmax, eps);
            dcg_pupdate_kernel<<<ugrid, kThreads, 0, s>>>(lo, hi, solver == PF2_SOLVER_CG ? A->r : A->z, A->p, A->st);
            PF2_TRY(dist_halo(d, A->p, A->halo));
            c->launches += 3;
        }
        PF2_LAUNCH_CHECK();
        enq += todo;
        PF2_CUDA(cudaMemcpyAsync(&A->h_st[slot], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        PF2_CUDA(cudaEventRecord(A->ev[slot], s));
        if (have_prev) {
            PF2_CUDA(cudaEventSynchronize(A->ev[slot ^ 1]));
            last = A->h_st[slot ^ 1];
            if (A->pev_armed[slot ^ 1]) {
                A->pev_armed[slot ^ 1] = false;
                float ms = 0;
                if (!last.done && cudaEventElapsedTime(&ms, A->pev[slot ^ 1][0], A->pev[slot ^ 1][1]) == cudaSuccess) { A->prof_ms[0] += ms; A->prof_samples++; }
                else cudaGetLastError();
            }
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
    PF2_CUDA(cudaStreamSynchronize(s));
    A->pev_armed[0] = A->pev_armed[1] = false;
    PF2_CUDA(cudaMemcpyAsync(&A->h_st[0], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
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

}  // namespace pf2

using namespace pf2;

extern "C" {

int pf2_dist_unique_id(char out[128]) {
    PF2_TRY(nccl::load());
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    PF2_NCCL(nccl::GetUniqueId(&id));
    memcpy(out, &id, 128);
    return PF2_OK;
}

int pf2_dist_create(pf2_ctx* ctx, int rank, int nranks, const char id_bytes[128], pf2_dist** out) {
    PF2_CHECK(ctx && out && id_bytes && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
    PF2_TRY(nccl::load());
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_dist* d = new pf2_dist();
    d->ctx = ctx; d->rank = rank; d->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    PF2_NCCL(nccl::CommInitRank(&d->comm, nranks, id, rank));
    *out = d;
    return PF2_OK;
}

int pf2_dist_destroy(pf2_dist* d) {
    if (!d) return PF2_OK;
    cudaStreamSynchronize(d->ctx->stream);
    if (d->comm) nccl::CommDestroy(d->comm);
    delete d;
    return PF2_OK;
}

int pf2_dist_allreduce_sum(pf2_dist* d, double* dev, int count) { return dist_allreduce(d, dev, count); }

int pf2_dist_halo(pf2_dist* d, double* vec_dev, const int halo[6]) { return dist_halo(d, vec_dev, halo); }

int pf2_csr_set_partition(pf2_csr* A, pf2_dist* d, int own_lo, int own_hi, const int halo[6]) {
    PF2_CHECK(A && own_lo >= 0 && own_lo <= own_hi && own_hi <= A->rows, "bad owned range");
    A->dist = d; A->own_lo = own_lo; A->own_hi = own_hi;
    for (int i = 0; i < 6; i++) A->halo[i] = halo ? halo[i] : 0;
    return PF2_OK;
}

}  // extern "C"
