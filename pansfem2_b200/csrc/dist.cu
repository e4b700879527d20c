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
#include "p2p.cuh"

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

namespace pf2 {
constexpr int kMaxRanks = kMaxRanksT;
// Peer-memory view of the box (handed to the kernels by value): see P2PView in types.cuh.
//   arena (one per rank, cudaMalloc + cudaIpc): slots[2][world][4] fp64 | flags[2][world] u64 | halo_flags[2] u64
typedef P2PView P2P;
}  // namespace pf2

struct pf2_dist {
    pf2_ctx* ctx = nullptr;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    // peer-memory backend (optional)
    bool p2p = false;
    void* arena = nullptr;
    pf2::P2P view;
    unsigned long long* epoch = nullptr;      // device: [0] allreduce epoch, [1] halo epoch
    std::vector<void*> opened;                // peers' arenas, mapped once per partition object
    void* peer_arena[pf2::kMaxRanksT] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
};

namespace pf2 {

int spmv_dot(pf2_csr* A, const double* x, double* y, const CgState* st, double* dot_out);
int ensure_workspace_pub(pf2_csr* A);
int plan_spmv_pub(pf2_csr* A);

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
                double* __restrict__ z, double* __restrict__ p, CgState* st, double* partials, unsigned int* ticket,
                const double* __restrict__ y0) {
    // y0 = A x0 of a warm start (x keeps x0, ghost entries included; r = b - y0); nullptr: the reference's x0 = 0
    double v[3] = { 0.0, 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i < lo || i >= hi) { if (!y0) x[i] = 0.0; r[i] = 0.0; z[i] = 0.0; p[i] = 0.0; dvec[i] = 1.0; continue; }
        const double bi = b[i];
        double ri = bi;
        if (y0) ri = bi - y0[i]; else x[i] = 0.0;
        r[i] = ri;
        double zi = ri;
        if (MODE == 1) {
            const int dp = diagpos[i];
            const double d = dp >= 0 ? data[indptr[i] + dp] : 0.0;
            dvec[i] = d;
            zi = ri / d;
        }
        z[i] = zi; p[i] = zi;
        v[0] += bi * bi;
        v[1] += zi * ri;
        v[2] += ri * ri;
    }
    if (grid_sum_last<3>(v, partials, ticket) && threadIdx.x == 0) { st->red[0] = v[0]; st->red[1] = v[1]; st->red[2] = v[2]; }
}

// phase 0: after the allreduce of {b.b, z.r} ; phase 1: after the allreduce of {z.r, r.r}
__global__ void dcg_scalars_kernel(CgState* st, int phase, int maxit, double eps) {
    if (phase == 0 || phase == 3) {       // 3: warm start -- r.r is its own sum and the solve may begin converged
        st->bb = st->red[0]; st->rr = (phase == 3) ? st->red[2] : st->red[0]; st->rho = st->red[1]; st->pAp = 0.0; st->beta = 0.0;
        st->iter = 0; st->maxit = maxit; st->eps = eps;
        st->done = (phase == 3 && sqrt(st->red[2]) < eps * sqrt(st->red[0])) ? 1 : 0;
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

// K2 with the {z.r, r.r} allreduce and the scalar tail fused into the last CTA (peer-memory backend)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
p2p_update_kernel(int lo, int hi, const double* __restrict__ p, const double* __restrict__ y, const double* __restrict__ dvec,
                  double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, CgState* st, double* partials,
                  unsigned int* ticket, const P2PView* p2p, unsigned long long* epoch_ctr) {
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
    if (grid_sum_last<2>(v, partials, ticket)) {
        __shared__ double sv[4];
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) { sv[0] = v[0]; sv[1] = v[1]; }
            __syncwarp();
            p2p_allreduce_warp(*p2p, epoch_ctr, sv, 2);
            if (threadIdx.x == 0) {
                const double zr = sv[0], rr = sv[1];
                st->beta = zr / st->rho;
                st->rho = zr;
                st->rr = rr;
                st->iter = st->iter + 1;
                if (sqrt(rr) < st->eps * sqrt(st->bb)) st->done = 1;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads)
dcg_pupdate_kernel(int lo, int hi, const double* __restrict__ z, double* __restrict__ p, const CgState* __restrict__ st) {
    if (st->done) return;
    const double beta = st->beta;
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) p[i] = beta * p[i] + z[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Peer-memory collectives (NVLink / NVSwitch, no NCCL on the iteration path).
//   allreduce of <= 4 fp64: one warp; lane r stores this rank's values into rank r's arena, fences, raises its flag there;
//   then every lane waits for rank r's flag in the LOCAL arena and lane 0 sums the slots in rank order, so the result is
//   bitwise identical on every rank.  Slots and flags are double-buffered by epoch parity: a rank cannot run two epochs
//   ahead of a peer because finishing an epoch needs that peer's flag.
// ---------------------------------------------------------------------------------------------------------------
// allreduce of st->pAp after the SpMV (phase 0), of {bb, zr} after init (phase 1; phase 3 = warm start: {bb, zr, rr}), of {zr, rr} + the
// scalar tail (phase 2)
__global__ void p2p_cg_scalars_kernel(P2P P, unsigned long long* epoch_ctr, CgState* st, int phase, int maxit, double eps) {
    if (phase != 1 && phase != 3 && st->done) return;
    __shared__ double v[4];
    if (threadIdx.x == 0) {
        if (phase == 0) { v[0] = st->pAp; }
        else { v[0] = st->red[0]; v[1] = st->red[1]; v[2] = st->red[2]; }
    }
    __syncwarp();
    p2p_allreduce_warp(P, epoch_ctr, v, phase == 0 ? 1 : (phase == 3 ? 3 : 2));
    if (threadIdx.x == 0) {
        if (phase == 0) st->pAp = v[0];
        else if (phase == 1 || phase == 3) {
            st->bb = v[0]; st->rr = (phase == 3) ? v[2] : v[0]; st->rho = v[1]; st->pAp = 0.0; st->beta = 0.0;
            st->iter = 0; st->maxit = maxit; st->eps = eps;
            st->done = (phase == 3 && sqrt(v[2]) < eps * sqrt(v[0])) ? 1 : 0;
        } else {
            const double zr = v[0], rr = v[1];
            st->beta = zr / st->rho;
            st->rho = zr;
            st->rr = rr;
            st->iter = st->iter + 1;
            if (sqrt(rr) < st->eps * sqrt(st->bb)) st->done = 1;
        }
    }
}

// p = beta p + z on the owned rows, with the halo exchange FUSED: boundary planes are stored straight into the neighbours'
// ghost ranges over NVLink; the last CTA publishes the halo epoch to both neighbours and waits for theirs, so that the
// next kernel on this stream (the SpMV) sees complete ghosts.  init = 1: p already holds z (first exchange after set-up).
__global__ void __launch_bounds__(kThreads)
p2p_pupdate_halo_kernel(int lo, int hi, const double* __restrict__ z, double* __restrict__ p, const CgState* __restrict__ st, int init,
                        P2P P, int sendL, int cntL, int sendR, int cntR, unsigned long long* epoch_ctr, unsigned int* ticket) {
    if (!init && st->done) return;
    const double beta = init ? 0.0 : st->beta;
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        const double v = init ? p[i] : beta * p[i] + z[i];
        if (!init) p[i] = v;
        if (P.left_p && i >= sendL && i < sendL + cntL) P.left_p[P.left_recv_off + (i - sendL)] = v;
        if (P.right_p && i >= sendR && i < sendR + cntR) P.right_p[P.right_recv_off + (i - sendR)] = v;
    }
    __shared__ bool last;
    __threadfence_system();                 // this CTA's remote stores are visible system-wide before the ticket
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned long long epoch = epoch_ctr[1] + 1;
        // I am the RIGHT neighbour of rank-1 (slot 1 there) and the LEFT neighbour of rank+1 (slot 0 there)
        if (P.rank > 0) *(volatile unsigned long long*)(P.halo_flags[P.rank - 1] + 1) = epoch;
        if (P.rank < P.world - 1) *(volatile unsigned long long*)(P.halo_flags[P.rank + 1] + 0) = epoch;
        if (!P.defer_halo_wait) {
            // wait here for the neighbours' planes (products that do not order their slices); otherwise the boundary slices of the
            // next product wait, after its interior slices
            const unsigned long long* mine = P.halo_flags[P.rank];
            if (P.rank > 0) p2p_wait_flag(P, mine + 0, epoch);
            if (P.rank < P.world - 1) p2p_wait_flag(P, mine + 1, epoch);
            __threadfence_system();
        }
        epoch_ctr[1] = epoch;
        *ticket = 0u;
    }
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

// ---- ILU0CG on a partitioned matrix: block-Jacobi ILU(0) per rank (ilu.cu restricts factorisation and sweeps to the owned block) -------
int ilu0_factor(pf2_csr* A);
int ilu0_apply(pf2_csr* A, double* v, const CgState* st, const double* factors = nullptr);

// x = 0 ; r = b ; z = r (the sweeps turn it into M^-1 r) ; b.b over the owned rows
__global__ void __launch_bounds__(kThreads)
dcg_ilu_init_kernel(int lo, int hi, int n, const double* __restrict__ b, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                    double* __restrict__ p, CgState* st, double* partials, unsigned int* ticket) {
    double v[1] = { 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const bool own = (i >= lo && i < hi);
        const double bi = own ? b[i] : 0.0;
        x[i] = 0.0; r[i] = bi; z[i] = bi; p[i] = 0.0;
        v[0] += bi * bi;
    }
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) st->red[0] = v[0];
}
// z.r over the owned rows into red[slot]; init: p = z
__global__ void __launch_bounds__(kThreads)
dcg_ilu_dot_kernel(int lo, int hi, const double* __restrict__ z, const double* __restrict__ r, double* __restrict__ p, int init, int slot,
                   CgState* st, double* partials, unsigned int* ticket) {
    if (!init && st->done) return;
    double v[1] = { 0.0 };
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        v[0] += z[i] * r[i];
        if (init) p[i] = z[i];
    }
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) st->red[slot] = v[0];
}
// x += alpha p ; r -= alpha y ; z = r ; r.r over the owned rows into red[1]
__global__ void __launch_bounds__(kThreads)
dcg_ilu_update_kernel(int lo, int hi, const double* __restrict__ p, const double* __restrict__ y, double* __restrict__ x, double* __restrict__ r,
                      double* __restrict__ z, CgState* st, double* partials, unsigned int* ticket) {
    if (st->done) return;
    const double alpha = st->rho / st->pAp;
    double v[1] = { 0.0 };
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        x[i] = x[i] + alpha * p[i];
        const double ri = r[i] + (-alpha) * y[i];
        r[i] = ri; z[i] = ri;
        v[0] += ri * ri;
    }
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) st->red[1] = v[0];
}

static int solve_dist_ilu(pf2_csr* A, const double* b, double* x, int itrmax, double eps, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    pf2_dist* d = A->dist;
    PF2_TRY(ensure_workspace_pub(A));
    PF2_TRY(ilu0_factor(A));
    const int n = A->rows, lo = A->own_lo, hi = A->own_hi;
    cudaStream_t s = c->stream;
    const bool p2p = d->p2p && A->p2p_ready;
    if (p2p && A->p2p_view.defer_halo_wait) {
        A->p2p_view.defer_halo_wait = 0;
        PF2_CUDA(cudaMemcpyAsync(A->p2p_dev, &A->p2p_view, sizeof(P2PView), cudaMemcpyHostToDevice, s));
    }
    const int grid = std::min(c->grid_for(n, 2), c->sm_count * 4);
    const int ugrid = std::min(c->grid_for(hi - lo, 2), c->sm_count * 4);
    // the two scalar reductions of an iteration: peer memory -> one-warp LL allreduce kernel, else NCCL + a one-thread kernel
    auto reduce_scalars = [&](int phase01) -> int {       // phase01 = 0: after set-up {b.b, z.r}; 1: inside an iteration {z.r, r.r}
        if (p2p) p2p_cg_scalars_kernel<<<1, 32, 0, s>>>(A->p2p_view, d->epoch, A->st, phase01 == 0 ? 1 : 2, itrmax, eps);
        else {
            PF2_TRY(dist_allreduce(d, A->st->red, 2));
            dcg_scalars_kernel<<<1, 1, 0, s>>>(A->st, phase01, itrmax, eps);
        }
        c->launches++;
        return PF2_OK;
    };
    auto push_p = [&](int init) -> int {
        if (p2p) p2p_pupdate_halo_kernel<<<ugrid, kThreads, 0, s>>>(lo, hi, A->z, A->p, A->st, init, A->p2p_view, A->halo[0], A->halo[2], A->halo[3], A->halo[5],
                                                                   d->epoch, c->red.ticket + 1);
        else {
            if (!init) dcg_pupdate_kernel<<<ugrid, kThreads, 0, s>>>(lo, hi, A->z, A->p, A->st);
            PF2_TRY(dist_halo(d, A->p, A->halo));
        }
        c->launches++;
        return PF2_OK;
    };
    dcg_ilu_init_kernel<<<grid, kThreads, 0, s>>>(lo, hi, n, b, x, A->r, A->z, A->p, A->st, c->red.partials, c->red.ticket);
    PF2_TRY(ilu0_apply(A, A->z, nullptr));
    dcg_ilu_dot_kernel<<<ugrid, kThreads, 0, s>>>(lo, hi, A->z, A->r, A->p, 1, 1, A->st, c->red.partials, c->red.ticket);
    c->launches += 2;
    PF2_TRY(reduce_scalars(0));
    PF2_TRY(push_p(1));
    PF2_LAUNCH_CHECK();
    const int chunk = 4;
    int enq = 0, slot = 0;
    bool have_prev = false, finished = false;
    CgState last;
    memset(&last, 0, sizeof last);
    while (!finished) {
        const int todo = std::min(chunk, itrmax - enq);
        for (int k = 0; k < todo; k++) {
            PF2_TRY(spmv_dot(A, A->p, A->y, A->st, &A->st->pAp));
            if (!p2p) PF2_TRY(dist_allreduce(d, &A->st->pAp, 1));
            dcg_ilu_update_kernel<<<ugrid, kThreads, 0, s>>>(lo, hi, A->p, A->y, x, A->r, A->z, A->st, c->red.partials, c->red.ticket);
            PF2_TRY(ilu0_apply(A, A->z, A->st));
            dcg_ilu_dot_kernel<<<ugrid, kThreads, 0, s>>>(lo, hi, A->z, A->r, A->p, 0, 0, A->st, c->red.partials, c->red.ticket);
            c->launches += 2;
            PF2_TRY(reduce_scalars(1));
            PF2_TRY(push_p(0));
        }
        PF2_LAUNCH_CHECK();
        enq += todo;
        PF2_CUDA(cudaMemcpyAsync(&A->h_st[slot], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        PF2_CUDA(cudaEventRecord(A->ev[slot], s));
        if (have_prev) {
            PF2_CUDA(cudaEventSynchronize(A->ev[slot ^ 1]));
            last = A->h_st[slot ^ 1];
            if (last.done) finished = true;
        }
        if (!finished && (enq >= itrmax || todo == 0)) finished = true;
        have_prev = true;
        slot ^= 1;
    }
    PF2_CUDA(cudaMemcpyAsync(&A->h_st[0], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    last = A->h_st[0];
    A->total_iters += last.iter;
    if (iters_out) *iters_out = last.iter;
    if (relres_out) *relres_out = sqrt(last.rr) / sqrt(last.bb);
    if (p2p) {
        unsigned long long aborted = 0;
        PF2_CUDA(cudaMemcpy(&aborted, d->epoch + 2, sizeof aborted, cudaMemcpyDeviceToHost));
        if (aborted) { set_error("partitioned PCG: a wait on a peer GPU timed out (a rank left the solve?)"); return PF2_E_CUDA; }
    }
    if (!last.done) {
        set_error("Convergence:faild after %d iterations (relres %.3e)", last.iter, sqrt(last.rr) / sqrt(last.bb));
        return PF2_E_NOCONV;
    }
    return PF2_OK;
}

int spmv(pf2_csr* A, const double* x, double* y);

// ---------------------------------------------------------------------------------------------------------------
// Single-reduction PCG under the partition (opt-in: pf2_csr_set_cg_variant(A, 1) / PF2_CG_SINGLE_REDUCTION=1; peer-memory backend).
// Chronopoulos-Gear form of CG.h:124-154 / 420-453: with u = M^-1 r, w = A u the direction and its image follow recurrences
//     p = u + beta p,  s = w + beta s  (= A p),  x += alpha p,  r -= alpha s,
//     gamma = u.r, delta = w.u,  beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old),
// so gamma, delta and r.r are reduced TOGETHER: one cross-GPU sum and two kernels per iteration (update + halo push | product + sum + tail)
// instead of two sums and three kernels.  Same iterates as the reference in exact arithmetic; round-off differs, so the parity bar is the
// solver tolerance (tests/test_gpu_dist.py), not bit equality, and the reference's recurrences stay the default.
// Storage: A->p holds u (its ghost ranges are what the neighbours write), A->z the direction p, A->y holds w, A->cg1_s holds s.
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads)
cg1_init_kernel(int lo, int hi, int n, const double* __restrict__ b, const long long* __restrict__ indptr, const int* __restrict__ diagpos,
                const double* __restrict__ data, double* __restrict__ dvec, double* __restrict__ x, double* __restrict__ r,
                double* __restrict__ u, double* __restrict__ pd, double* __restrict__ sd, CgState* st, double* partials, unsigned int* ticket,
                const double* __restrict__ y0, int maxit, double eps) {
    double v[3] = { 0.0, 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        pd[i] = 0.0; sd[i] = 0.0;
        // ghost entries of u are NOT touched here: the neighbours' set-up pushes may already have landed (no cross-GPU sum separates this
        // kernel from them, unlike the reference-form set-up), and every ghost row is written by its owner before the first product
        if (i < lo || i >= hi) { if (!y0) x[i] = 0.0; r[i] = 0.0; dvec[i] = 1.0; continue; }
        const double bi = b[i];
        double ri = bi;
        if (y0) ri = bi - y0[i]; else x[i] = 0.0;
        r[i] = ri;
        double ui = ri;
        if (MODE == 1) {
            const int dp = diagpos[i];
            const double d = dp >= 0 ? data[indptr[i] + dp] : 0.0;
            dvec[i] = d;
            ui = ri / d;
        }
        u[i] = ui;
        v[0] += bi * bi;
        v[1] += ui * ri;
        v[2] += ri * ri;
    }
    if (grid_sum_last<3>(v, partials, ticket) && threadIdx.x == 0) {
        st->red[0] = v[0]; st->red[1] = v[1]; st->red[2] = v[2];
        st->rho = 0.0; st->pAp = 0.0; st->beta = 0.0; st->zr_new = 0.0; st->rr = 0.0; st->bb = 0.0;
        st->iter = 0; st->done = 0; st->maxit = maxit; st->eps = eps;
        st->pad = 1;                     // the first product's tail takes {b.b, u.r, r.r} and starts the recurrences (beta = 0)
    }
}

// update of p, s, x, r, u on the owned rows with the halo exchange of u fused (as p2p_pupdate_halo_kernel): partial {u.r, r.r} -> st->red
template <int MODE>
__global__ void __launch_bounds__(kThreads)
cg1_update_halo_kernel(int lo, int hi, const double* __restrict__ w, const double* __restrict__ dvec, double* __restrict__ u,
                       double* __restrict__ pd, double* __restrict__ sd, double* __restrict__ x, double* __restrict__ r, CgState* st,
                       double* partials, unsigned int* ticket, P2P P, int sendL, int cntL, int sendR, int cntR, unsigned long long* epoch_ctr) {
    if (st->done) return;
    const double alpha = st->zr_new, beta = st->beta;
    double v[2] = { 0.0, 0.0 };
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        const double pi = u[i] + beta * pd[i];
        const double si = w[i] + beta * sd[i];
        pd[i] = pi; sd[i] = si;
        x[i] = x[i] + alpha * pi;
        const double ri = r[i] + (-alpha) * si;
        r[i] = ri;
        const double ui = (MODE == 0) ? ri : ri / dvec[i];
        u[i] = ui;
        if (P.left_p && i >= sendL && i < sendL + cntL) P.left_p[P.left_recv_off + (i - sendL)] = ui;
        if (P.right_p && i >= sendR && i < sendR + cntR) P.right_p[P.right_recv_off + (i - sendR)] = ui;
        v[0] += ui * ri;
        v[1] += ri * ri;
    }
    __threadfence_system();              // this CTA's remote stores are visible system-wide before its ticket
    if (!grid_sum_last<2>(v, partials, ticket)) return;
    if (threadIdx.x == 0) {
        st->red[0] = v[0]; st->red[1] = v[1];
        __threadfence_system();
        const unsigned long long epoch = epoch_ctr[1] + 1;
        if (P.rank > 0) *(volatile unsigned long long*)(P.halo_flags[P.rank - 1] + 1) = epoch;
        if (P.rank < P.world - 1) *(volatile unsigned long long*)(P.halo_flags[P.rank + 1] + 0) = epoch;
        const unsigned long long* mine = P.halo_flags[P.rank];
        if (P.rank > 0) p2p_wait_flag(P, mine + 0, epoch);
        if (P.rank < P.world - 1) p2p_wait_flag(P, mine + 1, epoch);
        __threadfence_system();
        epoch_ctr[1] = epoch;
    }
}

static int set_cg1_state(pf2_csr* A, CgState* st, cudaStream_t s) {
    if (A->p2p_view.cg1 == st) return PF2_OK;
    A->p2p_view.cg1 = st;
    PF2_CUDA(cudaMemcpyAsync(A->p2p_dev, &A->p2p_view, sizeof(P2PView), cudaMemcpyHostToDevice, s));
    return PF2_OK;
}

bool cg1_requested(const pf2_csr* A) {
    if (A->cg_variant >= 0) return A->cg_variant == 1;
    static const int env = [] { const char* e = getenv("PF2_CG_SINGLE_REDUCTION"); return (e && atoi(e) != 0) ? 1 : 0; }();
    return env == 1;
}

static int solve_dist_cg1(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    pf2_dist* d = A->dist;
    PF2_TRY(ensure_workspace_pub(A));
    const int n = A->rows, lo = A->own_lo, hi = A->own_hi;
    cudaStream_t s = c->stream;
    if (!A->cg1_s) PF2_CUDA(cudaMalloc((void**)&A->cg1_s, sizeof(double) * (size_t)n));
    plan_spmv_pub(A);
    if (A->p2p_view.cg1 != nullptr || A->p2p_view.defer_halo_wait) {
        A->p2p_view.cg1 = nullptr; A->p2p_view.defer_halo_wait = 0;
        PF2_CUDA(cudaMemcpyAsync(A->p2p_dev, &A->p2p_view, sizeof(P2PView), cudaMemcpyHostToDevice, s));
    }
    const double* y0 = nullptr;
    if (warm) { PF2_TRY(spmv(A, x, A->y)); y0 = A->y; }
    double* u = A->p; double* pd = A->z; double* w = A->y; double* sd = A->cg1_s;
    const int grid = std::min(c->grid_for(n, 2), c->sm_count * 4);
    if (solver == PF2_SOLVER_CG) cg1_init_kernel<0><<<grid, kThreads, 0, s>>>(lo, hi, n, b, A->indptr, A->diagpos, A->data, A->dvec, x, A->r, u, pd, sd, A->st, c->red.partials, c->red.ticket, y0, itrmax, eps);
    else cg1_init_kernel<1><<<grid, kThreads, 0, s>>>(lo, hi, n, b, A->indptr, A->diagpos, A->data, A->dvec, x, A->r, u, pd, sd, A->st, c->red.partials, c->red.ticket, y0, itrmax, eps);
    const int hgrid = std::min(c->grid_for(hi - lo, 2), c->sm_count * 4);
    // ghost ranges of u: the p-update kernel's set-up form pushes the boundary planes of the array it is given and waits for the neighbours'
    p2p_pupdate_halo_kernel<<<hgrid, kThreads, 0, s>>>(lo, hi, u, u, A->st, 1, A->p2p_view, A->halo[0], A->halo[2], A->halo[3], A->halo[5], d->epoch, c->red.ticket + 1);
    PF2_LAUNCH_CHECK();
    PF2_TRY(set_cg1_state(A, A->st, s));
    PF2_TRY(spmv_dot(A, u, w, A->st, &A->st->pAp));          // w0 = A u0; its tail sums {w.u, b.b, u.r, r.r} and sets alpha0, beta0 = 0
    c->launches += 3;

    const int chunk = 32;
    int enq = 0, slot = 0;
    bool have_prev = false, finished = false;
    CgState last;
    memset(&last, 0, sizeof last);
    int rc = PF2_OK;
    while (!finished && rc == PF2_OK) {
        const int todo = std::min(chunk, itrmax - enq);
        for (int k = 0; k < todo && rc == PF2_OK; k++) {
            const bool sample = (k == todo / 2) && enq > 0;      // one iteration per chunk is bracketed by events: update + halo | product + sum
            if (sample) cudaEventRecord(A->pev[slot][0], s);
            if (solver == PF2_SOLVER_CG) cg1_update_halo_kernel<0><<<hgrid, kThreads, 0, s>>>(lo, hi, w, A->dvec, u, pd, sd, x, A->r, A->st, c->red.partials, c->red.ticket, A->p2p_view, A->halo[0], A->halo[2], A->halo[3], A->halo[5], d->epoch);
            else cg1_update_halo_kernel<1><<<hgrid, kThreads, 0, s>>>(lo, hi, w, A->dvec, u, pd, sd, x, A->r, A->st, c->red.partials, c->red.ticket, A->p2p_view, A->halo[0], A->halo[2], A->halo[3], A->halo[5], d->epoch);
            if (sample) cudaEventRecord(A->pev[slot][1], s);
            rc = spmv_dot(A, u, w, A->st, &A->st->pAp);
            if (sample) { cudaEventRecord(A->pev[slot][2], s); A->pev_armed[slot] = true; }
            c->launches += 2;
        }
        if (rc != PF2_OK) break;
        if (cudaGetLastError() != cudaSuccess) { set_error("single-reduction PCG: launch failed"); rc = PF2_E_CUDA; break; }
        enq += todo;
        cudaMemcpyAsync(&A->h_st[slot], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, s);
        cudaEventRecord(A->ev[slot], s);
        if (have_prev) {
            cudaEventSynchronize(A->ev[slot ^ 1]);
            last = A->h_st[slot ^ 1];
            if (A->pev_armed[slot ^ 1]) {
                A->pev_armed[slot ^ 1] = false;
                float mu = 0, mp = 0;
                if (!last.done && cudaEventElapsedTime(&mu, A->pev[slot ^ 1][0], A->pev[slot ^ 1][1]) == cudaSuccess &&
                    cudaEventElapsedTime(&mp, A->pev[slot ^ 1][1], A->pev[slot ^ 1][2]) == cudaSuccess) { A->prof_ms[0] += mp; A->prof_ms[1] += mu; A->prof_samples++; }
                else cudaGetLastError();
            }
            if (last.done) finished = true;
        }
        if (!finished && (enq >= itrmax || todo == 0)) {
            cudaEventSynchronize(A->ev[slot]);
            last = A->h_st[slot];
            finished = true;
        }
        have_prev = true;
        slot ^= 1;
    }
    cudaStreamSynchronize(s);
    A->pev_armed[0] = A->pev_armed[1] = false;
    // products outside this solve (ILU0CG, BiCGSTAB, warm-start residuals) reduce their own dot only
    A->p2p_view.cg1 = nullptr;
    PF2_CUDA(cudaMemcpyAsync(A->p2p_dev, &A->p2p_view, sizeof(P2PView), cudaMemcpyHostToDevice, s));
    if (rc != PF2_OK) return rc;
    PF2_CUDA(cudaMemcpyAsync(&A->h_st[0], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    last = A->h_st[0];
    A->total_iters += last.iter;
    A->cg1_solves++;
    if (iters_out) *iters_out = last.iter;
    if (relres_out) *relres_out = sqrt(last.rr) / sqrt(last.bb);
    unsigned long long aborted = 0;
    PF2_CUDA(cudaMemcpy(&aborted, d->epoch + 2, sizeof aborted, cudaMemcpyDeviceToHost));
    if (aborted) { set_error("partitioned PCG: a wait on a peer GPU timed out (a rank left the solve?)"); return PF2_E_CUDA; }
    if (!last.done) {
        set_error("Convergence:faild after %d iterations (relres %.3e)", last.iter, sqrt(last.rr) / sqrt(last.bb));
        return PF2_E_NOCONV;
    }
    return PF2_OK;
}

// warm = 1: x holds the initial guess, ghost entries included (pf2_solve_x0); ILU0CG always starts from 0
int solve_dist(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    pf2_dist* d = A->dist;
    if (solver == PF2_SOLVER_ILU0CG) return solve_dist_ilu(A, b, x, itrmax, eps, iters_out, relres_out);
    if (d->p2p && A->p2p_ready && cg1_requested(A) && (solver == PF2_SOLVER_CG || solver == PF2_SOLVER_SCALINGCG))
        return solve_dist_cg1(A, solver, b, x, itrmax, eps, warm, iters_out, relres_out);
    PF2_TRY(ensure_workspace_pub(A));
    const int n = A->rows, lo = A->own_lo, hi = A->own_hi;
    const int grid = std::min(c->grid_for(n, 2), c->sm_count * 4);
    cudaStream_t s = c->stream;
    const double* y0 = nullptr;
    if (warm) { PF2_TRY(spmv(A, x, A->y)); y0 = A->y; }        // r0 = b - A x0: one extra product over the local rows
    if (solver == PF2_SOLVER_CG) dcg_init_kernel<0><<<grid, kThreads, 0, s>>>(lo, hi, n, b, A->indptr, A->diagpos, A->data, A->dvec, x, A->r, A->z, A->p, A->st, c->red.partials, c->red.ticket, y0);
    else dcg_init_kernel<1><<<grid, kThreads, 0, s>>>(lo, hi, n, b, A->indptr, A->diagpos, A->data, A->dvec, x, A->r, A->z, A->p, A->st, c->red.partials, c->red.ticket, y0);
    PF2_LAUNCH_CHECK();
    const bool p2p = d->p2p && A->p2p_ready;
    if (p2p) {
        // products that order their slices (SELL-32, natural row order) take over the wait for the neighbours' planes
        plan_spmv_pub(A);
        // opt-in (PF2_HALO_DEFER=1): measured SLOWER on 8 B200s (2-D 4 M dof: 0.0700 vs 0.0685 ms per iteration, hex8 12.8 M dof: 0.272 vs
        // 0.261; profiles/r02_dist_probe.md) -- the system-scope acquire in every boundary warp costs more than the overlap wins
        const int defer = (A->spmv_variant == 31 && A->sell_perm == nullptr && A->sell_d16 != nullptr && (A->sell_nb == 1 || A->sell_nb == 3) &&
                           getenv("PF2_HALO_DEFER") != nullptr && getenv("PF2_HALO_NODEFER") == nullptr) ? 1 : 0;
        if (defer != A->p2p_view.defer_halo_wait) {
            A->p2p_view.defer_halo_wait = defer;
            PF2_CUDA(cudaMemcpyAsync(A->p2p_dev, &A->p2p_view, sizeof(P2PView), cudaMemcpyHostToDevice, s));
        }
    }
    const int hgrid = std::min(c->grid_for(hi - lo, 2), c->sm_count * 4);
    if (p2p) {
        p2p_cg_scalars_kernel<<<1, 32, 0, s>>>(A->p2p_view, d->epoch, A->st, warm ? 3 : 1, itrmax, eps);
        p2p_pupdate_halo_kernel<<<hgrid, kThreads, 0, s>>>(lo, hi, A->z, A->p, A->st, 1, A->p2p_view, A->halo[0], A->halo[2], A->halo[3], A->halo[5],
                                                           d->epoch, c->red.ticket + 1);
    } else {
        PF2_TRY(dist_allreduce(d, A->st->red, warm ? 3 : 2));
        dcg_scalars_kernel<<<1, 1, 0, s>>>(A->st, warm ? 3 : 0, itrmax, eps);
        PF2_TRY(dist_halo(d, A->p, A->halo));
    }
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
            if (sample) PF2_CUDA(cudaEventRecord(A->pev[slot][1], s));
            if (!p2p) PF2_TRY(dist_allreduce(d, &A->st->pAp, 1));      // peer-memory backend: fused into the SpMV's last CTA
            if (p2p) {
                if (solver == PF2_SOLVER_CG) p2p_update_kernel<0><<<ugrid, kThreads, 0, s>>>(lo, hi, A->p, A->y, A->dvec, x, A->r, A->z, A->st, c->red.partials, c->red.ticket, A->p2p_dev, d->epoch);
                else p2p_update_kernel<1><<<ugrid, kThreads, 0, s>>>(lo, hi, A->p, A->y, A->dvec, x, A->r, A->z, A->st, c->red.partials, c->red.ticket, A->p2p_dev, d->epoch);
            } else if (solver == PF2_SOLVER_CG) dcg_update_kernel<0><<<ugrid, kThreads, 0, s>>>(lo, hi, A->p, A->y, A->dvec, x, A->r, A->z, A->st, c->red.partials, c->red.ticket);
            else dcg_update_kernel<1><<<ugrid, kThreads, 0, s>>>(lo, hi, A->p, A->y, A->dvec, x, A->r, A->z, A->st, c->red.partials, c->red.ticket);
            if (sample) PF2_CUDA(cudaEventRecord(A->pev[slot][2], s));      // update (NCCL backend: with the sum of p.Ap before it)
            if (p2p) {
                p2p_pupdate_halo_kernel<<<hgrid, kThreads, 0, s>>>(lo, hi, solver == PF2_SOLVER_CG ? A->r : A->z, A->p, A->st, 0, A->p2p_view,
                                                                   A->halo[0], A->halo[2], A->halo[3], A->halo[5], d->epoch, c->red.ticket + 1);
            } else {
                PF2_TRY(dist_allreduce(d, A->st->red, 2));
                dcg_scalars_kernel<<<1, 1, 0, s>>>(A->st, 1, itrmax, eps);
                dcg_pupdate_kernel<<<ugrid, kThreads, 0, s>>>(lo, hi, solver == PF2_SOLVER_CG ? A->r : A->z, A->p, A->st);
                PF2_TRY(dist_halo(d, A->p, A->halo));
            }
            if (sample) { PF2_CUDA(cudaEventRecord(A->pev[slot][3], s)); A->pev_armed[slot] = true; }      // p-update + halo exchange (+ sums)
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
                float ms[3] = { 0, 0, 0 };
                bool ok = !last.done;
                for (int j = 0; j < 3 && ok; j++) ok = cudaEventElapsedTime(&ms[j], A->pev[slot ^ 1][j], A->pev[slot ^ 1][j + 1]) == cudaSuccess;
                if (ok) { for (int j = 0; j < 3; j++) A->prof_ms[j] += ms[j]; A->prof_samples++; }
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
    if (p2p) {
        unsigned long long aborted = 0;
        PF2_CUDA(cudaMemcpy(&aborted, d->epoch + 2, sizeof aborted, cudaMemcpyDeviceToHost));
        if (aborted) { set_error("partitioned PCG: a wait on a peer GPU timed out (a rank left the solve?)"); return PF2_E_CUDA; }
    }
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
    for (void* p : d->opened) cudaIpcCloseMemHandle(p);
    if (d->arena) cudaFree(d->arena);
    if (d->epoch) cudaFree(d->epoch);
    if (d->comm) nccl::CommDestroy(d->comm);
    delete d;
    return PF2_OK;
}

int pf2_dist_allreduce_sum(pf2_dist* d, double* dev, int count) { return dist_allreduce(d, dev, count); }

// ---- peer-memory backend ------------------------------------------------------------------------------------------
// export: this rank's IPC handles {arena (64 B), Krylov slab of A (64 B)} ; import: all ranks' handles + halo descriptors.
//   arena: slots[2][world][4] fp64 | flags[2][world] u64 | halo_flags[2] u64 | (16-byte aligned) ll[2][world][4][2] u64
static size_t arena_ll_offset(int world) { return ((sizeof(double) * 2 * world * 4 + sizeof(unsigned long long) * (2 * world + 2)) + 15) & ~(size_t)15; }
static size_t arena_bytes(int world) { return arena_ll_offset(world) + sizeof(unsigned long long) * 2 * world * 4 * 2; }

int pf2_csr_p2p_export(pf2_csr* A, char handles_out[128]) {
    PF2_CHECK(A && A->dist, "call pf2_csr_set_partition first");
    pf2_dist* d = A->dist;
    pf2_ctx* c = A->ctx;
    PF2_CUDA(cudaSetDevice(c->device));
    PF2_TRY(ensure_workspace_pub(A));
    if (!d->arena) {
        PF2_CUDA(cudaMalloc(&d->arena, arena_bytes(d->nranks)));
        PF2_CUDA(cudaMemset(d->arena, 0, arena_bytes(d->nranks)));
        PF2_CUDA(cudaMalloc((void**)&d->epoch, 4 * sizeof(unsigned long long)));      // [0] allreduce epoch, [1] halo epoch, [2] abort word
        PF2_CUDA(cudaMemset(d->epoch, 0, 4 * sizeof(unsigned long long)));
    }
    cudaIpcMemHandle_t h0, h1;
    PF2_CUDA(cudaIpcGetMemHandle(&h0, d->arena));
    PF2_CUDA(cudaIpcGetMemHandle(&h1, A->slab));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handles_out, &h0, 64);
    memcpy(handles_out + 64, &h1, 64);
    return PF2_OK;
}

// all_handles: world x 128 bytes (as exported); all_meta: world x 8 ints = each rank's {row halo descriptor[6], local rows, 0}
int pf2_csr_p2p_import(pf2_csr* A, const char* all_handles, const int* all_halo) {
    PF2_CHECK(A && A->dist && A->dist->arena && all_handles && all_halo, "export first");
    pf2_dist* d = A->dist;
    PF2_CHECK(d->nranks <= kMaxRanks, "peer-memory backend supports up to 8 ranks (one box)");
    PF2_CUDA(cudaSetDevice(A->ctx->device));
    const int world = d->nranks, me = d->rank;
    P2P v;
    memset(&v, 0, sizeof v);
    v.rank = me; v.world = world;
    const size_t np = (((size_t)A->rows) + 31) & ~(size_t)31;      // slab layout: r | p | z | y | dvec (solver.cu)
    for (int r = 0; r < world; r++) {
        void* base = nullptr;
        if (r == me) base = d->arena;
        else if (d->peer_arena[r]) base = d->peer_arena[r];       // a second matrix of the same partition: the arenas are already mapped
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + (size_t)r * 128, 64);
            PF2_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            d->opened.push_back(base);
            d->peer_arena[r] = base;
        }
        v.slots[r] = (double*)base;
        v.flags[r] = (unsigned long long*)((char*)base + sizeof(double) * 2 * world * 4);
        v.halo_flags[r] = v.flags[r] + 2 * world;
        v.ll[r] = (unsigned long long*)((char*)base + arena_ll_offset(world));
    }
    (void)np;
    for (int side = 0; side < 2; side++) {
        const int nb = side == 0 ? me - 1 : me + 1;
        if (nb < 0 || nb >= world) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)nb * 128 + 64, 64);
        void* base = nullptr;
        PF2_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        A->p2p_opened[side] = base;
        // the neighbour's p vector starts one padded vector into its slab; its row count is implied by its own padding,
        // so the neighbour publishes the p offset through all_halo? -> no: p offset = np_nb doubles; np_nb is sent as halo[6]
        const int* hn = all_halo + (size_t)nb * 8;
        const long long np_nb = ((long long)hn[6] + 31) & ~31LL;
        double* pvec = (double*)base + np_nb;
        if (side == 0) { v.left_p = pvec; v.left_recv_off = hn[4]; }     // my left plane lands in their RIGHT ghost range (recvR_off)
        else { v.right_p = pvec; v.right_recv_off = hn[1]; }             // my right plane lands in their LEFT ghost range (recvL_off)
    }
    A->pcg_dist_ok = true;
    for (int r = 0; r < world; r++) if (all_halo[(size_t)r * 8 + 7] == 0) A->pcg_dist_ok = false;
    v.abort = (unsigned int*)(d->epoch + 2);
    v.epoch = d->epoch;
    v.own_lo = A->own_lo; v.own_hi = A->own_hi;
    v.sendL = A->halo[0]; v.cntL = me > 0 ? A->halo[2] : 0; v.sendR = A->halo[3]; v.cntR = me < world - 1 ? A->halo[5] : 0;
    v.defer_halo_wait = 0;
    A->p2p_view = v;
    if (!A->p2p_dev) PF2_CUDA(cudaMalloc((void**)&A->p2p_dev, sizeof(P2PView)));
    PF2_CUDA(cudaMemcpy(A->p2p_dev, &v, sizeof(P2PView), cudaMemcpyHostToDevice));
    A->p2p_epoch = d->epoch;
    A->p2p_ready = true;
    d->p2p = true;
    d->view = v;
    return PF2_OK;
}

// Unmap the neighbours' Krylov slabs of this matrix.  Every rank calls it (and the ranks synchronise) BEFORE any of them destroys
// its matrix: freeing memory a peer still has mapped is undefined.
int pf2_csr_p2p_release(pf2_csr* A) {
    if (!A) return PF2_OK;
    PF2_CUDA(cudaSetDevice(A->ctx->device));
    PF2_CUDA(cudaStreamSynchronize(A->ctx->stream));
    for (int side = 0; side < 2; side++) {
        if (A->p2p_opened[side]) { PF2_CUDA(cudaIpcCloseMemHandle(A->p2p_opened[side])); A->p2p_opened[side] = nullptr; }
    }
    A->p2p_ready = false;
    return PF2_OK;
}

int pf2_dist_halo(pf2_dist* d, double* vec_dev, const int halo[6]) { return dist_halo(d, vec_dev, halo); }

int pf2_csr_set_partition(pf2_csr* A, pf2_dist* d, int own_lo, int own_hi, const int halo[6]) {
    PF2_CHECK(A && own_lo >= 0 && own_lo <= own_hi && own_hi <= A->rows, "bad owned range");
    // the ILU(0) level schedule and factors depend on the owned block (block-Jacobi ILU per rank): rebuild them lazily
    if (A->level_rows) {
        cudaFree(A->level_rows); cudaFree(A->level_rows_u); cudaFree(A->level_ptr); cudaFree(A->level_ptr_u);
        A->level_rows = A->level_rows_u = A->level_ptr = A->level_ptr_u = nullptr;
    }
    A->ilu_valid = false;
    A->dist = d; A->own_lo = own_lo; A->own_hi = own_hi;
    for (int i = 0; i < 6; i++) A->halo[i] = halo ? halo[i] : 0;
    return PF2_OK;
}

}  // extern "C"
