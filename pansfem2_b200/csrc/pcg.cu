// pcg.cu -- host side of the persistent PCG kernel (pcg_persistent.cuh): picks the instantiation for the matrix' SELL-32 mirror
//           (16-bit deltas / block deltas / absolute columns, streaming or L2-resident loads, partitioned or not), launches it
//           cooperatively once per solve and reads the result back.
#include "types.cuh"
#include "pcg_persistent.cuh"

namespace pf2 {

int plan_spmv_pub(pf2_csr* A);
int sell_refresh(pf2_csr* A);

// ---- persistent PCG kernel (pcg_persistent.cuh) ---------------------------------------------------------------------------
// PF2_E_UNSUPPORTED without an error message = "not applicable to this matrix": the caller falls back to the three-kernel loop.
int ensure_workspace_pub(pf2_csr* A);

// Which loop runs (pf2_csr_set_pcg_mode / PF2_PCG override).  Measured on B200 (profiles/r02_pcg_tune.md): a grid-wide exchange of the
// persistent kernel costs 3 us on a 20-CTA grid and 6-10 us on a full one, a kernel boundary 2-3 us whatever the grid, and the
// stand-alone kernels keep a few per cent more of the HBM rate; so the persistent kernel is the default only where launches dominate
// (small systems: 60x40 sample 1.5x faster) and stays selectable everywhere else, partitioned matrices included.
static bool pcg_enabled(const pf2_csr* A) {
    if (A->pcg_mode >= 0) return A->pcg_mode != 0;
    static const int env = getenv("PF2_PCG") ? atoi(getenv("PF2_PCG")) : -1;
    if (env >= 0) return env != 0;
    static const long long auto_rows = getenv("PF2_PCG_AUTO_ROWS") ? atoll(getenv("PF2_PCG_AUTO_ROWS")) : 50000;
    return A->dist == nullptr && (long long)A->rows <= auto_rows;
}

template <class IDX, int NB, int MODE, bool DIST, bool CS>
static int pcg_launch(pf2_csr* A, PcgArgs& args) {
    pf2_ctx* c = A->ctx;
    const void* fn = (const void*)pcg_persistent_kernel<IDX, NB, MODE, DIST, CS>;
    const int wave = c->wave_grid(fn, kThreads);
    const int s_lo = args.own_lo / kSellC, s_hi = (args.own_hi + kSellC - 1) / kSellC;
    const int want = std::max(1, (s_hi - s_lo + (kThreads / 32) - 1) / (kThreads / 32));
    int grid = std::min(std::min(wave, want), kPcgMaxCtas);
    static const int cap = getenv("PF2_PCG_GRID") ? atoi(getenv("PF2_PCG_GRID")) : 0;      // tuning / tests: CTAs of the cooperative grid
    if (cap > 0) grid = std::min(grid, cap);
    A->pcg_grid = grid;
    void* params[] = { (void*)&args };
    PF2_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), params, 0, c->stream));
    c->launches++;
    return PF2_OK;
}

int pcg_persistent_solve(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out) {
    pf2_ctx* c = A->ctx;
    if (!pcg_enabled(A) || (solver != PF2_SOLVER_CG && solver != PF2_SOLVER_SCALINGCG)) return PF2_E_UNSUPPORTED;
    const bool dist = A->dist != nullptr;
    // partitioned: NCCL backend keeps the host-ordered loop; with peer memory EVERY rank must take the same path (the protocols differ),
    // so the ranks agreed at import time on whether all of their slabs qualify (pf2_csr_pcg_capable -> meta[7])
    if (dist && (!A->p2p_ready || !A->pcg_dist_ok)) return PF2_E_UNSUPPORTED;
    PF2_TRY(plan_spmv_pub(A));
    if (A->spmv_variant != 31 || A->sell_nb == 2 || (A->sell_perm && dist)) return PF2_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return PF2_E_UNSUPPORTED;      // the vector phases use 128-bit accesses
    PF2_CUDA(cudaSetDevice(c->device));
    PF2_TRY(ensure_workspace_pub(A));
    if (!A->sell_values_valid) PF2_TRY(sell_refresh(A));
    if (!A->pcg_sync) {
        PF2_CUDA(cudaMalloc(&A->pcg_sync, sizeof(PcgSync)));
        PF2_CUDA(cudaHostAlloc(&A->h_pcg_sync, sizeof(unsigned long long) * 8, cudaHostAllocDefault));
    }
    PF2_CHECK((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(b) & 7) == 0, "x and b must be 8-byte aligned");
    PcgArgs a;
    memset(&a, 0, sizeof a);
    a.rows = A->rows; a.nslices = (A->rows + kSellC - 1) / kSellC;
    a.own_lo = dist ? A->own_lo : 0; a.own_hi = dist ? A->own_hi : A->rows;
    a.itrmax = itrmax; a.warm = warm; a.eps = eps;
    a.slice_ptr = A->sell_ptr; a.sell_val = A->sell_val; a.perm = A->sell_perm;
    a.sell_idx = A->sell_b32 ? (const void*)A->sell_b32 : A->sell_d16 ? (const void*)A->sell_d16 : (const void*)A->sell_idx;
    a.indptr = A->indptr; a.diagpos = A->diagpos; a.data = A->data;
    a.b = b; a.x = x; a.r = A->r; a.z = A->z; a.p = A->p; a.y = A->y; a.dvec = A->dvec;
    a.st = A->st; a.sync = (PcgSync*)A->pcg_sync;
    a.p2p = A->p2p_dev; a.epoch = A->p2p_epoch;
    a.sendL = A->halo[0]; a.cntL = A->halo[2]; a.sendR = A->halo[3]; a.cntR = A->halo[5];
    if (dist) {
        if (A->p2p_view.rank == 0) a.cntL = 0;
        if (A->p2p_view.rank == A->p2p_view.world - 1) a.cntR = 0;
    }
    // matrix stream: evict-first when it cannot stay in L2 anyway, plain loads when the slab's mirror is small enough to live there
    static const double cs_mb = getenv("PF2_PCG_CS_MB") ? atof(getenv("PF2_PCG_CS_MB")) : 48.0;
    const double idx_bytes = A->sell_b32 ? 4.0 / A->sell_nb : A->sell_d16 ? 2.0 / A->sell_nb : 4.0;
    const bool cs = (double)A->sell_entries * (8.0 + idx_bytes) > cs_mb * 1.0e6;
    const bool wide = A->sell_b32 != nullptr || (A->sell_d16 == nullptr);      // 4-byte index stream
    const int mode = solver == PF2_SOLVER_SCALINGCG ? 1 : 0;
    PF2_CUDA(cudaMemsetAsync(A->pcg_sync, 0, sizeof(PcgSync), c->stream));
    PF2_CUDA(cudaEventRecord(A->pev[0][0], c->stream));
    int rc = PF2_OK;
#define PCG5(IDXT, NBV, M, D, CSV) rc = pcg_launch<IDXT, NBV, M, D, CSV>(A, a)
#define PCG4(IDXT, NBV, M, D) { if (cs) PCG5(IDXT, NBV, M, D, true); else PCG5(IDXT, NBV, M, D, false); }
#define PCG3(IDXT, NBV, M) { if (dist) PCG4(IDXT, NBV, M, true) else PCG4(IDXT, NBV, M, false) }
#define PCG2(IDXT, NBV) { if (mode) PCG3(IDXT, NBV, 1) else PCG3(IDXT, NBV, 0) }
    if (A->sell_nb == 3) { if (wide) PCG2(int, 3) else PCG2(short, 3) }
    else { if (wide) PCG2(int, 1) else PCG2(short, 1) }
#undef PCG2
#undef PCG3
#undef PCG4
#undef PCG5
    PF2_TRY(rc);
    PF2_CUDA(cudaEventRecord(A->pev[0][1], c->stream));
    PF2_CUDA(cudaMemcpyAsync(&A->h_st[0], A->st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaMemcpyAsync(A->h_pcg_sync, &((PcgSync*)A->pcg_sync)->t_ns[0], sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    const CgState last = A->h_st[0];
    const unsigned long long* t_ns = (const unsigned long long*)A->h_pcg_sync;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, A->pev[0][0], A->pev[0][1]) == cudaSuccess) A->pcg_kernel_ms += ms; else cudaGetLastError();
    A->pcg_iters += last.iter; A->pcg_solves++;
    for (int j = 0; j < 3; j++) { A->pcg_phase_ns[j] += (double)t_ns[j]; A->pcg_wait_ns[j] += (double)t_ns[4 + j]; }
    A->total_iters += last.iter;
    if (iters_out) *iters_out = last.iter;
    if (relres_out) *relres_out = sqrt(last.rr) / sqrt(last.bb);
    if (last.done == 2) { set_error("persistent PCG: a synchronisation wait timed out (a peer rank left the solve?)"); return PF2_E_CUDA; }
    if (!last.done) {
        set_error("Convergence:faild after %d iterations (relres %.3e)", last.iter, sqrt(last.rr) / sqrt(last.bb));
        return PF2_E_NOCONV;
    }
    return PF2_OK;
}

}  // namespace pf2

// 1 when this matrix' SELL-32 mirror is one the persistent kernel's partitioned instantiation handles (natural row order, no 2-wide
// block deltas); the ranks of a partition exchange it so that all of them take the same path
extern "C" int pf2_csr_pcg_capable(pf2_csr* A, int* out) {
    using namespace pf2;
    PF2_CHECK(A && out, "null argument");
    PF2_TRY(plan_spmv_pub(A));
    *out = (A->spmv_variant == 31 && A->sell_perm == nullptr && A->sell_nb != 2) ? 1 : 0;
    return PF2_OK;
}

// diagnostics: per-CTA %globaltimer stamps of iteration kPcgDbgIter of the last persistent solve (6 x 2048 u64)
extern "C" int pf2_csr_pcg_debug(pf2_csr* A, unsigned long long* out_host) {
    using namespace pf2;
    PF2_CHECK(A && A->pcg_sync && out_host, "no persistent solve yet");
    PF2_CUDA(cudaMemcpy(out_host, &((PcgSync*)A->pcg_sync)->dbg[0][0], sizeof(unsigned long long) * 6 * kPcgMaxCtas, cudaMemcpyDeviceToHost));
    return PF2_OK;
}
