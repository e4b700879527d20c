// types.cuh -- the opaque handle types behind the C ABI.  All arrays are device pointers unless prefixed h_.
#pragma once
#include "common.cuh"

namespace pf2 {
// decoded PF2_EQ_CODE (assemble_generic.cu)
struct EqInfo {
    int phys, shape, quad, quad2;   // fields of the code with the defaults filled in
    int kind;                       // KIND_ELAST2D / KIND_HEAT2D / KIND_SOLID3D
    int dim, npe, ndof;
    bool fast;                      // one of the three specialised selections (element.cuh)
    int legacy;                     // PF2_EQ_PLANESTRAIN / _SOLID / _HEAT of the specialised kernel when fast
    int npass;
    double cn[2], lam[2], mu[2];
};
int decode_eq(int eq, double V, EqInfo* out);
}  // namespace pf2

struct pf2_mesh {
    pf2_ctx* ctx = nullptr;
    int dim = 0, nnode = 0, npe = 0, nelem = 0;
    int own_elem_lo = 0, own_elem_hi = 0;   // multi-GPU: owned elements (compliance is summed over them only)
    double* coords = nullptr;   // nnode*dim, node-major (AoS as std::vector<Vector<T>>)
    bool shares_coords = false; // pf2_mesh_create_on_nodes: the coordinates belong to another mesh
    int* conn = nullptr;        // nelem*npe
};

struct pf2_dofmap {
    pf2_ctx* ctx = nullptr;
    int nnode = 0, ndof = 0, kdegree = 0;
    int* n2g = nullptr;         // nnode*ndof : -1 = Dirichlet, else global row (node-major, dof-minor)
    double* ufix = nullptr;     // nnode*ndof : prescribed value on fixed dofs, 0 elsewhere
};

namespace pf2 {
constexpr int kGatherTile = 128;   // nodes (= threads) per CTA of the gather assembly (assemble_gather.cuh)
constexpr int kGather3Tile = 32;   // nodes per CTA of the hex8 gather assembly (three threads per node)
// the hex8 gather plan is built on request (PF2_ASSEMBLE_GATHER3D=1 before pf2_csr_pattern): 4 bytes per element node on top of the pattern
inline bool gather3d_requested() {
    const char* e = getenv("PF2_ASSEMBLE_GATHER3D");
    return e != nullptr && atoi(e) != 0;
}
constexpr int kCsrPad = 8;   // spare entries behind indptr / indices / data (see spmv_tma.cuh)
// device-resident state of one Krylov solve (CG.h:124-154 / 420-453 / 320-352)
struct CgState {
    double rho;      // z.r of the current residual (Mrkrk)
    double pAp;
    double rr;       // r.r
    double bb;       // b.b
    double beta;
    double zr_new;   // scratch for the ILU path
    double red[4];   // partial sums awaiting an allreduce (multi-GPU)
    int iter;        // iterations completed
    int done;        // 1 = converged (x frozen), kernels early-exit
    int maxit;
    int pad;
    double eps;
};
}  // namespace pf2

struct pf2_dist;
namespace pf2 {
constexpr int kMaxRanksT = 8;
struct P2PView {          // must match pf2::P2P in dist.cu (kept POD here so that pf2_csr can embed it)
    int rank, world;
    double* slots[kMaxRanksT];
    unsigned long long* flags[kMaxRanksT];
    unsigned long long* halo_flags[kMaxRanksT];
    double* left_p;
    double* right_p;
    int left_recv_off, right_recv_off;
    unsigned long long* ll[kMaxRanksT];    // LL words of the cross-GPU sums: [parity][sender][4 terms][2 halves] in every rank's arena
    unsigned int* abort;                   // local device word: a bounded spin timed out (a peer left the solve)
    // halo geometry of the local matrix (rows): owned range, the two boundary planes this rank sends; defer = 1: the wait for the
    // neighbours' planes happens in the boundary slices of the next product instead of at the end of the p-update
    int own_lo, own_hi, sendL, cntL, sendR, cntR, defer_halo_wait, pad;
    unsigned long long* epoch;             // local device words: [0] allreduce epoch, [1] halo epoch
    // single-reduction CG (solve_dist_cg1): non-null while such a solve runs -- the product's last CTA then reduces {w.u, u.r, r.r} in ONE
    // cross-GPU sum and runs the scalar tail (alpha, beta, stopping test) on this state
    CgState* cg1;
};
}  // namespace pf2

struct pf2_csr {
    pf2_ctx* ctx = nullptr;
    int rows = 0;
    // row-block partition (multi-GPU): rows [own_lo, own_hi) are owned, the rest are ghost rows whose x entries arrive
    // by halo exchange; single GPU: [0, rows)
    int own_lo = 0, own_hi = 0;
    pf2_dist* dist = nullptr;
    int halo[6] = { 0, 0, 0, 0, 0, 0 };   // sendL_off, recvL_off, cntL, sendR_off, recvR_off, cntR (row offsets)
    bool p2p_ready = false;               // peer-memory backend: neighbours' p vectors and all arenas are mapped
    pf2::P2PView p2p_view;
    pf2::P2PView* p2p_dev = nullptr;      // device copy handed to the fused kernels (nullptr: single GPU or NCCL backend)
    unsigned long long* p2p_epoch = nullptr;
    // partitioned PCG recurrences: 0 = the reference's (two cross-GPU sums per iteration), 1 = single-reduction (Chronopoulos-Gear: one sum,
    // two kernels per iteration; same iterates in exact arithmetic, round-off differs), -1 = from the environment (PF2_CG_SINGLE_REDUCTION)
    int cg_variant = -1;
    double* cg1_s = nullptr;              // s = A p of the single-reduction recurrences
    long long cg1_solves = 0;             // solves that ran them (pf2_csr_pcg_stats out[11])
    bool pcg_dist_ok = false;             // every rank's slab qualifies for the persistent kernel's partitioned instantiation
    void* p2p_opened[2] = { nullptr, nullptr };   // IPC mappings of the left / right neighbour's Krylov slab
    long long nnz = 0;
    long long* indptr = nullptr;   // rows+1 (int64: config 5 has nnz > 2^31)
    int* indices = nullptr;        // nnz, sorted within a row
    double* data = nullptr;        // nnz
    double* F = nullptr;           // rows: right-hand side assembled next to K
    int* diagpos = nullptr;        // rows: offset of the diagonal inside the row, -1 if structurally absent (CSR.h:155-167)
    int max_row = 0;               // longest row
    // scatter map of the symbolic phase (pattern-built matrices only)
    int* bmap = nullptr;           // [(e*npe+a)*npe + b] : column offset of node b's first free dof in node a's rows
    int map_npe = 0, map_ndof = 0, map_nelem = 0;
    // gather plan of the numeric phase (assemble_gather.cuh): node -> adjacent elements in ascending order, rows before each node,
    // and the shared-memory bytes the largest node tile needs (0: plan not usable, the scatter kernels run)
    int* n2e_ptr = nullptr;        // nnode + 1
    int* n2e = nullptr;            // nelem * npe
    int* node_row0 = nullptr;      // nnode + 1 : free dofs in the nodes before this one (= first row of the node)
    int gather_nnode = 0;
    size_t gather_smem = 0;
    // SpMV plan
    int spmv_variant = 0;          // 0 = not planned
    int tma_stages = 3, tma_ctas_per_sm = 4;   // TMA pipeline depth and residency target
    // SELL-32 mirror (variant 31): slice pointers, column-major indices / values, stored entries
    long long* sell_ptr = nullptr;
    int* sell_perm = nullptr;      // SELL-C-sigma: row of every slot (slice*32 + lane), -1 = padding; null = natural order
    int* sell_idx = nullptr;       // absolute columns (dropped when the 16-bit delta form is usable)
    short* sell_d16 = nullptr;     // column - row, 2 bytes per stored entry (|delta| <= 32767)
    int sell_nb = 1;               // > 1: ONE node-unit delta per run of sell_nb consecutive columns, in sell_d16 or (wide meshes) sell_b32
    int* sell_b32 = nullptr;
    int sell_max_delta = 0;
    double* sell_val = nullptr;
    long long sell_entries = 0;
    bool sell_values_valid = false;
    // Krylov workspace (lazily allocated)
    double *r = nullptr, *p = nullptr, *z = nullptr, *y = nullptr, *dvec = nullptr, *xw = nullptr, *bw = nullptr;
    pf2::CgState* st = nullptr;
    pf2::CgState* h_st = nullptr;  // pinned, 2 slots
    cudaEvent_t ev[2] = { nullptr, nullptr };
    // ILU(0)
    double* ilu = nullptr;         // factors in A's pattern
    bool ilu_valid = false;
    int* level_rows = nullptr;     // rows sorted by dependency level of the forward (unit-L) sweep
    int* level_rows_u = nullptr;   // ... of the backward (U) sweep
    std::vector<int> h_level_ptr, h_level_ptr_u;   // host: first row of each level in the arrays above
    int* level_ptr = nullptr;      // device copies of the level pointers (one-launch sweeps)
    int* level_ptr_u = nullptr;
    int level_width = 0;           // rows of the widest level
    unsigned int* ilu_ready = nullptr;   // sync-free sweeps: ready[i] == epoch <=> v[i] is final in the sweep numbered `epoch`
    unsigned int ilu_epoch = 0;
    double* slab = nullptr;        // r | p | z | y | dvec contiguous (one L2 access-policy window)
    // instrumentation: every chunk one iteration is bracketed by events (sampled per-kernel device time)
    cudaEvent_t pev[2][4] = { { nullptr, nullptr, nullptr, nullptr }, { nullptr, nullptr, nullptr, nullptr } };
    bool pev_armed[2] = { false, false };
    // matrix-free operator on a uniform structured mesh (spmv_mf.cuh, SpMV variant 41; opt-in)
    bool mf_ready = false;
    int mf_dim = 0, mf_ndof = 0, mf_eq = 0, mf_n[3] = { 0, 0, 0 }, mf_nelem = 0;
    double* mf_E = nullptr;           // per-element modulus of the last assembly
    const int* mf_n2g = nullptr;      // borrowed from the dof map
    double mf_ke0[576];               // unit-modulus element matrix (host copy, uploaded to constant memory)
    double mf_V = -1.0, mf_t = -1.0;  // parameters mf_ke0 was built with
    unsigned long long mf_version = 0;
    bool mf_nodal = false;            // single GPU: run the PCG in nodal numbering with the p-update fused into the operator
    double* mf_slab = nullptr;        // nodal-space vectors: b | D | x | r | z | y | p0 | p1
    // BiCGSTAB family workspace (bicgstab.cu): 12 vectors, device state, pinned mirror (2 slots), poll events
    double* bi_slab = nullptr;
    void* bi_st = nullptr;
    void* bi_hst = nullptr;
    cudaEvent_t bi_ev[2] = { nullptr, nullptr };
    // persistent PCG kernel (pcg_persistent.cuh): barrier / timer block, accumulated statistics
    void* pcg_sync = nullptr;
    void* h_pcg_sync = nullptr;        // pinned copy read back after every solve
    int pcg_mode = -1;                 // -1: from the environment (PF2_PCG, default on), 0: three kernels per iteration, 1: persistent kernel
    int pcg_grid = 0;                  // CTAs of the last persistent launch
    double pcg_kernel_ms = 0.0;        // CUDA-event time of the persistent kernels since the last statistics reset
    long long pcg_iters = 0, pcg_solves = 0;
    double pcg_phase_ns[3] = { 0, 0, 0 }, pcg_wait_ns[3] = { 0, 0, 0 };
    double prof_ms[3] = { 0, 0, 0 };   // spmv+dot, update, p-update
    long long prof_samples = 0;
    long long total_iters = 0;
};

struct pf2_filter {
    pf2_ctx* ctx = nullptr;
    int kind = 0, n = 0;
    long long nnb = 0;
    double beta = 1.0;              // HeavisideFilter ctor default (HeavisideFilter.h:37,46)
    long long* rowptr = nullptr;
    int* nbr = nullptr;
    double* w = nullptr;
    double* dr = nullptr;           // scratch: d rho / d s~ (Heaviside chain rule)
    double *hs = nullptr, *hr = nullptr, *hd = nullptr;   // staging for the _host entry points
    // multi-GPU: elements [sum_lo, sum_hi) are owned (volume sums run over them only); ghost planes arrive by halo exchange
    int sum_lo = 0, sum_hi = 0;
    pf2_dist* dist = nullptr;
    int ehalo[6] = { 0, 0, 0, 0, 0, 0 };
};


namespace pf2 {
struct OcState {
    double l0, l1, lambda;
    double eps, volscale, volshift;   // g = volscale * sum(rho) - volshift
    double partial;                   // multi-GPU: this rank's volume sum awaiting the allreduce
    int steps, done;
    int defer, pad;                   // defer = 1: the decision is taken by oc_decide_kernel after the allreduce
};

}  // namespace pf2

struct pf2_oc {
    pf2_ctx* ctx = nullptr;
    int n = 0, k = 0;
    long long n_global = 0;         // design variables of the whole (partitioned) problem; = n on one GPU
    double iota, lmin, lmax, leps, move;
    double previousvalue = 0.0, epsvalue = 1.0e-5;   // OC.h:49-50
    double* xnew = nullptr;
    double* rho = nullptr;
    pf2::OcState* st = nullptr;
    pf2::OcState* h_st = nullptr;   // pinned, 2 slots
    cudaEvent_t ev[2] = { nullptr, nullptr };
};


namespace pf2 {
constexpr int kMaxM = 4;

struct MmaSmall {
    double a0, a[kMaxM], c[kMaxM], d[kMaxM];
    double y[kMaxM], lam[kMaxM], s[kMaxM], mu[kMaxM], z, zeta;
    double dy[kMaxM], dlam[kMaxM], ds[kMaxM], dmu[kMaxM], dz, dzeta;
    double b[kMaxM];
    double eps, tau, tymax, dwl, dwl1;
    double red[32];          // sums / max of the current pass (allreduced across ranks in a partitioned run)
    int accept, ll, halvings, newton;
};

struct MmaParams {
    double raa0, albefa, move, asyinit, asydecr, asyincr;
};

}  // namespace pf2

struct pf2_mma {
    pf2_ctx* ctx = nullptr;
    int n = 0, m = 0, k = 0;
    bool conlin = false;                                  // CONLIN<T> instead of MMA<T>
    double previousvalue = 0.0, epsvalue = 1.0e-5;       // MMA.h:67-68
    pf2::MmaParams P = { 1.0e-5, 0.1, 0.5, 0.5, 0.7, 1.2 };   // MMA.h:80-85
    double *xmin = nullptr, *xmax = nullptr, *xkm1 = nullptr, *xkm2 = nullptr, *L = nullptr, *U = nullptr;
    double *alpha = nullptr, *beta = nullptr, *p0 = nullptr, *q0 = nullptr, *p = nullptr, *q = nullptr;
    double *x = nullptr, *gsi = nullptr, *ita = nullptr, *xn = nullptr, *gsin = nullptr, *itan = nullptr;
    double *dx = nullptr, *dgsi = nullptr, *dita = nullptr, *Dx = nullptr, *dtx = nullptr;
    double* gval = nullptr;
    // multi-GPU: variables [lo, hi) of the local arrays are owned by this rank
    pf2_dist* dist = nullptr;
    int lo = 0, hi = 0;
    pf2::MmaSmall* S = nullptr;
    pf2::MmaSmall* h_S = nullptr;   // pinned
};

