// simp.cu -- the device-resident SIMP design loop.
//
// One pf2_simp_iterate call is one pass of the loop body of sample/optimize/sample_optimize_density_oc.cpp:83-208
// (and ..._mma.cpp, identical up to the optimiser call) with every field -- design s, density rho, K, F, u,
// sensitivities -- resident in HBM from the first iteration to the last:
//    :85-88   beta doubling                        host scalar
//    :93      rho = filter(s)                      filter_apply (+ volume sum for :99-105)
//    :113-129 BCs, element loop, Assembling, loads assemble_device   (numbering and pattern were built once)
//    :131-133 CSR, ScalingCG, Disassembling        solve + disassemble
//    :136-162 reaction, compliance, sensitivities  compliance_sens_device
//    :168-169 filtered sensitivities               filter_sens (dfds and dgds in one pass)
//    :192-195 IsConvergence                        host compare of two scalars
//    :198-207 OC / MMA update                      oc_update / mma_update
// Only scalars (f, g, iteration counts, done flags) cross to the host.
#include "types.cuh"

namespace pf2 {
int assemble_device(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq, const double* modulus_dev, const double* rho_dev,
                    const double params[5], int nload, const int* load_node_dev, const int* load_dof_dev, const double* load_val_dev);
int compliance_sens_device(pf2_mesh* mesh, int eq, const double* u_nodal, const double* rho, const double params[6], double* f_dev,
                           double* dfdrho, double* r_nodal);
int solve_x0(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int warm, int* iters_out, double* relres_out);
int filter_apply(pf2_filter* f, const double* s, double* rho, double* sum_out, OcState* oc);
int filter_sens(pf2_filter* f, const double* s, const double* g1, double* out1, const double* g2, double c2, double* out2);
int oc_update(pf2_oc* oc, pf2_filter* filter, double weightlimit, double scale1, double* x, double f, const double* dfdx,
              const double* dgdx, int* steps_out, double* lambda_out);
int mma_update(pf2_mma* mm, double* xk, double f, const double* dfdx, const double* g_host, const double* dgdx, int* newton_out);
int dist_halo(pf2_dist* d, double* vec, const int halo[6]);
int dist_allreduce(pf2_dist* d, double* dev, int count);
}  // namespace pf2

using namespace pf2;

// ---- VTK staging: one kernel lays the point and cell fields out exactly as ExportToVTK.h writes them (coordinates and vectors padded
// to three components, ExportToVTK.h:31-37,112-117), so that ONE device-to-host copy into pinned memory feeds the writer ----
__global__ void vtk_stage_kernel(int nnode, int dim, int ndof, int nelem, const double* __restrict__ coords, const double* __restrict__ u,
                                 const double* __restrict__ r, const double* __restrict__ rho, double* __restrict__ out) {
    const size_t np3 = (size_t)nnode * 3;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < np3; k += (size_t)gridDim.x * blockDim.x) {
        const size_t i = k / 3;
        const int d = (int)(k % 3);
        out[k] = d < dim ? coords[i * dim + d] : 0.0;
        out[np3 + k] = d < ndof ? u[i * ndof + d] : 0.0;
        if (r) out[2 * np3 + k] = d < ndof ? r[i * ndof + d] : 0.0;
    }
    double* cell = out + (r ? 3 : 2) * np3;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)nelem; e += (size_t)gridDim.x * blockDim.x) cell[e] = rho[e];
}

struct pf2_simp {
    pf2_ctx* ctx = nullptr;
    pf2_mesh* mesh = nullptr;
    pf2_dofmap* map = nullptr;
    pf2_csr* A = nullptr;
    pf2_filter* filter = nullptr;
    pf2_oc* oc = nullptr;
    pf2_mma* mma = nullptr;
    int eq = 0, opt_kind = 0, solver = PF2_SOLVER_SCALINGCG;
    int n = 0, ndof = 0;
    double E0, E1, V, p, weightlimit, scale0, scale1, thick, beta;
    int beta_period = 0, itrmax = 100000;
    double cgeps = 1.0e-10;
    int k = 0;
    double beta0 = 0.0;          // Heaviside beta of pf2_simp_create (restored by pf2_simp_reset)
    int warm_start = 0;          // opt-in: PCG starts from the previous design iteration's displacements (pf2_simp_set_warm_start)
    bool have_solution = false;  // xsol holds a converged solution of this run
    int nload = 0;
    int *ld_node = nullptr, *ld_dof = nullptr;
    double* ld_val = nullptr;
    double *s = nullptr, *rho = nullptr, *xsol = nullptr, *u = nullptr, *r_nodal = nullptr, *dfdrho = nullptr, *dfds = nullptr, *dgds = nullptr;
    cudaEvent_t ev[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    double phase_ms[6] = { 0, 0, 0, 0, 0, 0 };
    // multi-GPU (row-block partition): this object drives ONE slab; n_global = elements of the whole problem
    pf2_dist* dist = nullptr;
    int ehalo[6] = { 0, 0, 0, 0, 0, 0 };
    long long n_global = 0;
    double* vtk_dev = nullptr;       // staging of pf2_simp_export_vtk (device) and its pinned host mirror
    double* vtk_host = nullptr;
    int* vtk_conn_host = nullptr;    // connectivity, downloaded once (it never changes)
};

static int simp_iterate(pf2_simp* S, int check_convergence, double stats[8]) {
    pf2_ctx* c = S->ctx;
    cudaStream_t s = c->stream;
    pf2_dist* d = S->dist;
    const double nglob = (double)(S->n_global ? S->n_global : S->n);
    PF2_CUDA(cudaSetDevice(c->device));
    NvtxRange nv_iter("pf2_simp_iterate");
    PF2_CUDA(cudaEventRecord(S->ev[0], s));
    NvtxRange nv_phase("pf2 filter");
    if (S->beta_period > 0 && S->k % S->beta_period == 0) S->beta *= 2.0;       // driver :85-88
    S->filter->beta = S->beta;
    if (d) PF2_TRY(dist_halo(d, S->s, S->ehalo));                              // ghost element planes of the design
    PF2_TRY(filter_apply(S->filter, S->s, S->rho, c->scalars + 1, nullptr));
    if (d) { PF2_TRY(dist_allreduce(d, c->scalars + 1, 1)); PF2_TRY(dist_halo(d, S->rho, S->ehalo)); }
    PF2_CUDA(cudaEventRecord(S->ev[1], s));
    nv_phase.next("pf2 assemble");
    const double ap[5] = { S->E0, S->E1, S->V, S->p, S->thick };
    PF2_TRY(assemble_device(S->A, S->mesh, S->map, S->eq, nullptr, S->rho, ap, S->nload, S->ld_node, S->ld_dof, S->ld_val));
    PF2_CUDA(cudaEventRecord(S->ev[2], s));
    nv_phase.next("pf2 solve");
    int iters = 0;
    double relres = 0.0;
    int rc = solve_x0(S->A, S->solver, S->A->F, S->xsol, S->itrmax, S->cgeps, (S->warm_start && S->have_solution) ? 1 : 0, &iters, &relres);
    if (rc != PF2_OK && rc != PF2_E_NOCONV) return rc;      // non-convergence: the reference prints and carries on
    S->have_solution = true;
    if (d) PF2_TRY(dist_halo(d, S->xsol, S->A->halo));                         // displacements of the ghost node planes
    PF2_TRY(pf2_disassemble(S->map, S->xsol, S->u));
    PF2_CUDA(cudaEventRecord(S->ev[3], s));
    nv_phase.next("pf2 compliance+sensitivity");
    const double sp[6] = { S->E0, S->E1, S->V, S->p, S->thick, S->scale0 };
    PF2_TRY(compliance_sens_device(S->mesh, S->eq, S->u, S->rho, sp, c->scalars, S->dfdrho, nullptr));
    if (d) { PF2_TRY(dist_allreduce(d, c->scalars, 1)); PF2_TRY(dist_halo(d, S->dfdrho, S->ehalo)); }
    PF2_CUDA(cudaEventRecord(S->ev[4], s));
    nv_phase.next("pf2 filter sensitivities");
    const double dgdrho = S->scale1 / (S->weightlimit * nglob);                 // driver :104
    PF2_TRY(filter_sens(S->filter, S->s, S->dfdrho, S->dfds, nullptr, dgdrho, S->dgds));
    if (d) { PF2_TRY(dist_halo(d, S->dfds, S->ehalo)); PF2_TRY(dist_halo(d, S->dgds, S->ehalo)); }
    PF2_CUDA(cudaEventRecord(S->ev[5], s));
    nv_phase.next("pf2 optimiser update");
    PF2_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    const double f = c->h_scalars[0];
    const double g = S->scale1 * c->h_scalars[1] / (S->weightlimit * nglob) - 1.0 * S->scale1;   // driver :99-105
    int converged = 0;
    if (S->oc) PF2_TRY(pf2_oc_is_convergence(S->oc, f, &converged));
    else PF2_TRY(pf2_mma_is_convergence(S->mma, f, &converged));
    int opt_steps = 0;
    if (!(check_convergence && converged)) {
        if (S->oc) PF2_TRY(oc_update(S->oc, S->filter, S->weightlimit, S->scale1, S->s, f, S->dfds, S->dgds, &opt_steps, nullptr));
        else PF2_TRY(mma_update(S->mma, S->s, f, S->dfds, &g, S->dgds, &opt_steps));
    }
    PF2_CUDA(cudaEventRecord(S->ev[6], s));
    PF2_CUDA(cudaEventSynchronize(S->ev[6]));
    for (int i = 0; i < 6; i++) {
        float ms = 0;
        PF2_CUDA(cudaEventElapsedTime(&ms, S->ev[i], S->ev[i + 1]));
        S->phase_ms[i] = ms;
    }
    if (stats) {
        stats[0] = f; stats[1] = g; stats[2] = (check_convergence && converged) ? 1.0 : 0.0; stats[3] = iters; stats[4] = relres;
        stats[5] = opt_steps; stats[6] = S->beta; stats[7] = S->k;
    }
    S->k++;
    return PF2_OK;
}

extern "C" {

int pf2_simp_create(pf2_ctx* ctx, pf2_mesh* mesh, pf2_dofmap* map, pf2_csr* A, pf2_filter* filter, int eq, int opt_kind,
                    const double* optp, const double params[12], int nload, const int* load_node_host,
                    const int* load_dof_host, const double* load_val_host, pf2_simp** out) {
    PF2_CHECK(ctx && mesh && map && A && filter && optp && params && out, "null argument");
    PF2_CHECK(filter->n == mesh->nelem, "filter size must equal the element count");
    PF2_CHECK(opt_kind == PF2_OPT_OC || opt_kind == PF2_OPT_MMA || opt_kind == PF2_OPT_CONLIN, "unknown optimiser");
    PF2_CUDA(cudaSetDevice(ctx->device));
    pf2_simp* S = new pf2_simp();
    S->ctx = ctx; S->mesh = mesh; S->map = map; S->A = A; S->filter = filter; S->eq = eq; S->opt_kind = opt_kind;
    S->n = mesh->nelem; S->ndof = map->ndof;
    S->E0 = params[0]; S->E1 = params[1]; S->V = params[2]; S->p = params[3]; S->weightlimit = params[4];
    S->scale0 = params[5]; S->scale1 = params[6]; S->thick = params[7]; S->beta = params[8];
    S->beta_period = (int)params[9]; S->itrmax = (int)params[10]; S->cgeps = params[11];
    S->beta0 = S->beta;
    const size_t n = (size_t)S->n, nd = (size_t)mesh->nnode * map->ndof;
    PF2_TRY(dev_alloc(&S->s, n)); PF2_TRY(dev_alloc(&S->rho, n)); PF2_TRY(dev_alloc(&S->dfdrho, n));
    PF2_TRY(dev_alloc(&S->dfds, n)); PF2_TRY(dev_alloc(&S->dgds, n));
    PF2_TRY(dev_alloc(&S->xsol, (size_t)A->rows)); PF2_TRY(dev_alloc(&S->u, nd)); PF2_TRY(dev_alloc(&S->r_nodal, nd));
    PF2_CUDA(cudaMemsetAsync(S->s, 0, sizeof(double) * n, ctx->stream));
    S->nload = nload;
    if (nload > 0) {
        PF2_TRY(dev_alloc(&S->ld_node, (size_t)nload)); PF2_TRY(dev_alloc(&S->ld_dof, (size_t)nload)); PF2_TRY(dev_alloc(&S->ld_val, (size_t)nload));
        PF2_CUDA(cudaMemcpyAsync(S->ld_node, load_node_host, sizeof(int) * (size_t)nload, cudaMemcpyHostToDevice, ctx->stream));
        PF2_CUDA(cudaMemcpyAsync(S->ld_dof, load_dof_host, sizeof(int) * (size_t)nload, cudaMemcpyHostToDevice, ctx->stream));
        PF2_CUDA(cudaMemcpyAsync(S->ld_val, load_val_host, sizeof(double) * (size_t)nload, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (opt_kind == PF2_OPT_OC) {
        PF2_TRY(pf2_oc_create(ctx, S->n, optp[0], optp[1], optp[2], optp[3], optp[4], &S->oc));
    } else if (opt_kind == PF2_OPT_MMA) {
        std::vector<double> xmin(n, optp[11]), xmax(n, optp[12]);
        PF2_TRY(pf2_mma_create(ctx, S->n, 1, optp[7], &optp[8], &optp[9], &optp[10], xmin.data(), xmax.data(), &S->mma));
        PF2_TRY(pf2_mma_set_parameters(S->mma, optp[0], optp[1], optp[2], optp[3], optp[4], optp[5], optp[6]));
    } else {
        std::vector<double> xmin(n, optp[6]), xmax(n, optp[7]);
        PF2_TRY(pf2_conlin_create(ctx, S->n, 1, optp[2], &optp[3], &optp[4], &optp[5], xmin.data(), xmax.data(), &S->mma));
        PF2_TRY(pf2_conlin_set_parameters(S->mma, optp[0], optp[1]));
    }
    for (int i = 0; i < 7; i++) PF2_CUDA(cudaEventCreate(&S->ev[i]));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = S;
    return PF2_OK;
}

int pf2_simp_destroy(pf2_simp* S) {
    if (!S) return PF2_OK;
    cudaStreamSynchronize(S->ctx->stream);
    pf2_oc_destroy(S->oc);
    pf2_mma_destroy(S->mma);
    void* ptrs[] = { S->s, S->rho, S->xsol, S->u, S->r_nodal, S->dfdrho, S->dfds, S->dgds, S->ld_node, S->ld_dof, S->ld_val };
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < 7; i++) if (S->ev[i]) cudaEventDestroy(S->ev[i]);
    if (S->vtk_dev) cudaFree(S->vtk_dev);
    if (S->vtk_host) cudaFreeHost(S->vtk_host);
    if (S->vtk_conn_host) free(S->vtk_conn_host);
    delete S;
    return PF2_OK;
}

int pf2_simp_set_design(pf2_simp* S, const double* s_host) {
    PF2_CUDA(cudaMemcpyAsync(S->s, s_host, sizeof(double) * (size_t)S->n, cudaMemcpyHostToDevice, S->ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(S->ctx->stream));
    return PF2_OK;
}
// Back to iteration 0 of the driver (sample_optimize_density_oc.cpp:78-83): design, Heaviside beta, the optimiser's history
// (previousvalue, iteration count: MMA re-initialises its asymptotes for k < 2, MMA.h:133-141) and the warm-start state.
int pf2_simp_reset(pf2_simp* S, const double* s_host) {
    PF2_CHECK(S && s_host, "null argument");
    PF2_TRY(pf2_simp_set_design(S, s_host));
    S->k = 0; S->beta = S->beta0; S->have_solution = false;
    if (S->oc) { S->oc->k = 0; S->oc->previousvalue = 0.0; }
    if (S->mma) { S->mma->k = 0; S->mma->previousvalue = 0.0; }
    return PF2_OK;
}
int pf2_simp_set_warm_start(pf2_simp* S, int on) {
    PF2_CHECK(S, "null argument");
    S->warm_start = on ? 1 : 0;
    if (!on) S->have_solution = false;
    return PF2_OK;
}
int pf2_simp_set_partition(pf2_simp* S, pf2_dist* d, int own_elem_lo, int own_elem_hi, const int elem_halo[6], long long n_global_elems) {
    PF2_CHECK(S && d && elem_halo && own_elem_lo >= 0 && own_elem_lo <= own_elem_hi && own_elem_hi <= S->n && n_global_elems >= S->n - 0, "bad partition");
    PF2_CHECK(S->A->dist == d, "call pf2_csr_set_partition on the matrix first");
    S->dist = d; S->n_global = n_global_elems;
    for (int i = 0; i < 6; i++) { S->ehalo[i] = elem_halo[i]; S->filter->ehalo[i] = elem_halo[i]; }
    S->filter->dist = d; S->filter->sum_lo = own_elem_lo; S->filter->sum_hi = own_elem_hi;
    S->mesh->own_elem_lo = own_elem_lo; S->mesh->own_elem_hi = own_elem_hi;
    if (S->oc) S->oc->n_global = n_global_elems;
    if (S->mma) { S->mma->dist = d; S->mma->lo = own_elem_lo; S->mma->hi = own_elem_hi; }
    return PF2_OK;
}

int pf2_simp_set_solver(pf2_simp* S, int solver) {
    PF2_CHECK(solver >= 0 && solver <= 2, "unknown solver");
    S->solver = solver;
    return PF2_OK;
}

int pf2_simp_iterate(pf2_simp* S, int check_convergence, double stats[8]) { return simp_iterate(S, check_convergence, stats); }

int pf2_simp_iterate_host(pf2_simp* S, int check_convergence, const double* s_in_host, double* s_out_host, double* rho_out_host, double stats[8]) {
    cudaStream_t s = S->ctx->stream;
    const size_t bytes = sizeof(double) * (size_t)S->n;
    if (s_in_host) PF2_CUDA(cudaMemcpyAsync(S->s, s_in_host, bytes, cudaMemcpyHostToDevice, s));
    PF2_TRY(simp_iterate(S, check_convergence, stats));
    if (s_out_host) PF2_CUDA(cudaMemcpyAsync(s_out_host, S->s, bytes, cudaMemcpyDeviceToHost, s));
    if (rho_out_host) PF2_CUDA(cudaMemcpyAsync(rho_out_host, S->rho, bytes, cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    return PF2_OK;
}

int pf2_simp_get(pf2_simp* S, double* s_host, double* rho_host, double* u_nodal_host, double* r_nodal_host) {
    pf2_ctx* c = S->ctx;
    cudaStream_t s = c->stream;
    const size_t nb = sizeof(double) * (size_t)S->n, ndb = sizeof(double) * (size_t)S->mesh->nnode * S->ndof;
    if (s_host) PF2_CUDA(cudaMemcpyAsync(s_host, S->s, nb, cudaMemcpyDeviceToHost, s));
    if (rho_host) PF2_CUDA(cudaMemcpyAsync(rho_host, S->rho, nb, cudaMemcpyDeviceToHost, s));
    if (u_nodal_host) PF2_CUDA(cudaMemcpyAsync(u_nodal_host, S->u, ndb, cudaMemcpyDeviceToHost, s));
    if (r_nodal_host) {
        // reaction forces r = K_full(rho) u (driver :136-150), recomputed on demand (the VTK dump is off the hot path)
        const double sp[6] = { S->E0, S->E1, S->V, S->p, S->thick, S->scale0 };
        PF2_TRY(compliance_sens_device(S->mesh, S->eq, S->u, S->rho, sp, c->scalars + 2, nullptr, S->r_nodal));
        PF2_CUDA(cudaMemcpyAsync(r_nodal_host, S->r_nodal, ndb, cudaMemcpyDeviceToHost, s));
    }
    PF2_CUDA(cudaStreamSynchronize(s));
    return PF2_OK;
}

// The drivers' per-iteration dump (sample_optimize_density_oc.cpp:175-184): MakeHeadderToVTK, AddPointsToVTK, AddElementToVTK,
// AddElementTypes, AddPointVectors u [, r], AddElementScalers rho as "s" (ExportToVTK.h:19-137), written from the device-resident
// state.  The fields are staged on the device in the writer's layout and cross in one copy; numbers are formatted like the
// reference's `ostream << double` (default float format = %g with 6 significant digits), so the file is byte-identical to the one
// the reference's writers produce from the same values.
int pf2_simp_export_vtk(pf2_simp* S, const char* path, int cell_type, int with_reactions) {
    PF2_CHECK(S && path, "null argument");
    pf2_ctx* c = S->ctx;
    cudaStream_t s = c->stream;
    PF2_CUDA(cudaSetDevice(c->device));
    const int nnode = S->mesh->nnode, nelem = S->n, npe = S->mesh->npe, dim = S->mesh->dim, ndof = S->ndof;
    const size_t np3 = (size_t)nnode * 3, total = 3 * np3 + (size_t)nelem;
    if (!S->vtk_dev) {
        PF2_TRY(dev_alloc(&S->vtk_dev, total));
        PF2_CUDA(cudaHostAlloc((void**)&S->vtk_host, total * sizeof(double), cudaHostAllocDefault));
        S->vtk_conn_host = (int*)malloc(sizeof(int) * (size_t)nelem * npe);
        PF2_CHECK(S->vtk_conn_host, "out of host memory");
        PF2_CUDA(cudaMemcpyAsync(S->vtk_conn_host, S->mesh->conn, sizeof(int) * (size_t)nelem * npe, cudaMemcpyDeviceToHost, s));
    }
    if (with_reactions) {
        const double sp[6] = { S->E0, S->E1, S->V, S->p, S->thick, S->scale0 };
        PF2_TRY(compliance_sens_device(S->mesh, S->eq, S->u, S->rho, sp, c->scalars + 2, nullptr, S->r_nodal));      // driver :136-150
    }
    vtk_stage_kernel<<<c->grid_for((long long)np3), kThreads, 0, s>>>(nnode, dim, ndof, nelem, S->mesh->coords, S->u, with_reactions ? S->r_nodal : nullptr,
                                                                    S->rho, S->vtk_dev);
    PF2_LAUNCH_CHECK();
    c->launches++;
    const size_t used = (with_reactions ? 3 : 2) * np3 + (size_t)nelem;
    PF2_CUDA(cudaMemcpyAsync(S->vtk_host, S->vtk_dev, used * sizeof(double), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    FILE* f = fopen(path, "w");
    if (!f) { set_error("cannot open %s for writing", path); return PF2_E_INVALID; }
    fputs("# vtk DataFile Version 4.1\nvtk output\nASCII\nDATASET UNSTRUCTURED_GRID\n", f);                  // ExportToVTK.h:19-24
    fprintf(f, "\nPOINTS\t%d\tfloat\n", nnode);                                                                   // :29-40
    const double* P = S->vtk_host;
    for (int i = 0; i < nnode; i++) fprintf(f, "%g\t%g\t%g\t\n", P[3 * (size_t)i], P[3 * (size_t)i + 1], P[3 * (size_t)i + 2]);
    fprintf(f, "\nCELLS %d\t%lld\n", nelem, (long long)nelem * (npe + 1));                                        // :44-57
    for (int e = 0; e < nelem; e++) {
        fprintf(f, "%d\t", npe);
        for (int a = 0; a < npe; a++) fprintf(f, "%d\t", S->vtk_conn_host[(size_t)e * npe + a]);
        fputc('\n', f);
    }
    fprintf(f, "\nCELL_TYPES\t%d\n", nelem);                                                                       // :61-66
    for (int e = 0; e < nelem; e++) fprintf(f, "%d\n", cell_type);
    fprintf(f, "\nPOINT_DATA\t%d\n", nnode);                                                                       // :103-119
    const char* names[2] = { "u", "r" };
    for (int v = 0; v < (with_reactions ? 2 : 1); v++) {
        fprintf(f, "VECTORS %s float\n", names[v]);
        const double* V = S->vtk_host + (size_t)(v + 1) * np3;
        for (int i = 0; i < nnode; i++) fprintf(f, "%g\t%g\t%g\t\n", V[3 * (size_t)i], V[3 * (size_t)i + 1], V[3 * (size_t)i + 2]);
    }
    fprintf(f, "\nCELL_DATA\t%d\nSCALARS s float\nLOOKUP_TABLE default\n", nelem);                                // :125-135
    const double* R = S->vtk_host + (with_reactions ? 3 : 2) * np3;
    for (int e = 0; e < nelem; e++) fprintf(f, "%g\n", R[e]);
    if (fclose(f) != 0) { set_error("write to %s failed", path); return PF2_E_INVALID; }
    return PF2_OK;
}

int pf2_simp_phase_ms(pf2_simp* S, double ms[6]) {
    for (int i = 0; i < 6; i++) ms[i] = S->phase_ms[i];
    return PF2_OK;
}
int pf2_simp_cg_stats(pf2_simp* S, double* spmv_ms_avg, long long* spmv_calls) {
    double st[8];
    PF2_TRY(pf2_csr_solver_stats(S->A, st));
    if (spmv_ms_avg) *spmv_ms_avg = st[0];
    if (spmv_calls) *spmv_calls = (long long)st[4];
    return PF2_OK;
}

}  // extern "C"
