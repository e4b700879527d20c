// levelset.cu -- the device-resident level-set design loop.
//
// One pf2_levelset_iterate call is one pass of the loop body of sample/optimize/sample_optimize_levelset.cpp:75-192:
//    :76-95    BCs, PlaneStressStiffness<Q4, Gauss4Square> with E = Emin + str (E0 - Emin), Assembling, ScalingCG
//                                                   -> modulus_kernel + assemble_device + solve + disassemble
//    :110-122  objective += ue^T Ke ue ; TD = (1e-4 + str (1 - 1e-4)) ue^T Ke(E', c) ue            -> ls_energy_kernel
//    :123      TDN = InterpolateNodalFromElemental (General.h:193-207)                              -> ls_scatter_kernel, ls_nodal_kernel
//    :125-141  vol, ex, lambda, the convergence test, C = nelem / sum |TDN|                         -> ls_scalars_kernel (one thread)
//    :147-172  SetDirichlet(phi), T = Me/dt + Ke(tau nelem), Y = Me/dt phi + Fe(C (TDN - lambda))  -> T is design-independent:
//              (ReactionDiffusion.h:21-48, 82-107, 111-145)                                           assembled ONCE (rd_matrix_kernel);
//                                                                                                     per iteration only rd_rhs_kernel
//    :174-176  ScalingCG, Disassembling                                                             -> solve + ls_update_phi_kernel
//    :178-190  clamp phi to [-1, 1], InterpolateElementalFromNodal, str = (phi_e < 0 ? 0 : 1)      -> ls_update_phi_kernel, ls_str_kernel
// Only the four scalars {objective, vol, lambda, converged} cross to the host per iteration.
#include "types.cuh"
#include "element_generic.cuh"

namespace pf2 {
int assemble_device(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq, const double* modulus_dev, const double* rho_dev,
                    const double params[5], int nload, const int* load_node_dev, const int* load_dof_dev, const double* load_val_dev);
int solve(pf2_csr* A, int solver, const double* b, double* x, int itrmax, double eps, int* iters_out, double* relres_out);

// device-side scalars of the loop
struct LsState {
    double objective, vol, lambda, C;   // of the current iteration
    double tdn_sum, tdn_abs;            // sum TDN, sum |TDN|
    int converged;
    int t;
};

__global__ void ls_modulus_kernel(int nelem, const double* __restrict__ str, double E0, double Emin, double* __restrict__ Emod) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) Emod[e] = Emin + str[e] * (E0 - Emin);
}

// objective[t] = sum_e ue^T Ke(E_e, nu) ue ; TD_e = (1e-4 + str_e (1 - 1e-4)) ue^T Ke(E', c) ue ; vol = sum str
__global__ void __launch_bounds__(kThreads)
ls_energy_kernel(int nelem, ElemSpec sp_k, ElemSpec sp_td, double Etd, const double* __restrict__ coords, const int* __restrict__ conn,
                 const double* __restrict__ u, const double* __restrict__ str, const double* __restrict__ Emod, double* __restrict__ TD,
                 LsState* st, double* objective_hist, double* partials, unsigned int* ticket) {
    double v[2] = { 0.0, 0.0 };
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        const int* nd = conn + (size_t)e * 4;
        double X[4][2], ue[4][2], fe[4][2];
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const int node = nd[n];
            X[n][0] = coords[(size_t)node * 2]; X[n][1] = coords[(size_t)node * 2 + 1];
            ue[n][0] = u[(size_t)node * 2]; ue[n][1] = u[(size_t)node * 2 + 1];
        }
        const double wk = generic_energy<KIND_ELAST2D, SH_Q4, false>(X, ue, sp_k, 1.0, fe);
        const double wt = generic_energy<KIND_ELAST2D, SH_Q4, false>(X, ue, sp_td, 1.0, fe);
        const double s = str[e];
        v[0] += Emod[e] * wk;
        v[1] += s;
        TD[e] = (1.0e-4 + s * (1.0 - 1.0e-4)) * (Etd * wt);
    }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) {
        st->objective = v[0];
        st->vol = v[1] / (double)nelem;
        objective_hist[st->t] = v[0];
    }
}

// InterpolateNodalFromElemental: un[node] = (sum of the adjacent elements' values) / count
__global__ void ls_scatter_kernel(int nelem, const int* __restrict__ conn, const double* __restrict__ TD, double* __restrict__ TDN) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)nelem * 4; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(&TDN[conn[i]], TD[i >> 2]);
}
__global__ void ls_count_kernel(int nelem, const int* __restrict__ conn, int* __restrict__ count) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)nelem * 4; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(&count[conn[i]], 1);
}
__global__ void __launch_bounds__(kThreads)
ls_nodal_kernel(int nnode, const int* __restrict__ count, double* __restrict__ TDN, LsState* st, double* partials, unsigned int* ticket) {
    double v[2] = { 0.0, 0.0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnode; i += gridDim.x * blockDim.x) {
        const double x = TDN[i] / (double)count[i];
        TDN[i] = x;
        v[0] += x;
        v[1] += fabs(x);
    }
    if (grid_sum_last<2>(v, partials, ticket) && threadIdx.x == 0) { st->tdn_sum = v[0]; st->tdn_abs = v[1]; }
}

// driver :125-141 and :143-147
__global__ void ls_scalars_kernel(LsState* st, const double* __restrict__ objective, int nnode, int nelem, double Vmax, double volInit, double nvol,
                                  double p, double d, int check_convergence) {
    const int t = st->t;
    const double vol = st->vol;
    const double ex = Vmax + (volInit - Vmax) * fmax(0.0, 1.0 - (t + 1) / nvol);
    st->lambda = st->tdn_sum / (double)nnode * exp(p * ((vol - ex) / ex + d));
    st->C = (double)nelem / st->tdn_abs;
    int conv = 0;
    if (check_convergence && t > nvol && fabs(vol - Vmax) < 0.005) {
        conv = 1;
        for (int k = 1; k <= 5; k++) conv = conv && (fabs(objective[t] - objective[t - k]) < 0.01 * fabs(objective[t]));
    }
    st->converged = conv;
}

__device__ __forceinline__ void q4_shape(double r0, double r1, double (&N)[4]) {
    N[0] = 0.25 * (1.0 - r0) * (1.0 - r1); N[1] = 0.25 * (1.0 + r0) * (1.0 - r1);
    N[2] = 0.25 * (1.0 + r0) * (1.0 + r1); N[3] = 0.25 * (1.0 - r0) * (1.0 + r1);
}

// T = Me/dt + D Ke, one thread per (element, local node a): row a of the 4x4 element matrix, scattered through bmap
__global__ void rd_matrix_kernel(int nelem, const double* __restrict__ coords, const int* __restrict__ conn, const int* __restrict__ n2g,
                                 const int* __restrict__ bmap, const long long* __restrict__ indptr, double inv_dt, double D, double* __restrict__ data) {
    for (long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x; tid < (long long)nelem * 4; tid += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(tid >> 2), a = (int)(tid & 3);
        const int* nd = conn + (size_t)e * 4;
        const int row = n2g[nd[a]];
        if (row == -1) continue;
        double X[4][2];
#pragma unroll
        for (int n = 0; n < 4; n++) { X[n][0] = coords[(size_t)nd[n] * 2]; X[n][1] = coords[(size_t)nd[n] * 2 + 1]; }
        double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
        for (int g = 0; g < 4; g++) {
            double r0, r1, gx[4], gy[4], det, N[4];
            q4_gauss(g, r0, r1);
            q4_grad(X, r0, r1, gx, gy, det);
            q4_shape(r0, r1, N);
            double Na = N[0], ax = gx[0], ay = gy[0];
#pragma unroll
            for (int n = 1; n < 4; n++) if (n == a) { Na = N[n]; ax = gx[n]; ay = gy[n]; }
#pragma unroll
            for (int b = 0; b < 4; b++) acc[b] += (Na * N[b] * inv_dt + D * (ax * gx[b] + ay * gy[b])) * det;
        }
        const int* bm = bmap + ((size_t)e * 4 + a) * 4;
        const long long base = indptr[row];
#pragma unroll
        for (int b = 0; b < 4; b++) if (n2g[nd[b]] != -1) atomicAdd(&data[base + bm[b]], acc[b]);      // fixed columns: phi = 0, no lift
    }
}

// Y = Me/dt phi + Fe,  Fe_a = sum_g N_a C (TDN_g - lambda) J w   (driver :158-166)
__global__ void rd_rhs_kernel(int nelem, const double* __restrict__ coords, const int* __restrict__ conn, const int* __restrict__ n2g,
                              const double* __restrict__ phi, const double* __restrict__ TDN, const LsState* __restrict__ st, double inv_dt,
                              double* __restrict__ Y) {
    const double C = st->C, lambda = st->lambda;
    for (long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x; tid < (long long)nelem * 4; tid += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(tid >> 2), a = (int)(tid & 3);
        const int* nd = conn + (size_t)e * 4;
        const int row = n2g[nd[a]];
        if (row == -1) continue;
        double X[4][2], ph[4], tn[4];
#pragma unroll
        for (int n = 0; n < 4; n++) {
            X[n][0] = coords[(size_t)nd[n] * 2]; X[n][1] = coords[(size_t)nd[n] * 2 + 1];
            ph[n] = phi[nd[n]]; tn[n] = TDN[nd[n]];
        }
        double acc = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            double r0, r1, gx[4], gy[4], det, N[4];
            q4_gauss(g, r0, r1);
            q4_grad(X, r0, r1, gx, gy, det);
            q4_shape(r0, r1, N);
            double Na = N[0];
#pragma unroll
            for (int n = 1; n < 4; n++) if (n == a) Na = N[n];
            double pg = 0.0, ug = 0.0;
#pragma unroll
            for (int b = 0; b < 4; b++) { pg += N[b] * ph[b]; ug += N[b] * tn[b]; }
            acc += Na * (pg * inv_dt + C * (ug - lambda)) * det;
        }
        atomicAdd(&Y[row], acc);
    }
}

__global__ void ls_fix_phi_kernel(int nphi, const int* __restrict__ pnode, double* __restrict__ phi) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nphi; i += gridDim.x * blockDim.x) phi[pnode[i]] = 0.0;
}
// Disassembling + clamp (driver :176-180)
__global__ void ls_update_phi_kernel(int nnode, const int* __restrict__ n2g, const double* __restrict__ y, double* __restrict__ phi) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnode; i += gridDim.x * blockDim.x) {
        const int r = n2g[i];
        const double v = (r != -1) ? y[r] : phi[i];
        phi[i] = fmax(fmin(1.0, v), -1.0);
    }
}
// InterpolateElementalFromNodal + characteristic function (driver :182-190)
__global__ void ls_str_kernel(int nelem, const int* __restrict__ conn, const double* __restrict__ phi, double* __restrict__ str) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        const int* nd = conn + (size_t)e * 4;
        double v = 0.0;
#pragma unroll
        for (int n = 0; n < 4; n++) v += phi[nd[n]];
        v /= 4.0;
        str[e] = (v < 0.0) ? 0.0 : 1.0;
    }
}
__global__ void ls_advance_kernel(LsState* st) { st->t++; }

}  // namespace pf2

using namespace pf2;

struct pf2_levelset {
    pf2_ctx* ctx = nullptr;
    pf2_mesh* mesh = nullptr;
    pf2_dofmap* map_u = nullptr;
    pf2_csr* K = nullptr;
    pf2_dofmap* map_phi = nullptr;      // owned
    pf2_csr* T = nullptr;               // owned
    int nnode = 0, nelem = 0, tmax = 0, t = 0;
    double Vmax, tau, E0, Emin, nu, nvol, dt, d, p, volInit = 1.0, Etd = 0.0;
    int eq = 0;
    ElemSpec sp_k, sp_td;
    int nload = 0, nphi = 0;
    int *ld_node = nullptr, *ld_dof = nullptr, *pnode = nullptr, *count = nullptr;
    double* ld_val = nullptr;
    double *phi = nullptr, *str = nullptr, *Emod = nullptr, *xsol = nullptr, *u = nullptr, *TD = nullptr, *TDN = nullptr, *Y = nullptr, *ysol = nullptr;
    double* objective = nullptr;
    LsState* st = nullptr;
    bool vol_init_set = false;
};

static ElemSpec planestress_spec(double V) {
    ElemSpec sp;
    sp.npass = 1; sp.quad[0] = sp.quad[1] = PF2_QUAD_G4SQ; sp.wilson_taylor = 0;
    const double c = 1.0 / ((1.0 - V) * (1.0 + V));
    sp.cn[0] = sp.cn[1] = c; sp.lam[0] = sp.lam[1] = V * c; sp.mu[0] = sp.mu[1] = 0.5 * (1.0 - V) * c;
    return sp;
}

extern "C" {

int pf2_levelset_create(pf2_ctx* ctx, pf2_mesh* mesh, pf2_dofmap* map_u, pf2_csr* K, int nphi, const int* phi_fixed_nodes_host,
                        const double prm[9], int tmax, int nload, const int* load_node_host, const int* load_dof_host,
                        const double* load_val_host, pf2_levelset** out) {
    PF2_CHECK(ctx && mesh && map_u && K && prm && out && tmax > 0, "null argument");
    PF2_CHECK(mesh->dim == 2 && mesh->npe == 4 && map_u->ndof == 2, "the level-set loop is built for Q4 plane stress (sample_optimize_levelset.cpp)");
    PF2_CHECK(nphi == 0 || phi_fixed_nodes_host, "null phi boundary list");
    PF2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    pf2_levelset* L = new pf2_levelset();
    L->ctx = ctx; L->mesh = mesh; L->map_u = map_u; L->K = K; L->nnode = mesh->nnode; L->nelem = mesh->nelem; L->tmax = tmax;
    L->Vmax = prm[0]; L->tau = prm[1]; L->E0 = prm[2]; L->Emin = prm[3]; L->nu = prm[4]; L->nvol = prm[5]; L->dt = prm[6]; L->d = prm[7]; L->p = prm[8];
    L->eq = PF2_EQ_CODE(PF2_PHYS_PLANESTRESS, PF2_SHAPE_Q4, PF2_QUAD_G4SQ, 0);
    // topological-derivative material (driver :36-38, :120)
    const double nu = L->nu, E0 = L->E0;
    const double A1 = -1.5 * (1.0 - nu) * (1.0 - 14.0 * nu + 15.0 * pow(nu, 2.0)) * E0 / ((1.0 + nu) * (7.0 - 5.0 * nu) * pow(1.0 - 2.0 * nu, 2.0));
    const double A2 = 7.5 * (1.0 - nu) * E0 / ((1.0 + nu) * (7.0 - 5.0 * nu));
    const double cc = A1 / (A1 + 2.0 * A2);
    L->Etd = (A1 + 2.0 * A2) * (1.0 - pow(cc, 2.0));
    L->sp_k = planestress_spec(nu);
    L->sp_td = planestress_spec(cc);
    // phi system: one dof per node, phi = 0 on the listed nodes
    std::vector<int> pdof((size_t)std::max(nphi, 1), 0);
    std::vector<double> pval((size_t)std::max(nphi, 1), 0.0);
    int tdeg = 0;
    int rc = pf2_dofmap_create(ctx, mesh->nnode, 1, nphi, phi_fixed_nodes_host, pdof.data(), pval.data(), &tdeg, &L->map_phi);
    if (rc == PF2_OK) rc = pf2_csr_pattern(ctx, mesh, L->map_phi, &L->T);
    if (rc != PF2_OK) { delete L; return rc; }
    L->nload = nload; L->nphi = nphi;
    PF2_TRY(dev_alloc(&L->ld_node, (size_t)nload)); PF2_TRY(dev_alloc(&L->ld_dof, (size_t)nload)); PF2_TRY(dev_alloc(&L->ld_val, (size_t)nload));
    PF2_TRY(dev_alloc(&L->pnode, (size_t)nphi)); PF2_TRY(dev_alloc(&L->count, (size_t)L->nnode));
    if (nload > 0) {
        PF2_CUDA(cudaMemcpyAsync(L->ld_node, load_node_host, sizeof(int) * (size_t)nload, cudaMemcpyHostToDevice, s));
        PF2_CUDA(cudaMemcpyAsync(L->ld_dof, load_dof_host, sizeof(int) * (size_t)nload, cudaMemcpyHostToDevice, s));
        PF2_CUDA(cudaMemcpyAsync(L->ld_val, load_val_host, sizeof(double) * (size_t)nload, cudaMemcpyHostToDevice, s));
    }
    if (nphi > 0) PF2_CUDA(cudaMemcpyAsync(L->pnode, phi_fixed_nodes_host, sizeof(int) * (size_t)nphi, cudaMemcpyHostToDevice, s));
    PF2_TRY(dev_alloc(&L->phi, (size_t)L->nnode)); PF2_TRY(dev_alloc(&L->str, (size_t)L->nelem)); PF2_TRY(dev_alloc(&L->Emod, (size_t)L->nelem));
    PF2_TRY(dev_alloc(&L->xsol, (size_t)K->rows)); PF2_TRY(dev_alloc(&L->u, (size_t)L->nnode * 2)); PF2_TRY(dev_alloc(&L->TD, (size_t)L->nelem));
    PF2_TRY(dev_alloc(&L->TDN, (size_t)L->nnode)); PF2_TRY(dev_alloc(&L->Y, (size_t)L->T->rows)); PF2_TRY(dev_alloc(&L->ysol, (size_t)L->T->rows));
    PF2_TRY(dev_alloc(&L->objective, (size_t)tmax)); PF2_TRY(dev_alloc(&L->st, 1));
    PF2_CUDA(cudaMemsetAsync(L->st, 0, sizeof(LsState), s));
    PF2_CUDA(cudaMemsetAsync(L->objective, 0, sizeof(double) * (size_t)tmax, s));
    PF2_CUDA(cudaMemsetAsync(L->count, 0, sizeof(int) * (size_t)L->nnode, s));
    ls_count_kernel<<<ctx->grid_for((long long)L->nelem * 4), kThreads, 0, s>>>(L->nelem, mesh->conn, L->count);
    // T = Me/dt + tau nelem Ke does not depend on the design: assemble it once
    PF2_CUDA(cudaMemsetAsync(L->T->data, 0, sizeof(double) * (size_t)L->T->nnz, s));
    rd_matrix_kernel<<<ctx->grid_for((long long)L->nelem * 4), kThreads, 0, s>>>(L->nelem, mesh->coords, mesh->conn, L->map_phi->n2g, L->T->bmap,
                                                                                L->T->indptr, 1.0 / L->dt, L->tau * (double)L->nelem, L->T->data);
    PF2_LAUNCH_CHECK();
    ctx->launches += 2;
    L->T->sell_values_valid = false;
    // default state of the driver (:70-71): phi = 1, str = 1
    std::vector<double> ones((size_t)std::max(L->nnode, L->nelem), 1.0);
    PF2_CUDA(cudaMemcpyAsync(L->phi, ones.data(), sizeof(double) * (size_t)L->nnode, cudaMemcpyHostToDevice, s));
    PF2_CUDA(cudaMemcpyAsync(L->str, ones.data(), sizeof(double) * (size_t)L->nelem, cudaMemcpyHostToDevice, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    L->volInit = 1.0;
    *out = L;
    return PF2_OK;
}

int pf2_levelset_destroy(pf2_levelset* L) {
    if (!L) return PF2_OK;
    cudaSetDevice(L->ctx->device);
    void* bufs[] = { L->ld_node, L->ld_dof, L->ld_val, L->pnode, L->count, L->phi, L->str, L->Emod, L->xsol, L->u, L->TD, L->TDN, L->Y, L->ysol, L->objective, L->st };
    for (void* b : bufs) if (b) cudaFree(b);
    if (L->T) pf2_csr_destroy(L->T);
    if (L->map_phi) pf2_dofmap_destroy(L->map_phi);
    delete L;
    return PF2_OK;
}

// phi (nnode) and str (nelem) of the start of the run; volInit = mean(str) as the driver computes it (:72)
int pf2_levelset_set_state(pf2_levelset* L, const double* phi_host, const double* str_host) {
    PF2_CHECK(L && phi_host && str_host, "null argument");
    cudaStream_t s = L->ctx->stream;
    PF2_CUDA(cudaMemcpyAsync(L->phi, phi_host, sizeof(double) * (size_t)L->nnode, cudaMemcpyHostToDevice, s));
    PF2_CUDA(cudaMemcpyAsync(L->str, str_host, sizeof(double) * (size_t)L->nelem, cudaMemcpyHostToDevice, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    double v = 0.0;
    for (int i = 0; i < L->nelem; i++) v += str_host[i];
    L->volInit = v / (double)L->nelem;
    L->t = 0;
    PF2_CUDA(cudaMemsetAsync(L->st, 0, sizeof(LsState), s));
    return PF2_OK;
}

// stats[8] = { objective, vol, lambda, converged, cg iterations (u), cg relres (u), cg iterations (phi), t }
int pf2_levelset_iterate(pf2_levelset* L, int check_convergence, double stats[8]) {
    PF2_CHECK(L, "null handle");
    PF2_CHECK(L->t < L->tmax, "iteration budget (tmax) exhausted");
    pf2_ctx* c = L->ctx;
    cudaStream_t s = c->stream;
    PF2_CUDA(cudaSetDevice(c->device));
    const int ge = c->grid_for(L->nelem), gn = c->grid_for(L->nnode), g4 = c->grid_for((long long)L->nelem * 4);
    ls_modulus_kernel<<<ge, kThreads, 0, s>>>(L->nelem, L->str, L->E0, L->Emin, L->Emod);
    const double ap[5] = { 0.0, 0.0, L->nu, 1.0, 1.0 };
    PF2_TRY(assemble_device(L->K, L->mesh, L->map_u, L->eq, L->Emod, nullptr, ap, L->nload, L->ld_node, L->ld_dof, L->ld_val));
    int it_u = 0, it_p = 0;
    double rr_u = 0.0, rr_p = 0.0;
    int rc = solve(L->K, PF2_SOLVER_SCALINGCG, L->K->F, L->xsol, 100000, 1.0e-10, &it_u, &rr_u);
    if (rc != PF2_OK && rc != PF2_E_NOCONV) return rc;
    PF2_TRY(pf2_disassemble(L->map_u, L->xsol, L->u));
    ls_energy_kernel<<<ge, kThreads, 0, s>>>(L->nelem, L->sp_k, L->sp_td, L->Etd, L->mesh->coords, L->mesh->conn, L->u, L->str, L->Emod, L->TD, L->st,
                                            L->objective, c->red.partials, c->red.ticket);
    PF2_CUDA(cudaMemsetAsync(L->TDN, 0, sizeof(double) * (size_t)L->nnode, s));
    ls_scatter_kernel<<<g4, kThreads, 0, s>>>(L->nelem, L->mesh->conn, L->TD, L->TDN);
    ls_nodal_kernel<<<gn, kThreads, 0, s>>>(L->nnode, L->count, L->TDN, L->st, c->red.partials, c->red.ticket);
    ls_scalars_kernel<<<1, 1, 0, s>>>(L->st, L->objective, L->nnode, L->nelem, L->Vmax, L->volInit, L->nvol, L->p, L->d, check_convergence);
    PF2_LAUNCH_CHECK();
    c->launches += 5;
    LsState h;
    PF2_CUDA(cudaMemcpyAsync(&h, L->st, sizeof(LsState), cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    if (!h.converged) {
        ls_fix_phi_kernel<<<c->grid_for(std::max(L->nphi, 1)), kThreads, 0, s>>>(L->nphi, L->pnode, L->phi);
        PF2_CUDA(cudaMemsetAsync(L->Y, 0, sizeof(double) * (size_t)L->T->rows, s));
        rd_rhs_kernel<<<g4, kThreads, 0, s>>>(L->nelem, L->mesh->coords, L->mesh->conn, L->map_phi->n2g, L->phi, L->TDN, L->st, 1.0 / L->dt, L->Y);
        PF2_LAUNCH_CHECK();
        c->launches += 2;
        rc = solve(L->T, PF2_SOLVER_SCALINGCG, L->Y, L->ysol, 100000, 1.0e-10, &it_p, &rr_p);
        if (rc != PF2_OK && rc != PF2_E_NOCONV) return rc;
        ls_update_phi_kernel<<<gn, kThreads, 0, s>>>(L->nnode, L->map_phi->n2g, L->ysol, L->phi);
        ls_str_kernel<<<ge, kThreads, 0, s>>>(L->nelem, L->mesh->conn, L->phi, L->str);
        PF2_LAUNCH_CHECK();
        c->launches += 2;
    }
    ls_advance_kernel<<<1, 1, 0, s>>>(L->st);
    c->launches++;
    if (stats) {
        stats[0] = h.objective; stats[1] = h.vol; stats[2] = h.lambda; stats[3] = h.converged; stats[4] = it_u; stats[5] = rr_u; stats[6] = it_p;
        stats[7] = L->t;
    }
    L->t++;
    return PF2_OK;
}

int pf2_levelset_get(pf2_levelset* L, double* phi_host, double* str_host, double* u_host) {
    PF2_CHECK(L, "null handle");
    cudaStream_t s = L->ctx->stream;
    if (phi_host) PF2_CUDA(cudaMemcpyAsync(phi_host, L->phi, sizeof(double) * (size_t)L->nnode, cudaMemcpyDeviceToHost, s));
    if (str_host) PF2_CUDA(cudaMemcpyAsync(str_host, L->str, sizeof(double) * (size_t)L->nelem, cudaMemcpyDeviceToHost, s));
    if (u_host) PF2_CUDA(cudaMemcpyAsync(u_host, L->u, sizeof(double) * (size_t)L->nnode * 2, cudaMemcpyDeviceToHost, s));
    PF2_CUDA(cudaStreamSynchronize(s));
    return PF2_OK;
}

}  // extern "C"
