// assemble.cu -- NUMERIC phase of assembly and the reaction / compliance / sensitivity passes.
//
//   pf2_assemble         = the drivers' element loop (sample_optimize_density_oc.cpp:122-129):
//                          element routine (PlaneStrain.h:21-58 | Solid.h:21-64 | HeatTransfer.h:20-43)
//                          + Assembling(K,F,u,Ke,...) (Assembling.h:47-66) + nodal loads (Assembling.h:152-158)
//   pf2_compliance_sens  = reaction pass + compliance + sensitivity pass (sample_optimize_density_oc.cpp:136-162)
//   pf2_disassemble      = Disassembling (Assembling.h:163-171)
//
// Assembly kernel: one thread per (element, local node a) computes the NDOF rows of Ke that belong to node a in
// registers (batched B^T D B, element.cuh) and scatter-adds them into the precomputed CSR pattern with fp64 RED
// atomics; positions come from indptr[row] + bmap (pattern.cu).  The threads of one element sit in adjacent lanes so
// connectivity / coordinate / density loads are broadcast, and consecutive elements give coalesced 16/32-byte loads.
// Algorithmic bytes per element: connectivity 4*npe + coordinates 8*dim*npe (shared with neighbours through L2)
// + density 8 + map 4*npe^2 + values 8*(npe*ndof)^2 atomically added.
#include "types.cuh"
#include "element.cuh"
#include "assemble_gather.cuh"

namespace pf2 {

template <int EQ>
__global__ void __launch_bounds__(128)
assemble_kernel(int nelem, const double* __restrict__ coords, const int* __restrict__ conn, const int* __restrict__ n2g,
                const double* __restrict__ ufix, const int* __restrict__ bmap, const long long* __restrict__ indptr,
                const double* __restrict__ modulus, const double* __restrict__ rho, double E0, double E1, double V, double p,
                double t, double* __restrict__ data, double* __restrict__ F) {
    constexpr int DIM = ElemTraits<EQ>::DIM, NPE = ElemTraits<EQ>::NPE, NDOF = ElemTraits<EQ>::NDOF;
    const long long total = (long long)nelem * NPE;
    for (long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x; tid < total; tid += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(tid / NPE), a = (int)(tid % NPE);
        int nd[NPE];
#pragma unroll
        for (int n = 0; n < NPE; n++) nd[n] = conn[(size_t)e * NPE + n];
        int na = nd[0];
#pragma unroll
        for (int n = 1; n < NPE; n++) if (n == a) na = nd[n];
        // rows of node a
        int rows[NDOF];
        bool any = false;
#pragma unroll
        for (int i = 0; i < NDOF; i++) { rows[i] = n2g[(size_t)na * NDOF + i]; any |= (rows[i] != -1); }
        if (!any) continue;
        double X[NPE][DIM];
#pragma unroll
        for (int n = 0; n < NPE; n++)
#pragma unroll
            for (int k = 0; k < DIM; k++) X[n][k] = coords[(size_t)nd[n] * DIM + k];
        const double E = modulus ? modulus[e] : simp_modulus(rho[e], E0, E1, p);
        double acc[NDOF][NPE * NDOF];
        if constexpr (EQ == PF2_EQ_PLANESTRAIN) planestrain_rows(reinterpret_cast<const double(&)[4][2]>(X), a, V, t, reinterpret_cast<double(&)[2][8]>(acc));
        else if constexpr (EQ == PF2_EQ_HEAT) heat_rows(reinterpret_cast<const double(&)[4][2]>(X), a, t, reinterpret_cast<double(&)[1][4]>(acc));
        else solid_rows(reinterpret_cast<const double(&)[8][3]>(X), a, V, reinterpret_cast<double(&)[3][24]>(acc));
        const int* bm = bmap + ((size_t)e * NPE + a) * NPE;
#pragma unroll
        for (int b = 0; b < NPE; b++) {
            const int off = bm[b];
            int cfree[NDOF];
            int rank = 0;
#pragma unroll
            for (int j = 0; j < NDOF; j++) {
                const int c = n2g[(size_t)nd[b] * NDOF + j];
                cfree[j] = (c != -1) ? rank++ : -1;
            }
#pragma unroll
            for (int i = 0; i < NDOF; i++) {
                if (rows[i] == -1) continue;
                const long long base = indptr[rows[i]] + off;
#pragma unroll
                for (int j = 0; j < NDOF; j++) {
                    const double v = E * acc[i][b * NDOF + j];
                    if (cfree[j] >= 0) atomicAdd(&data[base + cfree[j]], v);                  // Assembling.h:55
                    else {
                        const double uf = ufix[(size_t)nd[b] * NDOF + j];
                        if (uf != 0.0) atomicAdd(&F[rows[i]], -(v * uf));                       // Assembling.h:59
                    }
                }
            }
        }
    }
}

// the specialised 2-D element routines as row providers of the gather kernel (assemble_gather.cuh)
struct ElemPlaneStrainQ4 {
    static constexpr int DIM = 2, NPE = 4, NDOF = 2;
    double V, t;
    __device__ __forceinline__ void rows(const double (&X)[4][2], int a, double (&acc)[2][8]) const { planestrain_rows(X, a, V, t, acc); }
};
struct ElemHeatQ4 {
    static constexpr int DIM = 2, NPE = 4, NDOF = 1;
    double t;
    __device__ __forceinline__ void rows(const double (&X)[4][2], int a, double (&acc)[1][4]) const { heat_rows(X, a, t, acc); }
};

// Assembling(F, q, nodetoglobal) (Assembling.h:152-158)
__global__ void loads_kernel(int nload, int ndof, const int* __restrict__ node, const int* __restrict__ dof,
                             const double* __restrict__ val, const int* __restrict__ n2g, double* F) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nload; i += gridDim.x * blockDim.x) {
        const int r = n2g[(size_t)node[i] * ndof + dof[i]];
        if (r != -1) atomicAdd(&F[r], val[i]);
    }
}

// Disassembling (Assembling.h:163-171): u[node][dof] = result[row] on free dofs, prescribed value on fixed dofs
__global__ void disassemble_kernel(size_t n, const int* __restrict__ n2g, const double* __restrict__ ufix,
                                   const double* __restrict__ x, double* __restrict__ u) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = n2g[i];
        u[i] = (r != -1) ? x[r] : ufix[i];
    }
}

// One thread per element: strain energy w = ue^T Ke(E=1) ue evaluated from the strains at the Gauss points
// (identical to forming Ke and multiplying, without the 64 / 576 entry matrix), then
//   f      += E_e * w                                  (= sum_n u_n . r_n, driver :153, before scale0)
//   dfdrho  = -scale0*p*(E1-E0)*rho^(p-1) * w          (driver :161)
//   r      += E_e * Ke(E=1) ue  scattered to nodes     (driver :145-150, optional)
template <int EQ>
__global__ void __launch_bounds__(kThreads)
sens_kernel(int nelem, const double* __restrict__ coords, const int* __restrict__ conn, const double* __restrict__ u,
            const double* __restrict__ rho, double E0, double E1, double V, double p, double t, double scale0,
            double* __restrict__ dfdrho, double* r_nodal, double* f_out, double* partials, unsigned int* ticket, int sum_lo, int sum_hi) {
    constexpr int DIM = ElemTraits<EQ>::DIM, NPE = ElemTraits<EQ>::NPE, NDOF = ElemTraits<EQ>::NDOF;
    const Iso c(V);
    double fsum = 0.0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        int nd[NPE];
        double X[NPE][DIM], ue[NPE][NDOF], fe[NPE][NDOF];
#pragma unroll
        for (int n = 0; n < NPE; n++) {
            nd[n] = conn[(size_t)e * NPE + n];
#pragma unroll
            for (int k = 0; k < DIM; k++) X[n][k] = coords[(size_t)nd[n] * DIM + k];
#pragma unroll
            for (int k = 0; k < NDOF; k++) { ue[n][k] = u[(size_t)nd[n] * NDOF + k]; fe[n][k] = 0.0; }
        }
        double w = 0.0;
        if constexpr (EQ == PF2_EQ_SOLID) {
#pragma unroll 1
            for (int g = 0; g < 8; g++) {
                double r0, r1, r2, gx[8], gy[8], gz[8], det;
                h8_gauss(g, r0, r1, r2);
                h8_grad(reinterpret_cast<const double(&)[8][3]>(X), r0, r1, r2, gx, gy, gz, det);
                double exx = 0, eyy = 0, ezz = 0, gxy = 0, gyz = 0, gzx = 0;
#pragma unroll
                for (int n = 0; n < 8; n++) {
                    const double ux = ue[n][0], uy = ue[n][1 % NDOF], uz = ue[n][2 % NDOF];
                    exx += gx[n] * ux; eyy += gy[n] * uy; ezz += gz[n] * uz;
                    gxy += gy[n] * ux + gx[n] * uy; gyz += gz[n] * uy + gy[n] * uz; gzx += gz[n] * ux + gx[n] * uz;
                }
                const double sxx = c.cn * exx + c.lam * (eyy + ezz), syy = c.cn * eyy + c.lam * (exx + ezz), szz = c.cn * ezz + c.lam * (exx + eyy);
                const double sxy = c.mu * gxy, syz = c.mu * gyz, szx = c.mu * gzx;
                w += (sxx * exx + syy * eyy + szz * ezz + sxy * gxy + syz * gyz + szx * gzx) * det;
                if (r_nodal) {
#pragma unroll
                    for (int n = 0; n < 8; n++) {
                        fe[n][0] += (gx[n] * sxx + gy[n] * sxy + gz[n] * szx) * det;
                        fe[n][1 % NDOF] += (gy[n] * syy + gx[n] * sxy + gz[n] * syz) * det;
                        fe[n][2 % NDOF] += (gz[n] * szz + gy[n] * syz + gx[n] * szx) * det;
                    }
                }
            }
        } else {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                double r0, r1, gx[4], gy[4], det;
                q4_gauss(g, r0, r1);
                q4_grad(reinterpret_cast<const double(&)[4][2]>(X), r0, r1, gx, gy, det);
                const double wg = det * t;
                if constexpr (EQ == PF2_EQ_PLANESTRAIN) {
                    double exx = 0, eyy = 0, gxy = 0;
#pragma unroll
                    for (int n = 0; n < 4; n++) {
                        const double ux = ue[n][0], uy = ue[n][1 % NDOF];
                        exx += gx[n] * ux; eyy += gy[n] * uy; gxy += gy[n] * ux + gx[n] * uy;
                    }
                    const double sxx = c.cn * exx + c.lam * eyy, syy = c.cn * eyy + c.lam * exx, sxy = c.mu * gxy;
                    w += (sxx * exx + syy * eyy + sxy * gxy) * wg;
                    if (r_nodal) {
#pragma unroll
                        for (int n = 0; n < 4; n++) {
                            fe[n][0] += (gx[n] * sxx + gy[n] * sxy) * wg;
                            fe[n][1 % NDOF] += (gy[n] * syy + gx[n] * sxy) * wg;
                        }
                    }
                } else {
                    double qx = 0, qy = 0;
#pragma unroll
                    for (int n = 0; n < 4; n++) { qx += gx[n] * ue[n][0]; qy += gy[n] * ue[n][0]; }
                    w += (qx * qx + qy * qy) * wg;
                    if (r_nodal) {
#pragma unroll
                        for (int n = 0; n < 4; n++) fe[n][0] += (gx[n] * qx + gy[n] * qy) * wg;
                    }
                }
            }
        }
        const double rh = rho[e];
        const double E = simp_modulus(rh, E0, E1, p);
        if (e >= sum_lo && e < sum_hi) fsum += E * w;
        if (dfdrho) dfdrho[e] = -scale0 * p * (-E0 + E1) * pow(rh, p - 1.0) * w;
        if (r_nodal) {
#pragma unroll
            for (int n = 0; n < NPE; n++)
#pragma unroll
                for (int k = 0; k < NDOF; k++) atomicAdd(&r_nodal[(size_t)nd[n] * NDOF + k], E * fe[n][k]);
        }
    }
    double v[1] = { fsum };
    if (grid_sum_last<1>(v, partials, ticket) && threadIdx.x == 0) *f_out = scale0 * v[0];
}

// single element through the same device code (per-element legacy call of the reference API, parity tests)
template <int EQ>
__global__ void element_matrix_kernel(const double* __restrict__ xe, double E, double V, double t, double* __restrict__ Ke) {
    constexpr int DIM = ElemTraits<EQ>::DIM, NPE = ElemTraits<EQ>::NPE, NDOF = ElemTraits<EQ>::NDOF;
    const int a = threadIdx.x;
    if (a >= NPE) return;
    double X[NPE][DIM];
    for (int n = 0; n < NPE; n++) for (int k = 0; k < DIM; k++) X[n][k] = xe[n * DIM + k];
    double acc[NDOF][NPE * NDOF];
    if constexpr (EQ == PF2_EQ_PLANESTRAIN) planestrain_rows(reinterpret_cast<const double(&)[4][2]>(X), a, V, t, reinterpret_cast<double(&)[2][8]>(acc));
    else if constexpr (EQ == PF2_EQ_HEAT) heat_rows(reinterpret_cast<const double(&)[4][2]>(X), a, t, reinterpret_cast<double(&)[1][4]>(acc));
    else solid_rows(reinterpret_cast<const double(&)[8][3]>(X), a, V, reinterpret_cast<double(&)[3][24]>(acc));
    constexpr int M = NPE * NDOF;
    for (int i = 0; i < NDOF; i++) for (int j = 0; j < M; j++) Ke[(a * NDOF + i) * M + j] = E * acc[i][j];
}

}  // namespace pf2

using namespace pf2;

namespace pf2 {
int assemble_generic_launch(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, const EqInfo& q, const double* modulus_dev, const double* rho_dev,
                            const double params[5]);
int sens_generic_launch(pf2_mesh* mesh, const EqInfo& q, const double* u_nodal, const double* rho, const double params[6], double* f_dev,
                        double* dfdrho, double* r_nodal);
int element_generic_launch(pf2_ctx* ctx, const EqInfo& q, const double* xe_dev, double E, double t, double* Ke_dev);
int mf_update(pf2_csr* A, pf2_mesh* mesh, const double* modulus_dev, const double* rho_dev, const double params[5]);
int advdiff_element_launch(pf2_ctx* ctx, const EqInfo& q, const double* xe_dev, double ax, double ay, double k, double* Ke_dev);
int element_general_launch(pf2_ctx* ctx, const EqInfo& q, const double* xe_dev, const double D[9], double t, double* Ke_dev);

// numeric assembly with the nodal loads already on the device (the design loop keeps them resident)
int assemble_device(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq, const double* modulus_dev, const double* rho_dev,
                    const double params[5], int nload, const int* load_node_dev, const int* load_dof_dev, const double* load_val_dev) {
    EqInfo q;
    PF2_TRY(decode_eq(eq, params[2], &q));
    PF2_CHECK(A->bmap != nullptr, "matrix was not built by pf2_csr_pattern");
    PF2_CHECK(q.phys != PF2_PHYS_ADVDIFF, "advection-diffusion selections are assembled by pf2_advdiff_assemble");
    PF2_CHECK(q.phys < PF2_PHYS_PLANE_D, "selections with a caller-supplied constitutive matrix have the per-element entry point only (pf2_element_matrix_d)");
    PF2_CHECK(q.npe == mesh->npe && q.dim == mesh->dim, "equation does not match the mesh's element type");
    PF2_CHECK(q.ndof == map->ndof, "equation does not match the dof map (the reference asserts doulist.size(), PlaneStrain.h:22)");
    PF2_CHECK(A->map_nelem == mesh->nelem && A->map_npe == mesh->npe && A->map_ndof == map->ndof, "pattern built for another mesh");
    PF2_CHECK(modulus_dev || rho_dev, "need a modulus or a density field");
    pf2_ctx* c = A->ctx;
    cudaStream_t s = c->stream;
    const double E0 = params[0], E1 = params[1], V = params[2], p = params[3], t = params[4];
    // 2-D selections: row-gather kernel (every entry written once, no memset, bitwise reproducible) when the plan fits
    const bool gather3 = q.fast && q.legacy == PF2_EQ_SOLID && gather3d_usable(A, mesh);
    const bool gather = (q.dim == 2 && gather_usable(A, mesh)) || gather3;
    if (!gather) {
        PF2_CUDA(cudaMemsetAsync(A->data, 0, sizeof(double) * (size_t)A->nnz, s));
        PF2_CUDA(cudaMemsetAsync(A->F, 0, sizeof(double) * (size_t)A->rows, s));
    }
    const long long work = (long long)mesh->nelem * mesh->npe;
    const int grid = (int)std::min<long long>((work + 127) / 128, (long long)c->sm_count * 32);
#define LAUNCH(EQ) assemble_kernel<EQ><<<grid, 128, 0, s>>>(mesh->nelem, mesh->coords, mesh->conn, map->n2g, map->ufix, A->bmap, A->indptr, \
                                                          modulus_dev, rho_dev, E0, E1, V, p, t, A->data, A->F)
    if (!q.fast) PF2_TRY(assemble_generic_launch(A, mesh, map, q, modulus_dev, rho_dev, params));
    else if (gather && q.legacy == PF2_EQ_PLANESTRAIN) PF2_TRY(assemble_gather_launch(A, mesh, map, ElemPlaneStrainQ4{ V, t }, modulus_dev, rho_dev, E0, E1, p));
    else if (gather && q.legacy == PF2_EQ_HEAT) PF2_TRY(assemble_gather_launch(A, mesh, map, ElemHeatQ4{ t }, modulus_dev, rho_dev, E0, E1, p));
    else if (gather3) PF2_TRY(assemble_gather_hex8_launch(A, mesh, map, V, modulus_dev, rho_dev, E0, E1, p));
    else {
        if (q.legacy == PF2_EQ_PLANESTRAIN) LAUNCH(PF2_EQ_PLANESTRAIN);
        else if (q.legacy == PF2_EQ_SOLID) LAUNCH(PF2_EQ_SOLID);
        else LAUNCH(PF2_EQ_HEAT);
        PF2_LAUNCH_CHECK();
        c->launches++;
    }
#undef LAUNCH
    if (nload > 0) {
        loads_kernel<<<c->grid_for(nload), kThreads, 0, s>>>(nload, map->ndof, load_node_dev, load_dof_dev, load_val_dev, map->n2g, A->F);
        PF2_LAUNCH_CHECK();
        c->launches++;
    }
    A->ilu_valid = false;
    A->sell_values_valid = false;
    if (A->mf_ready) PF2_TRY(mf_update(A, mesh, modulus_dev, rho_dev, params));
    return PF2_OK;
}

// f (device scalar), dfdrho and optionally r = K u
int compliance_sens_device(pf2_mesh* mesh, int eq, const double* u_nodal, const double* rho, const double params[6], double* f_dev,
                           double* dfdrho, double* r_nodal) {
    pf2_ctx* c = mesh->ctx;
    cudaStream_t s = c->stream;
    EqInfo q;
    PF2_TRY(decode_eq(eq, params[2], &q));
    PF2_CHECK(q.npe == mesh->npe && q.dim == mesh->dim, "equation does not match the mesh");
    PF2_CHECK(q.phys != PF2_PHYS_ADVDIFF, "no compliance / sensitivity pass for the (non-symmetric) advection-diffusion operator");
    PF2_CHECK(q.phys < PF2_PHYS_PLANE_D, "selections with a caller-supplied constitutive matrix have the per-element entry point only (pf2_element_matrix_d)");
    const int ndof = q.ndof;
    if (r_nodal) PF2_CUDA(cudaMemsetAsync(r_nodal, 0, sizeof(double) * (size_t)mesh->nnode * ndof, s));
    if (!q.fast) return sens_generic_launch(mesh, q, u_nodal, rho, params, f_dev, dfdrho, r_nodal);
    eq = q.legacy;
    const int grid = c->grid_for(mesh->nelem);
#define LAUNCH(EQ) sens_kernel<EQ><<<grid, kThreads, 0, s>>>(mesh->nelem, mesh->coords, mesh->conn, u_nodal, rho, params[0], params[1], \
                                                           params[2], params[3], params[4], params[5], dfdrho, r_nodal, f_dev, c->red.partials, c->red.ticket, mesh->own_elem_lo, mesh->own_elem_hi)
    if (eq == PF2_EQ_PLANESTRAIN) LAUNCH(PF2_EQ_PLANESTRAIN);
    else if (eq == PF2_EQ_SOLID) LAUNCH(PF2_EQ_SOLID);
    else LAUNCH(PF2_EQ_HEAT);
#undef LAUNCH
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}
}  // namespace pf2

extern "C" {

int pf2_assemble(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq, const double* modulus_dev, const double* rho_dev,
                 const double params[5], int nload, const int* load_node_host, const int* load_dof_host, const double* load_val_host) {
    PF2_CHECK(A && mesh && map && params, "null argument");
    pf2_ctx* c = A->ctx;
    cudaStream_t s = c->stream;
    PF2_CUDA(cudaSetDevice(c->device));
    int *dn = nullptr, *dd = nullptr;
    double* dv = nullptr;
    if (nload > 0) {
        PF2_TRY(dev_alloc(&dn, (size_t)nload)); PF2_TRY(dev_alloc(&dd, (size_t)nload)); PF2_TRY(dev_alloc(&dv, (size_t)nload));
        PF2_CUDA(cudaMemcpyAsync(dn, load_node_host, sizeof(int) * (size_t)nload, cudaMemcpyHostToDevice, s));
        PF2_CUDA(cudaMemcpyAsync(dd, load_dof_host, sizeof(int) * (size_t)nload, cudaMemcpyHostToDevice, s));
        PF2_CUDA(cudaMemcpyAsync(dv, load_val_host, sizeof(double) * (size_t)nload, cudaMemcpyHostToDevice, s));
    }
    int rc = assemble_device(A, mesh, map, eq, modulus_dev, rho_dev, params, nload, dn, dd, dv);
    if (nload > 0) {
        cudaStreamSynchronize(s);
        cudaFree(dn); cudaFree(dd); cudaFree(dv);
    }
    return rc;
}

int pf2_element_matrix(pf2_ctx* ctx, int eq, const double* xe_host, double E, double V, double t, double* Ke_host) {
    PF2_CHECK(ctx && xe_host && Ke_host, "bad arguments");
    EqInfo q;
    PF2_TRY(decode_eq(eq, V, &q));
    PF2_CHECK(q.phys < PF2_PHYS_PLANE_D, "this selection takes a constitutive matrix: pf2_element_matrix_d");
    const int npe = q.npe, dim = q.dim, m = npe * q.ndof;
    if (!ctx->elem_scratch) PF2_TRY(dev_alloc(&ctx->elem_scratch, (size_t)64 + 3600));      // hex20: 60 coordinates, 60 x 60 entries
    double *xe = ctx->elem_scratch, *Ke = ctx->elem_scratch + 64;
    PF2_CUDA(cudaMemcpyAsync(xe, xe_host, sizeof(double) * npe * dim, cudaMemcpyHostToDevice, ctx->stream));
    if (q.phys == PF2_PHYS_ADVDIFF) PF2_TRY(advdiff_element_launch(ctx, q, xe, E, V, t, Ke));      // (E, V, t) = (ax, ay, k)
    else if (!q.fast) PF2_TRY(element_generic_launch(ctx, q, xe, E, t, Ke));
    else {
        if (q.legacy == PF2_EQ_PLANESTRAIN) element_matrix_kernel<PF2_EQ_PLANESTRAIN><<<1, 32, 0, ctx->stream>>>(xe, E, V, t, Ke);
        else if (q.legacy == PF2_EQ_SOLID) element_matrix_kernel<PF2_EQ_SOLID><<<1, 32, 0, ctx->stream>>>(xe, E, V, t, Ke);
        else element_matrix_kernel<PF2_EQ_HEAT><<<1, 32, 0, ctx->stream>>>(xe, E, V, t, Ke);
        PF2_LAUNCH_CHECK();
        ctx->launches++;
    }
    PF2_CUDA(cudaMemcpyAsync(Ke_host, Ke, sizeof(double) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    return PF2_OK;
}

int pf2_element_matrix_d(pf2_ctx* ctx, int eq, const double* xe_host, const double D_host[9], double t, double* Ke_host) {
    PF2_CHECK(ctx && xe_host && D_host && Ke_host, "bad arguments");
    EqInfo q;
    PF2_TRY(decode_eq(eq, 0.0, &q));
    PF2_CHECK(q.phys >= PF2_PHYS_PLANE_D, "not a PF2_PHYS_PLANE_D* selection");
    const int npe = q.npe, m = npe * 2;
    if (!ctx->elem_scratch) PF2_TRY(dev_alloc(&ctx->elem_scratch, (size_t)64 + 3600));
    double *xe = ctx->elem_scratch, *Ke = ctx->elem_scratch + 64;
    PF2_CUDA(cudaMemcpyAsync(xe, xe_host, sizeof(double) * npe * 2, cudaMemcpyHostToDevice, ctx->stream));
    PF2_TRY(element_general_launch(ctx, q, xe, D_host, t, Ke));
    PF2_CUDA(cudaMemcpyAsync(Ke_host, Ke, sizeof(double) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    PF2_CUDA(cudaStreamSynchronize(ctx->stream));
    return PF2_OK;
}

int pf2_disassemble(pf2_dofmap* map, const double* x_dev, double* u_nodal_dev) {
    pf2_ctx* c = map->ctx;
    const size_t n = (size_t)map->nnode * map->ndof;
    disassemble_kernel<<<c->grid_for((long long)n), kThreads, 0, c->stream>>>(n, map->n2g, map->ufix, x_dev, u_nodal_dev);
    PF2_LAUNCH_CHECK();
    c->launches++;
    return PF2_OK;
}

int pf2_compliance_sens(pf2_mesh* mesh, int eq, const double* u_nodal_dev, const double* rho_dev, const double params[6],
                        double* f_out, double* dfdrho_dev, double* r_nodal_dev) {
    PF2_CHECK(mesh && u_nodal_dev && rho_dev && params, "null argument");
    pf2_ctx* c = mesh->ctx;
    PF2_TRY(compliance_sens_device(mesh, eq, u_nodal_dev, rho_dev, params, c->scalars, dfdrho_dev, r_nodal_dev));
    if (f_out) {
        PF2_CUDA(cudaMemcpyAsync(c->h_scalars, c->scalars, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        PF2_CUDA(cudaStreamSynchronize(c->stream));
        *f_out = c->h_scalars[0];
    }
    return PF2_OK;
}

}  // extern "C"
