// advdiff.cu -- batched assembly of the advection-diffusion systems and the per-element call.
//
//   pf2_advdiff_assemble : the element loops of sample/advection/sample_advectiondiffusion_static.cpp:42-55 and
//                          sample_advectiondiffusion_dynamic.cpp:51-70 (element routines of Advection.h, Assembling.h:22-66)
//   advdiff_element_launch: one element matrix (pf2_element_matrix with a PF2_PHYS_ADVDIFF code)
// Same thread mapping and scatter as assemble.cu: one thread per (element, local node) holds that node's row of the two
// group matrices in registers and adds it through the precomputed map; the system is non-symmetric, which the pattern
// (full element connectivity) and the scatter do not care about.  The solve is pf2_solve with a BiCGSTAB variant.
#include "types.cuh"
#include "element_advdiff.cuh"

namespace pf2 {

template <int SHAPE>
__global__ void __launch_bounds__(128)
advdiff_assemble_kernel(int nelem, AdvSpec sp, const double* __restrict__ vel, double cm, double ck, double cf, const double* __restrict__ coords,
                        const int* __restrict__ conn, const int* __restrict__ n2g, const double* __restrict__ ufix, const int* __restrict__ bmap,
                        const long long* __restrict__ indptr, const double* __restrict__ Tn, double* __restrict__ data, double* __restrict__ F) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE;
    const long long total = (long long)nelem * NPE;
    for (long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x; tid < total; tid += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(tid / NPE), a = (int)(tid % NPE);
        const int* nd = conn + (size_t)e * NPE;
        const int row = n2g[nd[a]];
        if (row == -1) continue;
        double X[NPE][2];
#pragma unroll
        for (int n = 0; n < NPE; n++) { X[n][0] = coords[(size_t)nd[n] * 2]; X[n][1] = coords[(size_t)nd[n] * 2 + 1]; }
        AdvSpec se = sp;
        if (vel) { se.ax = vel[(size_t)e * 2]; se.ay = vel[(size_t)e * 2 + 1]; }
        double accK[NPE], accM[NPE];
        advdiff_rows<SHAPE>(X, a, se, accK, accM);
        const int* bm = bmap + ((size_t)e * NPE + a) * NPE;
        const long long base = indptr[row];
        double fa = 0.0;
#pragma unroll
        for (int b = 0; b < NPE; b++) {
            const int nb = nd[b], col = n2g[nb];
            const double ke = cm * accM[b] + ck * accK[b];
            const double Tb = (col == -1) ? ufix[nb] : (Tn ? Tn[nb] : 0.0);     // SetDirichlet keeps the prescribed value in the field
            if (col != -1) atomicAdd(&data[base + bm[b]], ke);                  // Assembling.h:30
            else fa -= ke * Tb;                                                 // Assembling.h:34
            if (Tn) fa += (cm * accM[b] - cf * accK[b]) * Tb;                   // Fe = ((M + MS)/dt - (1 - theta)(A + D + AS)) Te, Assembling.h:38
        }
        if (fa != 0.0) atomicAdd(&F[row], fa);
    }
}

template <int SHAPE>
__global__ void advdiff_element_kernel(AdvSpec sp, const double* __restrict__ xe, double* __restrict__ Ke) {
    constexpr int NPE = ShapeTraits<SHAPE>::NPE;
    const int a = threadIdx.x;
    if (a >= NPE) return;
    double X[NPE][2];
    for (int n = 0; n < NPE; n++) { X[n][0] = xe[n * 2]; X[n][1] = xe[n * 2 + 1]; }
    double accK[NPE], accM[NPE];
    advdiff_rows<SHAPE>(X, a, sp, accK, accM);
    for (int b = 0; b < NPE; b++) Ke[a * NPE + b] = accK[b] + accM[b];
}

#define PF2_DISPATCH_ADV(shape, CALL)                     \
    do {                                                  \
        if ((shape) == PF2_SHAPE_T3) { CALL(SH_T3); }     \
        else if ((shape) == PF2_SHAPE_T6) { CALL(SH_T6); } \
        else if ((shape) == PF2_SHAPE_Q4) { CALL(SH_Q4); } \
        else { CALL(SH_Q8); }                             \
    } while (0)

int advdiff_element_launch(pf2_ctx* ctx, const EqInfo& q, const double* xe_dev, double ax, double ay, double k, double* Ke_dev) {
    const AdvSpec sp = { q.quad, q.quad2, ax, ay, k };
#define CALL(S) advdiff_element_kernel<S><<<1, 32, 0, ctx->stream>>>(sp, xe_dev, Ke_dev)
    PF2_DISPATCH_ADV(q.shape, CALL);
#undef CALL
    PF2_LAUNCH_CHECK();
    ctx->launches++;
    return PF2_OK;
}

}  // namespace pf2

using namespace pf2;

extern "C" int pf2_advdiff_assemble(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq, const double* vel_dev, const double prm[6],
                                    const double* T_nodal_dev) {
    PF2_CHECK(A && mesh && map && prm, "null argument");
    EqInfo q;
    PF2_TRY(decode_eq(eq, 0.0, &q));
    PF2_CHECK(q.phys == PF2_PHYS_ADVDIFF, "not an advection-diffusion selection (PF2_PHYS_ADVDIFF)");
    PF2_CHECK(A->bmap != nullptr, "matrix was not built by pf2_csr_pattern");
    PF2_CHECK(q.npe == mesh->npe && mesh->dim == 2, "equation does not match the mesh's element type");
    PF2_CHECK(map->ndof == 1, "the advection-diffusion routines take one dof per node (the reference asserts doulist.size() == 1, Advection.h:21)");
    PF2_CHECK(A->map_nelem == mesh->nelem && A->map_npe == mesh->npe && A->map_ndof == 1, "pattern built for another mesh");
    const double cm = prm[3], ck = prm[4], cf = prm[5];
    if (A->mf_ready) { set_error("matrix-free operator is set on this matrix; advection-diffusion systems are assembled"); return PF2_E_UNSUPPORTED; }
    pf2_ctx* c = A->ctx;
    cudaStream_t s = c->stream;
    PF2_CUDA(cudaSetDevice(c->device));
    PF2_CUDA(cudaMemsetAsync(A->data, 0, sizeof(double) * (size_t)A->nnz, s));
    PF2_CUDA(cudaMemsetAsync(A->F, 0, sizeof(double) * (size_t)A->rows, s));
    const AdvSpec sp = { q.quad, q.quad2, prm[0], prm[1], prm[2] };
    const long long work = (long long)mesh->nelem * mesh->npe;
    const int grid = (int)std::min<long long>((work + 127) / 128, (long long)c->sm_count * 32);
#define CALL(S) advdiff_assemble_kernel<S><<<grid, 128, 0, s>>>(mesh->nelem, sp, vel_dev, cm, ck, cf, mesh->coords, mesh->conn, map->n2g, map->ufix, \
                                                                A->bmap, A->indptr, T_nodal_dev, A->data, A->F)
    PF2_DISPATCH_ADV(q.shape, CALL);
#undef CALL
    PF2_LAUNCH_CHECK();
    c->launches++;
    A->ilu_valid = false;
    A->sell_values_valid = false;
    return PF2_OK;
}
