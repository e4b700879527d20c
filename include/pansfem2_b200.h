/* pansfem2_b200.h -- C ABI of libpansfem2_b200.so, the B200 (sm_100a) implementation of PANSFEM2's SIMP
 * topology-optimisation hot path.
 *
 * PANSFEM2 has no FFI of its own: its public surface is a header-only C++ template library.  This ABI is what the
 * header mirror under pansfem2_b200/src/ (same relative paths, names and signatures as the reference's src/) binds
 * for T = double; each entry point cites the reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes only; every handle is opaque; no exceptions cross the boundary
 *   - every function returns 0 on success or a PF2_E_* code; pf2_last_error() gives the message (thread local)
 *   - "_host" arguments are host pointers, "_dev" arguments are device pointers obtained from pf2_malloc
 *   - all work is enqueued on the context's stream; calls that return host results synchronise that stream
 *   - indices are int32 (as the reference, CSR.h:72-73); CSR row pointers are int64 on the device side so that
 *     config 5 (nnz = 3.45e9) is representable
 *   - there is NO CPU fallback: without a CUDA device pf2_ctx_create fails with PF2_E_NODEVICE
 */
#ifndef PANSFEM2_B200_H
#define PANSFEM2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF2_OK 0
#define PF2_E_INVALID 1    /* bad argument (the reference asserts, e.g. PlaneStrain.h:22) */
#define PF2_E_CUDA 2       /* CUDA runtime error */
#define PF2_E_NODEVICE 3   /* no usable CUDA device */
#define PF2_E_NOCONV 4     /* solver hit itrmax (the reference prints "Convergence:faild", CG.h:152,451) */
#define PF2_E_UNSUPPORTED 5

/* equation / element selection == the reference's template arguments
 *   PF2_EQ_PLANESTRAIN : PlaneStrainStiffness<double, ShapeFunction4Square, Gauss4Square>        (PlaneStrain.h:21)
 *   PF2_EQ_SOLID       : SolidLinearIsotropicElastic<double, ShapeFunction8Cubic, Gauss8Cubic>   (Solid.h:21)
 *   PF2_EQ_HEAT        : HeatTransfer<double, ShapeFunction4Square, Gauss4Square>                (HeatTransfer.h:20)
 * These three are the selections of the topology-optimisation drivers and run on specialised kernels. */
enum { PF2_EQ_PLANESTRAIN = 0, PF2_EQ_SOLID = 1, PF2_EQ_HEAT = 2 };
/* Any other <Equation, ShapeFunction, Integration> selection of the reference is an `eq` CODE built with PF2_EQ_CODE:
 *   physics : the three above, PlaneStressStiffness (PlaneStress.h:21), PlaneStrainStiffnessSRI (PlaneStrain.h:63;
 *             quad = ICD, the deviatoric rule, quad2 = ICV, the volumetric rule), and the scalar consistent mass matrix
 *             E * N N^T * t of ReactionDiffusionConsistentMass (ReactionDiffusion.h:21; E = 1, t = 1) and HeatCapacity
 *             (HeatTransfer.h:48; E = rho*c).  ReactionDiffusionStiffness (ReactionDiffusion.h:82) is PF2_PHYS_HEAT with E = D, t = 1.
 *             PlaneStrainStiffnessBbar (PlaneStrain.h:129; quad = ICD, quad2 = ICV like SRI) and the 2-dof consistent mass
 *             E * N^T N * t of PlaneStrainMass / PlaneStressMass (PlaneStrain.h:386, PlaneStress.h:63; E = rho), and
 *             PlaneStrainStiffnessWilsonTaylor (PlaneStrain.h:189; quadrilaterals, Gauss4Square or Gauss9Square).
 *   shape   : ShapeFunction3Triangle / 6Triangle / 4Square / 8Square / 4Tetrahedron / 8Cubic / 20Cubic (ShapeFunction.h)
 *   quad    : Gauss1Triangle / 3Triangle / 1Square / 4Square / 9Square / 1Tetrahedron / 8Cubic / 27Cubic (GaussIntegration.h)
 * 0 in a field means "the default of that physics" (Q4 + Gauss4Square, hex8 + Gauss8Cubic; SRI: Gauss4Square / Gauss1Square),
 * so the legacy values 0, 1, 2 are themselves valid codes.  The rule must belong to the shape's reference domain
 * (triangle, square, tetrahedron, cube), otherwise PF2_E_INVALID. */
enum { PF2_PHYS_PLANESTRAIN = 0, PF2_PHYS_SOLID = 1, PF2_PHYS_HEAT = 2, PF2_PHYS_PLANESTRESS = 3, PF2_PHYS_PLANESTRAIN_SRI = 4, PF2_PHYS_MASS = 5,
       PF2_PHYS_PLANESTRAIN_BBAR = 6, PF2_PHYS_MASS2 = 7, PF2_PHYS_PLANESTRAIN_WT = 8, PF2_PHYS_ADVDIFF = 9,
       PF2_PHYS_PLANE_D = 10, PF2_PHYS_PLANE_D_BBAR = 11, PF2_PHYS_PLANE_D_WT = 12 };
/* PF2_PHYS_ADVDIFF: the scalar advection-diffusion routines of Advection.h on any 2-D shape / rule; the quad2 field of the code is
 * the MASK of the routines to sum: Advection (Advection.h:19), Diffusion (:135), AdvectionSUPG (:47), AdvectionShockCapturing (:91),
 * Mass (:161), MassSUPG (:188).  pf2_element_matrix takes (E, V, t) = (ax, ay, k); systems are assembled by pf2_advdiff_assemble. */
/* PF2_PHYS_PLANE_D / _D_BBAR / _D_WT: PlaneStiffness, PlaneStiffnessBbar (quad = ICD, quad2 = ICV) and PlaneStiffnessWilsonTaylor of
 * Homogenization.h:141-280 - plane elements with a caller-supplied 3 x 3 constitutive matrix; one element through pf2_element_matrix_d. */
enum { PF2_ADV_ADVECTION = 1, PF2_ADV_DIFFUSION = 2, PF2_ADV_SUPG = 4, PF2_ADV_SHOCK = 8, PF2_ADV_MASS = 16, PF2_ADV_MASS_SUPG = 32 };
enum { PF2_SHAPE_DEFAULT = 0, PF2_SHAPE_T3 = 1, PF2_SHAPE_T6 = 2, PF2_SHAPE_Q4 = 3, PF2_SHAPE_Q8 = 4, PF2_SHAPE_TET4 = 5,
       PF2_SHAPE_HEX8 = 6, PF2_SHAPE_HEX20 = 7,
       PF2_SHAPE_LINE2 = 8, PF2_SHAPE_LINE3 = 9 };      /* ShapeFunction2Line / 3Line (ShapeFunction.h:20-84): edges carrying surface loads */
enum { PF2_QUAD_DEFAULT = 0, PF2_QUAD_G1TRI = 1, PF2_QUAD_G3TRI = 2, PF2_QUAD_G1SQ = 3, PF2_QUAD_G4SQ = 4, PF2_QUAD_G9SQ = 5,
       PF2_QUAD_G1TET = 6, PF2_QUAD_G8CUBE = 7, PF2_QUAD_G27CUBE = 8,
       PF2_QUAD_G1LINE = 9, PF2_QUAD_G2LINE = 10 };     /* Gauss1Line / Gauss2Line (GaussIntegration.h:18-60) */
#define PF2_EQ_CODE(phys, shape, quad, quad2) ((phys) | ((shape) << 8) | ((quad) << 16) | ((quad2) << 24))
/* solver selection: CG (CG.h:124), ScalingCG (CG.h:420), ILU0CG (CG.h:320); for non-symmetric systems BiCGSTAB (CG.h:159),
 * BiCGSTAB2 (CG.h:199), ScalingBiCGSTAB (CG.h:458), ILU0BiCGSTAB (CG.h:357) */
enum { PF2_SOLVER_CG = 0, PF2_SOLVER_SCALINGCG = 1, PF2_SOLVER_ILU0CG = 2, PF2_SOLVER_BICGSTAB = 3, PF2_SOLVER_BICGSTAB2 = 4,
       PF2_SOLVER_SCALINGBICGSTAB = 5, PF2_SOLVER_ILU0BICGSTAB = 6 };
/* DensityFilter (DensityFilter.h:45-71), HeavisideFilter (HeavisideFilter.h:61-99),
 * SensitivityFilter / SensitivityFilter2 (SensitivityFilter.h:44-55, 88-99; sensitivities only) */
enum { PF2_FILTER_DENSITY = 0, PF2_FILTER_HEAVISIDE = 1, PF2_FILTER_SENS_SIGMUND = 2, PF2_FILTER_SENS_BORRVALL = 3 };
/* OC (OC.h:78), MMA (MMA.h:117), CONLIN (CONLIN.h:89) */
enum { PF2_OPT_OC = 0, PF2_OPT_MMA = 1, PF2_OPT_CONLIN = 2 };

/* decode an eq code: spatial dimension, nodes per element, dofs per node (any pointer may be NULL) */
int pf2_eq_describe(int eq, int* dim, int* npe, int* ndof);

typedef struct pf2_ctx pf2_ctx;
typedef struct pf2_mesh pf2_mesh;
typedef struct pf2_dofmap pf2_dofmap;
typedef struct pf2_csr pf2_csr;
typedef struct pf2_filter pf2_filter;
typedef struct pf2_oc pf2_oc;
typedef struct pf2_mma pf2_mma;
typedef struct pf2_simp pf2_simp;
typedef struct pf2_levelset pf2_levelset;

const char* pf2_last_error(void);
const char* pf2_version(void);

/* ---- context, memory ------------------------------------------------------------------------------------ */
/* stream: a cudaStream_t to enqueue on (e.g. torch's current stream) or NULL to create a private one. */
int pf2_ctx_create(int device, void* stream, pf2_ctx** out);
int pf2_ctx_destroy(pf2_ctx* ctx);
int pf2_ctx_sync(pf2_ctx* ctx);
int pf2_ctx_device_info(pf2_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
/* number of kernels this library has launched on the context since creation (bench.py's gpu_launches) */
int pf2_ctx_launch_count(pf2_ctx* ctx, long long* out);
int pf2_malloc(pf2_ctx* ctx, size_t bytes, void** dev_out);
int pf2_free(pf2_ctx* ctx, void* dev);
int pf2_memcpy_h2d(pf2_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int pf2_memcpy_d2h(pf2_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int pf2_memset(pf2_ctx* ctx, void* dst_dev, int value, size_t bytes);
/* pinned host staging (cudaHostAlloc) for the end-to-end path */
int pf2_host_alloc(size_t bytes, void** host_out);
int pf2_host_free(void* host);
/* device-side stopwatch on the context's stream (CUDA events) */
int pf2_timer_start(pf2_ctx* ctx);
int pf2_timer_stop(pf2_ctx* ctx, double* milliseconds);
/* write `bytes` of scratch to evict L2 between timed repetitions */
int pf2_flush_l2(pf2_ctx* ctx);

/* ---- mesh, boundary conditions, numbering ------------------------------------------------------------------ */
/* coords: nnode*dim doubles (std::vector<Vector<T>> flattened), conn: nelem*npe node ids
 * (std::vector<std::vector<int>> flattened).  dim/npe: 2/4 (Q4) or 3/8 (hex8). */
int pf2_mesh_create(pf2_ctx* ctx, int dim, int nnode, const double* coords_host, int npe, int nelem,
                    const int* conn_host, pf2_mesh** out);
/* a second element list over the nodes of `base` -- the edges that carry a surface load, a sub-region with a body force (the drivers'
 * `edges` lists, sample_planestrain.cpp:23,52); shares the node coordinates on the device: destroy it before `base` */
int pf2_mesh_create_on_nodes(pf2_mesh* base, int npe, int nelem, const int* conn_host, pf2_mesh** out);
int pf2_mesh_destroy(pf2_mesh* mesh);
/* SetDirichlet (BoundaryCondition.h:20-25) + Renumbering (Assembling.h:175-186): fixed dofs get -1, free dofs are
 * numbered node-major / dof-minor.  *kdegree_out = KDEGREE.  nfixed = 0 reproduces RemoveBoundaryConditions
 * (BoundaryCondition.h:66-72) followed by Renumbering. */
int pf2_dofmap_create(pf2_ctx* ctx, int nnode, int ndof, int nfixed, const int* fix_node_host, const int* fix_dof_host,
                      const double* fix_val_host, int* kdegree_out, pf2_dofmap** out);
int pf2_dofmap_destroy(pf2_dofmap* map);
int pf2_dofmap_get(pf2_dofmap* map, int* nodetoglobal_host /* nnode*ndof */);

/* ---- sparse matrix: LILCSR<T> build then CSR<T> (LILCSR.h:92-114, CSR.h:93-105) --------------------------- */
/* Symbolic phase, once per mesh + boundary conditions: the full element-connectivity pattern (the reference
 * inserts explicit zeros, Assembling.h:55), sorted columns, plus the precomputed scatter map. */
int pf2_csr_pattern(pf2_ctx* ctx, pf2_mesh* mesh, pf2_dofmap* map, pf2_csr** out);
/* CSR<T> from caller arrays (the per-element legacy path assembles on the host through LILCSR<T>). */
int pf2_csr_upload(pf2_ctx* ctx, int rows, const int* indptr_host, const int* indices_host, const double* data_host,
                   pf2_csr** out);
int pf2_csr_destroy(pf2_csr* A);
int pf2_csr_info(pf2_csr* A, int* rows, long long* nnz);
/* indptr as int64 (rows+1), indices int32 (nnz), data (nnz), F (rows): any may be NULL */
int pf2_csr_download(pf2_csr* A, long long* indptr_host, int* indices_host, double* data_host, double* F_host);
int pf2_csr_set_values(pf2_csr* A, const double* data_host);
/* device views for composing with other calls: right-hand side F (rows) and values (nnz) */
int pf2_csr_device_F(pf2_csr* A, double** F_dev);
int pf2_csr_device_data(pf2_csr* A, double** data_dev);

/* Numeric phase = the element loop of the drivers (sample_optimize_density_oc.cpp:122-129):
 *   Ke = element routine(E_i) ; Assembling(K,F,u,Ke,...) (Assembling.h:47-66) ; Assembling(F,q,...) (Assembling.h:152)
 * modulus_dev: per-element modulus E_i (nelem) or NULL to derive it from rho_dev by SIMP interpolation
 *   E_i = E1*rho^p + E0*(1-rho^p)  (driver :123).  params = {E0, E1, poisson, p, thickness}. */
int pf2_assemble(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq, const double* modulus_dev, const double* rho_dev,
                 const double params[5], int nload, const int* load_node_host, const int* load_dof_host,
                 const double* load_val_host);
/* The element loops of sample/advection/sample_advectiondiffusion_static.cpp:42-55 and ..._dynamic.cpp:51-70 as one launch:
 *   K  = cm*(M + MS) + ck*(A + D + AS + SC)                         over the routines selected in eq (PF2_PHYS_ADVDIFF),
 *   F  = (cm*(M + MS) - cf*(A + D + AS + SC)) * T_e - K * T_fixed    (Assembling.h:22-43; the first part only when T_nodal_dev is given)
 * static sample: cm = 0, ck = 1, cf = 0, T_nodal_dev = NULL; dynamic sample: cm = 1/dt, ck = theta, cf = 1 - theta.
 * prm = {ax, ay, k, cm, ck, cf}; vel_dev: per-element velocity (nelem*2, the dynamic sample's rotating field) or NULL for the uniform
 * (ax, ay).  T_nodal_dev: the nodal field (nnode); fixed nodes are read from the dof map's Dirichlet values.  One dof per node;
 * the matrix is non-symmetric: solve with PF2_SOLVER_BICGSTAB and friends, then pf2_disassemble. */
int pf2_advdiff_assemble(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq, const double* vel_dev, const double prm[6],
                         const double* T_nodal_dev);
/* one element matrix, host in / host out: the reference's per-element call kept for parity
 * (PlaneStrain.h:21-58, Solid.h:21-64, HeatTransfer.h:20-43).  xe: npe*dim, Ke_out: (npe*ndof)^2 row-major. */
int pf2_element_matrix(pf2_ctx* ctx, int eq, const double* xe_host, double E, double V, double t, double* Ke_host);
/* the same for the PF2_PHYS_PLANE_D* selections (Homogenization.h:141-280): D_host = the 3 x 3 constitutive matrix, row-major */
int pf2_element_matrix_d(pf2_ctx* ctx, int eq, const double* xe_host, const double D_host[9], double t, double* Ke_host);

/* ---- CSR<T>::operator* (CSR.h:109-122) ------------------------------------------------------------------------ */
int pf2_spmv(pf2_csr* A, const double* x_dev, double* y_dev);
int pf2_spmv_host(pf2_csr* A, const double* x_host, double* y_host);
/* kernel selection override for tests / tuning: 0 = auto; 1-5 vector, 11-15 shared-memory stream, 21-26 TMA pipeline */
int pf2_spmv_set_variant(pf2_csr* A, int variant);
int pf2_spmv_set_tma_tuning(pf2_csr* A, int stages, int ctas_per_sm);
/* micro-benchmark hook: run SpMV `reps` times with kernel variant `variant` (0 = auto), device-timed */
int pf2_spmv_bench(pf2_csr* A, int variant, int reps, int flush_l2, double* ms_per_spmv);

/* Opt-in matrix-free application of K on a UNIFORM structured mesh (SquareMesh.h numbering / x-major hex lattice, all elements
 * congruent; Q4 or hex8 stiffness selections): y = sum_e E_e Ke0 p_e with one shared unit-modulus element matrix, instead of
 * streaming the CSR.  Same product up to summation order; assembly, preconditioner and the Krylov recurrences are unchanged.
 * Call after pf2_csr_pattern; takes effect from the next pf2_assemble (SpMV variant 41).  PF2_E_UNSUPPORTED when the mesh does
 * not qualify (the matrix keeps its CSR kernels).  pf2_spmv_set_variant(A, 0) returns to the CSR kernels. */
int pf2_csr_matrix_free(pf2_csr* A, pf2_mesh* mesh, pf2_dofmap* map, int eq);

/* ---- CG / ScalingCG / ILU0CG (CG.h:124-154, 420-453, 320-352); ILU0 / PreILU0 (CG.h:258-315) ---------------- */
/* x0 = 0; stop when ||r||_2 < eps*||b||_2 on the recursive residual.  Returns PF2_E_NOCONV at itrmax (x holds
 * the last iterate, as the reference returns it). */
int pf2_solve(pf2_csr* A, int solver, const double* b_dev, double* x_dev, int itrmax, double eps, int* iters_out,
              double* relres_out);
int pf2_solve_host(pf2_csr* A, int solver, const double* b_host, double* x_host, int itrmax, double eps,
                   int* iters_out, double* relres_out);
/* The same solve from the initial guess held in x_dev (overload sanctioned by SURVEY.md section 7: the reference always starts
 * from x0 = 0, CG.h:423).  Recurrences and stopping rule ||r|| < eps*||b|| are unchanged; r0 = b - A*x0 costs one extra product.
 * On a partitioned matrix the ghost entries of x0 must be valid.  ILU0CG and the BiCGSTAB family ignore the guess. */
int pf2_solve_x0(pf2_csr* A, int solver, const double* b_dev, double* x_dev, int itrmax, double eps, int* iters_out,
                 double* relres_out);
/* How CG / ScalingCG iterate (CG.h:430-449): 1 = one persistent cooperative kernel per solve (grid barriers carry the dot
 * products; default where the SELL-32 mirror applies), 0 = three kernels per iteration, -1 = environment (PF2_PCG, default 1). */
int pf2_csr_set_pcg_mode(pf2_csr* A, int mode);
/* Recurrences of CG / ScalingCG on a ROW-PARTITIONED matrix with the peer-memory backend: 0 = the reference's (CG.h:430-449: two
 * cross-GPU sums, three kernels per iteration; default), 1 = single-reduction (Chronopoulos-Gear: p = u + beta p, s = w + beta s,
 * alpha = gamma / (delta - beta gamma / alpha_old); ONE sum, two kernels per iteration).  Same iterates in exact arithmetic and the
 * same stopping rule; round-off differs, so results agree to the solver tolerance rather than bitwise.  -1 = environment
 * (PF2_CG_SINGLE_REDUCTION=1).  Ignored on one GPU, with the NCCL backend and by the other solvers. */
int pf2_csr_set_cg_variant(pf2_csr* A, int variant);
/* per-kernel device times of the Krylov loop.  Three-kernel loop: CUDA events around one iteration per chunk; persistent kernel:
 * its in-kernel %globaltimer stamps per phase (barriers included), samples = iterations.
 * out = {spmv+dot ms, update ms, p-update ms, samples, total iterations, SpMV variant, rows, nnz} */
int pf2_csr_solver_stats(pf2_csr* A, double out[8]);
int pf2_csr_solver_stats_reset(pf2_csr* A);
/* persistent PCG kernel since the last reset: out = {CUDA-event ms of the kernel launches (whole solves), iterations, solves,
 * CTAs of the last launch, product / update / p-update phase ms per iteration, stored SELL entries, and the part of each of the
 * three phases CTA 0 spent inside the grid exchange (tail of the grid + exchange latency), solves that ran the single-reduction
 * recurrences (pf2_csr_set_cg_variant)} */
int pf2_csr_pcg_stats(pf2_csr* A, double out[12]);
/* diagnostics: %globaltimer of every CTA at the start / end of its share of the three phases in iteration 5 of the last persistent
 * solve: out_host[6][2048] (product start, end, update start, end, p-update start, end) */
int pf2_csr_pcg_debug(pf2_csr* A, unsigned long long* out_host);
/* ILU(0) factors of A (unit-L strictly lower + U with diagonal in A's pattern), cached on A until values change */
int pf2_ilu0_factor(pf2_csr* A);
int pf2_ilu0_download(pf2_csr* A, double* data_host);
int pf2_ilu0_solve_host(pf2_csr* A, const double* b_host, double* x_host);   /* PreILU0 with A's cached factors */
/* PreILU0(M, b) (CG.h:289-315) where the VALUES of M are the factors (what the reference's ILU0 returns) */
int pf2_preilu0_host(pf2_csr* M, const double* b_host, double* x_host);

/* ---- load vectors (PlaneStrainSurfaceForce / PlaneStrainBodyForce PlaneStrain.h:421,503; PlaneStressSurfaceForce / BodyForce
 *      PlaneStress.h:98,134; HeatTransferSurfaceFlux HeatTransfer.h:76) for a whole batch of elements, assembled on the device ------ */
/* The reference evaluates the caller's force functor at x_g = X_e^T N(r_g) (PlaneStrain.h:441,522).  A functor cannot cross the ABI, so
 * the batched form is two calls: pf2_integration_points hands out every x_g (xg_dev[nelem][ngauss][2]), the caller evaluates its
 * force density there in one go, pf2_load_vector integrates and assembles.  mesh: a pf2_mesh over the SAME node numbering whose
 * elements are the loaded edges (PF2_SHAPE_LINE2 / _LINE3 with PF2_QUAD_G1LINE / _G2LINE) or areas (T3, T6, Q4, Q8 with their rules). */
int pf2_integration_points(pf2_mesh* mesh, int shape, int quad, double* xg_dev);
/* F[row(n, i)] += sum_g N_n(r_g) f_i(x_g) m_g t w_g   (Assembling(F, Fe, ...) Assembling.h:132-147), i < ndof of the dof map;
 * m_g = |dX/dr| on edges (w_g = Weights[g][0]), det(dX/dr) on areas (w_g = Weights[g][0] * Weights[g][1]).
 * f_gauss_dev[nelem][ngauss][ndof]: the force density at the integration points, or NULL: the constant vector f_const[ndof]. */
int pf2_load_vector(pf2_mesh* mesh, pf2_dofmap* map, int shape, int quad, const double* f_const, const double* f_gauss_dev, double t,
                    double* F_dev);

/* Disassembling (Assembling.h:163-171): free dofs from the solution, fixed dofs keep their Dirichlet value */
int pf2_disassemble(pf2_dofmap* map, const double* x_dev, double* u_nodal_dev);

/* ---- filters -------------------------------------------------------------------------------------------------- */
/* neighbors / w as the reference's ragged lists flattened to CSR form (rowptr has n+1 entries) */
int pf2_filter_create(pf2_ctx* ctx, int kind, int n, const long long* rowptr_host, const int* nbr_host,
                      const double* w_host, pf2_filter** out);
int pf2_filter_destroy(pf2_filter* f);
int pf2_filter_set_beta(pf2_filter* f, double beta);                       /* HeavisideFilter::UpdateBeta */
int pf2_filter_apply(pf2_filter* f, const double* s_dev, double* rho_dev); /* GetFilteredVariables */
int pf2_filter_sens(pf2_filter* f, const double* s_dev, const double* dfdrho_dev, double* dfds_dev); /* GetFilteredSensitivitis */
int pf2_filter_apply_host(pf2_filter* f, const double* s_host, double* rho_host);
int pf2_filter_sens_host(pf2_filter* f, const double* s_host, const double* dfdrho_host, double* dfds_host);

/* ---- reaction / compliance / sensitivity passes (sample_optimize_density_oc.cpp:136-162) -------------------- */
/* f = scale0 * u^T K_full(rho) u ; dfdrho_i = -scale0*p*(E1-E0)*rho_i^(p-1) * ue^T Ke(E=1) ue ;
 * r_nodal_dev (optional) = K_full(rho) u.  params = {E0, E1, poisson, p, thickness, scale0}. */
int pf2_compliance_sens(pf2_mesh* mesh, int eq, const double* u_nodal_dev, const double* rho_dev, const double params[6],
                        double* f_out, double* dfdrho_dev, double* r_nodal_dev);

/* ---- OC (OC.h:46-107) ----------------------------------------------------------------------------------------- */
int pf2_oc_create(pf2_ctx* ctx, int n, double iota, double lambdamin, double lambdamax, double lambdaeps,
                  double movelimit, pf2_oc** out);
int pf2_oc_destroy(pf2_oc* oc);
int pf2_oc_is_convergence(pf2_oc* oc, double f, int* converged);           /* OC::IsConvergence OC.h:68-73 */
/* UpdateVariables with the drivers' constraint functor g(x) = scale1*sum(filter(x))/(weightlimit*n) - scale1
 * (sample_optimize_density_oc.cpp:198-207) evaluated on the device.  x_dev updated in place. */
int pf2_oc_update(pf2_oc* oc, pf2_filter* filter, double weightlimit, double scale1, double* x_dev, double f,
                  const double* dfdx_dev, const double* dgdx_dev, int* steps_out, double* lambda_out);
/* one OC candidate x+(lambda) (OC.h:85-92), host in / host out: lets a caller-supplied constraint functor drive the
 * bisection on the host exactly as OC<T>::UpdateVariables<F> does */
int pf2_oc_candidate_host(pf2_oc* oc, const double* x_host, const double* dfdx_host, const double* dgdx_host, double lambda,
                          double* xnew_host);
/* bookkeeping of UpdateVariables (OC.h:104-105) for callers that ran the bisection themselves */
int pf2_oc_commit(pf2_oc* oc, double f);

/* ---- MMA (MMA.h:64-509) --------------------------------------------------------------------------------------- */
int pf2_mma_create(pf2_ctx* ctx, int n, int m, double a0, const double* a_host, const double* c_host,
                   const double* d_host, const double* xmin_host, const double* xmax_host, pf2_mma** out);
int pf2_mma_destroy(pf2_mma* mma);
int pf2_mma_set_parameters(pf2_mma* mma, double raa0, double albefa, double move, double asyinit, double asydecr,
                           double asyincr, double epsvalue);
int pf2_mma_is_convergence(pf2_mma* mma, double f, int* converged);
/* UpdateVariables(xk, f, dfdx, g[m], dgdx[m][n]); dgdx_dev is m*n row-major.  x_dev updated in place. */
int pf2_mma_update(pf2_mma* mma, double* x_dev, double f, const double* dfdx_dev, const double* g_host,
                   const double* dgdx_dev, int* newton_steps_out);

/* ---- CONLIN (CONLIN.h:18-26): a pf2_mma handle in CONLIN mode; update / convergence through pf2_mma_update / _is_convergence */
int pf2_conlin_create(pf2_ctx* ctx, int n, int m, double a0, const double* a_host, const double* c_host,
                      const double* d_host, const double* xmin_host, const double* xmax_host, pf2_mma** out);
int pf2_conlin_set_parameters(pf2_mma* conlin, double move, double epsvalue);

/* ---- the device-resident design loop (sample_optimize_density_{oc,mma}.cpp:83-208) ------------------------- */
/* params[12] = {E0,E1,poisson,p,weightlimit,scale0,scale1,thickness,beta0,beta_period,cg_itrmax,cg_eps}
 * optp: OC {iota,lmin,lmax,leps,move} | MMA {raa0,albefa,move,asyinit,asydecr,asyincr,epsvalue,a0,a,c,d,xmin,xmax}
 *       | CONLIN {move,epsvalue,a0,a,c,d,xmin,xmax} */
int pf2_simp_create(pf2_ctx* ctx, pf2_mesh* mesh, pf2_dofmap* map, pf2_csr* A, pf2_filter* filter, int eq, int opt_kind,
                    const double* optp, const double params[12], int nload, const int* load_node_host,
                    const int* load_dof_host, const double* load_val_host, pf2_simp** out);
int pf2_simp_destroy(pf2_simp* S);
int pf2_simp_set_design(pf2_simp* S, const double* s_host);
int pf2_simp_set_solver(pf2_simp* S, int solver);
/* Restart the loop at iteration 0 (driver :78-83) with the design s_host: Heaviside beta, the optimiser's previousvalue and
 * iteration count (MMA / CONLIN rebuild their asymptotes for k < 2, MMA.h:133-141) and the warm-start state are reset. */
int pf2_simp_reset(pf2_simp* S, const double* s_host);
/* Opt-in: the PCG of design iteration k starts from the displacements of iteration k-1 instead of 0 (pf2_solve_x0; the
 * reference's drivers always start from 0, CG.h:423).  Same stopping rule; compliance and densities agree within the solver
 * tolerance (tests/test_gpu_parity.py). */
int pf2_simp_set_warm_start(pf2_simp* S, int on);
/* One design iteration; the design never leaves the device.  stats[8] = {f, g, converged, cg_iters, cg_relres,
 * optimizer_steps, beta, k}.  When `converged` is set the design was NOT updated (the driver breaks, :192-195). */
int pf2_simp_iterate(pf2_simp* S, int check_convergence, double stats[8]);
/* The same iteration through host buffers (end-to-end path): uploads s, runs, downloads s, rho. */
int pf2_simp_iterate_host(pf2_simp* S, int check_convergence, const double* s_in_host, double* s_out_host,
                          double* rho_out_host, double stats[8]);
int pf2_simp_get(pf2_simp* S, double* s_host, double* rho_host, double* u_nodal_host, double* r_nodal_host);
/* The drivers' VTK dump of one iteration (sample_optimize_density_oc.cpp:175-184 = MakeHeadderToVTK + AddPointsToVTK + AddElementToVTK +
 * AddElementTypes + AddPointVectors u [+ r] + AddElementScalers rho as "s", ExportToVTK.h:19-137) written from the device-resident state:
 * the fields are staged on the device in the writers' layout, cross in one copy and are formatted like `ostream << double`, so the file
 * is byte-identical to what the reference's writers produce from the same values.  cell_type: VTK cell type of every element (9 = quad,
 * 5 = triangle, 12 = hexahedron); with_reactions: also recompute and write the nodal reactions r = K(rho) u. */
int pf2_simp_export_vtk(pf2_simp* S, const char* path, int cell_type, int with_reactions);
/* per-phase device time of the last iteration, ms: {filter, assemble, solve, compliance+sens, filter-sens, update} */
int pf2_simp_phase_ms(pf2_simp* S, double ms[6]);
int pf2_simp_cg_stats(pf2_simp* S, double* spmv_ms_avg, long long* spmv_calls);

/* ---- the level-set design loop (sample_optimize_levelset.cpp:75-192; PlaneStress.h:21, ReactionDiffusion.h:21-145,
 *      General.h:193-235) ---------------------------------------------------------------------------------------- */
/* mesh: Q4; map_u / K: the 2-dof displacement numbering and its pattern (pf2_dofmap_create, pf2_csr_pattern);
 * phi_fixed_nodes: nodes where the level-set function is held at 0 (driver :63-68);
 * prm = { Vmax, tau, E0, Emin, nu, nvol, dt, d, p } (driver :25-34); tmax bounds the objective history.
 * The reaction-diffusion matrix T = Me/dt + tau*nelem*Ke is assembled once here (it does not depend on the design).
 * Initial state phi = 1, str = 1 (driver :70-71); pf2_levelset_set_state overrides it. */
int pf2_levelset_create(pf2_ctx* ctx, pf2_mesh* mesh, pf2_dofmap* map_u, pf2_csr* K, int nphi, const int* phi_fixed_nodes_host,
                        const double prm[9], int tmax, int nload, const int* load_node_host, const int* load_dof_host,
                        const double* load_val_host, pf2_levelset** out);
int pf2_levelset_destroy(pf2_levelset* ls);
int pf2_levelset_set_state(pf2_levelset* ls, const double* phi_host, const double* str_host);
/* one pass of the loop body; stats[8] = { objective (the driver prints objective/nelem), volume, lambda, converged,
 * CG iterations of the displacement solve, its relative residual, CG iterations of the phi solve, t }.
 * When the driver's convergence test fires (:128-137) the level-set update is skipped, as there. */
int pf2_levelset_iterate(pf2_levelset* ls, int check_convergence, double stats[8]);
/* phi (nnode), str (nelem), u (nnode*2) of the last iteration; any pointer may be NULL */
int pf2_levelset_get(pf2_levelset* ls, double* phi_host, double* str_host, double* u_host);

/* ---- multi-GPU: row-block (x-slab) partition over the GPUs of one box (no counterpart in the reference, whose only
 *      parallelism is the OpenMP loop of CSR.h:114) ----------------------------------------------------------------- */
typedef struct pf2_dist pf2_dist;
/* rank 0 creates the 128-byte NCCL id and hands it to the other ranks through any channel (torch.distributed, MPI, a file) */
int pf2_dist_unique_id(char out[128]);
int pf2_dist_create(pf2_ctx* ctx, int rank, int nranks, const char id[128], pf2_dist** out);
int pf2_dist_destroy(pf2_dist* d);
int pf2_dist_allreduce_sum(pf2_dist* d, double* dev, int count);
/* halo = {sendL_off, recvL_off, cntL, sendR_off, recvR_off, cntR}: contiguous ranges exchanged with rank-1 / rank+1 */
int pf2_dist_halo(pf2_dist* d, double* vec_dev, const int halo[6]);
/* Mark a LOCAL matrix (local mesh = owned element planes + one ghost plane per side) as one row block of a partitioned
 * system: rows [own_lo, own_hi) are owned; pf2_solve then runs the distributed PCG (halo exchange of p + allreduces). */
int pf2_csr_set_partition(pf2_csr* A, pf2_dist* d, int own_lo, int own_hi, const int halo[6]);
/* Peer-memory backend for the partitioned PCG (optional; NCCL otherwise): the halo exchange is fused into the p-update kernel
 * (boundary planes are stored straight into the neighbours' ghost ranges over NVLink) and the dot-product allreduces run
 * inside one-warp kernels over IPC-mapped arenas.  export: this rank's IPC handles {arena, Krylov slab} (2 x 64 bytes);
 * exchange them through any channel; import: all ranks' handles (world x 128 bytes) and all ranks' meta (world x 8 ints =
 * {row halo descriptor[6], local rows, pf2_csr_pcg_capable}). */
int pf2_csr_p2p_export(pf2_csr* A, char handles_out[128]);
/* meta[7] of pf2_csr_p2p_import: 1 when this rank's slab qualifies for the persistent PCG kernel's partitioned instantiation; the
 * kernel is used only when every rank reports 1 (all ranks must run the same protocol) */
int pf2_csr_pcg_capable(pf2_csr* A, int* out);
int pf2_csr_p2p_import(pf2_csr* A, const char* all_handles, const int* all_meta);
/* Unmap the neighbours' Krylov slabs again.  All ranks call it and synchronise before any of them destroys its matrix. */
int pf2_csr_p2p_release(pf2_csr* A);
/* Make a design loop built on a slab's local mesh one part of a partitioned loop: elements [own_elem_lo, own_elem_hi) are
 * owned, elem_halo = contiguous element ranges exchanged with the neighbours, n_global_elems = elements of the whole mesh
 * (the volume constraint is global). */
int pf2_simp_set_partition(pf2_simp* S, pf2_dist* d, int own_elem_lo, int own_elem_hi, const int elem_halo[6], long long n_global_elems);

#ifdef __cplusplus
}
#endif
#endif /* PANSFEM2_B200_H */
