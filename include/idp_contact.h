/* idp_contact.h — C ABI of libidp_contact.so: the B200 (sm_100a) implementation of the IPC contact hot path of
 * ipc-sim/IDP for codimensional triangle surfaces, instantiation <T=double, dim=3, shell=false, elasticIPC=false>.
 *
 * Every entry point below replaces one operator (or one piece of marshalling) of the reference's C++ boundary
 * "B2" (SURVEY.md §8b): the six free function templates of Library/FEM/IPC.h. Citations are relative to
 * /root/reference/Library. The reference-side binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes, int status codes, no exceptions, no ownership transfer of host buffers;
 *   - host arrays are dense: positions nV x 3 doubles with a caller-given stride (3 for packed xyz, 4 for the
 *     reference's 32-byte VECTOR<double,3>, Math/VECTOR.h:33-44), indices int32;
 *   - one context = one GPU = one CUDA stream; calls are synchronous for the calling host thread;
 *   - there is NO CPU fallback: without a CUDA device idp_create fails with IDP_ERR_CUDA.
 */
#ifndef IDP_CONTACT_H
#define IDP_CONTACT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct idp_ctx idp_ctx;

enum idp_status {
    IDP_OK = 0,
    IDP_ERR_CUDA = 1,                   /* CUDA runtime / allocation failure (see idp_last_error) */
    IDP_ERR_INVALID = 2,                /* bad argument or call order */
    IDP_ERR_NONPOSITIVE_DISTANCE = 3,   /* IPC.h:817-820,873-876,925-928: "distance detected during barrier evaluation" -> exit(-1) */
    IDP_ERR_CCD_ZERO_STEP = 4,          /* IPC.h:2014-2032: CCD returned a zero step -> exit(-1) */
    IDP_ERR_UNSUPPORTED_PRIMITIVE = 5,  /* rod / particle / NNExclusion / dim==2 / shell / elasticIPC inputs: rejected, never emulated */
    IDP_ERR_NCCL = 6,
    IDP_ERR_CCD_ITERATION_CAP = 7,      /* additive CCD exceeded IDP_ACCD_MAX_ITER trips (the reference would loop forever) */
    IDP_ERR_EIGEN_NO_CONVERGENCE = 8    /* makePD (Math/UTILS.h:9-27): the symmetric eigen-solver hit its iteration cap for some row */
};

/* stage ids for idp_stage_ms: names follow the reference's TIMER_FLAG scopes */
enum idp_stage {
    IDP_STAGE_CCS_BUILD_HASH = 0,   /* Compute_Constraint_Set_Build_Hash */
    IDP_STAGE_CCS_PT = 1,           /* Compute_Constraint_Set_PT (broad + narrow) */
    IDP_STAGE_CCS_EE = 2,           /* Compute_Constraint_Set_EE */
    IDP_STAGE_CCS_MERGE = 3,        /* Compute_Constraint_Set_Merge */
    IDP_STAGE_BARRIER = 4,          /* Compute_Barrier / _Gradient / _Hessian per-row kernel */
    IDP_STAGE_CSR = 5,              /* constructCSRMatrixFromTriplet */
    IDP_STAGE_CCD_BUILD_HASH = 6,   /* Compute_Intersection_Free_StepSize_Build_Hash */
    IDP_STAGE_CCD_PT = 7,           /* Compute_Intersection_Free_StepSize_PT */
    IDP_STAGE_CCD_EE = 8,           /* Compute_Intersection_Free_StepSize_EE */
    IDP_STAGE_MIN_DIST = 9,         /* Compute_Min_Dist */
    IDP_STAGE_UPLOAD = 10,          /* host -> device marshalling of positions / directions */
    IDP_STAGE_K_BARRIER = 11,       /* the k_barrier launch alone (dominant FP64 kernel) */
    IDP_STAGE_K_QUERY = 12,         /* the last broad-phase query launch alone */
    IDP_STAGE_K_ACCD = 13,          /* the last additive-CCD launch alone */
    IDP_STAGE_K_CLASSIFY = 14,      /* the last classification launch alone */
    IDP_STAGE_COMM = 15,            /* NCCL collectives, accumulated since the last idp_constraint_set */
    IDP_STAGE_COUNT = 16
};

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int idp_create(int device, idp_ctx** out);
void idp_destroy(idp_ctx* ctx);
const char* idp_last_error(idp_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the context's own stream */
int idp_set_stream(idp_ctx* ctx, void* cuda_stream);

/* ---- inputs --------------------------------------------------------------------------------------------- */
/* Surface primitives in the ordering contract of Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834):
 * boundaryNode ascending, boundaryEdge lexicographic with first-triangle orientation, boundaryTri in element order.
 * These are the boundaryNode / boundaryEdge / boundaryTri / DBCb arguments of Compute_Constraint_Set (FEM/IPC.h:19-36)
 * and Compute_Intersection_Free_StepSize (IPC.h:1879-1890). dbc may be NULL (no Dirichlet nodes). Call once per time step. */
int idp_set_mesh(idp_ctx* ctx, int nV, int nBN, const int* bnode, int nBE, const int* bedge2, int nBT, const int* btri3,
    const uint8_t* dbc);
/* The reference's rod / particle / NNExclusion containers (IPC.h:25-27): the B2 shim reports their sizes here;
 * any non-zero size returns IDP_ERR_UNSUPPORTED_PRIMITIVE (out of scope). */
int idp_declare_unsupported(idp_ctx* ctx, int n_rod, int n_particle, int n_nn_exclusion);
/* X (MESH_NODE<T,3>&, IPC.h:20) — current positions, nV rows, `stride` doubles per row (3 or 4). Call per iterate. */
int idp_set_positions(idp_ctx* ctx, const double* x, int stride);
/* nodeAttr.x0 (rest positions read at IPC.h:421-426, 880-885) — needed only for the edge-edge mollifier. */
int idp_set_rest_positions(idp_ctx* ctx, const double* x0, int stride);

/* ---- Compute_Constraint_Set (FEM/IPC.h:19-740) ------------------------------------------------------------- */
/* Builds the constraint set on the device for the current positions: spatial-hash broad phase (Grid/SPATIAL_HASH.h:28-291),
 * AABB filter (Math/Distance/CCD.h:149-185), distance-type classification (DISTANCE_TYPE.h), duplicate PP/PE merge
 * (IPC.h:599-654). Row order: [PT rows][EE and mollified rows][merged PP/PE rows in key order]; rows inside the first
 * two groups are sorted lexicographically (the reference's order there depends on unordered_set iteration). */
int idp_constraint_set(idp_ctx* ctx, double dhat2, double thickness, int* n_rows);
/* constraintSet (VECTOR<int,4>, 16 B/row) and stencilInfo (weight, dHat2) to the host. Either pointer may be NULL.
 * Sharded: the rows THIS rank holds (idp_last_count(ctx, 9) of them), not a collective. */
int idp_get_constraints(idp_ctx* ctx, int* rows4, double* info2);
/* The GLOBAL list (idp_last_count(ctx, 0) rows) in the order of the unsharded path, on every rank; dist2 = the per-row
 * values of the last idp_min_dist2 in the same order. Any pointer may be NULL. Sharded: a collective (all ranks call it
 * with the same NULL pattern); unsharded: the same as idp_get_constraints. */
int idp_gather_constraints(idp_ctx* ctx, int* rows4, double* info2, double* dist2);
/* Use a caller-supplied constraint set (the constraintSet / stencilInfo arguments of Compute_Barrier*, IPC.h:745-746). */
int idp_set_constraints(idp_ctx* ctx, int n_rows, const int* rows4, const double* info2);
/* Post-AABB candidate pairs of the last idp_constraint_set / idp_ccd_step, lexicographically sorted:
 * which = 0: PT (svI, sfI); 1: EE (eI, eJ); 2: CCD PT; 3: CCD EE. First call with pairs2 == NULL to get the count. */
int idp_get_candidates(idp_ctx* ctx, int which, long* n_pairs, int* pairs2);

/* ---- Compute_Barrier / _Gradient / _Hessian (FEM/IPC.h:742-941, 943-1256, 1258-1731) ----------------------- */
/* E += sum of weighted barrier values over the current rows (adds, like the reference). kappa = kappa[0]. */
int idp_barrier_energy(idp_ctx* ctx, double dhat2, double kappa, double thickness, double* E_inout);
/* g_accum[v*stride + a] += barrier gradient (nodeAttr.g accumulation, IPC.h:1034-1042); g_accum may be NULL to keep the
 * gradient on the device only (idp_gradient_device). */
int idp_barrier_gradient(idp_ctx* ctx, double dhat2, double kappa, double thickness, double* g_accum, int stride);
/* Per-row Hessians with optional PSD projection (makePD, Math/UTILS.h:9-27), sort-reduced into a device-resident scalar
 * CSR of size 3nV x 3nV with the pattern/ordering Eigen::setFromTriplets produces (Math/CSR_MATRIX.h:49-56):
 * duplicates summed, columns ascending per row, explicit zeros kept. */
int idp_barrier_hessian(idp_ctx* ctx, double dhat2, double kappa, double thickness, int project_spd, long* nnz);
/* E, g and H in one pass over the rows (what one Newton iteration needs); any of E_inout / nnz may be NULL. */
int idp_barrier_all(idp_ctx* ctx, double dhat2, double kappa, double thickness, int project_spd, double* E_inout, long* nnz);
/* copy the CSR to the host: exactly the (ptr, col, val) arrays CSR_MATRIX::Construct_From_CSR takes (CSR_MATRIX.h:33-47) */
int idp_get_hessian_csr(idp_ctx* ctx, int* ptr, int* col, double* val);
/* Asynchronous variants for hosts that overlap the hand-over with the next operator (the Newton step goes on with the line
 * search while the matrix travels): *_begin enqueues the device-to-host copies on a copy stream and returns; the output
 * buffers must stay valid (and should be pinned) until idp_transfers_end has returned, which completes every pending
 * transfer. The CSR crosses PCIe compactly (values + one column vertex per 3x3 block + block-row starts) and the scalar
 * (ptr, col) arrays of Construct_From_CSR are expanded by host threads inside idp_transfers_end while the values are still
 * in flight. Results are identical to the synchronous getters. */
int idp_get_constraints_begin(idp_ctx* ctx, int* rows4, double* info2);
int idp_get_hessian_csr_begin(idp_ctx* ctx, int* ptr, int* col, double* val);
int idp_transfers_end(idp_ctx* ctx);
/* device pointers (valid until the next idp_barrier_hessian / idp_barrier_all on this context) */
int idp_hessian_csr_device(idp_ctx* ctx, const int** d_ptr, const int** d_col, const double** d_val, long* nnz);
int idp_gradient_device(idp_ctx* ctx, const double** d_g_xyz);
/* g_accum[v*stride + a] += the gradient of the last idp_barrier_gradient / idp_barrier_all (sharded: the all-reduced one) */
int idp_get_gradient(idp_ctx* ctx, double* g_accum, int stride);

/* ---- the caller-side steps around the barrier Hessian, kept on the device -------------------- */
/* Terms assembled into the SAME device CSR by every following idp_barrier_hessian / idp_barrier_all (they persist until
 * replaced; n_elem = 0 / NULL removes a term; idp_set_mesh* clears both):
 *   flow term (FEM/Shell/INC_POTENTIAL.h:323-339, the `flow` branch of Compute_IncPotential_Hessian): per triangle element
 *     (`stride` ints per element, volume vol[e]) and axis d:  H[(v_i,d),(v_i,d)] += 2 h vol / 6 and
 *     H[(v_i,d),(v_j,d)] -= h vol / 6 for the two other vertices j of the element;
 *   lumped mass (`sysMtr.Get_Matrix() += M.Get_Matrix()`, INC_POTENTIAL.h:383-386, M from Shell/DISCRETE_SHELL.h:279-318):
 *     H[(v,d),(v,d)] += m[v] for every vertex with m[v] != 0.
 * With these the matrix leaving idp_barrier_hessian is what Construct_From_Triplet + `+= M` hand to Project_DBC (values;
 * the pattern additionally keeps explicit zeros inside every 3x3 block). Sharded contexts split the elements / vertices
 * by contiguous ranges like the query primitives. */
int idp_system_set_flow_term(idp_ctx* ctx, int n_elem, const int* elem3, int stride, const double* vol, double h);
int idp_system_set_mass(idp_ctx* ctx, const double* m_per_vertex);
/* Elastic terms of the discrete-shell system (FEM/Shell/INC_POTENTIAL.h:75-96, 213-232, 340-352: the non-flow branch), assembled
 * into the same device CSR as the barrier rows by every following idp_barrier_hessian / idp_barrier_all, each element's
 * Hessian projected to PSD when project_spd is set (makePD, Math/UTILS.h:9-27); they persist until replaced (n = 0 removes a
 * term; idp_set_mesh* clears them). Elements whose vertices are all Dirichlet nodes (mask of idp_set_mesh*) are skipped.
 *   membrane (FEM/Shell/MEMBRANE.h:8-315, neo-Hookean on the first fundamental form): per triangle (`stride` ints) the REST
 *     first fundamental form ib3 = (IB00, IB01, IB11) (elemAttr IB, not inverted), vol, lambda, mu (elasticityAttr), weight h^2;
 *   hinges (FEM/Shell/BENDING.h:52-80,176-213,438-497, KL = false): per hinge the stencil (v0; v1, v2; v3) of edgeStencil and
 *     info3 = (rest angle, rest edge length, rest height) of edgeInfo (Shell/DISCRETE_SHELL.h:169-211); k = bendingStiffMult *
 *     E t^3 / (24 (1 - nu^2)); energy h^2 k (theta - thetabar)^2 ebar / hbar. */
int idp_system_set_membrane(idp_ctx* ctx, int n_elem, const int* elem3, int stride, const double* ib3, const double* vol, const double* lambda,
    const double* mu, double h);
int idp_system_set_hinges(idp_ctx* ctx, int n_hinge, const int* stencil4, const double* info3, double k, double h);
/* Compute_Membrane_Energy + Compute_Bending_Energy (adds to *E_inout) and the matching gradients (g_accum[v*stride + a] +=;
 * g_accum may be NULL) at the current positions. The barrier operators above stay barrier-only. */
int idp_elastic_energy(idp_ctx* ctx, double* E_inout);
int idp_elastic_gradient(idp_ctx* ctx, double* g_accum, int stride);
/* Lagged friction (FEM/FRICTION.h:17-662), reusing the constraint rows:
 *   idp_friction_update  = Compute_Friction_Basis (:17-124): freezes the non-mollified rows of the CURRENT constraint set at the
 *     current positions with their closest-point weights, tangent bases and normal forces -b'(d) 2 sqrt(d) (lagged until the
 *     next update); *n_friction_rows = how many rows carry friction;
 *   idp_friction_set     = the Xn / epsv2*h*h / mu arguments of Compute_Friction_Potential / _Gradient / _Hessian (:172-662);
 *     mu = 0 switches the term off;
 *   idp_friction_energy / _gradient add mu * sum lambda f0(|u|) and its gradient at the current positions; the Hessian blocks
 *     (inner 2x2 matrix is PSD by construction, so no projection is needed) are assembled into the same device CSR by every
 *     following idp_barrier_hessian / idp_barrier_all (INC_POTENTIAL.h:375-377);
 *   idp_get_friction copies the frozen rows in the reference's layout (rows, closest-point parameters (2), tangent basis
 *     (3x2 column major), normal force); any pointer may be NULL.
 *   idp_friction_set_components = the compNodeRange / muComp arguments of Compute_Friction_Coef (:126-170): with n_comp > 0 every
 *     following idp_friction_update scales the normal force of a row by mu_comp[c0 + c1 * n_comp], c0 / c1 the components of its
 *     first and of its opposite primitive (vertex v belongs to the first component with v < comp_node_range[c]); the caller then
 *     passes mu = 1 like the reference does. n_comp = 0 switches it off. idp_set_mesh* clears the rows. */
int idp_friction_update(idp_ctx* ctx, double dhat2, double kappa, double thickness, long* n_friction_rows);
int idp_friction_set(idp_ctx* ctx, const double* xn, int stride, double epsv2_h2, double mu);
int idp_friction_set_components(idp_ctx* ctx, int n_comp, const int* comp_node_range, const double* mu_comp);
int idp_friction_energy(idp_ctx* ctx, double* E_inout);
int idp_friction_gradient(idp_ctx* ctx, double* g_accum, int stride);
int idp_get_friction(idp_ctx* ctx, long* n_rows, int* rows4, double* closest2, double* basis6, double* normal_force);
/* CSR_MATRIX::Project_DBC (Math/CSR_MATRIX.h:130-141) applied to the device CSR with the Dirichlet mask of idp_set_mesh:
 * every stored entry whose row or column vertex is a Dirichlet node becomes (row == col). Single-GPU contexts. */
int idp_project_dbc(idp_ctx* ctx);
/* the same with a caller-supplied vertex mask (nV bytes; NULL = the mask of idp_set_mesh*): Compute_IncPotential_Hessian projects
 * only the Dirichlet nodes that do not move (DBCb_fixed) while the augmented-Lagrangian penalty is active (INC_POTENTIAL.h:387-394) */
int idp_project_dbc_mask(idp_ctx* ctx, const uint8_t* mask);
/* The role of Solve_Direct (Math/DIRECT_SOLVER.h:14-88: CHOLMOD / SimplicialLDLT on the host) without moving the matrix:
 * conjugate gradients with a 3x3 block-Jacobi preconditioner on the device CSR, x0 = 0, stops when |r| <= rel_tol |rhs| or
 * after max_iter iterations (iterative: the caller states the tolerance; *rel_residual reports what was reached).
 * rhs / sol: 3 nV doubles on the host (sol may be NULL). The matrix must be symmetric positive definite (project_spd = 1,
 * mass and / or Dirichlet rows present). Single-GPU contexts. */
int idp_solve_pcg(idp_ctx* ctx, const double* rhs, double* sol, double rel_tol, int max_iter, int* iters, double* rel_residual);
/* Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834) on the device, once per mesh instead of once per time
 * step on the host with a std::map (Shell/IMPLICIT_EULER.h:222-241): sets the context's mesh from the element list
 * (`stride` ints per triangle) in the reference's ordering contract. x (nV rows, xstride doubles; may be NULL = topology
 * only) provides the areas BNArea / BEArea / BTArea and the "zero summed area is not a boundary node" rule. */
int idp_set_mesh_from_triangles(idp_ctx* ctx, int nV, int nF, const int* tri, int stride, const double* x, int xstride, const uint8_t* dbc);
/* the primitives (and, after idp_set_mesh_from_triangles, the areas) the context holds; any pointer may be NULL */
int idp_get_surface_primitives(idp_ctx* ctx, int* nBN, int* bnode, int* nBE, int* bedge2, int* nBT, int* btri3, double* BNArea, double* BEArea,
    double* BTArea);

/* ---- Compute_Intersection_Free_StepSize (FEM/IPC.h:1879-2244) ---------------------------------------------- */
/* searchDir: nV rows x 3 doubles (std::vector<T>, stride 3 in the reference). alpha is in/out and includes the
 * span clamp of the CCD hash build (Grid/SPATIAL_HASH.h:466-482). */
int idp_ccd_step(idp_ctx* ctx, const double* search_dir, int stride, double thickness, double* alpha_inout);
/* the same in two halves: upload the direction once, then filter steps against device-resident inputs */
int idp_set_search_direction(idp_ctx* ctx, const double* search_dir, int stride);
int idp_ccd_step_resident(idp_ctx* ctx, double thickness, double* alpha_inout);

/* ---- Compute_Min_Dist2 (FEM/IPC.h:2246-2388) ---------------------------------------------------------------- */
/* dist2 (n_rows doubles, may be NULL) and minDist2 = min - thickness^2. With zero rows nothing is written (IPC.h:2253). */
int idp_min_dist2(idp_ctx* ctx, double thickness, double* dist2, double* min_dist2);

/* ---- multi-GPU (one context per rank; NCCL only for the natural reductions) ---------------------------------- */
/* Sharded semantics (after idp_comm_init): every entry point is called by ALL ranks with the same arguments.
 *   idp_constraint_set      queries split by contiguous primitive ranges; each rank keeps the rows it produced (direct PT /
 *                           EE rows of its range + the merged PP/PE rows of its vertex slab); *n_rows is the GLOBAL count;
 *   idp_get_constraints     NOT a collective: the rows of this rank ([its direct PT][its direct EE][its merged PP/PE] rows);
 *   idp_gather_constraints  collective: gathers the global list, bit-identical to the single-GPU one, on every rank;
 *   idp_barrier_*           E and the gradient are all-reduced (every rank returns the global values); the Hessian CSR of a
 *                           rank holds the partial sums of its rows -- the global Hessian is the sum of the P CSRs (their
 *                           patterns are nearly disjoint: a primitive range touches a vertex slab);
 *   idp_ccd_step*           sharded queries, all-reduce(min) of the step;
 *   idp_min_dist2           all-reduce(min); dist2 = the values of this rank's rows (global order: idp_gather_constraints);
 *   idp_set_constraints     rows are replicated and evaluated by the owner of the vertex chunk of their smallest vertex. */
/* out_id: 128 bytes (ncclUniqueId) produced on rank 0 and broadcast by the host program */
int idp_comm_unique_id(void* out_id128);
int idp_comm_init(idp_ctx* ctx, int rank, int nranks, const void* id128);
/* In-process group: ctxs[r] becomes rank r of nranks (<= 8) contexts of THIS process, on the same or on different GPUs.
 * Same sharded semantics as idp_comm_init, but the exchanges go over peer memory (device-to-device copies and a
 * rank-ordered reduction kernel that loads the peers' buffers) instead of NCCL. Every context must then be driven by
 * its own host thread (the collectives meet at a host barrier). Used by the sharding parity tests on one GPU. */
int idp_comm_init_local(idp_ctx** ctxs, int nranks);
/* in-process group only: wake every rank waiting in a collective with IDP_ERR_INVALID (a host thread gave up) */
int idp_comm_abort(idp_ctx* ctx);
/* shard without a communicator (results stay partial; for tests of the partition logic). nranks <= 8. */
int idp_set_shard(idp_ctx* ctx, int rank, int nranks);

/* ---- instrumentation ------------------------------------------------------------------------------------------ */
long idp_kernel_launches(idp_ctx* ctx);     /* this library's own kernels launched since idp_reset_counters */
long idp_library_calls(idp_ctx* ctx);       /* CUB device-wide primitives invoked since idp_reset_counters */
void idp_reset_counters(idp_ctx* ctx);
float idp_stage_ms(idp_ctx* ctx, int stage); /* device time of the last execution of a stage (CUDA events) */
long idp_last_count(idp_ctx* ctx, int what); /* 0 rows, 1 PT cand, 2 EE cand, 3 CCD PT cand, 4 CCD EE cand, 5 ACCD trips, 6 nnz, 7 unique 3x3 blocks, 8 device allocations made so far, 9 rows held by this rank */
/* FP64 pipe microbenchmark: register-resident DFMA chains on every SM; returns measured TFLOP/s (roofline denominator) */
int idp_measure_fp64_tflops(idp_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
