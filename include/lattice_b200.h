/* lattice_b200.h -- C ABI of the B200-native beam-FEM hot path for pyLatticeDSO.
 *
 * The reference (Tcadart/pyLatticeDSO) has no FFI layer: its hot path is three
 * Python callables that delegate to FEniCSx/PETSc.  Each
 * entry point below names the reference interface (file:line under the
 * pyLatticeDSO checkout) whose arithmetic it replaces.  INTEGRATION.md shows
 * the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer (e.g. torch.Tensor.data_ptr());
 *     the caller owns all buffers; the library allocates only inside lat_ctx
 *     (workspace, freed by lat_ctx_destroy) and reports sizes by two-call
 *     queries; FP64 values, int32 indices, int64 sizes.
 *   - per-node DOF order [ux,uy,uz,rx,ry,rz]; global DOF = 6*node + d
 *     (pyLatticeDesign/point.py:68).
 *   - BSR: 6x6 blocks, block rows sorted, column indices sorted inside a row,
 *     block values row-major, 36 doubles per block.
 *   - return value: 0 ok; <0 argument error; >0 CUDA/NCCL error (see
 *     lat_last_error).  PCG non-convergence is NOT an error: it is reported in
 *     lat_pcg_result.info exactly like conjugate_gradient_solver.py:73,98,103,108.
 *   - one ctx <-> one device <-> one stream; a ctx is not thread-safe; work is
 *     enqueued on the ctx stream and only the calls documented as "syncs"
 *     wait for it.
 */
#ifndef LATTICE_B200_H
#define LATTICE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lat_ctx lat_ctx;

#define LAT_OK 0
#define LAT_ERR_ARG (-1)
#define LAT_ERR_STATE (-2)
#define LAT_ERR_UNSUPPORTED (-3)

/* assembly modes */
#define LAT_ASM_GATHER 0 /* deterministic: one thread per BSR block gathers its elements */
#define LAT_ASM_ATOMIC 1 /* one thread per element, warp-aggregated FP64 atomics          */
#define LAT_ASM_ROWS 2   /* deterministic: one thread per block row, 256-bit stores; fastest, the host default */

/* preconditioners */
#define LAT_PC_NONE 0
#define LAT_PC_JACOBI 1
#define LAT_PC_BLOCK6 2

int lat_version(void);

/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or 0 for
 * the legacy default stream. */
int lat_ctx_create(int device, void* stream, lat_ctx** out);
int lat_ctx_destroy(lat_ctx* ctx);
const char* lat_last_error(lat_ctx* ctx);
/* blocks until everything enqueued on the ctx stream is done */
int lat_ctx_sync(lat_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
int64_t lat_launch_count(lat_ctx* ctx);

/* ---- A1: 12x12 Timoshenko element stiffness ---------------------------------
 * Replaces the FFCx element kernel generated from SimulationBase.define_K_form
 * (pyLatticeSim/simulation_base.py:141-156,190-197,220-225) with the local frame
 * of BeamModel.calculate_local_coordinate_system (beam_model.py:197-216) and the
 * section constants of Material.compute_mechanical_properties
 * (material_definition.py:142-156).  drad != 0 returns dK_e/dr instead
 * (material_definition.py:207-223).
 * Ke: [n_elem][12][12] row-major, element DOFs [w1,th1,w2,th2] in global axes. */
int lat_elem_stiffness(lat_ctx* ctx, const double* x, const double* y, const double* z,
                       const int32_t* en0, const int32_t* en1, const double* rad,
                       int64_t n_elem, double young, double nu, double kappa,
                       int drad, double* Ke);

/* ---- A3: sparsity pattern ----------------------------------------------------
 * Replaces dolfinx create_matrix / PETSc preallocation behind
 * fem_petsc.assemble_matrix (simulation_base.py:480-481, schur_complement.py:69-71).
 * Builds the 6x6 block graph (node adjacency + self) by counting sort on the
 * row node followed by a segmented per-row sort/unique and scans.
 * Call 1: lat_bsr_pattern_build  -> *nnzb (kept inside ctx)       [syncs]
 * Call 2: lat_bsr_pattern_export -> rowptr[n_nodes+1], colidx[nnzb],
 *         elem_block[n_elem*4] = BSR block index of the (n0,n0),(n0,n1),(n1,n0),
 *         (n1,n1) quadrants of every element (the scatter map; may be NULL). */
int lat_bsr_pattern_build(lat_ctx* ctx, const int32_t* en0, const int32_t* en1,
                          int64_t n_elem, int64_t n_nodes, int64_t* nnzb);
int lat_bsr_pattern_export(lat_ctx* ctx, int32_t* rowptr, int32_t* colidx, int32_t* elem_block);

/* Scalar CSR structure of a BSR pattern: indptr[6n+1], indices[36 nnzb].
 * Bit-exact with scipy.sparse.coo_matrix((v,(i,j))).tocsr() + sum_duplicates +
 * sort_indices fed with all 144 entries of every element (oracle.assemble_csr). */
int lat_csr_structure(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx,
                      int64_t n_nodes, int32_t* indptr, int32_t* indices);
/* BSR block values -> CSR value order (same structure as lat_csr_structure) */
int lat_bsr_to_csr_values(lat_ctx* ctx, const int32_t* rowptr, int64_t n_nodes,
                          const double* bsr_vals, double* csr_vals);

/* ---- A1+A3: fused element generation + global assembly ---------------------
 * K = sum_e P_e^T K_e P_e into BSR values (fem_petsc.assemble_matrix without bcs,
 * simulation_base.py:480, schur_complement.py:69-71).  Requires the pattern of
 * the same mesh to be resident in ctx (lat_bsr_pattern_build).  vals[nnzb*36]
 * is overwritten.  drad != 0 assembles sum_e chain_e * dK_e/dr (chain may be NULL). */
int lat_assemble_bsr(lat_ctx* ctx, const double* x, const double* y, const double* z,
                     const int32_t* en0, const int32_t* en1, const double* rad,
                     const double* chain, int64_t n_elem, int64_t n_nodes,
                     double young, double nu, double kappa, int mode, int drad,
                     double* vals);

/* ---- A4: Dirichlet elimination ------------------------------------------------
 * fem_petsc.assemble_matrix(bcs) + apply_lifting + set_bc (simulation_base.py:480-492):
 *   vals_bc = rows/cols of constrained DOFs zeroed, unit diagonal (may alias vals);
 *   b = f - K[:,c] g on free rows, b[c] = g[c].
 * fixed: uint8[6n] (non-zero = constrained), g,f,b: double[6n]. */
int lat_apply_dirichlet(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx,
                        int64_t n_nodes, const double* vals, const uint8_t* fixed,
                        const double* g, const double* f, double* vals_bc, double* b);

/* u[c] = g[c] on constrained DOFs (fem_petsc.set_bc, simulation_base.py:492): makes the imposed
 * values exact after an iterative solve. */
int lat_set_dirichlet_values(lat_ctx* ctx, const uint8_t* fixed, const double* g, int64_t n_dof, double* u);

/* ---- A10: y = K x (reactions R = K_unconstrained u, simulation_base.py:582-645) */
int lat_bsr_spmv(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx,
                 const double* vals, int64_t n_nodes, const double* x, double* y);

/* ---- A5/A6: preconditioned conjugate gradients -------------------------------
 * Replaces KSP preonly+LU (simulation_base.py:502-511) and restates
 * conjugate_gradient_solver (pyLatticeSim/conjugate_gradient_solver.py:15-122).
 * reference_semantics = 1 reproduces that function's iteration exactly: x0 = 0,
 * alpha = min(rz/pAp, alpha_max) (:78-79), p <- z every restart_every iterations
 * (:89-90), stop when |r| <= tol |b| (:97) or |p| < mintol (|x| + 1e-12) (:102),
 * info 0/1/2 (:73,98,103,108).  reference_semantics = 0 is textbook PCG with the
 * single test |r| <= tol |b|, run in the Chronopoulos-Gear arrangement (one gather, one
 * reduction per iteration; same Krylov iterates in exact arithmetic). */
typedef struct {
  double tol;
  double mintol;
  double alpha_max;
  int64_t restart_every;
  int32_t maxiter;
  int32_t precond;
  int32_t reference_semantics;
  int32_t check_every; /* iterations between host polls of the device status (0 = default 32) */
  int32_t profile_iters; /* >0: the first profile_iters iterations are launched outside the CUDA graph
                            with CUDA events around each kernel (fills spmv_ms / update_ms) */
  int32_t reserved;      /* bit 1: experimental TMA-staged SpMV kernel; bit 2: no CUDA graph in the multi-GPU path;
                            bit 3: classic two-reduction recurrences instead of Chronopoulos-Gear (textbook mode);
                            bit 4 (lat_pcg_bsr_dist): NVLink peer-memory halo/all-reduce instead of NCCL;
                            bit 5: with bit 4, overlap the halo (side stream) with the product of the interior rows -- opt-in,
                            measured no faster (profiles/r01_overlap_ab.txt);
                            bit 6: with bit 4, use the separate halo kernel (k_p2p_halo) instead of the default fused halo
                            (<= 2 neighbours: the update kernel pushes the boundary entries of u into the neighbours' ghost
                            sections and only the product CTAs that read ghosts wait; 8-12 % faster per iteration) */
} lat_pcg_opts;

typedef struct {
  int32_t iters;
  int32_t info;      /* 0 converged, 1 maxiter, 2 alpha < 1e-6 seen (reference flag), 3 breakdown (p.Ap <= 0 / NaN),
                        4 peer time-out (multi-GPU peer-memory path), 5 the recurrence residual met tol but the
                        recomputed |b - A x| stayed above 2 tol |b| after two restarts (see true_relres) */
  double relres;     /* |r| / |b| at exit (recurrence residual) */
  double norm_b;
  double solve_ms;   /* device time of the iteration loop (CUDA events on the ctx stream) */
  int64_t launches;  /* kernels launched by this call */
  double spmv_ms;    /* mean duration of the fused SpMV kernel over the profiled iterations (0 if none) */
  double update_ms;  /* mean duration of the update kernel over the profiled iterations */
  int32_t profiled;  /* iterations actually profiled */
  int32_t reserved;  /* bits 0-7: restarts triggered by the true-residual safeguard; bit 8: the iteration batches
                        ran as CUDA graphs (always on one GPU; multi-GPU unless capture failed or opts bit 2) */
  double true_relres; /* |b - A x| / |b| recomputed after convergence (textbook mode; -1 if not measured) */
} lat_pcg_result;

/* x is overwritten (x0 = 0).  [syncs] */
int lat_pcg_bsr(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                int64_t n_nodes, const double* b, double* x, const lat_pcg_opts* opts,
                lat_pcg_result* result);

/* ---- matrix-free operator (B200 design choice, no reference counterpart) ---------------------
 * The same linear system as lat_assemble_bsr + lat_apply_dirichlet + lat_pcg_bsr, but K is never
 * stored: every product regenerates the element action from the geometry (beam_model.py:197-216 and
 * material_definition.py:147 restated as f = a*d + b*t(t.d) + c*(t x d) per 3-vector, see
 * csrc/matfree.cuh).  About a tenth of the HBM bytes of the assembled product and no 15 GB matrix at
 * octet 100^3; iterates agree with the assembled path to rounding (tests/test_gpu_matfree.py).
 *   lat_matfree_setup  needs the resident pattern of the same mesh (lat_bsr_pattern_build); `fixed` is the
 *                      uint8[6 n_nodes] Dirichlet mask (NULL: none).  Call again when radii / mask change.
 *   lat_matfree_apply  y = A u;  eliminated = 1: A = P K P + (I - P) (what the solver iterates on),
 *                      eliminated = 0: the raw stiffness K (reactions R = K u - f).
 *   lat_matfree_rhs    b = P (f - K g) + (I - P) g: the lifting of lat_apply_dirichlet (f may be NULL).
 *   lat_pcg_matfree    Chronopoulos-Gear PCG on A; opts as lat_pcg_bsr, reference_semantics must be 0
 *                      (LAT_ERR_UNSUPPORTED otherwise).  [syncs] */
int lat_matfree_setup(lat_ctx* ctx, const double* x, const double* y, const double* z, const int32_t* en0,
                      const int32_t* en1, const double* rad, int64_t n_elem, int64_t n_nodes, double young,
                      double nu, double kappa, const uint8_t* fixed);
int lat_matfree_apply(lat_ctx* ctx, const double* u, double* y, int eliminated);
int lat_matfree_rhs(lat_ctx* ctx, const double* g, const double* f, double* b);
int lat_pcg_matfree(lat_ctx* ctx, const double* b, double* x, const lat_pcg_opts* opts, lat_pcg_result* result);

/* ---- (e) multi-GPU: slab partition, NCCL over NVLink ---------------------------------------
 * The reference is single-process (MPI.COMM_SELF, utils_simulation.py:39); this is the new exchange
 * step of the sharded PCG.  One process per GPU.  Each rank holds the block rows of the nodes it
 * OWNS (n_owned, complete rows) over a local column space [owned nodes | ghost nodes]; ghost nodes are
 * ordered by owner rank (the order of `peer`), then by global node id, so a received halo lands
 * contiguously.  Per PCG iteration: one halo exchange of z (ncclSend/ncclRecv with the <= 2 slab
 * neighbours) of u = M^-1 r and ONE all-reduce of 3 doubles ((r,u), (Au,u), (r,r); Chronopoulos-Gear form). */
typedef struct {
  int32_t n_neighbors;
  int32_t pad;
  const int32_t* peer;       /* HOST  [n_neighbors] ranks */
  const int32_t* send_count; /* HOST  [n_neighbors] owned nodes sent to each peer */
  const int32_t* recv_count; /* HOST  [n_neighbors] ghost nodes received from each peer */
  const int32_t* send_idx;   /* DEVICE int32[sum send_count] local ids of the owned nodes to send, peer-major */
  int64_t n_owned, n_local;  /* nodes */
} lat_halo;

/* rank 0 creates the 128-byte NCCL id; the host layer broadcasts it (torch.distributed / MPI / file) */
int lat_nccl_unique_id(void* id128);
int lat_comm_create(lat_ctx* ctx, const void* id128, int nranks, int rank); /* collective */
int lat_comm_destroy(lat_ctx* ctx);
int lat_allreduce_sum(lat_ctx* ctx, double* buf, int64_t n);               /* in place, on the ctx stream */
int lat_halo_exchange(lat_ctx* ctx, const lat_halo* halo, double* vec);     /* vec: [6 n_local] */
/* b: [6 n_owned] (at least), x: [6 n_local] (ghost part is NOT updated: call lat_halo_exchange) [syncs] */
int lat_pcg_bsr_dist(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                     const lat_halo* halo, const double* b, double* x, const lat_pcg_opts* opts,
                     lat_pcg_result* result);
/* Matrix-free twin: the operator was set up (lat_matfree_setup) over the LOCAL mesh (n_local nodes, ghosts
 * included); products run on the owned rows, ghosts of u arrive through the same NCCL / peer-memory halo
 * (bit 4 of opts.reserved).  b: [6 n_owned] at least, x: [6 n_local].  [syncs] */
int lat_pcg_matfree_dist(lat_ctx* ctx, const lat_halo* halo, const double* b, double* x, const lat_pcg_opts* opts,
                         lat_pcg_result* result);

/* NVLink peer-memory path (no NCCL inside the iteration): each rank creates one arena (mailboxes, halo
 * flags and the ghosted vector u), the host layer all-gathers the 64-byte CUDA IPC handles, every rank
 * maps all arenas.  lat_pcg_bsr_dist with bit 4 of opts.reserved then pushes halos with st.global on the
 * mapped peer pointers and all-reduces the 3 CG sums through the mailboxes inside its own kernels
 * (halo_push inside k_cg_update, halo_wait inside k_cg_spmv<GHOST>, k_p2p_reduce); info = 4 reports a peer time-out.
 *   nb_rank[k], nb_dst_node0[k]: neighbour k (same order as lat_halo.peer) and the first node index, in
 *   THAT rank's local numbering, of the ghost segment this rank fills. */
int lat_p2p_arena_create(lat_ctx* ctx, int64_t n_local, void* handle64);
int lat_p2p_attach(lat_ctx* ctx, const void* handles /*[nranks][64]*/, int nranks, int rank, int n_neighbors,
                   const int32_t* nb_rank, const int64_t* nb_dst_node0);
int lat_p2p_destroy(lat_ctx* ctx);

/* ---- A11: compliance sensitivity ----------------------------------------------
 * g[group[e]] -= chain_e * u_e^T (dK_e/dr)(r_e) u_e   (LatticeOpti.calculate_gradient
 * compliance branch, lattice_opti.py:746-841, sign of :719 included; element form
 * of lattice_sim.py:1020-1054 with material_definition.py:163-231).
 * u: [6 n_nodes]; lambda: second vector for adjoint objectives (NULL -> lambda = u,
 * lattice_opti.py:843-902 uses lambda^T dS u); group: int32[n_elem] (<0 = skip);
 * chain: NULL -> 1.  g[n_groups] is overwritten.  q_elem (optional) receives the
 * per-element contribution. */
int lat_compliance_grad(lat_ctx* ctx, const double* x, const double* y, const double* z,
                        const int32_t* en0, const int32_t* en1, const double* rad,
                        const double* chain, const int32_t* group, int64_t n_elem,
                        double young, double nu, double kappa, const double* u,
                        const double* lambda, int64_t n_groups, double* g, double* q_elem);

/* ---- A7: batched per-cell Schur complement -----------------------------------
 * SchurComplement.calculate_schur_complement (schur_complement.py:75-147) via
 * get_schur_complement (utils_schur.py:22-53) for a batch of cells that share one
 * local mesh topology (same local connectivity, per-cell coordinates and radii):
 *   S_c = K_BB - K_BI K_II^-1 K_IB.
 * Local numbering: the first n_bnd_nodes local nodes are the boundary nodes in
 * cell.node_in_order_simulation order (cell.py:611-680); the remaining
 * n_loc_nodes - n_bnd_nodes are interior.
 *   xyz:  [n_cells][n_loc_nodes][3]   len0/len1: int32[n_loc_elem] local connectivity
 *   rad:  [n_cells][n_loc_elem]
 *   S:    [n_cells][6 n_bnd][6 n_bnd] row-major (C order like the reference ndarray)
 * drad_chain != NULL additionally writes dS/dr_g for n_grad radius groups:
 *   elem_group int32[n_loc_elem] (<0 none), drad_chain[n_loc_elem] = d rad_e / d r_group,
 *   dS: [n_cells][n_grad][6 n_bnd][6 n_bnd]  (analytic; replaces the central FD of
 *   lattice_sim.py:1020-1054). */
int lat_schur_batch(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                    const double* rad, int64_t n_cells, int32_t n_loc_nodes, int32_t n_bnd_nodes,
                    int32_t n_loc_elem, double young, double nu, double kappa, double* S,
                    const int32_t* elem_group, const double* drad_chain, int32_t n_grad, double* dS);

/* Same S_c through the strut pre-pass (no sensitivities): every chain of collinear elements between two joints
 * (the reference meshes a strut with ~18 elements, schur_complement.py / gmsh h = 0.05 cell size) is condensed
 * exactly onto its end joints first -- one thread per (cell, strut), O(elements) scalar work -- and the dense
 * condensation runs on the joint-only cell (BCC: 54 DOFs for any subdivision).
 *   chain_ptr int32[n_chains+1], chain_elem / chain_flip int32[n_loc_elem]: the elements of each chain in walking
 *   order (flip = 1: walked from its second to its first node); chain_a / chain_b int32[n_chains]: end joints in
 *   the REDUCED numbering (n_joints joints, the n_bnd_nodes boundary nodes first, in the order of S's rows).
 *   xyz / rad / len0 / len1 keep the FULL local numbering of lat_schur_batch. */
int lat_schur_batch_chains(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                           const double* rad, int64_t n_cells, int32_t n_loc_nodes, int32_t n_loc_elem,
                           const int32_t* chain_ptr, const int32_t* chain_elem, const int32_t* chain_flip,
                           const int32_t* chain_a, const int32_t* chain_b, int32_t n_chains, int32_t n_joints,
                           int32_t n_bnd_nodes, double young, double nu, double kappa, double* S);

/* Same S_c (and, optionally, the analytic sensitivities dS_c/dr_g) through the strut pre-pass with the warp-level
 * kernel for STAR cells: one interior joint (joint index n_bnd_nodes) joined to every boundary joint by exactly one strut
 * -- a BCC cell at ANY subdivision.  A half-warp owns a cell: S(k,l) = delta_kl D_k - O_k Kcc^-1 O_l^T from the 6x6 blocks of
 * the condensed struts, Kcc = sum_k C_k inverted in registers; no n x n matrix and no CTA barrier (k_schur_star).
 *   chain_group int32[n_chains]: radius group of every strut (all elements of a strut share it), drad_chain[n_loc_elem]:
 *   d rad_e / d r_group (NULL -> 1); dS: [n_cells][n_grad][6 n_bnd][6 n_bnd] or NULL.  The strut pre-pass is differentiated
 *   in forward mode, so the sensitivities no longer need lat_schur_batch's dense route over all strut-interior DOFs
 *   (replaces the central finite differences of lattice_sim.py:1020-1054).
 * Cells WITHOUT an interior joint (n_joints == n_bnd_nodes: Octet, Kelvin, Cubic, ...): nothing is left to eliminate after
 * the pre-pass, S is the assembled joint-only cell matrix and dS/dr_g the same assembly applied to the differentiated
 * pre-pass of group g's struts (k_schur_direct, one warp per cell).
 * Other topologies: S falls through to lat_schur_batch_chains; with dS != NULL the call returns LAT_ERR_UNSUPPORTED
 * (use lat_schur_batch).  [syncs: two small D2H copies of the chain ends] */
int lat_schur_batch_struts(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                           const double* rad, int64_t n_cells, int32_t n_loc_nodes, int32_t n_loc_elem,
                           const int32_t* chain_ptr, const int32_t* chain_elem, const int32_t* chain_flip,
                           const int32_t* chain_a, const int32_t* chain_b, int32_t n_chains, int32_t n_joints,
                           int32_t n_bnd_nodes, double young, double nu, double kappa, double* S,
                           const int32_t* chain_group, const double* drad_chain, int32_t n_grad, double* dS);

/* ---- joint-only global assembly (B200 design choice; exact static condensation of the struts) --------------
 * The reference applies loads and constraints to lattice points only (full_scale_lattice_simulation.py:77-153), so
 * eliminating the strut-interior nodes changes neither the joint displacements nor the joint reactions.  Every chain
 * of elements between two joints becomes one 12x12 super-element (k_chain_condense, as in lat_schur_batch_chains) and
 * the BSR matrix is assembled over the JOINT mesh: call lat_bsr_pattern_build(chain_a, chain_b, n_chains, n_joints)
 * first, then this; lat_apply_dirichlet / lat_pcg_bsr / lat_bsr_spmv work on the result unchanged.
 *   xyz: [n_nodes_full][3] row-major; len0/len1/rad: the subdivided mesh; chain_*: as lat_schur_batch_chains. */
int lat_assemble_bsr_struts(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                            const double* rad, int64_t n_nodes_full, int64_t n_elem, const int32_t* chain_ptr,
                            const int32_t* chain_elem, const int32_t* chain_flip, int64_t n_chains, int64_t n_joints,
                            double young, double nu, double kappa, double* vals);
/* Back-substitution of the joint-only solve: displacements of the strut-interior nodes from the joint displacements
 * (exact: no load on interior nodes).  u_full [6 n_nodes_full] must already hold the joints in its first 6 n_joints
 * entries (joints are the first nodes of the subdivided mesh); u_joints is a separate buffer.  One thread per strut,
 * max_chain_len <= 64 elements per strut (LAT_ERR_UNSUPPORTED otherwise). */
int lat_strut_recover(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1, const double* rad,
                      const int32_t* chain_ptr, const int32_t* chain_elem, const int32_t* chain_flip,
                      const int32_t* chain_a, const int32_t* chain_b, int64_t n_chains, int32_t max_chain_len,
                      double young, double nu, double kappa, const double* u_joints, double* u_full);

/* ---- A8: DDM interface operator ------------------------------------------------
 * y = sum_c B_c S_c B_c^T x  (LatticeSim.calculate_reaction_force_global ->
 * update_reaction_force_each_cell -> solve_sub_problem, lattice_sim.py:1180-1252,
 * cell.py:684-750).  S: [n_cells][nb][nb] (or one shared matrix when s_stride = 0),
 * gidx: int32[n_cells][nb] global free-DOF index of each local boundary DOF or -1
 * (fixed DOF: contributes u_fixed[c][k] instead of x, output dropped);
 * u_fixed may be NULL (= 0).  y[n_free] is overwritten. */
int lat_ddm_matvec(lat_ctx* ctx, const double* S, int64_t s_stride, const int32_t* gidx,
                   const double* u_fixed, int64_t n_cells, int32_t nb, int64_t n_free,
                   const double* x, double* y);

/* ---- A9: assembled interface operator -------------------------------------------------------
 * K_G = sum_c P_c^T S_c P_c as BSR over the interface (cell-boundary) nodes: what
 * LatticeSim.build_preconditioner + Cell.build_local_preconditioner (lattice_sim.py:1351-1415,
 * cell.py:783-827) assemble as COO triplets for SuperLU.  cell_nodes: int32[n_cells][n_bnd_nodes]
 * interface node index (Point.index_boundary) of each local boundary node in
 * cell.node_in_order_simulation order (-1 = absent).  The pattern (rowptr, colidx, nnzb) must contain
 * every (node, node) pair of every cell: build it with lat_bsr_pattern_build on the node pairs.
 * vals[nnzb*36] is overwritten.  FP64 atomics: summation order is not fixed.  [syncs] */
int lat_assemble_cells_bsr(lat_ctx* ctx, const double* S, int64_t s_stride, const int32_t* cell_nodes,
                           int64_t n_cells, int32_t n_bnd_nodes, const int32_t* rowptr, const int32_t* colidx,
                           int64_t nnzb, double* vals);
/* The same matrix by a gather over a PLAN, for repeated assemblies on one interface pattern (a design iteration changes S,
 * not the pattern): blk_ptr int32[nnzb+1] / contrib int64[blk_ptr[nnzb]] list for every BSR block its contributions
 * (cell * n_bnd_nodes + a) * n_bnd_nodes + b, sorted by block (the host builds it once: ddm.InterfaceProblem).  No
 * atomics: fixed summation order, bit-reproducible matrix, every entry of S read once.  vals[nnzb*36] is overwritten. */
int lat_assemble_cells_bsr_plan(lat_ctx* ctx, const double* S, int64_t s_stride, int32_t n_bnd_nodes,
                                const int32_t* blk_ptr, const int64_t* contrib, int64_t nnzb, double* vals);

/* ---- A11, cell form: q[c][j] = v_c^T dS_{m(c,j)} u_c ------------------------------------------
 * The per-(cell, geometry) term of LatticeOpti.calculate_gradient (lattice_opti.py:752-761,
 * u_cell @ (dS @ u_cell)) for all cells at once; the caller maps q to the parameter vector
 * (unit_cell / constant / linear, :758-839).  mats: [n_mats][nb][nb] row-major, the UNIQUE sensitivity
 * matrices (the reference caches one per (geometry, radii) key, lattice_sim.py:857-883); mat_index:
 * int32[n_cells][n_grad] index into mats (-1: none -> 0); U: [n_cells][nb] boundary displacements in
 * cell.node_in_order_simulation order; V: second vector for adjoint objectives (NULL -> V = U).
 * out[n_cells][n_grad] is overwritten. */
int lat_cell_quadform(lat_ctx* ctx, const double* mats, int64_t n_mats, const int32_t* mat_index, const double* U,
                      const double* V, int64_t n_cells, int32_t n_grad, int32_t nb, double* out);

/* ---- N4: reduced-basis / surrogate Schur pipeline (csrc/lattice_surrogate.cu) ------------------------------
 * All pointers are device pointers unless stated otherwise. */

/* Greedy reduced basis of Schur snapshots: reduce_basis_greedy (greedy_algorithm.py:35-155, the while loop of
 * :112-124 with its dgemv/dger deflation).  snaps: [n_snap][len], snapshot s = vec(S_s) in any fixed order (the
 * reference uses Fortran order, :99).  Out: basis [n_snap][len] capacity, row kk = kk-th basis vector
 * (reducedbasis[kk], unit 2-norm); coef [n_snap][n_snap] capacity, row kk = newcoef of step kk; mainelem
 * int32[n_snap] (s_I of each step, device); norms [n_snap] = |vec(S_s)|_2; *k_out (HOST) = basis size.  [syncs] */
int lat_greedy_basis(lat_ctx* ctx, const double* snaps, int64_t n_snap, int64_t len, double tol, double* basis,
                     double* coef, int32_t* mainelem, double* norms, int32_t* k_out);
/* Upper-triangular solve U X = R in place: dtrtrs of greedy_algorithm.py:129.  U [n][ldu], R [n][m], row-major. */
int lat_upper_solve(lat_ctx* ctx, const double* U, int32_t n, int32_t ldu, double* R, int32_t m);
/* Least-squares coefficients alphas[s][:] = argmin |B^T a - V_s| : la.lstsq(basisF, v) of greedy_algorithm.py:135-138
 * and project_to_reduced_basis (:233-266), by the normal equations of the (near-orthonormal) basis.
 * basis [k][len] (rows = basis vectors), V [n][len], alphas [n][k].  [syncs] */
int lat_basis_project(lat_ctx* ctx, const double* basis, int32_t k, int64_t len, const double* V, int64_t n,
                      double* alphas);
/* ThinPlateSplineRBF.__init__ (utils_rbf.py:22-61): assemble [[Phi + reg I, P], [P^T, 0]] and solve by LU with
 * partial pivoting.  x_train [N][d], y [N][m]; wcp [(N+d+1)][m]: rows [0,N) = W, then CP.  [syncs] */
int lat_rbf_fit(lat_ctx* ctx, const double* x_train, int32_t N, int32_t d, const double* y, int32_t m, double reg,
                double* wcp);
/* ThinPlateSplineRBF.evaluate / .gradient (utils_rbf.py:82-144) for M queries xq [M][d]:
 * f [M][m] (or NULL), grad [M][d][m] (or NULL). */
int lat_rbf_eval(lat_ctx* ctx, const double* x_train, int32_t N, int32_t d, const double* wcp, int32_t m,
                 const double* xq, int64_t M, double* f, double* grad);
/* alpha look-up of the "nearest_neighbor" (mode 0, lattice_sim.py:939-942) and "linear" (mode 1, the np.interp branch
 * of evaluate_alphas_linear_surrogate, :781-792; d = 1 only) surrogates.  alpha_train [N][m], out [M][m]. */
int lat_alpha_lookup(lat_ctx* ctx, int32_t mode, const double* x_train, int32_t N, int32_t d,
                     const double* alpha_train, int32_t m, const double* xq, int64_t M, double* out);
/* N-parameter "linear" surrogate (LinearNDInterpolator inside the convex hull of the centres, NearestNDInterpolator
 * outside: evaluate_alphas_linear_surrogate, lattice_sim.py:794-807) on a Delaunay triangulation of the centres given
 * in scipy.spatial.Delaunay's layout: simplices int32 [n_simplices][d+1], transform [n_simplices][d+1][d]. */
int lat_alpha_simplex(lat_ctx* ctx, const int32_t* simplices, const double* transform, int32_t n_simplices, int32_t d,
                      const double* x_train, int32_t N, const double* alpha_train, int32_t m, const double* xq,
                      int64_t M, double* out);
/* Basis in the layout of lat_basis_expand: basis [len][k] row-major (the npz's basis_reduced_ortho) ->
 * basisP [4*ceil(k/4)][len]; n_fortran = n applies the reference's order='F' reshape of every column
 * (lattice_sim.py:973-976): basisP[kk][a*n + b] = basis[a + n*b][kk]; 0 = keep the order. */
int lat_basis_prepare(lat_ctx* ctx, const double* basis, int64_t len, int32_t k, int32_t n_fortran, double* basisP);
/* out[q][:] = sum_kk alphas[q][kk] basisP[kk][:]  -- the GEMM of get_schur_complement_from_reduced_basis_batch
 * (lattice_sim.py:961-976) and of _compute_schur_gradients_RBF (:1075-1080) for M queries, on the FP64 tensor cores
 * (DMMA).  alphas [M][lda]; out [M][len], 32-byte aligned, len % 4 == 0. */
int lat_basis_expand(lat_ctx* ctx, const double* basisP, int32_t k, int64_t len, const double* alphas, int64_t M,
                     int32_t lda, double* out);

/* ---- two-level preconditioner: 6x6 block-Jacobi + rigid-body-mode coarse space (csrc/coarse.cuh) ----
 * Stands where the reference passes SuperLU's factorisation of the global interface matrix to its PCG as the
 * preconditioner M (lattice_sim.py:1333-1415, used by solve_DDM :1148-1160 and LatticeOpti._solve_adjoint_vector,
 * lattice_opti.py:1636-1645): M^-1 = D^-1 + Z E^+ Z^T with E = Z^T A Z, Z = the six rigid-body modes of every
 * aggregate of nodes about its centroid, constrained DOFs masked.  The coarse space is resident in the context:
 *   lat_coarse_setup       node_agg [n_nodes]: aggregate of every local node; agg_ptr [n_agg+1] / agg_nodes: the nodes
 *                          whose residual is restricted and whose u is corrected, grouped by aggregate (single GPU: every
 *                          node once; sharded: the OWNED nodes, ghosts only appear in node_agg); fixed [6 n_nodes] or
 *                          NULL; centers [3 n_agg] or NULL (NULL: centroid of the listed nodes; ranks of a sharded solve
 *                          pass common points, e.g. the box centres -- the coarse space does not depend on them).
 *                          Deactivates any inverse. [syncs]
 *   lat_coarse_galerkin    E [6 n_agg][6 n_agg] = Z^T A Z over the n_nodes block rows given (all rows; sharded: the owned
 *                          rows, E is then summed over the ranks by the caller), with or without the Dirichlet
 *                          elimination (the masked rows / columns do not take part)
 *   lat_coarse_set_inverse registers Einv (device, [6 n_agg]^2 row-major, symmetric, borrowed until replaced; the
 *                          host forms it from E with a dense library factorisation) -- from then on lat_pcg_bsr,
 *                          lat_pcg_matfree and their _dist forms on this context (same local n_nodes, precond as given,
 *                          textbook mode) add the coarse correction in every iteration and report bit 10 in
 *                          result.reserved; NULL switches it off.  The sharded solvers all-reduce the coarse residual
 *                          (6 n_agg doubles) once per iteration and use the separate halo kernel / NCCL exchange.
 *                          LAT_ERR_UNSUPPORTED with reference_semantics.
 *   lat_coarse_apply       u += Z Einv Z^T r (one application of the coarse correction on this rank's listed nodes, no
 *                          all-reduce; r, u 16-byte aligned) */
int lat_coarse_setup(lat_ctx* ctx, const double* x, const double* y, const double* z, int64_t n_nodes,
                     const int32_t* node_agg, const int32_t* agg_ptr, const int32_t* agg_nodes, int32_t n_agg,
                     const uint8_t* fixed, const double* centers);
int lat_coarse_galerkin(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                        int64_t n_nodes, double* E);
int lat_coarse_set_inverse(lat_ctx* ctx, const double* einv);
int lat_coarse_apply(lat_ctx* ctx, const double* r, double* u);

#ifdef __cplusplus
}
#endif
#endif /* LATTICE_B200_H */
