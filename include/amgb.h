/*
 * amgb.h -- C ABI of the B200-native AMG-PCG theta-sweep hot path.
 *
 * This is the drop-in boundary for the one path of MatteoCaldana/AMG-ANN that
 * this library replaces.  Every entry point cites the
 * reference call site it stands in for.  Reference paths are relative to
 * code/data-generation/ of the reference tree.
 *
 *   reference call                                              replaced by
 *   ----------------------------------------------------------  ---------------------------
 *   PETScWrappers::MPI::SparseMatrix (PETSc AIJ, 1 rank)        amgb_matrix_upload_csr
 *     built at testcase2-.../src/main.cpp:239-249
 *   PreconditionBoomerAMG::AdditionalData                       amgb_boomeramg_data
 *     common/amg_solver.h:20, t2 main.cpp:447-453
 *   preconditioner.initialize(A, data)  common/amg_solver.h:48  amgb_precond_initialize
 *   preconditioner.vmult (PCApply in CG) common/amg_solver.h:54 amgb_precond_vmult
 *   SolverControl(n, tol) + SolverCG::solve  amg_solver.h:33,54 amgb_cg_solve
 *   hypre setup printout scraped by BoomerAMGParser             amgb_precond_level_stats
 *     common/parser.h:181-266, amg_solver.h:71-78
 *   -ksp_monitor residuals scraped by PETScOutputParser         res_hist of amgb_cg_solve
 *     common/parser.h:149-155, amg_solver.h:83-86
 *   ViewMaker::make_view (MatGetRow loop) common/view_maker.h:26-74   amgb_make_view
 *
 * Conventions: plain pointers and sizes only; pointers are HOST pointers unless
 * the name says _device; every function returns AMGB_OK (0) or a negative
 * amgb_status; no exceptions cross the ABI; no hidden global state (everything
 * hangs off an amgb_ctx; one ctx per host thread); there is NO CPU fallback:
 * without a usable CUDA device amgb_ctx_create fails with AMGB_ERR_NO_DEVICE.
 */
#ifndef AMGB_H
#define AMGB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMGB_VERSION 100

typedef enum amgb_status {
  AMGB_OK = 0,
  AMGB_ERR_BAD_ARG = -1,
  AMGB_ERR_NO_DEVICE = -2,      /* no CUDA device / driver: the library has no CPU path */
  AMGB_ERR_CUDA = -3,           /* a CUDA runtime call failed; see amgb_last_error */
  AMGB_ERR_OOM = -4,
  AMGB_ERR_UNSUPPORTED = -5,    /* option exists in the reference API but not on device */
  AMGB_ERR_NO_CONVERGENCE = -6, /* SolverControl::NoConvergence equivalent */
  AMGB_ERR_BREAKDOWN = -7,      /* zero pivot / NaN in the iteration */
  AMGB_ERR_RANGE = -8,          /* index out of range (level, buffer capacity) */
  AMGB_ERR_COMM = -9            /* NCCL failure */
} amgb_status;

/* deal.II PETScWrappers::PreconditionBoomerAMG::AdditionalData::RelaxationType,
 * same order as the deal.II enum. */
typedef enum amgb_relaxation_type {
  AMGB_RELAX_Jacobi = 0,
  AMGB_RELAX_sequentialGaussSeidel = 1,
  AMGB_RELAX_seqboundaryGaussSeidel = 2,
  AMGB_RELAX_SORJacobi = 3,
  AMGB_RELAX_backwardSORJacobi = 4,
  AMGB_RELAX_symmetricSORJacobi = 5,
  AMGB_RELAX_l1scaledSORJacobi = 6,
  AMGB_RELAX_GaussianElimination = 7,
  AMGB_RELAX_l1GaussSeidel = 8,
  AMGB_RELAX_backwardl1GaussSeidel = 9,
  AMGB_RELAX_CG = 10,
  AMGB_RELAX_Chebyshev = 11,
  AMGB_RELAX_FCFJacobi = 12,
  AMGB_RELAX_l1scaledJacobi = 13,
  AMGB_RELAX_None = 14
} amgb_relaxation_type;

/* hypre coarsen_type / interp_type numbers (PETSc -pc_hypre_boomeramg_coarsen_type). */
enum { AMGB_COARSEN_CLJP = 0, AMGB_COARSEN_FALGOUT = 6, AMGB_COARSEN_PMIS = 8 };
enum { AMGB_INTERP_CLASSICAL = 0, AMGB_INTERP_DIRECT = 3, AMGB_INTERP_EXT_I = 6 };

/* What to do when the requested smoother is inherently sequential on one rank
 * (the reference default: hybrid symmetric Gauss-Seidel, hypre relax type 6). */
enum {
  AMGB_SMOOTHER_SUBSTITUTE = 0, /* run C/F-ordered l1-scaled Jacobi instead (hypre's own GPU choice) */
  AMGB_SMOOTHER_STRICT = 1,     /* return AMGB_ERR_UNSUPPORTED */
  AMGB_SMOOTHER_MULTICOLOR = 2  /* Gauss-Seidel types 3 / 4 / 6 (forward / backward / symmetric): run the
                                   same sweep in MULTICOLOUR order -- a true Gauss-Seidel sweep whose
                                   unknowns are ordered by the colours of a greedy colouring of the
                                   level's graph (colour by colour, every colour in parallel) instead of
                                   by index; reported as 103 / 104 / 106 by amgb_precond_effective_relax.
                                   The other sequential types (l1 variants, 1, 2) are substituted as under
                                   AMGB_SMOOTHER_SUBSTITUTE.  Needs a structurally symmetric operator. */
};
/* relax type codes reported for the multicolour sweeps (outside hypre's numbering) */
enum { AMGB_RELAX_MC_FORWARD = 103, AMGB_RELAX_MC_BACKWARD = 104, AMGB_RELAX_MC_SYMMETRIC = 106 };

/*
 * Parameter pack of initialize().  The first twelve fields are deal.II's
 * AdditionalData field for field, in constructor order (the reference passes the
 * first five: t2 main.cpp:447-453, t3 main.cpp:458-464).  The rest are the PETSc
 * PCHYPRE defaults that deal.II leaves implicit, made
 * explicit so they can be pinned in tests.  Fill with amgb_boomeramg_data_default.
 */
typedef struct amgb_boomeramg_data {
  int32_t symmetric_operator;               /* default 0 (reference always passes 1) */
  double strong_threshold;                  /* theta, default 0.25 */
  double max_row_sum;                       /* default 0.9 */
  uint32_t aggressive_coarsening_num_levels;/* default 0; >0: second PMIS on S2 + multipass interp. */
  int32_t output_details;                   /* collect level statistics */
  int32_t relaxation_type_up;               /* amgb_relaxation_type, default SORJacobi */
  int32_t relaxation_type_down;
  int32_t relaxation_type_coarse;           /* default GaussianElimination */
  uint32_t n_sweeps_coarse;                 /* default 1 */
  double tol;                               /* default 0.0 */
  uint32_t max_iter;                        /* default 1 */
  int32_t w_cycle;                          /* default 0 */
  /* ---- explicit PCHYPRE/hypre knobs ---- */
  int32_t coarsen_type;       /* AMGB_COARSEN_*; device default PMIS (Falgout is serial) */
  int32_t interp_type;        /* AMGB_INTERP_*; default classical modified */
  int32_t relax_order;        /* 1 = C/F relaxation (PCHYPRE default) */
  uint32_t n_sweeps;          /* grid sweeps down and up, default 1 */
  int32_t max_levels;         /* default 25 */
  int32_t max_coarse_size;    /* default 9 */
  double relax_weight;        /* default 1.0 (Jacobi-type smoothers) */
  int32_t smoother_policy;    /* AMGB_SMOOTHER_SUBSTITUTE / _STRICT */
  int32_t options_via_string; /* 1: theta and max_row_sum are rounded through
                                 std::to_string (6 decimals) as deal.II forwards
                                 them to PETSc (Appendix A.1, hard part H4) */
  int32_t keep_setup_intermediates; /* 1: keep strength masks etc. for the parity accessors */
  int32_t dist_replicate_below;     /* row-partitioned path: levels (other than the finest) with at
                                       most this many rows are gathered and handled redundantly on
                                       every rank, without further exchanges (default 262144; 0: never) */
  int32_t reserved[6];
} amgb_boomeramg_data;

typedef struct amgb_ctx amgb_ctx;
typedef struct amgb_matrix amgb_matrix;
typedef struct amgb_precond amgb_precond;

/* ---- context ------------------------------------------------------------ */
/* device_id: CUDA ordinal.  stream: a cudaStream_t to enqueue on (e.g. the
 * caller's current stream), or NULL to let the context create its own. */
int amgb_ctx_create(amgb_ctx** out, int device_id, void* stream);
int amgb_ctx_destroy(amgb_ctx* ctx);
int amgb_ctx_synchronize(amgb_ctx* ctx);
/* Grow the context's memory pool to `bytes` now (one allocation, freed into the pool), so that
 * initialize() does not grow it step by step: growing a pool stalls the whole device, which is what
 * made sweeps with several contexts on one GPU erratic (4.5 ... 7.4 s for a config-2 sweep instead of
 * 4.53 s).  A hierarchy needs about 100 bytes per matrix entry (setup intermediates included).
 * Optional: the first amgb_precond_initialize() of a context reserves that much by itself when it is
 * less than a third of the free device memory (AMGB_NO_AUTO_RESERVE=1 turns that off). */
int amgb_ctx_reserve(amgb_ctx* ctx, int64_t bytes);
/* Last error text of this context (never NULL). */
const char* amgb_last_error(const amgb_ctx* ctx);
const char* amgb_status_string(int status);
int amgb_version(void);
/* Number of kernels of this library launched on ctx since creation / last reset. */
int amgb_ctx_kernel_launches(const amgb_ctx* ctx, int64_t* count);
int amgb_ctx_reset_kernel_launches(amgb_ctx* ctx);

/* ---- matrix (resident across the theta sweep) ---------------------------- */
/* CSR with ascending column ids per row and a stored diagonal in every row
 * (PETSc AIJ as deal.II builds it).  rowptr has n+1 entries.  The 32-bit variant
 * matches PetscInt of a default PETSc build.  The upload is complete when the call
 * returns and the matrix is read-only afterwards: preconditioners and solves of SEVERAL
 * contexts (streams, host threads) of the same device may use it at the same time, which
 * is how the independent systems of a theta sweep are kept in flight together. */
int amgb_matrix_upload_csr(amgb_ctx* ctx, int64_t n, const int32_t* rowptr,
                           const int32_t* col, const double* val, amgb_matrix** out);
int amgb_matrix_upload_csr64(amgb_ctx* ctx, int64_t n, const int64_t* rowptr,
                             const int32_t* col, const double* val, amgb_matrix** out);
/* Wrap CSR arrays that already live in device memory (no copy; caller keeps ownership). */
int amgb_matrix_wrap_device_csr(amgb_ctx* ctx, int64_t n, int64_t nnz,
                                const int32_t* rowptr_device, const int32_t* col_device,
                                const double* val_device, amgb_matrix** out);
/* On-device assembly of the Q1 diffusion system of the reference's structured test case
 * (ref testcase2-diffusion-structured/src/main.cpp:255-320; mu = 10^epsv on a
 * pattern_size^mode pattern, :101-113; Dirichlet rows as :312-318; lexicographic node
 * numbering; m cells per direction on [-1,1]^3).  Bit-identical to the host generator of
 * include/amgb_gen.h, but the matrix is born in HBM.  rhs_device /
 * x0_device: device arrays of n = (m+1)^3 doubles, or NULL. */
int amgb_matrix_assemble_poisson_q1(amgb_ctx* ctx, int32_t m, int32_t pattern_size, int32_t mode,
                                    const double* epsv, int64_t n_epsv, amgb_matrix** out,
                                    double* rhs_device, double* x0_device);
/* On-device assembly of the Q1 vector-elasticity system of the reference's testcase 3 (3 DoFs per node,
 * interleaved; ref testcase3-elasticity-structured/src/main.cpp:320-342; E = 1000 * young[pattern cell],
 * nu = 0.29, :48-49,88-99; constrained DoFs condensed out, :264-268; lexicographic node numbering).
 * Matrix and initial guess are bit-identical to the host generator of include/amgb_gen.h; the body
 * force is evaluated with the device's sin/cos, so the right-hand side agrees to rounding.
 * rhs_device / x0_device: device arrays of n = 3 (m+1)^3 doubles, or NULL. */
int amgb_matrix_assemble_elasticity_q1(amgb_ctx* ctx, int32_t m, int32_t pattern_size, int32_t mode,
                                       const double* young, int64_t n_young, amgb_matrix** out,
                                       double* rhs_device, double* x0_device);
/* Same, returning the right-hand side and the initial guess in HOST arrays (n doubles, or NULL). */
int amgb_matrix_assemble_poisson_q1_hostvec(amgb_ctx* ctx, int32_t m, int32_t pattern_size, int32_t mode,
                                            const double* epsv, int64_t n_epsv, amgb_matrix** out,
                                            double* rhs_host, double* x0_host);
/* CSR of a resident matrix back to host arrays (any pointer may be NULL). */
int amgb_matrix_download_csr(const amgb_matrix* A, int32_t* rowptr, int32_t* col, double* val);
/* ---- DoF numberings of the step before the path ---- */
/* The numbering deal.II's DoFHandler::distribute_dofs gives Q1 DoFs on
 * GridGenerator::subdivided_hyper_cube(coarse_cells) refined `refinements` times (ref testcase2-
 * diffusion-structured/src/main.cpp:423-425,230-232): active cells in refinement-tree order, the 8
 * vertices of a cell in lexicographic order, a vertex numbered when first met.  Restated from memory of
 * deal.II; computed on the device in closed form.  new_to_lex: HOST array of (m+1)^3 entries,
 * m = coarse_cells << refinements; new_to_lex[k] = lexicographic id of the node that DoF k sits on. */
int amgb_numbering_dealii_q1(amgb_ctx* ctx, int32_t coarse_cells, int32_t refinements, int32_t* new_to_lex);
/* B = Q A Q^T: rows and columns renumbered (new_to_old: HOST array, a permutation of 0..n-1; anything
 * else is AMGB_ERR_BAD_ARG), columns of every row ascending again.  Applies any DoF renumbering
 * (the one above, Cuthill-McKee from include/amgb_gen.h, ...) to a resident matrix on the device. */
int amgb_matrix_permute(amgb_ctx* ctx, const amgb_matrix* A, const int32_t* new_to_old, amgb_matrix** out);
int amgb_matrix_destroy(amgb_matrix* A);
int amgb_matrix_dims(const amgb_matrix* A, int64_t* n, int64_t* nnz);
/* y = A x (host vectors); the plain SpMV, exposed for parity tests. */
int amgb_matrix_vmult(amgb_ctx* ctx, const amgb_matrix* A, double* y, const double* x);

/* ---- preconditioner ------------------------------------------------------ */
int amgb_boomeramg_data_default(amgb_boomeramg_data* data);
/* ref common/amg_solver.h:48.  Builds the whole hierarchy on the device. */
int amgb_precond_initialize(amgb_ctx* ctx, const amgb_matrix* A,
                            const amgb_boomeramg_data* data, amgb_precond** out);
int amgb_precond_destroy(amgb_precond* P);
/* dst = M^{-1} src: one V(1,1) cycle from a zero initial guess (PCApply_HYPRE). */
int amgb_precond_vmult(amgb_precond* P, double* dst, const double* src);
int amgb_precond_vmult_device(amgb_precond* P, double* dst_device, const double* src_device);

/* Level statistics: what BoomerAMGParser scrapes (common/parser.h:248-256).
 * Arrays need capacity >= max_levels. */
int amgb_precond_num_levels(const amgb_precond* P, int32_t* n_levels);
int amgb_precond_level_stats(const amgb_precond* P, int32_t capacity, int32_t* n_levels,
                             int64_t* rows, int64_t* nnz, double* sparsity,
                             double* grid_complexity, double* operator_complexity,
                             double* memory_complexity);
/* The remaining columns of hypre's "Operator Matrix Information" table (par_stats.c): entries
 * per row and row sums of one level (computed on request; single-device hierarchies). */
int amgb_precond_level_row_stats(const amgb_precond* P, int32_t level, int32_t* min_entries,
                                 int32_t* max_entries, double* min_row_sum, double* max_row_sum);
/* Chebyshev smoother (RelaxationType::Chebyshev, hypre relax type 16) of a level: the
 * CG/Lanczos estimates of the extreme eigenvalues of D^-1/2 A D^-1/2 and the coefficients
 * of the polynomial (coefs needs room for 4; n_coefs = hypre's cheby_order, default 2). */
int amgb_precond_level_cheby(const amgb_precond* P, int32_t level, double* max_eig, double* min_eig,
                             double* coefs, int32_t* n_coefs);
/* Multicolour Gauss-Seidel (AMGB_SMOOTHER_MULTICOLOR): colour of every point of a level (original
 * numbering, 0 .. n_colors-1), or AMGB_ERR_RANGE when the level has no colouring. */
int amgb_precond_get_colors(const amgb_precond* P, int32_t level, int32_t* colors, int32_t* n_colors);
/* hypre relax type actually run on the device for down/up/coarse (after the
 * smoother policy was applied). */
int amgb_precond_effective_relax(const amgb_precond* P, int32_t* down, int32_t* up,
                                 int32_t* coarse);

/* Parity accessors (integer outputs must be bit-exact vs the oracle).
 * level 0 is the finest.  Query sizes with amgb_precond_level_dims first. */
int amgb_precond_level_dims(const amgb_precond* P, int32_t level, int64_t* n,
                            int64_t* nnz_A, int64_t* n_coarse, int64_t* nnz_P);
int amgb_precond_get_strength_mask(const amgb_precond* P, int32_t level, uint8_t* mask /*nnz_A*/);
int amgb_precond_get_cf_marker(const amgb_precond* P, int32_t level, int32_t* cf /*n*/);
int amgb_precond_get_A_csr(const amgb_precond* P, int32_t level, int32_t* rowptr,
                           int32_t* col, double* val);
int amgb_precond_get_P_csr(const amgb_precond* P, int32_t level, int32_t* rowptr,
                           int32_t* col, double* val);

/* ---- PCG ----------------------------------------------------------------- */
/* ref common/amg_solver.h:33,38,54 and SURVEY.md Appendix A.4: PETSc KSPCG, left
 * preconditioning, stops when the PRECONDITIONED residual norm ||M^{-1} r||_2 is
 * <= abs_tol (absolute!), initial guess x honoured.  res_hist receives
 * min(hist_cap, n_iters+1) norms (entry 0 = before the first step).  Returns
 * AMGB_ERR_NO_CONVERGENCE if max_steps is hit (x and the history are still
 * written). */
int amgb_cg_solve(amgb_ctx* ctx, const amgb_matrix* A, double* x, const double* b,
                  amgb_precond* P, int64_t max_steps, double abs_tol, double* res_hist,
                  int64_t hist_cap, int64_t* n_iters);
/* Same with x and b already on the device. */
int amgb_cg_solve_device(amgb_ctx* ctx, const amgb_matrix* A, double* x_device,
                         const double* b_device, amgb_precond* P, int64_t max_steps,
                         double abs_tol, double* res_hist, int64_t hist_cap,
                         int64_t* n_iters);

/* ---- pooling ------------------------------------------------------------- */
/* ref common/view_maker.h:26-74.  Outputs are V*V row-major (V*bin_row+bin_col);
 * count is integer-exact, max_pp/max_np are order-independent hence exact,
 * sum is accumulated in a fixed tree over the CSR order.  t_us (may be NULL)
 * receives the device time of the pass in microseconds. */
int amgb_make_view(amgb_ctx* ctx, const amgb_matrix* A, int32_t view_size, double* sum,
                   int64_t* count, double* max_pp, double* max_np, double* t_us);

/* ANN-ready image: pooling followed, still on the device, by the per-channel normalisation
 * of ref code/data-modeling/train_ann.py:133-172 (norm_view) and the channels-last stacking
 * of its "sum+max+c" view type (:247-256).  out: V*V*4 doubles, [bin][channel] with channels
 * (sum, max_pp, max_np, count).  count_channel_as_reference != 0 reproduces the reference's
 * own count channel, which normalize_view_df (:189-193) computes from the max_np view (its
 * loop variable is left at the last view type); 0 normalises the true counts.
 * An all-zero channel gives NaN under the scaled modes, as numpy does (the reference then
 * rejects the image, :200-209). */
enum {
  AMGB_NORM_NOTHING = 0,  /* "nothing"  */
  AMGB_NORM_PURE = 1,     /* "pure":     x / max|x|                        */
  AMGB_NORM_RESC = 2,     /* "resc":     (x / count) / max|.|              */
  AMGB_NORM_PURE_LOG = 3, /* "pure_log": sign(x) log(|x|+1) / max|.|      */
  AMGB_NORM_RESC_LOG = 4, /* "resc_log"                                    */
  AMGB_NORM_MEAN = 5      /* "mean":     x / count                         */
};
int amgb_make_view_normalized(amgb_ctx* ctx, const amgb_matrix* A, int32_t view_size, int32_t mode,
                              int32_t count_channel_as_reference, double* out, double* t_us);

/* ---- row-partitioned (multi-GPU) path ------------------------------------- */
/* For systems that do not fit one device (BASELINE config 5: >= 100 M DoFs, nnz > 2^31) the
 * matrix is partitioned by contiguous global row ranges, one rank per GPU, like a PETSc
 * MPIAIJ matrix (the reference itself runs on one rank only: t2 main.cpp:221-224,525).
 * Every function below is COLLECTIVE: all ranks of the communicator call it in the same
 * order.  Integer outputs (strength masks, C/F markers, patterns) and the operator values
 * are identical to the single-device ones for any number of ranks; residual histories
 * agree to rounding (the order of the dot-product reduction changes).
 *
 * Communicators: NCCL (one process per GPU; the 128-byte unique id is created on rank 0
 * with amgb_nccl_unique_id and distributed by the caller, e.g. through torch.distributed
 * or MPI_Bcast), or an in-process group whose ranks are host threads (one amgb_ctx each, on
 * the same or on different devices). */
typedef struct amgb_comm amgb_comm;
typedef struct amgb_local_group amgb_local_group;
typedef struct amgb_dist_matrix amgb_dist_matrix;
#define AMGB_NCCL_UNIQUE_ID_BYTES 128
int amgb_nccl_unique_id(void* out, int capacity);
int amgb_comm_create_nccl(amgb_ctx* ctx, int nranks, int rank, const void* unique_id, amgb_comm** out);
int amgb_local_group_create(int nranks, amgb_local_group** out);
int amgb_local_group_destroy(amgb_local_group* g);
/* A rank that fails calls this so that the other ranks' collectives return AMGB_ERR_COMM
 * instead of waiting for it. */
int amgb_local_group_abort(amgb_local_group* g);
int amgb_comm_create_local(amgb_local_group* g, int rank, amgb_comm** out);
int amgb_comm_destroy(amgb_comm* c);
int amgb_comm_rank(const amgb_comm* c);
int amgb_comm_size(const amgb_comm* c);

/* Owned rows [row_begin,row_end) of the global n_global x n_global matrix: rowptr_local has
 * row_end-row_begin+1 entries starting at 0, col_global holds GLOBAL column ids (ascending
 * per row, diagonal stored).  Ranges must ascend with the rank and tile [0,n_global).
 * Host pointers.  n_global < 2^31; the local nnz must be < 2^31 (the global nnz need not). */
int amgb_dist_matrix_create(amgb_ctx* ctx, amgb_comm* comm, int64_t n_global, int64_t row_begin,
                            int64_t row_end, const int64_t* rowptr_local, const int32_t* col_global,
                            const double* val, amgb_dist_matrix** out);
/* The slab [row_begin,row_end) of the same system assembled on this rank's device
 * (rhs_device / x0_device: row_end-row_begin doubles, or NULL).  Collective. */
int amgb_dist_matrix_assemble_poisson_q1(amgb_ctx* ctx, amgb_comm* comm, int32_t m, int32_t pattern_size,
                                         int32_t mode, const double* epsv, int64_t n_epsv,
                                         int64_t row_begin, int64_t row_end, amgb_dist_matrix** out,
                                         double* rhs_device, double* x0_device);
int amgb_dist_matrix_destroy(amgb_dist_matrix* A);
/* Pooled image of the partitioned matrix (ref common/view_maker.h:26-74 run on an MPI matrix: every
 * rank bins the rows it owns, :41-65): local pass on every rank's device, then the four channels are
 * combined over the ranks in rank order (sum and count added, maxima maxed).  Collective; every rank
 * receives the whole V x V image.  count / max_pp / max_np are identical to amgb_make_view of the
 * assembled matrix, sum agrees to rounding. */
int amgb_dist_make_view(amgb_ctx* ctx, const amgb_dist_matrix* A, int32_t view_size, double* sum,
                        int64_t* count, double* max_pp, double* max_np, double* t_us);
/* initialize() on the partitioned matrix; the result is used with the amgb_precond_*
 * queries (level statistics are global) and destroyed with amgb_precond_destroy. */
int amgb_dist_precond_initialize(amgb_ctx* ctx, const amgb_dist_matrix* A,
                                 const amgb_boomeramg_data* data, amgb_precond** out);
/* cg.solve() on the partitioned system: x and b are the owned slabs.  SpMV halo exchange
 * and the dot-product all-reduce run on the communicator. */
int amgb_dist_cg_solve(amgb_ctx* ctx, double* x_local, const double* b_local, amgb_precond* P,
                       int64_t max_steps, double abs_tol, double* res_hist, int64_t hist_cap,
                       int64_t* n_iters);
int amgb_dist_cg_solve_device(amgb_ctx* ctx, double* x_local_device, const double* b_local_device,
                              amgb_precond* P, int64_t max_steps, double abs_tol, double* res_hist,
                              int64_t hist_cap, int64_t* n_iters);
/* Parity accessors: the OWNED part of a level with global ids (not collective). */
int amgb_dist_precond_level_dims(const amgb_precond* P, int32_t level, int64_t* n_global,
                                 int64_t* row_begin, int64_t* n_local, int64_t* nnz_local,
                                 int64_t* n_coarse_global, int64_t* coarse_begin, int64_t* nnz_P_local);
/* First level that is replicated on every rank (= number of levels if none): from there on
 * the single-device accessors (amgb_precond_get_cf_marker, _get_A_csr, _get_P_csr) apply. */
int amgb_dist_precond_replicated_from(const amgb_precond* P, int32_t* level);
int amgb_dist_precond_get_cf_marker(const amgb_precond* P, int32_t level, int32_t* cf_local);
int amgb_dist_precond_get_A_rows(const amgb_precond* P, int32_t level, int32_t* rowptr_local,
                                 int32_t* col_global, double* val);
int amgb_dist_precond_get_P_rows(const amgb_precond* P, int32_t level, int32_t* rowptr_local,
                                 int32_t* col_global, double* val);

/* ---- measurement hooks (bench.py) ---------------------------------------- */
/* Per-kernel-family device time accumulated with CUDA events on the context's
 * stream while profiling is enabled.  Families: see amgb_timer_name. */
int amgb_ctx_enable_timers(amgb_ctx* ctx, int enable);
int amgb_ctx_reset_timers(amgb_ctx* ctx);
int amgb_timer_count(void);
const char* amgb_timer_name(int family);
int amgb_ctx_get_timer(amgb_ctx* ctx, int family, double* total_ms, int64_t* launches,
                       double* algorithmic_bytes);
/* Same, restricted to the launches made for one multigrid level (0 = finest; only
 * launches timed while the timers were enabled are counted here). */
int amgb_ctx_get_timer_level(amgb_ctx* ctx, int family, int level, double* total_ms,
                             int64_t* launches, double* algorithmic_bytes);

/* Code routes taken by the host side of a context since creation / last reset (which SELL
 * layout, which SpGEMM table tier or overflow stage, captured cycle, fused coarse tail ...):
 * lets the parity tests at the benchmarked sizes assert that the routes which carry the
 * benchmark were the ones compared with the oracle.  Names: amgb_route_name. */
int amgb_route_count(void);
const char* amgb_route_name(int route);
int amgb_ctx_get_route(const amgb_ctx* ctx, int route, int64_t* count);
int amgb_ctx_reset_routes(amgb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* AMGB_H */
