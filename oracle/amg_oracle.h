/*
 * amg_oracle.h -- CPU oracle for the AMG-PCG theta-sweep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (amg-ann_b200/, include/)
 * may include, link or load this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker or as
 * the timed CPU baseline.
 *
 * PARITY UNPINNED for the AMG part: the reference's arithmetic for this path
 * lives in hypre BoomerAMG reached through PETSc (PCHYPRE, KSPCG) reached through
 * deal.II PETScWrappers (ref common/amg_solver.h:3-8,38-39,48,54).  None of the
 * three is vendored under /root/reference or installed here, the exact versions
 * are not recorded in the reference (only "deal.II >= 9.3.1" and the mk-2024.0
 * module bundle, ref environment/postInstall:4-12), and the reference ships no
 * tests, golden vectors or stored outputs (SURVEY.md sections 4 and 8c).  This
 * file therefore restates the published algorithms of those libraries as
 * documented in SURVEY.md Appendix A; it is a restatement, not hypre.
 * The pooling part (orc_make_view) IS pinned by source: it follows
 * ref common/view_maker.h:26-74 literally.
 *
 * Single thread, no dependencies, compiled -O2 -ffp-contract=off so that every
 * multiply-add is two roundings (the CUDA path uses __dmul_rn/__dadd_rn in the
 * parity-critical setup kernels for the same reason).
 */
#ifndef AMG_ORACLE_H
#define AMG_ORACLE_H

#include <stdint.h>

#include "amgb.h" /* amgb_boomeramg_data: the parameter pack is shared */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_hier orc_hier;

/* --- stage functions (unit parity) --- */
/* hypre_BoomerAMGCreateS, Appendix A.3 "Strength". mask[k]=1 iff entry k is a strong connection. */
int orc_strength(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
                 double theta, double max_row_sum, uint8_t* mask);
/* hypre_Rand sequence value for index i (seed 2747, multiplicative LCG 16807 mod 2^31-1). */
double orc_hypre_rand(int64_t i);
/* hypre_BoomerAMGCoarsenPMIS (coarsen type 8) on one rank. cf in {+1,-1,-3}. */
int orc_coarsen_pmis(int64_t n, const int32_t* rowptr, const int32_t* col,
                     const uint8_t* mask, int32_t* cf);
/* hypre_BoomerAMGCoarsenFalgout (coarsen type 6) on one rank: Ruge first+second
 * pass, then CLJP with CF_init=1. */
int orc_coarsen_cljp(int64_t n, const int32_t* rowptr, const int32_t* col,
                     const uint8_t* mask, int32_t* cf);
int orc_coarsen_falgout(int64_t n, const int32_t* rowptr, const int32_t* col,
                        const uint8_t* mask, int32_t* cf);

/* --- hierarchy --- */
int orc_setup(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
              const amgb_boomeramg_data* data, orc_hier** out);
void orc_destroy(orc_hier* h);
int orc_num_levels(const orc_hier* h);
int orc_level_dims(const orc_hier* h, int level, int64_t* n, int64_t* nnz_A,
                   int64_t* n_coarse, int64_t* nnz_P);
int orc_get_strength_mask(const orc_hier* h, int level, uint8_t* mask);
int orc_get_cf_marker(const orc_hier* h, int level, int32_t* cf);
int orc_get_A_csr(const orc_hier* h, int level, int32_t* rowptr, int32_t* col, double* val);
int orc_get_P_csr(const orc_hier* h, int level, int32_t* rowptr, int32_t* col, double* val);
int orc_level_stats(const orc_hier* h, int capacity, int32_t* n_levels, int64_t* rows,
                    int64_t* nnz, double* sparsity, double* grid_cx, double* op_cx,
                    double* mem_cx);
int orc_effective_relax(const orc_hier* h, int32_t* down, int32_t* up, int32_t* coarse);
/* multicolour Gauss-Seidel (smoother_policy = AMGB_SMOOTHER_MULTICOLOR): colours of a level */
int orc_get_colors(const orc_hier* h, int level, int32_t* colors, int32_t* n_colors);
/* Chebyshev smoother (hypre relax type 16, par_cheby.c) of a level: CG/Lanczos spectrum
 * estimates of D^-1/2 A D^-1/2 and the polynomial coefficients (n_coefs = order). */
int orc_level_cheby(const orc_hier* h, int level, double* max_eig, double* min_eig, double* coefs,
                    int32_t* n_coefs);
/* EISPACK tql1 as hypre_LINPACKcgtql1 takes its arguments (offdiag[1..n) used). */
int orc_tql1(int32_t n, double* diag, double* offdiag);

/* z = one V-cycle applied to r from a zero initial guess (PCApply_HYPRE). */
int orc_vmult(orc_hier* h, double* z, const double* r);
/* PETSc KSPCG as driven by deal.II (Appendix A.4). */
int orc_cg_solve(orc_hier* h, int64_t n, const int32_t* rowptr, const int32_t* col,
                 const double* val, double* x, const double* b, int64_t max_steps,
                 double abs_tol, double* res_hist, int64_t hist_cap, int64_t* n_iters);

int orc_spmv(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
             const double* x, double* y);

/* ref common/view_maker.h:26-74, literal. */
int orc_make_view(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
                  int32_t view_size, double* sum, int64_t* count, double* max_pp,
                  double* max_np);

/* deal.II forwards theta / max_row_sum to PETSc through std::to_string (A.1, H4). */
double orc_option_roundtrip(double v);

#ifdef __cplusplus
}
#endif
#endif
