/*
 * amgb_gen.h -- synthetic finite-element system generators (host only, no CUDA).
 *
 * These stand in for the reference's deal.II drivers, which cannot be built here
 * (deal.II/PETSc/hypre are not vendored).  They produce the
 * CSR system the hot path consumes, with the matrix conventions of the reference:
 *
 *  - Poisson / discontinuous diffusion, Q1 hexahedra on [-1,1]^3:
 *      ref: code/data-generation/testcase2-diffusion-structured/src/main.cpp:255-320
 *      (assembly), :101-113 (piecewise-constant mu = 10^eps on a ps^mode pattern),
 *      :312-318 (Dirichlet rows via apply_boundary_values(..., false): the row is
 *      zeroed keeping its pattern, diagonal = |first non-zero diagonal entry|,
 *      columns are NOT eliminated), :239-249 (every pattern entry is stored,
 *      explicit zeros included).
 *  - Q1 vector elasticity (3 DoF per node, node-major interleaved):
 *      ref: code/data-generation/testcase3-linear-elasticity/src/main.cpp:320-342
 *      (cell matrix), :88-109 (mu, lambda from the Young-modulus pattern),
 *      :264-272 (Dirichlet values condensed through AffineConstraints with
 *      keep_constrained_dofs=false: constrained rows/columns reduce to a diagonal).
 *
 * Numbering is lexicographic: node (ix,iy,iz) -> ix + (m+1)*(iy + (m+1)*iz).
 * All functions return 0 on success, a negative amgb status on error.
 */
#ifndef AMGB_GEN_H
#define AMGB_GEN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* n = c*(m+1)^3 rows, nnz stored entries, for an m x m x m cell mesh with c DoF/node.
 * kind 0 = Poisson (pattern = full 27-point, nnz = (3m+1)^3),
 * kind 1 = elasticity with constrained rows reduced to a diagonal. */
int amgb_gen_sizes(int kind, int m, int64_t* n, int64_t* nnz);

/* Row range variant: sizes of rows [row_begin,row_end) (z-slab partitions for
 * the row-partitioned multi-GPU path). */
int amgb_gen_poisson_q1_range_sizes(int m, int64_t row_begin, int64_t row_end,
                                    int64_t* nnz);

/* Poisson / diffusion.  epsv has pattern_size^mode entries (mode in 1..3); mu of a
 * cell = 10^epsv[pattern index of the cell centre].  rowptr has (row_end-row_begin+1)
 * entries and is local (starts at 0); col holds GLOBAL column ids.
 * rhs/x0 may be NULL.  x0 = 0 on interior rows and the boundary value on
 * Dirichlet rows (ref t2 main.cpp:319, :446). */
int amgb_gen_poisson_q1(int m, int pattern_size, int mode, const double* epsv,
                        int64_t n_epsv, int64_t row_begin, int64_t row_end,
                        int64_t* rowptr, int32_t* col, double* val, double* rhs,
                        double* x0);

/* Elasticity.  young[pattern_size^mode] is the per-pattern-cell factor 10^e
 * (ref t3 main.cpp:200-209); mu = 1000*young/(1+nu), lambda = mu*nu/(1-2nu),
 * nu = 0.29 (ref t3 main.cpp:48-49).  Whole matrix only (nnz < 2^31). */
int amgb_gen_elasticity_q1(int m, int pattern_size, int mode, const double* young,
                           int64_t n_young, int64_t* rowptr, int32_t* col,
                           double* val, double* rhs, double* x0);

/* std::default_random_engine + uniform_real_distribution(0,max), as
 * ref common/myutils.h:47-54 (libstdc++: minstd_rand0, generate_canonical). */
int amgb_gen_random_vec(int64_t seed, int64_t len, double max, double* out);

/* ref t2 datagen.py:17-27 build_epsv_binary: checkerboard 0/1 pattern scaled by
 * `contrast_exp`. */
int amgb_gen_checkerboard_epsv(int pattern_size, int mode, double contrast_exp,
                               double* out);

/* Cuthill-McKee renumbering of a symmetric-pattern CSR matrix as deal.II's
 * DoFRenumbering::Cuthill_McKee / SparsityTools::reorder_Cuthill_McKee does it (ref testcase1-.../
 * src/main.cpp:183-184, testcase3-.../src/main.cpp:249-250; restated from memory of deal.II): start
 * from the unnumbered point of smallest coordination number (lowest index among ties), then level by
 * level: the unnumbered neighbours of the previous front, ordered by coordination number and, among
 * equals, by index.  Components that are not connected are started the same way.  new_to_old[k] = old
 * index of the point numbered k; reversed != 0 gives reverse Cuthill-McKee.  Apply with
 * amgb_matrix_permute (include/amgb.h). */
int amgb_gen_cuthill_mckee(int64_t n, const int64_t* rowptr, const int32_t* col, int reversed,
                           int32_t* new_to_old);

#ifdef __cplusplus
}
#endif
#endif
