// Solve phase: CSR SpMV family with fused epilogues (plain, residual,
// Jacobi-type C/F half sweep, SpMV+dot, prolong-correct), the V(1,1) cycle
// driver, the dense coarsest-grid solve and PCG with device-resident scalars.
//
// Reference semantics being replaced (all inside hypre/PETSc):
//   hypre_ParCSRMatrixMatvec{,T}, hypre_BoomerAMGRelax (types 0 / 18 with C/F
//   ordering through hypre_BoomerAMGRelaxIF), hypre_BoomerAMGCycle,
//   hypre_GaussElimSolve, KSPSolve_CG driven by deal.II's SolverControl;
//   call sites ref common/amg_solver.h:48,54.
//
// Every kernel here is HBM-bound; algorithmic bytes per launch follow
// SURVEY.md 8(d) and are passed to AMGB_LAUNCH for the roofline report.
#include <cmath>
#include <cstring>

#include "amgb_internal.cuh"

namespace amgb {

constexpr int kBlock = 256;

// ---------------------------------------------------------------------------
// CSR row kernel: LANES threads cooperate on one row, coalesced val/col loads,
// shuffle reduction, epilogue functor decides what to do with (row, A_row . x).
// ---------------------------------------------------------------------------
struct EpiStore {
  double* y;
  __device__ bool active(int64_t) const { return true; }
  __device__ void store(int64_t row, double s) const { y[row] = s; }
  __device__ void skip(int64_t) const {}
};

struct EpiResidual {  // r = f - A u
  const double* f;
  double* r;
  __device__ bool active(int64_t) const { return true; }
  __device__ void store(int64_t row, double s) const { r[row] = f[row] - s; }
  __device__ void skip(int64_t) const {}
};

struct EpiAdd {  // u += P e
  double* u;
  __device__ bool active(int64_t) const { return true; }
  __device__ void store(int64_t row, double s) const { u[row] += s; }
  __device__ void skip(int64_t) const {}
};

// Jacobi-type half sweep, out-of-place: rows with cf == pts (or all rows if
// pts == 0) get out = u + w (f - A u) inv_relax, the others are copied.
struct EpiJacobi {
  const double* f;
  const double* u;
  const double* inv_relax;
  const int32_t* cf;  // -3 counts as -1 (end of hypre_BoomerAMGBuildInterp)
  double* out;
  double w;
  int pts;
  __device__ bool active(int64_t row) const {
    if (pts == 0) return true;
    const int c = cf[row];
    return pts > 0 ? c > 0 : c < 0;
  }
  __device__ void store(int64_t row, double s) const {
    out[row] = u[row] + w * (f[row] - s) * inv_relax[row];
  }
  __device__ void skip(int64_t row) const { out[row] = u[row]; }
};

template <int LANES, class Epi>
__global__ void __launch_bounds__(kBlock)
csr_rows_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                const double* __restrict__ val, const double* __restrict__ x, Epi epi) {
  const int64_t row = ((int64_t)blockIdx.x * kBlock + threadIdx.x) / LANES;
  const int lane = threadIdx.x % LANES;
  double s = 0.0;
  bool act = false;
  if (row < n) {
    act = epi.active(row);
    if (act) {
      const int b = rp[row], e = rp[row + 1];
      for (int k = b + lane; k < e; k += LANES) s += val[k] * x[col[k]];
    }
  }
#pragma unroll
  for (int d = LANES / 2; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d, LANES);
  if (lane == 0 && row < n) {
    if (act) epi.store(row, s); else epi.skip(row);
  }
}

// SpMV with fused dot: w = A p and partial[blockIdx] = sum over the block's rows of p_i w_i.
template <int LANES>
__global__ void __launch_bounds__(kBlock)
csr_spmv_dot_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                    const double* __restrict__ val, const double* __restrict__ x,
                    double* __restrict__ y, double* __restrict__ partial) {
  const int64_t row = ((int64_t)blockIdx.x * kBlock + threadIdx.x) / LANES;
  const int lane = threadIdx.x % LANES;
  double s = 0.0;
  if (row < n) {
    const int b = rp[row], e = rp[row + 1];
    for (int k = b + lane; k < e; k += LANES) s += val[k] * x[col[k]];
  }
#pragma unroll
  for (int d = LANES / 2; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d, LANES);
  double c = 0.0;
  if (lane == 0 && row < n) {
    y[row] = s;
    c = x[row] * s;
  }
  // fixed-tree block reduction (deterministic for a given n)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  __shared__ double ws[kBlock / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kBlock / 32; ++i) t += ws[i];
    partial[blockIdx.x] = t;
  }
}

static int pick_lanes(const DeviceCsr& A) {
  const double avg = A.n > 0 ? double(A.nnz) / double(A.n) : 1.0;
  int lanes = 1;
  while (lanes < 32 && lanes * 4 < avg) lanes <<= 1;
  return lanes;
}

template <class Epi>
static int launch_rows(amgb_ctx* ctx, const DeviceCsr& A, const double* x, Epi epi, int family,
                       double bytes) {
  if (A.n == 0) return AMGB_OK;
  const int lanes = pick_lanes(A);
  const unsigned grid = (unsigned)div_up(A.n * lanes, kBlock);
  switch (lanes) {
    case 1: AMGB_LAUNCH(ctx, family, bytes, (csr_rows_kernel<1, Epi>), grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, epi); break;
    case 2: AMGB_LAUNCH(ctx, family, bytes, (csr_rows_kernel<2, Epi>), grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, epi); break;
    case 4: AMGB_LAUNCH(ctx, family, bytes, (csr_rows_kernel<4, Epi>), grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, epi); break;
    case 8: AMGB_LAUNCH(ctx, family, bytes, (csr_rows_kernel<8, Epi>), grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, epi); break;
    case 16: AMGB_LAUNCH(ctx, family, bytes, (csr_rows_kernel<16, Epi>), grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, epi); break;
    default: AMGB_LAUNCH(ctx, family, bytes, (csr_rows_kernel<32, Epi>), grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, epi); break;
  }
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

// plain CSR bytes, rp = 4 B.
static double csr_bytes(const DeviceCsr& A) { return 12.0 * A.nnz + 4.0 * (A.n + 1); }

int spmv(amgb_ctx* ctx, const DeviceCsr& A, const double* x, double* y, int family) {
  return launch_rows(ctx, A, x, EpiStore{y}, family, csr_bytes(A) + 8.0 * A.ncols + 8.0 * A.n);
}

static int spmv_dot(amgb_ctx* ctx, const DeviceCsr& A, const double* x, double* y, double* partial,
                    int64_t* nblocks) {
  const int lanes = pick_lanes(A);
  const unsigned grid = (unsigned)div_up(A.n * lanes, kBlock);
  *nblocks = grid;
  const double bytes = csr_bytes(A) + 16.0 * A.n;
  switch (lanes) {
    case 1: AMGB_LAUNCH(ctx, F_SPMV, bytes, csr_spmv_dot_kernel<1>, grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, y, partial); break;
    case 2: AMGB_LAUNCH(ctx, F_SPMV, bytes, csr_spmv_dot_kernel<2>, grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, y, partial); break;
    case 4: AMGB_LAUNCH(ctx, F_SPMV, bytes, csr_spmv_dot_kernel<4>, grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, y, partial); break;
    case 8: AMGB_LAUNCH(ctx, F_SPMV, bytes, csr_spmv_dot_kernel<8>, grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, y, partial); break;
    case 16: AMGB_LAUNCH(ctx, F_SPMV, bytes, csr_spmv_dot_kernel<16>, grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, y, partial); break;
    default: AMGB_LAUNCH(ctx, F_SPMV, bytes, csr_spmv_dot_kernel<32>, grid, kBlock, 0, A.n, A.rp.p, A.col.p, A.val.p, x, y, partial); break;
  }
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Level auxiliaries: diagonal position, 1/l1 or 1/diag for the smoother.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
level_aux_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                 const double* __restrict__ val, int relax_type, double* __restrict__ inv_relax) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  double diag = 0.0, l1 = 0.0;
  for (int k = rp[i]; k < rp[i + 1]; ++k) {
    const double v = val[k];
    if (col[k] == i) diag = v;
    l1 += fabs(v);
  }
  double inv = 0.0;  // rows with a zero diagonal are skipped (hypre_BoomerAMGRelax)
  if (diag != 0.0) inv = relax_type == 18 ? 1.0 / l1 : 1.0 / diag;
  inv_relax[i] = inv;
}

int level_aux(amgb_precond* P, int level) {
  amgb_ctx* ctx = P->ctx;
  Level& L = P->lv[level];
  const int64_t n = L.A.n;
  AMGB_TRY(L.inv_relax.alloc(ctx, n));
  // one smoother family per hierarchy: down == up is enforced at initialize
  AMGB_LAUNCH(ctx, F_AUX, csr_bytes(L.A) + 8.0 * n, level_aux_kernel, (unsigned)div_up(n, kBlock), kBlock,
              0, n, L.A.rp.p, L.A.col.p, L.A.val.p, P->relax_down, L.inv_relax.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_TRY(L.tmp.alloc(ctx, n));
  if (level > 0) {
    AMGB_TRY(L.u.alloc(ctx, n));
    AMGB_TRY(L.f.alloc(ctx, n));
  }
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Dense coarsest-grid operator: hypre_gselim order (no pivoting), one block.
// Each row update sequence is the sequential one, so the result is bit-identical
// to the scalar loop.
// ---------------------------------------------------------------------------
__global__ void csr_to_dense_kernel(int64_t n, const int32_t* __restrict__ rp,
                                    const int32_t* __restrict__ col, const double* __restrict__ val,
                                    double* __restrict__ M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = rp[i]; k < rp[i + 1]; ++k) M[i * n + col[k]] = val[k];
}

__global__ void __launch_bounds__(kBlock) dense_factor_kernel(int n, double* __restrict__ M) {
  for (int k = 0; k + 1 < n; ++k) {
    const double pivot = M[(int64_t)k * n + k];
    if (pivot != 0.0) {
      for (int j = k + 1 + threadIdx.x; j < n; j += kBlock) {
        const double mjk = M[(int64_t)j * n + k];
        if (mjk != 0.0) {
          const double factor = mjk / pivot;
          for (int m = k + 1; m < n; ++m)
            M[(int64_t)j * n + m] = __dsub_rn(M[(int64_t)j * n + m], __dmul_rn(factor, M[(int64_t)k * n + m]));
          M[(int64_t)j * n + k] = factor;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kBlock)
dense_solve_kernel(int n, const double* __restrict__ M, const double* __restrict__ f,
                   double* __restrict__ x) {
  for (int i = threadIdx.x; i < n; i += kBlock) x[i] = f[i];
  __syncthreads();
  for (int k = 0; k + 1 < n; ++k) {
    if (M[(int64_t)k * n + k] != 0.0) {
      const double xk = x[k];
      for (int j = k + 1 + threadIdx.x; j < n; j += kBlock) {
        const double l = M[(int64_t)j * n + k];
        if (l != 0.0) x[j] = __dsub_rn(x[j], __dmul_rn(l, xk));
      }
    }
    __syncthreads();
  }
  for (int k = n - 1; k > 0; --k) {
    const double d = M[(int64_t)k * n + k];
    if (d != 0.0) {
      if (threadIdx.x == 0) x[k] = x[k] / d;
      __syncthreads();
      const double xk = x[k];
      for (int j = threadIdx.x; j < k; j += kBlock) {
        const double u = M[(int64_t)j * n + k];
        if (u != 0.0) x[j] = __dsub_rn(x[j], __dmul_rn(xk, u));
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && n > 0 && M[0] != 0.0) x[0] = x[0] / M[0];
}

constexpr int64_t kMaxDenseCoarse = 1024;  // same limit as the oracle

static int setup_dense(amgb_precond* P) {
  amgb_ctx* ctx = P->ctx;
  Level& C = P->lv.back();
  P->dense_ok = false;
  if (P->relax_coarse != 9 || C.A.n > kMaxDenseCoarse) return AMGB_OK;
  const int64_t n = C.A.n;
  AMGB_TRY(P->dense.alloc_zero(ctx, n * n));
  AMGB_LAUNCH(ctx, F_COARSE, csr_bytes(C.A), csr_to_dense_kernel, (unsigned)div_up(n, 128), 128, 0, n,
              C.A.rp.p, C.A.col.p, C.A.val.p, P->dense.p);
  AMGB_LAUNCH(ctx, F_COARSE, 8.0 * n * n, dense_factor_kernel, 1, kBlock, 0, (int)n, P->dense.p);
  AMGB_CHECK_LAUNCH(ctx);
  P->dense_ok = true;
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// V-cycle (hypre_BoomerAMGCycle): pre-smooth, residual, restrict, recurse,
// prolong-correct, post-smooth; C/F ordering per hypre_BoomerAMGRelaxIF.
// ---------------------------------------------------------------------------
struct CycleVecs {
  double* u;
  const double* f;
};

static int half_sweep(amgb_precond* P, Level& L, const double* f, const double* u_in, double* u_out,
                      int pts) {
  const bool fine = &L == &P->lv[0];
  EpiJacobi epi{f, u_in, L.inv_relax.p, L.cf.p, u_out, P->data.relax_weight, L.cf.p ? pts : 0};
  // bytes: the rows touched; for a C/F half sweep roughly the selected share.
  // Reported as the full-sweep formula of SURVEY.md 8(d) scaled by the share of
  // rows relaxed is not knowable cheaply; count vectors fully and the matrix by
  // share = n_coarse/n (C) or 1 - n_coarse/n (F).
  double share = 1.0;
  if (epi.pts != 0 && L.A.n > 0) {
    const double c = double(L.n_coarse) / double(L.A.n);
    share = pts > 0 ? c : 1.0 - c;
  }
  const double bytes = share * csr_bytes(L.A) + 32.0 * L.A.n;
  return launch_rows(P->ctx, L.A, u_in, epi, fine ? F_SMOOTH_L0 : F_SMOOTH, bytes);
}

// One hypre_BoomerAMGRelaxIF call; the result ends in `u` (L.tmp is scratch).
static int relax_if(amgb_precond* P, Level& L, const double* f, double* u, int cycle_param) {
  if (P->data.relax_order == 1 && cycle_param < 3 && L.cf.p) {
    const int p0 = cycle_param < 2 ? 1 : -1;
    AMGB_TRY(half_sweep(P, L, f, u, L.tmp.p, p0));
    AMGB_TRY(half_sweep(P, L, f, L.tmp.p, u, -p0));
  } else {
    AMGB_TRY(half_sweep(P, L, f, u, L.tmp.p, 0));
    AMGB_CUDA(P->ctx, cudaMemcpyAsync(u, L.tmp.p, L.A.n * sizeof(double), cudaMemcpyDeviceToDevice,
                                      P->ctx->stream));
  }
  return AMGB_OK;
}

static int cycle(amgb_precond* P, int l, double* u, const double* f) {
  amgb_ctx* ctx = P->ctx;
  Level& L = P->lv[l];
  const int nl = (int)P->lv.size();
  if (l == nl - 1) {
    if (P->relax_coarse == 9 && P->dense_ok) {
      AMGB_LAUNCH(ctx, F_COARSE, 8.0 * L.A.n * L.A.n, dense_solve_kernel, 1, kBlock, 0, (int)L.A.n,
                  P->dense.p, f, u);
      AMGB_CHECK_LAUNCH(ctx);
    } else {
      const unsigned sweeps = P->data.n_sweeps_coarse ? P->data.n_sweeps_coarse : 1u;
      for (unsigned s = 0; s < sweeps; ++s) AMGB_TRY(relax_if(P, L, f, u, 3));
    }
    return AMGB_OK;
  }
  for (unsigned s = 0; s < P->data.n_sweeps; ++s) AMGB_TRY(relax_if(P, L, f, u, 1));
  // residual into tmp, restriction into the coarse rhs
  AMGB_TRY(launch_rows(ctx, L.A, u, EpiResidual{f, L.tmp.p}, l == 0 ? F_RESIDUAL_L0 : F_RESIDUAL,
                       csr_bytes(L.A) + 24.0 * L.A.n));
  Level& C = P->lv[l + 1];
  AMGB_TRY(launch_rows(ctx, L.R, L.tmp.p, EpiStore{C.f.p}, l == 0 ? F_RESTRICT_L0 : F_RESTRICT,
                       csr_bytes(L.R) + 8.0 * L.A.n + 8.0 * C.A.n));
  AMGB_CUDA(ctx, cudaMemsetAsync(C.u.p, 0, C.A.n * sizeof(double), ctx->stream));
  AMGB_TRY(cycle(P, l + 1, C.u.p, C.f.p));
  if (P->data.w_cycle && l + 1 < nl - 1) AMGB_TRY(cycle(P, l + 1, C.u.p, C.f.p));
  AMGB_TRY(launch_rows(ctx, L.P, C.u.p, EpiAdd{u}, l == 0 ? F_PROLONG_L0 : F_PROLONG,
                       csr_bytes(L.P) + 8.0 * C.A.n + 16.0 * L.A.n));
  for (unsigned s = 0; s < P->data.n_sweeps; ++s) AMGB_TRY(relax_if(P, L, f, u, 2));
  return AMGB_OK;
}

int vcycle_apply(amgb_precond* P, double* z_dev, const double* r_dev) {
  amgb_ctx* ctx = P->ctx;
  AMGB_CUDA(ctx, cudaMemsetAsync(z_dev, 0, P->lv[0].A.n * sizeof(double), ctx->stream));
  const unsigned iters = P->data.max_iter ? P->data.max_iter : 1u;
  for (unsigned it = 0; it < iters; ++it) AMGB_TRY(cycle(P, 0, z_dev, r_dev));
  return AMGB_OK;
}

void destroy_solve_state(amgb_precond* P) {
  if (P->vcycle_graph) {
    cudaGraphExecDestroy(P->vcycle_graph);
    P->vcycle_graph = nullptr;
  }
}

int finish_solve_setup(amgb_precond* P) {
  for (int l = 0; l < (int)P->lv.size(); ++l) AMGB_TRY(level_aux(P, l));
  return setup_dense(P);
}

// ---------------------------------------------------------------------------
// PCG vector kernels.  Scalars stay on the device:
//   sc[0]=beta sc[1]=beta_old sc[2]=(p,w) sc[3]=alpha sc[4]=dp
//   fl[0]=done fl[1]=iterations fl[2]=status
// Dot products use fixed-size chunks reduced in a fixed tree, then the chunk
// partials are summed in index order by one block: the value does not depend
// on the grid size.
// ---------------------------------------------------------------------------
constexpr int kDotItems = 8;
constexpr int kDotChunk = kBlock * kDotItems;

__device__ __forceinline__ double block_sum(double c, double* ws) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kBlock / 32; ++i) t += ws[i];
  }
  __syncthreads();
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(kBlock)
dot2_kernel(int64_t n, const double* __restrict__ z, const double* __restrict__ r,
            double* __restrict__ pzz, double* __restrict__ pzr) {
  __shared__ double ws[kBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kDotChunk;
  double a = 0.0, b = 0.0;
#pragma unroll
  for (int k = 0; k < kDotItems; ++k) {
    const int64_t i = base + (int64_t)k * kBlock + threadIdx.x;
    if (i < n) {
      const double zi = z[i];
      a += zi * zi;
      b += zi * r[i];
    }
  }
  const double ta = block_sum(a, ws);
  const double tb = block_sum(b, ws);
  if (threadIdx.x == 0) {
    pzz[blockIdx.x] = ta;
    pzr[blockIdx.x] = tb;
  }
}

// sums m partials in index order with one block: thread t owns a contiguous
// slice, slices are combined by the fixed block tree.
__device__ double ordered_sum(const double* __restrict__ p, int64_t m, double* ws) {
  const int64_t per = (m + kBlock - 1) / kBlock;
  const int64_t b = (int64_t)threadIdx.x * per;
  const int64_t e = b + per < m ? b + per : m;
  double s = 0.0;
  for (int64_t i = b; i < e; ++i) s += p[i];
  return block_sum(s, ws);
}

__global__ void __launch_bounds__(kBlock)
finalize_alpha_kernel(const double* __restrict__ partial, int64_t m, double* sc, int* fl) {
  __shared__ double ws[kBlock / 32];
  const double pw = ordered_sum(partial, m, ws);
  if (threadIdx.x == 0) {
    sc[2] = pw;
    if (pw == 0.0 || pw != pw) {
      fl[0] = 1;
      fl[2] = AMGB_ERR_BREAKDOWN;
      sc[3] = 0.0;
    } else {
      sc[3] = sc[0] / pw;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
finalize_beta_kernel(const double* __restrict__ pzz, const double* __restrict__ pzr, int64_t m,
                     double* sc, int* fl, double* hist, int64_t hist_cap, double abs_tol, int first) {
  __shared__ double ws[kBlock / 32];
  const double zz = ordered_sum(pzz, m, ws);
  const double zr = ordered_sum(pzr, m, ws);
  if (threadIdx.x == 0) {
    const double dp = sqrt(zz);
    sc[4] = dp;
    sc[1] = sc[0];
    sc[0] = zr;
    const int it = first ? 0 : fl[1] + 1;
    fl[1] = it;
    if (it < hist_cap) hist[it] = dp;
    if (dp != dp) {
      fl[0] = 1;
      fl[2] = AMGB_ERR_BREAKDOWN;
    } else if (dp <= abs_tol) {
      fl[0] = 1;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
update_p_kernel(int64_t n, const double* __restrict__ z, double* __restrict__ p,
                const double* __restrict__ sc, int first) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  if (first) {
    p[i] = z[i];
  } else {
    const double bb = sc[0] / sc[1];
    p[i] = z[i] + bb * p[i];
  }
}

__global__ void __launch_bounds__(kBlock)
axpy2_kernel(int64_t n, const double* __restrict__ p, const double* __restrict__ w,
             double* __restrict__ x, double* __restrict__ r, const double* __restrict__ sc) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const double alpha = sc[3];
  x[i] += alpha * p[i];
  r[i] -= alpha * w[i];
}

struct PcgFlags {
  int done, iters, status, pad;
};

static int cg_device(amgb_ctx* ctx, const amgb_matrix* A, double* x, const double* b, amgb_precond* P,
                     int64_t max_steps, double abs_tol, double* res_hist, int64_t hist_cap,
                     int64_t* n_iters) {
  const int64_t n = A->A.n;
  if (P->lv.empty() || P->lv[0].A.n != n)
    return set_error(ctx, AMGB_ERR_BAD_ARG, "preconditioner was initialised for another matrix");
  DevBuf<double> r, z, p, w, sc, hist, pa, pb;
  DevBuf<int> fl;
  AMGB_TRY(r.alloc(ctx, n));
  AMGB_TRY(z.alloc(ctx, n));
  AMGB_TRY(p.alloc(ctx, n));
  AMGB_TRY(w.alloc(ctx, n));
  AMGB_TRY(sc.alloc_zero(ctx, 8));
  AMGB_TRY(fl.alloc_zero(ctx, 4));
  int64_t cap = hist_cap > 0 && res_hist ? hist_cap : 1;
  if (cap > max_steps + 1) cap = max_steps + 1;
  if (cap < 1) cap = 1;
  AMGB_TRY(hist.alloc_zero(ctx, cap));
  const int64_t dot_blocks = div_up(n, kDotChunk);
  const int64_t spmv_blocks = div_up(n * 32, kBlock);  // upper bound for any LANES
  AMGB_TRY(pa.alloc(ctx, spmv_blocks > dot_blocks ? spmv_blocks : dot_blocks));
  AMGB_TRY(pb.alloc(ctx, dot_blocks));
  const unsigned vgrid = (unsigned)div_up(n, kBlock);

  // r = b - A x ; z = M^{-1} r ; dp = ||z|| ; beta = (z, r)
  AMGB_TRY(launch_rows(ctx, A->A, x, EpiResidual{b, r.p}, F_RESIDUAL, csr_bytes(A->A) + 24.0 * n));
  AMGB_TRY(vcycle_apply(P, z.p, r.p));
  AMGB_LAUNCH(ctx, F_VEC, 16.0 * n, dot2_kernel, (unsigned)dot_blocks, kBlock, 0, n, z.p, r.p, pa.p, pb.p);
  AMGB_LAUNCH(ctx, F_VEC, 16.0 * dot_blocks, finalize_beta_kernel, 1, kBlock, 0, pa.p, pb.p, dot_blocks,
              sc.p, fl.p, hist.p, cap, abs_tol, 1);
  AMGB_CHECK_LAUNCH(ctx);
  PcgFlags* hf = (PcgFlags*)ctx->pinned;
  auto read_flags = [&]() -> int {
    AMGB_CUDA(ctx, cudaMemcpyAsync(hf, fl.p, sizeof(PcgFlags), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AMGB_OK;
  };
  AMGB_TRY(read_flags());
  int64_t it = 0;
  while (!hf->done && it < max_steps) {
    AMGB_LAUNCH(ctx, F_VEC, 24.0 * n, update_p_kernel, vgrid, kBlock, 0, n, z.p, p.p, sc.p, it == 0 ? 1 : 0);
    int64_t nb = 0;
    AMGB_TRY(spmv_dot(ctx, A->A, p.p, w.p, pa.p, &nb));
    AMGB_LAUNCH(ctx, F_VEC, 8.0 * nb, finalize_alpha_kernel, 1, kBlock, 0, pa.p, nb, sc.p, fl.p);
    AMGB_LAUNCH(ctx, F_VEC, 48.0 * n, axpy2_kernel, vgrid, kBlock, 0, n, p.p, w.p, x, r.p, sc.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(vcycle_apply(P, z.p, r.p));
    AMGB_LAUNCH(ctx, F_VEC, 16.0 * n, dot2_kernel, (unsigned)dot_blocks, kBlock, 0, n, z.p, r.p, pa.p, pb.p);
    AMGB_LAUNCH(ctx, F_VEC, 16.0 * dot_blocks, finalize_beta_kernel, 1, kBlock, 0, pa.p, pb.p, dot_blocks,
                sc.p, fl.p, hist.p, cap, abs_tol, 0);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(read_flags());
    ++it;
  }
  *n_iters = hf->iters;
  const int status = hf->status;
  const bool done = hf->done != 0;
  if (res_hist && hist_cap > 0) {
    int64_t k = hf->iters + 1;
    if (k > cap) k = cap;
    AMGB_CUDA(ctx, cudaMemcpyAsync(res_hist, hist.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (status != 0) return set_error(ctx, status, "PCG breakdown at iteration %d", hf->iters);
  if (!done)
    return set_error(ctx, AMGB_ERR_NO_CONVERGENCE, "PCG did not reach %g in %lld steps", abs_tol,
                     (long long)max_steps);
  return AMGB_OK;
}

}  // namespace amgb

using namespace amgb;

extern "C" {

int amgb_precond_vmult_device(amgb_precond* P, double* dst_device, const double* src_device) {
  if (!P || !dst_device || !src_device) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(P->ctx->device);
  return vcycle_apply(P, dst_device, src_device);
}

int amgb_precond_vmult(amgb_precond* P, double* dst, const double* src) {
  if (!P || !dst || !src) return AMGB_ERR_BAD_ARG;
  amgb_ctx* ctx = P->ctx;
  cudaSetDevice(ctx->device);
  const int64_t n = P->lv[0].A.n;
  DevBuf<double> d, s;
  AMGB_TRY(d.alloc(ctx, n));
  AMGB_TRY(s.alloc(ctx, n));
  AMGB_CUDA(ctx, cudaMemcpyAsync(s.p, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_TRY(vcycle_apply(P, d.p, s.p));
  AMGB_CUDA(ctx, cudaMemcpyAsync(dst, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int amgb_cg_solve_device(amgb_ctx* ctx, const amgb_matrix* A, double* x_device, const double* b_device,
                         amgb_precond* P, int64_t max_steps, double abs_tol, double* res_hist,
                         int64_t hist_cap, int64_t* n_iters) {
  if (!ctx || !A || !x_device || !b_device || !P || !n_iters || max_steps < 0) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  return cg_device(ctx, A, x_device, b_device, P, max_steps, abs_tol, res_hist, hist_cap, n_iters);
}

int amgb_cg_solve(amgb_ctx* ctx, const amgb_matrix* A, double* x, const double* b, amgb_precond* P,
                  int64_t max_steps, double abs_tol, double* res_hist, int64_t hist_cap,
                  int64_t* n_iters) {
  if (!ctx || !A || !x || !b || !P || !n_iters || max_steps < 0) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  const int64_t n = A->A.n;
  DevBuf<double> dx, db;
  AMGB_TRY(dx.alloc(ctx, n));
  AMGB_TRY(db.alloc(ctx, n));
  AMGB_CUDA(ctx, cudaMemcpyAsync(dx.p, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(db.p, b, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const int rc = cg_device(ctx, A, dx.p, db.p, P, max_steps, abs_tol, res_hist, hist_cap, n_iters);
  if (rc == AMGB_OK || rc == AMGB_ERR_NO_CONVERGENCE) {
    AMGB_CUDA(ctx, cudaMemcpyAsync(x, dx.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return rc;
}

}  // extern "C"
