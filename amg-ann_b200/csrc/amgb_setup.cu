// AMG setup on the device: strength of connection, PMIS C/F splitting,
// modified classical interpolation, explicit transpose and the Galerkin triple
// product as two hash SpGEMMs.  This replaces hypre_BoomerAMGSetup reached from
// `preconditioner.initialize(system_matrix, data)` (ref common/amg_solver.h:48);
// the algorithms are restated in SURVEY.md Appendix A.3 and, operation for
// operation, in oracle/amg_oracle.cpp, against which the integer outputs are
// bit-exact and the weights/operators bit-identical.
//
// Bit-exactness rules used throughout (hard parts H3/H5/H6 of SURVEY.md 7.3):
//  * every multiply that feeds an add is written __dmul_rn/__dadd_rn so nvcc can
//    not contract it into an FMA (the oracle is built with -ffp-contract=off);
//  * sums whose result feeds a comparison (row sums, interpolation sums, SpGEMM
//    accumulators) are accumulated in the same sequential order as the oracle:
//    ascending column within a row, ascending k for C(i,c) = sum_k a_ik b_kc.
#include <algorithm>
#include <cstdlib>
#include <chrono>
#include <climits>
#include <cstring>
#include <string>

#include "amgb_internal.cuh"

namespace amgb {

constexpr int kBlock = 256;

// ---------------------------------------------------------------------------
// Strength (hypre_BoomerAMGCreateS).  One thread per row, sequential in the
// oracle's order: diagonal first, then off-diagonals left to right.
// ---------------------------------------------------------------------------
// (c, v, mk are indexed by the global entry number k: either the arrays themselves or the
// warp's shared-memory copy shifted by its first entry)
__device__ __forceinline__ void strength_row(int i, int b, int e, const int32_t* c, const double* v, uint8_t* mk,
                                             double theta, double max_row_sum, int32_t* __restrict__ has_strong,
                                             double* __restrict__ diagv) {
  double diag = 0.0;
  for (int k = b; k < e; ++k)
    if (c[k] == i) diag = v[k];
  diagv[i] = diag;
  double row_scale = 0.0, row_sum = diag;
  if (diag < 0) {
    for (int k = b; k < e; ++k)
      if (c[k] != i) {
        const double a = v[k];
        row_scale = row_scale < a ? a : row_scale;
        row_sum = __dadd_rn(row_sum, a);
      }
  } else {
    for (int k = b; k < e; ++k)
      if (c[k] != i) {
        const double a = v[k];
        row_scale = a < row_scale ? a : row_scale;
        row_sum = __dadd_rn(row_sum, a);
      }
  }
  int any = 0;
  const bool all_weak = fabs(row_sum) > __dmul_rn(fabs(diag), max_row_sum) && max_row_sum < 1.0;
  const double thr = __dmul_rn(theta, row_scale);
  for (int k = b; k < e; ++k) {
    uint8_t m = 0;
    if (!all_weak && c[k] != i) {
      const double a = v[k];
      m = diag < 0 ? (a > thr) : (a < thr);
    }
    mk[k] = m;
    any |= m;
  }
  has_strong[i] = any;
}

// The 32 rows of a warp are one contiguous run of entries: it is copied to shared memory
// with coalesced loads (a thread walking its own row straight from global memory touches a
// new line at nearly every step), the rows are evaluated there, and the mask goes back the
// same way.  Runs longer than the stage (sized by the host from the average row) are evaluated in place.
constexpr int kStrengthBlock = 128;

__global__ void __launch_bounds__(kStrengthBlock)
strength_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                const double* __restrict__ val, double theta, double max_row_sum,
                uint8_t* __restrict__ mask, int32_t* __restrict__ has_strong,
                double* __restrict__ diagv, int stage) {
  // per warp: `stage` values, column ids and mask bytes (stage is a multiple of 8)
  extern __shared__ double strength_smem[];
  constexpr int kW = kStrengthBlock / 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* my_val = strength_smem + (size_t)w * stage;
  int32_t* my_col = reinterpret_cast<int32_t*>(strength_smem + (size_t)kW * stage) + (size_t)w * stage;
  uint8_t* my_mask = reinterpret_cast<uint8_t*>(reinterpret_cast<int32_t*>(strength_smem + (size_t)kW * stage) +
                                                (size_t)kW * stage) + (size_t)w * stage;
  const int64_t r0 = ((int64_t)blockIdx.x * (kStrengthBlock / 32) + w) * 32;
  if (r0 >= n) return;
  const int64_t r1 = r0 + 32 < n ? r0 + 32 : n;
  const int eb = rp[r0], cnt = rp[r1] - eb;
  const int64_t i = r0 + lane;
  const bool staged = cnt <= stage;
  if (staged) {
    for (int t = lane; t < cnt; t += 32) {
      my_col[t] = col[eb + t];
      my_val[t] = val[eb + t];
    }
    __syncwarp();
  }
  if (i < n) {
    if (staged) strength_row((int)i, rp[i], rp[i + 1], my_col - eb, my_val - eb, my_mask - eb, theta, max_row_sum,
                             has_strong, diagv);
    else strength_row((int)i, rp[i], rp[i + 1], col, val, mask, theta, max_row_sum, has_strong, diagv);
  }
  if (staged) {
    __syncwarp();
    for (int t = lane; t < cnt; t += 32) mask[eb + t] = my_mask[t];
  }
}

// ---------------------------------------------------------------------------
// PMIS (hypre_BoomerAMGCoarsenPMIS, one rank).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double hypre_rand_at(int64_t i) {
  // hypre_Rand after hypre_SeedRand(2747): seed_{i} = 2747 * 16807^(i+1) mod (2^31-1)
  const unsigned long long m = 2147483647ull;
  unsigned long long e = (unsigned long long)i + 1ull, base = 16807ull, r = 2747ull;
  while (e) {
    if (e & 1ull) r = (r * base) % m;
    base = (base * base) % m;
    e >>= 1;
  }
  return (double)r / 2147483647.0;
}

__global__ void __launch_bounds__(kBlock)
pmis_influence_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                      const uint8_t* __restrict__ mask, int32_t* __restrict__ influence) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  for (int k = rp[i]; k < rp[i + 1]; ++k)
    if (mask[k]) atomicAdd(&influence[col[k]], 1);
}

__global__ void __launch_bounds__(kBlock)
pmis_init_kernel(int64_t n, const int32_t* __restrict__ influence, const int32_t* __restrict__ has_strong,
                 const int32_t* __restrict__ gid, double* __restrict__ measure, int32_t* __restrict__ cf) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  // the random part depends on the GLOBAL index only, so the splitting is the same for
  // any number of devices (SURVEY.md 8e "Determinism")
  double m = (double)influence[i] + hypre_rand_at(gid ? (int64_t)gid[i] : i);
  int c = 0;
  if (!has_strong[i]) {
    c = -3;  // special F point: no strong connections
    m = 0.0;
  } else if (m < 1.0) {
    c = -1;  // nobody depends on it
    m = 0.0;
  }
  measure[i] = m;
  cf[i] = c;
}

// tentative independent-set membership: undecided points with measure > 1
__global__ void __launch_bounds__(kBlock)
pmis_mark_kernel(int64_t n, const int32_t* __restrict__ cf, const double* __restrict__ measure,
                 int32_t* __restrict__ mark) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  mark[i] = (cf[i] == 0 && measure[i] > 1.0) ? 1 : 0;
}

// hypre_BoomerAMGIndepSet: along every strong connection i -> j between two
// candidates the smaller measure is knocked out (writes of 0 only: benign race,
// result independent of scheduling).
__global__ void __launch_bounds__(kBlock)
pmis_knockout_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                     const uint8_t* __restrict__ mask, const int32_t* __restrict__ cf,
                     const double* __restrict__ measure, int32_t* __restrict__ mark) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  if (cf[i] != 0) return;
  const double mi = measure[i];
  if (!(mi > 1.0)) return;
  bool lose = false;
  for (int k = rp[i]; k < rp[i + 1]; ++k) {
    if (!mask[k]) continue;
    const int j = col[k];
    const double mj = measure[j];
    if (mj > 1.0) {
      if (mi > mj) mark[j] = 0;
      else if (mj > mi) lose = true;
    }
  }
  if (lose) mark[i] = 0;
}

__global__ void __launch_bounds__(kBlock)
pmis_set_c_kernel(int64_t n, const int32_t* __restrict__ mark, int32_t* __restrict__ cf) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  if (cf[i] == 0 && mark[i]) cf[i] = 1;
}

// undecided points that strongly depend on a C point become F; decided points
// leave the graph (measure = 0); counts the points still undecided.
__global__ void __launch_bounds__(kBlock)
pmis_set_f_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                  const uint8_t* __restrict__ mask, int32_t* __restrict__ cf,
                  double* __restrict__ measure, int32_t* __restrict__ undecided, int64_t own_begin,
                  int64_t own_end) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  int und = 0;
  if (i < n) {
    int c = cf[i];
    if (c == 0) {
      for (int k = rp[i]; k < rp[i + 1]; ++k)
        if (mask[k] && cf[col[k]] > 0) {  // cf of others only moves 0 -> -1 here, never to > 0
          c = -1;
          break;
        }
      if (c != 0) cf[i] = c;
    }
    if (c != 0) measure[i] = 0.0; else und = (i >= own_begin && i < own_end) ? 1 : 0;
  }
  const unsigned b = __ballot_sync(0xffffffffu, und);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(undecided, __popc(b));
}

__global__ void __launch_bounds__(kBlock)
cpoint_flag_kernel(int64_t n, const int32_t* __restrict__ cf, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) flag[i] = cf[i] > 0 ? 1 : 0;
}

// Small levels: the whole coarsening -- influence counts, measures, every round -- in ONE block with
// block barriers between the phases (the phases are the kernels above, row for row).  A level of a
// few thousand rows is launch-bound: ~6 rounds of four kernels, a memset and a host read each become
// one launch and no host round trip.  Same splitting: no phase depends on the order of the threads.
constexpr int kPmisSmallThreads = 256;
constexpr int64_t kPmisSmallRows = 2048;

__global__ void __launch_bounds__(kPmisSmallThreads)
pmis_small_kernel(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                  const uint8_t* __restrict__ mask, const int32_t* __restrict__ has_strong,
                  int32_t* influence /* zeroed */, double* measure, int32_t* mark, int32_t* cf) {
  const int t = threadIdx.x;
  for (int i = t; i < n; i += kPmisSmallThreads)
    for (int k = rp[i]; k < rp[i + 1]; ++k)
      if (mask[k]) atomicAdd(&influence[col[k]], 1);
  __syncthreads();
  for (int i = t; i < n; i += kPmisSmallThreads) {
    double m = (double)influence[i] + hypre_rand_at(i);
    int c = 0;
    if (!has_strong[i]) {
      c = -3;
      m = 0.0;
    } else if (m < 1.0) {
      c = -1;
      m = 0.0;
    }
    measure[i] = m;
    cf[i] = c;
  }
  __syncthreads();
  for (int round = 0; round < 100000; ++round) {
    for (int i = t; i < n; i += kPmisSmallThreads) mark[i] = (cf[i] == 0 && measure[i] > 1.0) ? 1 : 0;
    __syncthreads();
    for (int i = t; i < n; i += kPmisSmallThreads) {
      if (cf[i] != 0) continue;
      const double mi = measure[i];
      if (!(mi > 1.0)) continue;
      bool lose = false;
      for (int k = rp[i]; k < rp[i + 1]; ++k) {
        if (!mask[k]) continue;
        const int j = col[k];
        const double mj = measure[j];
        if (mj > 1.0) {
          if (mi > mj) mark[j] = 0;
          else if (mj > mi) lose = true;
        }
      }
      if (lose) mark[i] = 0;
    }
    __syncthreads();
    for (int i = t; i < n; i += kPmisSmallThreads)
      if (cf[i] == 0 && mark[i]) cf[i] = 1;
    __syncthreads();
    int und = 0;
    for (int i = t; i < n; i += kPmisSmallThreads) {
      int c = cf[i];
      if (c == 0) {
        for (int k = rp[i]; k < rp[i + 1]; ++k)
          if (mask[k] && cf[col[k]] > 0) {
            c = -1;
            break;
          }
        if (c != 0) cf[i] = c;
      }
      if (c != 0) measure[i] = 0.0; else und = 1;
    }
    if (__syncthreads_count(und) == 0) break;
  }
}

int coarsen_pmis(amgb_ctx* ctx, const DeviceCsr& A, const uint8_t* mask, const int32_t* has_strong, int32_t* cf,
                 DistHooks* hooks) {
  const int64_t n = A.n;
  if (!hooks && n <= kPmisSmallRows && !std::getenv("AMGB_NO_SMALL_LEVELS")) {
    DevBuf<int32_t> influence, mark;
    DevBuf<double> measure;
    AMGB_TRY(influence.alloc_zero(ctx, n));
    AMGB_TRY(mark.alloc(ctx, n));
    AMGB_TRY(measure.alloc(ctx, n));
    AMGB_LAUNCH(ctx, F_COARSEN, 40.0 * A.nnz + 60.0 * n, pmis_small_kernel, 1, kPmisSmallThreads, 0, (int)n, A.rp.p,
                A.col.p, mask, has_strong, influence.p, measure.p, mark.p, cf);
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  }
  const int64_t own_begin = hooks ? hooks->own_begin : 0, own_end = hooks ? hooks->own_end : n;
  const unsigned grid = (unsigned)div_up(n, kBlock);
  DevBuf<int32_t> influence, mark, undecided;
  DevBuf<double> measure;
  AMGB_TRY(influence.alloc_zero(ctx, n));
  AMGB_TRY(mark.alloc(ctx, n));
  AMGB_TRY(measure.alloc(ctx, n));
  AMGB_TRY(undecided.alloc(ctx, 1));
  const double row_bytes = 5.0 * A.nnz + 4.0 * (n + 1);
  AMGB_LAUNCH(ctx, F_COARSEN, row_bytes, pmis_influence_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
              influence.p);
  AMGB_LAUNCH(ctx, F_COARSEN, 20.0 * n, pmis_init_kernel, grid, kBlock, 0, n, influence.p, has_strong,
              hooks ? hooks->gid : (const int32_t*)nullptr, measure.p, cf);
  AMGB_CHECK_LAUNCH(ctx);
  if (hooks) {  // influence counts / states of ghost points come from their owners
    AMGB_TRY(hooks->sync_f64(measure.p));
    AMGB_TRY(hooks->sync_i32(cf));
  }
  for (int round = 0; round < 100000; ++round) {
    AMGB_CUDA(ctx, cudaMemsetAsync(undecided.p, 0, sizeof(int32_t), ctx->stream));
    AMGB_LAUNCH(ctx, F_COARSEN, 16.0 * n, pmis_mark_kernel, grid, kBlock, 0, n, cf, measure.p, mark.p);
    AMGB_LAUNCH(ctx, F_COARSEN, row_bytes, pmis_knockout_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
                cf, measure.p, mark.p);
    AMGB_LAUNCH(ctx, F_COARSEN, 12.0 * n, pmis_set_c_kernel, grid, kBlock, 0, n, mark.p, cf);
    if (hooks) AMGB_TRY(hooks->sync_i32(cf));
    AMGB_LAUNCH(ctx, F_COARSEN, row_bytes, pmis_set_f_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask, cf,
                measure.p, undecided.p, own_begin, own_end);
    AMGB_CHECK_LAUNCH(ctx);
    if (hooks) {
      AMGB_TRY(hooks->sync_i32(cf));
      AMGB_TRY(hooks->sync_f64(measure.p));
    }
    int32_t und = 0;
    AMGB_TRY(read_i32(ctx, undecided.p, &und));
    if (hooks) {
      int64_t total = und;
      AMGB_TRY(hooks->allreduce_sum(&total));
      und = total > 0 ? 1 : 0;
    }
    if (und == 0) return AMGB_OK;
  }
  return set_error(ctx, AMGB_ERR_BREAKDOWN, "PMIS did not terminate");
}

// ---------------------------------------------------------------------------
// CLJP (coarsen type 0, hypre_BoomerAMGCoarsen): the parallel coarsening of the Falgout
// family -- under Falgout (the reference's PCHYPRE default, serial on one rank) this routine
// is the third stage.  Operation for operation oracle/amg_oracle.cpp::coarsen_cljp (restated
// from memory like the rest; parity unpinned).  The measure of a point is the number of
// strong connections INTO it that are still in the graph, plus the PMIS tie breaker; the
// independent-set step is the PMIS one; what differs are the edge removals and measure
// decrements after every round.  Measures change by exact decrements of 1.0 (atomicAdd of
// -1.0 is exact here), so the splitting does not depend on the order of the threads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
cljp_init_kernel(int64_t n, const int32_t* __restrict__ influence, const int32_t* __restrict__ has_strong,
                 double* __restrict__ measure, int32_t* __restrict__ cf) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const bool any = has_strong[i] != 0;
  cf[i] = any ? 0 : -3;  // rows without strong connections never coarsen (special F points)
  measure[i] = any ? (double)influence[i] + hypre_rand_at(i) : 0.0;
}

__global__ void __launch_bounds__(kBlock)
cljp_edges_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                  const uint8_t* __restrict__ mask, const int32_t* __restrict__ cf, uint8_t* __restrict__ edge) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  for (int k = rp[i]; k < rp[i + 1]; ++k) edge[k] = mask[k] && cf[col[k]] != -3;
}

// (1) undecided points nobody depends on any more, whose own dependencies are all accounted
// for, become F; decided points leave the graph; counts the undecided ones
__global__ void __launch_bounds__(kBlock)
cljp_set_f_kernel(int64_t n, const int32_t* __restrict__ rp, const uint8_t* __restrict__ edge,
                  int32_t* __restrict__ cf, double* __restrict__ measure, int32_t* __restrict__ undecided) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  int und = 0;
  if (i < n) {
    int c = cf[i];
    if (c == 0 && measure[i] < 1.0) {
      bool open = false;
      for (int k = rp[i]; k < rp[i + 1] && !open; ++k) open = edge[k] != 0;
      if (!open) {
        c = -1;
        cf[i] = -1;
      }
    }
    if (c != 0) measure[i] = 0.0; else und = 1;
  }
  const unsigned b = __ballot_sync(0xffffffffu, und);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(undecided, __popc(b));
}

// (3) new C points drop their dependency edges and their targets lose one; (4) undecided
// points drop the edges to C points and the edges to points they share a C point with
__global__ void __launch_bounds__(kBlock)
cljp_update_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                   const uint8_t* __restrict__ mask, const int32_t* __restrict__ cf,
                   const int32_t* __restrict__ newc, uint8_t* __restrict__ edge, double* __restrict__ measure) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const int b = rp[i], e = rp[i + 1];
  if (newc[i]) {
    for (int k = b; k < e; ++k)
      if (edge[k]) {
        edge[k] = 0;
        if (cf[col[k]] == 0) atomicAdd(&measure[col[k]], -1.0);
      }
    return;
  }
  if (cf[i] != 0) return;
  for (int k = b; k < e; ++k)
    if (mask[k] && cf[col[k]] > 0) edge[k] = 0;
  for (int k = b; k < e; ++k) {
    if (!edge[k]) continue;
    const int j = col[k];
    bool common = false;
    for (int k2 = rp[j]; k2 < rp[j + 1] && !common; ++k2) {
      if (!mask[k2]) continue;
      const int c = col[k2];
      if (cf[c] <= 0) continue;
      // is c a strong dependency of i as well?  (row i is sorted by column)
      int lo = b, hi = e;
      while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (col[mid] < c) lo = mid + 1; else hi = mid;
      }
      common = lo < e && col[lo] == c && mask[lo];
    }
    if (common) {
      edge[k] = 0;
      atomicAdd(&measure[j], -1.0);
    }
  }
}

int coarsen_cljp(amgb_ctx* ctx, const DeviceCsr& A, const uint8_t* mask, const int32_t* has_strong, int32_t* cf) {
  const int64_t n = A.n;
  const unsigned grid = (unsigned)div_up(n, kBlock);
  DevBuf<int32_t> influence, mark, undecided;
  DevBuf<double> measure;
  DevBuf<uint8_t> edge;
  AMGB_TRY(influence.alloc_zero(ctx, n));
  AMGB_TRY(mark.alloc(ctx, n));
  AMGB_TRY(measure.alloc(ctx, n));
  AMGB_TRY(undecided.alloc(ctx, 1));
  AMGB_TRY(edge.alloc(ctx, A.nnz));
  const double row_bytes = 5.0 * A.nnz + 4.0 * (n + 1);
  AMGB_LAUNCH(ctx, F_COARSEN, row_bytes, pmis_influence_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
              influence.p);
  AMGB_LAUNCH(ctx, F_COARSEN, 20.0 * n, cljp_init_kernel, grid, kBlock, 0, n, (const int32_t*)influence.p, has_strong,
              measure.p, cf);
  AMGB_LAUNCH(ctx, F_COARSEN, row_bytes + A.nnz, cljp_edges_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
              (const int32_t*)cf, edge.p);
  AMGB_CHECK_LAUNCH(ctx);
  for (int round = 0; round < 100000; ++round) {
    AMGB_CUDA(ctx, cudaMemsetAsync(undecided.p, 0, sizeof(int32_t), ctx->stream));
    AMGB_LAUNCH(ctx, F_COARSEN, A.nnz + 16.0 * n, cljp_set_f_kernel, grid, kBlock, 0, n, A.rp.p,
                (const uint8_t*)edge.p, cf, measure.p, undecided.p);
    AMGB_CHECK_LAUNCH(ctx);
    int32_t und = 0;
    AMGB_TRY(read_i32(ctx, undecided.p, &und));
    if (und == 0) return AMGB_OK;
    AMGB_LAUNCH(ctx, F_COARSEN, 16.0 * n, pmis_mark_kernel, grid, kBlock, 0, n, (const int32_t*)cf,
                (const double*)measure.p, mark.p);
    AMGB_LAUNCH(ctx, F_COARSEN, row_bytes, pmis_knockout_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
                (const int32_t*)cf, (const double*)measure.p, mark.p);
    AMGB_LAUNCH(ctx, F_COARSEN, 12.0 * n, pmis_set_c_kernel, grid, kBlock, 0, n, (const int32_t*)mark.p, cf);
    AMGB_LAUNCH(ctx, F_COARSEN, 2.0 * row_bytes, cljp_update_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
                (const int32_t*)cf, (const int32_t*)mark.p, edge.p, measure.p);
    AMGB_CHECK_LAUNCH(ctx);
  }
  return set_error(ctx, AMGB_ERR_BREAKDOWN, "CLJP did not terminate");
}

// ---------------------------------------------------------------------------
// Interpolation type 0, modified classical (hypre_BoomerAMGBuildInterp).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
interp_count_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                    const uint8_t* __restrict__ mask, const int32_t* __restrict__ cf,
                    int32_t* __restrict__ count, int64_t row_begin, int64_t row_end) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  int c = 0;
  if (i < row_begin || i >= row_end) {
    c = 0;
  } else if (cf[i] > 0) {
    c = 1;
  } else {
    for (int k = rp[i]; k < rp[i + 1]; ++k)
      if (mask[k] && cf[col[k]] > 0) ++c;
  }
  count[i] = c;
}

constexpr int kInterpStage = 48;  // P-row entries staged in shared memory per warp
constexpr int kInterpLaneMax = 16;     // longest P row of the lane-parallel route
constexpr int kInterpLaneStride = 17;  // (odd stride: the lanes of phase A hit distinct banks)

// position of coarse column cc in the sorted slice pcol[0..len), or -1
__device__ __forceinline__ int find_sorted(const int32_t* pcol, int len, int cc) {
  int lo = 0, hi = len;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (pcol[mid] < cc) lo = mid + 1; else hi = mid;
  }
  return (lo < len && pcol[lo] == cc) ? lo : -1;
}

// One warp per row.  Row i's entries are walked in order (the oracle's order);
// for a strong F neighbour k the warp reads row k cooperatively (coalesced), the
// qualifying entries are found with a ballot and summed in lane (= column)
// order, and each qualifying lane then updates its own, distinct, P entry.  The
// P row holds FINE column ids while it is being built (sorted ascending, so
// membership is a binary search) and is renumbered at the end.
__global__ void __launch_bounds__(kBlock)
interp_fill_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                   const double* __restrict__ val, const uint8_t* __restrict__ mask,
                   const int32_t* __restrict__ cf, const int32_t* __restrict__ f2c,
                   const double* __restrict__ diagv, const int32_t* __restrict__ prp, int32_t* pcol,
                   double* pval, int64_t row_begin, const uint8_t* __restrict__ only) {
  const int64_t i = row_begin + (((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  if (only && !only[i]) return;  // second stage of the grouped kernel: the rows it left behind
  const unsigned full = 0xffffffffu;
  const int jb = prp[i];
  const int len = prp[i + 1] - jb;
  if (cf[i] > 0) {
    if (lane == 0) {
      pcol[jb] = f2c[i];
      pval[jb] = 1.0;
    }
    return;
  }
  const int b = rp[i], e = rp[i + 1];
  // the row of P under construction lives in shared memory when it is short (the usual
  // case): it is searched and updated once per strong F neighbour
  __shared__ int32_t s_c[kBlock / 32][kInterpStage];
  __shared__ double s_v[kBlock / 32][kInterpStage];
  const bool staged = len <= kInterpStage;
  int32_t* myc = staged ? s_c[threadIdx.x >> 5] : pcol + jb;
  double* myv = staged ? s_v[threadIdx.x >> 5] : pval + jb;
  // Short P rows (the usual case) take the lane-parallel route: every strong F neighbour of
  // a 32-entry chunk is reduced by its own lane (phase A), the matched entries of its row
  // parked in s_m, and lane p then accumulates P entry p over the chunk in entry order
  // (phase B).  Same operands, same order as the one-neighbour-at-a-time route below.
  __shared__ double s_m[kBlock / 32][32 * kInterpLaneStride];
  __shared__ int s_lane[kBlock / 32][32], s_b1[kBlock / 32][32], s_e1[kBlock / 32][32];
  __shared__ unsigned s_has[kBlock / 32][32];
  const bool lane_parallel = len <= kInterpLaneMax;
  double* mym = s_m[threadIdx.x >> 5];
  double acc = 0.0;  // lane p's P entry (lane_parallel)
  // phase 0: diagonal, and the strong C neighbours (fine ids) in row order
  double diagonal = 0.0;
  {
    int base = 0;
    for (int kb = b; kb < e; kb += 32) {
      const int k = kb + lane;
      bool isc = false;
      int i1 = -1;
      double a = 0.0;
      if (k < e) {
        i1 = col[k];
        a = val[k];
        isc = i1 != (int)i && mask[k] && cf[i1] > 0;
      }
      const unsigned dm = __ballot_sync(full, k < e && i1 == (int)i);
      if (dm) diagonal = __shfl_sync(full, a, __ffs(dm) - 1);
      const unsigned cm = __ballot_sync(full, isc);
      if (isc) {
        const int pos = base + __popc(cm & ((1u << lane) - 1u));
        myc[pos] = i1;
        myv[pos] = 0.0;
      }
      base += __popc(cm);
    }
  }
  __syncwarp();
  // phase 1: entries of row i in order.  Each accumulator keeps the sequential order of
  // its own terms: P(i,p) receives the C entry p and the distributions of the strong F
  // neighbours in entry order; the diagonal receives the weak entries and the strong F
  // neighbours that share no C point with i, in entry order, at the end of the chunk.
  // The row of the NEXT strong F neighbour (values, columns, C/F markers of the columns)
  // is fetched while the current one is reduced, so its latency is off the critical path.
  int seen_c = 0;
  for (int kb = b; kb < e; kb += 32) {
    const int k = kb + lane;
    int my_i1 = -1, my_c1 = -3, my_strong = 0, my_b1 = 0, my_len1 = 0;
    double my_a = 0.0, my_sgn = 1.0;
    if (k < e) {
      my_i1 = col[k];
      my_a = val[k];
      my_strong = mask[k];
      my_c1 = cf[my_i1];
    }
    const bool offd = k < e && my_i1 != (int)i;
    const bool is_c = offd && my_strong && my_c1 > 0;
    const bool is_sf = offd && my_strong && my_c1 <= 0 && my_c1 != -3;
    const bool is_weak = offd && !my_strong && my_c1 != -3;
    if (is_sf) {
      my_b1 = rp[my_i1];
      my_len1 = rp[my_i1 + 1] - my_b1;
      my_sgn = diagv[my_i1] < 0 ? -1.0 : 1.0;
    }
    const unsigned sfm = __ballot_sync(full, is_sf), cm = __ballot_sync(full, is_c);
    const unsigned wm = __ballot_sync(full, is_weak);
    if (lane_parallel) {
      // phase A: one lane per (strong F neighbour k, C point p of the P row) pair looks a_kp up
      // in row k by binary search (rows are sorted); the hits are parked in s_m and flagged
      const int wi = threadIdx.x >> 5;
      const int32_t* sc = s_c[wi];
      const int nsf = __popc(sfm);
      if (is_sf) {
        const int t = __popc(sfm & ((1u << lane) - 1u));
        s_lane[wi][t] = lane;
        s_b1[wi][t] = my_b1;
        s_e1[wi][t] = my_sgn < 0 ? -(my_b1 + my_len1) - 1 : my_b1 + my_len1;  // sign of a_kk rides along
        s_has[wi][lane] = 0u;
      }
      __syncwarp();
      for (int q = lane; q < nsf * len; q += 32) {
        const int t = q / len, pos = q - t * len;
        const int lt = s_lane[wi][t], want = sc[pos];
        const int ee = s_e1[wi][t];
        const int e1 = ee < 0 ? -(ee + 1) : ee;
        int lo = s_b1[wi][t], hi = e1;
        while (lo < hi) {
          const int mid = lo + ((hi - lo) >> 1);  // entry numbers reach past 2^30: lo + hi would overflow
          if (col[mid] < want) lo = mid + 1; else hi = mid;
        }
        if (lo < e1 && col[lo] == want) {
          const double v = val[lo];
          if (ee < 0 ? v > 0 : v < 0) {
            mym[lt * kInterpLaneStride + pos] = v;
            atomicOr(&s_has[wi][lt], 1u << pos);
          }
        }
      }
      __syncwarp();
      // the sum of neighbour k over the shared C points, in column order
      double my_sum = 0.0;
      unsigned my_has = 0;
      if (is_sf) {
        my_has = s_has[wi][lane];
        for (unsigned m = my_has; m; m &= m - 1)
          my_sum = __dadd_rn(my_sum, mym[lane * kInterpLaneStride + __ffs(m) - 1]);
      }
      const double my_dist = (is_sf && my_sum != 0) ? my_a / my_sum : 0.0;
      const unsigned zs = __ballot_sync(full, is_sf && my_sum == 0);
      __syncwarp();
      // phase B: P entry `lane` over the chunk, in entry order
      for (unsigned both = sfm | cm; both; both &= both - 1) {
        const int t = __ffs(both) - 1;
        if ((cm >> t) & 1u) {
          const double a = __shfl_sync(full, my_a, t);
          if (lane == seen_c) acc = __dadd_rn(acc, a);
          ++seen_c;
        } else {
          const double dist = __shfl_sync(full, my_dist, t);
          const unsigned has = __shfl_sync(full, my_has, t);
          if ((has >> lane) & 1u) acc = __dadd_rn(acc, __dmul_rn(dist, mym[t * kInterpLaneStride + lane]));
        }
      }
      __syncwarp();  // s_m is rewritten by the next chunk
      for (unsigned m = wm | zs; m; m &= m - 1)
        diagonal = __dadd_rn(diagonal, __shfl_sync(full, my_a, __ffs(m) - 1));
      continue;
    }
    unsigned zero_sum = 0, rest = sfm;
    double nv = 0.0;  // prefetched row of the lowest strong F neighbour in `rest`
    int ncol = -1, ncf = 0;
    auto prefetch = [&](unsigned m) {
      nv = 0.0;
      ncol = -1;
      ncf = 0;
      if (!m) return;
      const int nt = __ffs(m) - 1;
      const int pb = __shfl_sync(full, my_b1, nt), pl = __shfl_sync(full, my_len1, nt);
      if (pl <= 32 && lane < pl) {
        nv = val[pb + lane];
        ncol = col[pb + lane];
        ncf = cf[ncol];
      }
    };
    prefetch(rest);
    for (unsigned both = sfm | cm; both; both &= both - 1) {
      const int t = __ffs(both) - 1;
      const double a = __shfl_sync(full, my_a, t);
      if ((cm >> t) & 1u) {
        if (lane == 0) myv[seen_c] = __dadd_rn(myv[seen_c], a);
        ++seen_c;
        __syncwarp();
        continue;
      }
      // strong F neighbour: distribute a_ik over the C points i and k share
      const int b1 = __shfl_sync(full, my_b1, t), len1 = __shfl_sync(full, my_len1, t);
      const double sgn = __shfl_sync(full, my_sgn, t);
      const int e1 = b1 + len1;
      double sum = 0.0;
      if (len1 <= 32) {
        // the usual case: row k fits one warp-wide read (already in registers)
        const double v = nv;
        const int i2 = ncol, f2 = ncf;
        rest &= rest - 1;
        prefetch(rest);
        int pos = -1;
        if (i2 >= 0 && sgn * v < 0 && f2 > 0) pos = find_sorted(myc, len, i2);
        for (unsigned m = __ballot_sync(full, pos >= 0); m; m &= m - 1)
          sum = __dadd_rn(sum, __shfl_sync(full, v, __ffs(m) - 1));
        if (sum != 0) {
          const double distribute = a / sum;
          if (pos >= 0) myv[pos] = __dadd_rn(myv[pos], __dmul_rn(distribute, v));
          __syncwarp();
        } else {
          zero_sum |= 1u << t;
        }
      } else {
        for (int k1b = b1; k1b < e1; k1b += 32) {
          const int k1 = k1b + lane;
          double v = 0.0;
          bool q = false;
          if (k1 < e1) {
            v = val[k1];
            if (sgn * v < 0) {
              const int i2 = col[k1];
              q = cf[i2] > 0 && find_sorted(myc, len, i2) >= 0;
            }
          }
          for (unsigned m = __ballot_sync(full, q); m; m &= m - 1)
            sum = __dadd_rn(sum, __shfl_sync(full, v, __ffs(m) - 1));
        }
        if (sum != 0) {
          const double distribute = a / sum;
          for (int k1b = b1; k1b < e1; k1b += 32) {
            const int k1 = k1b + lane;
            if (k1 < e1) {
              const double v = val[k1];
              if (sgn * v < 0) {
                const int i2 = col[k1];
                if (cf[i2] > 0) {
                  const int pos = find_sorted(myc, len, i2);
                  if (pos >= 0) myv[pos] = __dadd_rn(myv[pos], __dmul_rn(distribute, v));
                }
              }
            }
          }
          __syncwarp();
        } else {
          zero_sum |= 1u << t;
        }
        rest &= rest - 1;
        prefetch(rest);
      }
    }
    for (unsigned m = wm | zero_sum; m; m &= m - 1)
      diagonal = __dadd_rn(diagonal, __shfl_sync(full, my_a, __ffs(m) - 1));
  }
  if (lane_parallel && lane < len) myv[lane] = acc;
  __syncwarp();
  // phase 2: scale, renumber to coarse ids
  const double nd = -diagonal;
  for (int t = lane; t < len; t += 32) {
    pval[jb + t] = diagonal == 0.0 ? 0.0 : myv[t] / nd;
    pcol[jb + t] = f2c[myc[t]];
  }
}

// ---------------------------------------------------------------------------
// Grouped interpolation kernel (first stage of build_interp): 8 lanes per row, 4 rows per
// warp, working on a COMPACTED copy of the operator.
//
// What row i needs from the row of a strong F neighbour k are the entries a_kc with c a C
// point and sign(a_kk) a_kc < 0 -- on a 27-point operator about 4 of 27.  They are gathered
// once per level (interp_ac_*: "A_C", ascending columns like A), so that the pair (i, k)
// costs a walk over ~4 entries by ONE lane instead of a warp-wide pass over 27; the eight
// lanes of a group take the strong F neighbours of an 8-entry chunk of row i side by side.
// Every sum keeps the oracle's order: sum_k runs over A_C[k] in column order; P(i,p) and the
// diagonal receive their terms entry by entry of row i (the chunk's entries are applied one
// after the other, the matches of one entry in parallel -- they address distinct P entries).
// Rows with more than 32 interpolation points, or with a neighbour matching more than 8 of
// them, are left to interp_fill_kernel (flagged in `todo`).
// ---------------------------------------------------------------------------
constexpr int kIgLanes = 8;
constexpr int kIgRows = kBlock / kIgLanes;  // rows per block
constexpr int kIgMaxP = 32;                 // interpolation points per row
constexpr int kIgMaxMatch = 8;              // matches of one neighbour

__global__ void __launch_bounds__(kBlock)
interp_ac_count_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                       const double* __restrict__ val, const int32_t* __restrict__ cf,
                       const double* __restrict__ diagv, int32_t* __restrict__ count) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k >= n) return;
  int c = 0;
  if (cf[k] <= 0) {  // only F rows are ever distributed
    const bool neg = diagv[k] < 0;
    for (int e = rp[k]; e < rp[k + 1]; ++e) {
      const double v = val[e];
      c += (cf[col[e]] > 0 && (neg ? v > 0 : v < 0)) ? 1 : 0;
    }
  }
  count[k] = c;
}

__global__ void __launch_bounds__(kBlock)
interp_ac_fill_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                      const double* __restrict__ val, const int32_t* __restrict__ cf,
                      const double* __restrict__ diagv, const int32_t* __restrict__ acrp,
                      int32_t* __restrict__ accol, double* __restrict__ acval) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k >= n || cf[k] > 0) return;
  const bool neg = diagv[k] < 0;
  int w = acrp[k];
  for (int e = rp[k]; e < rp[k + 1]; ++e) {
    const double v = val[e];
    const int c = col[e];
    if (cf[c] > 0 && (neg ? v > 0 : v < 0)) {
      accol[w] = c;
      acval[w] = v;
      ++w;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
interp_fill_group_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                         const double* __restrict__ val, const uint8_t* __restrict__ mask,
                         const int32_t* __restrict__ cf, const int32_t* __restrict__ f2c,
                         const int32_t* __restrict__ acrp, const int32_t* __restrict__ accol,
                         const double* __restrict__ acval, const int32_t* __restrict__ prp,
                         int32_t* __restrict__ pcol, double* __restrict__ pval, int64_t row_begin,
                         uint8_t* __restrict__ todo, int32_t* __restrict__ n_todo) {
  __shared__ int32_t s_cs[kIgRows][kIgMaxP];                      // interpolation points (fine ids, ascending)
  __shared__ double s_pv[kIgRows][kIgMaxP];                       // P entries under construction
  __shared__ int32_t s_mpos[kIgRows][kIgLanes][kIgMaxMatch];      // matches of the chunk's neighbours: P position
  __shared__ double s_mval[kIgRows][kIgLanes][kIgMaxMatch];       // ... and a_kc
  const int g = threadIdx.x / kIgLanes, q = threadIdx.x % kIgLanes;
  const int64_t i = row_begin + (int64_t)blockIdx.x * kIgRows + g;
  const unsigned gm = 0xffu << ((threadIdx.x & 31) / kIgLanes * kIgLanes);
  if (i >= n) return;  // (whole groups leave together)
  const int jb = prp[i], len = prp[i + 1] - jb;
  if (cf[i] > 0) {
    if (q == 0) {
      pcol[jb] = f2c[i];
      pval[jb] = 1.0;
    }
    return;
  }
  if (len > kIgMaxP) {
    if (q == 0) {
      todo[i] = 1;
      atomicAdd(n_todo, 1);
    }
    return;
  }
  int32_t* cs = s_cs[g];
  double* pv = s_pv[g];
  const int b = rp[i], e = rp[i + 1];
  // phase 0: diagonal, interpolation points in row order
  double diagonal = 0.0;
  {
    int base = 0;
    for (int kb = b; kb < e; kb += kIgLanes) {
      const int k = kb + q;
      int i1 = -1;
      double a = 0.0;
      bool isc = false;
      if (k < e) {
        i1 = col[k];
        a = val[k];
        isc = i1 != (int)i && mask[k] && cf[i1] > 0;
      }
      const unsigned sh = (threadIdx.x & 31) / kIgLanes * kIgLanes;
      const unsigned dm = (__ballot_sync(gm, k < e && i1 == (int)i) >> sh) & 0xffu;
      if (dm) diagonal = __shfl_sync(gm, a, __ffs(dm) - 1, kIgLanes);
      const unsigned cm = (__ballot_sync(gm, isc) >> sh) & 0xffu;
      if (isc) {
        const int pos = base + __popc(cm & ((1u << q) - 1u));
        cs[pos] = i1;
        pv[pos] = 0.0;
      }
      base += __popc(cm);
    }
  }
  __syncwarp(gm);
  // phase 1: the entries of row i, 8 at a time
  bool overflow = false;
  int seen_c = 0;
  for (int kb = b; kb < e; kb += kIgLanes) {
    const int k = kb + q;
    int i1 = -1, c1 = -3, strong = 0;
    double a = 0.0;
    if (k < e) {
      i1 = col[k];
      a = val[k];
      strong = mask[k];
      c1 = cf[i1];
    }
    const bool offd = k < e && i1 != (int)i;
    const bool is_c = offd && strong && c1 > 0;
    const bool is_sf = offd && strong && c1 <= 0 && c1 != -3;
    const bool is_weak = offd && !strong && c1 != -3;
    // my neighbour's share: sum over the interpolation points it meets, in column order
    double dist = 0.0;
    int nmatch = 0;
    bool zero = false;
    if (is_sf) {
      double sum = 0.0;
      const int t0 = acrp[i1], nac = acrp[i1 + 1] - t0;
      // the first entries of the compacted row in one batch of independent loads (the row has ~4)
      constexpr int kAcPre = 4;
      int pc[kAcPre];
      double pvv[kAcPre];
#pragma unroll
      for (int m = 0; m < kAcPre; ++m) {
        pc[m] = m < nac ? accol[t0 + m] : -1;
        pvv[m] = m < nac ? acval[t0 + m] : 0.0;
      }
      auto take = [&](int c2, double v) {
        const int pos = find_sorted(cs, len, c2);
        if (pos >= 0) {
          sum = __dadd_rn(sum, v);
          if (nmatch < kIgMaxMatch) {
            s_mpos[g][q][nmatch] = pos;
            s_mval[g][q][nmatch] = v;
          }
          ++nmatch;
        }
      };
#pragma unroll
      for (int m = 0; m < kAcPre; ++m)
        if (m < nac) take(pc[m], pvv[m]);
      for (int t = t0 + kAcPre; t < t0 + nac; ++t) take(accol[t], acval[t]);
      if (nmatch > kIgMaxMatch) overflow = true;
      if (sum != 0) dist = a / sum; else zero = true;
    }
    const unsigned sh = (threadIdx.x & 31) / kIgLanes * kIgLanes;
    if ((__ballot_sync(gm, overflow) >> sh) & 0xffu) {
      overflow = true;
      break;
    }
    __syncwarp(gm);
    // apply the chunk's entries one after the other
    const unsigned cmask = (__ballot_sync(gm, is_c) >> sh) & 0xffu;
    const unsigned fmask = (__ballot_sync(gm, is_sf && !zero) >> sh) & 0xffu;
    const unsigned dmask = (__ballot_sync(gm, is_weak || (is_sf && zero)) >> sh) & 0xffu;
    for (unsigned todo_m = cmask | fmask | dmask; todo_m; todo_m &= todo_m - 1) {
      const int t = __ffs(todo_m) - 1;
      const double at = __shfl_sync(gm, a, t, kIgLanes);
      if ((cmask >> t) & 1u) {
        if (q == 0) pv[seen_c] = __dadd_rn(pv[seen_c], at);
        ++seen_c;
      } else if ((fmask >> t) & 1u) {
        const double dt = __shfl_sync(gm, dist, t, kIgLanes);
        const int nm = __shfl_sync(gm, nmatch, t, kIgLanes);
        if (q < nm) {  // nm <= 8: one match per lane, distinct P entries
          const int pos = s_mpos[g][t][q];
          pv[pos] = __dadd_rn(pv[pos], __dmul_rn(dt, s_mval[g][t][q]));
        }
      } else {
        diagonal = __dadd_rn(diagonal, at);  // (kept identically by every lane of the group)
      }
      __syncwarp(gm);
    }
  }
  if (overflow) {
    if (q == 0) {
      todo[i] = 1;
      atomicAdd(n_todo, 1);
    }
    return;
  }
  // phase 2: scale, renumber to coarse ids
  const double nd = -diagonal;
  for (int t = q; t < len; t += kIgLanes) {
    pval[jb + t] = diagonal == 0.0 ? 0.0 : pv[t] / nd;
    pcol[jb + t] = f2c[cs[t]];
  }
}

// The same kernel with warp-uniform control flow (see spgemm_flat_kernel), opt-in (AMGB_INTERP_UNIFORM=1): measured
// 11.6 -> 9.0 ms on level 0 at m = 200 and green on the whole GPU suite, but it went in shortly before the
// session's last bench run hung and could not be re-measured under three lanes afterwards, so the kernel of
// the last complete bench run stays the default.  The four rows of a warp run every loop to the
// warp's maximum trip count with their own lanes predicated off, and ballots / shuffles / barriers name
// the full warp -- with group masks the four groups were serialised through every collective (ncu,
// round-2 start: 15.7 of 32 threads per issued instruction).
__global__ void __launch_bounds__(kBlock)
interp_fill_group_uniform_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                         const double* __restrict__ val, const uint8_t* __restrict__ mask,
                         const int32_t* __restrict__ cf, const int32_t* __restrict__ f2c,
                         const int32_t* __restrict__ acrp, const int32_t* __restrict__ accol,
                         const double* __restrict__ acval, const int32_t* __restrict__ prp,
                         int32_t* __restrict__ pcol, double* __restrict__ pval, int64_t row_begin,
                         uint8_t* __restrict__ todo, int32_t* __restrict__ n_todo) {
  __shared__ int32_t s_cs[kIgRows][kIgMaxP];                      // interpolation points (fine ids, ascending)
  __shared__ double s_pv[kIgRows][kIgMaxP];                       // P entries under construction
  __shared__ int32_t s_mpos[kIgRows][kIgLanes][kIgMaxMatch];      // matches of the chunk's neighbours: P position
  __shared__ double s_mval[kIgRows][kIgLanes][kIgMaxMatch];       // ... and a_kc
  const unsigned full = 0xffffffffu;
  const int g = threadIdx.x / kIgLanes, q = threadIdx.x % kIgLanes;
  const int64_t i = row_begin + (int64_t)blockIdx.x * kIgRows + g;
  const unsigned sh = (threadIdx.x & 31) / kIgLanes * kIgLanes;
  bool live = i < n;  // group-uniform
  int jb = 0, len = 0;
  if (live) {
    jb = prp[i];
    len = prp[i + 1] - jb;
    if (cf[i] > 0) {
      if (q == 0) {
        pcol[jb] = f2c[i];
        pval[jb] = 1.0;
      }
      live = false;
    } else if (len > kIgMaxP) {
      if (q == 0) {
        todo[i] = 1;
        atomicAdd(n_todo, 1);
      }
      live = false;
    }
  }
  int32_t* cs = s_cs[g];
  double* pv = s_pv[g];
  int b = 0, e = 0;
  if (live) {
    b = rp[i];
    e = rp[i + 1];
  }
  const int maxlen = __reduce_max_sync(full, e - b);
  // phase 0: diagonal, interpolation points in row order
  double diagonal = 0.0;
  {
    int base = 0;
    for (int kb = 0; kb < maxlen; kb += kIgLanes) {
      const int k = b + kb + q;
      int i1 = -1;
      double a = 0.0;
      bool isc = false;
      if (k < e) {
        i1 = col[k];
        a = val[k];
        isc = i1 != (int)i && mask[k] && cf[i1] > 0;
      }
      const unsigned dm = (__ballot_sync(full, k < e && i1 == (int)i) >> sh) & 0xffu;
      const double dv = __shfl_sync(full, a, dm ? __ffs(dm) - 1 : 0, kIgLanes);
      if (dm) diagonal = dv;
      const unsigned cm = (__ballot_sync(full, isc) >> sh) & 0xffu;
      if (isc) {
        const int pos = base + __popc(cm & ((1u << q) - 1u));
        cs[pos] = i1;
        pv[pos] = 0.0;
      }
      base += __popc(cm);
    }
  }
  __syncwarp();
  // phase 1: the entries of row i, 8 at a time
  bool overflow = false;
  int seen_c = 0;
  for (int kb = 0; kb < maxlen; kb += kIgLanes) {
    const int k = b + kb + q;
    int i1 = -1, c1 = -3, strong = 0;
    double a = 0.0;
    if (k < e) {
      i1 = col[k];
      a = val[k];
      strong = mask[k];
      c1 = cf[i1];
    }
    const bool offd = k < e && i1 != (int)i;
    const bool is_c = offd && strong && c1 > 0;
    const bool is_sf = offd && strong && c1 <= 0 && c1 != -3;
    const bool is_weak = offd && !strong && c1 != -3;
    // my neighbour's share: sum over the interpolation points it meets, in column order
    double dist = 0.0;
    int nmatch = 0;
    bool zero = false;
    if (is_sf) {
      double sum = 0.0;
      const int t0 = acrp[i1], nac = acrp[i1 + 1] - t0;
      // the first entries of the compacted row in one batch of independent loads (the row has ~4)
      constexpr int kAcPre = 4;
      int pc[kAcPre];
      double pvv[kAcPre];
#pragma unroll
      for (int m = 0; m < kAcPre; ++m) {
        pc[m] = m < nac ? accol[t0 + m] : -1;
        pvv[m] = m < nac ? acval[t0 + m] : 0.0;
      }
      auto take = [&](int c2, double v) {
        const int pos = find_sorted(cs, len, c2);
        if (pos >= 0) {
          sum = __dadd_rn(sum, v);
          if (nmatch < kIgMaxMatch) {
            s_mpos[g][q][nmatch] = pos;
            s_mval[g][q][nmatch] = v;
          }
          ++nmatch;
        }
      };
#pragma unroll
      for (int m = 0; m < kAcPre; ++m)
        if (m < nac) take(pc[m], pvv[m]);
      for (int t = t0 + kAcPre; t < t0 + nac; ++t) take(accol[t], acval[t]);
      if (nmatch > kIgMaxMatch) overflow = true;
      if (sum != 0) dist = a / sum; else zero = true;
    }
    if ((__ballot_sync(full, overflow) >> sh) & 0xffu) {  // this row goes to the warp-per-row kernel
      overflow = true;
      e = b;  // (nothing more is read; the rest of the loops run empty for this group)
    }
    __syncwarp();
    // apply the chunk's entries one after the other
    // (every lane of the warp takes part in the ballots; an overflowed group then drops its bits)
    unsigned cmask = (__ballot_sync(full, is_c) >> sh) & 0xffu;
    unsigned fmask = (__ballot_sync(full, is_sf && !zero) >> sh) & 0xffu;
    unsigned dmask = (__ballot_sync(full, is_weak || (is_sf && zero)) >> sh) & 0xffu;
    if (overflow) cmask = fmask = dmask = 0u;
    const unsigned any_m = cmask | fmask | dmask;
    const int tmax = __reduce_max_sync(full, 32 - __clz(any_m));
    for (int t = 0; t < tmax; ++t) {
      const double at = __shfl_sync(full, a, t, kIgLanes);
      const double dt = __shfl_sync(full, dist, t, kIgLanes);
      const int nm = __shfl_sync(full, nmatch, t, kIgLanes);
      if ((cmask >> t) & 1u) {
        if (q == 0) pv[seen_c] = __dadd_rn(pv[seen_c], at);
        ++seen_c;
      } else if ((fmask >> t) & 1u) {
        if (q < nm) {  // nm <= 8: one match per lane, distinct P entries
          const int pos = s_mpos[g][t][q];
          pv[pos] = __dadd_rn(pv[pos], __dmul_rn(dt, s_mval[g][t][q]));
        }
      } else if ((dmask >> t) & 1u) {
        diagonal = __dadd_rn(diagonal, at);  // (kept identically by every lane of the group)
      }
      __syncwarp();
    }
  }
  if (!live) return;
  if (overflow) {
    if (q == 0) {
      todo[i] = 1;
      atomicAdd(n_todo, 1);
    }
    return;
  }
  // phase 2: scale, renumber to coarse ids
  const double nd = -diagonal;
  for (int t = q; t < len; t += kIgLanes) {
    pval[jb + t] = diagonal == 0.0 ? 0.0 : pv[t] / nd;
    pcol[jb + t] = f2c[cs[t]];
  }
}

// ---------------------------------------------------------------------------
// Explicit transpose R = P^T with rows sorted by fine index.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
transpose_count_kernel(int64_t nnz, const int32_t* __restrict__ pcol, int32_t* __restrict__ count) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k < nnz) atomicAdd(&count[pcol[k]], 1);
}

__global__ void __launch_bounds__(kBlock)
transpose_fill_kernel(int64_t n, const int32_t* __restrict__ prp, const int32_t* __restrict__ pcol,
                      const double* __restrict__ pval, const int32_t* __restrict__ rrp,
                      int32_t* __restrict__ cursor, int32_t* __restrict__ rcol, double* __restrict__ rval) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  for (int k = prp[i]; k < prp[i + 1]; ++k) {
    const int c = pcol[k];
    const int w = rrp[c] + atomicAdd(&cursor[c], 1);
    rcol[w] = (int)i;
    rval[w] = pval[k];
  }
}

// the atomic fill leaves each row in arbitrary order.  One warp per row: the row's column
// ids go to shared memory, every lane ranks the entries it holds in registers against all
// of them (ids are distinct) and writes them back at their rank.  Rows longer than
// kSortStage fall back to an insertion sort by one lane.
constexpr int kSortStage = 256;

template <int PER>  // entries per lane: rows of up to 32 * PER entries
__device__ __forceinline__ void rank_sort_row(int32_t* key, int lane, int b, int len, int32_t* __restrict__ col,
                                              double* __restrict__ val) {
  int c[PER], rank[PER];
  double v[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int t = j * 32 + lane;
    c[j] = t < len ? col[b + t] : INT_MAX;
    v[j] = t < len ? val[b + t] : 0.0;
    rank[j] = 0;
    if (t < len) key[t] = c[j];
  }
  __syncwarp();  // the whole row is in registers / shared memory from here on
  for (int u = 0; u < len; ++u) {
    const int k = key[u];
#pragma unroll
    for (int j = 0; j < PER; ++j) rank[j] += k < c[j] ? 1 : 0;
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    if (j * 32 + lane < len) {
      col[b + rank[j]] = c[j];
      val[b + rank[j]] = v[j];
    }
  }
}

__global__ void __launch_bounds__(kBlock)
sort_rows_kernel(int64_t n, const int32_t* __restrict__ rp, int32_t* __restrict__ col,
                 double* __restrict__ val) {
  __shared__ int32_t s_key[kBlock / 32][kSortStage];
  const int64_t i = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (i >= n) return;
  const int b = rp[i], e = rp[i + 1], len = e - b;
  if (len <= 1) return;
  if (len <= 64) {
    rank_sort_row<2>(s_key[w], lane, b, len, col, val);
    return;
  }
  if (len <= kSortStage) {
    rank_sort_row<kSortStage / 32>(s_key[w], lane, b, len, col, val);
    return;
  }
  if (lane != 0) return;
  for (int a = b + 1; a < e; ++a) {
    const int c = col[a];
    const double v = val[a];
    int t = a - 1;
    while (t >= b && col[t] > c) {
      col[t + 1] = col[t];
      val[t + 1] = val[t];
      --t;
    }
    col[t + 1] = c;
    val[t + 1] = v;
  }
}

int transpose_csr(amgb_ctx* ctx, const DeviceCsr& P, DeviceCsr& R) {
  R.n = P.ncols;
  R.ncols = P.n;
  R.nnz = P.nnz;
  DevBuf<int32_t> count, cursor;
  AMGB_TRY(count.alloc_zero(ctx, R.n));
  AMGB_TRY(cursor.alloc_zero(ctx, R.n));
  AMGB_TRY(R.rp.alloc(ctx, R.n + 1));
  AMGB_TRY(R.col.alloc(ctx, R.nnz));
  AMGB_TRY(R.val.alloc(ctx, R.nnz));
  if (P.nnz > 0)
    AMGB_LAUNCH(ctx, F_TRANSPOSE, 8.0 * P.nnz, transpose_count_kernel, (unsigned)div_up(P.nnz, kBlock), kBlock,
                0, P.nnz, P.col.p, count.p);
  AMGB_TRY(exclusive_scan_i32(ctx, count.p, R.rp.p, R.n));
  AMGB_LAUNCH(ctx, F_TRANSPOSE, 24.0 * P.nnz + 4.0 * P.n, transpose_fill_kernel, (unsigned)div_up(P.n, kBlock),
              kBlock, 0, P.n, P.rp.p, P.col.p, P.val.p, R.rp.p, cursor.p, R.col.p, R.val.p);
  AMGB_LAUNCH(ctx, F_TRANSPOSE, 24.0 * R.nnz + 4.0 * R.n, sort_rows_kernel, (unsigned)div_up(R.n * 32, kBlock),
              kBlock, 0, R.n, R.rp.p, R.col.p, R.val.p);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// DoF numberings of the step before the path as permutations applied on
// the device.  The generators number nodes lexicographically; the reference's matrices come
// out of deal.II in the order DoFHandler::distribute_dofs hands out indices: active cells in
// refinement-tree order (GridGenerator::subdivided_hyper_cube(c) cells lexicographic, the 8
// children of a refined cell consecutive, x fastest), the 8 vertices of a cell in
// lexicographic order, a vertex numbered when it is first met (ref testcase2-diffusion-
// structured/src/main.cpp:423-425 mesh, :230-232 distribute_dofs + subdomain_wise).  Restated
// from memory of deal.II (not installed here); closed form, no sort: the traversal rank of a
// cell is (coarse lexicographic index) * 8^r + Morton code of its position inside the coarse
// cell; a cell introduces a vertex iff it has the smallest rank among the cells that share it;
// an exclusive scan over the cells in traversal order turns the per-cell counts into numbers.
// ---------------------------------------------------------------------------
struct TreeGrid {
  int c, r, m;  // coarse cells per direction, refinements, m = c << r
  __device__ __forceinline__ long long rank_of(int cx, int cy, int cz) const {
    const int X = cx >> r, Y = cy >> r, Z = cz >> r, mask = (1 << r) - 1;
    const int lx = cx & mask, ly = cy & mask, lz = cz & mask;
    long long mo = 0;
    for (int b = 0; b < r; ++b)
      mo |= ((long long)((lx >> b) & 1) << (3 * b)) | ((long long)((ly >> b) & 1) << (3 * b + 1)) |
            ((long long)((lz >> b) & 1) << (3 * b + 2));
    return (((long long)X + (long long)c * (Y + (long long)c * Z)) << (3 * r)) + mo;
  }
  __device__ __forceinline__ void cell_of(long long t, int& cx, int& cy, int& cz) const {
    const long long mo = t & ((1ll << (3 * r)) - 1);
    long long cc = t >> (3 * r);
    const int X = (int)(cc % c);
    cc /= c;
    const int Y = (int)(cc % c), Z = (int)(cc / c);
    int lx = 0, ly = 0, lz = 0;
    for (int b = 0; b < r; ++b) {
      lx |= (int)((mo >> (3 * b)) & 1) << b;
      ly |= (int)((mo >> (3 * b + 1)) & 1) << b;
      lz |= (int)((mo >> (3 * b + 2)) & 1) << b;
    }
    cx = (X << r) + lx;
    cy = (Y << r) + ly;
    cz = (Z << r) + lz;
  }
  // is cell (cx,cy,cz) with rank t the first one in traversal order that touches vertex (vx,vy,vz)?
  __device__ __forceinline__ bool introduces(long long t, int vx, int vy, int vz) const {
    for (int dz = -1; dz <= 0; ++dz)
      for (int dy = -1; dy <= 0; ++dy)
        for (int dx = -1; dx <= 0; ++dx) {
          const int ox = vx + dx, oy = vy + dy, oz = vz + dz;
          if (ox < 0 || oy < 0 || oz < 0 || ox >= m || oy >= m || oz >= m) continue;
          if (rank_of(ox, oy, oz) < t) return false;
        }
    return true;
  }
};

__global__ void __launch_bounds__(kBlock)
tree_count_kernel(TreeGrid g, long long ncells, int32_t* __restrict__ count) {
  const long long t = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (t >= ncells) return;
  int cx, cy, cz, c = 0;
  g.cell_of(t, cx, cy, cz);
  for (int v = 0; v < 8; ++v) c += g.introduces(t, cx + (v & 1), cy + ((v >> 1) & 1), cz + ((v >> 2) & 1)) ? 1 : 0;
  count[t] = c;
}

__global__ void __launch_bounds__(kBlock)
tree_number_kernel(TreeGrid g, long long ncells, const int32_t* __restrict__ first, int32_t* __restrict__ new_to_lex) {
  const long long t = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (t >= ncells) return;
  int cx, cy, cz;
  g.cell_of(t, cx, cy, cz);
  int next = first[t];
  const long long N = g.m + 1;
  for (int v = 0; v < 8; ++v) {
    const int vx = cx + (v & 1), vy = cy + ((v >> 1) & 1), vz = cz + ((v >> 2) & 1);
    if (g.introduces(t, vx, vy, vz)) new_to_lex[next++] = (int32_t)(vx + N * (vy + N * (long long)vz));
  }
}

// ---- B = Q A Q^T for a renumbering given as new -> old ----
__global__ void __launch_bounds__(kBlock)
invert_perm_kernel(int64_t n, const int32_t* __restrict__ new_to_old, int32_t* __restrict__ old_to_new,
                   int32_t* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const int32_t o = new_to_old[i];
  if (o < 0 || o >= n) {
    atomicAdd(bad, 1);
    return;
  }
  if (atomicExch(&old_to_new[o], (int32_t)i) != -1) atomicAdd(bad, 1);  // hit twice: not a permutation
}

__global__ void __launch_bounds__(kBlock)
perm_row_len_kernel(int64_t n, const int32_t* __restrict__ new_to_old, const int32_t* __restrict__ rp,
                    int32_t* __restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) len[i] = rp[new_to_old[i] + 1] - rp[new_to_old[i]];
}

__global__ void __launch_bounds__(kBlock)
perm_fill_kernel(int64_t n, const int32_t* __restrict__ new_to_old, const int32_t* __restrict__ old_to_new,
                 const int32_t* __restrict__ rp, const int32_t* __restrict__ col, const double* __restrict__ val,
                 const int32_t* __restrict__ brp, int32_t* __restrict__ bcol, double* __restrict__ bval) {
  const int64_t i = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int o = new_to_old[i], b = rp[o], len = rp[o + 1] - b, w = brp[i];
  for (int t = lane; t < len; t += 32) {
    bcol[w + t] = old_to_new[col[b + t]];
    bval[w + t] = val[b + t];
  }
}

int permute_csr(amgb_ctx* ctx, const DeviceCsr& A, const int32_t* new_to_old_device, DeviceCsr& B) {
  const int64_t n = A.n;
  const unsigned grid = (unsigned)div_up(n, kBlock);
  DevBuf<int32_t> inv, len, bad;
  AMGB_TRY(inv.alloc(ctx, n));
  AMGB_TRY(len.alloc(ctx, n));
  AMGB_TRY(bad.alloc_zero(ctx, 1));
  AMGB_CUDA(ctx, cudaMemsetAsync(inv.p, 0xff, n * sizeof(int32_t), ctx->stream));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * n, invert_perm_kernel, grid, kBlock, 0, n, new_to_old_device, inv.p, bad.p);
  AMGB_CHECK_LAUNCH(ctx);
  int32_t nbad = 0;
  AMGB_TRY(read_i32(ctx, bad.p, &nbad));
  if (nbad) return set_error(ctx, AMGB_ERR_BAD_ARG, "renumbering is not a permutation of 0..n-1");
  B.n = B.ncols = n;
  B.nnz = A.nnz;
  AMGB_TRY(B.rp.alloc(ctx, n + 1));
  AMGB_TRY(B.col.alloc(ctx, A.nnz));
  AMGB_TRY(B.val.alloc(ctx, A.nnz));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * n, perm_row_len_kernel, grid, kBlock, 0, n, new_to_old_device,
              (const int32_t*)A.rp.p, len.p);
  AMGB_TRY(exclusive_scan_i32(ctx, len.p, B.rp.p, n));
  AMGB_LAUNCH(ctx, F_AUX, 24.0 * A.nnz, perm_fill_kernel, (unsigned)div_up(n * 32, kBlock), kBlock, 0, n,
              new_to_old_device, (const int32_t*)inv.p, (const int32_t*)A.rp.p, (const int32_t*)A.col.p,
              (const double*)A.val.p, (const int32_t*)B.rp.p, B.col.p, B.val.p);
  AMGB_LAUNCH(ctx, F_AUX, 24.0 * A.nnz, sort_rows_kernel, (unsigned)div_up(n * 32, kBlock), kBlock, 0, n,
              (const int32_t*)B.rp.p, B.col.p, B.val.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// SpGEMM C = A*B: two-phase hash (count, then fill).  G lanes cooperate on one output
// row (G = 8 for short B rows such as A*P, 32 otherwise), so a warp works on 32/G rows
// at once; hash tables live in shared memory and are sized per row, with two fall-back
// stages for the few rows that do not fit (a big shared-memory table, then a table in
// global memory).
//
// Numeric phase, bit-exactness: the entries k of A's row are walked in order; for one k
// the lanes of the group spread over B's row k, whose columns are distinct, so no two
// lanes touch the same accumulator within a step and every C(i,c) accumulates a_ik*b_kc
// in ascending k exactly like the oracle's Gustavson loop (a group barrier orders the
// steps).  The A-row entries (column, value, B-row extent) are loaded G at a time by the
// lanes and broadcast with shuffles, so the only dependent global load inside a step is
// B's row itself.  Rows are finally ordered by column with a rank sort over the
// (compacted) table; `sorted = false` skips that: the result is then only usable as the
// inner operand T of R*(A*P), whose value does not depend on T's column order.
// ---------------------------------------------------------------------------
constexpr int kSpThreads = 128;
constexpr unsigned kEmpty = 0xffffffffu;

__device__ __forceinline__ unsigned group_mask(int G) {
  return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
}

__device__ __forceinline__ int hash_slot(unsigned key, int lgH) {
  return lgH == 0 ? 0 : (int)((key * 2654435761u) >> (32 - lgH));
}

__device__ __forceinline__ int ceil_log2(int v) {  // smallest l with (1<<l) >= v, v >= 1
  return v <= 1 ? 0 : 32 - __clz(v - 1);
}

// insert key; returns 1 if it was not present, 0 if it was, -1 if the table is full
__device__ __forceinline__ int hash_insert(unsigned* keys, int H, int lgH, unsigned key) {
  int h = hash_slot(key, lgH);
  for (int probes = 0; probes < H; ++probes) {
    const unsigned old = atomicCAS(&keys[h], kEmpty, key);
    if (old == kEmpty) return 1;
    if (old == key) return 0;
    h = (h + 1) & (H - 1);
  }
  return -1;
}

template <int G>
__device__ __forceinline__ int group_sum(int v, unsigned gm) {
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(gm, v, d);
  return v;
}

// upper bound on the number of entries of row i of A*B (= number of products)
template <int G>
__device__ __forceinline__ int row_upper_bound(int b, int e, int gl, unsigned gm,
                                               const int32_t* __restrict__ acol,
                                               const int32_t* __restrict__ brp) {
  int ub = 0;
  for (int k = b + gl; k < e; k += G) {
    const int kk = acol[k];
    ub += brp[kk + 1] - brp[kk];
  }
  return group_sum<G>(ub, gm);
}

// number of distinct columns of row i of A*B, or -1 if the table overflowed
template <int G>
__device__ __forceinline__ int symbolic_row(unsigned* keys, int H, int lgH, int b, int e, int gl,
                                            unsigned gm, const int32_t* __restrict__ acol,
                                            const int32_t* __restrict__ brp,
                                            const int32_t* __restrict__ bcol) {
  for (int t = gl; t < H; t += G) keys[t] = kEmpty;
  __syncwarp(gm);
  int cnt = 0, fail = 0;
  for (int kb = b; kb < e; kb += G) {
    const int k = kb + gl;
    int my_bb = 0, my_len = 0;
    if (k < e) {
      const int kk = acol[k];
      my_bb = brp[kk];
      my_len = brp[kk + 1] - my_bb;
    }
    const int steps = min(G, e - kb);
    for (int t = 0; t < steps; ++t) {
      const int bb = __shfl_sync(gm, my_bb, t, G), len = __shfl_sync(gm, my_len, t, G);
      for (int m = gl; m < len; m += G) {
        const int r = hash_insert(keys, H, lgH, (unsigned)bcol[bb + m]);
        if (r < 0) fail = 1; else cnt += r;
      }
    }
    if (__any_sync(gm, fail)) {  // group-uniform: the table is full
      fail = 1;
      break;
    }
  }
  __syncwarp(gm);
  cnt = group_sum<G>(cnt, gm);
  return fail ? -1 : cnt;
}

// rows: list of row ids (nullptr: identity).  Rows whose table would not fit CAP try the
// full CAP table (ub overestimates the distinct count several times in a Galerkin product);
// a full table is detected by the probe limit and the row goes to the overflow list.
template <int G, int CAP>
__global__ void __launch_bounds__(kSpThreads)
spgemm_symbolic_kernel(int64_t nrows, const int32_t* __restrict__ rows, const int32_t* __restrict__ arp,
                       const int32_t* __restrict__ acol, const int32_t* __restrict__ brp,
                       const int32_t* __restrict__ bcol, int32_t* __restrict__ count,
                       int32_t* __restrict__ ovf_rows, int32_t* __restrict__ ovf_info /* [0]=count [1]=max ub */) {
  extern __shared__ unsigned smem_keys[];
  const int g = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned gm = group_mask(G);
  const int64_t idx = (int64_t)blockIdx.x * (kSpThreads / G) + g;
  if (idx >= nrows) return;
  const int i = rows ? rows[idx] : (int)idx;
  unsigned* keys = smem_keys + (size_t)g * CAP;
  const int b = arp[i], e = arp[i + 1];
  const int ub = row_upper_bound<G>(b, e, gl, gm, acol, brp);
  if (ub == 0) {
    if (gl == 0) count[i] = 0;
    return;
  }
  int lgH = ceil_log2(2 * ub);
  if (lgH < 3) lgH = 3;
  if ((1 << lgH) > CAP) lgH = ceil_log2(CAP);
  const int cnt = symbolic_row<G>(keys, 1 << lgH, lgH, b, e, gl, gm, acol, brp, bcol);
  if (gl == 0) {
    if (cnt < 0) {
      const int w = atomicAdd(&ovf_info[0], 1);
      ovf_rows[w] = i;
      atomicMax(&ovf_info[1], ub);
      count[i] = 0;
    } else {
      count[i] = cnt;
    }
  }
}

// oversized rows: one warp per row from the overflow list, table in global memory
__global__ void __launch_bounds__(kSpThreads)
spgemm_symbolic_global_kernel(const int32_t* __restrict__ ovf_rows, int novf, int H, int lgH,
                              unsigned* __restrict__ tables, const int32_t* __restrict__ arp,
                              const int32_t* __restrict__ acol, const int32_t* __restrict__ brp,
                              const int32_t* __restrict__ bcol, int32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * kSpThreads + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * kSpThreads) >> 5;
  unsigned* keys = tables + (size_t)warp * H;
  for (int idx = warp; idx < novf; idx += nwarps) {
    const int i = ovf_rows[idx];
    const int cnt = symbolic_row<32>(keys, H, lgH, arp[i], arp[i + 1], lane, 0xffffffffu, acol, brp, bcol);
    if (lane == 0) count[i] = cnt;
    __syncwarp();
  }
}

// accumulate row i of A*B into the (keys, vals) table, ascending k
template <int G>
__device__ __forceinline__ void accumulate_row(unsigned* keys, double* vals, int H, int lgH, int b, int e, int gl,
                                               unsigned gm, const int32_t* __restrict__ acol,
                                               const double* __restrict__ aval, const int32_t* __restrict__ brp,
                                               const int32_t* __restrict__ bcol, const double* __restrict__ bval) {
  for (int t = gl; t < H; t += G) keys[t] = kEmpty;
  __syncwarp(gm);
  for (int kb = b; kb < e; kb += G) {
    const int k = kb + gl;
    int my_bb = 0, my_len = 0;
    double my_a = 0.0;
    if (k < e) {
      const int kk = acol[k];
      my_a = aval[k];
      my_bb = brp[kk];
      my_len = brp[kk + 1] - my_bb;
    }
    const int steps = min(G, e - kb);
    for (int t = 0; t < steps; ++t) {
      const int bb = __shfl_sync(gm, my_bb, t, G), len = __shfl_sync(gm, my_len, t, G);
      const double a = __shfl_sync(gm, my_a, t, G);
      for (int m = gl; m < len; m += G) {
        const unsigned key = (unsigned)bcol[bb + m];
        const double prod = __dmul_rn(a, bval[bb + m]);
        int h = hash_slot(key, lgH);
        for (;;) {
          const unsigned old = atomicCAS(&keys[h], kEmpty, key);
          if (old == kEmpty) {
            vals[h] = __dadd_rn(0.0, prod);
            break;
          }
          if (old == key) {
            vals[h] = __dadd_rn(vals[h], prod);
            break;
          }
          h = (h + 1) & (H - 1);
        }
      }
      __syncwarp(gm);  // orders the accumulation over k
    }
  }
}

// in-place compaction of the occupied slots to the front (slot order); G slots at a time:
// a chunk is read into registers before anything of it is overwritten, and the write
// positions never run ahead of the chunk being read
template <int G>
__device__ __forceinline__ void compact_table(unsigned* keys, double* vals, int H, int gl, unsigned gm) {
  int out = 0;
  const int sh = (threadIdx.x & 31) / G * G;  // first lane of my group inside the warp
  for (int tb = 0; tb < H; tb += G) {
    const unsigned key = keys[tb + gl];
    const double v = vals[tb + gl];
    const bool occ = key != kEmpty;
    const unsigned om = (__ballot_sync(gm, occ) >> sh) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
    __syncwarp(gm);
    if (occ) {
      const int w = out + __popc(om & ((1u << gl) - 1u));
      keys[w] = key;
      vals[w] = v;
    }
    out += __popc(om);
    __syncwarp(gm);
  }
}

// SORT: ascending columns through a rank sort of the compacted table (out_n is small here:
// rows with big tables go to the second stage, which uses a bitonic network)
template <int G, bool SORT>
__device__ __forceinline__ void emit_row(unsigned* keys, double* vals, int H, int gl, unsigned gm, int out_b,
                                         int out_n, int32_t* __restrict__ ccol, double* __restrict__ cval) {
  compact_table<G>(keys, vals, H, gl, gm);
  for (int t = gl; t < out_n; t += G) {
    const unsigned key = keys[t];
    int pos = t;
    if (SORT) {
      pos = 0;
      for (int u = 0; u < out_n; ++u) pos += keys[u] < key ? 1 : 0;
    }
    ccol[out_b + pos] = (int)key;
    cval[out_b + pos] = vals[t];
  }
}

template <int G, int CAP, bool SORT>
__global__ void __launch_bounds__(kSpThreads)
spgemm_numeric_kernel(int64_t n, const int32_t* __restrict__ arp, const int32_t* __restrict__ acol,
                      const double* __restrict__ aval, const int32_t* __restrict__ brp,
                      const int32_t* __restrict__ bcol, const double* __restrict__ bval,
                      const int32_t* __restrict__ crp, int32_t* __restrict__ ccol,
                      double* __restrict__ cval, int32_t* __restrict__ ovf_rows,
                      int32_t* __restrict__ ovf_info) {
  extern __shared__ unsigned char smem_raw[];
  const int g = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned gm = group_mask(G);
  const int64_t i = (int64_t)blockIdx.x * (kSpThreads / G) + g;
  if (i >= n) return;
  const int out_b = crp[i], out_n = crp[i + 1] - out_b;
  if (out_n == 0) return;
  int lgH = ceil_log2(2 * out_n);
  if ((1 << lgH) < G) lgH = ceil_log2(G);  // the compaction reads the table G slots at a time
  if ((1 << lgH) > CAP) {
    if (gl == 0) {
      const int w = atomicAdd(&ovf_info[0], 1);
      ovf_rows[w] = (int)i;
      atomicMax(&ovf_info[1], out_n);
    }
    return;
  }
  constexpr int kGroups = kSpThreads / G;
  double* vals = reinterpret_cast<double*>(smem_raw) + (size_t)g * CAP;
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw + sizeof(double) * (size_t)kGroups * CAP) + (size_t)g * CAP;
  const int H = 1 << lgH;
  accumulate_row<G>(keys, vals, H, lgH, arp[i], arp[i + 1], gl, gm, acol, aval, brp, bcol, bval);
  emit_row<G, SORT>(keys, vals, H, gl, gm, out_b, out_n, ccol, cval);
}

// second stage: one warp per listed row; bitonic sort of the whole table by key
// (empty = 0xffffffff sorts last)
__device__ __forceinline__ void big_row(unsigned* keys, double* vals, int H, int lgH, int i,
                                        const int32_t* __restrict__ arp, const int32_t* __restrict__ acol,
                                        const double* __restrict__ aval, const int32_t* __restrict__ brp,
                                        const int32_t* __restrict__ bcol, const double* __restrict__ bval,
                                        const int32_t* __restrict__ crp, int32_t* __restrict__ ccol,
                                        double* __restrict__ cval) {
  const int lane = threadIdx.x & 31;
  const unsigned gm = 0xffffffffu;
  accumulate_row<32>(keys, vals, H, lgH, arp[i], arp[i + 1], lane, gm, acol, aval, brp, bcol, bval);
  for (int kk = 2; kk <= H; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < H; t += 32) {
        const int x = t ^ j;
        if (x > t) {
          const unsigned kt = keys[t], kx = keys[x];
          const bool up = (t & kk) == 0;
          if ((kt > kx) == up && kt != kx) {
            keys[t] = kx;
            keys[x] = kt;
            const double vt = vals[t];
            vals[t] = vals[x];
            vals[x] = vt;
          }
        }
      }
      __syncwarp(gm);
    }
  }
  const int out_b = crp[i], out_n = crp[i + 1] - out_b;
  for (int t = lane; t < out_n; t += 32) {
    ccol[out_b + t] = (int)keys[t];
    cval[out_b + t] = vals[t];
  }
  __syncwarp(gm);
}

template <int CAP>
__global__ void __launch_bounds__(kSpThreads)
spgemm_numeric_big_kernel(const int32_t* __restrict__ rows, int nrows, const int32_t* __restrict__ arp,
                          const int32_t* __restrict__ acol, const double* __restrict__ aval,
                          const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
                          const double* __restrict__ bval, const int32_t* __restrict__ crp,
                          int32_t* __restrict__ ccol, double* __restrict__ cval, int32_t* __restrict__ ovf_rows,
                          int32_t* __restrict__ ovf_info) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int kWarpsPerBlock = kSpThreads / 32;
  const int w = threadIdx.x >> 5;
  const int idx = blockIdx.x * kWarpsPerBlock + w;
  if (idx >= nrows) return;
  const int i = rows[idx];
  const int out_n = crp[i + 1] - crp[i];
  int lgH = ceil_log2(2 * out_n);
  if (lgH < 5) lgH = 5;
  if ((1 << lgH) > CAP) {
    if ((threadIdx.x & 31) == 0) {
      const int p = atomicAdd(&ovf_info[0], 1);
      ovf_rows[p] = i;
      atomicMax(&ovf_info[1], out_n);
    }
    return;
  }
  double* vals = reinterpret_cast<double*>(smem_raw) + (size_t)w * CAP;
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw + sizeof(double) * (size_t)kWarpsPerBlock * CAP) + (size_t)w * CAP;
  big_row(keys, vals, 1 << lgH, lgH, i, arp, acol, aval, brp, bcol, bval, crp, ccol, cval);
}

__global__ void __launch_bounds__(kSpThreads)
spgemm_numeric_global_kernel(const int32_t* __restrict__ ovf_rows, int novf, int H, int lgH,
                             unsigned* __restrict__ key_tables, double* __restrict__ val_tables,
                             const int32_t* __restrict__ arp, const int32_t* __restrict__ acol,
                             const double* __restrict__ aval, const int32_t* __restrict__ brp,
                             const int32_t* __restrict__ bcol, const double* __restrict__ bval,
                             const int32_t* __restrict__ crp, int32_t* __restrict__ ccol,
                             double* __restrict__ cval) {
  const int warp = (blockIdx.x * kSpThreads + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * kSpThreads) >> 5;
  unsigned* keys = key_tables + (size_t)warp * H;
  double* vals = val_tables + (size_t)warp * H;
  for (int idx = warp; idx < novf; idx += nwarps)
    big_row(keys, vals, H, lgH, ovf_rows[idx], arp, acol, aval, brp, bcol, bval, crp, ccol, cval);
}


// ---------------------------------------------------------------------------
// Row-per-thread SpGEMM for products whose B rows are SHORT and whose product rows are
// small (A*P on the finest levels of a 27-point operator: ~85 products, ~18 distinct
// columns per row).  The sub-warp hash kernels above spend most of their issue slots on
// idle lanes and per-entry barriers there (a P row has 1-4 entries for 8 lanes).  Here a
// thread owns a row and a private hash table in shared memory, laid out bank-owned: slot s
// of thread t lives at word s*blockDim + t, so thread t only ever touches bank t%32 and the
// tables of a warp never conflict, whatever slots the lanes probe.  No atomics, no
// barriers; the entries k of A's row are walked in order, so every C(i,c) accumulates
// a_ik*b_kc in ascending k from 0.0 -- the oracle's Gustavson loop verbatim.  The emission
// order (table order) is deterministic too: insertion is sequential.  Rows that outgrow the
// table go to the overflow list of the second stage like in the sub-warp kernels.
// ---------------------------------------------------------------------------
constexpr int kRtThreads = 128;
constexpr int kRtSymLg = 7, kRtSymSlots = 1 << kRtSymLg;   // count pass: 128 keys per row (64 KB per block)
constexpr int kRtNumLg = 6, kRtNumSlots = 1 << kRtNumLg;   // numeric pass: 64 (key, value) slots per row (96 KB)
constexpr int kRtNumMaxRow = 48;                           // longest product row accumulated here
constexpr int kRtSymMaxRow = 112;                          // distinct columns the count table is allowed to hold

__global__ void __launch_bounds__(kRtThreads)
spgemm_rowthread_count_kernel(int64_t n, const int32_t* __restrict__ arp, const int32_t* __restrict__ acol,
                              const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
                              int32_t* __restrict__ count, int32_t* __restrict__ ovf_rows,
                              int32_t* __restrict__ ovf_info) {
  extern __shared__ unsigned rt_smem[];
  const int64_t i = (int64_t)blockIdx.x * kRtThreads + threadIdx.x;
  if (i >= n) return;
  unsigned* keys = rt_smem + threadIdx.x;
#pragma unroll 8
  for (int t = 0; t < kRtSymSlots; ++t) keys[t * kRtThreads] = kEmpty;
  int cnt = 0, ub = 0;
  bool full = false;
  const int ab = arp[i], ae = arp[i + 1];
  int mb = 0, me = 0;
  if (ab < ae) {
    const int kk = acol[ab];
    mb = brp[kk];
    me = brp[kk + 1];
  }
  for (int k = ab; k < ae; ++k) {
    int mb2 = 0, me2 = 0;  // the next B row's extent is fetched while this one is inserted
    if (k + 1 < ae) {
      const int kk2 = acol[k + 1];
      mb2 = brp[kk2];
      me2 = brp[kk2 + 1];
    }
    ub += me - mb;
    if (!full)
      for (int m = mb; m < me; ++m) {
        const unsigned key = (unsigned)bcol[m];
        int h = (int)((key * 2654435761u) >> (32 - kRtSymLg));
        for (;;) {
          const unsigned cur = keys[h * kRtThreads];
          if (cur == key) break;
          if (cur == kEmpty) {
            keys[h * kRtThreads] = key;
            ++cnt;
            break;
          }
          h = (h + 1) & (kRtSymSlots - 1);
        }
        if (cnt > kRtSymMaxRow) {  // (an empty slot always remains: the probe loop terminates)
          full = true;
          break;
        }
      }
    mb = mb2;
    me = me2;
  }
  if (full) {
    const int w = atomicAdd(&ovf_info[0], 1);
    ovf_rows[w] = (int)i;
    atomicMax(&ovf_info[1], ub);
    count[i] = 0;
  } else {
    count[i] = cnt;
  }
}

__global__ void __launch_bounds__(kRtThreads)
spgemm_rowthread_numeric_kernel(int64_t n, const int32_t* __restrict__ arp, const int32_t* __restrict__ acol,
                                const double* __restrict__ aval, const int32_t* __restrict__ brp,
                                const int32_t* __restrict__ bcol, const double* __restrict__ bval,
                                const int32_t* __restrict__ crp, int32_t* __restrict__ ccol,
                                double* __restrict__ cval, int32_t* __restrict__ ovf_rows,
                                int32_t* __restrict__ ovf_info) {
  extern __shared__ unsigned char rt_smem_raw[];
  const int64_t i = (int64_t)blockIdx.x * kRtThreads + threadIdx.x;
  if (i >= n) return;
  const int out_b = crp[i], out_n = crp[i + 1] - out_b;
  if (out_n == 0) return;
  if (out_n > kRtNumMaxRow) {
    const int w = atomicAdd(&ovf_info[0], 1);
    ovf_rows[w] = (int)i;
    atomicMax(&ovf_info[1], out_n);
    return;
  }
  double* vals = reinterpret_cast<double*>(rt_smem_raw) + threadIdx.x;
  unsigned* keys = reinterpret_cast<unsigned*>(rt_smem_raw + sizeof(double) * (size_t)kRtNumSlots * kRtThreads) + threadIdx.x;
#pragma unroll 8
  for (int t = 0; t < kRtNumSlots; ++t) keys[t * kRtThreads] = kEmpty;
  const int ab = arp[i], ae = arp[i + 1];
  int mb = 0, me = 0;
  double a = 0.0;
  if (ab < ae) {
    const int kk = acol[ab];
    a = aval[ab];
    mb = brp[kk];
    me = brp[kk + 1];
  }
  for (int k = ab; k < ae; ++k) {
    int mb2 = 0, me2 = 0;
    double a2 = 0.0;
    if (k + 1 < ae) {
      const int kk2 = acol[k + 1];
      a2 = aval[k + 1];
      mb2 = brp[kk2];
      me2 = brp[kk2 + 1];
    }
    for (int m = mb; m < me; ++m) {
      const unsigned key = (unsigned)bcol[m];
      const double prod = __dmul_rn(a, bval[m]);
      int h = (int)((key * 2654435761u) >> (32 - kRtNumLg));
      for (;;) {
        const unsigned cur = keys[h * kRtThreads];
        if (cur == key) {
          vals[h * kRtThreads] = __dadd_rn(vals[h * kRtThreads], prod);
          break;
        }
        if (cur == kEmpty) {
          keys[h * kRtThreads] = key;
          vals[h * kRtThreads] = __dadd_rn(0.0, prod);
          break;
        }
        h = (h + 1) & (kRtNumSlots - 1);
      }
    }
    a = a2;
    mb = mb2;
    me = me2;
  }
  int w = out_b;
  for (int t = 0; t < kRtNumSlots; ++t) {
    const unsigned key = keys[t * kRtThreads];
    if (key != kEmpty) {
      ccol[w] = (int)key;
      cval[w] = vals[t * kRtThreads];
      ++w;
    }
  }
}

// ---------------------------------------------------------------------------
// Flattened first stage for products with SHORT B rows (A*P; B rows of 1-4 entries).
// The sub-warp kernels above walk A's row entry by entry: one dependent global load of a
// B row, one group barrier and mostly idle lanes per entry (27 steps for a 27-point row).
// Here the products of (up to) 32 A entries are numbered p = 0..np-1 in (k, m) order by
// group-wide scans of the B-row lengths; a byte map p -> A entry and the per-entry
// (B offset - first p, a_ik) descriptors live in shared memory, so that lane gl of the
// group takes the products p = base + gl: every lane is busy, the loads of several
// batches are in flight together and nothing waits on a barrier per A entry.
// Order: the products of one batch with the same column are found with match.any and
// added by the first of their lanes in lane order = ascending p = ascending k, starting
// from the table value; batches follow each other.  Every C(i,c) therefore accumulates
// a_ik*b_kc in ascending k from 0.0, the oracle's Gustavson loop, bit for bit.
// Rows whose chunk has more than S products (or whose table would not fit CAP) go to the
// overflow list of the second stage like in the sub-warp kernels.
// ---------------------------------------------------------------------------
constexpr int kFlatChunk = 32;  // A-row entries numbered together

// entry-wise variant: the table only (the chunk descriptors sit between values and keys, the map is not touched)
template <int G, int CAP>
constexpr size_t entry_smem_bytes() {
  constexpr size_t groups = kSpThreads / G;
  return groups * (sizeof(double) * (CAP + kFlatChunk) + sizeof(unsigned) * CAP);
}

template <int G, int CAP, int S, bool NUMERIC>
constexpr size_t flat_smem_bytes() {
  constexpr size_t groups = kSpThreads / G;
  return groups * ((NUMERIC ? sizeof(double) * (CAP + kFlatChunk) : 0) + sizeof(unsigned) * CAP +
                   sizeof(int) * kFlatChunk + S);
}

// Control flow is WARP-UNIFORM: the groups of a warp run every loop to the warp's maximum trip
// count with their own lanes predicated off, and every collective (scan shuffles, match.any,
// votes, barriers) names the full warp -- the group is folded into the matched value.  With
// group masks the compiler has to run the groups' collectives one group after the other (ncu:
// 8 of 32 threads per issued instruction in the batch loop), which costs more than it saves.
// ENTRY (numeric pass only): the entry-by-entry accumulation of the sub-warp kernels (one B row per
// step, the group's lanes spread over it) inside the same warp-uniform skeleton.
template <int G, int CAP, int S, bool NUMERIC, bool SORT, bool ENTRY = false>
__global__ void __launch_bounds__(kSpThreads)
spgemm_flat_kernel(int64_t n, const int32_t* __restrict__ arp, const int32_t* __restrict__ acol,
                   const double* __restrict__ aval, const int32_t* __restrict__ brp,
                   const int32_t* __restrict__ bcol, const double* __restrict__ bval,
                   const int32_t* __restrict__ crp, int32_t* __restrict__ ccol, double* __restrict__ cval,
                   int32_t* __restrict__ count, int32_t* __restrict__ ovf_rows, int32_t* __restrict__ ovf_info) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int kGroups = kSpThreads / G;
  constexpr int R = kFlatChunk / G;  // A entries per lane and chunk
  constexpr int U = 4;               // batches whose B loads are in flight together
  static_assert(kFlatChunk <= 256, "byte map");
  const unsigned full = 0xffffffffu;
  const int g = threadIdx.x / G, gl = threadIdx.x % G;
  const int lane = threadIdx.x & 31;
  const int sh = lane - gl;  // first lane of my group
  const unsigned glow = G == 32 ? full : ((1u << G) - 1u);
  const int64_t i = (int64_t)blockIdx.x * kGroups + g;
  constexpr size_t off8 = NUMERIC ? sizeof(double) * (size_t)kGroups * (CAP + kFlatChunk) : 0;
  double* vals = reinterpret_cast<double*>(smem_raw) + (size_t)g * CAP;
  double* da = reinterpret_cast<double*>(smem_raw) + (size_t)kGroups * CAP + (size_t)g * kFlatChunk;
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw + off8) + (size_t)g * CAP;
  int* db = reinterpret_cast<int*>(smem_raw + off8 + sizeof(unsigned) * (size_t)kGroups * CAP) + (size_t)g * kFlatChunk;
  unsigned char* kmap = smem_raw + off8 + (sizeof(unsigned) * CAP + sizeof(int) * kFlatChunk) * (size_t)kGroups + (size_t)g * S;

  // group state (uniform within the group): live = this group still accumulates its row
  bool live = i < n;
  int b = 0, e = 0, out_b = 0, out_n = 0, lgH = 3;
  bool table_ready = false;
  if (live) {
    b = arp[i];
    e = arp[i + 1];
  }
  if (NUMERIC) {
    if (live) {
      out_b = crp[i];
      out_n = crp[i + 1] - out_b;
      if (out_n == 0) live = false;
    }
    if (live) {
      lgH = ceil_log2(2 * out_n);
      if ((1 << lgH) < G) lgH = ceil_log2(G);  // the compaction reads the table G slots at a time
      if ((1 << lgH) > CAP) {
        if (gl == 0) {
          const int w = atomicAdd(&ovf_info[0], 1);
          ovf_rows[w] = (int)i;
          atomicMax(&ovf_info[1], out_n);
        }
        live = false;
      }
    }
    if (live)
      for (int t = gl; t < (1 << lgH); t += G) keys[t] = kEmpty;
    table_ready = true;
  } else if (live && b == e) {
    if (gl == 0) count[i] = 0;
    live = false;
  }
  if (!live) e = b;
  int cnt = 0, fail = 0;
  const int maxlen = __reduce_max_sync(full, e - b);
  if constexpr (ENTRY) {
    static_assert(NUMERIC, "the count pass has no entry-by-entry variant here");
    const int H = 1 << lgH;
    for (int kb = 0; kb < maxlen; kb += G) {
      const int k = b + kb + gl;
      int my_bb = 0, my_len = 0;
      double my_a = 0.0;
      if (k < e) {
        const int kk = acol[k];
        my_a = aval[k];
        my_bb = brp[kk];
        my_len = brp[kk + 1] - my_bb;
      }
      const int steps = min(G, maxlen - kb);  // warp-uniform
      for (int t = 0; t < steps; ++t) {
        const int bb = __shfl_sync(full, my_bb, t, G), len = __shfl_sync(full, my_len, t, G);
        const double a = __shfl_sync(full, my_a, t, G);
        for (int m = gl; m < len; m += G) {  // (len = 0 past the end of my row)
          const unsigned key = (unsigned)bcol[bb + m];
          const double prod = __dmul_rn(a, bval[bb + m]);
          int h = hash_slot(key, lgH);
          for (;;) {
            const unsigned old = atomicCAS(&keys[h], kEmpty, key);
            if (old == kEmpty) {
              vals[h] = __dadd_rn(0.0, prod);
              break;
            }
            if (old == key) {
              vals[h] = __dadd_rn(vals[h], prod);
              break;
            }
            h = (h + 1) & (H - 1);
          }
        }
        __syncwarp();  // orders the accumulation over k
      }
    }
  } else
  for (int kb = 0; kb < maxlen; kb += kFlatChunk) {
    // ---- number the products of this chunk
    int bb[R], len[R], off[R];
    double a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int k = b + kb + r * G + gl;
      bb[r] = 0;
      len[r] = 0;
      a[r] = 0.0;
      if (k < e) {
        const int kk = acol[k];
        if (NUMERIC) a[r] = aval[k];
        bb[r] = brp[kk];
        len[r] = brp[kk + 1] - bb[r];
      }
    }
    int np = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int incl = len[r];
#pragma unroll
      for (int d = 1; d < G; d <<= 1) {
        const int t = __shfl_up_sync(full, incl, d, G);
        if (gl >= d) incl += t;
      }
      off[r] = np + incl - len[r];
      np += __shfl_sync(full, incl, G - 1, G);
    }
    if (np > S) {  // group-uniform: the row goes to the second stage
      fail = 1;
      np = 0;
      e = b;
    }
    if (!table_ready && np > 0) {  // count pass: the table follows the products of the (usually only) chunk
      lgH = e - b <= kFlatChunk ? ceil_log2(2 * np) : ceil_log2(CAP);
      if (lgH < 3) lgH = 3;
      if ((1 << lgH) > CAP) lgH = ceil_log2(CAP);
      for (int t = gl; t < (1 << lgH); t += G) keys[t] = kEmpty;
      table_ready = true;
    }
    __syncwarp();  // the previous chunk's batches are done with the map
    if (np > 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int idx = r * G + gl;
        if (len[r] > 0) {
          db[idx] = bb[r] - off[r];
          if (NUMERIC) da[idx] = a[r];
          for (int m = 0; m < len[r]; ++m) kmap[off[r] + m] = (unsigned char)idx;
        }
      }
    }
    __syncwarp();
    const int H = 1 << lgH;
    const int npmax = __reduce_max_sync(full, np);
    // ---- batches of G products per group, U batches loaded together
    for (int base = 0; base < npmax; base += G * U) {
      unsigned key[U];
      double prod[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = base + u * G + gl;
        key[u] = 0x80000000u | (unsigned)lane;  // no column: matches nothing
        prod[u] = 0.0;
        if (p < np) {
          const int idx = kmap[p];
          const int addr = db[idx] + p;
          key[u] = (unsigned)bcol[addr];
          if (NUMERIC) prod[u] = __dmul_rn(da[idx], bval[addr]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (base + u * G >= npmax) break;  // warp-uniform
        const bool has = base + u * G + gl < np;
        if (!NUMERIC) {
          if (has) {
            const int r = hash_insert(keys, H, lgH, key[u]);
            if (r < 0) fail = 1; else cnt += r;
          }
        } else {
          // equal columns of MY group: the group's first lane rides in the upper half of the matched value
          const unsigned same = __match_any_sync(full, ((unsigned long long)sh << 32) | key[u]);
          const bool lead = has && (same & (0u - same)) == (1u << lane);
          unsigned rest = 0;
          int h = 0;
          double acc = 0.0;
          if (lead) {
            h = hash_slot(key[u], lgH);
            for (;;) {
              const unsigned old = atomicCAS(&keys[h], kEmpty, key[u]);
              if (old == kEmpty) {
                acc = __dadd_rn(0.0, prod[u]);
                break;
              }
              if (old == key[u]) {
                acc = __dadd_rn(vals[h], prod[u]);
                break;
              }
              h = (h + 1) & (H - 1);
            }
            rest = same & (same - 1);
          }
          while (__any_sync(full, rest != 0)) {  // the other products of my column, in lane order
            const int src = rest ? __ffs(rest) - 1 : lane;
            const double v = __shfl_sync(full, prod[u], src);
            if (rest) {
              acc = __dadd_rn(acc, v);
              rest &= rest - 1;
            }
          }
          if (lead) vals[h] = acc;
          __syncwarp();  // orders the batches
        }
      }
    }
  }
  __syncwarp();
  fail = ((__ballot_sync(full, fail != 0) >> sh) & glow) != 0;
  if (!NUMERIC) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) cnt += __shfl_xor_sync(full, cnt, d, G);
    int ub = 0;
    if (__any_sync(full, fail)) {  // upper bound of the rows handed to the second stage
      const int b0 = i < n ? arp[i] : 0, e0 = i < n ? arp[i + 1] : 0;
      const int lmax = __reduce_max_sync(full, fail ? e0 - b0 : 0);
      for (int k0 = 0; k0 < lmax; k0 += G) {
        const int k = b0 + k0 + gl;
        if (fail && k < e0) {
          const int kk = acol[k];
          ub += brp[kk + 1] - brp[kk];
        }
      }
#pragma unroll
      for (int d = G / 2; d > 0; d >>= 1) ub += __shfl_xor_sync(full, ub, d, G);
    }
    if (gl == 0 && i < n && (live || fail)) {
      if (fail) {
        const int w = atomicAdd(&ovf_info[0], 1);
        ovf_rows[w] = (int)i;
        atomicMax(&ovf_info[1], ub);
        count[i] = 0;
      } else {
        count[i] = cnt;
      }
    }
    return;
  }
  if (fail) {
    if (gl == 0) {
      const int w = atomicAdd(&ovf_info[0], 1);
      ovf_rows[w] = (int)i;
      atomicMax(&ovf_info[1], out_n);
    }
    live = false;
  }
  // ---- in-place compaction of the occupied slots to the front, then the row (rank sort if SORT)
  const int H = live ? (1 << lgH) : 0;
  const int Hmax = __reduce_max_sync(full, H);
  int outp = 0;
  for (int tb = 0; tb < Hmax; tb += G) {
    const int idx = tb + gl;
    const unsigned key = idx < H ? keys[idx] : kEmpty;
    const double v = idx < H ? vals[idx] : 0.0;
    const bool occ = key != kEmpty;
    const unsigned om = (__ballot_sync(full, occ) >> sh) & glow;
    if (occ) {  // (write positions never run ahead of the chunk just read)
      const int w = outp + __popc(om & ((1u << gl) - 1u));
      keys[w] = key;
      vals[w] = v;
    }
    outp += __popc(om);
    __syncwarp();
  }
  if (!live) return;
  for (int t = gl; t < out_n; t += G) {
    const unsigned key = keys[t];
    int pos = t;
    if (SORT) {
      pos = 0;
      for (int u = 0; u < out_n; ++u) pos += keys[u] < key ? 1 : 0;
    }
    ccol[out_b + pos] = (int)key;
    cval[out_b + pos] = vals[t];
  }
}

// 64-bit total of the row counts: the 32-bit scan below would wrap silently
__global__ void __launch_bounds__(kBlock)
sum_counts_kernel(int64_t n, const int32_t* __restrict__ count, unsigned long long* __restrict__ total) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  unsigned long long v = i < n ? (unsigned long long)count[i] : 0ull;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, v);
}

static int read_ovf(amgb_ctx* ctx, const int32_t* ovf_info, int* novf, int* maxv) {
  int32_t* info = (int32_t*)ctx->pinned;
  AMGB_CUDA(ctx, cudaMemcpyAsync(info, ovf_info, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *novf = info[0];
  *maxv = info[1];
  return AMGB_OK;
}

// G: lanes per row; CAP_SYM / CAP_NUM: first-stage table capacities per row group
// rowthread: the first stage of both passes is the row-per-thread kernel pair (unsorted output only)
template <int G, int CAP_SYM, int CAP_NUM>
static int spgemm_impl(amgb_ctx* ctx, const DeviceCsr& A, const DeviceCsr& B, DeviceCsr& C, bool sorted,
                       bool rowthread = false, int flat = 0 /* 1: flattened count pass, 2: flattened numeric pass, 4: entry-wise numeric pass in the warp-uniform kernel */) {
  constexpr int kBigSym = 8192, kBigNum = 2048;  // second-stage capacities (one warp per row)
  constexpr int kFlatS = 2 * CAP_NUM;            // products of one 32-entry chunk of an A row (flat kernels)
  const int64_t n = A.n;
  C.n = n;
  C.ncols = B.ncols;
  DevBuf<int32_t> count, ovf1, ovf2, info;
  AMGB_TRY(count.alloc(ctx, n));
  AMGB_TRY(ovf1.alloc(ctx, n));
  AMGB_TRY(ovf2.alloc(ctx, n));
  AMGB_TRY(info.alloc_zero(ctx, 4));
  AMGB_TRY(C.rp.alloc(ctx, n + 1));
  constexpr int kGroups = kSpThreads / G;
  const unsigned grid = (unsigned)div_up(n, kGroups);
  const double in_bytes = 12.0 * A.nnz + 4.0 * A.n + 12.0 * B.nnz + 4.0 * B.n;
  const int fb_blocks = ctx->sm_count * 2;
  const int fb_warps = fb_blocks * kSpThreads / 32;
  int novf = 0, maxv = 0;
  // ---- symbolic
  {
    if (rowthread) {
      const size_t smem = sizeof(unsigned) * (size_t)kRtSymSlots * kRtThreads;
      AMGB_CUDA(ctx, cudaFuncSetAttribute(spgemm_rowthread_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
      AMGB_LAUNCH(ctx, F_SPGEMM, in_bytes, spgemm_rowthread_count_kernel, (unsigned)div_up(n, kRtThreads), kRtThreads,
                  smem, n, A.rp.p, A.col.p, B.rp.p, B.col.p, count.p, ovf1.p, info.p);
    } else if (flat & 1) {
      auto kern = spgemm_flat_kernel<G, CAP_SYM, kFlatS, false, false>;
      constexpr size_t smem = flat_smem_bytes<G, CAP_SYM, kFlatS, false>();
      if (smem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMGB_LAUNCH(ctx, F_SPGEMM, in_bytes, kern, grid, kSpThreads, smem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p,
                  B.val.p, (const int32_t*)nullptr, (int32_t*)nullptr, (double*)nullptr, count.p, ovf1.p, info.p);
    } else {
      auto kern = spgemm_symbolic_kernel<G, CAP_SYM>;
      const size_t smem = sizeof(unsigned) * (size_t)kGroups * CAP_SYM;
      if (smem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMGB_LAUNCH(ctx, F_SPGEMM, in_bytes, kern, grid, kSpThreads, smem, n, (const int32_t*)nullptr, A.rp.p, A.col.p,
                  B.rp.p, B.col.p, count.p, ovf1.p, info.p);
    }
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(read_ovf(ctx, info.p, &novf, &maxv));
    if (novf > 0) {
      ctx->routes[R_SPGEMM_SYM_BIG]++;
      auto big = spgemm_symbolic_kernel<32, kBigSym>;
      const size_t bsmem = sizeof(unsigned) * (size_t)(kSpThreads / 32) * kBigSym;
      AMGB_CUDA(ctx, cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
      AMGB_LAUNCH(ctx, F_SPGEMM, 0.0, big, (unsigned)div_up(novf, kSpThreads / 32), kSpThreads, bsmem, (int64_t)novf,
                  (const int32_t*)ovf1.p, A.rp.p, A.col.p, B.rp.p, B.col.p, count.p, ovf2.p, info.p + 2);
      AMGB_CHECK_LAUNCH(ctx);
      AMGB_TRY(read_ovf(ctx, info.p + 2, &novf, &maxv));
      if (novf > 0) {
        int lgH = 5;
        const long long want = std::min<long long>(2ll * maxv, 2ll * B.ncols);
        while ((1ll << lgH) < want) ++lgH;
        const int H = 1 << lgH;
        ctx->routes[R_SPGEMM_SYM_GLOBAL]++;
        DevBuf<unsigned> tables;
        AMGB_TRY(tables.alloc(ctx, (size_t)fb_warps * H));
        AMGB_LAUNCH(ctx, F_SPGEMM, 0.0, spgemm_symbolic_global_kernel, fb_blocks, kSpThreads, 0, ovf2.p, novf, H, lgH,
                    tables.p, A.rp.p, A.col.p, B.rp.p, B.col.p, count.p);
        AMGB_CHECK_LAUNCH(ctx);
      }
    }
  }
  DevBuf<unsigned long long> total;
  AMGB_TRY(total.alloc_zero(ctx, 1));
  AMGB_LAUNCH(ctx, F_SCAN, 4.0 * n, sum_counts_kernel, (unsigned)div_up(n, kBlock), kBlock, 0, n,
              (const int32_t*)count.p, total.p);
  AMGB_TRY(exclusive_scan_i32(ctx, count.p, C.rp.p, n));
  int64_t nnz64 = 0;
  AMGB_TRY(read_i64(ctx, (const int64_t*)total.p, &nnz64));
  if (nnz64 > (int64_t)INT32_MAX)
    return set_error(ctx, AMGB_ERR_RANGE,
                     "a sparse product of the setup has %lld entries: row pointers are 32-bit, partition the system "
                     "over more devices", (long long)nnz64);
  const int32_t nnz = (int32_t)nnz64;
  C.nnz = nnz;
  AMGB_TRY(C.col.alloc(ctx, nnz));
  AMGB_TRY(C.val.alloc(ctx, nnz));
  AMGB_CUDA(ctx, cudaMemsetAsync(info.p, 0, 4 * sizeof(int32_t), ctx->stream));
  // ---- numeric
  {
    const size_t smem = (sizeof(unsigned) + sizeof(double)) * (size_t)kGroups * CAP_NUM;
    const double bytes = in_bytes + 12.0 * nnz + 4.0 * n;
    if (rowthread) {
      const size_t rsmem = (sizeof(unsigned) + sizeof(double)) * (size_t)kRtNumSlots * kRtThreads;
      AMGB_CUDA(ctx, cudaFuncSetAttribute(spgemm_rowthread_numeric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)rsmem));
      AMGB_LAUNCH(ctx, F_SPGEMM, bytes, spgemm_rowthread_numeric_kernel, (unsigned)div_up(n, kRtThreads), kRtThreads,
                  rsmem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p, B.val.p, C.rp.p, C.col.p, C.val.p, ovf1.p,
                  info.p);
    } else if ((flat & 2) && sorted) {
      auto kern = spgemm_flat_kernel<G, CAP_NUM, kFlatS, true, true>;
      constexpr size_t fsmem = flat_smem_bytes<G, CAP_NUM, kFlatS, true>();
      if (fsmem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      AMGB_LAUNCH(ctx, F_SPGEMM, bytes, kern, grid, kSpThreads, fsmem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p,
                  B.val.p, C.rp.p, C.col.p, C.val.p, (int32_t*)nullptr, ovf1.p, info.p);
    } else if (flat & 2) {
      auto kern = spgemm_flat_kernel<G, CAP_NUM, kFlatS, true, false>;
      constexpr size_t fsmem = flat_smem_bytes<G, CAP_NUM, kFlatS, true>();
      if (fsmem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      AMGB_LAUNCH(ctx, F_SPGEMM, bytes, kern, grid, kSpThreads, fsmem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p,
                  B.val.p, C.rp.p, C.col.p, C.val.p, (int32_t*)nullptr, ovf1.p, info.p);
    } else if ((flat & 4) && sorted) {
      auto kern = spgemm_flat_kernel<G, CAP_NUM, kFlatS, true, true, true>;
      constexpr size_t fsmem = entry_smem_bytes<G, CAP_NUM>();
      if (fsmem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      AMGB_LAUNCH(ctx, F_SPGEMM, bytes, kern, grid, kSpThreads, fsmem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p,
                  B.val.p, C.rp.p, C.col.p, C.val.p, (int32_t*)nullptr, ovf1.p, info.p);
    } else if (flat & 4) {
      auto kern = spgemm_flat_kernel<G, CAP_NUM, kFlatS, true, false, true>;
      constexpr size_t fsmem = entry_smem_bytes<G, CAP_NUM>();
      if (fsmem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      AMGB_LAUNCH(ctx, F_SPGEMM, bytes, kern, grid, kSpThreads, fsmem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p,
                  B.val.p, C.rp.p, C.col.p, C.val.p, (int32_t*)nullptr, ovf1.p, info.p);
    } else if (sorted) {
      auto kern = spgemm_numeric_kernel<G, CAP_NUM, true>;
      if (smem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMGB_LAUNCH(ctx, F_SPGEMM, bytes, kern, grid, kSpThreads, smem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p,
                  B.val.p, C.rp.p, C.col.p, C.val.p, ovf1.p, info.p);
    } else {
      auto kern = spgemm_numeric_kernel<G, CAP_NUM, false>;
      if (smem > 48 * 1024) AMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMGB_LAUNCH(ctx, F_SPGEMM, bytes, kern, grid, kSpThreads, smem, n, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p,
                  B.val.p, C.rp.p, C.col.p, C.val.p, ovf1.p, info.p);
    }
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(read_ovf(ctx, info.p, &novf, &maxv));
    if (novf > 0) {
      ctx->routes[R_SPGEMM_NUM_BIG]++;
      auto big = spgemm_numeric_big_kernel<kBigNum>;
      const size_t bsmem = (sizeof(unsigned) + sizeof(double)) * (size_t)(kSpThreads / 32) * kBigNum;
      AMGB_CUDA(ctx, cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
      AMGB_LAUNCH(ctx, F_SPGEMM, 0.0, big, (unsigned)div_up(novf, kSpThreads / 32), kSpThreads, bsmem,
                  (const int32_t*)ovf1.p, novf, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p, B.val.p, C.rp.p, C.col.p,
                  C.val.p, ovf2.p, info.p + 2);
      AMGB_CHECK_LAUNCH(ctx);
      AMGB_TRY(read_ovf(ctx, info.p + 2, &novf, &maxv));
      if (novf > 0) {
        int lgH = 5;
        while ((1ll << lgH) < 2ll * maxv) ++lgH;
        const int H = 1 << lgH;
        ctx->routes[R_SPGEMM_NUM_GLOBAL]++;
        DevBuf<unsigned> kt;
        DevBuf<double> vt;
        AMGB_TRY(kt.alloc(ctx, (size_t)fb_warps * H));
        AMGB_TRY(vt.alloc(ctx, (size_t)fb_warps * H));
        AMGB_LAUNCH(ctx, F_SPGEMM, 0.0, spgemm_numeric_global_kernel, fb_blocks, kSpThreads, 0, ovf2.p, novf, H, lgH,
                    kt.p, vt.p, A.rp.p, A.col.p, A.val.p, B.rp.p, B.col.p, B.val.p, C.rp.p, C.col.p, C.val.p);
        AMGB_CHECK_LAUNCH(ctx);
      }
    }
  }
  return AMGB_OK;
}

// sorted = false: the caller does not need ascending columns (inner operand of R*(A*P))
int spgemm(amgb_ctx* ctx, const DeviceCsr& A, const DeviceCsr& B, DeviceCsr& C, bool sorted) {
  const double avg_a = A.n > 0 ? double(A.nnz) / double(A.n) : 0.0;
  const double avg_b = B.n > 0 ? double(B.nnz) / double(B.n) : 0.0;
  if (avg_b > 8.0) {
    ctx->routes[R_SPGEMM_G32]++;
    // AMGB_SPGEMM_FLAT=2: the flattened first stage with one warp per row here too (experiment)
    const char* f32 = std::getenv("AMGB_SPGEMM_FLAT");
    return spgemm_impl<32, 1024, 512>(ctx, A, B, C, sorted, false, (f32 && f32[0] == '2') ? 3 : 0);
  }
  // short rows of B (A*P): 8 lanes per row; the table tier follows the expected row of the
  // product (about a third of the products are distinct on the FE stencils measured: 27-point
  // Poisson 120 -> 35).  Rows that outgrow their tier go to the one-warp-per-row second stage,
  // which is much slower, so wide operators (3 DoF/node elasticity, coarse levels) start higher.
  double est = avg_a * avg_b / 3.0;
  if (const char* e = std::getenv("AMGB_SPGEMM_TIER")) est = e[0] == '0' ? 0.0 : (e[0] == '1' ? 100.0 : 1000.0);
  // Row-per-thread first stage for products that may stay unsorted (the inner operand A*P): opt-in
  // (AMGB_SPGEMM_ROWTHREAD=1).  Measured on B200 at m=200, level 0: 24 ms against 18 ms for the sub-warp
  // kernels -- fully divergent global loads (one sector per lane and instruction) and two blocks per SM
  // cost more than the idle lanes of the sub-warp kernels; kept because its tables never overflow
  // into atomics and as a second implementation the parity tests compare with the oracle.
  const char* rt_env = std::getenv("AMGB_SPGEMM_ROWTHREAD");
  if (rt_env && rt_env[0] == '1' && !sorted) {
    ctx->routes[R_SPGEMM_ROWREG]++;
    return spgemm_impl<8, 256, 128>(ctx, A, B, C, sorted, true);
  }
  ctx->routes[est <= 50.0 ? R_SPGEMM_G8_T128 : (est <= 110.0 ? R_SPGEMM_G8_T256 : R_SPGEMM_G8_T512)]++;
  // First stage: the flattened kernels cost per PRODUCT (batches of 8 per row), the entry-by-entry
  // accumulation per A ENTRY (one step each); both run in the warp-uniform kernel.  Measured on
  // B200 at m = 200 (profiles/r2_spgemm_variants.md): the flattened count pass wins everywhere
  // (level 0: 3.9 against 4.9 ms for the sub-warp kernel); the flattened numeric pass wins where
  // B's rows are short or A's rows long (level 0 at theta = 0.7, 1.85 entries per P row: 18.0
  // against 18.2 ms for the level's four kernels; level 1, 46 entries per A row: 4.7 against 5.0)
  // and loses at 2.7 entries per P row (level 0 at theta = 0.25: 25.7 against 22.9 ms).
  // AMGB_SPGEMM_FLAT = 0: the sub-warp kernels of round 1; 1: flattened, both passes; 4: flattened
  // count pass + entry-by-entry numeric pass; 2: as 1, and the one-warp-per-row products too (the
  // parity tests compare every variant with the oracle).
  const char* fl_env = std::getenv("AMGB_SPGEMM_FLAT");
  int flat = 1 | ((avg_b <= 2.0 || avg_a > 40.0) ? 2 : 4);
  if (fl_env) flat = fl_env[0] == '0' ? 0 : (fl_env[0] == '4' ? 5 : 3);
  if (flat) ctx->routes[R_SPGEMM_FLAT]++;
  if (est <= 50.0) return spgemm_impl<8, 256, 128>(ctx, A, B, C, sorted, false, flat);
  if (est <= 110.0) return spgemm_impl<8, 512, 256>(ctx, A, B, C, sorted, false, flat);
  return spgemm_impl<8, 1024, 512>(ctx, A, B, C, sorted, false, flat);
}

// ---- setup stages as host functions (shared with the row-partitioned driver) ----
int run_strength(amgb_ctx* ctx, const DeviceCsr& A, double theta, double max_row_sum, uint8_t* mask,
                 int32_t* has_strong, double* diagv) {
  // stage: 32 average rows plus a quarter, at most about 100 KB per block
  const double avg = A.n > 0 ? double(A.nnz) / double(A.n) : 1.0;
  int stage = (int)(40.0 * avg) / 128 * 128 + 128;
  stage = std::min(std::max(stage, 256), 2048);
  const size_t smem = (size_t)(kStrengthBlock / 32) * stage * 13;
  if (smem > 48 * 1024)  // per device: set whenever it is needed
    AMGB_CUDA(ctx, cudaFuncSetAttribute(strength_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * 13));
  AMGB_LAUNCH(ctx, F_STRENGTH, 13.0 * A.nnz + 8.0 * A.n, strength_kernel, (unsigned)div_up(A.n, kStrengthBlock),
              kStrengthBlock, smem, A.n, A.rp.p, A.col.p, A.val.p, theta, max_row_sum, mask, has_strong, diagv, stage);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

int number_coarse_points(amgb_ctx* ctx, int64_t n, const int32_t* cf, int32_t* f2c, int32_t* n_coarse) {
  DevBuf<int32_t> flag;
  AMGB_TRY(flag.alloc(ctx, n));
  AMGB_LAUNCH(ctx, F_INTERP, 8.0 * n, cpoint_flag_kernel, (unsigned)div_up(n, kBlock), kBlock, 0, n, cf, flag.p);
  AMGB_TRY(exclusive_scan_i32(ctx, flag.p, f2c, n));
  return read_i32(ctx, f2c + n, n_coarse);
}

int build_interp(amgb_ctx* ctx, const DeviceCsr& A, const uint8_t* mask, const int32_t* cf, const int32_t* col_id,
                 const double* diagv, int64_t row_begin, int64_t row_end, int64_t n_coarse_cols, DeviceCsr& P) {
  const int64_t n = A.n;
  DevBuf<int32_t> pcount;
  AMGB_TRY(pcount.alloc(ctx, n));
  AMGB_LAUNCH(ctx, F_INTERP, 5.0 * A.nnz + 12.0 * n, interp_count_kernel, (unsigned)div_up(n, kBlock), kBlock, 0, n,
              A.rp.p, A.col.p, mask, cf, pcount.p, row_begin, row_end);
  P.n = n;
  P.ncols = n_coarse_cols;
  AMGB_TRY(P.rp.alloc(ctx, n + 1));
  AMGB_TRY(exclusive_scan_i32(ctx, pcount.p, P.rp.p, n));
  int32_t nnzp = 0;
  AMGB_TRY(read_i32(ctx, P.rp.p + n, &nnzp));
  P.nnz = nnzp;
  AMGB_TRY(P.col.alloc(ctx, nnzp));
  AMGB_TRY(P.val.alloc(ctx, nnzp));
  const int64_t rows = row_end - row_begin;
  const double fill_bytes = 13.0 * A.nnz + 12.0 * nnzp + 16.0 * n;
  if (std::getenv("AMGB_INTERP_WARP")) {  // A/B and parity aid: the warp-per-row kernel alone
    AMGB_LAUNCH(ctx, F_INTERP, fill_bytes, interp_fill_kernel, (unsigned)div_up(rows * 32, kBlock), kBlock, 0,
                row_end, A.rp.p, A.col.p, A.val.p, mask, cf, col_id, diagv, P.rp.p, P.col.p, P.val.p, row_begin,
                (const uint8_t*)nullptr);
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  }
  // compacted C-column, sign-filtered rows of the F points
  DevBuf<int32_t> account, acrp, accol, n_todo;
  DevBuf<double> acval;
  DevBuf<uint8_t> todo;
  AMGB_TRY(account.alloc(ctx, n));
  AMGB_TRY(acrp.alloc(ctx, n + 1));
  AMGB_LAUNCH(ctx, F_INTERP, 12.0 * A.nnz + 16.0 * n, interp_ac_count_kernel, (unsigned)div_up(n, kBlock), kBlock, 0, n,
              A.rp.p, A.col.p, A.val.p, cf, diagv, account.p);
  AMGB_TRY(exclusive_scan_i32(ctx, account.p, acrp.p, n));
  int32_t nnzac = 0;
  AMGB_TRY(read_i32(ctx, acrp.p + n, &nnzac));
  AMGB_TRY(accol.alloc(ctx, nnzac));
  AMGB_TRY(acval.alloc(ctx, nnzac));
  AMGB_TRY(todo.alloc_zero(ctx, n));
  AMGB_TRY(n_todo.alloc_zero(ctx, 1));
  AMGB_LAUNCH(ctx, F_INTERP, 12.0 * A.nnz + 12.0 * nnzac, interp_ac_fill_kernel, (unsigned)div_up(n, kBlock), kBlock, 0,
              n, A.rp.p, A.col.p, A.val.p, cf, diagv, (const int32_t*)acrp.p, accol.p, acval.p);
  auto group_kernel = std::getenv("AMGB_INTERP_UNIFORM") ? interp_fill_group_uniform_kernel : interp_fill_group_kernel;
  AMGB_LAUNCH(ctx, F_INTERP, fill_bytes, group_kernel, (unsigned)div_up(rows, kIgRows), kBlock, 0, row_end,
              A.rp.p, A.col.p, A.val.p, mask, cf, col_id, (const int32_t*)acrp.p, (const int32_t*)accol.p,
              (const double*)acval.p, (const int32_t*)P.rp.p, P.col.p, P.val.p, row_begin, todo.p, n_todo.p);
  AMGB_CHECK_LAUNCH(ctx);
  int32_t left = 0;
  AMGB_TRY(read_i32(ctx, n_todo.p, &left));
  if (left > 0) {  // rows with many interpolation points: the warp-per-row kernel
    AMGB_LAUNCH(ctx, F_INTERP, 0.0, interp_fill_kernel, (unsigned)div_up(rows * 32, kBlock), kBlock, 0, row_end, A.rp.p,
                A.col.p, A.val.p, mask, cf, col_id, diagv, P.rp.p, P.col.p, P.val.p, row_begin, (const uint8_t*)todo.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `todo` goes out of scope
  }
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Aggressive coarsening (levels < aggressive_coarsening_num_levels; ref t3 main.cpp:456):
// second PMIS on the distance-two strength graph S2 of the first-stage C points
// (hypre_BoomerAMGCreate2ndS, num_paths = 1) and multipass interpolation
// (hypre_BoomerAMGBuildMultipass).  Operation for operation the oracle's
// aggressive_second_pass / interp_multipass.  S2 is the pattern product A1*B
// (A1: strong rows of the C points; B: unit rows for C points, strong C neighbours for
// F points) computed with the SpGEMM above; the passes of the interpolation are sparse
// products W_k * P_(<k) whose ordered accumulation is again the SpGEMM's.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
agg_rowlen_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                  const uint8_t* __restrict__ mask, const int32_t* __restrict__ cf, const int32_t* __restrict__ f2c,
                  int32_t* __restrict__ len_a1, int32_t* __restrict__ len_b) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  int s = 0, sc = 0;
  for (int k = rp[i]; k < rp[i + 1]; ++k)
    if (mask[k]) {
      ++s;
      sc += cf[col[k]] > 0;
    }
  if (cf[i] > 0) {
    len_a1[f2c[i]] = s;
    len_b[i] = 1;
  } else {
    len_b[i] = sc;
  }
}

__global__ void __launch_bounds__(kBlock)
agg_fill_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                const uint8_t* __restrict__ mask, const int32_t* __restrict__ cf, const int32_t* __restrict__ f2c,
                const int32_t* __restrict__ a1_rp, int32_t* __restrict__ a1_col, double* __restrict__ a1_val,
                const int32_t* __restrict__ b_rp, int32_t* __restrict__ b_col, double* __restrict__ b_val) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  int w = b_rp[i];
  if (cf[i] > 0) {
    b_col[w] = f2c[i];
    b_val[w] = 1.0;
    int a = a1_rp[f2c[i]];
    for (int k = rp[i]; k < rp[i + 1]; ++k)
      if (mask[k]) {
        a1_col[a] = col[k];
        a1_val[a] = 1.0;
        ++a;
      }
  } else {
    for (int k = rp[i]; k < rp[i + 1]; ++k)
      if (mask[k] && cf[col[k]] > 0) {
        b_col[w] = f2c[col[k]];
        b_val[w] = 1.0;
        ++w;
      }
  }
}

__global__ void __launch_bounds__(kBlock)
offdiag_mask_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                    uint8_t* __restrict__ mask, int32_t* __restrict__ has_strong) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  int any = 0;
  for (int k = rp[i]; k < rp[i + 1]; ++k) {
    const uint8_t m = col[k] != (int)i;
    mask[k] = m;
    any |= m;
  }
  has_strong[i] = any;
}

__global__ void __launch_bounds__(kBlock)
correct_cf_kernel(int64_t n, int32_t* __restrict__ cf, const int32_t* __restrict__ f2c,
                  const int32_t* __restrict__ cf2) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n && cf[i] > 0) cf[i] = cf2[f2c[i]];
}

// cf (in: first-stage splitting with its numbering f2c; out: corrected splitting)
static int aggressive_second_pass(amgb_ctx* ctx, const DeviceCsr& A, const uint8_t* mask, int32_t* cf,
                                  const int32_t* f2c, int32_t nc1) {
  const int64_t n = A.n;
  const unsigned grid = (unsigned)div_up(n, kBlock);
  DeviceCsr A1, B, M;
  DevBuf<int32_t> len_a1, len_b;
  AMGB_TRY(len_a1.alloc(ctx, nc1));
  AMGB_TRY(len_b.alloc(ctx, n));
  AMGB_LAUNCH(ctx, F_COARSEN, 5.0 * A.nnz + 12.0 * n, agg_rowlen_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
              (const int32_t*)cf, f2c, len_a1.p, len_b.p);
  A1.n = nc1;
  A1.ncols = n;
  B.n = n;
  B.ncols = nc1;
  AMGB_TRY(A1.rp.alloc(ctx, nc1 + 1));
  AMGB_TRY(B.rp.alloc(ctx, n + 1));
  AMGB_TRY(exclusive_scan_i32(ctx, len_a1.p, A1.rp.p, nc1));
  AMGB_TRY(exclusive_scan_i32(ctx, len_b.p, B.rp.p, n));
  int32_t nnz_a1 = 0, nnz_b = 0;
  AMGB_TRY(read_i32(ctx, A1.rp.p + nc1, &nnz_a1));
  AMGB_TRY(read_i32(ctx, B.rp.p + n, &nnz_b));
  A1.nnz = nnz_a1;
  B.nnz = nnz_b;
  AMGB_TRY(A1.col.alloc(ctx, nnz_a1));
  AMGB_TRY(A1.val.alloc(ctx, nnz_a1));
  AMGB_TRY(B.col.alloc(ctx, nnz_b));
  AMGB_TRY(B.val.alloc(ctx, nnz_b));
  AMGB_LAUNCH(ctx, F_COARSEN, 5.0 * A.nnz + 12.0 * (nnz_a1 + nnz_b), agg_fill_kernel, grid, kBlock, 0, n, A.rp.p,
              A.col.p, mask, (const int32_t*)cf, f2c, (const int32_t*)A1.rp.p, A1.col.p, A1.val.p,
              (const int32_t*)B.rp.p, B.col.p, B.val.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_TRY(spgemm(ctx, A1, B, M, true));
  DevBuf<uint8_t> mask2;
  DevBuf<int32_t> has2, cf2;
  AMGB_TRY(mask2.alloc(ctx, M.nnz));
  AMGB_TRY(has2.alloc(ctx, nc1));
  AMGB_TRY(cf2.alloc(ctx, nc1));
  AMGB_LAUNCH(ctx, F_COARSEN, 5.0 * M.nnz + 8.0 * nc1, offdiag_mask_kernel, (unsigned)div_up(nc1, kBlock), kBlock, 0,
              (int64_t)nc1, M.rp.p, M.col.p, mask2.p, has2.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_TRY(coarsen_pmis(ctx, M, mask2.p, has2.p, cf2.p, nullptr));
  AMGB_LAUNCH(ctx, F_COARSEN, 12.0 * n, correct_cf_kernel, grid, kBlock, 0, n, cf, f2c, (const int32_t*)cf2.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

// ---- multipass interpolation ----
__global__ void __launch_bounds__(kBlock)
mp_init_kernel(int64_t n, const int32_t* __restrict__ cf, int32_t* __restrict__ pass, int32_t* __restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const int c = cf[i] > 0;
  pass[i] = c ? 0 : -1;
  len[i] = c;
}

__global__ void __launch_bounds__(kBlock)
mp_assign_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                 const uint8_t* __restrict__ mask, int32_t* __restrict__ pass, int k, int32_t* __restrict__ newly) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  int hit = 0;
  if (i < n && pass[i] < 0) {
    // neighbours of pass k-1 are final; points assigned in this sweep get k != k-1
    for (int e = rp[i]; e < rp[i + 1]; ++e)
      if (mask[e] && pass[col[e]] == k - 1) {
        hit = 1;
        break;
      }
    if (hit) pass[i] = k;
  }
  const unsigned b = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(newly, __popc(b));
}

__global__ void __launch_bounds__(kBlock)
mp_count_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                const uint8_t* __restrict__ mask, const int32_t* __restrict__ pass, int k,
                int32_t* __restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  int c = 0;
  if (pass[i] == k)
    for (int e = rp[i]; e < rp[i + 1]; ++e) c += (col[e] != (int)i && mask[e] && pass[col[e]] == k - 1);
  len[i] = c;
}

// one thread per row: the four sign sums in entry order, then the weights (oracle order)
__global__ void __launch_bounds__(kBlock)
mp_fill_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
               const double* __restrict__ val, const uint8_t* __restrict__ mask, const int32_t* __restrict__ pass,
               int k, const int32_t* __restrict__ colmap, const int32_t* __restrict__ wrp, int32_t* __restrict__ wcol,
               double* __restrict__ wval) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n || pass[i] != k) return;
  double diag = 0.0, nneg = 0.0, npos = 0.0, cneg = 0.0, cpos = 0.0;
  for (int e = rp[i]; e < rp[i + 1]; ++e) {
    const int j = col[e];
    const double v = val[e];
    if (j == (int)i) {
      diag = v;
      continue;
    }
    const bool in = mask[e] && pass[j] == k - 1;
    if (v < 0) {
      nneg = __dadd_rn(nneg, v);
      if (in) cneg = __dadd_rn(cneg, v);
    } else if (v > 0) {
      npos = __dadd_rn(npos, v);
      if (in) cpos = __dadd_rn(cpos, v);
    }
  }
  if (cpos == 0) diag = __dadd_rn(diag, npos);
  if (cneg == 0) diag = __dadd_rn(diag, nneg);
  const double alfa = (cneg != 0 && diag != 0) ? (nneg / cneg) / diag : 0.0;
  const double beta = (cpos != 0 && diag != 0) ? (npos / cpos) / diag : 0.0;
  int w = wrp[i];
  for (int e = rp[i]; e < rp[i + 1]; ++e) {
    const int j = col[e];
    if (j == (int)i || !(mask[e] && pass[j] == k - 1)) continue;
    const double v = val[e];
    wcol[w] = colmap ? colmap[j] : j;
    wval[w] = v < 0 ? __dmul_rn(-alfa, v) : __dmul_rn(-beta, v);
    ++w;
  }
}

__global__ void __launch_bounds__(kBlock)
unit_rows_kernel(int64_t n, const int32_t* __restrict__ cf, const int32_t* __restrict__ f2c,
                 const int32_t* __restrict__ prp, int32_t* __restrict__ pcol, double* __restrict__ pval) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n && cf[i] > 0) {
    pcol[prp[i]] = f2c[i];
    pval[prp[i]] = 1.0;
  }
}

__global__ void __launch_bounds__(kBlock)
merge_len_kernel(int64_t n, const int32_t* __restrict__ xrp, const int32_t* __restrict__ yrp,
                 int32_t* __restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) len[i] = (xrp[i + 1] - xrp[i]) + (yrp[i + 1] - yrp[i]);
}

__global__ void __launch_bounds__(kBlock)
merge_fill_kernel(int64_t n, const int32_t* __restrict__ xrp, const int32_t* __restrict__ xcol,
                  const double* __restrict__ xval, const int32_t* __restrict__ yrp, const int32_t* __restrict__ ycol,
                  const double* __restrict__ yval, const int32_t* __restrict__ crp, int32_t* __restrict__ ccol,
                  double* __restrict__ cval) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  int w = crp[i];
  for (int k = xrp[i]; k < xrp[i + 1]; ++k, ++w) {
    ccol[w] = xcol[k];
    cval[w] = xval[k];
  }
  for (int k = yrp[i]; k < yrp[i + 1]; ++k, ++w) {
    ccol[w] = ycol[k];
    cval[w] = yval[k];
  }
}

static int csr_from_lengths(amgb_ctx* ctx, int64_t n, int64_t ncols, const int32_t* len, DeviceCsr& M) {
  M.n = n;
  M.ncols = ncols;
  AMGB_TRY(M.rp.alloc(ctx, n + 1));
  AMGB_TRY(exclusive_scan_i32(ctx, len, M.rp.p, n));
  int32_t nnz = 0;
  AMGB_TRY(read_i32(ctx, M.rp.p + n, &nnz));
  M.nnz = nnz;
  AMGB_TRY(M.col.alloc(ctx, nnz));
  AMGB_TRY(M.val.alloc(ctx, nnz));
  return AMGB_OK;
}

static int build_interp_multipass(amgb_ctx* ctx, const DeviceCsr& A, const uint8_t* mask, const int32_t* cf,
                                  const int32_t* f2c, int64_t nc, DeviceCsr& P) {
  const int64_t n = A.n;
  const unsigned grid = (unsigned)div_up(n, kBlock);
  DevBuf<int32_t> pass, len, newly;
  AMGB_TRY(pass.alloc(ctx, n));
  AMGB_TRY(len.alloc(ctx, n));
  AMGB_TRY(newly.alloc(ctx, 1));
  AMGB_LAUNCH(ctx, F_INTERP, 12.0 * n, mp_init_kernel, grid, kBlock, 0, n, cf, pass.p, len.p);
  DeviceCsr cur;
  AMGB_TRY(csr_from_lengths(ctx, n, nc, len.p, cur));
  AMGB_LAUNCH(ctx, F_INTERP, 24.0 * nc, unit_rows_kernel, grid, kBlock, 0, n, cf, f2c, (const int32_t*)cur.rp.p,
              cur.col.p, cur.val.p);
  AMGB_CHECK_LAUNCH(ctx);
  int npass = 0;
  for (int k = 1; k < 1000; ++k) {
    AMGB_CUDA(ctx, cudaMemsetAsync(newly.p, 0, sizeof(int32_t), ctx->stream));
    AMGB_LAUNCH(ctx, F_INTERP, 5.0 * A.nnz + 8.0 * n, mp_assign_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
                pass.p, k, newly.p);
    AMGB_CHECK_LAUNCH(ctx);
    int32_t nn = 0;
    AMGB_TRY(read_i32(ctx, newly.p, &nn));
    if (nn == 0) break;
    npass = k;
  }
  for (int k = 1; k <= npass; ++k) {
    AMGB_LAUNCH(ctx, F_INTERP, 5.0 * A.nnz + 8.0 * n, mp_count_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p, mask,
                (const int32_t*)pass.p, k, len.p);
    DeviceCsr W, Pk, merged;
    AMGB_TRY(csr_from_lengths(ctx, n, k == 1 ? nc : n, len.p, W));
    AMGB_LAUNCH(ctx, F_INTERP, 13.0 * A.nnz + 12.0 * W.nnz, mp_fill_kernel, grid, kBlock, 0, n, A.rp.p, A.col.p,
                A.val.p, mask, (const int32_t*)pass.p, k, k == 1 ? f2c : (const int32_t*)nullptr,
                (const int32_t*)W.rp.p, W.col.p, W.val.p);
    AMGB_CHECK_LAUNCH(ctx);
    const DeviceCsr* add = &W;
    if (k > 1) {
      AMGB_TRY(spgemm(ctx, W, cur, Pk, true));
      add = &Pk;
    }
    AMGB_LAUNCH(ctx, F_INTERP, 16.0 * n, merge_len_kernel, grid, kBlock, 0, n, (const int32_t*)cur.rp.p,
                (const int32_t*)add->rp.p, len.p);
    AMGB_TRY(csr_from_lengths(ctx, n, nc, len.p, merged));
    AMGB_LAUNCH(ctx, F_INTERP, 24.0 * merged.nnz, merge_fill_kernel, grid, kBlock, 0, n, (const int32_t*)cur.rp.p,
                (const int32_t*)cur.col.p, (const double*)cur.val.p, (const int32_t*)add->rp.p,
                (const int32_t*)add->col.p, (const double*)add->val.p, (const int32_t*)merged.rp.p, merged.col.p,
                merged.val.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cur = std::move(merged);
  }
  P = std::move(cur);
  return AMGB_OK;
}

// deal.II forwards theta / max_row_sum to PETSc through std::to_string (6 decimals)
static double option_roundtrip(double v) { return std::strtod(std::to_string(v).c_str(), nullptr); }

static int hypre_relax_type(int dealii_type, bool symmetric_operator) {
  switch (dealii_type) {
    case AMGB_RELAX_Jacobi: return 0;
    case AMGB_RELAX_sequentialGaussSeidel: return 1;
    case AMGB_RELAX_seqboundaryGaussSeidel: return 2;
    case AMGB_RELAX_SORJacobi: return symmetric_operator ? 6 : 3;
    case AMGB_RELAX_backwardSORJacobi: return 4;
    case AMGB_RELAX_symmetricSORJacobi: return 6;
    case AMGB_RELAX_l1scaledSORJacobi: return 8;
    case AMGB_RELAX_GaussianElimination: return 9;
    case AMGB_RELAX_l1GaussSeidel: return 13;
    case AMGB_RELAX_backwardl1GaussSeidel: return 14;
    case AMGB_RELAX_CG: return 15;
    case AMGB_RELAX_Chebyshev: return 16;
    case AMGB_RELAX_FCFJacobi: return 17;
    case AMGB_RELAX_l1scaledJacobi: return 18;
    default: return -1;
  }
}

int resolve_options(amgb_precond* P) {
  amgb_ctx* ctx = P->ctx;
  const amgb_boomeramg_data& d = P->data;
  if (d.aggressive_coarsening_num_levels != 0 && P->dist)
    return set_error(ctx, AMGB_ERR_UNSUPPORTED, "aggressive coarsening is not available on the row-partitioned path");
  if (d.coarsen_type != AMGB_COARSEN_PMIS && d.coarsen_type != AMGB_COARSEN_CLJP)
    return set_error(ctx, AMGB_ERR_UNSUPPORTED,
                     "coarsen_type %d: PMIS (8) and CLJP (0) run on the device; the Ruge-Stueben passes of Falgout "
                     "(6) are sequential -- CLJP is its parallel third stage", d.coarsen_type);
  if (d.coarsen_type == AMGB_COARSEN_CLJP && (P->dist || d.aggressive_coarsening_num_levels != 0))
    return set_error(ctx, AMGB_ERR_UNSUPPORTED,
                     "CLJP coarsening: single device, without aggressive levels");
  if (d.interp_type != AMGB_INTERP_CLASSICAL)
    return set_error(ctx, AMGB_ERR_UNSUPPORTED, "interp_type %d not available", d.interp_type);
  if (d.max_levels < 1) return set_error(ctx, AMGB_ERR_BAD_ARG, "max_levels < 1");
  P->theta_eff = d.options_via_string ? option_roundtrip(d.strong_threshold) : d.strong_threshold;
  P->mrs_eff = d.options_via_string ? option_roundtrip(d.max_row_sum) : d.max_row_sum;
  const bool sym = d.symmetric_operator != 0;
  int types[3] = {hypre_relax_type(d.relaxation_type_down, sym), hypre_relax_type(d.relaxation_type_up, sym),
                  hypre_relax_type(d.relaxation_type_coarse, sym)};
  for (int t = 0; t < 3; ++t) {
    const bool jacobi_like = types[t] == 0 || types[t] == 18;
    const bool ok = jacobi_like || (t == 2 && types[t] == 9);
    if (ok) continue;
    if (types[t] == 16 && t < 2) continue;  // Chebyshev on the way down / up (amgb_cheby.cu)
    const bool sequential = types[t] == 1 || types[t] == 2 || types[t] == 3 || types[t] == 4 || types[t] == 6 ||
                            types[t] == 8 || types[t] == 13 || types[t] == 14;
    if (d.smoother_policy == AMGB_SMOOTHER_MULTICOLOR && (types[t] == 3 || types[t] == 4 || types[t] == 6)) {
      if (P->dist)
        return set_error(ctx, AMGB_ERR_UNSUPPORTED, "multicolour Gauss-Seidel is not available on the row-partitioned path");
      types[t] += 100;  // the same Gauss-Seidel sweep in multicolour order (AMGB_RELAX_MC_*)
      continue;
    }
    if (sequential && (d.smoother_policy == AMGB_SMOOTHER_SUBSTITUTE || d.smoother_policy == AMGB_SMOOTHER_MULTICOLOR)) {
      types[t] = 18;  // l1-scaled Jacobi, C/F ordered when relax_order == 1
      continue;
    }
    return set_error(ctx, AMGB_ERR_UNSUPPORTED, "relaxation type (hypre %d) not available on the device",
                     types[t]);
  }
  const bool mc_pair = types[0] >= AMGB_RELAX_MC_FORWARD && types[1] >= AMGB_RELAX_MC_FORWARD;  // (forward down, backward up)
  if (types[0] != types[1] && !mc_pair)
    return set_error(ctx, AMGB_ERR_UNSUPPORTED, "different up/down smoothers are not supported");
  if ((types[0] >= AMGB_RELAX_MC_FORWARD) != (types[1] >= AMGB_RELAX_MC_FORWARD))
    return set_error(ctx, AMGB_ERR_UNSUPPORTED, "multicolour Gauss-Seidel on the way down needs it on the way up too");
  if (types[2] >= AMGB_RELAX_MC_FORWARD && types[0] < AMGB_RELAX_MC_FORWARD)
    return set_error(ctx, AMGB_ERR_UNSUPPORTED,
                     "multicolour Gauss-Seidel on the coarsest grid only: the levels are numbered either by colour or C/F");
  P->relax_down = types[0];
  P->relax_up = types[1];
  P->relax_coarse = types[2];
  return AMGB_OK;
}

// The level loop of the single-device setup, starting from P->lv[level0].A (which must be
// set).  Also used by the row-partitioned driver for the replicated coarse levels.
// AMGB_TRACE=1: host wall time of every setup stage (stream synchronised at the stage
// boundaries), to stderr.  A diagnosis aid: allocation and host round trips show up here,
// not in the per-kernel timers.
struct StageTrace {
  amgb_ctx* ctx;
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit StageTrace(amgb_ctx* c) : ctx(c), on(std::getenv("AMGB_TRACE") != nullptr) {
    if (on) {
      cudaStreamSynchronize(ctx->stream);
      t0 = std::chrono::steady_clock::now();
    }
  }
  void mark(int level, const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[amgb trace] level %d %-12s %9.3f ms\n", level, what,
                 std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

int build_levels_from(amgb_precond* P, int level0) {
  amgb_ctx* ctx = P->ctx;
  const amgb_boomeramg_data& d = P->data;
  StageTrace trace(ctx);
  for (int level = level0;; ++level) {
    Level& L = P->lv[level];
    const int64_t n = L.A.n;
    ctx->cur_level = level;
    if (level == d.max_levels - 1 || n <= d.max_coarse_size) break;
    DevBuf<int32_t> has_strong;
    DevBuf<double> diagv;
    AMGB_TRY(L.mask.alloc(ctx, L.A.nnz));
    AMGB_TRY(has_strong.alloc(ctx, n));
    AMGB_TRY(diagv.alloc(ctx, n));
    AMGB_TRY(L.cf.alloc(ctx, n));
    trace.mark(level, "alloc");
    AMGB_TRY(run_strength(ctx, L.A, P->theta_eff, P->mrs_eff, L.mask.p, has_strong.p, diagv.p));
    trace.mark(level, "strength");
    if (d.coarsen_type == AMGB_COARSEN_CLJP) AMGB_TRY(coarsen_cljp(ctx, L.A, L.mask.p, has_strong.p, L.cf.p));
    else AMGB_TRY(coarsen_pmis(ctx, L.A, L.mask.p, has_strong.p, L.cf.p, nullptr));
    trace.mark(level, "coarsen");
    // coarse numbering: ascending fine index of the C points
    AMGB_TRY(L.f2c.alloc(ctx, n + 1));
    int32_t nc = 0;
    AMGB_TRY(number_coarse_points(ctx, n, L.cf.p, L.f2c.p, &nc));
    const bool aggressive = (unsigned)level < d.aggressive_coarsening_num_levels;
    if (aggressive && nc > 0 && nc < n) {
      AMGB_TRY(aggressive_second_pass(ctx, L.A, L.mask.p, L.cf.p, L.f2c.p, nc));
      AMGB_TRY(number_coarse_points(ctx, n, L.cf.p, L.f2c.p, &nc));
    }
    if (nc == 0 || nc == n) {
      // coarsening stalled: this level is the coarsest
      L.mask.release();
      L.cf.release();
      L.f2c.release();
      break;
    }
    L.n_coarse = nc;
    if (aggressive) AMGB_TRY(build_interp_multipass(ctx, L.A, L.mask.p, L.cf.p, L.f2c.p, nc, L.P));
    else AMGB_TRY(build_interp(ctx, L.A, L.mask.p, L.cf.p, L.f2c.p, diagv.p, 0, n, nc, L.P));
    trace.mark(level, "interp");
    AMGB_TRY(transpose_csr(ctx, L.P, L.R));
    trace.mark(level, "transpose");
    // Galerkin product A_c = R (A P)
    DeviceCsr T;
    AMGB_TRY(spgemm(ctx, L.A, L.P, T, false));
    trace.mark(level, "A*P");
    P->lv.emplace_back();
    Level& Lc = P->lv[level + 1];
    AMGB_TRY(spgemm(ctx, P->lv[level].R, T, Lc.A, true));
    trace.mark(level, "R*(AP)");
    if (!d.keep_setup_intermediates) P->lv[level].mask.release();
  }
  return AMGB_OK;
}

int build_hierarchy(amgb_precond* P) {
  amgb_ctx* ctx = P->ctx;
  AMGB_TRY(resolve_options(P));
  const amgb_boomeramg_data& d = P->data;
  const DeviceCsr& A0 = P->mat->A;
  P->lv.clear();
  P->lv.reserve(d.max_levels + 1);
  P->lv.emplace_back();
  {
    Level& L = P->lv[0];
    L.A.n = A0.n;
    L.A.ncols = A0.ncols;
    L.A.nnz = A0.nnz;
    L.A.rp.wrap(ctx, A0.rp.p, A0.rp.n);
    L.A.col.wrap(ctx, A0.col.p, A0.col.n);
    L.A.val.wrap(ctx, A0.val.p, A0.val.n);
  }
  AMGB_TRY(build_levels_from(P, 0));
  P->st_rows.clear();
  P->st_nnz.clear();
  P->st_nnzP.clear();
  for (auto& L : P->lv) {
    P->st_rows.push_back(L.A.n);
    P->st_nnz.push_back(L.A.nnz);
    P->st_nnzP.push_back(L.P.nnz);
  }
  ctx->cur_level = 0;
  StageTrace trace(ctx);
  AMGB_TRY(finish_solve_setup(P));
  trace.mark(-1, "solve setup");
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

}  // namespace amgb

using namespace amgb;

extern "C" {

int amgb_precond_initialize(amgb_ctx* ctx, const amgb_matrix* A, const amgb_boomeramg_data* data,
                            amgb_precond** out) {
  if (!ctx || !A || !data || !out) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  amgb_precond* P = new amgb_precond;
  P->ctx = ctx;
  P->mat = A;
  P->data = *data;
  if (std::getenv("AMGB_NO_GRAPH")) P->use_graph = false;  // profiling aid: plain launches
  // PCG iterations as a WHILE node of one graph: opt-in.  Measured on B200 (m=200 sweep): 0.8 % faster
  // with one system in flight, but graphs with conditional nodes do not overlap with the work of other
  // streams, which costs the 6 % that three systems in flight gain.
  P->graph_loop = std::getenv("AMGB_PCG_GRAPH_LOOP") != nullptr;
  // The context's private pool is grown ONCE to what a hierarchy of this matrix takes (about 100 bytes
  // per entry: Galerkin intermediates, CSR + SELL operators, work vectors) instead of allocation by
  // allocation: growing a pool stalls the whole device, so with several contexts running systems side
  // by side a theta sweep at m = 200 took 4.5 ... 7.4 s instead of 4.53 s (bench.py --reserve-gb 0).
  // Skipped when that is more than a third of the free memory (amgb_ctx_reserve is the explicit call).
  if (ctx->pool && !std::getenv("AMGB_NO_AUTO_RESERVE")) {
    const uint64_t want = 100ull * (uint64_t)A->A.nnz + (64ull << 20);
    uint64_t have = 0;
    if (cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &have) == cudaSuccess && have < want / 2) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && want < free_b / 3) {
        const int rrc = amgb_ctx_reserve(ctx, (int64_t)want);
        if (rrc != AMGB_OK) (void)cudaGetLastError();  // best effort: the setup allocates on demand
      }
    }
    (void)cudaGetLastError();
  }
  const int rc = build_hierarchy(P);
  if (rc != AMGB_OK) {
    cudaStreamSynchronize(ctx->stream);
    (void)cudaGetLastError();
    delete P;
    return rc;
  }
  *out = P;
  return AMGB_OK;
}

int amgb_precond_destroy(amgb_precond* P) {
  if (!P) return AMGB_OK;
  cudaSetDevice(P->ctx->device);
  destroy_solve_state(P);
  if (P->dist) amgb_dist_state_destroy(P->dist);
  P->dist = nullptr;
  delete P;
  return AMGB_OK;
}

int amgb_numbering_dealii_q1(amgb_ctx* ctx, int32_t coarse_cells, int32_t refinements, int32_t* new_to_lex) {
  if (!ctx || !new_to_lex || coarse_cells < 1 || refinements < 0 || refinements > 10) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  const int64_t m = (int64_t)coarse_cells << refinements, ncells = m * m * m, n = (m + 1) * (m + 1) * (m + 1);
  if (n >= (int64_t(1) << 31)) return set_error(ctx, AMGB_ERR_RANGE, "%lld nodes: ids are 32-bit", (long long)n);
  TreeGrid g{coarse_cells, refinements, (int)m};
  DevBuf<int32_t> count, first, out;
  AMGB_TRY(count.alloc(ctx, ncells));
  AMGB_TRY(first.alloc(ctx, ncells + 1));
  AMGB_TRY(out.alloc(ctx, n));
  const unsigned grid = (unsigned)div_up(ncells, kBlock);
  AMGB_LAUNCH(ctx, F_AUX, 4.0 * ncells, tree_count_kernel, grid, kBlock, 0, g, (long long)ncells, count.p);
  AMGB_TRY(exclusive_scan_i32(ctx, count.p, first.p, ncells));
  AMGB_LAUNCH(ctx, F_AUX, 8.0 * ncells + 4.0 * n, tree_number_kernel, grid, kBlock, 0, g, (long long)ncells,
              (const int32_t*)first.p, out.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaMemcpyAsync(new_to_lex, out.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int amgb_matrix_permute(amgb_ctx* ctx, const amgb_matrix* A, const int32_t* new_to_old, amgb_matrix** out) {
  if (!ctx || !A || !new_to_old || !out) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  DevBuf<int32_t> perm;
  AMGB_TRY(perm.alloc(ctx, A->A.n));
  AMGB_CUDA(ctx, cudaMemcpyAsync(perm.p, new_to_old, A->A.n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  amgb_matrix* M = new amgb_matrix;
  M->ctx = ctx;
  const int rc = permute_csr(ctx, A->A, perm.p, M->A);
  if (rc != AMGB_OK) {
    cudaStreamSynchronize(ctx->stream);
    delete M;
    return rc;
  }
  *out = M;
  return AMGB_OK;
}

int amgb_precond_num_levels(const amgb_precond* P, int32_t* n_levels) {
  if (!P || !n_levels) return AMGB_ERR_BAD_ARG;
  *n_levels = (int32_t)P->lv.size();
  return AMGB_OK;
}

int amgb_precond_level_stats(const amgb_precond* P, int32_t capacity, int32_t* n_levels, int64_t* rows,
                             int64_t* nnz, double* sparsity, double* grid_complexity,
                             double* operator_complexity, double* memory_complexity) {
  if (!P) return AMGB_ERR_BAD_ARG;
  const int nl = (int)P->lv.size();
  if (n_levels) *n_levels = nl;
  if (capacity < nl) return AMGB_ERR_RANGE;
  double sr = 0, sa = 0, sp = 0;
  for (int l = 0; l < nl; ++l) {
    if (rows) rows[l] = P->st_rows[l];
    if (nnz) nnz[l] = P->st_nnz[l];
    if (sparsity) sparsity[l] = double(P->st_nnz[l]) / (double(P->st_rows[l]) * double(P->st_rows[l]));
    sr += double(P->st_rows[l]);
    sa += double(P->st_nnz[l]);
    sp += double(P->st_nnzP[l]);
  }
  if (grid_complexity) *grid_complexity = sr / double(P->st_rows[0]);
  if (operator_complexity) *operator_complexity = sa / double(P->st_nnz[0]);
  if (memory_complexity) *memory_complexity = (sa + sp) / double(P->st_nnz[0]);
  return AMGB_OK;
}

int amgb_precond_get_colors(const amgb_precond* P, int32_t level, int32_t* colors, int32_t* n_colors) {
  if (!P) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->lv.size() || !P->lv[level].color.p) return AMGB_ERR_RANGE;
  const Level& L = P->lv[level];
  if (n_colors) *n_colors = (int32_t)L.color_ptr.size() - 1;
  if (colors) {
    cudaSetDevice(P->ctx->device);
    AMGB_CUDA(P->ctx, cudaMemcpyAsync(colors, L.color.p, (size_t)L.A.n * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                      P->ctx->stream));
    AMGB_CUDA(P->ctx, cudaStreamSynchronize(P->ctx->stream));
  }
  return AMGB_OK;
}

int amgb_precond_effective_relax(const amgb_precond* P, int32_t* down, int32_t* up, int32_t* coarse) {
  if (!P) return AMGB_ERR_BAD_ARG;
  if (down) *down = P->relax_down;
  if (up) *up = P->relax_up;
  // (Gaussian elimination needs a coarsest grid of at most 1024 rows; beyond that the coarsest level is
  // relaxed with the sweeps of the way down, and that is what is reported)
  if (coarse) *coarse = (P->relax_coarse == 9 && !P->dense_ok && !P->lv.empty()) ? P->relax_down : P->relax_coarse;
  return AMGB_OK;
}

int amgb_precond_level_dims(const amgb_precond* P, int32_t level, int64_t* n, int64_t* nnz_A,
                            int64_t* n_coarse, int64_t* nnz_P) {
  if (!P) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = P->lv[level];
  if (n) *n = L.A.n;
  if (nnz_A) *nnz_A = L.A.nnz;
  if (n_coarse) *n_coarse = L.P.ncols;
  if (nnz_P) *nnz_P = L.P.nnz;
  return AMGB_OK;
}

// per-level row statistics of hypre's setup printout (par_stats.c): entries per row and row sums
__global__ void __launch_bounds__(256)
row_stats_kernel(int64_t n, const int32_t* __restrict__ rp, const double* __restrict__ val, int32_t* __restrict__ imm,
                 unsigned long long* __restrict__ dmm) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int b = rp[i], e = rp[i + 1];
  double s = 0.0;
  for (int k = b; k < e; ++k) s += val[k];
  atomicMin(&imm[0], e - b);
  atomicMax(&imm[1], e - b);
  // order-preserving map of doubles to unsigned integers, so min / max are integer atomics
  unsigned long long u = (unsigned long long)__double_as_longlong(s);
  u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  atomicMin(&dmm[0], u);
  atomicMax(&dmm[1], u);
}

static int d2h(amgb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return AMGB_OK;
  AMGB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int amgb_precond_get_strength_mask(const amgb_precond* P, int32_t level, uint8_t* mask) {
  if (!P || !mask) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = P->lv[level];
  if (!L.mask.p)
    return set_error(P->ctx, AMGB_ERR_RANGE, "no strength mask kept for level %d (set keep_setup_intermediates)", level);
  cudaSetDevice(P->ctx->device);
  return d2h(P->ctx, mask, L.mask.p, L.A.nnz);
}

int amgb_precond_get_cf_marker(const amgb_precond* P, int32_t level, int32_t* cf) {
  if (!P || !cf) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = P->lv[level];
  if (!L.cf.p) return set_error(P->ctx, AMGB_ERR_RANGE, "level %d is the coarsest: no C/F splitting", level);
  cudaSetDevice(P->ctx->device);
  return d2h(P->ctx, cf, L.cf.p, L.A.n * sizeof(int32_t));
}

static int get_csr(amgb_ctx* ctx, const DeviceCsr& M, int32_t* rowptr, int32_t* col, double* val) {
  cudaSetDevice(ctx->device);
  if (rowptr) AMGB_TRY(d2h(ctx, rowptr, M.rp.p, (M.n + 1) * sizeof(int32_t)));
  if (col) AMGB_TRY(d2h(ctx, col, M.col.p, M.nnz * sizeof(int32_t)));
  if (val) AMGB_TRY(d2h(ctx, val, M.val.p, M.nnz * sizeof(double)));
  return AMGB_OK;
}

int amgb_precond_level_row_stats(const amgb_precond* P, int32_t level, int32_t* min_entries, int32_t* max_entries,
                                 double* min_row_sum, double* max_row_sum) {
  if (!P) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->lv.size()) return AMGB_ERR_RANGE;
  const DeviceCsr& A = P->lv[level].A;
  amgb_ctx* ctx = P->ctx;
  if (!A.rp.p || P->dist) return set_error(ctx, AMGB_ERR_RANGE, "level %d: operator not held in CSR form", level);
  cudaSetDevice(ctx->device);
  DevBuf<int32_t> imm;
  DevBuf<unsigned long long> dmm;
  AMGB_TRY(imm.alloc(ctx, 2));
  AMGB_TRY(dmm.alloc(ctx, 2));
  const int32_t i0[2] = {INT32_MAX, INT32_MIN};
  const unsigned long long d0[2] = {~0ull, 0ull};
  AMGB_CUDA(ctx, cudaMemcpyAsync(imm.p, i0, sizeof i0, cudaMemcpyHostToDevice, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(dmm.p, d0, sizeof d0, cudaMemcpyHostToDevice, ctx->stream));
  AMGB_LAUNCH(ctx, F_AUX, 8.0 * A.nnz + 4.0 * A.n, row_stats_kernel, (unsigned)div_up(A.n, 256), 256, 0, A.n,
              (const int32_t*)A.rp.p, (const double*)A.val.p, imm.p, dmm.p);
  AMGB_CHECK_LAUNCH(ctx);
  int32_t ih[2];
  unsigned long long dh[2];
  AMGB_TRY(d2h(ctx, ih, imm.p, sizeof ih));
  AMGB_TRY(d2h(ctx, dh, dmm.p, sizeof dh));
  auto back = [](unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    std::memcpy(&d, &u, sizeof d);
    return d;
  };
  if (min_entries) *min_entries = ih[0];
  if (max_entries) *max_entries = ih[1];
  if (min_row_sum) *min_row_sum = back(dh[0]);
  if (max_row_sum) *max_row_sum = back(dh[1]);
  return AMGB_OK;
}

int amgb_precond_get_A_csr(const amgb_precond* P, int32_t level, int32_t* rowptr, int32_t* col, double* val) {
  if (!P) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->lv.size()) return AMGB_ERR_RANGE;
  return get_csr(P->ctx, P->lv[level].A, rowptr, col, val);
}

int amgb_precond_get_P_csr(const amgb_precond* P, int32_t level, int32_t* rowptr, int32_t* col, double* val) {
  if (!P) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level + 1 >= (int)P->lv.size()) return AMGB_ERR_RANGE;
  return get_csr(P->ctx, P->lv[level].P, rowptr, col, val);
}

}  // extern "C"
