// Solve phase: SELL-32 operators in a C/F-permuted numbering, SpMV family with
// fused epilogues (plain, residual, Jacobi-type C/F half sweep, SpMV+dot,
// prolong-correct), the V(1,1) cycle driver captured in a CUDA graph, the dense
// coarsest-grid solve and PCG with device-resident scalars.
//
// Reference semantics being replaced (all inside hypre/PETSc):
//   hypre_ParCSRMatrixMatvec{,T}, hypre_BoomerAMGRelax (types 0 / 18 with C/F
//   ordering through hypre_BoomerAMGRelaxIF), hypre_BoomerAMGCycle,
//   hypre_GaussElimSolve, KSPSolve_CG driven by deal.II's SolverControl;
//   call sites ref common/amg_solver.h:48,54.
//
// Layout (DESIGN.md "Solve-phase layout"): on every level the unknowns are
// renumbered C points first (ascending), then F points (ascending).  Operators
// are stored as SELL-32: a warp owns 32 consecutive rows, entry j of the 32 rows
// is one contiguous 256 B (values) + 128 B (columns) segment, so every matrix
// load is a full-line coalesced request and there is no row reduction.  A C/F
// half sweep is then a contiguous slice range with no idle lanes, and the second
// half sweep reads the freshly relaxed set from the output buffer through a
// split-source gather (col < n_C ? x_lo : x_hi), which removes all full-vector
// copies.  All kernels are HBM-bound; algorithmic bytes per launch follow
// SURVEY.md 8(d) (defined on plain CSR) and go to the roofline report.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "amgb_dist.cuh"
#include "amgb_internal.cuh"

namespace amgb {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

// ---------------------------------------------------------------------------
// Plain CSR SpMV on the user's matrix (amgb_matrix_vmult, parity tests only).
// ---------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(kBlock)
csr_spmv_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t row = ((int64_t)blockIdx.x * kBlock + threadIdx.x) / LANES;
  const int lane = threadIdx.x % LANES;
  double s = 0.0;
  if (row < n) {
    const int b = rp[row], e = rp[row + 1];
    for (int k = b + lane; k < e; k += LANES) s += val[k] * x[col[k]];
  }
#pragma unroll
  for (int d = LANES / 2; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d, LANES);
  if (lane == 0 && row < n) y[row] = s;
}

int spmv(amgb_ctx* ctx, const DeviceCsr& A, const double* x, double* y, int family) {
  if (A.n == 0) return AMGB_OK;
  const double avg = double(A.nnz) / double(A.n);
  const double bytes = 12.0 * A.nnz + 4.0 * (A.n + 1) + 8.0 * A.ncols + 8.0 * A.n;
  if (avg <= 12.0) {
    AMGB_LAUNCH(ctx, family, bytes, csr_spmv_kernel<4>, (unsigned)div_up(A.n * 4, kBlock), kBlock, 0, A.n,
                A.rp.p, A.col.p, A.val.p, x, y);
  } else if (avg <= 48.0) {
    AMGB_LAUNCH(ctx, family, bytes, csr_spmv_kernel<8>, (unsigned)div_up(A.n * 8, kBlock), kBlock, 0, A.n,
                A.rp.p, A.col.p, A.val.p, x, y);
  } else {
    AMGB_LAUNCH(ctx, family, bytes, csr_spmv_kernel<32>, (unsigned)div_up(A.n * 32, kBlock), kBlock, 0, A.n,
                A.rp.p, A.col.p, A.val.p, x, y);
  }
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// C/F permutation and CSR -> SELL-32 conversion (once per level, end of setup).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
build_perm_kernel(int64_t n, const int32_t* __restrict__ cf, const int32_t* __restrict__ f2c, int nC,
                  int32_t* __restrict__ perm, int32_t* __restrict__ inv_perm) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  int ni = (int)i;
  if (cf) {
    const int rank = f2c[i];  // number of C points before i
    ni = cf[i] > 0 ? rank : nC + ((int)i - rank);
  }
  inv_perm[i] = ni;
  perm[ni] = (int)i;
}

// Rows of similar length next to each other (SELL-C-sigma): inside the C block and inside the F block
// of the C/F numbering, every window of kSortWindow consecutive positions is ordered by descending
// (row length of A, row length of P), ties in the order they had.  A slice then holds rows of (nearly)
// one length: on the irregular coarse operators the padding of A goes from 1.2-1.5x to 1.03x, on the
// finest level (all rows 27 long) the second key takes P's padding from 2.2x to 1.1x.  The numbering
// of the solve phase is ours to choose.  Opt-in: see finish_solve_setup_range for the measurement.
constexpr int kSortWindow = 1024;

__global__ void __launch_bounds__(kSortWindow)
perm_window_sort_kernel(int lo, int hi, const int32_t* __restrict__ arp, const int32_t* __restrict__ prp,
                        int32_t* __restrict__ perm) {
  __shared__ unsigned s_key[kSortWindow];
  __shared__ int32_t s_row[kSortWindow];
  const int t = threadIdx.x;
  const int pos = lo + blockIdx.x * kSortWindow + t;
  int row = -1;
  unsigned key = 0;  // (positions past the block's end sort last)
  if (pos < hi) {
    row = perm[pos];
    const int la = min(arp[row + 1] - arp[row], 32767);
    const int lp = prp ? min(prp[row + 1] - prp[row], 63) : 0;
    key = 0x80000000u | ((unsigned)la << 16) | ((unsigned)lp << 10) | (unsigned)(kSortWindow - 1 - t);
  }
  s_key[t] = key;
  s_row[t] = row;
  __syncthreads();
  for (int kk = 2; kk <= kSortWindow; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const int x = t ^ j;
      if (x > t) {
        const unsigned kt = s_key[t], kx = s_key[x];
        const bool desc = (t & kk) == 0;  // descending runs first: the whole window ends up descending
        if ((kt < kx) == desc) {
          s_key[t] = kx;
          s_key[x] = kt;
          const int32_t r = s_row[t];
          s_row[t] = s_row[x];
          s_row[x] = r;
        }
      }
      __syncthreads();
    }
  }
  if (pos < hi) perm[pos] = s_row[t];
}

__global__ void __launch_bounds__(kBlock)
invert_perm_kernel(int64_t n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv_perm) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) inv_perm[perm[i]] = (int)i;
}

// ---------------------------------------------------------------------------
// Multicolour Gauss-Seidel (smoother_policy = AMGB_SMOOTHER_MULTICOLOR, relax types 103 / 104 / 106):
// greedy colouring of the level's graph, operation for operation oracle/amg_oracle.cpp::multicolor.
// Rounds: every uncoloured point whose priority (a hash of the index, ties to the larger index) is the
// largest among its uncoloured neighbours takes the smallest colour none of its coloured neighbours
// has.  Two neighbours are never selected in the same round and the selection is written to `next`
// before it is applied, so the colouring does not depend on the order of the threads.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t color_priority(int64_t i) {
  uint32_t h = (uint32_t)(i + 1) * 2654435761u;
  h ^= h >> 15;
  h *= 2246822519u;
  h ^= h >> 13;
  return h;
}

__global__ void __launch_bounds__(kBlock)
color_round_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                   const int32_t* __restrict__ color, int32_t* __restrict__ next, int32_t* __restrict__ info) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  next[i] = -1;
  if (color[i] >= 0) return;
  const uint32_t pi = color_priority(i);
  bool is_max = true;
  unsigned long long used = 0ull;
  for (int k = rp[i]; k < rp[i + 1]; ++k) {
    const int j = col[k];
    if (j == (int)i) continue;
    const int cj = color[j];
    if (cj >= 0) {
      used |= 1ull << cj;
    } else {
      const uint32_t pj = color_priority(j);
      if (pj > pi || (pj == pi && j > (int)i)) is_max = false;
    }
  }
  if (!is_max) return;
  const int c = __ffsll((long long)~used) - 1;  // lowest colour not in use (-1: all 64 taken)
  if (c < 0) {
    info[1] = 1;
    return;
  }
  next[i] = c;
  atomicAdd(&info[0], 1);
  atomicMax(&info[2], c + 1);
}

__global__ void __launch_bounds__(kBlock)
color_apply_kernel(int64_t n, const int32_t* __restrict__ next, int32_t* __restrict__ color) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n && next[i] >= 0) color[i] = next[i];
}

// a row must not read a point of its own colour (it is being written by the same launch): holds by
// construction when the pattern is structurally symmetric, checked for any other operator
__global__ void __launch_bounds__(kBlock)
color_check_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                   const int32_t* __restrict__ color, int32_t* __restrict__ info) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const int ci = color[i];
  for (int k = rp[i]; k < rp[i + 1]; ++k) {
    const int j = col[k];
    if (j != (int)i && color[j] == ci) {
      info[3] = 1;
      return;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
color_flag_kernel(int64_t n, const int32_t* __restrict__ color, int c, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) flag[i] = color[i] == c ? 1 : 0;
}

__global__ void __launch_bounds__(kBlock)
color_place_kernel(int64_t n, const int32_t* __restrict__ color, int c, const int32_t* __restrict__ pos, int base,
                   int32_t* __restrict__ perm, int32_t* __restrict__ inv_perm) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n || color[i] != c) return;
  const int ni = base + pos[i];
  inv_perm[i] = ni;
  perm[ni] = (int)i;
}

// colours of level L (kept in L.color) and the solve numbering by (colour, index)
static int color_level(amgb_ctx* ctx, Level& L) {
  const int64_t n = L.A.n;
  const unsigned grid = (unsigned)div_up(n, kBlock);
  DevBuf<int32_t> next, info, pos;
  AMGB_TRY(L.color.alloc(ctx, n));
  AMGB_TRY(next.alloc(ctx, n));
  AMGB_TRY(pos.alloc(ctx, n + 1));
  AMGB_TRY(info.alloc_zero(ctx, 4));
  AMGB_CUDA(ctx, cudaMemsetAsync(L.color.p, 0xff, (size_t)n * sizeof(int32_t), ctx->stream));  // -1
  int32_t* h = (int32_t*)ctx->pinned;
  int64_t done = 0;
  int ncolors = 0;
  for (int round = 0; done < n; ++round) {
    if (round > 100000) return set_error(ctx, AMGB_ERR_BREAKDOWN, "multicolouring did not terminate");
    AMGB_LAUNCH(ctx, F_AUX, 8.0 * L.A.nnz + 12.0 * n, color_round_kernel, grid, kBlock, 0, n, (const int32_t*)L.A.rp.p,
                (const int32_t*)L.A.col.p, (const int32_t*)L.color.p, next.p, info.p);
    AMGB_LAUNCH(ctx, F_AUX, 8.0 * n, color_apply_kernel, grid, kBlock, 0, n, (const int32_t*)next.p, L.color.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_CUDA(ctx, cudaMemcpyAsync(h, info.p, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[1]) return set_error(ctx, AMGB_ERR_RANGE, "multicolouring needs more than 64 colours");
    if (h[0] == done) return set_error(ctx, AMGB_ERR_BREAKDOWN, "multicolouring stalled");
    done = h[0];
    ncolors = h[2];
  }
  AMGB_LAUNCH(ctx, F_AUX, 8.0 * L.A.nnz + 8.0 * n, color_check_kernel, grid, kBlock, 0, n, (const int32_t*)L.A.rp.p,
              (const int32_t*)L.A.col.p, (const int32_t*)L.color.p, info.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaMemcpyAsync(h, info.p, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h[3])
    return set_error(ctx, AMGB_ERR_UNSUPPORTED,
                     "multicolour Gauss-Seidel needs a structurally symmetric operator (two coupled points share a colour)");
  L.color_ptr.assign(1, 0);
  for (int c = 0; c < ncolors; ++c) {
    AMGB_LAUNCH(ctx, F_AUX, 8.0 * n, color_flag_kernel, grid, kBlock, 0, n, (const int32_t*)L.color.p, c, next.p);
    AMGB_TRY(exclusive_scan_i32(ctx, next.p, pos.p, n));
    int32_t cnt = 0;
    AMGB_TRY(read_i32(ctx, pos.p + n, &cnt));
    AMGB_LAUNCH(ctx, F_AUX, 16.0 * n, color_place_kernel, grid, kBlock, 0, n, (const int32_t*)L.color.p, c,
                (const int32_t*)pos.p, L.color_ptr.back(), L.perm.p, L.inv_perm.p);
    AMGB_CHECK_LAUNCH(ctx);
    L.color_ptr.push_back(L.color_ptr.back() + cnt);
  }
  if (L.color_ptr.back() != (int)n) return set_error(ctx, AMGB_ERR_BREAKDOWN, "multicolouring lost points");
  return AMGB_OK;
}

// one warp per slice of 32/T rows: width = ceil(longest row / T), in 32-element units
// col_lt >= 0: only the entries whose (mapped) column is < col_lt are kept, and only in rows >= col_lt
// (the F rows x C columns block of a C/F-permuted operator)
template <int T>
__global__ void __launch_bounds__(kBlock)
sell_width_kernel(int64_t n, int64_t nslices, const int32_t* __restrict__ perm,
                  const int32_t* __restrict__ rp, int32_t* __restrict__ width, const int32_t* __restrict__ col,
                  const int32_t* __restrict__ colmap, int col_lt) {
  const int64_t s = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int64_t row = s * (32 / T) + lane / T;
  int len = 0;
  if (row < n) {
    const int old = perm ? perm[row] : (int)row;
    len = rp[old + 1] - rp[old];
    if (col_lt >= 0) {
      int kept = 0;
      if (row >= col_lt && lane % T == 0)
        for (int k = rp[old]; k < rp[old] + len; ++k) kept += (colmap ? colmap[col[k]] : col[k]) < col_lt;
      len = kept;  // (the other lanes of the row contribute 0 to the max below)
    }
  }
  len = (len + T - 1) / T;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
  if (lane == 0) width[s] = len;
}

template <int T>
__global__ void __launch_bounds__(kBlock)
sell_fill_kernel(int64_t n, int64_t ncols, int64_t nslices, const int32_t* __restrict__ perm,
                 const int32_t* __restrict__ colmap, const int32_t* __restrict__ rp,
                 const int32_t* __restrict__ col, const double* __restrict__ val,
                 const int32_t* __restrict__ slice_ptr, int32_t* __restrict__ scol,
                 double* __restrict__ sval, int col_lt) {
  const int64_t s = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int64_t row = s * (32 / T) + lane / T;
  const int t = lane % T;
  int b = 0, len = 0;
  if (row < n) {
    const int old = perm ? perm[row] : (int)row;
    b = rp[old];
    len = rp[old + 1] - b;
  }
  const int w = slice_ptr[s + 1] - slice_ptr[s];
  const int64_t base = (int64_t)slice_ptr[s] * 32 + lane;
  const int pad_col = row < ncols ? (int)row : 0;
  if (col_lt >= 0) {
    // filtered block: lane t of the row walks the whole row and keeps the kept entries number t, t+T, ...
    int kept = 0, ju = 0;
    if (row >= col_lt)
      for (int k = b; k < b + len; ++k) {
        const int c = colmap ? colmap[col[k]] : col[k];
        if (c >= col_lt) continue;
        if (kept % T == t) {
          scol[base + (int64_t)ju * 32] = c;
          sval[base + (int64_t)ju * 32] = val[k];
          ++ju;
        }
        ++kept;
      }
    for (; ju < w; ++ju) {
      scol[base + (int64_t)ju * 32] = 0;
      sval[base + (int64_t)ju * 32] = 0.0;
    }
    return;
  }
  for (int ju = 0; ju < w; ++ju) {
    const int j = ju * T + t;
    int c = pad_col;
    double v = 0.0;
    if (j < len) {
      c = col[b + j];
      if (colmap) c = colmap[c];
      v = val[b + j];
    }
    scol[base + (int64_t)ju * 32] = c;
    sval[base + (int64_t)ju * 32] = v;
  }
}

// lanes per row: 1 on big regular levels (pure streaming, no reduction), more
// where rows are long or the level is too small to hide a serial row walk
static int pick_T(const DeviceCsr& A) {
  const double avg = A.n > 0 ? double(A.nnz) / double(A.n) : 1.0;
  if (A.n >= (1 << 20) && avg <= 32.0) return 1;
  if (A.n < 8192) return avg > 2.0 ? 32 : 1;
  // entries per lane: big levels have parallelism to spare, so fewer lanes share a row
  // (less padding, longer independent load batches); small levels need the lanes
  // (measured on B200 at m=200: level 1 (1.3 M rows) 3.7 -> 4.7 TB/s going from 8 to 32)
  const int per_lane = A.n >= (1 << 18) ? 32 : (A.n >= (1 << 15) ? 16 : 8);
  int T = 1;
  while (T < 32 && T * per_lane < avg) T <<= 1;
  return T;
}

#define AMGB_DISPATCH_T(T_, CALL)     \
  switch (T_) {                       \
    case 1: { constexpr int TT = 1; CALL; } break;   \
    case 2: { constexpr int TT = 2; CALL; } break;   \
    case 4: { constexpr int TT = 4; CALL; } break;   \
    case 8: { constexpr int TT = 8; CALL; } break;   \
    case 16: { constexpr int TT = 16; CALL; } break; \
    default: { constexpr int TT = 32; CALL; } break; \
  }

// col_lt >= 0: the block (rows >= col_lt) x (columns < col_lt) only, T lanes per row as given
static int csr_to_sell(amgb_ctx* ctx, const DeviceCsr& A, const int32_t* row_perm, const int32_t* colmap,
                       Sell& S, int col_lt = -1, int T_block = 1) {
  S.n = A.n;
  S.ncols = A.ncols;
  S.nnz = A.nnz;
  S.T = col_lt >= 0 ? T_block : pick_T(A);
  ctx->routes[S.T == 1 ? R_SELL_T1_STREAM : R_SELL_T_MULTI]++;
  S.nslices = div_up(A.n * S.T, 32);
  DevBuf<int32_t> width;
  AMGB_TRY(width.alloc(ctx, S.nslices));
  AMGB_TRY(S.slice_ptr.alloc(ctx, S.nslices + 1));
  const unsigned grid = (unsigned)div_up(S.nslices * 32, kBlock);
  AMGB_DISPATCH_T(S.T, AMGB_LAUNCH(ctx, F_AUX, 8.0 * A.n + (col_lt >= 0 ? 4.0 * A.nnz : 0.0), sell_width_kernel<TT>, grid,
                                   kBlock, 0, A.n, S.nslices, row_perm, A.rp.p, width.p, (const int32_t*)A.col.p, colmap,
                                   col_lt));
  AMGB_TRY(exclusive_scan_i32(ctx, width.p, S.slice_ptr.p, S.nslices));
  int32_t total = 0;
  AMGB_TRY(read_i32(ctx, S.slice_ptr.p + S.nslices, &total));
  S.padded = (int64_t)total * 32;
  AMGB_TRY(S.col.alloc(ctx, S.padded));
  AMGB_TRY(S.val.alloc(ctx, S.padded));
  AMGB_DISPATCH_T(S.T, AMGB_LAUNCH(ctx, F_AUX, 12.0 * A.nnz + 12.0 * S.padded, sell_fill_kernel<TT>, grid,
                                   kBlock, 0, A.n, A.ncols, S.nslices, row_perm, colmap, A.rp.p, A.col.p,
                                   A.val.p, S.slice_ptr.p, S.col.p, S.val.p, col_lt));
  AMGB_CHECK_LAUNCH(ctx);
  if (col_lt >= 0) {  // entries really stored, for the byte accounting (padding included: it is read)
    S.nnz = S.padded;
  }
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// SELL row kernel.  One warp per slice, T lanes per row.  The gather source is
// split: columns < split come from x_lo, the others from x_hi (the two are the
// same array for ordinary products).  Epilogue functors decide what happens to
// (row, A_row . x).
// ---------------------------------------------------------------------------
struct EpiStore {
  double* y;
  __device__ __forceinline__ void store(int row, double s) const { y[row] = s; }
};

struct EpiResidual {  // r = f - A u
  const double* f;
  double* r;
  __device__ __forceinline__ void store(int row, double s) const { r[row] = f[row] - s; }
};

struct EpiAdd {  // u += P e
  double* u;
  __device__ __forceinline__ void store(int row, double s) const { u[row] += s; }
};

struct EpiJacobi {  // out = u_old + w (f - A u) inv_relax  (hypre relax types 0 / 18)
  const double* f;
  const double* u_old;
  const double* inv_relax;
  double* out;
  double w;
  __device__ __forceinline__ void store(int row, double s) const {
    out[row] = u_old[row] + w * (f[row] - s) * inv_relax[row];
  }
};

struct EpiJacobiZero {  // the same from the zero guess: out = 0 + w (f - A_FC u_C) inv_relax
  const double* f;
  const double* inv_relax;
  double* out;
  double w;
  __device__ __forceinline__ void store(int row, double s) const {
    out[row] = 0.0 + w * (f[row] - s) * inv_relax[row];
  }
};

// Chebyshev sweep (hypre_ParCSRRelax_Cheby_Solve with scaling), u += D^-1/2 p(D^-1/2 A D^-1/2) D^-1/2 (f - A u):
// first pass r = ds (f - A u), t = ds (c_k r); middle passes q = c_i r + ds (A t), t' = ds q;
// last pass u' = u + ds (c_0 r + ds (A t)).
struct EpiChebyFirst {
  const double* f;
  const double* ds;
  double* r;
  double* t;
  double ck;
  __device__ __forceinline__ void store(int row, double s) const {
    const double d = ds[row], rr = d * (f[row] - s);
    r[row] = rr;
    t[row] = d * (ck * rr);
  }
};

struct EpiChebyMid {
  const double* r;
  const double* ds;
  double* t;
  double ci;
  __device__ __forceinline__ void store(int row, double s) const {
    const double d = ds[row];
    t[row] = d * (ci * r[row] + d * s);
  }
};

struct EpiChebyLast {
  const double* u;
  const double* r;
  const double* ds;
  double* out;
  double c0;
  __device__ __forceinline__ void store(int row, double s) const {
    const double d = ds[row];
    out[row] = u[row] + d * (c0 * r[row] + d * s);
  }
};

// STREAM: evict-first loads for operators that are far larger than L2; small
// levels use plain loads so that they stay L2-resident across V-cycles.
template <bool STREAM, class V>
__device__ __forceinline__ V ld_mat(const V* p) {
  if constexpr (STREAM) return __ldcs(p);
  else return __ldg(p);
}

// SPLIT: columns < split are gathered from x_lo, the others from x_hi (second
// C/F half sweep); otherwise a single source, one load per entry.
template <bool STREAM, bool SPLIT>
__device__ __forceinline__ double sell_lane_dot(int w, int64_t base, const int32_t* __restrict__ col,
                                                const double* __restrict__ val,
                                                const double* __restrict__ x_lo,
                                                const double* __restrict__ x_hi, int split) {
  constexpr int U = 8;  // entries in flight per lane and array
  double s0 = 0.0, s1 = 0.0;
  int j = 0;
  for (; j + U <= w; j += U) {
    const int64_t p = base + (int64_t)j * 32;
    int c[U];
    double v[U], x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = ld_mat<STREAM>(col + p + u * 32);
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_mat<STREAM>(val + p + u * 32);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* src = x_hi;
      if (SPLIT) src = c[u] < split ? x_lo : x_hi;
      x[u] = src[c[u]];
    }
#pragma unroll
    for (int u = 0; u < U; u += 2) {
      s0 += v[u] * x[u];
      s1 += v[u + 1] * x[u + 1];
    }
  }
  if (j < w) {
    // tail: the same batch, predicated, so that its loads are in flight together too
    // (27-entry stencil rows leave 3 entries here; short coarse-level rows live here entirely)
    const int64_t p = base + (int64_t)j * 32;
    const int rem = w - j;
    int c[U];
    double v[U], x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = u < rem ? ld_mat<STREAM>(col + p + u * 32) : -1;
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = u < rem ? ld_mat<STREAM>(val + p + u * 32) : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* src = x_hi;
      if (SPLIT) src = c[u] < split ? x_lo : x_hi;
      x[u] = u < rem ? src[c[u]] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; u += 2) {
      s0 += v[u] * x[u];
      s1 += v[u + 1] * x[u + 1];
    }
  }
  return s0 + s1;
}

template <int T, bool SPLIT, class Epi>
__global__ void __launch_bounds__(kBlock, 4)  // <= 64 registers: room for 8 loads in flight per array
sell_rows_kernel(int s_begin, int s_end, int row_lo, int row_hi, const int32_t* __restrict__ slice_ptr,
                 const int32_t* __restrict__ col, const double* __restrict__ val,
                 const double* __restrict__ x_lo, const double* __restrict__ x_hi, int split, Epi epi) {
  const int s = s_begin + (int)(((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5);
  if (s >= s_end) return;
  const int lane = threadIdx.x & 31;
  const int row = s * (32 / T) + lane / T;
  const bool act = row >= row_lo && row < row_hi;
  double sum = 0.0;
  if (act) {
    const int sp = slice_ptr[s];
    const int w = slice_ptr[s + 1] - sp;
    sum = sell_lane_dot<T == 1, SPLIT>(w, (int64_t)sp * 32 + lane, col, val, x_lo, x_hi, split);
  }
#pragma unroll
  for (int d = T / 2; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (act && lane % T == 0) epi.store(row, sum);
}

// w = A p fused with the partial dot (p, w): one partial per block, fixed tree.
template <int T>
__global__ void __launch_bounds__(kBlock)
sell_spmv_dot_kernel(int s_begin, int s_end, int row_lo, int row_hi, const int32_t* __restrict__ slice_ptr,
                     const int32_t* __restrict__ col, const double* __restrict__ val,
                     const double* __restrict__ x, double* __restrict__ y, double* __restrict__ partial) {
  const int s = s_begin + (int)(((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int row = s * (32 / T) + lane / T;
  const bool act = s < s_end && row >= row_lo && row < row_hi;
  double sum = 0.0;
  if (act) {
    const int sp = slice_ptr[s];
    const int w = slice_ptr[s + 1] - sp;
    sum = sell_lane_dot<T == 1, false>(w, (int64_t)sp * 32 + lane, col, val, x, x, 0);
  }
#pragma unroll
  for (int d = T / 2; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  double c = 0.0;
  if (act && lane % T == 0) {
    y[row] = sum;
    c = x[row] * sum;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  __shared__ double ws[kWarps];
  if (lane == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t += ws[i];
    partial[blockIdx.x] = t;
  }
}

// w = A p on rows [row_lo, row_hi) with one partial of (p, w) per block, written from partial[*count] on
static int launch_spmv_dot(amgb_ctx* ctx, const Sell& S, int row_lo, int row_hi, const double* x, double* y,
                           double* partial, int64_t* count, double bytes) {
  if (row_hi <= row_lo) return AMGB_OK;
  const int rps = 32 / S.T;
  const int s_begin = row_lo / rps, s_end = (int)div_up(row_hi, rps);
  const unsigned grid = (unsigned)div_up((int64_t)(s_end - s_begin) * 32, kBlock);
  AMGB_DISPATCH_T(S.T, AMGB_LAUNCH(ctx, F_SPMV, bytes, sell_spmv_dot_kernel<TT>, grid, kBlock, 0, s_begin, s_end,
                                   row_lo, row_hi, S.slice_ptr.p, S.col.p, S.val.p, x, y, partial + *count));
  AMGB_CHECK_LAUNCH(ctx);
  *count += grid;
  return AMGB_OK;
}

template <class Epi>
static int launch_sell(amgb_ctx* ctx, const Sell& S, int row_lo, int row_hi, const double* x_lo,
                       const double* x_hi, int split, Epi epi, int family, double bytes) {
  if (row_hi <= row_lo) return AMGB_OK;
  const int rps = 32 / S.T;
  const int s_begin = row_lo / rps, s_end = (int)div_up(row_hi, rps);
  const unsigned grid = (unsigned)div_up((int64_t)(s_end - s_begin) * 32, kBlock);
  if (x_lo != x_hi) {
    AMGB_DISPATCH_T(S.T, AMGB_LAUNCH(ctx, family, bytes, (sell_rows_kernel<TT, true, Epi>), grid, kBlock, 0,
                                     s_begin, s_end, row_lo, row_hi, S.slice_ptr.p, S.col.p, S.val.p, x_lo,
                                     x_hi, split, epi));
  } else {
    AMGB_DISPATCH_T(S.T, AMGB_LAUNCH(ctx, family, bytes, (sell_rows_kernel<TT, false, Epi>), grid, kBlock, 0,
                                     s_begin, s_end, row_lo, row_hi, S.slice_ptr.p, S.col.p, S.val.p, x_lo,
                                     x_hi, split, epi));
  }
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Level auxiliaries in the new numbering: 1/l1 or 1/diag for the smoother.
// ---------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kBlock)
sell_aux_kernel(int nslices, int n, const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ col,
                const double* __restrict__ val, int relax_type, double* __restrict__ inv_relax) {
  const int s = (int)(((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int row = s * (32 / T) + lane / T;
  double diag = 0.0, l1 = 0.0;
  if (row < n) {
    const int sp = slice_ptr[s];
    const int w = slice_ptr[s + 1] - sp;
    const int64_t base = (int64_t)sp * 32 + lane;
    for (int j = 0; j < w; ++j) {
      const double v = val[base + (int64_t)j * 32];
      if (col[base + (int64_t)j * 32] == row && v != 0.0) diag = v;
      l1 += fabs(v);
    }
  }
#pragma unroll
  for (int d = T / 2; d > 0; d >>= 1) {
    diag += __shfl_xor_sync(0xffffffffu, diag, d);  // exactly one lane holds the diagonal
    l1 += __shfl_xor_sync(0xffffffffu, l1, d);
  }
  if (row < n && lane % T == 0) {
    double inv = 0.0;  // rows with a zero diagonal are skipped (hypre_BoomerAMGRelax)
    if (diag != 0.0) inv = relax_type == 18 ? 1.0 / l1 : 1.0 / diag;
    inv_relax[row] = inv;
  }
}

// ---------------------------------------------------------------------------
// Row-partitioned path: which rows gather from the halo part of the vector (columns >=
// halo_begin)?  With a contiguous row partition they sit at the two ends of the C block and
// of the F block; the rows in between form the interior ranges of Sell::ilo/ihi.
//   out[0] = end of the leading halo rows of block 0, out[1] = start of its trailing ones,
//   out[2], out[3] the same for block 1 (initialised to {0, bsplit, bsplit, n}).
// ---------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kBlock)
sell_halo_rows_kernel(int nslices, int n, const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ col,
                      int halo_begin, int bsplit, int* __restrict__ out) {
  const int s = (int)(((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int row = s * (32 / T) + lane / T;
  bool touches = false;
  if (row < n) {
    const int sp = slice_ptr[s];
    const int w = slice_ptr[s + 1] - sp;
    const int64_t base = (int64_t)sp * 32 + lane;
    for (int j = 0; j < w; ++j) touches |= col[base + (int64_t)j * 32] >= halo_begin;
  }
  if (!touches) return;
  if (row < bsplit) {
    if (row < bsplit / 2) atomicMax(&out[0], row + 1); else atomicMin(&out[1], row);
  } else {
    if (row < bsplit + (n - bsplit) / 2) atomicMax(&out[2], row + 1); else atomicMin(&out[3], row);
  }
}

static int find_interior_rows(amgb_ctx* ctx, Sell& S, int halo_begin, int bsplit) {
  const int n = (int)S.n;
  DevBuf<int> out;
  AMGB_TRY(out.alloc(ctx, 4));
  const int init[4] = {0, bsplit, bsplit, n};
  AMGB_CUDA(ctx, cudaMemcpyAsync(out.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  AMGB_DISPATCH_T(S.T, AMGB_LAUNCH(ctx, F_AUX, 4.0 * S.nnz, sell_halo_rows_kernel<TT>,
                                   (unsigned)div_up(S.nslices * 32, kBlock), kBlock, 0, (int)S.nslices, n,
                                   S.slice_ptr.p, S.col.p, halo_begin, bsplit, out.p));
  AMGB_CHECK_LAUNCH(ctx);
  int h[4];
  AMGB_CUDA(ctx, cudaMemcpyAsync(h, out.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  S.bsplit = bsplit;
  S.ilo[0] = h[0];
  S.ihi[0] = std::max(h[0], h[1]);
  S.ilo[1] = h[2];
  S.ihi[1] = std::max(h[2], h[3]);
  S.has_interior = true;
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

// Large coarsest grids (coarsening stalled): one elimination step k as two
// grid-wide kernels.  Every entry still receives its updates in ascending k with
// two roundings each, so the factors equal the sequential loop bit for bit.
__global__ void __launch_bounds__(kBlock)
dense_step_factors_kernel(int n, int k, double* __restrict__ M, double* __restrict__ fac) {
  const int j = k + 1 + blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const double pivot = M[(int64_t)k * n + k];
  double f = 0.0;
  if (pivot != 0.0) {
    const double mjk = M[(int64_t)j * n + k];
    if (mjk != 0.0) {
      f = mjk / pivot;
      M[(int64_t)j * n + k] = f;
    }
  }
  fac[j] = f;
}

__global__ void __launch_bounds__(kBlock)
dense_step_update_kernel(int n, int k, double* __restrict__ M, const double* __restrict__ fac) {
  const int m = k + 1 + blockIdx.x * kBlock + threadIdx.x;
  const int j = k + 1 + blockIdx.y;
  if (m >= n) return;
  const double f = fac[j];
  if (f != 0.0)
    M[(int64_t)j * n + m] = __dsub_rn(M[(int64_t)j * n + m], __dmul_rn(f, M[(int64_t)k * n + m]));
}

// Blocked forward/back substitution with the factors of hypre_gselim: 32-wide
// diagonal blocks are solved by one warp with shuffles, the panel below/above
// by one thread per row walking the block's columns in order.  Per entry of x
// the sequence of updates (ascending k forward, descending k backward, two
// roundings each) is the sequential one.
constexpr int kDenseThreads = 512;

__global__ void __launch_bounds__(kDenseThreads)
dense_solve_kernel(int n, const double* __restrict__ M, const double* __restrict__ f,
                   double* __restrict__ x) {
  extern __shared__ double xs[];        // n entries
  __shared__ double D[32][33];          // current diagonal block
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned full = 0xffffffffu;
  for (int i = tid; i < n; i += kDenseThreads) xs[i] = f[i];
  // forward elimination: x[j] -= L[j,k] x[k], k ascending
  for (int kb = 0; kb < n; kb += 32) {
    const int kend = min(kb + 32, n), bw = kend - kb;
    for (int t = tid; t < 32 * 32; t += kDenseThreads) {
      const int r = t >> 5, c = t & 31;
      D[r][c] = (r < bw && c < bw) ? M[(int64_t)(kb + r) * n + kb + c] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double xv = lane < bw ? xs[kb + lane] : 0.0;
      for (int k = 0; k < bw; ++k) {
        const double xk = __shfl_sync(full, xv, k);
        if (D[k][k] != 0.0 && lane > k && lane < bw) {
          const double l = D[lane][k];
          if (l != 0.0) xv = __dsub_rn(xv, __dmul_rn(l, xk));
        }
      }
      if (lane < bw) xs[kb + lane] = xv;
    }
    __syncthreads();
    for (int j = kend + tid; j < n; j += kDenseThreads) {
      double lrow[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) lrow[k] = k < bw ? M[(int64_t)j * n + kb + k] : 0.0;
      double xv = xs[j];
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < bw && D[k][k] != 0.0 && lrow[k] != 0.0) xv = __dsub_rn(xv, __dmul_rn(lrow[k], xs[kb + k]));
      xs[j] = xv;
    }
    __syncthreads();
  }
  // back substitution: x[k] /= U[k,k]; x[j] -= x[k] U[j,k], k descending
  for (int khi = n; khi > 0;) {
    const int klo = max(khi - 32, 0), bw = khi - klo;
    for (int t = tid; t < 32 * 32; t += kDenseThreads) {
      const int r = t >> 5, c = t & 31;
      D[r][c] = (r < bw && c < bw) ? M[(int64_t)(klo + r) * n + klo + c] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double xv = lane < bw ? xs[klo + lane] : 0.0;
      for (int k = bw - 1; k >= 0; --k) {
        const double d = D[k][k];
        if (d != 0.0 && lane == k) xv = xv / d;
        const double xk = __shfl_sync(full, xv, k);
        if (d != 0.0 && lane < k) {
          const double u = D[lane][k];
          if (u != 0.0) xv = __dsub_rn(xv, __dmul_rn(xk, u));
        }
      }
      if (lane < bw) xs[klo + lane] = xv;
    }
    __syncthreads();
    for (int j = tid; j < klo; j += kDenseThreads) {
      double urow[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) urow[k] = k < bw ? M[(int64_t)j * n + klo + k] : 0.0;
      double xv = xs[j];
#pragma unroll
      for (int k = 31; k >= 0; --k)
        if (k < bw && D[k][k] != 0.0 && urow[k] != 0.0) xv = __dsub_rn(xv, __dmul_rn(xs[klo + k], urow[k]));
      xs[j] = xv;
    }
    __syncthreads();
    khi = klo;
  }
  for (int i = tid; i < n; i += kDenseThreads) x[i] = xs[i];
}

constexpr int64_t kMaxDenseCoarse = 1024;  // same limit as the oracle

// CA: the whole coarsest operator (on the row-partitioned path: gathered from all ranks)
static int setup_dense_from(amgb_precond* P, const DeviceCsr& CA) {
  amgb_ctx* ctx = P->ctx;
  P->dense_ok = false;
  if (P->relax_coarse != 9 || CA.n > kMaxDenseCoarse) return AMGB_OK;
  const int64_t n = CA.n;
  AMGB_TRY(P->dense.alloc_zero(ctx, n * n));
  // the coarsest level has no C/F splitting: its permutation is the identity
  AMGB_LAUNCH(ctx, F_COARSE, 12.0 * CA.nnz, csr_to_dense_kernel, (unsigned)div_up(n, 128), 128, 0, n,
              CA.rp.p, CA.col.p, CA.val.p, P->dense.p);
  if (n <= 96) {
    AMGB_LAUNCH(ctx, F_COARSE, 8.0 * n * n, dense_factor_kernel, 1, kBlock, 0, (int)n, P->dense.p);
  } else {
    ctx->routes[R_DENSE_STEPWISE]++;
    DevBuf<double> fac;
    AMGB_TRY(fac.alloc(ctx, n));
    for (int k = 0; k + 1 < (int)n; ++k) {
      const int rem = (int)n - k - 1;
      AMGB_LAUNCH(ctx, F_COARSE, 16.0 * rem, dense_step_factors_kernel, (unsigned)div_up(rem, kBlock), kBlock, 0,
                  (int)n, k, P->dense.p, fac.p);
      AMGB_LAUNCH(ctx, F_COARSE, 16.0 * rem * rem, dense_step_update_kernel,
                  dim3((unsigned)div_up(rem, kBlock), (unsigned)rem), kBlock, 0, (int)n, k, P->dense.p, fac.p);
    }
  }
  AMGB_CHECK_LAUNCH(ctx);
  P->dense_ok = true;
  return AMGB_OK;
}

static int setup_dense(amgb_precond* P) { return setup_dense_from(P, P->lv.back().A); }

int finish_solve_setup(amgb_precond* P) { return finish_solve_setup_range(P, 0); }

int finish_solve_setup_range(amgb_precond* P, int l0) {
  amgb_ctx* ctx = P->ctx;
  const int nl = (int)P->lv.size();
  // 1. permutations (C points first) on every level
  for (int l = l0; l < nl; ++l) {
    Level& L = P->lv[l];
    const int64_t n = L.A.n;
    AMGB_TRY(L.perm.alloc(ctx, n));
    AMGB_TRY(L.inv_perm.alloc(ctx, n));
    AMGB_LAUNCH(ctx, F_AUX, 16.0 * n, build_perm_kernel, (unsigned)div_up(n, kBlock), kBlock, 0, n,
                (const int32_t*)L.cf.p, (const int32_t*)L.f2c.p, (int)L.n_coarse, L.perm.p, L.inv_perm.p);
    AMGB_CHECK_LAUNCH(ctx);
  }
  // multicolour Gauss-Seidel: the solve numbering is by colour instead of C/F (the coarsest level keeps
  // the identity when it is solved directly)
  const bool mc = P->relax_down >= AMGB_RELAX_MC_FORWARD || P->relax_coarse >= AMGB_RELAX_MC_FORWARD;
  if (mc)
    for (int l = l0; l < nl; ++l) {
      if (l == nl - 1 && P->relax_coarse == 9 && P->lv[l].A.n <= kMaxDenseCoarse) continue;
      ctx->cur_level = l;
      AMGB_TRY(color_level(ctx, P->lv[l]));
    }
  // the coarsest grid is factorised first: the tail is only planned on top of a dense solve
  AMGB_TRY(setup_dense(P));
  P->tail_from = P->dense_ok ? tail_plan(P, l0) : -1;
  if (P->tail_from >= 0) AMGB_TRY(tail_pack(P));
  const int n_sell = P->tail_from >= 0 ? P->tail_from : nl;  // levels with SELL operators of their own
  // 1b. rows of similar length next to each other on the levels with SELL operators: opt-in
  // (AMGB_ROW_SORT=1).  Measured on B200 at m = 200: the padding goes away as computed (level-0
  // prolongation 162 -> 129 us) but the coarse-level sweeps are bound by their gathers, not by the
  // matrix stream, and lose what locality the window order costs (level 1: 7.14 -> 7.31 ms per solve);
  // one solve 104.5 -> 103.6 ms at theta = 0.25, 287 -> 290 ms at theta = 0.7, and 2 ms more setup.
  if (std::getenv("AMGB_ROW_SORT") && !mc)
    for (int l = l0; l < n_sell; ++l) {
      Level& L = P->lv[l];
      const int n = (int)L.A.n;
      if (n < 2 * kSortWindow || !L.cf.p) continue;
      const int nC = (int)L.n_coarse;
      const int32_t* prp = l + 1 < nl ? L.P.rp.p : nullptr;
      ctx->cur_level = l;
      AMGB_LAUNCH(ctx, F_AUX, 16.0 * nC, perm_window_sort_kernel, (unsigned)div_up(nC, kSortWindow), kSortWindow, 0, 0, nC,
                  (const int32_t*)L.A.rp.p, prp, L.perm.p);
      AMGB_LAUNCH(ctx, F_AUX, 16.0 * (n - nC), perm_window_sort_kernel, (unsigned)div_up(n - nC, kSortWindow),
                  kSortWindow, 0, nC, n, (const int32_t*)L.A.rp.p, prp, L.perm.p);
      AMGB_LAUNCH(ctx, F_AUX, 8.0 * n, invert_perm_kernel, (unsigned)div_up(n, kBlock), kBlock, 0, (int64_t)n,
                  (const int32_t*)L.perm.p, L.inv_perm.p);
      AMGB_CHECK_LAUNCH(ctx);
    }
  // 2. operators: A (rows, cols permuted), P (fine rows, coarse cols), R = P^T
  for (int l = l0; l < nl; ++l) {
    Level& L = P->lv[l];
    const int64_t n = L.A.n;
    ctx->cur_level = l;
    if (l >= n_sell) {  // tail level: lives in the packed blob; only the first one exchanges vectors
      if (l + 1 < nl) {
        L.R.rp.release();
        L.R.col.release();
        L.R.val.release();
        if (!P->data.keep_setup_intermediates) {
          L.P.rp.release();
          L.P.col.release();
          L.P.val.release();
        }
      }
      L.n_solve = L.n_vec = n;
      if (l == n_sell && l > 0) {
        AMGB_TRY(L.u.alloc(ctx, n));
        AMGB_TRY(L.f.alloc(ctx, n));
      }
      L.f2c.release();
      continue;
    }
    AMGB_TRY(csr_to_sell(ctx, L.A, L.perm.p, L.inv_perm.p, L.As));
    if (l + 1 < nl && L.n_coarse > 0 && L.n_coarse < n && P->data.relax_order == 1 &&
        (P->relax_down == 0 || P->relax_down == 18) && !std::getenv("AMGB_NO_AFC")) {
      // (the rows of the block are ~6 times shorter than those of A: a quarter of its lanes per row)
      AMGB_TRY(csr_to_sell(ctx, L.A, L.perm.p, L.inv_perm.p, L.Afc, (int)L.n_coarse, L.As.T >= 8 ? L.As.T / 4 : 1));
    }
    if (l + 1 < nl) {
      Level& C = P->lv[l + 1];
      AMGB_TRY(csr_to_sell(ctx, L.P, L.perm.p, C.inv_perm.p, L.Ps));
      AMGB_TRY(csr_to_sell(ctx, L.R, C.perm.p, L.inv_perm.p, L.Rs));
      L.R.rp.release();
      L.R.col.release();
      L.R.val.release();
      if (!P->data.keep_setup_intermediates) {
        L.P.rp.release();
        L.P.col.release();
        L.P.val.release();
      }
    }
    L.n_solve = L.n_vec = n;
    if (P->relax_down == 16) AMGB_TRY(cheby_setup_level(P, l));
    // (with Chebyshev on the way down and up, inv_relax only serves a Jacobi-type coarse relaxation)
    // scaling of the Jacobi-type relaxation of this level: the coarsest level is relaxed with the COARSE
    // type when that is a Jacobi-type one (relax_coarse may differ from relax_down: 0 vs 18); with
    // Chebyshev on the way down and up, 1/diag only serves such a coarse relaxation (or the sweeps that
    // stand in for Gaussian elimination when the coarsest grid is too large for it)
    const bool coarsest = l == nl - 1;
    int aux_type = P->relax_down == 16 ? 0 : P->relax_down;
    if (coarsest && (P->relax_coarse == 0 || P->relax_coarse == 18)) aux_type = P->relax_coarse;
    AMGB_TRY(L.inv_relax.alloc(ctx, n));
    AMGB_DISPATCH_T(L.As.T, AMGB_LAUNCH(ctx, F_AUX, L.As.csr_bytes() + 8.0 * n, sell_aux_kernel<TT>,
                                        (unsigned)div_up(L.As.nslices * 32, kBlock), kBlock, 0,
                                        (int)L.As.nslices, (int)n, L.As.slice_ptr.p, L.As.col.p, L.As.val.p,
                                        aux_type, L.inv_relax.p));
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(L.tmp.alloc(ctx, n));
    if (l > 0) {
      AMGB_TRY(L.u.alloc(ctx, n));
      AMGB_TRY(L.f.alloc(ctx, n));
    }
    L.f2c.release();
  }
  ctx->cur_level = 0;
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// V-cycle (hypre_BoomerAMGCycle): pre-smooth, residual, restrict, recurse,
// prolong-correct, post-smooth; C/F ordering per hypre_BoomerAMGRelaxIF.
// ---------------------------------------------------------------------------
// Row-partitioned path: refresh the halo part of dst with the owners' values, an owner's
// value of point p being (p < split ? lo[p] : hi[p]).  No-op on a single device.
__global__ void gather_kernel(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ in,
                              double* __restrict__ out);

static inline bool partitioned_level(const amgb_precond* P, int l) {
  return P->dist && l < P->dist->replicated_from;
}

static int halo(amgb_precond* P, int l, const double* lo, const double* hi, int split, double* dst) {
  if (!partitioned_level(P, l)) return AMGB_OK;
  P->ctx->cur_level = l;
  amgb_dist_state* ds = P->dist;
  if (ds->window_slot >= 0) {
    AMGB_TRY(peer_put(P->ctx, ds->vpeer[l], lo, hi, split));
    return peer_get(P->ctx, ds->vpeer[l], dst);
  }
  return plan_sync_split(P->ctx, ds->comm, ds->dl[l].vplan, lo, hi, split, dst);
}

// every rank's block of a vector partitioned by `starts`, on every rank: mine is full + starts[rank]
static int gather_blocks(amgb_precond* P, const std::vector<int64_t>& starts, const double* mine, double* full) {
  amgb_dist_state* ds = P->dist;
  amgb_ctx* ctx = P->ctx;
  if (ds->window_slot >= 0) {
    const int me = ds->comm->rank;
    AMGB_TRY(peer_put(ctx, ds->gather_peer, mine, mine, 0));
    double* own = full + starts[me];
    const size_t bytes = (size_t)(starts[me + 1] - starts[me]) * sizeof(double);
    if (own != mine && bytes)
      AMGB_CUDA(ctx, cudaMemcpyAsync(own, mine, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return peer_get(ctx, ds->gather_peer, full);
  }
  return allgather_f64(ctx, ds->comm, starts, mine, full);
}

static int reduce_scalars(amgb_precond* P, double* red, int count) {
  amgb_dist_state* ds = P->dist;
  if (ds->window_slot >= 0) return peer_allreduce(P->ctx, ds->red_peer, red, count);
  return ds->comm->allreduce_sum_f64(P->ctx, red, count);
}

// One hypre_BoomerAMGRelaxIF call.  Reads `u` (halo fresh), leaves the relaxed vector in
// `out` (halo stale).  On the row-partitioned path the halo of `u` is overwritten between
// the two half sweeps.
// One SELL product on rows [row_lo, row_hi) whose gather source needs its halo refreshed
// first (halo of h_dst <- the owners' p < h_split ? h_lo[p] : h_hi[p], on level `lp`).
// With peer windows on a large level the exchange is overlapped: put, rows that reference
// no halo column, wait + unpack, the remaining (boundary) rows.
template <class Epi>
static int launch_sell_halo(amgb_precond* P, int lp, const Sell& S, int row_lo, int row_hi, const double* x_lo,
                            const double* x_hi, int split, Epi epi, int family, double bytes, const double* h_lo,
                            const double* h_hi, int h_split, double* h_dst) {
  amgb_ctx* ctx = P->ctx;
  amgb_dist_state* ds = P->dist;
  const int lvl = ctx->cur_level;
  const bool overlap = partitioned_level(P, lp) && ds->window_slot >= 0 && S.has_interior &&
                       S.n >= ds->overlap_min_rows && row_hi > row_lo;
  if (!overlap) {
    AMGB_TRY(halo(P, lp, h_lo, h_hi, h_split, h_dst));
    ctx->cur_level = lvl;
    return launch_sell(ctx, S, row_lo, row_hi, x_lo, x_hi, split, epi, family, bytes);
  }
  const double per_row = bytes / double(row_hi - row_lo);
  AMGB_TRY(peer_put(ctx, ds->vpeer[lp], h_lo, h_hi, h_split));
  for (int b = 0; b < 2; ++b) {
    const int a = std::max(row_lo, S.ilo[b]), e = std::min(row_hi, S.ihi[b]);
    if (a < e) AMGB_TRY(launch_sell(ctx, S, a, e, x_lo, x_hi, split, epi, family, per_row * (e - a)));
  }
  AMGB_TRY(peer_get(ctx, ds->vpeer[lp], h_dst));
  const int cut[4] = {S.ilo[0], S.ihi[0], S.ilo[1], S.ihi[1]};
  int from = row_lo;
  for (int b = 0; b < 3; ++b) {  // the three gaps around the two interior ranges
    const int to = b < 2 ? std::min(row_hi, std::max(from, cut[2 * b])) : row_hi;
    if (from < to) AMGB_TRY(launch_sell(ctx, S, from, to, x_lo, x_hi, split, epi, family, per_row * (to - from)));
    if (b < 2) from = std::max(from, std::min(row_hi, std::max(cut[2 * b + 1], cut[2 * b])));
  }
  return AMGB_OK;
}

// Jacobi-type sweep from a zero guess on rows [lo, hi): A u = 0, so the matrix is not read.
// "0.0 +" keeps the sign of zero that the general formula produces.
__global__ void __launch_bounds__(kBlock)
jacobi_zero_kernel(int lo, int hi, const double* __restrict__ f, const double* __restrict__ inv_relax, double w,
                   double* __restrict__ out) {
  const int row = lo + (int)((int64_t)blockIdx.x * kBlock + threadIdx.x);
  if (row < hi) out[row] = 0.0 + w * (f[row] - 0.0) * inv_relax[row];
}

static int relax_zero(amgb_ctx* ctx, const Level& L, int lo, int hi, const double* f, double w, double* out) {
  if (hi <= lo) return AMGB_OK;
  AMGB_LAUNCH(ctx, F_VEC, 24.0 * (hi - lo), jacobi_zero_kernel, (unsigned)div_up(hi - lo, kBlock), kBlock, 0, lo, hi,
              f, (const double*)L.inv_relax.p, w, out);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

// first pass of a Chebyshev sweep from the zero guess: r = ds f, t = ds (c_k r); no matrix pass
__global__ void __launch_bounds__(kBlock)
cheby_zero_kernel(int n, const double* __restrict__ f, const double* __restrict__ ds, double ck,
                  double* __restrict__ r, double* __restrict__ t) {
  const int row = (int)((int64_t)blockIdx.x * kBlock + threadIdx.x);
  if (row >= n) return;
  const double d = ds[row], rr = d * (f[row] - 0.0);
  r[row] = rr;
  t[row] = d * (ck * rr);
}

__global__ void __launch_bounds__(kBlock)
add_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  const int row = (int)((int64_t)blockIdx.x * kBlock + threadIdx.x);
  if (row < n) out[row] = a[row] + b[row];
}

// One Chebyshev sweep on all points of the level (par_cycle.c: no C/F ordering for type 16)
static int relax_cheby(amgb_precond* P, int l, const double* f, double* u, double* out, bool u_is_zero) {
  Level& L = P->lv[l];
  amgb_ctx* ctx = P->ctx;
  ctx->cur_level = l;
  const int n = (int)L.n_solve;
  if (n == 0) return AMGB_OK;
  const int fam = l == 0 ? F_SMOOTH_L0 : F_SMOOTH;
  const double mat = L.As.csr_bytes();
  const int k = L.cheby_degree;
  const double* ds = L.cheby_ds.p;
  double* r = L.cheby_r.p;
  double* cur = L.cheby_t[0].p;
  double* nxt = L.cheby_t[1].p;
  const unsigned vgrid = (unsigned)div_up(n, kBlock);
  if (u_is_zero) {
    AMGB_LAUNCH(ctx, F_VEC, 32.0 * n, cheby_zero_kernel, vgrid, kBlock, 0, n, f, ds, L.cheby_coefs[k], r, cur);
    AMGB_CHECK_LAUNCH(ctx);
  } else {  // (row-partitioned path: the halo of the gather source is refreshed first)
    AMGB_TRY(launch_sell_halo(P, l, L.As, 0, n, u, u, 0, EpiChebyFirst{f, ds, r, cur, L.cheby_coefs[k]}, fam,
                              mat + 40.0 * n, u, u, 0, u));
  }
  for (int i = k - 1; i >= 1; --i) {
    AMGB_TRY(launch_sell_halo(P, l, L.As, 0, n, cur, cur, 0, EpiChebyMid{r, ds, nxt, L.cheby_coefs[i]}, fam,
                              mat + 32.0 * n, cur, cur, 0, cur));
    std::swap(cur, nxt);
  }
  if (k == 0) {
    AMGB_LAUNCH(ctx, F_VEC, 24.0 * n, add_kernel, vgrid, kBlock, 0, n, (const double*)u, (const double*)cur, out);
    AMGB_CHECK_LAUNCH(ctx);
  } else {
    AMGB_TRY(launch_sell_halo(P, l, L.As, 0, n, cur, cur, 0, EpiChebyLast{u, r, ds, out, L.cheby_coefs[0]}, fam,
                              mat + 40.0 * n, cur, cur, 0, cur));
  }
  return AMGB_OK;
}

// Gauss-Seidel in multicolour order (relax types 103 / 104 / 106; oracle relax_multicolor): the level is
// numbered by colour, a colour is a contiguous row range whose rows only read other colours, so one
// in-place Jacobi-type launch per colour IS the Gauss-Seidel sweep.  Forward = colours ascending,
// backward = descending, symmetric = both.
static int relax_multicolor(amgb_precond* P, int l, int type, const double* f, const double* u, double* out,
                            bool u_is_zero) {
  Level& L = P->lv[l];
  amgb_ctx* ctx = P->ctx;
  ctx->cur_level = l;
  const int n = (int)L.n_solve;
  if (n == 0) return AMGB_OK;
  if (u_is_zero) AMGB_CUDA(ctx, cudaMemsetAsync(out, 0, (size_t)n * sizeof(double), ctx->stream));
  else AMGB_CUDA(ctx, cudaMemcpyAsync(out, u, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  const int nc = (int)L.color_ptr.size() - 1;
  const int fam = l == 0 ? F_SMOOTH_L0 : F_SMOOTH;
  const double mat = L.As.csr_bytes();
  auto sweep = [&](int c) -> int {
    const int lo = L.color_ptr[c], hi = L.color_ptr[c + 1];
    const double share = n > 0 ? double(hi - lo) / double(n) : 0.0;
    return launch_sell(ctx, L.As, lo, hi, out, out, 0, EpiJacobi{f, out, L.inv_relax.p, out, 1.0}, fam,
                       share * mat + 32.0 * (hi - lo));
  };
  if (type != AMGB_RELAX_MC_BACKWARD)
    for (int c = 0; c < nc; ++c) AMGB_TRY(sweep(c));
  if (type != AMGB_RELAX_MC_FORWARD)
    for (int c = nc - 1; c >= 0; --c) AMGB_TRY(sweep(c));
  return AMGB_OK;
}

static int relax_if(amgb_precond* P, int l, const double* f, double* u, double* out, int cycle_param,
                    bool u_is_zero = false) {
  {
    const int type = cycle_param == 1 ? P->relax_down
                                      : (cycle_param == 2 ? P->relax_up
                                                          : (P->relax_coarse == 9 ? P->relax_down : P->relax_coarse));
    if (type >= AMGB_RELAX_MC_FORWARD) return relax_multicolor(P, l, type, f, u, out, u_is_zero);
  }
  if (P->relax_down == 16 && (cycle_param < 3 || P->relax_coarse == 9))
    return relax_cheby(P, l, f, u, out, u_is_zero);
  // on the row-partitioned path the halo of `u` is refreshed here, overlapped with the rows that do not need it
  Level& L = P->lv[l];
  amgb_ctx* ctx = P->ctx;
  ctx->cur_level = l;
  const int n = (int)L.n_solve, nC = (int)L.n_coarse;
  const double w = P->data.relax_weight;
  const int fam = l == 0 ? F_SMOOTH_L0 : F_SMOOTH;
  const double mat = L.As.csr_bytes();
  EpiJacobi epi{f, u, L.inv_relax.p, out, w};
  // C/F ordering needs a splitting on this level (any rank may own no C or no F points)
  const bool cf_order = P->data.relax_order == 1 && cycle_param < 3 &&
                       (partitioned_level(P, l) ? l + 1 < (int)P->lv.size() : (nC > 0 && nC < n));
  if (cf_order) {
    const double share_c = n > 0 ? double(nC) / double(n) : 0.0;
    // half sweep = the rows touched + 4 vectors on those rows
    const double bytes_c = share_c * mat + 32.0 * nC, bytes_f = (1.0 - share_c) * mat + 32.0 * (n - nC);
    if (cycle_param < 2) {  // down: C then F
      if (u_is_zero) AMGB_TRY(relax_zero(ctx, L, 0, nC, f, w, out));
      else AMGB_TRY(launch_sell_halo(P, l, L.As, 0, nC, u, u, 0, epi, fam, bytes_c, u, u, 0, u));
      if (u_is_zero && L.Afc.n > 0 && !partitioned_level(P, l)) {
        // from the zero guess the F columns multiply zeros: the F rows x C columns block is all there is
        AMGB_TRY(launch_sell(ctx, L.Afc, nC, n, out, out, 0, EpiJacobiZero{f, L.inv_relax.p, out, w}, fam,
                             12.0 * L.Afc.padded + 4.0 * (n - nC) + 24.0 * (n - nC) + 8.0 * nC));
        return AMGB_OK;
      }
      // halo C points: fresh; halo F points: old
      AMGB_TRY(launch_sell_halo(P, l, L.As, nC, n, out, u, nC, epi, fam, bytes_f, out, u, nC, u));
    } else {  // up: F then C
      if (u_is_zero) AMGB_TRY(relax_zero(ctx, L, nC, n, f, w, out));
      else AMGB_TRY(launch_sell_halo(P, l, L.As, nC, n, u, u, 0, epi, fam, bytes_f, u, u, 0, u));
      // halo F points: fresh; halo C points: old
      AMGB_TRY(launch_sell_halo(P, l, L.As, 0, nC, u, out, nC, epi, fam, bytes_c, u, out, nC, out));
    }
  } else if (u_is_zero) {
    AMGB_TRY(relax_zero(ctx, L, 0, n, f, w, out));
  } else {
    AMGB_TRY(launch_sell_halo(P, l, L.As, 0, n, u, u, 0, epi, fam, mat + 32.0 * n, u, u, 0, u));
  }
  return AMGB_OK;
}

// On entry `u` holds the initial guess (u_is_zero: it is identically zero, halo included),
// `alt` is scratch of the same size; on exit the result is in `u` (the two pointers are
// swapped as needed) and, on the row-partitioned path, its halo is stale.
static int cycle(amgb_precond* P, int l, double*& u, double*& alt, const double* f, bool u_is_zero) {
  amgb_ctx* ctx = P->ctx;
  Level& L = P->lv[l];
  const int nl = (int)P->lv.size();
  const int n = (int)L.n_solve;
  if (l == P->tail_from && u_is_zero) {  // this level and everything below it: one kernel
    ctx->cur_level = l;
    return tail_cycle(P, f, u);
  }
  if (l == nl - 1) {
    if (P->relax_coarse == 9 && P->dense_ok) {
      if (partitioned_level(P, l)) {
        // replicated dense solve: gather the right-hand side, solve, keep the owned block
        amgb_dist_state* ds = P->dist;
        const int nfull = (int)ds->coarse_n;
        AMGB_TRY(gather_blocks(P, ds->coarse_starts, f, ds->coarse_f.p));
        AMGB_LAUNCH(ctx, F_COARSE, 8.0 * nfull * nfull, dense_solve_kernel, 1, kDenseThreads,
                    (size_t)nfull * sizeof(double), nfull, P->dense.p, ds->coarse_f.p, ds->coarse_x.p);
        AMGB_CHECK_LAUNCH(ctx);
        if (n > 0)
          AMGB_CUDA(ctx, cudaMemcpyAsync(u, ds->coarse_x.p + ds->coarse_starts[ds->comm->rank], (size_t)n * sizeof(double),
                                         cudaMemcpyDeviceToDevice, ctx->stream));
        return AMGB_OK;
      }
      AMGB_LAUNCH(ctx, F_COARSE, 8.0 * n * n, dense_solve_kernel, 1, kDenseThreads, (size_t)n * sizeof(double), n,
                  P->dense.p, f, u);
      AMGB_CHECK_LAUNCH(ctx);
    } else {
      const unsigned sweeps = P->data.n_sweeps_coarse ? P->data.n_sweeps_coarse : 1u;
      for (unsigned s = 0; s < sweeps; ++s) {
        AMGB_TRY(relax_if(P, l, f, u, alt, 3, u_is_zero && s == 0));
        std::swap(u, alt);
      }
    }
    return AMGB_OK;
  }
  for (unsigned s = 0; s < P->data.n_sweeps; ++s) {
    AMGB_TRY(relax_if(P, l, f, u, alt, 1, u_is_zero && s == 0));
    std::swap(u, alt);
  }
  // residual into the scratch buffer, restriction into the coarse rhs
  ctx->cur_level = l;
  AMGB_TRY(launch_sell_halo(P, l, L.As, 0, n, u, u, 0, EpiResidual{f, alt}, l == 0 ? F_RESIDUAL_L0 : F_RESIDUAL,
                            L.As.csr_bytes() + 24.0 * n, u, u, 0, u));
  Level& C = P->lv[l + 1];
  const int ncrs = (int)C.n_solve;
  if (partitioned_level(P, l) && !partitioned_level(P, l + 1)) {
    // onto the first replicated level: restrict to my coarse points (natural order), gather the
    // whole right-hand side on every rank, then into that level's C/F numbering
    amgb_dist_state* ds = P->dist;
    const DistLevel& D = ds->dl[l];
    // (with peer windows the owned block is produced in place inside the full vector)
    double* own = ds->window_slot >= 0 ? ds->repl_full.p + D.cstarts[ds->comm->rank] : ds->repl_own.p;
    AMGB_TRY(launch_sell_halo(P, l, L.Rs, 0, (int)D.nc_own, alt, alt, 0, EpiStore{own},
                              l == 0 ? F_RESTRICT_L0 : F_RESTRICT, L.Rs.csr_bytes() + 8.0 * n + 8.0 * D.nc_own, alt,
                              alt, 0, alt));
    AMGB_TRY(gather_blocks(P, D.cstarts, own, ds->repl_full.p));
    AMGB_LAUNCH(ctx, F_VEC, 20.0 * ncrs, gather_kernel, (unsigned)div_up(ncrs, kBlock), kBlock, 0, (int64_t)ncrs,
                (const int32_t*)C.perm.p, (const double*)ds->repl_full.p, C.f.p);
    AMGB_CHECK_LAUNCH(ctx);
  } else {
    AMGB_TRY(launch_sell_halo(P, l, L.Rs, 0, ncrs, alt, alt, 0, EpiStore{C.f.p}, l == 0 ? F_RESTRICT_L0 : F_RESTRICT,
                              L.Rs.csr_bytes() + 8.0 * n + 8.0 * ncrs, alt, alt, 0, alt));
  }
  if (C.n_vec > 0 && l + 1 != P->tail_from)  // (the fused tail starts from the zero guess by construction)
    AMGB_CUDA(ctx, cudaMemsetAsync(C.u.p, 0, (size_t)C.n_vec * sizeof(double), ctx->stream));
  double* cu = C.u.p;
  double* calt = C.tmp.p;
  AMGB_TRY(cycle(P, l + 1, cu, calt, C.f.p, true));
  if (P->data.w_cycle && l + 1 < nl - 1) AMGB_TRY(cycle(P, l + 1, cu, calt, C.f.p, false));
  ctx->cur_level = l;
  AMGB_TRY(launch_sell_halo(P, l + 1, L.Ps, 0, n, cu, cu, 0, EpiAdd{u}, l == 0 ? F_PROLONG_L0 : F_PROLONG,
                            L.Ps.csr_bytes() + 8.0 * ncrs + 16.0 * n, cu, cu, 0, cu));
  for (unsigned s = 0; s < P->data.n_sweeps; ++s) {
    AMGB_TRY(relax_if(P, l, f, u, alt, 2));
    std::swap(u, alt);
  }
  return AMGB_OK;
}

static int vcycle_launches(amgb_precond* P, double* z, const double* r) {
  amgb_ctx* ctx = P->ctx;
  const size_t n = (size_t)P->lv[0].n_vec;
  if (n) AMGB_CUDA(ctx, cudaMemsetAsync(z, 0, n * sizeof(double), ctx->stream));
  double* u = z;
  double* alt = P->lv[0].tmp.p;
  const unsigned iters = P->data.max_iter ? P->data.max_iter : 1u;
  for (unsigned it = 0; it < iters; ++it) AMGB_TRY(cycle(P, 0, u, alt, r, it == 0));
  if (u != z && n) AMGB_CUDA(ctx, cudaMemcpyAsync(z, u, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->cur_level = 0;
  return AMGB_OK;
}

int vcycle_apply(amgb_precond* P, double* z, const double* r) {
  amgb_ctx* ctx = P->ctx;
  if (!P->use_graph || ctx->timers_on || P->capturing) return vcycle_launches(P, z, r);
  if (P->vcycle_graph && (P->graph_z != z || P->graph_r != r)) destroy_solve_state(P);
  if (!P->vcycle_graph) {
    // capture the whole cycle once: on the small levels it is launch-latency bound
    const int64_t before = ctx->launches;
    int64_t fam_before[F_COUNT];
    double bytes_before[F_COUNT];
    for (int f = 0; f < F_COUNT; ++f) {
      fam_before[f] = ctx->fam_launches[f];
      bytes_before[f] = ctx->fam_bytes[f];
    }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      (void)cudaGetLastError();
      P->use_graph = false;
      return vcycle_launches(P, z, r);
    }
    const int rc = vcycle_launches(P, z, r);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    P->graph_kernels = ctx->launches - before;
    for (int f = 0; f < F_COUNT; ++f) {
      P->graph_fam_launches[f] = ctx->fam_launches[f] - fam_before[f];
      P->graph_fam_bytes[f] = ctx->fam_bytes[f] - bytes_before[f];
      ctx->fam_launches[f] = fam_before[f];
      ctx->fam_bytes[f] = bytes_before[f];
    }
    ctx->launches = before;
    if (rc != AMGB_OK || ce != cudaSuccess || !graph ||
        cudaGraphInstantiate(&P->vcycle_graph, graph, 0) != cudaSuccess) {
      (void)cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      P->vcycle_graph = nullptr;
      P->use_graph = false;
      if (rc != AMGB_OK) return rc;
      return vcycle_launches(P, z, r);
    }
    cudaGraphDestroy(graph);
    P->graph_z = z;
    P->graph_r = r;
  }
  AMGB_CUDA(ctx, cudaGraphLaunch(P->vcycle_graph, ctx->stream));
  ctx->routes[R_CYCLE_GRAPH]++;
  ctx->launches += P->graph_kernels;
  for (int f = 0; f < F_COUNT; ++f) {
    ctx->fam_launches[f] += P->graph_fam_launches[f];
    ctx->fam_bytes[f] += P->graph_fam_bytes[f];
  }
  return AMGB_OK;
}

void destroy_solve_state(amgb_precond* P) {
  if (P->vcycle_graph) {
    cudaGraphExecDestroy(P->vcycle_graph);
    P->vcycle_graph = nullptr;
  }
  P->graph_z = nullptr;
  P->graph_r = nullptr;
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
    for (int i = 0; i < kWarps; ++i) t += ws[i];
  }
  __syncthreads();
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(kBlock)
dot2_kernel(int64_t n, const double* __restrict__ z, const double* __restrict__ r,
            double* __restrict__ pzz, double* __restrict__ pzr) {
  __shared__ double ws[kWarps];
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
  __shared__ double ws[kWarps];
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

// The PCG loop can run as a WHILE node of a CUDA graph (cg_device): the kernel that closes an
// iteration then also decides whether another one runs.
struct PcgCond {
  cudaGraphConditionalHandle handle;
  int on;  // 0: the host loop reads the flags instead
};

// beta = (z, r), dp = ||z||: history, convergence / breakdown flags, iteration counter
__device__ __forceinline__ void pcg_close_iteration(double zz, double zr, double* sc, int* fl, double* hist,
                                                    int64_t hist_cap, double abs_tol, int first, int64_t max_steps,
                                                    PcgCond c) {
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
  if (c.on) cudaGraphSetConditional(c.handle, (!fl[0] && it < max_steps) ? 1u : 0u);
}

__global__ void __launch_bounds__(kBlock)
finalize_beta_kernel(const double* __restrict__ pzz, const double* __restrict__ pzr, int64_t m,
                     double* sc, int* fl, double* hist, int64_t hist_cap, double abs_tol, int first,
                     int64_t max_steps, PcgCond c) {
  __shared__ double ws[kWarps];
  const double zz = ordered_sum(pzz, m, ws);
  const double zr = ordered_sum(pzr, m, ws);
  if (threadIdx.x == 0) pcg_close_iteration(zz, zr, sc, fl, hist, hist_cap, abs_tol, first, max_steps, c);
}

// Row-partitioned path: the block partials are summed locally (same fixed tree), the 1-2
// local sums are all-reduced over the ranks, and the scalar updates read the global sums.
__global__ void __launch_bounds__(kBlock)
local_sums_kernel(const double* __restrict__ p1, const double* __restrict__ p2, int64_t m, double* __restrict__ red) {
  __shared__ double ws[kWarps];
  const double a = ordered_sum(p1, m, ws);
  const double b = p2 ? ordered_sum(p2, m, ws) : 0.0;
  if (threadIdx.x == 0) {
    red[0] = a;
    red[1] = b;
  }
}

__global__ void finalize_alpha_red_kernel(const double* __restrict__ red, double* sc, int* fl) {
  const double pw = red[0];
  sc[2] = pw;
  if (pw == 0.0 || pw != pw) {
    fl[0] = 1;
    fl[2] = AMGB_ERR_BREAKDOWN;
    sc[3] = 0.0;
  } else {
    sc[3] = sc[0] / pw;
  }
}

__global__ void finalize_beta_red_kernel(const double* __restrict__ red, double* sc, int* fl, double* hist,
                                         int64_t hist_cap, double abs_tol, int first, int64_t max_steps, PcgCond c) {
  pcg_close_iteration(red[0], red[1], sc, fl, hist, hist_cap, abs_tol, first, max_steps, c);
}

// The two streaming vector updates of a PCG step: two elements per thread as 16-byte accesses,
// several blocks' worth per thread (grid-stride), so enough loads are in flight per SM.
__global__ void __launch_bounds__(kBlock)
update_p_kernel(int64_t n, const double* __restrict__ z, double* __restrict__ p,
                const double* __restrict__ sc, const int* __restrict__ fl) {
  const bool first = fl[1] == 0;  // first step: no previous direction
  const double bb = first ? 0.0 : sc[0] / sc[1];
  const int64_t n2 = n >> 1;
  const double2* __restrict__ z2 = reinterpret_cast<const double2*>(z);
  double2* __restrict__ p2 = reinterpret_cast<double2*>(p);
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    const double2 zz = z2[i];
    double2 pp = p2[i];
    if (first) {
      pp = zz;
    } else {
      pp.x = zz.x + bb * pp.x;
      pp.y = zz.y + bb * pp.y;
    }
    p2[i] = pp;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = first ? z[n - 1] : z[n - 1] + bb * p[n - 1];
}

__global__ void __launch_bounds__(kBlock)
axpy2_kernel(int64_t n, const double* __restrict__ p, const double* __restrict__ w,
             double* __restrict__ x, double* __restrict__ r, const double* __restrict__ sc) {
  const double alpha = sc[3];
  const int64_t n2 = n >> 1;
  const double2* __restrict__ p2 = reinterpret_cast<const double2*>(p);
  const double2* __restrict__ w2 = reinterpret_cast<const double2*>(w);
  double2* __restrict__ x2 = reinterpret_cast<double2*>(x);
  double2* __restrict__ r2 = reinterpret_cast<double2*>(r);
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    const double2 pp = p2[i], ww = w2[i];
    double2 xx = x2[i], rr = r2[i];
    xx.x += alpha * pp.x;
    xx.y += alpha * pp.y;
    rr.x -= alpha * ww.x;
    rr.y -= alpha * ww.y;
    x2[i] = xx;
    r2[i] = rr;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    x[n - 1] += alpha * p[n - 1];
    r[n - 1] -= alpha * w[n - 1];
  }
}

// out[k] = in[perm[k]]  /  out[perm[k]] = in[k]
__global__ void __launch_bounds__(kBlock)
gather_kernel(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ in,
              double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k < n) out[k] = in[perm[k]];
}

__global__ void __launch_bounds__(kBlock)
scatter_kernel(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ in,
               double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k < n) out[perm[k]] = in[k];
}

struct PcgFlags {
  int done, iters, status, pad;
};

// Records  prologue -> WHILE(cond) { body } -> epilogue  as one graph and launches it on the
// context's stream.  The body is captured on a second stream into the body graph of the
// conditional node (a stream carries one capture at a time), with the context's stream
// pointer swapped for the duration, so every launch helper records into the right graph.
// *ran = false (and nothing executed) when the graph could not be built: the caller then
// runs the host loop.
template <class Pro, class Body, class Epi>
static int pcg_graph_loop(amgb_ctx* ctx, amgb_precond* P, PcgCond& cond, Pro&& prologue, Body&& body, Epi&& epilogue,
                          bool* ran) {
  *ran = false;
  if (!ctx->stream2 && cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) {
    (void)cudaGetLastError();
    P->graph_loop = false;
    return AMGB_OK;
  }
  cudaStream_t main_stream = ctx->stream;
  cudaGraph_t G = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool outer_open = false, inner_open = false;
  int rc = AMGB_OK;
  bool ok = cudaGraphCreate(&G, 0) == cudaSuccess &&
            cudaGraphConditionalHandleCreate(&cond.handle, G, 0, 0) == cudaSuccess;
  cond.on = 1;
  const bool was_capturing = P->capturing;
  P->capturing = true;  // the cycle is recorded inline, not as its own executable graph
  if (ok) ok = outer_open = cudaStreamBeginCaptureToGraph(main_stream, G, nullptr, nullptr, 0,
                                                          cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  if (ok) rc = prologue();
  if (ok && rc == AMGB_OK) {
    cudaStreamCaptureStatus st;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    cudaGraph_t cg = nullptr;
    ok = cudaStreamGetCaptureInfo(main_stream, &st, nullptr, &cg, &deps, &ndeps) == cudaSuccess &&
         st == cudaStreamCaptureStatusActive;
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = cond.handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t cnode = nullptr;
    if (ok) ok = cudaGraphAddNode(&cnode, G, deps, ndeps, &np) == cudaSuccess;
    if (ok) ok = cudaStreamUpdateCaptureDependencies(main_stream, &cnode, 1, cudaStreamSetCaptureDependencies) ==
                 cudaSuccess;
    if (ok) {
      cudaGraph_t bg = np.conditional.phGraph_out[0];
      const int64_t l0 = ctx->launches;
      int64_t fam_l0[F_COUNT];
      double fam_b0[F_COUNT];
      for (int f = 0; f < F_COUNT; ++f) {
        fam_l0[f] = ctx->fam_launches[f];
        fam_b0[f] = ctx->fam_bytes[f];
      }
      ctx->stream = ctx->stream2;
      ok = inner_open = cudaStreamBeginCaptureToGraph(ctx->stream, bg, nullptr, nullptr, 0,
                                                      cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) rc = body();
      if (inner_open) {
        cudaGraph_t got = nullptr;
        if (cudaStreamEndCapture(ctx->stream, &got) != cudaSuccess) ok = false;
        inner_open = false;
      }
      ctx->stream = main_stream;
      P->loop_kernels = ctx->launches - l0;
      for (int f = 0; f < F_COUNT; ++f) {
        P->loop_fam_launches[f] = ctx->fam_launches[f] - fam_l0[f];
        P->loop_fam_bytes[f] = ctx->fam_bytes[f] - fam_b0[f];
      }
    }
    if (ok && rc == AMGB_OK) rc = epilogue();
  }
  if (outer_open) {
    cudaGraph_t got = nullptr;
    if (cudaStreamEndCapture(main_stream, &got) != cudaSuccess) ok = false;
  }
  P->capturing = was_capturing;
  if (ok && rc == AMGB_OK) ok = cudaGraphInstantiate(&exec, G, 0) == cudaSuccess;
  if (ok && rc == AMGB_OK) ok = cudaGraphLaunch(exec, main_stream) == cudaSuccess;
  if (ok && rc == AMGB_OK) {
    const cudaError_t e = cudaStreamSynchronize(main_stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "PCG graph", __FILE__, __LINE__);
    *ran = true;
  } else {
    (void)cudaGetLastError();
    if (rc == AMGB_OK) P->graph_loop = false;  // graph API refused: host loop from now on
  }
  if (exec) cudaGraphExecDestroy(exec);
  if (G) cudaGraphDestroy(G);
  return rc;
}

// History entries kept on the device: PCG needs tens of iterations; a buffer of max_steps + 1
// = n + 1 entries (SolverControl(n, tol), ref amg_solver.h:33) would cost a 65 MB allocation and
// memset inside the timed solve at m=200.
constexpr int64_t kMaxHistory = 1 << 16;

// x, b in the USER numbering (device; on the row-partitioned path: the owned slab);
// everything inside runs in the permuted one.
//
// The iteration is either driven by the host (one 16-byte flag read per step), or -- whenever
// the V-cycle is capturable -- recorded ONCE as a CUDA graph
//     prologue -> WHILE(cond) { one PCG step } -> scatter
// whose condition is set on the device by the kernel that closes a step
// (pcg_close_iteration): no host round trip between the first residual and the solution.
static int cg_device(amgb_ctx* ctx, const amgb_matrix* A, double* x_user, const double* b_user,
                     amgb_precond* P, int64_t max_steps, double abs_tol, double* res_hist, int64_t hist_cap,
                     int64_t* n_iters) {
  if (P->lv.empty()) return set_error(ctx, AMGB_ERR_BAD_ARG, "preconditioner is not initialised");
  amgb_dist_state* ds = P->dist;
  if (!ds && (!A || P->lv[0].A.n != A->A.n || P->mat != A))
    return set_error(ctx, AMGB_ERR_BAD_ARG, "preconditioner was initialised for another matrix");
  Level& L0 = P->lv[0];
  const int64_t n = L0.n_solve, nv = L0.n_vec;  // rows here, vector length (rows + halo)
  const Sell& As = L0.As;
  DevBuf<double> x, b, r, z, p, w, sc, hist, pa, pb;
  DevBuf<int> fl;
  AMGB_TRY(x.alloc(ctx, nv));
  AMGB_TRY(b.alloc(ctx, nv));
  AMGB_TRY(r.alloc(ctx, nv));
  AMGB_TRY(z.alloc(ctx, nv));
  AMGB_TRY(p.alloc(ctx, nv));
  AMGB_TRY(w.alloc(ctx, nv));
  AMGB_TRY(sc.alloc_zero(ctx, 8));
  AMGB_TRY(fl.alloc_zero(ctx, 4));
  int64_t cap = hist_cap > 0 && res_hist ? hist_cap : 1;
  if (cap > max_steps + 1) cap = max_steps + 1;
  if (cap > kMaxHistory) cap = kMaxHistory;
  if (cap < 1) cap = 1;
  AMGB_TRY(hist.alloc_zero(ctx, cap));
  const int64_t dot_blocks = div_up(n, kDotChunk);
  const int64_t spmv_blocks = div_up(As.nslices * 32, kBlock);
  // (+16: the overlapped product is launched in up to five row ranges, each rounding up to whole blocks)
  AMGB_TRY(pa.alloc(ctx, (spmv_blocks > dot_blocks ? spmv_blocks : dot_blocks) + 16));
  AMGB_TRY(pb.alloc(ctx, dot_blocks + 1));
  const unsigned vgrid = (unsigned)div_up(n, kBlock);
  // streaming updates: two elements per thread, at most 16 blocks per SM (grid-stride beyond that)
  const unsigned sgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(div_up(n / 2 + 1, kBlock), (int64_t)ctx->sm_count * 16));
  double* red = ds ? ds->red.p : nullptr;
  PcgCond cond{};
  // the captured cycle belongs to this call's (z, r): dropped on every way out
  struct SolveStateGuard {
    amgb_precond* P;
    ~SolveStateGuard() { destroy_solve_state(P); }
  } guard{P};

  auto finalize_beta = [&](int first) -> int {
    if (ds) {
      AMGB_LAUNCH(ctx, F_VEC, 16.0 * dot_blocks, local_sums_kernel, 1, kBlock, 0, (const double*)pa.p,
                  (const double*)pb.p, dot_blocks, red);
      AMGB_CHECK_LAUNCH(ctx);
      AMGB_TRY(reduce_scalars(P, red, 2));
      AMGB_LAUNCH(ctx, F_VEC, 16.0, finalize_beta_red_kernel, 1, 1, 0, (const double*)red, sc.p, fl.p, hist.p, cap,
                  abs_tol, first, max_steps, cond);
    } else {
      AMGB_LAUNCH(ctx, F_VEC, 16.0 * dot_blocks, finalize_beta_kernel, 1, kBlock, 0, pa.p, pb.p, dot_blocks, sc.p,
                  fl.p, hist.p, cap, abs_tol, first, max_steps, cond);
    }
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  };
  // r = b - A x ; z = M^{-1} r ; dp = ||z|| ; beta = (z, r)
  auto prologue = [&]() -> int {
    AMGB_LAUNCH(ctx, F_VEC, 20.0 * n, gather_kernel, vgrid, kBlock, 0, n, L0.perm.p, x_user, x.p);
    AMGB_LAUNCH(ctx, F_VEC, 20.0 * n, gather_kernel, vgrid, kBlock, 0, n, L0.perm.p, b_user, b.p);
    AMGB_TRY(launch_sell_halo(P, 0, As, 0, (int)n, x.p, x.p, 0, EpiResidual{b.p, r.p}, F_RESIDUAL_L0,
                              As.csr_bytes() + 24.0 * n, x.p, x.p, 0, x.p));
    AMGB_TRY(vcycle_apply(P, z.p, r.p));
    AMGB_LAUNCH(ctx, F_VEC, 16.0 * n, dot2_kernel, (unsigned)dot_blocks, kBlock, 0, n, z.p, r.p, pa.p, pb.p);
    return finalize_beta(1);
  };
  // one PCG step: p, w = A p, alpha, x, r, z = M^{-1} r, beta
  auto body = [&]() -> int {
    AMGB_LAUNCH(ctx, F_VEC, 24.0 * n, update_p_kernel, sgrid, kBlock, 0, n, z.p, p.p, sc.p, (const int*)fl.p);
    int64_t np = 0;  // partials of (p, w)
    const double spmv_bytes = As.csr_bytes() + 16.0 * n;
    if (ds && ds->window_slot >= 0 && As.has_interior && As.n >= ds->overlap_min_rows && n > 0) {
      // halo of p in flight while the rows that do not gather from it are multiplied
      const double per_row = spmv_bytes / double(n);
      AMGB_TRY(peer_put(ctx, ds->vpeer[0], p.p, p.p, 0));
      int edge[6] = {0, As.ilo[0], As.ihi[0], As.ilo[1], As.ihi[1], (int)n};
      for (int k = 1; k < 6; ++k) edge[k] = std::max(edge[k], edge[k - 1]);
      for (int k = 1; k < 5; k += 2)
        AMGB_TRY(launch_spmv_dot(ctx, As, edge[k], edge[k + 1], p.p, w.p, pa.p, &np, per_row * (edge[k + 1] - edge[k])));
      AMGB_TRY(peer_get(ctx, ds->vpeer[0], p.p));
      for (int k = 0; k < 6; k += 2)
        AMGB_TRY(launch_spmv_dot(ctx, As, edge[k], edge[k + 1], p.p, w.p, pa.p, &np, per_row * (edge[k + 1] - edge[k])));
    } else {
      AMGB_TRY(halo(P, 0, p.p, p.p, 0, p.p));
      AMGB_TRY(launch_spmv_dot(ctx, As, 0, (int)n, p.p, w.p, pa.p, &np, spmv_bytes));
    }
    if (ds) {
      AMGB_LAUNCH(ctx, F_VEC, 8.0 * np, local_sums_kernel, 1, kBlock, 0, (const double*)pa.p,
                  (const double*)nullptr, np, red);
      AMGB_CHECK_LAUNCH(ctx);
      AMGB_TRY(reduce_scalars(P, red, 1));
      AMGB_LAUNCH(ctx, F_VEC, 8.0, finalize_alpha_red_kernel, 1, 1, 0, (const double*)red, sc.p, fl.p);
    } else {
      AMGB_LAUNCH(ctx, F_VEC, 8.0 * np, finalize_alpha_kernel, 1, kBlock, 0, pa.p, np, sc.p, fl.p);
    }
    AMGB_LAUNCH(ctx, F_VEC, 48.0 * n, axpy2_kernel, sgrid, kBlock, 0, n, p.p, w.p, x.p, r.p, sc.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(vcycle_apply(P, z.p, r.p));
    AMGB_LAUNCH(ctx, F_VEC, 16.0 * n, dot2_kernel, (unsigned)dot_blocks, kBlock, 0, n, z.p, r.p, pa.p, pb.p);
    return finalize_beta(0);
  };
  auto epilogue = [&]() -> int {
    AMGB_LAUNCH(ctx, F_VEC, 20.0 * n, scatter_kernel, vgrid, kBlock, 0, n, L0.perm.p, x.p, x_user);
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  };
  PcgFlags* hf = (PcgFlags*)ctx->pinned;
  auto read_flags = [&]() -> int {
    AMGB_CUDA(ctx, cudaMemcpyAsync(hf, fl.p, sizeof(PcgFlags), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AMGB_OK;
  };

  bool done_by_graph = false;
  if (P->use_graph && P->graph_loop && !ctx->timers_on) {
    int rc = pcg_graph_loop(ctx, P, cond, prologue, body, epilogue, &done_by_graph);
    if (rc != AMGB_OK) return rc;
    if (done_by_graph) {
      AMGB_TRY(read_flags());
      // the launch counters saw the captured step once; it ran hf->iters times
      const int64_t extra = (int64_t)hf->iters - 1;
      ctx->launches += extra * P->loop_kernels;
      for (int f = 0; f < F_COUNT; ++f) {
        ctx->fam_launches[f] += extra * P->loop_fam_launches[f];
        ctx->fam_bytes[f] += double(extra) * P->loop_fam_bytes[f];
      }
      ctx->routes[R_PCG_DEVICE_LOOP]++;
    }
  }
  if (!done_by_graph) {
    cond.on = 0;
    // One PCG step = ONE graph launch (vector kernels and the whole cycle recorded together): three
    // host calls per step (launch, flag copy, synchronise) instead of nine.
    const bool step_graph = P->use_graph && !ctx->timers_on;
    const bool was_capturing = P->capturing;
    if (step_graph) P->capturing = true;  // the single cycle of the prologue is not worth an executable graph
    const int prc = prologue();
    P->capturing = was_capturing;
    AMGB_TRY(prc);
    AMGB_TRY(read_flags());
    int64_t it = 0;
    cudaGraphExec_t step_exec = nullptr;
    struct ExecGuard {
      cudaGraphExec_t* e;
      ~ExecGuard() { if (*e) cudaGraphExecDestroy(*e); }
    } exec_guard{&step_exec};
    int64_t step_kernels = 0, step_fam[F_COUNT] = {0};
    double step_bytes[F_COUNT] = {0};
    if (step_graph && !hf->done && it < max_steps) {
      const int64_t l0 = ctx->launches;
      int64_t fam_l0[F_COUNT];
      double fam_b0[F_COUNT];
      for (int f = 0; f < F_COUNT; ++f) {
        fam_l0[f] = ctx->fam_launches[f];
        fam_b0[f] = ctx->fam_bytes[f];
      }
      cudaGraph_t g = nullptr;
      int brc = AMGB_OK;
      if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        P->capturing = true;
        brc = body();
        P->capturing = was_capturing;
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
        if (brc == AMGB_OK && ce == cudaSuccess && g) {
          if (cudaGraphInstantiate(&step_exec, g, 0) != cudaSuccess) step_exec = nullptr;
        }
        if (g) cudaGraphDestroy(g);
      }
      (void)cudaGetLastError();
      step_kernels = ctx->launches - l0;
      ctx->launches = l0;
      for (int f = 0; f < F_COUNT; ++f) {
        step_fam[f] = ctx->fam_launches[f] - fam_l0[f];
        step_bytes[f] = ctx->fam_bytes[f] - fam_b0[f];
        ctx->fam_launches[f] = fam_l0[f];
        ctx->fam_bytes[f] = fam_b0[f];
      }
      AMGB_TRY(brc);
    }
    while (!hf->done && it < max_steps) {
      if (step_exec) {
        AMGB_CUDA(ctx, cudaGraphLaunch(step_exec, ctx->stream));
        ctx->routes[R_CYCLE_GRAPH]++;
        ctx->launches += step_kernels;
        for (int f = 0; f < F_COUNT; ++f) {
          ctx->fam_launches[f] += step_fam[f];
          ctx->fam_bytes[f] += step_bytes[f];
        }
      } else {
        AMGB_TRY(body());
      }
      AMGB_TRY(read_flags());
      ++it;
    }
    AMGB_TRY(epilogue());
  }
  *n_iters = hf->iters;
  const int status = hf->status;
  const bool done = hf->done != 0;
  const int iters = hf->iters;
  if (res_hist && hist_cap > 0) {
    int64_t k = iters + 1;
    if (k > cap) k = cap;
    AMGB_CUDA(ctx, cudaMemcpyAsync(res_hist, hist.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ds) AMGB_TRY(peer_check(ctx, ds->peer_err));
  if (status != 0) return set_error(ctx, status, "PCG breakdown at iteration %d", iters);
  if (!done)
    return set_error(ctx, AMGB_ERR_NO_CONVERGENCE, "PCG did not reach %g in %lld steps", abs_tol,
                     (long long)max_steps);
  return AMGB_OK;
}

// z_user = M^{-1} r_user with both vectors in the user numbering (device)
static int vmult_user(amgb_precond* P, double* dst, const double* src) {
  amgb_ctx* ctx = P->ctx;
  Level& L0 = P->lv[0];
  const int64_t n = L0.n_solve;
  const unsigned vgrid = (unsigned)div_up(n, kBlock);
  DevBuf<double> r, z;
  AMGB_TRY(r.alloc(ctx, L0.n_vec));
  AMGB_TRY(z.alloc(ctx, L0.n_vec));
  AMGB_LAUNCH(ctx, F_VEC, 20.0 * n, gather_kernel, vgrid, kBlock, 0, n, L0.perm.p, src, r.p);
  AMGB_CHECK_LAUNCH(ctx);
  const bool g = P->use_graph;
  P->use_graph = false;  // one-off buffers: not worth a capture
  const int rc = vcycle_apply(P, z.p, r.p);
  P->use_graph = g;
  AMGB_TRY(rc);
  AMGB_LAUNCH(ctx, F_VEC, 20.0 * n, scatter_kernel, vgrid, kBlock, 0, n, L0.perm.p, z.p, dst);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (P->dist) AMGB_TRY(peer_check(ctx, P->dist->peer_err));
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Row-partitioned solve-phase setup.  Solve numbering of a rank on a level:
//   [owned C points | owned F points | halo points referenced by owned rows, by global id]
// The operators hold the owned rows only; the halo part of a vector is refreshed from the
// owners (HaloPlan vplan) before every kernel that gathers from it.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
dist_perm_kernel(int64_t nloc, const int32_t* __restrict__ cf_own, const int32_t* __restrict__ f2c_own, int nC,
                 int32_t* __restrict__ perm, int32_t* __restrict__ inv_perm) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= nloc) return;
  const int rank = f2c_own[i] - f2c_own[0];  // owned C points before i
  const int ni = cf_own[i] > 0 ? rank : nC + ((int)i - rank);
  inv_perm[i] = ni;
  perm[ni] = (int)i;
}

__global__ void __launch_bounds__(kBlock)
mark_cols_kernel(int64_t nnz, const int32_t* __restrict__ col, int32_t* __restrict__ need) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k < nnz) need[col[k]] = 1;
}

__device__ __forceinline__ int64_t lower_bound_i32(const int32_t* __restrict__ a, int64_t n, int32_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kBlock)
mark_gids_kernel(int64_t n, const int32_t* __restrict__ ids, const int32_t* __restrict__ gid, int64_t next,
                 int32_t* __restrict__ need) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k >= n) return;
  const int64_t e = lower_bound_i32(gid, next, ids[k]);
  if (e < next && gid[e] == ids[k]) need[e] = 1;
}

__global__ void __launch_bounds__(kBlock)
colmap_kernel(int64_t next, int64_t o0, int64_t nloc, const int32_t* __restrict__ need,
              const int32_t* __restrict__ pos, const int32_t* __restrict__ inv_perm,
              const int32_t* __restrict__ gid, int32_t* __restrict__ colmap, int32_t* __restrict__ halo_gid) {
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= next) return;
  int c = -1;
  if (e >= o0 && e < o0 + nloc) {
    c = inv_perm[e - o0];
  } else if (need[e]) {
    c = (int)nloc + pos[e];
    halo_gid[pos[e]] = gid[e];
  }
  colmap[e] = c;
}

// compact coarse column t of Phat -> solve index on the next level
__global__ void __launch_bounds__(kBlock)
tcmap_kernel(int64_t nct, const int32_t* __restrict__ tc_gid, const int32_t* __restrict__ gid_next, int64_t next_n,
             const int32_t* __restrict__ colmap_next, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (t >= nct) return;
  const int64_t e = lower_bound_i32(gid_next, next_n, tc_gid[t]);
  out[t] = (e < next_n && gid_next[e] == tc_gid[t]) ? colmap_next[e] : -1;
}

static void csr_view(amgb_ctx* ctx, const DeviceCsr& M, int64_t row0, int64_t rows, int64_t nnz, int64_t ncols,
                     DeviceCsr& V) {
  V.n = rows;
  V.ncols = ncols;
  V.nnz = nnz;
  V.rp.wrap(ctx, M.rp.p + row0, rows + 1);
  V.col.wrap(ctx, M.col.p, M.col.n);
  V.val.wrap(ctx, M.val.p, M.val.n);
}

// natural (ascending global id) position -> C/F-permuted index of a replicated level
__global__ void __launch_bounds__(kBlock)
tcmap_replicated_kernel(int64_t nct, const int32_t* __restrict__ tc_gid, const int32_t* __restrict__ inv_perm_next,
                        int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (t < nct) out[t] = inv_perm_next[tc_gid[t]];
}

int finish_solve_setup_dist(amgb_precond* P) {
  amgb_ctx* ctx = P->ctx;
  amgb_dist_state* ds = P->dist;
  amgb_comm* comm = ds->comm;
  const int nl = (int)P->lv.size();
  const int nd = std::min(nl, ds->replicated_from);  // partitioned levels are [0, nd)
  if (const char* e = std::getenv("AMGB_OVERLAP_MIN_ROWS")) ds->overlap_min_rows = std::atoll(e);
  auto interior = [&](Sell& S, int64_t halo_begin, int64_t bsplit) -> int {
    if (comm->size < 2 || S.n < ds->overlap_min_rows) return AMGB_OK;
    return find_interior_rows(ctx, S, (int)halo_begin, (int)bsplit);
  };
  // replicated tail first: its numbering is needed by the last partitioned level
  if (nd < nl) AMGB_TRY(finish_solve_setup_range(P, nd));
  // 1. numbering, halo lists, vector plans
  for (int l = 0; l < nd; ++l) {
    Level& L = P->lv[l];
    DistLevel& D = ds->dl[l];
    const int64_t nloc = D.nloc, o0 = D.o0, next = D.next;
    ctx->cur_level = l;
    AMGB_TRY(L.perm.alloc(ctx, nloc));
    AMGB_TRY(L.inv_perm.alloc(ctx, nloc));
    if (l + 1 < nl) {  // (a rank may own nothing here: cf is then an empty array)
      AMGB_LAUNCH(ctx, F_AUX, 16.0 * nloc, dist_perm_kernel, (unsigned)div_up(nloc, kBlock), kBlock, 0, nloc,
                  (const int32_t*)(L.cf.p + o0), (const int32_t*)(L.f2c.p + o0), (int)L.n_coarse, L.perm.p,
                  L.inv_perm.p);
    } else {
      L.n_coarse = 0;
      AMGB_LAUNCH(ctx, F_AUX, 16.0 * nloc, build_perm_kernel, (unsigned)div_up(nloc, kBlock), kBlock, 0, nloc,
                  (const int32_t*)nullptr, (const int32_t*)nullptr, 0, L.perm.p, L.inv_perm.p);
    }
    AMGB_CHECK_LAUNCH(ctx);
    DevBuf<int32_t> need, pos, halo_gid;
    AMGB_TRY(need.alloc_zero(ctx, next));
    AMGB_TRY(pos.alloc(ctx, next + 1));
    // owned rows of Ahat are the entry range [rp[o0], rp[o0 + nloc])
    int32_t eb = 0, ee = 0;
    AMGB_TRY(read_i32(ctx, L.A.rp.p + o0, &eb));
    AMGB_TRY(read_i32(ctx, L.A.rp.p + o0 + nloc, &ee));
    AMGB_LAUNCH(ctx, F_AUX, 8.0 * (ee - eb), mark_cols_kernel, (unsigned)div_up(ee - eb, kBlock), kBlock, 0,
                (int64_t)(ee - eb), (const int32_t*)(L.A.col.p + eb), need.p);
    if (l > 0) {  // coarse points my fine rows interpolate from
      const DeviceCsr& Pp = ds->dl[l - 1].Pown;
      AMGB_LAUNCH(ctx, F_AUX, 8.0 * Pp.nnz, mark_gids_kernel, (unsigned)div_up(Pp.nnz, kBlock), kBlock, 0, Pp.nnz,
                  (const int32_t*)Pp.col.p, (const int32_t*)D.gid.p, next, need.p);
    }
    AMGB_CHECK_LAUNCH(ctx);
    if (nloc) AMGB_CUDA(ctx, cudaMemsetAsync(need.p + o0, 0, nloc * sizeof(int32_t), ctx->stream));
    AMGB_TRY(exclusive_scan_i32(ctx, need.p, pos.p, next));
    int32_t nh = 0;
    AMGB_TRY(read_i32(ctx, pos.p + next, &nh));
    D.nhalo = nh;
    AMGB_TRY(halo_gid.alloc(ctx, nh));
    AMGB_TRY(D.colmap.alloc(ctx, next));
    AMGB_LAUNCH(ctx, F_AUX, 20.0 * next, colmap_kernel, (unsigned)div_up(next, kBlock), kBlock, 0, next, o0, nloc,
                (const int32_t*)need.p, (const int32_t*)pos.p, (const int32_t*)L.inv_perm.p,
                (const int32_t*)D.gid.p, D.colmap.p, halo_gid.p);
    AMGB_CHECK_LAUNCH(ctx);
    L.n_solve = nloc;
    L.n_vec = nloc + nh;
    AMGB_TRY(build_vector_plan(ctx, comm, halo_gid.p, nh, nloc, D.own.starts, L.inv_perm.p, D.vplan));
  }
  // 2. operators (owned rows) in SELL, smoother diagonals, work vectors
  for (int l = 0; l < nd; ++l) {
    Level& L = P->lv[l];
    DistLevel& D = ds->dl[l];
    const int64_t nloc = D.nloc, o0 = D.o0;
    ctx->cur_level = l;
    DeviceCsr Av;
    csr_view(ctx, L.A, o0, nloc, D.own.M.nnz, L.n_vec, Av);
    AMGB_TRY(csr_to_sell(ctx, Av, L.perm.p, D.colmap.p, L.As));
    AMGB_TRY(interior(L.As, nloc, L.n_coarse));
    if (l + 1 < nl) {
      Level& C = P->lv[l + 1];
      const bool to_replicated = l + 1 >= nd;
      DevBuf<int32_t> tcmap, natural, natural_inv;
      AMGB_TRY(tcmap.alloc(ctx, D.nct));
      if (to_replicated) {
        // the next level lives whole on every rank: columns of P address its C/F numbering
        // directly; the restriction writes my coarse points in natural order (gathered in cycle())
        AMGB_LAUNCH(ctx, F_AUX, 12.0 * D.nct, tcmap_replicated_kernel, (unsigned)div_up(D.nct, kBlock), kBlock, 0,
                    D.nct, (const int32_t*)D.tc_gid.p, (const int32_t*)C.inv_perm.p, tcmap.p);
        AMGB_TRY(natural.alloc(ctx, D.nc_own));
        AMGB_TRY(natural_inv.alloc(ctx, D.nc_own));
        AMGB_LAUNCH(ctx, F_AUX, 8.0 * D.nc_own, build_perm_kernel, (unsigned)div_up(D.nc_own, kBlock), kBlock, 0,
                    D.nc_own, (const int32_t*)nullptr, (const int32_t*)nullptr, 0, natural.p, natural_inv.p);
      } else {
        DistLevel& Dn = ds->dl[l + 1];
        AMGB_LAUNCH(ctx, F_AUX, 12.0 * D.nct, tcmap_kernel, (unsigned)div_up(D.nct, kBlock), kBlock, 0, D.nct,
                    (const int32_t*)D.tc_gid.p, (const int32_t*)Dn.gid.p, Dn.next, (const int32_t*)Dn.colmap.p,
                    tcmap.p);
      }
      AMGB_CHECK_LAUNCH(ctx);
      DeviceCsr Pv;
      csr_view(ctx, L.P, o0, nloc, D.Pown.nnz, C.n_vec, Pv);
      AMGB_TRY(csr_to_sell(ctx, Pv, L.perm.p, tcmap.p, L.Ps));
      if (!to_replicated) AMGB_TRY(interior(L.Ps, C.n_solve, L.n_coarse));
      L.R.ncols = L.n_vec;
      AMGB_TRY(csr_to_sell(ctx, L.R, to_replicated ? natural.p : C.perm.p, D.colmap.p, L.Rs));
      AMGB_TRY(interior(L.Rs, nloc, to_replicated ? 0 : C.n_coarse));
      AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // tcmap goes out of scope
      L.R.rp.release();
      L.R.col.release();
      L.R.val.release();
      L.P.rp.release();
      L.P.col.release();
      L.P.val.release();
      if (!P->data.keep_setup_intermediates) {
        D.Pown.rp.release();
        D.Pown.col.release();
        D.Pown.val.release();
        D.Pown.nnz = 0;
      }
      if (to_replicated) {
        AMGB_TRY(ds->repl_own.alloc(ctx, D.nc_own));
        AMGB_TRY(ds->repl_full.alloc(ctx, D.nc_global));
      }
    }
    if (P->relax_down == 16) AMGB_TRY(cheby_setup_level(P, l));  // collective: inner products over all ranks
    // scaling of the Jacobi-type relaxation of this level: the coarsest level is relaxed with the COARSE
    // type when that is a Jacobi-type one (relax_coarse may differ from relax_down: 0 vs 18); with
    // Chebyshev on the way down and up, 1/diag only serves such a coarse relaxation (or the sweeps that
    // stand in for Gaussian elimination when the coarsest grid is too large for it)
    const bool coarsest = l == nl - 1;
    int aux_type = P->relax_down == 16 ? 0 : P->relax_down;
    if (coarsest && (P->relax_coarse == 0 || P->relax_coarse == 18)) aux_type = P->relax_coarse;
    AMGB_TRY(L.inv_relax.alloc(ctx, nloc));
    AMGB_DISPATCH_T(L.As.T, AMGB_LAUNCH(ctx, F_AUX, L.As.csr_bytes() + 8.0 * nloc, sell_aux_kernel<TT>,
                                        (unsigned)div_up(L.As.nslices * 32, kBlock), kBlock, 0,
                                        (int)L.As.nslices, (int)nloc, L.As.slice_ptr.p, L.As.col.p, L.As.val.p,
                                        aux_type, L.inv_relax.p));
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(L.tmp.alloc(ctx, L.n_vec));
    if (l > 0) {
      AMGB_TRY(L.u.alloc(ctx, L.n_vec));
      AMGB_TRY(L.f.alloc(ctx, L.n_vec));
    }
    // the extended matrix is only needed by the setup; the solve runs on the SELL copy
    if (!P->data.keep_setup_intermediates) {
      L.A.rp.release();
      L.A.col.release();
      L.A.val.release();
    }
    L.f2c.release();
  }
  ctx->cur_level = 0;
  // 3. coarsest level still partitioned: replicated dense factorisation
  if (nd == nl) {
    const OwnedCsr& own = ds->dl[nl - 1].own;
    ds->coarse_n = own.n_global;
    ds->coarse_starts = own.starts;
    P->dense_ok = false;
    if (P->relax_coarse == 9 && own.n_global <= kMaxDenseCoarse) {
      DeviceCsr full;
      AMGB_TRY(allgather_rows(ctx, comm, own, full));
      AMGB_TRY(setup_dense_from(P, full));
      AMGB_TRY(ds->coarse_f.alloc(ctx, own.n_global));
      AMGB_TRY(ds->coarse_x.alloc(ctx, own.n_global));
      AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  AMGB_TRY(ds->red.alloc_zero(ctx, 8));
  // 4. peer-memory plans: halo of every partitioned level, the all-gather below the last
  //    partitioned level (replication cut or coarsest right-hand side), the PCG scalars
  {
    const int S = comm->size, me = comm->rank;
    std::vector<PeerSpec> specs(nd + 2);
    for (int l = 0; l < nd; ++l) {
      const HaloPlan& pl = ds->dl[l].vplan;
      specs[l].send_cnt = pl.send_cnt;
      specs[l].recv_cnt = pl.recv_cnt;
      specs[l].dst_off = pl.recv_off;
      specs[l].send_idx = pl.send_idx.p;
    }
    const std::vector<int64_t>& gs = nd < nl ? ds->dl[nd - 1].cstarts : ds->coarse_starts;
    PeerSpec& g = specs[nd];
    PeerSpec& r = specs[nd + 1];
    g.send_cnt.assign(S, 0);
    g.recv_cnt.assign(S, 0);
    g.dst_off.assign(S, 0);
    r.send_cnt.assign(S, 8);
    r.recv_cnt.assign(S, 8);
    r.dst_off.assign(S, 0);
    r.send_cnt[me] = r.recv_cnt[me] = 0;
    if ((int)gs.size() == S + 1) {
      for (int q = 0; q < S; ++q) {
        if (q == me) continue;
        g.send_cnt[q] = gs[me + 1] - gs[me];
        g.recv_cnt[q] = gs[q + 1] - gs[q];
        g.dst_off[q] = gs[q];
      }
    }
    std::vector<PeerPlan> plans;
    AMGB_TRY(build_peer_plans(ctx, comm, specs, plans, &ds->window_slot, &ds->peer_err));
    if (ds->window_slot >= 0) {
      ds->vpeer.assign(plans.begin(), plans.begin() + nd);
      ds->gather_peer = plans[nd];
      ds->red_peer = plans[nd + 1];
      // the whole cycle is kernels now: capture it (measured on 2 x B200, m=200: 84 -> 79 ms per
      // solve; with NCCL send/recv nodes the graph was slower than plain launches)
      if (comm->capturable() && !std::getenv("AMGB_NO_GRAPH")) P->use_graph = true;
    }
  }
  return AMGB_OK;
}

}  // namespace amgb

using namespace amgb;

extern "C" {

int amgb_precond_vmult_device(amgb_precond* P, double* dst_device, const double* src_device) {
  if (!P || !dst_device || !src_device) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(P->ctx->device);
  return vmult_user(P, dst_device, src_device);
}

int amgb_precond_vmult(amgb_precond* P, double* dst, const double* src) {
  if (!P || !dst || !src) return AMGB_ERR_BAD_ARG;
  amgb_ctx* ctx = P->ctx;
  cudaSetDevice(ctx->device);
  const int64_t n = P->lv[0].n_solve;  // row-partitioned path: the owned slab
  DevBuf<double> d, s;
  AMGB_TRY(d.alloc(ctx, n));
  AMGB_TRY(s.alloc(ctx, n));
  AMGB_CUDA(ctx, cudaMemcpyAsync(s.p, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_TRY(vmult_user(P, d.p, s.p));
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

// Row-partitioned PCG: x, b are this rank's owned slabs.  Collective over the communicator.
int amgb_dist_cg_solve_device(amgb_ctx* ctx, double* x_local_device, const double* b_local_device, amgb_precond* P,
                              int64_t max_steps, double abs_tol, double* res_hist, int64_t hist_cap,
                              int64_t* n_iters) {
  if (!ctx || !P || !P->dist || !n_iters || max_steps < 0) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  return cg_device(ctx, nullptr, x_local_device, b_local_device, P, max_steps, abs_tol, res_hist, hist_cap, n_iters);
}

int amgb_dist_cg_solve(amgb_ctx* ctx, double* x_local, const double* b_local, amgb_precond* P, int64_t max_steps,
                       double abs_tol, double* res_hist, int64_t hist_cap, int64_t* n_iters) {
  if (!ctx || !P || !P->dist || !n_iters || max_steps < 0) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  const int64_t n = P->lv[0].n_solve;
  DevBuf<double> dx, db;
  AMGB_TRY(dx.alloc(ctx, n));
  AMGB_TRY(db.alloc(ctx, n));
  if (n) {
    AMGB_CUDA(ctx, cudaMemcpyAsync(dx.p, x_local, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AMGB_CUDA(ctx, cudaMemcpyAsync(db.p, b_local, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  const int rc = cg_device(ctx, nullptr, dx.p, db.p, P, max_steps, abs_tol, res_hist, hist_cap, n_iters);
  if ((rc == AMGB_OK || rc == AMGB_ERR_NO_CONVERGENCE) && n) {
    AMGB_CUDA(ctx, cudaMemcpyAsync(x_local, dx.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return rc;
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
