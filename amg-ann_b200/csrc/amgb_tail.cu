// Coarse tail of the V-cycle in ONE kernel.
//
// Below a few hundred rows a level is pure launch latency: pre-smoothing (2 launches),
// residual, restriction, memset, prolongation, post-smoothing (2) -- eight dependent kernels
// of ~2.5 us each for a few kilobytes of operator, times four or five such levels per cycle.
// Here every level from `tail_from` down to the coarsest is packed once per setup into one
// contiguous blob (CSR in the level's C/F numbering, 16-bit columns, 1/l1 or 1/diag), and a
// single thread block copies the blob to shared memory at the start of each cycle and runs
// the whole recursion -- C/F Jacobi half sweeps, residual, restriction, the Gaussian-
// elimination solve on the coarsest grid, prolongation, post-smoothing -- with block barriers
// between the phases.  The tail levels get no SELL operators and no work vectors of their own.
//
// Replaces, for those levels, the per-level launches of cycle() in amgb_solve.cu
// (hypre_BoomerAMGCycle; ref common/amg_solver.h:54 through PCApply).  Supported: the
// reference's cycle shape (V-cycle, one sweep down and up, C/F-ordered Jacobi-type relaxation
// types 0 / 18, Gaussian elimination on a coarsest grid of at most 64 rows, one cycle per
// application); any other option combination keeps the per-level path.
#include <algorithm>
#include <cstdlib>

#include "amgb_internal.cuh"

namespace amgb {

constexpr int kTailThreads = 1024;
constexpr int kTailLanes = 4;                       // lanes per row
constexpr int kTailGroups = kTailThreads / kTailLanes;
constexpr size_t kTailSmemBudget = 200 * 1024;      // of the 227 KB a block may have
constexpr int kTailMaxRows = 1024;                  // rows of the largest tail level (one block scan in the pack kernel)
constexpr int kTailMaxDense = 64;

static size_t align16(size_t b) { return (b + 15) / 16 * 16; }

// ---------------------------------------------------------------------------
// Pack: one block per operator.  Rows in the new (C points first) numbering, entries in the
// order of the setup CSR (ascending old column).
// ---------------------------------------------------------------------------
struct TailPackOp {
  int rows;                 // rows of the packed operator
  const int32_t* row_src;   // new row -> old row (perm of the row space), or nullptr: identity
  const int32_t* col_map;   // old col -> new col (inv_perm of the column space), or nullptr
  const int32_t* rp;        // setup CSR
  const int32_t* col;
  const double* val;
  int off_rp, off_col, off_val;  // byte offsets into the blob
  int off_inv;              // >= 0: also write 1/l1 (relax_type 18) or 1/diag of every row there
  int relax_type;
};

__global__ void __launch_bounds__(kTailThreads)
tail_pack_kernel(const TailPackOp* __restrict__ ops, unsigned char* __restrict__ blob) {
  const TailPackOp op = ops[blockIdx.x];
  __shared__ int s_warp[kTailThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int32_t* out_rp = reinterpret_cast<int32_t*>(blob + op.off_rp);
  uint16_t* out_col = reinterpret_cast<uint16_t*>(blob + op.off_col);
  double* out_val = reinterpret_cast<double*>(blob + op.off_val);
  int len = 0, src = 0, b = 0;
  if (tid < op.rows) {
    src = op.row_src ? op.row_src[tid] : tid;
    b = op.rp[src];
    len = op.rp[src + 1] - b;
  }
  // block-wide exclusive scan of the row lengths (rows <= kTailMaxRows = blockDim)
  int incl = len;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = s_warp[lane];
    int vi = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, vi, d);
      if (lane >= d) vi += t;
    }
    s_warp[lane] = vi - v;
  }
  __syncthreads();
  const int start = incl - len + s_warp[wid];
  if (tid < op.rows) {
    out_rp[tid] = start;
    if (tid == op.rows - 1) out_rp[op.rows] = start + len;
    double diag = 0.0, l1 = 0.0;
    for (int j = 0; j < len; ++j) {
      int c = op.col[b + j];
      if (op.col_map) c = op.col_map[c];
      const double v = op.val[b + j];
      out_col[start + j] = (uint16_t)c;
      out_val[start + j] = v;
      if (c == tid && v != 0.0) diag = v;
      l1 += fabs(v);
    }
    if (op.off_inv >= 0) {
      double inv = 0.0;  // rows with a zero diagonal are skipped (hypre_BoomerAMGRelax)
      if (diag != 0.0) inv = op.relax_type == 18 ? 1.0 / l1 : 1.0 / diag;
      reinterpret_cast<double*>(blob + op.off_inv)[tid] = inv;
    }
  }
}

// ---------------------------------------------------------------------------
// The cycle.
// ---------------------------------------------------------------------------
struct SmOp {
  const int32_t* rp;
  const uint16_t* col;
  const double* val;
};

__device__ __forceinline__ SmOp sm_op(const unsigned char* sm, const TailOpDesc& o) {
  SmOp r;
  r.rp = reinterpret_cast<const int32_t*>(sm + o.off_rp);
  r.col = reinterpret_cast<const uint16_t*>(sm + o.off_col);
  r.val = reinterpret_cast<const double*>(sm + o.off_val);
  return r;
}

// rows [lo, hi) of op: s = sum_j val_j * X(col_j), four lanes per row; EPI(row, s) by lane 0 of the group.
// The trip count is uniform over the block, so the shuffles are convergent.
template <class X, class Epi>
__device__ __forceinline__ void tail_rows(const SmOp& op, int lo, int hi, X x, Epi epi) {
  const int g = threadIdx.x / kTailLanes, q = threadIdx.x % kTailLanes;
  for (int base = lo; base < hi; base += kTailGroups) {
    const int row = base + g;
    double s = 0.0;
    if (row < hi) {
      const int b = op.rp[row], e = op.rp[row + 1];
      for (int k = b + q; k < e; k += kTailLanes) s += op.val[k] * x((int)op.col[k]);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (row < hi && q == 0) epi(row, s);
  }
}

__global__ void __launch_bounds__(kTailThreads, 1)
cycle_tail_kernel(TailDesc d, const unsigned char* __restrict__ blob, const double* __restrict__ dense,
                  const double* __restrict__ f_in, double* __restrict__ u_out) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x;
  {
    const int4* src = reinterpret_cast<const int4*>(blob);
    int4* dst = reinterpret_cast<int4*>(sm);
    const int n16 = d.blob_bytes / 16;
    for (int i = tid; i < n16; i += kTailThreads) dst[i] = src[i];
    double* M = reinterpret_cast<double*>(sm + d.off_dense);
    for (int i = tid; i < d.dense_n * d.dense_n; i += kTailThreads) M[i] = dense[i];
    double* f0 = reinterpret_cast<double*>(sm + d.lv[0].off_f);
    for (int i = tid; i < d.lv[0].n; i += kTailThreads) f0[i] = f_in[i];
  }
  __syncthreads();
  const double w = d.w;
  const int last = d.nlev - 1;
  // ---- way down
  for (int l = 0; l < last; ++l) {
    const TailLevelDesc& L = d.lv[l];
    const int n = L.n, nC = L.nC;
    const SmOp A = sm_op(sm, L.A), R = sm_op(sm, L.R);
    const double* inv = reinterpret_cast<const double*>(sm + L.off_inv);
    const double* f = reinterpret_cast<const double*>(sm + L.off_f);
    double* u = reinterpret_cast<double*>(sm + L.off_u);
    double* t = reinterpret_cast<double*>(sm + L.off_t);
    if (l == 0) {  // pre-smoothing of the C points from the zero guess (deeper levels: done by the restriction above)
      for (int i = tid; i < nC; i += kTailThreads) t[i] = 0.0 + w * (f[i] - 0.0) * inv[i];
      __syncthreads();
    }
    // F points: the C values are fresh, the F values still zero
    tail_rows(A, nC, n, [&](int c) { return c < nC ? t[c] : 0.0; },
              [&](int row, double s) { t[row] = 0.0 + w * (f[row] - s) * inv[row]; });
    __syncthreads();
    // residual (parked in u)
    tail_rows(A, 0, n, [&](int c) { return t[c]; }, [&](int row, double s) { u[row] = f[row] - s; });
    __syncthreads();
    // restriction; the C points of the next level are pre-smoothed on the spot
    const TailLevelDesc& N = d.lv[l + 1];
    double* fn = reinterpret_cast<double*>(sm + N.off_f);
    if (l + 1 < last) {
      double* tn = reinterpret_cast<double*>(sm + N.off_t);
      const double* invn = reinterpret_cast<const double*>(sm + N.off_inv);
      const int nCn = N.nC;
      tail_rows(R, 0, N.n, [&](int c) { return u[c]; }, [&](int row, double s) {
        fn[row] = s;
        if (row < nCn) tn[row] = 0.0 + w * (s - 0.0) * invn[row];
      });
    } else {
      tail_rows(R, 0, N.n, [&](int c) { return u[c]; }, [&](int row, double s) { fn[row] = s; });
    }
    __syncthreads();
  }
  // ---- coarsest grid: forward / back substitution with the factors of hypre_gselim (one warp)
  {
    const int n = d.dense_n;
    const double* M = reinterpret_cast<const double*>(sm + d.off_dense);
    double* x = reinterpret_cast<double*>(sm + d.lv[last].off_f);  // solved in place
    if (tid < 32) {
      for (int k = 0; k + 1 < n; ++k) {
        if (M[k * n + k] != 0.0) {
          const double xk = x[k];
          for (int j = k + 1 + tid; j < n; j += 32) {
            const double lj = M[j * n + k];
            if (lj != 0.0) x[j] = __dsub_rn(x[j], __dmul_rn(lj, xk));
          }
        }
        __syncwarp();
      }
      for (int k = n - 1; k >= 0; --k) {
        const double dk = M[k * n + k];
        if (dk != 0.0) {
          if (tid == 0) x[k] = x[k] / dk;
          __syncwarp();
          const double xk = x[k];
          for (int j = tid; j < k; j += 32) {
            const double uj = M[j * n + k];
            if (uj != 0.0) x[j] = __dsub_rn(x[j], __dmul_rn(xk, uj));
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  // ---- way up
  for (int l = last - 1; l >= 0; --l) {
    const TailLevelDesc& L = d.lv[l];
    const int n = L.n, nC = L.nC;
    const SmOp A = sm_op(sm, L.A), Pp = sm_op(sm, L.P);
    const double* inv = reinterpret_cast<const double*>(sm + L.off_inv);
    const double* f = reinterpret_cast<const double*>(sm + L.off_f);
    double* u = reinterpret_cast<double*>(sm + L.off_u);
    double* t = reinterpret_cast<double*>(sm + L.off_t);
    const TailLevelDesc& N = d.lv[l + 1];
    const double* e = reinterpret_cast<const double*>(sm + (l + 1 == last ? N.off_f : N.off_u));
    tail_rows(Pp, 0, n, [&](int c) { return e[c]; }, [&](int row, double s) { t[row] += s; });
    __syncthreads();
    // post-smoothing: F points, then C points with the fresh F values
    tail_rows(A, nC, n, [&](int c) { return t[c]; },
              [&](int row, double s) { u[row] = t[row] + w * (f[row] - s) * inv[row]; });
    __syncthreads();
    tail_rows(A, 0, nC, [&](int c) { return c < nC ? t[c] : u[c]; },
              [&](int row, double s) { u[row] = t[row] + w * (f[row] - s) * inv[row]; });
    __syncthreads();
  }
  const double* res = reinterpret_cast<const double*>(sm + (last == 0 ? d.lv[0].off_f : d.lv[0].off_u));
  for (int i = tid; i < d.lv[0].n; i += kTailThreads) u_out[i] = res[i];
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
static size_t op_bytes(int rows, int64_t nnz) {
  return align16(4 * (size_t)(rows + 1)) + align16(2 * (size_t)nnz) + align16(8 * (size_t)nnz);
}

// First level of the tail (levels [from, nl) run in the fused kernel), or -1.
int tail_plan(const amgb_precond* P, int l0) {
  const amgb_boomeramg_data& d = P->data;
  const int nl = (int)P->lv.size();
  if (std::getenv("AMGB_NO_TAIL")) return -1;
  const bool jacobi = (P->relax_down == 0 || P->relax_down == 18) && P->relax_up == P->relax_down;
  if (!jacobi || P->relax_coarse != 9 || d.relax_order != 1 || d.n_sweeps != 1 || d.w_cycle ||
      (d.max_iter != 0 && d.max_iter != 1))
    return -1;
  if (nl - l0 < 1 || P->lv[nl - 1].A.n > kTailMaxDense) return -1;
  const int64_t nd = P->lv[nl - 1].A.n;
  size_t bytes = align16(8 * nd * nd) + align16(8 * nd);  // factors, right-hand side / solution
  int from = nl - 1;
  // (level 0 stays out: its SELL operator also serves the PCG residual and product)
  for (int l = nl - 2; l >= std::max(l0, 1); --l) {
    const Level& L = P->lv[l];
    if (L.A.n > kTailMaxRows || L.A.n > 65535) break;
    const size_t add = op_bytes((int)L.A.n, L.A.nnz) + op_bytes((int)L.A.n, L.P.nnz) +
                       op_bytes((int)P->lv[l + 1].A.n, L.P.nnz) + 4 * align16(8 * (size_t)L.A.n);
    if (bytes + add > kTailSmemBudget || nl - l > kTailMaxLevels) break;
    bytes += add;
    from = l;
  }
  return nl - from >= 3 ? from : -1;  // fewer than two relaxed levels: not worth a special path
}

// Packs levels [P->tail_from, nl) (needs the setup CSR of A, P, R and the permutations).
int tail_pack(amgb_precond* P) {
  amgb_ctx* ctx = P->ctx;
  const int nl = (int)P->lv.size(), from = P->tail_from;
  TailDesc& D = P->tail_desc;
  D = TailDesc();
  D.nlev = nl - from;
  D.w = P->data.relax_weight;
  const int aux_type = P->relax_down;
  std::vector<TailPackOp> ops;
  size_t off = 0;
  auto place = [&](int rows, int64_t nnz, TailOpDesc& o) {
    o.rows = rows;
    o.nnz = (int)nnz;
    o.off_rp = (int)off;
    off += align16(4 * (size_t)(rows + 1));
    o.off_col = (int)off;
    off += align16(2 * (size_t)nnz);
    o.off_val = (int)off;
    off += align16(8 * (size_t)nnz);
  };
  for (int l = from; l < nl - 1; ++l) {
    Level& L = P->lv[l];
    Level& C = P->lv[l + 1];
    TailLevelDesc& T = D.lv[l - from];
    T.n = (int)L.A.n;
    T.nC = (int)L.n_coarse;
    place(T.n, L.A.nnz, T.A);
    place(T.n, L.P.nnz, T.P);
    place((int)C.A.n, L.R.nnz, T.R);
    T.off_inv = (int)off;
    off += align16(8 * (size_t)T.n);
    TailPackOp a{T.n, L.perm.p, L.inv_perm.p, L.A.rp.p, L.A.col.p, L.A.val.p, T.A.off_rp, T.A.off_col, T.A.off_val,
                 T.off_inv, aux_type};
    TailPackOp p{T.n, L.perm.p, C.inv_perm.p, L.P.rp.p, L.P.col.p, L.P.val.p, T.P.off_rp, T.P.off_col, T.P.off_val, -1, 0};
    TailPackOp r{(int)C.A.n, C.perm.p, L.inv_perm.p, L.R.rp.p, L.R.col.p, L.R.val.p, T.R.off_rp, T.R.off_col,
                 T.R.off_val, -1, 0};
    ops.push_back(a);
    ops.push_back(p);
    ops.push_back(r);
  }
  D.blob_bytes = (int)off;
  // shared memory behind the blob: vectors of every level, factors of the coarsest grid
  for (int l = from; l < nl - 1; ++l) {
    TailLevelDesc& T = D.lv[l - from];
    T.off_u = (int)off;
    off += align16(8 * (size_t)T.n);
    T.off_f = (int)off;
    off += align16(8 * (size_t)T.n);
    T.off_t = (int)off;
    off += align16(8 * (size_t)T.n);
  }
  TailLevelDesc& Last = D.lv[D.nlev - 1];
  Last.n = (int)P->lv[nl - 1].A.n;
  Last.nC = 0;
  Last.off_f = (int)off;
  off += align16(8 * (size_t)Last.n);
  Last.off_u = Last.off_f;
  D.dense_n = Last.n;
  D.off_dense = (int)off;
  off += align16(8 * (size_t)Last.n * Last.n);
  P->tail_smem = off;
  if (off > 227 * 1024) return set_error(ctx, AMGB_ERR_RANGE, "coarse tail does not fit shared memory (%zu bytes)", off);
  AMGB_TRY(P->tail_blob.alloc(ctx, std::max<size_t>(D.blob_bytes, 16)));
  if (!ops.empty()) {
    DevBuf<TailPackOp> dops;
    AMGB_TRY(dops.alloc(ctx, ops.size()));
    AMGB_CUDA(ctx, cudaMemcpyAsync(dops.p, ops.data(), ops.size() * sizeof(TailPackOp), cudaMemcpyHostToDevice,
                                   ctx->stream));
    AMGB_LAUNCH(ctx, F_AUX, 2.0 * D.blob_bytes, tail_pack_kernel, (unsigned)ops.size(), kTailThreads, 0,
                (const TailPackOp*)dops.p, P->tail_blob.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `ops` (pageable) and `dops` go out of scope
  }
  AMGB_CUDA(ctx, cudaFuncSetAttribute(cycle_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return AMGB_OK;
}

// u = (cycle from the zero guess on level tail_from with right-hand side f); f, u: device vectors of that level
int tail_cycle(amgb_precond* P, const double* f, double* u) {
  amgb_ctx* ctx = P->ctx;
  ctx->routes[R_CYCLE_TAIL_FUSED]++;
  AMGB_LAUNCH(ctx, F_TAIL, (double)P->tail_desc.blob_bytes, cycle_tail_kernel, 1, kTailThreads, P->tail_smem,
              P->tail_desc, (const unsigned char*)P->tail_blob.p, (const double*)P->dense.p, f, u);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

}  // namespace amgb
