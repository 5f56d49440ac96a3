// Matrix -> V x V four-channel image pooling, replacing ViewMaker::make_view
// (ref common/view_maker.h:26-74: a single-threaded MatGetRow loop).
//
// One streaming, atomic-free pass over the CSR entries:
//   * the V row-bins are contiguous row ranges (ref view_maker.h:41-45,52), hence contiguous
//     ENTRY ranges [rp[r0], rp[r1]); a CTA is given a slice of the entry range of ONE row-bin,
//     so it only needs V-wide accumulators and never looks at row pointers again;
//   * a warp covers 256 consecutive entries per iteration (3 KB) with four pairs of load
//     instructions, each one contiguous run (8 bytes of columns and 16 bytes of values per
//     lane), 8 entries per lane;
//   * the rows of a row-bin of a banded matrix meet the same two or three column bins over and
//     over: the warp caches four bins (ids warp-uniform, partial sum / count / maxima
//     lane-private in registers), so the steady state is loads, a bin computation and
//     predicated register adds -- no cross-lane traffic at all;
//   * a bin that is not cached is installed in a free slot, or makes the warp flush its slots
//     first: a butterfly over the warp and ONE writer (lane 0) per bin into the WARP-PRIVATE
//     shared-memory accumulators -- no two writers ever share an address, no atomics anywhere;
//   * warps, then slices, are combined in index order.
// count is integer-exact, max_pp / max_np are order-independent hence exact, sum has a
// fixed (input-independent) association.
#include <climits>

#include <vector>

#include "amgb_dist.cuh"
#include "amgb_internal.cuh"

namespace amgb {

constexpr int kPoolBlock = 256;
constexpr int kPoolWarps = kPoolBlock / 32;
constexpr int kMaxView = 512;
constexpr int kPoolPerLane = 8;
constexpr int kPoolWarpChunk = 32 * kPoolPerLane;            // entries per warp and iteration
constexpr int kPoolBlockChunk = kPoolWarps * kPoolWarpChunk;  // entries per CTA and iteration
constexpr size_t kPoolSmemMax = (size_t)kPoolWarps * kMaxView * (3 * sizeof(double) + sizeof(int));

struct BinMap {
  int q, q1, p, t, V;
  double rq, rq1;  // 1/q, 1/q1
  // floor(i / d) for 0 <= i < 2^31 without an integer division: the quotient is a bin index
  // (< 2^10), so the double product is within one of it and a single correction makes it exact
  __device__ __forceinline__ static int div(int i, int d, double rd) {
    int b = __double2int_rz(__int2double_rn(i) * rd);
    const long long r = (long long)i - (long long)b * d;
    if (r < 0) --b;
    else if (r >= d) ++b;
    return b;
  }
  __device__ __forceinline__ int bin(int i) const { return i < t ? div(i, q1, rq1) : div(i - t, q, rq) + p; }
  __host__ __device__ int row_begin(int br) const { return br < p ? br * q1 : t + (br - p) * q; }
};

// One (bin, partial) slot of a lane.  The bin ids of the slots are WARP-UNIFORM (every lane caches
// the same kPoolSlots bins), the partials are lane-private registers.
struct PoolSlot {
  double s, pp, np;
  int cnt;
  __device__ __forceinline__ void clear() { s = 0.0; pp = 0.0; np = 0.0; cnt = 0; }
  // (plain compare + select: fmax() carries NaN quieting, ten instructions a piece on sm_100 and two
  // of them per slot and entry made the kernel issue-bound; a NaN entry is ignored either way)
  __device__ __forceinline__ void add(double v) {
    s += v;
    ++cnt;
    pp = v > pp ? v : pp;    // (v < 0 leaves pp, which starts at 0, alone: max(pp, max(v, 0)) == max(pp, v))
    const double nv = -v;
    np = nv > np ? nv : np;
  }
};

// butterfly over the warp, then lane 0 adds the total to the warp's accumulators of `bin`
__device__ __forceinline__ void pool_flush_slot(PoolSlot& a, int bin, int lane, double* w_sum, int* w_cnt, double* w_pp,
                                                double* w_np) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a.s += __shfl_xor_sync(full, a.s, d);
    a.cnt += __shfl_xor_sync(full, a.cnt, d);
    a.pp = fmax(a.pp, __shfl_xor_sync(full, a.pp, d));
    a.np = fmax(a.np, __shfl_xor_sync(full, a.np, d));
  }
  if (lane == 0) {
    w_sum[bin] += a.s;
    w_cnt[bin] += a.cnt;
    w_pp[bin] = fmax(w_pp[bin], a.pp);
    w_np[bin] = fmax(w_np[bin], a.np);
  }
  a.clear();
}

constexpr int kPoolSlots = 4;

template <bool VEC>
__global__ void __launch_bounds__(kPoolBlock, 3)
pool_entries_kernel(BinMap bm, int tiles, long long g0, int nloc, const int32_t* __restrict__ rp,
                    const int32_t* __restrict__ col,
                    const double* __restrict__ val, double* __restrict__ part_sum,
                    long long* __restrict__ part_cnt, double* __restrict__ part_pp,
                    double* __restrict__ part_np) {
  extern __shared__ unsigned char pool_smem[];
  const int V = bm.V;
  double* s_sum = reinterpret_cast<double*>(pool_smem);          // [warps][V]
  double* s_pp = s_sum + kPoolWarps * V;
  double* s_np = s_pp + kPoolWarps * V;
  int* s_cnt = reinterpret_cast<int*>(s_np + kPoolWarps * V);
  const int br = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  for (int t = threadIdx.x; t < kPoolWarps * V; t += kPoolBlock) {
    s_sum[t] = 0.0;
    s_pp[t] = 0.0;
    s_np[t] = 0.0;
    s_cnt[t] = 0;
  }
  __syncthreads();
  // rows [g0, g0 + nloc) of the global matrix are held here (a single device holds all of them)
  const long long rb0 = (long long)bm.row_begin(br) - g0, rb1 = (long long)bm.row_begin(br + 1) - g0;
  const int e_lo = rp[rb0 < 0 ? 0 : (rb0 > nloc ? nloc : (int)rb0)], e_hi = rp[rb1 < 0 ? 0 : (rb1 > nloc ? nloc : (int)rb1)];
  // whole CTA-chunks (absolute multiples of kPoolBlockChunk, so every lane's entry pairs are
  // 8-byte / 16-byte aligned) that overlap the entry range of this row-bin, split over the tiles
  const int c_first = e_lo / kPoolBlockChunk, c_last = (int)(((long long)e_hi + kPoolBlockChunk - 1) / kPoolBlockChunk);
  const int per = (c_last - c_first + tiles - 1) / tiles;
  const int c0 = c_first + tile * per, c1 = min(c0 + per, c_last);
  double* w_sum = s_sum + warp * V;
  double* w_pp = s_pp + warp * V;
  double* w_np = s_np + warp * V;
  int* w_cnt = s_cnt + warp * V;
  // The rows of one row-bin of a banded matrix meet two or three column bins, over and over:
  // the warp caches kPoolSlots bins (ids warp-uniform, partial sums lane-private in registers)
  // and streams the entries into them without any cross-lane traffic.  A bin that is not cached
  // is installed (free slot) or makes the warp flush all slots into its shared-memory
  // accumulators first (butterfly + one writer per bin).
  // (a slot is kept as the COLUMN RANGE of its bin: an entry is matched with two integer compares, the
  // division of BinMap::bin is only paid when a bin is installed)
  int key[kPoolSlots], lo[kPoolSlots], hi[kPoolSlots];
  PoolSlot acc[kPoolSlots];
#pragma unroll
  for (int t = 0; t < kPoolSlots; ++t) {
    key[t] = -2;  // free
    lo[t] = hi[t] = 0;
    acc[t].clear();
  }
  for (int chunk = c0; chunk < c1; ++chunk) {
    // entries 64 j + 2 lane + {0, 1} of the warp's 256, j = 0..3: every load instruction of the warp is
    // one contiguous run (8 bytes per lane for the columns, 16 for the values)
    const long long wbase = (long long)chunk * kPoolBlockChunk + warp * kPoolWarpChunk + 2 * lane;
    int c[kPoolPerLane];
    double v[kPoolPerLane];
    // (warp-uniform: the warp's 256 entries lie inside the row-bin's entry range -- everywhere but at
    // its two ends -- so the four load pairs need no per-lane range checks)
    const long long w0 = wbase - 2 * lane;
    if (VEC && w0 >= e_lo && w0 + kPoolWarpChunk <= e_hi) {
#pragma unroll
      for (int j = 0; j < kPoolPerLane; j += 2) {
        const int2 cc = __ldcs(reinterpret_cast<const int2*>(col + wbase + 32 * j));
        const double2 vv = __ldcs(reinterpret_cast<const double2*>(val + wbase + 32 * j));
        c[j] = cc.x;
        c[j + 1] = cc.y;
        v[j] = vv.x;
        v[j + 1] = vv.y;
      }
    } else
#pragma unroll
    for (int j = 0; j < kPoolPerLane; j += 2) {
      const long long k = wbase + 32 * j;
      if (VEC && k >= e_lo && k + 2 <= e_hi) {
        const int2 cc = __ldcs(reinterpret_cast<const int2*>(col + k));
        const double2 vv = __ldcs(reinterpret_cast<const double2*>(val + k));
        c[j] = cc.x;
        c[j + 1] = cc.y;
        v[j] = vv.x;
        v[j + 1] = vv.y;
      } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool ok = k + u >= e_lo && k + u < e_hi;
          c[j + u] = ok ? __ldcs(col + k + u) : -1;  // (outside the row-bin's entry range: only at its two ends)
          v[j + u] = ok ? __ldcs(val + k + u) : 0.0;
        }
      }
    }
    // Every pass adds the entries whose column lies in a cached bin and consumes them (c = -1); what is
    // left names a bin to install -- in a free slot, or after flushing all slots when none is free.
    // One new bin per pass, so a chunk with k uncached bins takes k + 1 passes; the steady state of a
    // banded matrix is a single pass and a single vote.
    for (;;) {
      int missing = -1;  // a column of mine that no cached bin covers
#pragma unroll
      for (int j = 0; j < kPoolPerLane; ++j) {
        if (c[j] < 0) continue;
        bool hit = false;
#pragma unroll
        for (int t = 0; t < kPoolSlots; ++t)
          if (c[j] >= lo[t] && c[j] < hi[t]) {
            acc[t].add(v[j]);
            hit = true;
          }
        if (hit) c[j] = -1; else missing = c[j];
      }
      const unsigned mm = __ballot_sync(full, missing >= 0);
      if (mm == 0) break;
      const int nb = bm.bin(__shfl_sync(full, missing, __ffs(mm) - 1));
      bool placed = false;
#pragma unroll
      for (int t = 0; t < kPoolSlots; ++t)
        if (!placed && key[t] < 0) {
          key[t] = nb;
          placed = true;
        }
      if (!placed) {  // all slots busy: flush them and start over with the new bin
#pragma unroll
        for (int t = 0; t < kPoolSlots; ++t) {
          pool_flush_slot(acc[t], key[t], lane, w_sum, w_cnt, w_pp, w_np);
          key[t] = -2;
        }
        key[0] = nb;
      }
#pragma unroll
      for (int t = 0; t < kPoolSlots; ++t) {
        lo[t] = key[t] >= 0 ? bm.row_begin(key[t]) : 0;
        hi[t] = key[t] >= 0 ? bm.row_begin(key[t] + 1) : 0;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kPoolSlots; ++t)
    if (key[t] >= 0) pool_flush_slot(acc[t], key[t], lane, w_sum, w_cnt, w_pp, w_np);
  __syncthreads();
  const size_t out = ((size_t)br * tiles + tile) * V;
  for (int bc = threadIdx.x; bc < V; bc += kPoolBlock) {
    double s = 0.0, pp = 0.0, np = 0.0;
    long long cn = 0;
#pragma unroll
    for (int w = 0; w < kPoolWarps; ++w) {
      s += s_sum[w * V + bc];
      cn += s_cnt[w * V + bc];
      pp = fmax(pp, s_pp[w * V + bc]);
      np = fmax(np, s_np[w * V + bc]);
    }
    part_sum[out + bc] = s;
    part_cnt[out + bc] = cn;
    part_pp[out + bc] = pp;
    part_np[out + bc] = np;
  }
}

__global__ void __launch_bounds__(kPoolBlock)
pool_finalize_kernel(int V, int tiles, const double* __restrict__ part_sum,
                     const long long* __restrict__ part_cnt, const double* __restrict__ part_pp,
                     const double* __restrict__ part_np, double* __restrict__ sum,
                     long long* __restrict__ cnt, double* __restrict__ pp, double* __restrict__ np) {
  const int idx = blockIdx.x * kPoolBlock + threadIdx.x;
  if (idx >= V * V) return;
  const int br = idx / V, bc = idx % V;
  double s = 0.0, a = 0.0, b = 0.0;
  long long c = 0;
  for (int t = 0; t < tiles; ++t) {
    const size_t o = ((size_t)br * tiles + t) * V + bc;
    s += part_sum[o];
    c += part_cnt[o];
    a = fmax(a, part_pp[o]);
    b = fmax(b, part_np[o]);
  }
  sum[idx] = s;
  cnt[idx] = c;
  pp[idx] = a;
  np[idx] = b;
}

}  // namespace amgb

using namespace amgb;

// Pooling into device buffers (vv = V*V entries each); ms: device time of the pass.
// M: the rows [g0, g0 + M.n) of the n_global x n_global matrix, GLOBAL column ids
static int pool_to_device(amgb_ctx* ctx, const amgb::DeviceCsr& M, int64_t n_global, int64_t g0, int V, double* d_sum,
                          long long* d_cnt, double* d_pp, double* d_np, float* ms_out) {
  const int n = (int)n_global;
  BinMap bm;
  bm.V = V;
  bm.q = n / V;
  bm.q1 = bm.q + 1;
  bm.p = n % V;
  bm.t = bm.q1 * bm.p;
  bm.rq = bm.q > 0 ? 1.0 / bm.q : 0.0;  // (V > n: every row is a bin of its own, q is never divided by)
  bm.rq1 = 1.0 / bm.q1;
  // slices per row-bin: enough CTAs for every SM, at least a few CTA-chunks of entries each
  int tiles = (int)div_up((int64_t)ctx->sm_count * 12, V);
  const int64_t chunks_per_bin = div_up(div_up(M.nnz, V), kPoolBlockChunk);
  if (tiles > chunks_per_bin / 4) tiles = (int)(chunks_per_bin / 4);
  if (tiles < 1) tiles = 1;
  const size_t vv = (size_t)V * V;
  const size_t np = (size_t)V * tiles * V;
  DevBuf<double> p_sum, p_pp, p_np;
  DevBuf<long long> p_cnt;
  AMGB_TRY(p_sum.alloc(ctx, np));
  AMGB_TRY(p_pp.alloc(ctx, np));
  AMGB_TRY(p_np.alloc(ctx, np));
  AMGB_TRY(p_cnt.alloc(ctx, np));
  const size_t smem = (size_t)kPoolWarps * V * (3 * sizeof(double) + sizeof(int));
  // (the limit is per function and device, shared by every host thread: always the maximum)
  AMGB_CUDA(ctx, cudaFuncSetAttribute(pool_entries_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmemMax));
  AMGB_CUDA(ctx, cudaFuncSetAttribute(pool_entries_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmemMax));
  struct Events {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } ev;
  AMGB_CUDA(ctx, cudaEventCreate(&ev.a));
  AMGB_CUDA(ctx, cudaEventCreate(&ev.b));
  AMGB_CUDA(ctx, cudaEventRecord(ev.a, ctx->stream));
  // pooling reads 12*nnz + 4*(n+1), writes 28*V^2
  const double bytes = 12.0 * M.nnz + 4.0 * (M.n + 1) + 28.0 * vv;
  const bool vec = (reinterpret_cast<uintptr_t>(M.col.p) % 16 == 0) && (reinterpret_cast<uintptr_t>(M.val.p) % 16 == 0);
  if (vec) {
    AMGB_LAUNCH(ctx, F_POOL, bytes, pool_entries_kernel<true>, (unsigned)(V * tiles), kPoolBlock, smem, bm, tiles,
                (long long)g0, (int)M.n, M.rp.p, M.col.p, M.val.p, p_sum.p, p_cnt.p, p_pp.p, p_np.p);
  } else {
    AMGB_LAUNCH(ctx, F_POOL, bytes, pool_entries_kernel<false>, (unsigned)(V * tiles), kPoolBlock, smem, bm, tiles,
                (long long)g0, (int)M.n, M.rp.p, M.col.p, M.val.p, p_sum.p, p_cnt.p, p_pp.p, p_np.p);
  }
  AMGB_LAUNCH(ctx, F_POOL, 28.0 * np + 28.0 * vv, pool_finalize_kernel, (unsigned)div_up(vv, kPoolBlock),
              kPoolBlock, 0, V, tiles, p_sum.p, p_cnt.p, p_pp.p, p_np.p, d_sum, d_cnt, d_pp, d_np);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaEventRecord(ev.b, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev.a, ev.b);
  if (ms_out) *ms_out = ms;
  return AMGB_OK;
}

extern "C" int amgb_make_view(amgb_ctx* ctx, const amgb_matrix* A, int32_t view_size, double* sum,
                              int64_t* count, double* max_pp, double* max_np, double* t_us) {
  if (!ctx || !A || !sum || !count || !max_pp || !max_np || view_size < 1) return AMGB_ERR_BAD_ARG;
  if (view_size > kMaxView)
    return set_error(ctx, AMGB_ERR_UNSUPPORTED, "view_size %d > %d", view_size, kMaxView);
  cudaSetDevice(ctx->device);
  const size_t vv = (size_t)view_size * view_size;
  DevBuf<double> d_sum, d_pp, d_np;
  DevBuf<long long> d_cnt;
  AMGB_TRY(d_sum.alloc(ctx, vv));
  AMGB_TRY(d_pp.alloc(ctx, vv));
  AMGB_TRY(d_np.alloc(ctx, vv));
  AMGB_TRY(d_cnt.alloc(ctx, vv));
  float ms = 0.f;
  AMGB_TRY(pool_to_device(ctx, A->A, A->A.n, 0, view_size, d_sum.p, d_cnt.p, d_pp.p, d_np.p, &ms));
  AMGB_CUDA(ctx, cudaMemcpyAsync(sum, d_sum.p, vv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(count, d_cnt.p, vv * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(max_pp, d_pp.p, vv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(max_np, d_np.p, vv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (t_us) *t_us = (double)ms * 1000.0;
  return AMGB_OK;
}

// Pooled image of a ROW-PARTITIONED matrix (ref common/view_maker.h:41-65 on an MPI matrix): every
// rank pools the rows it owns into the global V x V bins (global row and column ids), then the four
// channels are combined over the ranks in rank order -- sums and counts added, maxima maxed -- so
// every rank returns the same image.  Collective.
extern "C" int amgb_dist_make_view(amgb_ctx* ctx, const amgb_dist_matrix* A, int32_t view_size, double* sum,
                                   int64_t* count, double* max_pp, double* max_np, double* t_us) {
  if (!ctx || !A || !sum || !count || !max_pp || !max_np || view_size < 1) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  amgb_comm* comm = A->comm;
  const size_t vv = (size_t)view_size * view_size;
  int rc = view_size > kMaxView ? set_error(ctx, AMGB_ERR_UNSUPPORTED, "view_size %d > %d", view_size, kMaxView) : AMGB_OK;
  // layout of one rank's contribution: [sum | max_pp | max_np | count (as int64)]
  std::vector<double> mine(4 * vv, 0.0);
  float ms = 0.f;
  if (rc == AMGB_OK) {
    DevBuf<double> d_sum, d_pp, d_np;
    DevBuf<long long> d_cnt;
    rc = d_sum.alloc(ctx, vv);
    if (rc == AMGB_OK) rc = d_pp.alloc(ctx, vv);
    if (rc == AMGB_OK) rc = d_np.alloc(ctx, vv);
    if (rc == AMGB_OK) rc = d_cnt.alloc(ctx, vv);
    if (rc == AMGB_OK)
      rc = pool_to_device(ctx, A->own.M, A->own.n_global, A->own.g0, view_size, d_sum.p, d_cnt.p, d_pp.p, d_np.p, &ms);
    if (rc == AMGB_OK) {
      cudaError_t e = cudaMemcpyAsync(mine.data(), d_sum.p, vv * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(mine.data() + vv, d_pp.p, vv * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(mine.data() + 2 * vv, d_np.p, vv * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(mine.data() + 3 * vv, d_cnt.p, vv * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "pooled image download", __FILE__, __LINE__);
    }
  }
  // a local failure is agreed on before the exchange, so no rank waits for one that has left
  int64_t good = rc == AMGB_OK ? 1 : 0;
  const int arc = allreduce_min_i64_host(ctx, comm, &good);
  if (rc == AMGB_OK && arc != AMGB_OK) rc = arc;
  if (rc == AMGB_OK && !good) rc = set_error(ctx, AMGB_ERR_COMM, "another rank could not pool its rows");
  if (rc != AMGB_OK) return rc;
  std::vector<double> all(4 * vv * (size_t)comm->size);
  AMGB_TRY(comm->allgather_host(ctx, mine.data(), mine.size() * sizeof(double), all.data()));
  for (size_t i = 0; i < vv; ++i) {
    double s = 0.0, pp = 0.0, np = 0.0;
    int64_t c = 0;
    for (int q = 0; q < comm->size; ++q) {
      const double* r = all.data() + (size_t)q * 4 * vv;
      s += r[i];
      pp = pp > r[vv + i] ? pp : r[vv + i];
      np = np > r[2 * vv + i] ? np : r[2 * vv + i];
      c += reinterpret_cast<const int64_t*>(r + 3 * vv)[i];
    }
    sum[i] = s;
    max_pp[i] = pp;
    max_np[i] = np;
    count[i] = c;
  }
  double t = ms * 1000.0;  // device time of the local pass, max over the ranks
  std::vector<double> ts(comm->size);
  AMGB_TRY(comm->allgather_host(ctx, &t, sizeof t, ts.data()));
  for (double v : ts) t = v > t ? v : t;
  if (t_us) *t_us = t;
  return AMGB_OK;
}

namespace amgb {

// ref data-modeling/train_ann.py:133-172 (norm_view) applied per channel, and the
// channels-last stacking of "sum+max+c" (:247-256).  One block per channel: transform,
// max |.| over the image (order-independent), scale.
__device__ __forceinline__ double view_transform(int mode, double x, double cnt) {
  const double resc = cnt > 0 ? x / cnt : 0.0;
  switch (mode) {
    case AMGB_NORM_RESC:
    case AMGB_NORM_MEAN: return resc;
    case AMGB_NORM_PURE_LOG: return log(fabs(x) + 1.0) * ((x > 0) - (x < 0));
    case AMGB_NORM_RESC_LOG: return log(fabs(resc) + 1.0) * ((resc > 0) - (resc < 0));
    default: return x;  // pure, nothing
  }
}

__global__ void __launch_bounds__(kPoolBlock)
view_normalize_kernel(int vv, int mode, int count_as_reference, const double* __restrict__ sum,
                      const long long* __restrict__ cnt, const double* __restrict__ pp,
                      const double* __restrict__ np, double* __restrict__ out) {
  __shared__ double red[kPoolWarps];
  __shared__ double vmax;
  const int c = blockIdx.x;
  const double* src = c == 0 ? sum : (c == 1 ? pp : np);
  const bool from_count = c == 3 && !count_as_reference;
  double m = 0.0;
  for (int i = threadIdx.x; i < vv; i += kPoolBlock) {
    const double x = from_count ? (double)cnt[i] : src[i];
    m = fmax(m, fabs(view_transform(mode, x, (double)cnt[i])));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kPoolWarps; ++w) t = fmax(t, red[w]);
    vmax = t;
  }
  __syncthreads();
  const bool scaled = mode == AMGB_NORM_PURE || mode == AMGB_NORM_RESC || mode == AMGB_NORM_PURE_LOG ||
                      mode == AMGB_NORM_RESC_LOG;
  const double den = vmax;
  for (int i = threadIdx.x; i < vv; i += kPoolBlock) {
    const double x = from_count ? (double)cnt[i] : src[i];
    const double t = view_transform(mode, x, (double)cnt[i]);
    out[(size_t)i * 4 + c] = scaled ? t / den : t;  // 0/0 -> NaN like numpy (the reference then rejects the image)
  }
}

}  // namespace amgb

extern "C" int amgb_make_view_normalized(amgb_ctx* ctx, const amgb_matrix* A, int32_t view_size, int32_t mode,
                                         int32_t count_channel_as_reference, double* out, double* t_us) {
  if (!ctx || !A || !out || view_size < 1) return AMGB_ERR_BAD_ARG;
  if (mode < AMGB_NORM_NOTHING || mode > AMGB_NORM_MEAN) return AMGB_ERR_BAD_ARG;
  if (view_size > kMaxView)
    return set_error(ctx, AMGB_ERR_UNSUPPORTED, "view_size %d > %d", view_size, kMaxView);
  cudaSetDevice(ctx->device);
  const size_t vv = (size_t)view_size * view_size;
  DevBuf<double> d_sum, d_pp, d_np, d_out;
  DevBuf<long long> d_cnt;
  AMGB_TRY(d_sum.alloc(ctx, vv));
  AMGB_TRY(d_pp.alloc(ctx, vv));
  AMGB_TRY(d_np.alloc(ctx, vv));
  AMGB_TRY(d_cnt.alloc(ctx, vv));
  AMGB_TRY(d_out.alloc(ctx, vv * 4));
  float ms = 0.f;
  AMGB_TRY(pool_to_device(ctx, A->A, A->A.n, 0, view_size, d_sum.p, d_cnt.p, d_pp.p, d_np.p, &ms));
  AMGB_LAUNCH(ctx, F_POOL, 60.0 * vv, view_normalize_kernel, 4, kPoolBlock, 0, (int)vv, mode,
              count_channel_as_reference, d_sum.p, d_cnt.p, d_pp.p, d_np.p, d_out.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, vv * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (t_us) *t_us = (double)ms * 1000.0;
  return AMGB_OK;
}
