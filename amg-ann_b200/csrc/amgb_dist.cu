// Row-partitioned setup: index spaces, exchange plans, ghost-row fetches and the
// partitioned hierarchy driver.  The numerical stages are the single-device ones
// (amgb_setup.cu) run on each rank's extended index space; see amgb_dist.cuh and
// DESIGN.md "Row partition" for why the owned outputs are bit-identical to the
// single-device hierarchy for any number of ranks.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "amgb_dist.cuh"

namespace amgb {

constexpr int kBlock = 256;

static unsigned grid_for(int64_t n) { return (unsigned)div_up(n, kBlock); }

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
minmax_kernel(int64_t n, const int32_t* __restrict__ a, int32_t* __restrict__ mm) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  int lo = INT_MAX, hi = INT_MIN;
  if (i < n) lo = hi = a[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

__global__ void __launch_bounds__(kBlock)
mark_ids_kernel(int64_t n, const int32_t* __restrict__ ids, int64_t wlo, int32_t* __restrict__ mark) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) mark[ids[i] - wlo] = 1;
}

__global__ void __launch_bounds__(kBlock)
mark_range_kernel(int64_t b, int64_t e, int64_t wlo, int32_t* __restrict__ mark) {
  const int64_t i = b + (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < e) mark[i - wlo] = 1;
}

__global__ void __launch_bounds__(kBlock)
compact_ids_kernel(int64_t wlen, int64_t wlo, const int32_t* __restrict__ mark, const int32_t* __restrict__ pos,
                   int32_t* __restrict__ gid) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < wlen && mark[i]) gid[pos[i]] = (int32_t)(wlo + i);
}

__global__ void lower_bound_kernel(int64_t n, const int32_t* __restrict__ sorted, int nq, const int64_t* __restrict__ q,
                                   int64_t* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)sorted[mid] < q[t]) lo = mid + 1; else hi = mid;
  }
  out[t] = lo;
}

__global__ void __launch_bounds__(kBlock)
req_to_index_kernel(int64_t n, const int32_t* __restrict__ req_gid, int64_t g0, int64_t own_begin,
                    int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) idx[i] = (int32_t)(own_begin + ((int64_t)req_gid[i] - g0));
}

template <class T>
__global__ void __launch_bounds__(kBlock)
pack_kernel(int64_t n, const int32_t* __restrict__ idx, const T* __restrict__ src, T* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

__global__ void __launch_bounds__(kBlock)
pack_split_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ lo,
                  const double* __restrict__ hi, int split, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) {
    const int p = idx[i];
    out[i] = p < split ? lo[p] : hi[p];
  }
}

__global__ void __launch_bounds__(kBlock)
pack_rowlen_kernel(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ rp,
                   int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) out[i] = rp[idx[i] + 1] - rp[idx[i]];
}

// one warp per requested row: copy its entries into the packed send arrays
__global__ void __launch_bounds__(kBlock)
pack_rows_kernel(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ rp,
                 const int32_t* __restrict__ col, const double* __restrict__ val,
                 const int32_t* __restrict__ out_off, int32_t* __restrict__ out_col, double* __restrict__ out_val) {
  const int64_t k = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= n) return;
  const int b = rp[idx[k]], len = rp[idx[k] + 1] - b, o = out_off[k];
  for (int t = lane; t < len; t += 32) {
    out_col[o + t] = col[b + t];
    out_val[o + t] = val[b + t];
  }
}

__global__ void gather_i32_kernel(int n, const int32_t* __restrict__ a, const int64_t* __restrict__ at,
                                  int32_t* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = a[at[t]];
}

// values of an int32 device array at a few host-given positions
static int read_at(amgb_ctx* ctx, const int32_t* a, const std::vector<int64_t>& at, std::vector<int32_t>& out) {
  const int n = (int)at.size();
  out.resize(n);
  if (n == 0) return AMGB_OK;
  DevBuf<int64_t> dat;
  DevBuf<int32_t> dout;
  AMGB_TRY(dat.alloc(ctx, n));
  AMGB_TRY(dout.alloc(ctx, n));
  AMGB_CUDA(ctx, cudaMemcpyAsync(dat.p, at.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * n, gather_i32_kernel, (unsigned)div_up(n, 128), 128, 0, n, a, dat.p, dout.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaMemcpyAsync(out.data(), dout.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Dense-window index space: mark global ids, scan, compact.  The result is the sorted
// list of distinct ids, and pos[g - wlo] is the index of id g in it.
// ---------------------------------------------------------------------------
struct IndexSpace {
  int64_t wlo = 0, wlen = 0, count = 0;
  DevBuf<int32_t> mark, pos, gid;
  int begin(amgb_ctx* ctx, int64_t lo, int64_t hi_inclusive) {
    wlo = lo;
    wlen = hi_inclusive - lo + 1;
    if (wlen < 0) wlen = 0;
    AMGB_TRY(mark.alloc_zero(ctx, wlen));
    AMGB_TRY(pos.alloc(ctx, wlen + 1));
    return AMGB_OK;
  }
  int add_ids(amgb_ctx* ctx, const int32_t* ids, int64_t n) {
    AMGB_LAUNCH(ctx, F_AUX, 8.0 * n, mark_ids_kernel, grid_for(n), kBlock, 0, n, ids, wlo, mark.p);
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  }
  int add_range(amgb_ctx* ctx, int64_t b, int64_t e) {
    AMGB_LAUNCH(ctx, F_AUX, 4.0 * (e - b), mark_range_kernel, grid_for(e - b), kBlock, 0, b, e, wlo, mark.p);
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  }
  int finish(amgb_ctx* ctx) {
    AMGB_TRY(exclusive_scan_i32(ctx, mark.p, pos.p, wlen));
    int32_t c = 0;
    AMGB_TRY(read_i32(ctx, pos.p + wlen, &c));
    count = c;
    AMGB_TRY(gid.alloc(ctx, count));
    AMGB_LAUNCH(ctx, F_AUX, 12.0 * wlen, compact_ids_kernel, grid_for(wlen), kBlock, 0, wlen, wlo, mark.p, pos.p,
                gid.p);
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  }
};

// min / max over several id arrays (skips empty ones); returns lo > hi if all are empty
static int minmax_ids(amgb_ctx* ctx, std::initializer_list<std::pair<const int32_t*, int64_t>> arrays, int64_t* lo,
                      int64_t* hi) {
  DevBuf<int32_t> mm;
  AMGB_TRY(mm.alloc(ctx, 2));
  const int32_t init[2] = {INT_MAX, INT_MIN};
  AMGB_CUDA(ctx, cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  for (auto& a : arrays)
    if (a.second > 0) AMGB_LAUNCH(ctx, F_AUX, 4.0 * a.second, minmax_kernel, grid_for(a.second), kBlock, 0, a.second,
                                  a.first, mm.p);
  AMGB_CHECK_LAUNCH(ctx);
  int32_t h[2];
  AMGB_CUDA(ctx, cudaMemcpyAsync(h, mm.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *lo = h[0];
  *hi = h[1];
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Exchange plans
// ---------------------------------------------------------------------------
// gid: sorted global ids of my index space; owned ids are [starts[rank], starts[rank+1])
// and sit at [own_begin, own_begin + n_own).
static int build_plan(amgb_ctx* ctx, amgb_comm* comm, const int32_t* gid, int64_t n_space, int64_t own_begin,
                      int64_t n_own, const std::vector<int64_t>& starts, HaloPlan& pl) {
  const int S = comm->size, me = comm->rank;
  pl.size = S;
  pl.rank = me;
  pl.n_space = n_space;
  pl.own_begin = own_begin;
  pl.n_own = n_own;
  // run of every owner in my index space
  DevBuf<int64_t> dq, dout;
  AMGB_TRY(dq.alloc(ctx, S + 1));
  AMGB_TRY(dout.alloc(ctx, S + 1));
  AMGB_CUDA(ctx, cudaMemcpyAsync(dq.p, starts.data(), (S + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_LAUNCH(ctx, F_AUX, 16.0 * (S + 1), lower_bound_kernel, 1, 64 > S + 1 ? 64 : S + 1, 0, n_space, gid, S + 1, dq.p,
              dout.p);
  AMGB_CHECK_LAUNCH(ctx);
  std::vector<int64_t> pos(S + 1);
  AMGB_CUDA(ctx, cudaMemcpyAsync(pos.data(), dout.p, (S + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  pl.recv_off.assign(S, 0);
  pl.recv_cnt.assign(S, 0);
  pl.recv_total = 0;
  for (int q = 0; q < S; ++q) {
    pl.recv_off[q] = pos[q];
    pl.recv_cnt[q] = q == me ? 0 : pos[q + 1] - pos[q];
    pl.recv_total += pl.recv_cnt[q];
  }
  if (pos[me] != own_begin || pos[me + 1] - pos[me] != n_own)
    return set_error(ctx, AMGB_ERR_COMM, "index space does not hold the owned range contiguously");
  // what every rank needs from every other
  std::vector<int64_t> all((size_t)S * S);
  AMGB_TRY(comm->allgather_host(ctx, pl.recv_cnt.data(), S * sizeof(int64_t), all.data()));
  pl.send_off.assign(S, 0);
  pl.send_cnt.assign(S, 0);
  pl.send_total = 0;
  for (int p = 0; p < S; ++p) {
    pl.send_cnt[p] = all[(size_t)p * S + me];
    pl.send_off[p] = pl.send_total;
    pl.send_total += pl.send_cnt[p];
  }
  // send the requested ids to their owners
  DevBuf<int32_t> req;
  AMGB_TRY(req.alloc(ctx, pl.send_total));
  std::vector<size_t> sc(S), sd(S), rc(S), rd(S);
  for (int q = 0; q < S; ++q) {
    sc[q] = (size_t)pl.recv_cnt[q] * 4;
    sd[q] = (size_t)pl.recv_off[q] * 4;
    rc[q] = (size_t)pl.send_cnt[q] * 4;
    rd[q] = (size_t)pl.send_off[q] * 4;
  }
  AMGB_TRY(comm->alltoallv(ctx, gid, sc.data(), sd.data(), req.p, rc.data(), rd.data()));
  AMGB_TRY(pl.send_idx.alloc(ctx, pl.send_total));
  AMGB_LAUNCH(ctx, F_AUX, 8.0 * pl.send_total, req_to_index_kernel, grid_for(pl.send_total), kBlock, 0, pl.send_total,
              req.p, starts[me], own_begin, pl.send_idx.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_TRY(pl.sendbuf.alloc(ctx, (size_t)pl.send_total * 8));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `req` is freed in stream order; keep the plan simple
  return AMGB_OK;
}

static int plan_exchange(amgb_ctx* ctx, amgb_comm* comm, HaloPlan& pl, void* dst, int eb) {
  const int S = pl.size;
  std::vector<size_t> sc(S), sd(S), rc(S), rd(S);
  for (int q = 0; q < S; ++q) {
    sc[q] = (size_t)pl.send_cnt[q] * eb;
    sd[q] = (size_t)pl.send_off[q] * eb;
    rc[q] = (size_t)pl.recv_cnt[q] * eb;
    rd[q] = (size_t)pl.recv_off[q] * eb;
  }
  return comm->alltoallv(ctx, pl.sendbuf.p, sc.data(), sd.data(), dst, rc.data(), rd.data());
}

int plan_sync(amgb_ctx* ctx, amgb_comm* comm, HaloPlan& pl, void* per_point, int eb) {
  if (eb == 4) {
    AMGB_LAUNCH(ctx, F_AUX, 12.0 * pl.send_total, pack_kernel<int32_t>, grid_for(pl.send_total), kBlock, 0,
                pl.send_total, pl.send_idx.p, (const int32_t*)per_point, (int32_t*)pl.sendbuf.p);
  } else {
    AMGB_LAUNCH(ctx, F_AUX, 20.0 * pl.send_total, pack_kernel<double>, grid_for(pl.send_total), kBlock, 0,
                pl.send_total, pl.send_idx.p, (const double*)per_point, (double*)pl.sendbuf.p);
  }
  AMGB_CHECK_LAUNCH(ctx);
  return plan_exchange(ctx, comm, pl, per_point, eb);
}

int plan_sync_split(amgb_ctx* ctx, amgb_comm* comm, HaloPlan& pl, const double* lo, const double* hi, int split,
                    double* dst) {
  AMGB_LAUNCH(ctx, F_EXCHANGE, 20.0 * pl.send_total, pack_split_kernel, grid_for(pl.send_total), kBlock, 0, pl.send_total,
              pl.send_idx.p, lo, hi, split, (double*)pl.sendbuf.p);
  AMGB_CHECK_LAUNCH(ctx);
  return plan_exchange(ctx, comm, pl, dst, 8);
}

// Rows of the non-owned points of the plan's index space, from their owners.  `rp` is
// indexed by my index space (rows of my owned points are the ones served); the result has
// one row per non-owned point, in index-space order, with the owners' column ids.
static int fetch_rows(amgb_ctx* ctx, amgb_comm* comm, HaloPlan& pl, const int32_t* rp, const int32_t* col,
                      const double* val, DeviceCsr& ghost) {
  const int S = pl.size, me = pl.rank;
  const int64_t ns = pl.send_total, nr = pl.recv_total;
  // ghost row index of an owner's run
  std::vector<int64_t> goff(S);
  for (int q = 0; q < S; ++q) goff[q] = q <= me ? pl.recv_off[q] : pl.recv_off[q] - pl.n_own;
  DevBuf<int32_t> slen, soff, rlen;
  AMGB_TRY(slen.alloc(ctx, ns));
  AMGB_TRY(soff.alloc(ctx, ns + 1));
  AMGB_TRY(rlen.alloc(ctx, nr));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * ns, pack_rowlen_kernel, grid_for(ns), kBlock, 0, ns, pl.send_idx.p, rp, slen.p);
  AMGB_CHECK_LAUNCH(ctx);
  std::vector<size_t> sc(S), sd(S), rc(S), rd(S);
  for (int q = 0; q < S; ++q) {
    sc[q] = (size_t)pl.send_cnt[q] * 4;
    sd[q] = (size_t)pl.send_off[q] * 4;
    rc[q] = (size_t)pl.recv_cnt[q] * 4;
    rd[q] = (size_t)goff[q] * 4;
  }
  AMGB_TRY(comm->alltoallv(ctx, slen.p, sc.data(), sd.data(), rlen.p, rc.data(), rd.data()));
  ghost.n = nr;
  AMGB_TRY(ghost.rp.alloc(ctx, nr + 1));
  AMGB_TRY(exclusive_scan_i32(ctx, slen.p, soff.p, ns));
  AMGB_TRY(exclusive_scan_i32(ctx, rlen.p, ghost.rp.p, nr));
  // entry offsets at the run boundaries
  std::vector<int64_t> at;
  for (int q = 0; q < S; ++q) at.push_back(pl.send_off[q]);
  at.push_back(ns);
  std::vector<int32_t> sb, rb;
  AMGB_TRY(read_at(ctx, soff.p, at, sb));
  at.clear();
  for (int q = 0; q < S; ++q) at.push_back(goff[q]);
  at.push_back(nr);
  AMGB_TRY(read_at(ctx, ghost.rp.p, at, rb));
  // runs are contiguous and in rank order (my own run is empty): run q ends where q+1 starts
  const int64_t send_entries = sb[S], recv_entries = rb[S];
  ghost.nnz = recv_entries;
  AMGB_TRY(ghost.col.alloc(ctx, recv_entries));
  AMGB_TRY(ghost.val.alloc(ctx, recv_entries));
  DevBuf<int32_t> scol;
  DevBuf<double> sval;
  AMGB_TRY(scol.alloc(ctx, send_entries));
  AMGB_TRY(sval.alloc(ctx, send_entries));
  AMGB_LAUNCH(ctx, F_AUX, 24.0 * send_entries, pack_rows_kernel, grid_for(ns * 32), kBlock, 0, ns, pl.send_idx.p, rp,
              col, val, soff.p, scol.p, sval.p);
  AMGB_CHECK_LAUNCH(ctx);
  // send runs are contiguous per peer (send_off ascending); ghost runs likewise
  for (int pass = 0; pass < 2; ++pass) {
    const size_t eb = pass == 0 ? 4 : 8;
    for (int q = 0; q < S; ++q) {
      const int64_t s_b = sb[q], s_e = sb[q + 1];
      const int64_t r_b = rb[q], r_e = rb[q + 1];
      sc[q] = pl.send_cnt[q] ? (size_t)(s_e - s_b) * eb : 0;
      sd[q] = (size_t)s_b * eb;
      rc[q] = pl.recv_cnt[q] ? (size_t)(r_e - r_b) * eb : 0;
      rd[q] = (size_t)r_b * eb;
    }
    if (pass == 0) AMGB_TRY(comm->alltoallv(ctx, scol.p, sc.data(), sd.data(), ghost.col.p, rc.data(), rd.data()));
    else AMGB_TRY(comm->alltoallv(ctx, sval.p, sc.data(), sd.data(), ghost.val.p, rc.data(), rd.data()));
  }
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Extended local matrix of a level
// ---------------------------------------------------------------------------
// source of extended row e: owned row, ghost row (index in `ghost`), or none
__global__ void __launch_bounds__(kBlock)
ext_row_len_kernel(int64_t next, const int32_t* __restrict__ gid, int64_t g0, int64_t nloc,
                   const int32_t* __restrict__ own_rp, int64_t wlo1, int64_t wlen1, const int32_t* __restrict__ mark1,
                   const int32_t* __restrict__ pos1, int64_t own_begin1, const int32_t* __restrict__ ghost_rp,
                   int32_t* __restrict__ len, int32_t* __restrict__ src) {
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= next) return;
  const int64_t g = gid[e];
  int l = 0, s = INT_MIN;  // s >= 0: owned row s; s < 0 (not INT_MIN): ghost row -s-1
  if (g >= g0 && g < g0 + nloc) {
    s = (int)(g - g0);
    l = own_rp[s + 1] - own_rp[s];
  } else if (g >= wlo1 && g < wlo1 + wlen1 && mark1[g - wlo1]) {
    const int e1 = pos1[g - wlo1];
    const int r = e1 < own_begin1 ? e1 : e1 - (int)nloc;
    s = -r - 1;
    l = ghost_rp[r + 1] - ghost_rp[r];
  }
  len[e] = l;
  src[e] = s;
}

// one warp per extended row; column ids are mapped through pos[gid - wlo]
__global__ void __launch_bounds__(kBlock)
ext_fill_kernel(int64_t next, const int32_t* __restrict__ src, const int32_t* __restrict__ own_rp,
                const int32_t* __restrict__ own_col, const double* __restrict__ own_val,
                const int32_t* __restrict__ ghost_rp, const int32_t* __restrict__ ghost_col,
                const double* __restrict__ ghost_val, int64_t wlo, const int32_t* __restrict__ pos,
                const int32_t* __restrict__ rp, int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t e = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= next) return;
  const int s = src[e];
  if (s == INT_MIN) return;
  const int32_t* scol;
  const double* sval;
  int b, len;
  if (s >= 0) {
    b = own_rp[s];
    len = own_rp[s + 1] - b;
    scol = own_col;
    sval = own_val;
  } else {
    const int r = -s - 1;
    b = ghost_rp[r];
    len = ghost_rp[r + 1] - b;
    scol = ghost_col;
    sval = ghost_val;
  }
  const int o = rp[e];
  for (int t = lane; t < len; t += 32) {
    col[o + t] = pos[scol[b + t] - wlo];
    val[o + t] = sval[b + t];
  }
}

__global__ void __launch_bounds__(kBlock)
map_ids_kernel(int64_t n, int32_t* __restrict__ ids, const int32_t* __restrict__ table) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) ids[i] = table[ids[i]];
}

// Builds E, the extended square matrix Ahat (rows O U H1 complete, H2 rows empty) and the
// per-point plan of a level from the owned rows.  `extra`: further global ids that must be
// points of E (columns of the previous level's owned P rows).
static int make_ext_level(amgb_ctx* ctx, amgb_comm* comm, DistLevel& D, const int32_t* extra, int64_t n_extra,
                          DeviceCsr& Ahat) {
  const OwnedCsr& own = D.own;
  const int64_t g0 = own.g0, nloc = own.M.n;
  D.nloc = nloc;
  // stage 1: owned U H1 (U extra)
  int64_t lo, hi;
  AMGB_TRY(minmax_ids(ctx, {{own.M.col.p, own.M.nnz}, {extra, n_extra}}, &lo, &hi));
  if (nloc > 0) {
    lo = std::min<int64_t>(lo, g0);
    hi = std::max<int64_t>(hi, g0 + nloc - 1);
  }
  if (lo > hi) {  // nothing owned, nothing referenced
    lo = g0;
    hi = g0 - 1;
  }
  IndexSpace s1;
  AMGB_TRY(s1.begin(ctx, lo, hi));
  AMGB_TRY(s1.add_range(ctx, g0, g0 + nloc));
  AMGB_TRY(s1.add_ids(ctx, own.M.col.p, own.M.nnz));
  AMGB_TRY(s1.add_ids(ctx, extra, n_extra));
  AMGB_TRY(s1.finish(ctx));
  int32_t ob1 = 0;
  {
    std::vector<int32_t> v;
    // position of the first id >= g0 (valid also when nothing is owned)
    DevBuf<int64_t> dq, dout;
    AMGB_TRY(dq.alloc(ctx, 1));
    AMGB_TRY(dout.alloc(ctx, 1));
    AMGB_CUDA(ctx, cudaMemcpyAsync(dq.p, &g0, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    AMGB_LAUNCH(ctx, F_AUX, 16.0, lower_bound_kernel, 1, 32, 0, s1.count, s1.gid.p, 1, dq.p, dout.p);
    AMGB_CHECK_LAUNCH(ctx);
    int64_t p = 0;
    AMGB_TRY(read_i64(ctx, dout.p, &p));
    ob1 = (int32_t)p;
  }
  HaloPlan plan1;
  AMGB_TRY(build_plan(ctx, comm, s1.gid.p, s1.count, ob1, nloc, own.starts, plan1));
  DeviceCsr ghost;
  // rows are served by index-space position: shift the owned row pointer accordingly
  AMGB_TRY(fetch_rows(ctx, comm, plan1, own.M.rp.p - ob1, own.M.col.p, own.M.val.p, ghost));
  // stage 2: add the ghost rows' columns (H2)
  int64_t lo2, hi2;
  AMGB_TRY(minmax_ids(ctx, {{ghost.col.p, ghost.nnz}}, &lo2, &hi2));
  if (lo2 <= hi2) {
    lo = std::min(lo, lo2);
    hi = std::max(hi, hi2);
  }
  IndexSpace s2;
  AMGB_TRY(s2.begin(ctx, lo, hi));
  AMGB_TRY(s2.add_range(ctx, g0, g0 + nloc));
  AMGB_TRY(s2.add_ids(ctx, own.M.col.p, own.M.nnz));
  AMGB_TRY(s2.add_ids(ctx, extra, n_extra));
  AMGB_TRY(s2.add_ids(ctx, ghost.col.p, ghost.nnz));
  AMGB_TRY(s2.finish(ctx));
  D.next = s2.count;
  {
    DevBuf<int64_t> dq, dout;
    AMGB_TRY(dq.alloc(ctx, 1));
    AMGB_TRY(dout.alloc(ctx, 1));
    AMGB_CUDA(ctx, cudaMemcpyAsync(dq.p, &g0, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    AMGB_LAUNCH(ctx, F_AUX, 16.0, lower_bound_kernel, 1, 32, 0, s2.count, s2.gid.p, 1, dq.p, dout.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(read_i64(ctx, dout.p, &D.o0));
  }
  // Ahat
  const int64_t next = D.next;
  DevBuf<int32_t> len, src;
  AMGB_TRY(len.alloc(ctx, next));
  AMGB_TRY(src.alloc(ctx, next));
  AMGB_LAUNCH(ctx, F_AUX, 16.0 * next, ext_row_len_kernel, grid_for(next), kBlock, 0, next, s2.gid.p, g0, nloc,
              own.M.rp.p, s1.wlo, s1.wlen, s1.mark.p, s1.pos.p, (int64_t)ob1, ghost.rp.p, len.p, src.p);
  AMGB_CHECK_LAUNCH(ctx);
  Ahat.n = Ahat.ncols = next;
  AMGB_TRY(Ahat.rp.alloc(ctx, next + 1));
  AMGB_TRY(exclusive_scan_i32(ctx, len.p, Ahat.rp.p, next));
  int32_t nnz = 0;
  AMGB_TRY(read_i32(ctx, Ahat.rp.p + next, &nnz));
  Ahat.nnz = nnz;
  AMGB_TRY(Ahat.col.alloc(ctx, nnz));
  AMGB_TRY(Ahat.val.alloc(ctx, nnz));
  AMGB_LAUNCH(ctx, F_AUX, 24.0 * nnz, ext_fill_kernel, grid_for(next * 32), kBlock, 0, next, src.p, own.M.rp.p,
              own.M.col.p, own.M.val.p, ghost.rp.p, ghost.col.p, ghost.val.p, s2.wlo, s2.pos.p, Ahat.rp.p, Ahat.col.p,
              Ahat.val.p);
  AMGB_CHECK_LAUNCH(ctx);
  D.gid = std::move(s2.gid);
  AMGB_TRY(build_plan(ctx, comm, D.gid.p, next, D.o0, nloc, own.starts, D.plan));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Solve-phase vector plan and coarsest-level gathers
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
req_to_table_kernel(int64_t n, const int32_t* __restrict__ req_gid, int64_t g0, const int32_t* __restrict__ table,
                    int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) idx[i] = table[(int64_t)req_gid[i] - g0];
}

int build_vector_plan(amgb_ctx* ctx, amgb_comm* comm, const int32_t* halo_gid, int64_t n_halo, int64_t n_own,
                      const std::vector<int64_t>& starts, const int32_t* own_index, HaloPlan& pl) {
  const int S = comm->size, me = comm->rank;
  pl.size = S;
  pl.rank = me;
  pl.n_space = n_own + n_halo;
  pl.own_begin = 0;
  pl.n_own = n_own;
  DevBuf<int64_t> dq, dout;
  AMGB_TRY(dq.alloc(ctx, S + 1));
  AMGB_TRY(dout.alloc(ctx, S + 1));
  AMGB_CUDA(ctx, cudaMemcpyAsync(dq.p, starts.data(), (S + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_LAUNCH(ctx, F_AUX, 16.0 * (S + 1), lower_bound_kernel, 1, 64 > S + 1 ? 64 : S + 1, 0, n_halo, halo_gid, S + 1,
              dq.p, dout.p);
  AMGB_CHECK_LAUNCH(ctx);
  std::vector<int64_t> pos(S + 1);
  AMGB_CUDA(ctx, cudaMemcpyAsync(pos.data(), dout.p, (S + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  pl.recv_off.assign(S, 0);
  pl.recv_cnt.assign(S, 0);
  pl.recv_total = 0;
  for (int q = 0; q < S; ++q) {
    pl.recv_off[q] = n_own + pos[q];
    pl.recv_cnt[q] = pos[q + 1] - pos[q];
    pl.recv_total += pl.recv_cnt[q];
  }
  if (pl.recv_cnt[me] != 0) return set_error(ctx, AMGB_ERR_COMM, "halo list contains owned points");
  std::vector<int64_t> all((size_t)S * S);
  AMGB_TRY(comm->allgather_host(ctx, pl.recv_cnt.data(), S * sizeof(int64_t), all.data()));
  pl.send_off.assign(S, 0);
  pl.send_cnt.assign(S, 0);
  pl.send_total = 0;
  for (int p = 0; p < S; ++p) {
    pl.send_cnt[p] = all[(size_t)p * S + me];
    pl.send_off[p] = pl.send_total;
    pl.send_total += pl.send_cnt[p];
  }
  DevBuf<int32_t> req;
  AMGB_TRY(req.alloc(ctx, pl.send_total));
  std::vector<size_t> sc(S), sd(S), rc(S), rd(S);
  for (int q = 0; q < S; ++q) {
    sc[q] = (size_t)pl.recv_cnt[q] * 4;
    sd[q] = (size_t)pos[q] * 4;
    rc[q] = (size_t)pl.send_cnt[q] * 4;
    rd[q] = (size_t)pl.send_off[q] * 4;
  }
  AMGB_TRY(comm->alltoallv(ctx, halo_gid, sc.data(), sd.data(), req.p, rc.data(), rd.data()));
  AMGB_TRY(pl.send_idx.alloc(ctx, pl.send_total));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * pl.send_total, req_to_table_kernel, grid_for(pl.send_total), kBlock, 0,
              pl.send_total, req.p, starts[me], own_index, pl.send_idx.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_TRY(pl.sendbuf.alloc(ctx, (size_t)pl.send_total * 8));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

__global__ void __launch_bounds__(kBlock)
row_len_kernel(int64_t n, const int32_t* __restrict__ rp, int32_t* __restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) len[i] = rp[i + 1] - rp[i];
}

int allgather_rows(amgb_ctx* ctx, amgb_comm* comm, const OwnedCsr& own, DeviceCsr& full) {
  const int S = comm->size;
  const int64_t n = own.n_global, nloc = own.M.n;
  int64_t mine[1] = {own.M.nnz};
  std::vector<int64_t> nnzs(S);
  AMGB_TRY(comm->allgather_host(ctx, mine, sizeof(int64_t), nnzs.data()));
  std::vector<int64_t> eoff(S + 1, 0);
  for (int q = 0; q < S; ++q) eoff[q + 1] = eoff[q] + nnzs[q];
  DevBuf<int32_t> len, lens;
  AMGB_TRY(len.alloc(ctx, nloc));
  AMGB_TRY(lens.alloc(ctx, n));
  AMGB_LAUNCH(ctx, F_AUX, 8.0 * nloc, row_len_kernel, grid_for(nloc), kBlock, 0, nloc, own.M.rp.p, len.p);
  AMGB_CHECK_LAUNCH(ctx);
  full.n = full.ncols = n;
  full.nnz = eoff[S];
  AMGB_TRY(full.rp.alloc(ctx, n + 1));
  AMGB_TRY(full.col.alloc(ctx, full.nnz));
  AMGB_TRY(full.val.alloc(ctx, full.nnz));
  std::vector<size_t> sc(S), sd(S, 0), rc(S), rd(S);
  for (int q = 0; q < S; ++q) {
    sc[q] = (size_t)nloc * 4;
    rc[q] = (size_t)(own.starts[q + 1] - own.starts[q]) * 4;
    rd[q] = (size_t)own.starts[q] * 4;
  }
  AMGB_TRY(comm->alltoallv(ctx, len.p, sc.data(), sd.data(), lens.p, rc.data(), rd.data()));
  AMGB_TRY(exclusive_scan_i32(ctx, lens.p, full.rp.p, n));
  for (int pass = 0; pass < 2; ++pass) {
    const size_t eb = pass == 0 ? 4 : 8;
    for (int q = 0; q < S; ++q) {
      sc[q] = (size_t)own.M.nnz * eb;
      rc[q] = (size_t)nnzs[q] * eb;
      rd[q] = (size_t)eoff[q] * eb;
    }
    if (pass == 0) AMGB_TRY(comm->alltoallv(ctx, own.M.col.p, sc.data(), sd.data(), full.col.p, rc.data(), rd.data()));
    else AMGB_TRY(comm->alltoallv(ctx, own.M.val.p, sc.data(), sd.data(), full.val.p, rc.data(), rd.data()));
  }
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int allgather_f64(amgb_ctx* ctx, amgb_comm* comm, const std::vector<int64_t>& starts, const double* mine,
                  double* full) {
  const int S = comm->size, me = comm->rank;
  std::vector<size_t> sc(S), sd(S, 0), rc(S), rd(S);
  for (int q = 0; q < S; ++q) {
    sc[q] = (size_t)(starts[me + 1] - starts[me]) * 8;
    rc[q] = (size_t)(starts[q + 1] - starts[q]) * 8;
    rd[q] = (size_t)starts[q] * 8;
  }
  return comm->alltoallv(ctx, mine, sc.data(), sd.data(), full, rc.data(), rd.data());
}

struct Hooks : DistHooks {
  amgb_ctx* ctx;
  amgb_comm* comm;
  HaloPlan* plan;
  int sync_i32(int32_t* a) override { return plan_sync(ctx, comm, *plan, a, 4); }
  int sync_f64(double* a) override { return plan_sync(ctx, comm, *plan, a, 8); }
  int allreduce_sum(int64_t* v) override { return allreduce_sum_i64_host(ctx, comm, v); }
};

// global coarse id of the owned C points, -1 elsewhere (ghost entries are then synced)
__global__ void __launch_bounds__(kBlock)
coarse_gid_kernel(int64_t next, int64_t o0, int64_t nloc, const int32_t* __restrict__ cf,
                  const int32_t* __restrict__ f2c, int64_t c0, int32_t* __restrict__ cgid) {
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= next) return;
  int v = -1;
  if (e >= o0 && e < o0 + nloc && cf[e] > 0) v = (int)(c0 + (f2c[e] - f2c[o0]));
  cgid[e] = v;
}

// Phat: rows over E (owned from Pown, others from the fetched ghost rows), columns mapped
// to the compact coarse space
__global__ void __launch_bounds__(kBlock)
phat_len_kernel(int64_t next, int64_t o0, int64_t nloc, const int32_t* __restrict__ own_rp,
                const int32_t* __restrict__ ghost_rp, int32_t* __restrict__ len) {
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= next) return;
  if (e >= o0 && e < o0 + nloc) {
    len[e] = own_rp[e + 1] - own_rp[e];
  } else {
    const int64_t r = e < o0 ? e : e - nloc;
    len[e] = ghost_rp[r + 1] - ghost_rp[r];
  }
}

__global__ void __launch_bounds__(kBlock)
phat_fill_kernel(int64_t next, int64_t o0, int64_t nloc, const int32_t* __restrict__ own_rp,
                 const int32_t* __restrict__ own_col, const double* __restrict__ own_val,
                 const int32_t* __restrict__ ghost_rp, const int32_t* __restrict__ ghost_col,
                 const double* __restrict__ ghost_val, int64_t wlo, const int32_t* __restrict__ pos,
                 const int32_t* __restrict__ rp, int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= next) return;
  const int32_t* scol;
  const double* sval;
  int b, len;
  if (e >= o0 && e < o0 + nloc) {
    b = own_rp[e];
    len = own_rp[e + 1] - b;
    scol = own_col;
    sval = own_val;
  } else {
    const int64_t r = e < o0 ? e : e - nloc;
    b = ghost_rp[r];
    len = ghost_rp[r + 1] - b;
    scol = ghost_col;
    sval = ghost_val;
  }
  const int o = rp[e];
  for (int t = 0; t < len; ++t) {
    col[o + t] = pos[scol[b + t] - wlo];
    val[o + t] = sval[b + t];
  }
}

static int copy_csr_rows(amgb_ctx* ctx, const DeviceCsr& src, int64_t row0, int64_t rows, DeviceCsr& dst);

__global__ void __launch_bounds__(kBlock)
rebase_rp_kernel(int64_t n1, const int32_t* __restrict__ in, int32_t base, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n1) out[i] = in[i] - base;
}

// dst <- rows [row0, row0+rows) of src (deep copy, row pointer rebased to 0)
static int copy_csr_rows(amgb_ctx* ctx, const DeviceCsr& src, int64_t row0, int64_t rows, DeviceCsr& dst) {
  std::vector<int32_t> b;
  AMGB_TRY(read_at(ctx, src.rp.p, {row0, row0 + rows}, b));
  dst.n = rows;
  dst.ncols = src.ncols;
  dst.nnz = b[1] - b[0];
  AMGB_TRY(dst.rp.alloc(ctx, rows + 1));
  AMGB_TRY(dst.col.alloc(ctx, dst.nnz));
  AMGB_TRY(dst.val.alloc(ctx, dst.nnz));
  AMGB_LAUNCH(ctx, F_AUX, 8.0 * rows, rebase_rp_kernel, grid_for(rows + 1), kBlock, 0, rows + 1, src.rp.p + row0, b[0],
              dst.rp.p);
  AMGB_CHECK_LAUNCH(ctx);
  if (dst.nnz) {
    AMGB_CUDA(ctx, cudaMemcpyAsync(dst.col.p, src.col.p + b[0], dst.nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                   ctx->stream));
    AMGB_CUDA(ctx, cudaMemcpyAsync(dst.val.p, src.val.p + b[0], dst.nnz * sizeof(double), cudaMemcpyDeviceToDevice,
                                   ctx->stream));
  }
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// The partitioned hierarchy driver (mirror of build_hierarchy in amgb_setup.cu)
// ---------------------------------------------------------------------------
int build_hierarchy_dist(amgb_precond* P) {
  amgb_ctx* ctx = P->ctx;
  amgb_dist_state* ds = P->dist;
  amgb_comm* comm = ds->comm;
  const int S = comm->size, me = comm->rank;
  AMGB_TRY(resolve_options(P));
  const amgb_boomeramg_data& d = P->data;
  P->lv.clear();
  ds->dl.clear();
  ds->replicated_from = 1 << 30;
  P->lv.reserve(d.max_levels + 1);
  ds->dl.reserve(d.max_levels + 1);
  P->st_rows.clear();
  P->st_nnz.clear();
  P->st_nnzP.clear();
  // level 0: the user's slab (deep-copied pointers are wrapped, not owned)
  ds->dl.emplace_back();
  {
    const OwnedCsr& u = ds->mat->own;
    OwnedCsr& o = ds->dl[0].own;
    o.n_global = u.n_global;
    o.g0 = u.g0;
    o.starts = u.starts;
    o.M.n = u.M.n;
    o.M.ncols = u.M.ncols;
    o.M.nnz = u.M.nnz;
    o.M.rp.wrap(ctx, u.M.rp.p, u.M.rp.n);
    o.M.col.wrap(ctx, u.M.col.p, u.M.col.n);
    o.M.val.wrap(ctx, u.M.val.p, u.M.val.n);
  }
  DevBuf<int32_t> extra;  // global coarse ids referenced by the previous level's owned P rows
  int64_t n_extra = 0;
  for (int level = 0;; ++level) {
    ctx->cur_level = level;
    DistLevel& D = ds->dl[level];
    P->lv.emplace_back();
    Level& L = P->lv[level];
    int64_t nnz_global = D.own.M.nnz;
    AMGB_TRY(allreduce_sum_i64_host(ctx, comm, &nnz_global));
    P->st_rows.push_back(D.own.n_global);
    P->st_nnz.push_back(nnz_global);
    P->st_nnzP.push_back(0);
    if (level > 0 && level < d.max_levels - 1 && D.own.n_global > d.max_coarse_size &&
        D.own.n_global <= d.dist_replicate_below) {
      // small level: gather it on every rank and finish the hierarchy there with the
      // single-device code (identical on all ranks, and identical to the single-device
      // hierarchy); no exchanges from here down
      AMGB_TRY(allgather_rows(ctx, comm, D.own, L.A));
      ds->replicated_from = level;
      AMGB_TRY(build_levels_from(P, level));
      for (size_t l = (size_t)level; l < P->lv.size(); ++l) {
        if (l > (size_t)level) {
          P->st_rows.push_back(P->lv[l].A.n);
          P->st_nnz.push_back(P->lv[l].A.nnz);
          P->st_nnzP.push_back(0);
        }
        P->st_nnzP[l] = P->lv[l].P.nnz;
      }
      break;
    }
    AMGB_TRY(make_ext_level(ctx, comm, D, extra.p, n_extra, L.A));
    if (level == d.max_levels - 1 || D.own.n_global <= d.max_coarse_size) break;
    const int64_t next = D.next, o0 = D.o0, nloc = D.nloc;
    Hooks hooks;
    hooks.ctx = ctx;
    hooks.comm = comm;
    hooks.plan = &D.plan;
    hooks.own_begin = o0;
    hooks.own_end = o0 + nloc;
    hooks.gid = D.gid.p;
    DevBuf<int32_t> has_strong;
    DevBuf<double> diagv;
    AMGB_TRY(L.mask.alloc(ctx, L.A.nnz));
    AMGB_TRY(has_strong.alloc(ctx, next));
    AMGB_TRY(diagv.alloc(ctx, next));
    AMGB_TRY(L.cf.alloc(ctx, next));
    AMGB_TRY(run_strength(ctx, L.A, P->theta_eff, P->mrs_eff, L.mask.p, has_strong.p, diagv.p));
    // H2 rows are empty here, so their has_strong / diagonal come from the owners
    AMGB_TRY(hooks.sync_i32(has_strong.p));
    AMGB_TRY(hooks.sync_f64(diagv.p));
    AMGB_TRY(coarsen_pmis(ctx, L.A, L.mask.p, has_strong.p, L.cf.p, &hooks));
    AMGB_TRY(L.f2c.alloc(ctx, next + 1));
    int32_t nce = 0;
    AMGB_TRY(number_coarse_points(ctx, next, L.cf.p, L.f2c.p, &nce));
    std::vector<int32_t> fb;
    AMGB_TRY(read_at(ctx, L.f2c.p, {o0, o0 + nloc}, fb));
    D.nc_own = fb[1] - fb[0];
    std::vector<int64_t> counts(S);
    AMGB_TRY(comm->allgather_host(ctx, &D.nc_own, sizeof(int64_t), counts.data()));
    D.cstarts.assign(S + 1, 0);
    for (int q = 0; q < S; ++q) D.cstarts[q + 1] = D.cstarts[q] + counts[q];
    D.nc_global = D.cstarts[S];
    D.c0 = D.cstarts[me];
    if (D.nc_global == 0 || D.nc_global == D.own.n_global) {
      L.mask.release();
      L.cf.release();
      L.f2c.release();
      break;
    }
    L.n_coarse = D.nc_own;
    // interpolation of the owned rows, columns = global coarse ids
    DevBuf<int32_t> cgid;
    AMGB_TRY(cgid.alloc(ctx, next));
    AMGB_LAUNCH(ctx, F_INTERP, 12.0 * next, coarse_gid_kernel, grid_for(next), kBlock, 0, next, o0, nloc, L.cf.p,
                L.f2c.p, D.c0, cgid.p);
    AMGB_CHECK_LAUNCH(ctx);
    AMGB_TRY(hooks.sync_i32(cgid.p));
    AMGB_TRY(build_interp(ctx, L.A, L.mask.p, L.cf.p, cgid.p, diagv.p, o0, o0 + nloc, D.nc_global, D.Pown));
    int64_t nnzp_global = D.Pown.nnz;
    AMGB_TRY(allreduce_sum_i64_host(ctx, comm, &nnzp_global));
    P->st_nnzP.back() = nnzp_global;
    // ghost rows of P (H1 and H2), compact coarse column space, Phat
    DeviceCsr ghostP;
    AMGB_TRY(fetch_rows(ctx, comm, D.plan, D.Pown.rp.p, D.Pown.col.p, D.Pown.val.p, ghostP));
    int64_t clo, chi;
    AMGB_TRY(minmax_ids(ctx, {{D.Pown.col.p, D.Pown.nnz}, {ghostP.col.p, ghostP.nnz}}, &clo, &chi));
    if (clo > chi) {
      clo = D.c0;
      chi = D.c0 - 1;
    }
    IndexSpace cs;
    AMGB_TRY(cs.begin(ctx, clo, chi));
    AMGB_TRY(cs.add_ids(ctx, D.Pown.col.p, D.Pown.nnz));
    AMGB_TRY(cs.add_ids(ctx, ghostP.col.p, ghostP.nnz));
    AMGB_TRY(cs.finish(ctx));
    const int64_t nct = cs.count;
    int64_t tc0 = 0;
    {
      DevBuf<int64_t> dq, dout;
      AMGB_TRY(dq.alloc(ctx, 1));
      AMGB_TRY(dout.alloc(ctx, 1));
      AMGB_CUDA(ctx, cudaMemcpyAsync(dq.p, &D.c0, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
      AMGB_LAUNCH(ctx, F_AUX, 16.0, lower_bound_kernel, 1, 32, 0, nct, cs.gid.p, 1, dq.p, dout.p);
      AMGB_CHECK_LAUNCH(ctx);
      AMGB_TRY(read_i64(ctx, dout.p, &tc0));
    }
    DeviceCsr& Phat = L.P;
    {
      DevBuf<int32_t> len;
      AMGB_TRY(len.alloc(ctx, next));
      AMGB_LAUNCH(ctx, F_AUX, 12.0 * next, phat_len_kernel, grid_for(next), kBlock, 0, next, o0, nloc, D.Pown.rp.p,
                  ghostP.rp.p, len.p);
      Phat.n = next;
      Phat.ncols = nct;
      AMGB_TRY(Phat.rp.alloc(ctx, next + 1));
      AMGB_TRY(exclusive_scan_i32(ctx, len.p, Phat.rp.p, next));
      int32_t nnz = 0;
      AMGB_TRY(read_i32(ctx, Phat.rp.p + next, &nnz));
      Phat.nnz = nnz;
      AMGB_TRY(Phat.col.alloc(ctx, nnz));
      AMGB_TRY(Phat.val.alloc(ctx, nnz));
      AMGB_LAUNCH(ctx, F_AUX, 24.0 * nnz, phat_fill_kernel, grid_for(next), kBlock, 0, next, o0, nloc, D.Pown.rp.p,
                  D.Pown.col.p, D.Pown.val.p, ghostP.rp.p, ghostP.col.p, ghostP.val.p, cs.wlo, cs.pos.p, Phat.rp.p,
                  Phat.col.p, Phat.val.p);
      AMGB_CHECK_LAUNCH(ctx);
    }
    // Galerkin product: That = Ahat * Phat (rows O U H1), A_c(owned coarse rows) = Rhat * That
    DeviceCsr Rhat, T;
    AMGB_TRY(transpose_csr(ctx, Phat, Rhat));
    AMGB_TRY(spgemm(ctx, L.A, Phat, T, false));
    // owned coarse rows of Rhat are the contiguous compact range [tc0, tc0 + nc_own)
    AMGB_TRY(copy_csr_rows(ctx, Rhat, tc0, D.nc_own, L.R));
    ds->dl.emplace_back();
    DistLevel& Dn = ds->dl[level + 1];
    OwnedCsr& on = Dn.own;
    on.n_global = D.nc_global;
    on.g0 = D.c0;
    on.starts = D.cstarts;
    AMGB_TRY(spgemm(ctx, L.R, T, on.M, true));
    // compact coarse columns -> global coarse ids
    AMGB_LAUNCH(ctx, F_AUX, 12.0 * on.M.nnz, map_ids_kernel, grid_for(on.M.nnz), kBlock, 0, on.M.nnz, on.M.col.p,
                cs.gid.p);
    AMGB_CHECK_LAUNCH(ctx);
    on.M.ncols = D.nc_global;
    // the next level must know the coarse points my P rows reference
    n_extra = D.Pown.nnz;
    AMGB_TRY(extra.alloc(ctx, n_extra));
    if (n_extra)
      AMGB_CUDA(ctx, cudaMemcpyAsync(extra.p, D.Pown.col.p, n_extra * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                     ctx->stream));
    // keep the compact coarse ids of this level for the solve-phase column maps
    D.tc_gid = std::move(cs.gid);
    D.nct = nct;
    D.tc0 = tc0;
    if (!d.keep_setup_intermediates) L.mask.release();
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->cur_level = 0;
  AMGB_TRY(finish_solve_setup_dist(P));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

}  // namespace amgb

using namespace amgb;

void amgb_dist_state_destroy(amgb_dist_state* s) { delete s; }

extern "C" {

// COLLECTIVE, including its failures: every rank first contributes its row range and a local
// verdict to one all-gather, so a rank with bad arguments cannot leave while its peers wait in a
// collective for it -- all ranks return an error together (the offending rank with its own
// message, the others AMGB_ERR_COMM naming it).
int amgb_dist_matrix_create(amgb_ctx* ctx, amgb_comm* comm, int64_t n_global, int64_t row_begin, int64_t row_end,
                            const int64_t* rowptr_local, const int32_t* col_global, const double* val,
                            amgb_dist_matrix** out) {
  if (!ctx || !comm || !out) return AMGB_ERR_BAD_ARG;  // (no communicator to agree on anything)
  *out = nullptr;
  cudaSetDevice(ctx->device);
  int local = AMGB_OK;
  int64_t nloc = 0, nnz = 0;
  if (!rowptr_local || row_begin < 0 || row_end < row_begin || row_end > n_global) {
    local = set_error(ctx, AMGB_ERR_BAD_ARG, "bad row range [%lld, %lld) of %lld or null row pointers",
                      (long long)row_begin, (long long)row_end, (long long)n_global);
  } else if (n_global >= (int64_t(1) << 31)) {
    local = set_error(ctx, AMGB_ERR_RANGE, "n_global=%lld: global ids are 32-bit", (long long)n_global);
  } else {
    nloc = row_end - row_begin;
    nnz = rowptr_local[nloc];
    if (nnz >= (int64_t(1) << 31))
      local = set_error(ctx, AMGB_ERR_RANGE, "local nnz=%lld does not fit 32-bit row pointers: use more ranks",
                        (long long)nnz);
    else if (nnz > 0 && (!col_global || !val))
      local = set_error(ctx, AMGB_ERR_BAD_ARG, "null column / value arrays");
  }
  const int64_t mine[2] = {row_begin, (int64_t)local};
  std::vector<int64_t> all(2 * (size_t)comm->size);
  AMGB_TRY(comm->allgather_host(ctx, mine, sizeof mine, all.data()));
  for (int q = 0; q < comm->size; ++q)
    if (all[2 * q + 1] != AMGB_OK) {
      if (local != AMGB_OK) return local;
      return set_error(ctx, AMGB_ERR_COMM, "rank %d rejected its part of the matrix (status %lld)", q,
                       (long long)all[2 * q + 1]);
    }
  for (int q = 0; q + 1 < comm->size; ++q)
    if (all[2 * q] > all[2 * (q + 1)])  // (the same verdict on every rank: computed from the gathered ranges)
      return set_error(ctx, AMGB_ERR_BAD_ARG, "row ranges must ascend with the rank");
  amgb_dist_matrix* M = new amgb_dist_matrix;
  M->ctx = ctx;
  M->comm = comm;
  OwnedCsr& o = M->own;
  o.n_global = n_global;
  o.g0 = row_begin;
  o.starts.assign(comm->size + 1, 0);
  for (int q = 0; q < comm->size; ++q) o.starts[q] = all[2 * q];
  o.starts[comm->size] = n_global;
  o.M.n = nloc;
  o.M.ncols = n_global;
  o.M.nnz = nnz;
  std::vector<int32_t> rp32(nloc + 1);
  for (int64_t i = 0; i <= nloc; ++i) rp32[i] = (int32_t)rowptr_local[i];
  int rc = o.M.rp.alloc(ctx, nloc + 1);
  if (rc == AMGB_OK) rc = o.M.col.alloc(ctx, nnz);
  if (rc == AMGB_OK) rc = o.M.val.alloc(ctx, nnz);
  cudaError_t e = cudaSuccess;
  if (rc == AMGB_OK) {
    e = cudaMemcpyAsync(o.M.rp.p, rp32.data(), (nloc + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz)
      e = cudaMemcpyAsync(o.M.col.p, col_global, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz)
      e = cudaMemcpyAsync(o.M.val.p, val, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "distributed matrix upload", __FILE__, __LINE__);
  }
  // an allocation or copy failure is a local event too: agree on it before anyone goes on
  int64_t good = rc == AMGB_OK ? 1 : 0;
  const int arc = allreduce_min_i64_host(ctx, comm, &good);
  if (rc == AMGB_OK && arc != AMGB_OK) rc = arc;
  if (rc == AMGB_OK && !good) rc = set_error(ctx, AMGB_ERR_COMM, "another rank could not upload its part of the matrix");
  if (rc != AMGB_OK) {
    delete M;
    return rc;
  }
  *out = M;
  return AMGB_OK;
}

int amgb_dist_matrix_destroy(amgb_dist_matrix* A) {
  if (!A) return AMGB_OK;
  cudaSetDevice(A->ctx->device);
  delete A;
  return AMGB_OK;
}

int amgb_dist_precond_initialize(amgb_ctx* ctx, const amgb_dist_matrix* A, const amgb_boomeramg_data* data,
                                 amgb_precond** out) {
  if (!ctx || !A || !data || !out) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  amgb_precond* P = new amgb_precond;
  P->ctx = ctx;
  P->mat = nullptr;
  P->data = *data;
  // The V-cycle (kernels + halo exchanges) can be captured in a CUDA graph when the
  // communicator only enqueues stream work (NCCL).  Measured on 2 x B200 (m=160) the graph
  // with NCCL send/recv nodes is ~15 % SLOWER than plain launches, so it is opt-in.
  P->use_graph = A->comm->capturable() && std::getenv("AMGB_DIST_GRAPH") != nullptr;
  P->graph_loop = std::getenv("AMGB_PCG_GRAPH_LOOP") != nullptr;
  P->dist = new amgb_dist_state;
  P->dist->comm = A->comm;
  P->dist->mat = A;
  const int rc = build_hierarchy_dist(P);
  if (rc != AMGB_OK) {
    cudaStreamSynchronize(ctx->stream);
    (void)cudaGetLastError();
    amgb_precond_destroy(P);
    return rc;
  }
  *out = P;
  return AMGB_OK;
}

// ---- parity accessors: the OWNED part of every level, global ids ----
int amgb_dist_precond_level_dims(const amgb_precond* P, int32_t level, int64_t* n_global, int64_t* row_begin,
                                 int64_t* n_local, int64_t* nnz_local, int64_t* n_coarse_global,
                                 int64_t* coarse_begin, int64_t* nnz_P_local) {
  if (!P || !P->dist) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->dist->dl.size()) return AMGB_ERR_RANGE;
  const DistLevel& D = P->dist->dl[level];
  if (n_global) *n_global = D.own.n_global;
  if (row_begin) *row_begin = D.own.g0;
  if (n_local) *n_local = D.own.M.n;
  if (nnz_local) *nnz_local = D.own.M.nnz;
  if (n_coarse_global) *n_coarse_global = D.nc_global;
  if (coarse_begin) *coarse_begin = D.c0;
  if (nnz_P_local) *nnz_P_local = D.Pown.nnz;
  return AMGB_OK;
}

static int d2h_sync(amgb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return AMGB_OK;
  AMGB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int amgb_dist_precond_replicated_from(const amgb_precond* P, int32_t* level) {
  if (!P || !P->dist || !level) return AMGB_ERR_BAD_ARG;
  const int nl = (int)P->lv.size();
  *level = P->dist->replicated_from < nl ? P->dist->replicated_from : nl;
  return AMGB_OK;
}

int amgb_dist_precond_get_cf_marker(const amgb_precond* P, int32_t level, int32_t* cf_local) {
  if (!P || !P->dist || !cf_local) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->dist->dl.size() || level >= P->dist->replicated_from) return AMGB_ERR_RANGE;
  const DistLevel& D = P->dist->dl[level];
  const Level& L = P->lv[level];
  if (level + 1 >= (int)P->lv.size())
    return set_error(P->ctx, AMGB_ERR_RANGE, "level %d is the coarsest: no C/F splitting", level);
  if (D.nloc == 0) return AMGB_OK;  // this rank owns nothing on the level
  cudaSetDevice(P->ctx->device);
  return d2h_sync(P->ctx, cf_local, L.cf.p + D.o0, D.nloc * sizeof(int32_t));
}

int amgb_dist_precond_get_A_rows(const amgb_precond* P, int32_t level, int32_t* rowptr_local, int32_t* col_global,
                                 double* val) {
  if (!P || !P->dist) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->dist->dl.size()) return AMGB_ERR_RANGE;
  const DeviceCsr& M = P->dist->dl[level].own.M;
  cudaSetDevice(P->ctx->device);
  if (rowptr_local) AMGB_TRY(d2h_sync(P->ctx, rowptr_local, M.rp.p, (M.n + 1) * sizeof(int32_t)));
  if (col_global) AMGB_TRY(d2h_sync(P->ctx, col_global, M.col.p, M.nnz * sizeof(int32_t)));
  if (val) AMGB_TRY(d2h_sync(P->ctx, val, M.val.p, M.nnz * sizeof(double)));
  return AMGB_OK;
}

// rowptr_local has n_local+1 entries (rebased to 0); columns are global coarse ids
int amgb_dist_precond_get_P_rows(const amgb_precond* P, int32_t level, int32_t* rowptr_local, int32_t* col_global,
                                 double* val) {
  if (!P || !P->dist) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level + 1 >= (int)P->dist->dl.size() || level >= P->dist->replicated_from) return AMGB_ERR_RANGE;
  const DistLevel& D = P->dist->dl[level];
  const DeviceCsr& M = D.Pown;  // rows over E; the owned block is [o0, o0 + nloc)
  cudaSetDevice(P->ctx->device);
  std::vector<int32_t> rp(D.nloc + 1);
  AMGB_TRY(d2h_sync(P->ctx, rp.data(), M.rp.p + D.o0, (D.nloc + 1) * sizeof(int32_t)));
  const int32_t base = rp[0], nnz = rp[D.nloc] - base;
  if (rowptr_local)
    for (int64_t i = 0; i <= D.nloc; ++i) rowptr_local[i] = rp[i] - base;
  if (col_global) AMGB_TRY(d2h_sync(P->ctx, col_global, M.col.p + base, (size_t)nnz * sizeof(int32_t)));
  if (val) AMGB_TRY(d2h_sync(P->ctx, val, M.val.p + base, (size_t)nnz * sizeof(double)));
  return AMGB_OK;
}

}  // extern "C"
