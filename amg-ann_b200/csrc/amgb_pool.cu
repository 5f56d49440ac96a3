// Matrix -> V x V four-channel image pooling, replacing ViewMaker::make_view
// (ref common/view_maker.h:26-74: a single-threaded MatGetRow loop).
//
// Atomic-free and deterministic:
//   * the V row-bins are contiguous row ranges (ref view_maker.h:41-45,52), so a
//     CTA is given a tile of rows inside ONE row-bin and only needs V-wide
//     accumulators;
//   * each lane walks one CSR row; columns are ascending, so the row decomposes
//     into runs of equal column-bin which the lane sums in CSR order in
//     registers;
//   * the warp advances over the column-bins in lock step (min over the lanes'
//     next bin), reduces the 32 run partials with a fixed shuffle tree and lane 0
//     adds the result to the warp-private shared-memory accumulators -- no two
//     writers ever share an address;
//   * warps, then tiles, are combined in index order.
// count is integer-exact, max_pp / max_np are order-independent hence exact,
// sum has a fixed (input-independent) association.
#include <climits>

#include "amgb_internal.cuh"

namespace amgb {

constexpr int kPoolBlock = 256;
constexpr int kPoolWarps = kPoolBlock / 32;
constexpr int kMaxView = 512;

struct BinMap {
  int q, q1, p, t, V;
  __device__ __forceinline__ int bin(int i) const { return i < t ? i / q1 : (i - t) / q + p; }
  __host__ __device__ int row_begin(int br) const { return br < p ? br * q1 : t + (br - p) * q; }
};

__global__ void __launch_bounds__(kPoolBlock)
pool_tiles_kernel(BinMap bm, int tiles, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
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
  for (int t = threadIdx.x; t < kPoolWarps * V; t += kPoolBlock) {
    s_sum[t] = 0.0;
    s_pp[t] = 0.0;
    s_np[t] = 0.0;
    s_cnt[t] = 0;
  }
  __syncthreads();
  const int r0 = bm.row_begin(br), r1 = bm.row_begin(br + 1);
  const int per = (r1 - r0 + tiles - 1) / tiles;
  const int tb = r0 + tile * per;
  const int te = tb + per < r1 ? tb + per : r1;
  double* w_sum = s_sum + warp * V;
  double* w_pp = s_pp + warp * V;
  double* w_np = s_np + warp * V;
  int* w_cnt = s_cnt + warp * V;
  for (int base = tb + warp * 32; base < te; base += kPoolBlock) {
    const int row = base + lane;
    int k = 0, e = 0;
    if (row < te) {
      k = rp[row];
      e = rp[row + 1];
    }
    int nb = k < e ? bm.bin(col[k]) : INT_MAX;
    for (;;) {
      int cur = nb;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) cur = min(cur, __shfl_xor_sync(0xffffffffu, cur, d));
      if (cur == INT_MAX) break;
      double s = 0.0, pp = 0.0, np = 0.0;
      int c = 0;
      if (nb == cur) {
        int b = cur;
        while (k < e) {
          b = bm.bin(col[k]);
          if (b != cur) break;
          const double v = val[k];
          s += v;
          ++c;
          pp = fmax(pp, fmax(v, 0.0));
          np = fmax(np, fmax(-v, 0.0));
          ++k;
        }
        nb = k < e ? b : INT_MAX;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, d);
        c += __shfl_xor_sync(0xffffffffu, c, d);
        pp = fmax(pp, __shfl_xor_sync(0xffffffffu, pp, d));
        np = fmax(np, __shfl_xor_sync(0xffffffffu, np, d));
      }
      if (lane == 0) {
        w_sum[cur] += s;
        w_cnt[cur] += c;
        w_pp[cur] = fmax(w_pp[cur], pp);
        w_np[cur] = fmax(w_np[cur], np);
      }
    }
  }
  __syncthreads();
  const size_t out = ((size_t)br * tiles + tile) * V;
  for (int bc = threadIdx.x; bc < V; bc += kPoolBlock) {
    double s = 0.0, pp = 0.0, np = 0.0;
    long long c = 0;
#pragma unroll
    for (int w = 0; w < kPoolWarps; ++w) {
      s += s_sum[w * V + bc];
      c += s_cnt[w * V + bc];
      pp = fmax(pp, s_pp[w * V + bc]);
      np = fmax(np, s_np[w * V + bc]);
    }
    part_sum[out + bc] = s;
    part_cnt[out + bc] = c;
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
static int pool_to_device(amgb_ctx* ctx, const amgb_matrix* A, int V, double* d_sum, long long* d_cnt, double* d_pp,
                          double* d_np, float* ms_out) {
  const int n = (int)A->A.n;
  BinMap bm;
  bm.V = V;
  bm.q = n / V;
  bm.q1 = bm.q + 1;
  bm.p = n % V;
  bm.t = bm.q1 * bm.p;
  int tiles = (int)div_up((int64_t)ctx->sm_count * 4, V);
  const int max_rows_per_bin = bm.q1;
  const int max_tiles = (int)div_up(max_rows_per_bin, kPoolBlock);
  if (tiles > max_tiles) tiles = max_tiles;
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
  AMGB_CUDA(ctx, cudaFuncSetAttribute(pool_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  AMGB_CUDA(ctx, cudaEventCreate(&e0));
  AMGB_CUDA(ctx, cudaEventCreate(&e1));
  AMGB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  // pooling reads 12*nnz + 4*(n+1), writes 28*V^2
  const double bytes = 12.0 * A->A.nnz + 4.0 * (n + 1) + 28.0 * vv;
  AMGB_LAUNCH(ctx, F_POOL, bytes, pool_tiles_kernel, (unsigned)(V * tiles), kPoolBlock, smem, bm, tiles,
              A->A.rp.p, A->A.col.p, A->A.val.p, p_sum.p, p_cnt.p, p_pp.p, p_np.p);
  AMGB_LAUNCH(ctx, F_POOL, 28.0 * np + 28.0 * vv, pool_finalize_kernel, (unsigned)div_up(vv, kPoolBlock),
              kPoolBlock, 0, V, tiles, p_sum.p, p_cnt.p, p_pp.p, p_np.p, d_sum, d_cnt, d_pp, d_np);
  cudaError_t le = cudaGetLastError();
  if (le == cudaSuccess) le = cudaEventRecord(e1, ctx->stream);
  if (le == cudaSuccess) le = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  if (le == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (le != cudaSuccess) return cuda_fail(ctx, le, "pooling", __FILE__, __LINE__);
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
  AMGB_TRY(pool_to_device(ctx, A, view_size, d_sum.p, d_cnt.p, d_pp.p, d_np.p, &ms));
  AMGB_CUDA(ctx, cudaMemcpyAsync(sum, d_sum.p, vv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(count, d_cnt.p, vv * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(max_pp, d_pp.p, vv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(max_np, d_np.p, vv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (t_us) *t_us = (double)ms * 1000.0;
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
  AMGB_TRY(pool_to_device(ctx, A, view_size, d_sum.p, d_cnt.p, d_pp.p, d_np.p, &ms));
  AMGB_LAUNCH(ctx, F_POOL, 60.0 * vv, view_normalize_kernel, 4, kPoolBlock, 0, (int)vv, mode,
              count_channel_as_reference, d_sum.p, d_cnt.p, d_pp.p, d_np.p, d_out.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, vv * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (t_us) *t_us = (double)ms * 1000.0;
  return AMGB_OK;
}
