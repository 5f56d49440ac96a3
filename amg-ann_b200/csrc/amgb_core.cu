// Context, matrix residency, error reporting, prefix sums and the measurement
// hooks of libamgb.so.  ABI documentation: include/amgb.h.
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "amgb_internal.cuh"

namespace amgb {

int set_error(amgb_ctx* ctx, int status, const char* fmt, ...) {
  if (ctx) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    ctx->err = buf;
  }
  return status;
}

int cuda_fail(amgb_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
  const int st = (e == cudaErrorMemoryAllocation) ? AMGB_ERR_OOM : AMGB_ERR_CUDA;
  return set_error(ctx, st, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
}

int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

LaunchScope::LaunchScope(amgb_ctx* c, int family, double bytes) : ctx(c), timed(false) {
  ctx->launches++;
  ctx->fam_launches[family]++;
  ctx->fam_bytes[family] += bytes;
  if (ctx->timers_on) {
    timed = true;
    rec.family = family;
    rec.bytes = bytes;
    rec.level = ctx->cur_level < 0 ? 0 : (ctx->cur_level >= kTimerLevels ? kTimerLevels - 1 : ctx->cur_level);
    auto get = [&](cudaEvent_t* ev) {
      if (!ctx->free_events.empty()) {
        *ev = ctx->free_events.back();
        ctx->free_events.pop_back();
      } else {
        cudaEventCreate(ev);
      }
    };
    get(&rec.a);
    get(&rec.b);
    cudaEventRecord(rec.a, ctx->stream);
  }
}

LaunchScope::~LaunchScope() {
  if (timed) {
    cudaEventRecord(rec.b, ctx->stream);
    ctx->recs.push_back(rec);
  }
}

static void drain_timers(amgb_ctx* ctx) {
  if (ctx->recs.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ctx->fam_ms[r.family] += ms;
      ctx->lvl_ms[r.family][r.level] += ms;
      ctx->lvl_bytes[r.family][r.level] += r.bytes;
      ctx->lvl_launches[r.family][r.level] += 1;
    }
    ctx->free_events.push_back(r.a);
    ctx->free_events.push_back(r.b);
  }
  ctx->recs.clear();
}

// ---------------------------------------------------------------------------
// Exclusive scan of int32 counts: three hand-written passes
// (per-block reduce -> scan of block sums -> per-block scan with offset).
// Deterministic; integer, so order does not matter anyway.
// ---------------------------------------------------------------------------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // items per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_sums[kScanBlock / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int s = lane < kScanBlock / 32 ? warp_sums[lane] : 0;
    const int si = warp_incl_scan(s);
    if (lane < kScanBlock / 32) warp_sums[lane] = si - s;
    if (lane == kScanBlock / 32 - 1) block_total = si;
  }
  __syncthreads();
  *total = block_total;
  const int r = incl - v + warp_sums[wid];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(const int32_t* __restrict__ in,
                                                                 int64_t n,
                                                                 int32_t* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + (int64_t)k * kScanBlock + threadIdx.x;
    if (i < n) s += in[i];
  }
  int total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of m block sums; writes grand total to sums[m]
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(int32_t* sums, int64_t m) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < m; base += kScanBlock) {
    const int64_t i = base + threadIdx.x;
    const int v = i < m ? sums[i] : 0;
    int total;
    const int ex = block_excl_scan(v, &total);
    const int carry = carry_s;
    if (i < m) sums[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[m] = carry_s;
}

__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const int32_t* __restrict__ in,
                                                                int64_t n,
                                                                const int32_t* __restrict__ block_sums,
                                                                int64_t nblocks,
                                                                int32_t* __restrict__ out) {
  // thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile (blocked layout)
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  int total;
  int ex = block_excl_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
  if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) out[n] = block_sums[nblocks];
}

// short inputs: the whole scan in ONE block, tile after tile with a running carry (one launch and no
// block-sum buffer instead of three launches: the coarse levels of a hierarchy are launch-bound)
constexpr int64_t kScanSmall = 8 * kScanTile;

__global__ void __launch_bounds__(kScanBlock) scan_small_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                int32_t* __restrict__ out) {
  int carry = 0;
  for (int64_t tile = 0; tile < n; tile += kScanTile) {
    const int64_t base = tile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const int64_t i = base + k;
      v[k] = i < n ? in[i] : 0;
      s += v[k];
    }
    int total;
    int ex = block_excl_scan(s, &total) + carry;  // (ends with a barrier: the tile is read before it is written)
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const int64_t i = base + k;
      if (i < n) out[i] = ex;
      ex += v[k];
    }
    carry += total;
  }
  if (threadIdx.x == 0) out[n] = carry;
}

int exclusive_scan_i32(amgb_ctx* ctx, const int32_t* in, int32_t* out, int64_t n) {
  if (n <= 0) {
    AMGB_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(int32_t), ctx->stream));
    return AMGB_OK;
  }
  if (n <= kScanSmall) {
    AMGB_LAUNCH(ctx, F_SCAN, 8.0 * n, scan_small_kernel, 1, kScanBlock, 0, in, n, out);
    AMGB_CHECK_LAUNCH(ctx);
    return AMGB_OK;
  }
  const int64_t nblocks = div_up(n, kScanTile);
  DevBuf<int32_t> sums;
  AMGB_TRY(sums.alloc(ctx, nblocks + 1));
  AMGB_LAUNCH(ctx, F_SCAN, 4.0 * n, scan_reduce_kernel, (unsigned)nblocks, kScanBlock, 0, in, n, sums.p);
  AMGB_LAUNCH(ctx, F_SCAN, 8.0 * nblocks, scan_sums_kernel, 1, kScanBlock, 0, sums.p, nblocks);
  AMGB_LAUNCH(ctx, F_SCAN, 8.0 * n, scan_apply_kernel, (unsigned)nblocks, kScanBlock, 0, in, n, sums.p,
              nblocks, out);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

int read_i32(amgb_ctx* ctx, const int32_t* dptr, int32_t* host) {
  AMGB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, dptr, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *host = *(int32_t*)ctx->pinned;
  return AMGB_OK;
}

int read_i64(amgb_ctx* ctx, const int64_t* dptr, int64_t* host) {
  AMGB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, dptr, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *host = *(int64_t*)ctx->pinned;
  return AMGB_OK;
}

__global__ void rowptr64_to_32_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out,
                                      int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}

static const char* kFamilyNames[F_COUNT] = {
    "spmv",     "smooth", "residual",  "restrict", "prolong", "vec",  "coarse", "strength",
    "coarsen",  "interp", "transpose", "spgemm",   "scan",    "aux",  "pool",
    "smooth_l0", "residual_l0", "restrict_l0", "prolong_l0", "exchange", "tail"};

static const char* kRouteNames[R_COUNT] = {
    "sell_t1_stream",   "sell_t_multi",      "spgemm_g8_t128",  "spgemm_g8_t256",    "spgemm_g8_t512",
    "spgemm_g32",       "spgemm_sym_big",    "spgemm_sym_global", "spgemm_num_big",  "spgemm_num_global",
    "spgemm_rowreg",    "cycle_graph",       "cycle_tail_fused", "pcg_device_loop",  "dense_stepwise",
    "spgemm_flat"};

}  // namespace amgb

using namespace amgb;

extern "C" {

int amgb_route_count(void) { return R_COUNT; }

const char* amgb_route_name(int route) { return (route >= 0 && route < R_COUNT) ? kRouteNames[route] : ""; }

int amgb_ctx_get_route(const amgb_ctx* ctx, int route, int64_t* count) {
  if (!ctx || !count || route < 0 || route >= R_COUNT) return AMGB_ERR_BAD_ARG;
  *count = ctx->routes[route];
  return AMGB_OK;
}

int amgb_ctx_reset_routes(amgb_ctx* ctx) {
  if (!ctx) return AMGB_ERR_BAD_ARG;
  for (int r = 0; r < R_COUNT; ++r) ctx->routes[r] = 0;
  return AMGB_OK;
}

int amgb_version(void) { return AMGB_VERSION; }

const char* amgb_status_string(int status) {
  switch (status) {
    case AMGB_OK: return "AMGB_OK";
    case AMGB_ERR_BAD_ARG: return "AMGB_ERR_BAD_ARG";
    case AMGB_ERR_NO_DEVICE: return "AMGB_ERR_NO_DEVICE";
    case AMGB_ERR_CUDA: return "AMGB_ERR_CUDA";
    case AMGB_ERR_OOM: return "AMGB_ERR_OOM";
    case AMGB_ERR_UNSUPPORTED: return "AMGB_ERR_UNSUPPORTED";
    case AMGB_ERR_NO_CONVERGENCE: return "AMGB_ERR_NO_CONVERGENCE";
    case AMGB_ERR_BREAKDOWN: return "AMGB_ERR_BREAKDOWN";
    case AMGB_ERR_RANGE: return "AMGB_ERR_RANGE";
    case AMGB_ERR_COMM: return "AMGB_ERR_COMM";
    default: return "AMGB_ERR_UNKNOWN";
  }
}

const char* amgb_last_error(const amgb_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

int amgb_ctx_create(amgb_ctx** out, int device_id, void* stream) {
  if (!out) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    (void)cudaGetLastError();
    return AMGB_ERR_NO_DEVICE;  // no CPU path exists: fail loudly
  }
  if (device_id < 0 || device_id >= count) return AMGB_ERR_BAD_ARG;
  if (cudaSetDevice(device_id) != cudaSuccess) return AMGB_ERR_NO_DEVICE;
  amgb_ctx* ctx = new amgb_ctx;
  ctx->device = device_id;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) {
    delete ctx;
    return AMGB_ERR_NO_DEVICE;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return AMGB_ERR_CUDA;
    }
    ctx->own_stream = true;
  }
  // keep freed blocks in the stream-ordered pool: the theta sweep re-allocates
  // the same hierarchy shapes over and over
  // (a pool of its own per context: see amgb_ctx::pool; the default pool is the fallback)
  uint64_t thr = UINT64_MAX;
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device_id;
  if (!std::getenv("AMGB_SHARED_POOL") && cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
    cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
  } else {
    (void)cudaGetLastError();
    ctx->pool = nullptr;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess)
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  ctx->pinned_bytes = 4096;
  if (cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) != cudaSuccess) {
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return AMGB_ERR_CUDA;
  }
  *out = ctx;
  return AMGB_OK;
}

int amgb_ctx_destroy(amgb_ctx* ctx) {
  if (!ctx) return AMGB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drain_timers(ctx);
  for (auto ev : ctx->free_events) cudaEventDestroy(ev);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);  // blocks still held by live objects stay valid until freed
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AMGB_OK;
}

int amgb_ctx_reserve(amgb_ctx* ctx, int64_t bytes) {
  if (!ctx || bytes < 0) return AMGB_ERR_BAD_ARG;
  if (bytes == 0) return AMGB_OK;
  cudaSetDevice(ctx->device);
  // one allocation of the whole amount, freed at once: the pool keeps the memory (release threshold =
  // infinity), so the hierarchy of the first initialize() is carved out of it instead of growing the
  // pool allocation by allocation
  DevBuf<char> b;
  AMGB_TRY(b.alloc(ctx, (size_t)bytes));
  b.release();
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int amgb_ctx_synchronize(amgb_ctx* ctx) {
  if (!ctx) return AMGB_ERR_BAD_ARG;
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int amgb_ctx_kernel_launches(const amgb_ctx* ctx, int64_t* count) {
  if (!ctx || !count) return AMGB_ERR_BAD_ARG;
  *count = ctx->launches;
  return AMGB_OK;
}

int amgb_ctx_reset_kernel_launches(amgb_ctx* ctx) {
  if (!ctx) return AMGB_ERR_BAD_ARG;
  ctx->launches = 0;
  return AMGB_OK;
}

int amgb_ctx_enable_timers(amgb_ctx* ctx, int enable) {
  if (!ctx) return AMGB_ERR_BAD_ARG;
  drain_timers(ctx);
  ctx->timers_on = enable != 0;
  return AMGB_OK;
}

int amgb_ctx_reset_timers(amgb_ctx* ctx) {
  if (!ctx) return AMGB_ERR_BAD_ARG;
  drain_timers(ctx);
  for (int f = 0; f < F_COUNT; ++f) {
    ctx->fam_ms[f] = 0;
    ctx->fam_launches[f] = 0;
    ctx->fam_bytes[f] = 0;
    for (int l = 0; l < kTimerLevels; ++l) {
      ctx->lvl_ms[f][l] = 0;
      ctx->lvl_bytes[f][l] = 0;
      ctx->lvl_launches[f][l] = 0;
    }
  }
  return AMGB_OK;
}

int amgb_timer_count(void) { return F_COUNT; }

const char* amgb_timer_name(int family) {
  return (family >= 0 && family < F_COUNT) ? kFamilyNames[family] : "";
}

int amgb_ctx_get_timer(amgb_ctx* ctx, int family, double* total_ms, int64_t* launches,
                       double* algorithmic_bytes) {
  if (!ctx || family < 0 || family >= F_COUNT) return AMGB_ERR_BAD_ARG;
  drain_timers(ctx);
  if (total_ms) *total_ms = ctx->fam_ms[family];
  if (launches) *launches = ctx->fam_launches[family];
  if (algorithmic_bytes) *algorithmic_bytes = ctx->fam_bytes[family];
  return AMGB_OK;
}

int amgb_ctx_get_timer_level(amgb_ctx* ctx, int family, int level, double* total_ms, int64_t* launches,
                             double* algorithmic_bytes) {
  if (!ctx || family < 0 || family >= F_COUNT || level < 0 || level >= kTimerLevels) return AMGB_ERR_BAD_ARG;
  drain_timers(ctx);
  if (total_ms) *total_ms = ctx->lvl_ms[family][level];
  if (launches) *launches = ctx->lvl_launches[family][level];
  if (algorithmic_bytes) *algorithmic_bytes = ctx->lvl_bytes[family][level];
  return AMGB_OK;
}

// ---- matrix ---------------------------------------------------------------

static int upload_common(amgb_ctx* ctx, int64_t n, int64_t nnz, const void* rowptr, bool rp64,
                         const int32_t* col, const double* val, amgb_matrix** out) {
  if (!ctx || !rowptr || !col || !val || !out || n < 1) return AMGB_ERR_BAD_ARG;
  if (n >= (int64_t(1) << 31) || nnz >= (int64_t(1) << 31))
    return set_error(ctx, AMGB_ERR_RANGE,
                     "n=%lld nnz=%lld: a single-device matrix needs n, nnz < 2^31 "
                     "(row-partition larger systems)", (long long)n, (long long)nnz);
  cudaSetDevice(ctx->device);
  amgb_matrix* M = new amgb_matrix;
  M->ctx = ctx;
  M->A.n = M->A.ncols = n;
  M->A.nnz = nnz;
  int rc = M->A.rp.alloc(ctx, n + 1);
  if (rc == AMGB_OK) rc = M->A.col.alloc(ctx, nnz);
  if (rc == AMGB_OK) rc = M->A.val.alloc(ctx, nnz);
  if (rc != AMGB_OK) {
    delete M;
    return rc;
  }
  cudaError_t e;
  if (rp64) {
    DevBuf<int64_t> tmp;
    rc = tmp.alloc(ctx, n + 1);
    if (rc != AMGB_OK) { delete M; return rc; }
    e = cudaMemcpyAsync(tmp.p, rowptr, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      AMGB_LAUNCH(ctx, F_AUX, 12.0 * (n + 1), rowptr64_to_32_kernel, (unsigned)div_up(n + 1, 256), 256, 0,
                  tmp.p, M->A.rp.p, n + 1);
      e = cudaGetLastError();
    }
  } else {
    e = cudaMemcpyAsync(M->A.rp.p, rowptr, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(M->A.col.p, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(M->A.val.p, val, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    delete M;
    return cuda_fail(ctx, e, "matrix upload", __FILE__, __LINE__);
  }
  *out = M;
  return AMGB_OK;
}

int amgb_matrix_upload_csr(amgb_ctx* ctx, int64_t n, const int32_t* rowptr, const int32_t* col,
                           const double* val, amgb_matrix** out) {
  if (!rowptr || n < 1) return AMGB_ERR_BAD_ARG;
  return upload_common(ctx, n, rowptr[n], rowptr, false, col, val, out);
}

int amgb_matrix_upload_csr64(amgb_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col,
                             const double* val, amgb_matrix** out) {
  if (!rowptr || n < 1) return AMGB_ERR_BAD_ARG;
  return upload_common(ctx, n, rowptr[n], rowptr, true, col, val, out);
}

int amgb_matrix_wrap_device_csr(amgb_ctx* ctx, int64_t n, int64_t nnz, const int32_t* rowptr_device,
                                const int32_t* col_device, const double* val_device,
                                amgb_matrix** out) {
  if (!ctx || !rowptr_device || !col_device || !val_device || !out || n < 1 || nnz < 0)
    return AMGB_ERR_BAD_ARG;
  if (n >= (int64_t(1) << 31) || nnz >= (int64_t(1) << 31)) return AMGB_ERR_RANGE;
  amgb_matrix* M = new amgb_matrix;
  M->ctx = ctx;
  M->A.n = M->A.ncols = n;
  M->A.nnz = nnz;
  M->A.rp.wrap(ctx, const_cast<int32_t*>(rowptr_device), n + 1);
  M->A.col.wrap(ctx, const_cast<int32_t*>(col_device), nnz);
  M->A.val.wrap(ctx, const_cast<double*>(val_device), nnz);
  *out = M;
  return AMGB_OK;
}

int amgb_matrix_destroy(amgb_matrix* A) {
  if (!A) return AMGB_OK;
  cudaSetDevice(A->ctx->device);
  delete A;
  return AMGB_OK;
}

int amgb_matrix_dims(const amgb_matrix* A, int64_t* n, int64_t* nnz) {
  if (!A) return AMGB_ERR_BAD_ARG;
  if (n) *n = A->A.n;
  if (nnz) *nnz = A->A.nnz;
  return AMGB_OK;
}

int amgb_matrix_vmult(amgb_ctx* ctx, const amgb_matrix* A, double* y, const double* x) {
  if (!ctx || !A || !y || !x) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  const int64_t n = A->A.n;
  DevBuf<double> dx, dy;
  AMGB_TRY(dx.alloc(ctx, n));
  AMGB_TRY(dy.alloc(ctx, n));
  AMGB_CUDA(ctx, cudaMemcpyAsync(dx.p, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_TRY(spmv(ctx, A->A, dx.p, dy.p, F_SPMV));
  AMGB_CUDA(ctx, cudaMemcpyAsync(y, dy.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

int amgb_boomeramg_data_default(amgb_boomeramg_data* d) {
  if (!d) return AMGB_ERR_BAD_ARG;
  std::memset(d, 0, sizeof *d);
  d->symmetric_operator = 0;
  d->strong_threshold = 0.25;
  d->max_row_sum = 0.9;
  d->aggressive_coarsening_num_levels = 0;
  d->output_details = 0;
  d->relaxation_type_up = AMGB_RELAX_SORJacobi;
  d->relaxation_type_down = AMGB_RELAX_SORJacobi;
  d->relaxation_type_coarse = AMGB_RELAX_GaussianElimination;
  d->n_sweeps_coarse = 1;
  d->tol = 0.0;
  d->max_iter = 1;
  d->w_cycle = 0;
  d->coarsen_type = AMGB_COARSEN_PMIS;
  d->interp_type = AMGB_INTERP_CLASSICAL;
  d->relax_order = 1;
  d->n_sweeps = 1;
  d->max_levels = 25;
  d->max_coarse_size = 9;
  d->relax_weight = 1.0;
  d->smoother_policy = AMGB_SMOOTHER_SUBSTITUTE;
  d->options_via_string = 1;
  d->keep_setup_intermediates = 0;
  d->dist_replicate_below = 262144;
  return AMGB_OK;
}

}  // extern "C"
