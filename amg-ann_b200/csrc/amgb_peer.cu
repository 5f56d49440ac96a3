// Peer-memory exchanges of the row-partitioned solve phase: halo refresh, the all-gather at
// the replication cut and the PCG scalar reductions as plain kernels that store into the
// other GPUs' memory over NVLink (amgb_dist.cuh "PeerPlan", DESIGN.md "Peer windows").
//
// Protocol of one exchange with epoch e (= completed exchanges of the plan + 1):
//   put     every block stores its share of the send list into the peers' staging slot e&1;
//           the block that finishes last publishes e in my flag word at every peer
//           (bar.sync -> fence.sys -> ticket; last: fence.sys -> st.release.sys)
//   get     a few blocks poll the peers' flag words in MY window (ld.acquire.sys) until all
//           hold >= e, copy slot e&1 of my staging into the destination vector, and the
//           last one advances the plan's counter
// A slot is re-used by exchange e+2; the sender can only get there after it has waited for
// my flag of exchange e+1, which I publish after my unpack of e is complete (stream order).
#include <cstring>

#include "amgb_dist.cuh"

namespace amgb {

constexpr int kBlock = 256;
constexpr long long kSpinLimit = 20000000000ll;  // clock64 ticks (~10 s): a lost peer must not hang the GPU

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long e) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < e) {
    if (clock64() - t0 > kSpinLimit) return false;
    __nanosleep(32);
  }
  return true;
}

__global__ void __launch_bounds__(kBlock)
peer_put_kernel(int n, const int32_t* __restrict__ idx, const double* __restrict__ lo, const double* __restrict__ hi,
                int split, PeerTab t) {
  const unsigned long long e = *(volatile unsigned long long*)t.ctr + 1;
  const long long slot = (long long)(e & 1ull);
  const int k = blockIdx.x * kBlock + threadIdx.x;
  if (k < n) {
    int j = 0;
    while (k >= t.send_end[j]) ++j;
    const int i = k - (j ? t.send_end[j - 1] : 0);
    const int p = idx ? idx[k] : i;
    t.r_stage[j][slot * t.r_stride[j] + i] = p < split ? lo[p] : hi[p];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(t.ticket, 1u);
    if (prev == gridDim.x - 1) {
      *t.ticket = 0;
      __threadfence_system();
      for (int j = 0; j < t.npeers; ++j) st_release_sys(t.r_flag[j], e);
    }
  }
}

// wait + unpack in one launch: a handful of blocks (never enough to fill the device, so the
// peers' puts always find room when ranks share a GPU) poll the flags, then copy grid-stride;
// the block that finishes last advances the plan's counter.
constexpr int kGetBlocks = 48;

__global__ void __launch_bounds__(kBlock)
peer_get_kernel(int n, PeerTab t, double* __restrict__ dst) {
  const unsigned long long e = *(volatile unsigned long long*)t.ctr + 1;
  const long long slot = (long long)(e & 1ull);
  if (threadIdx.x < t.npeers && !wait_flag(t.l_flag[threadIdx.x], e)) atomicExch(t.err, 1);
  __syncthreads();
  const double* __restrict__ src = t.staging + slot * t.stride;
  for (int k = blockIdx.x * kBlock + threadIdx.x; k < n; k += gridDim.x * kBlock) {
    int j = 0;
    while (k >= t.recv_end[j]) ++j;
    dst[t.dst_off[j] + (k - (j ? t.recv_end[j - 1] : 0))] = src[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    // the put kernel's ticket is back at zero by now (its last block reset it before the flags went out)
    const unsigned prev = atomicAdd(t.ticket + 1, 1u);
    if (prev == gridDim.x - 1) {
      t.ticket[1] = 0;
      *(volatile unsigned long long*)t.ctr = e;
    }
  }
}

// put + wait + ordered sum in one block: every rank stores its partials into every other
// rank's window, then sums the S contributions in rank order
__global__ void __launch_bounds__(64)
peer_allreduce_kernel(PeerTab t, double* __restrict__ red, int count) {
  __shared__ double mine[8];
  const unsigned long long e = *(volatile unsigned long long*)t.ctr + 1;
  const long long slot = (long long)(e & 1ull);
  const int tid = threadIdx.x;
  if (tid < count) mine[tid] = red[tid];
  __syncthreads();
  for (int w = tid; w < t.npeers * count; w += blockDim.x) {
    const int j = w / count, i = w - j * count;
    t.r_stage[j][slot * t.r_stride[j] + i] = mine[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < t.npeers) {
    st_release_sys(t.r_flag[tid], e);
    if (!wait_flag(t.l_flag[tid], e)) atomicExch(t.err, 1);
  }
  __syncthreads();
  if (tid < count) {
    double s = 0.0;
    for (int pos = 0; pos <= t.npeers; ++pos) {
      if (pos == t.me_pos) {
        s += mine[tid];
      } else {
        const int j = pos < t.me_pos ? pos : pos - 1;
        s += t.staging[slot * t.stride + 8 * j + tid];
      }
    }
    red[tid] = s;
  }
  if (tid == 0) *(volatile unsigned long long*)t.ctr = e;
}

int peer_put(amgb_ctx* ctx, const PeerPlan& pl, const double* lo, const double* hi, int split) {
  if (pl.tab.npeers == 0) return AMGB_OK;
  // at least one block: it carries the flags even when nothing is sent
  const unsigned grid = pl.send_total > 0 ? (unsigned)div_up(pl.send_total, kBlock) : 1u;
  AMGB_LAUNCH(ctx, F_EXCHANGE, 20.0 * pl.send_total, peer_put_kernel, grid, kBlock, 0, (int)pl.send_total, pl.send_idx, lo,
              hi, split, pl.tab);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

int peer_get(amgb_ctx* ctx, const PeerPlan& pl, double* dst) {
  AMGB_TRY(pl.comm->launch_fence());
  if (pl.tab.npeers == 0) return AMGB_OK;
  int64_t grid = div_up(pl.recv_total, kBlock);
  grid = grid < 1 ? 1 : (grid > kGetBlocks ? kGetBlocks : grid);
  AMGB_LAUNCH(ctx, F_EXCHANGE, 16.0 * pl.recv_total, peer_get_kernel, (unsigned)grid, kBlock, 0, (int)pl.recv_total, pl.tab,
              dst);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

int peer_allreduce(amgb_ctx* ctx, const PeerPlan& pl, double* red, int count) {
  if (count > 8) return set_error(ctx, AMGB_ERR_BAD_ARG, "peer_allreduce: at most 8 values");
  AMGB_TRY(pl.comm->launch_fence());
  if (pl.tab.npeers == 0) return AMGB_OK;
  AMGB_LAUNCH(ctx, F_EXCHANGE, 16.0 * count * (pl.tab.npeers + 1), peer_allreduce_kernel, 1, 64, 0, pl.tab, red, count);
  AMGB_CHECK_LAUNCH(ctx);
  return AMGB_OK;
}

int peer_check(amgb_ctx* ctx, const int* err_word) {
  if (!err_word) return AMGB_OK;
  int h = 0;
  AMGB_CUDA(ctx, cudaMemcpyAsync(&h, err_word, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h) return set_error(ctx, AMGB_ERR_COMM, "a peer-memory exchange timed out (a rank of the communicator is gone)");
  return AMGB_OK;
}

// Window layout of one rank:  [256 B: error word] then per plan [256 B control: counter @0,
// tickets @8 (put) and @12 (get), flag of source rank q @64+8q] [staging: 2 slots x recv_total doubles].
constexpr size_t kCtrlBytes = 256;
static size_t round256(size_t b) { return (b + 255) / 256 * 256; }

int build_peer_plans(amgb_ctx* ctx, amgb_comm* comm, const std::vector<PeerSpec>& specs, std::vector<PeerPlan>& plans,
                     int* window_slot, int** err_word) {
  const int S = comm->size, me = comm->rank, np = (int)specs.size();
  plans.assign(np, PeerPlan());
  *window_slot = -1;
  *err_word = nullptr;
  if (S < 2) return AMGB_OK;
  // my layout, published to everyone: per plan [ctrl_off, stage_off, stride, run offset of source 0..S-1]
  const int cols = 3 + S;
  std::vector<int64_t> mine((size_t)np * cols, 0), all((size_t)np * cols * S, 0);
  size_t off = kCtrlBytes;
  bool fits = S - 1 <= kMaxPeers && 64 + 8 * (size_t)S <= kCtrlBytes;
  for (int i = 0; i < np; ++i) {
    const PeerSpec& sp = specs[i];
    int64_t* row = &mine[(size_t)i * cols];
    int64_t total = 0;
    for (int q = 0; q < S; ++q) {
      row[3 + q] = total;
      if (q != me) total += sp.recv_cnt[q];
    }
    if (total >= (int64_t(1) << 31)) fits = false;
    row[0] = (int64_t)off;
    off += kCtrlBytes;
    row[1] = (int64_t)off;
    row[2] = total;
    off += round256((size_t)total * 2 * sizeof(double));
  }
  int slot = -1;
  AMGB_TRY(comm->window_acquire(ctx, fits ? off : 0, mine.data(), mine.size() * sizeof(int64_t), all.data(), &slot));
  // `fits` is a function of the communicator size and of per-rank totals; make the verdict global
  int64_t good = fits && slot >= 0 ? 1 : 0;
  AMGB_TRY(allreduce_min_i64_host(ctx, comm, &good));
  if (!good) {
    if (slot >= 0) comm->window_release(slot);
    return AMGB_OK;
  }
  const PeerWindow& w = comm->windows[slot];
  for (int i = 0; i < np; ++i) {
    const PeerSpec& sp = specs[i];
    PeerPlan& pl = plans[i];
    PeerTab& t = pl.tab;
    std::memset(&t, 0, sizeof t);
    const int64_t* my = &mine[(size_t)i * cols];
    char* ctrl = w.base + my[0];
    t.staging = (double*)(w.base + my[1]);
    t.stride = my[2];
    t.ctr = (unsigned long long*)ctrl;
    t.ticket = (unsigned*)(ctrl + 8);
    t.err = (int*)w.base;
    int64_t s_end = 0, r_end = 0;
    for (int q = 0; q < S; ++q) {
      if (q == me) continue;
      if (sp.send_cnt[q] == 0 && sp.recv_cnt[q] == 0) continue;
      const int j = t.npeers++;
      if (q < me) t.me_pos = j + 1;
      s_end += sp.send_cnt[q];
      r_end += sp.recv_cnt[q];
      t.send_end[j] = (int)s_end;
      t.recv_end[j] = (int)r_end;
      t.dst_off[j] = sp.dst_off[q];
      const int64_t* theirs = &all[((size_t)q * np + i) * cols];
      char* pb = w.peer_base[q];
      t.r_stage[j] = (double*)(pb + theirs[1]) + theirs[3 + me];
      t.r_stride[j] = theirs[2];
      t.r_flag[j] = (unsigned long long*)(pb + theirs[0] + 64) + me;
      t.l_flag[j] = (const unsigned long long*)(ctrl + 64) + q;
    }
    pl.send_total = s_end;
    pl.recv_total = r_end;
    pl.send_idx = sp.send_idx;
    pl.comm = comm;
    pl.on = true;
  }
  // the kernels of the protocol are loaded now, not at their first launch (see launch_fence)
  cudaFuncAttributes fa;
  AMGB_CUDA(ctx, cudaFuncGetAttributes(&fa, peer_put_kernel));
  AMGB_CUDA(ctx, cudaFuncGetAttributes(&fa, peer_get_kernel));
  AMGB_CUDA(ctx, cudaFuncGetAttributes(&fa, peer_allreduce_kernel));
  *window_slot = slot;
  *err_word = (int*)w.base;
  return AMGB_OK;
}

}  // namespace amgb
