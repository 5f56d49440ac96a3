// Communicator back ends of the row-partitioned path: NCCL (one process per GPU) and an
// in-process thread group (parity tests on one GPU, or several GPUs driven by one process).
// Interface and semantics: amgb_comm.cuh.
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "amgb_comm.cuh"

namespace amgb {

// ---------------------------------------------------------------------------
// NCCL through dlopen: only the handful of entry points used here.
// ---------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0 };
enum { kNcclChar = 0, kNcclFloat64 = 8 };  // ncclDataType_t: ncclInt8/ncclChar = 0, ncclDouble = 8
enum { kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    bool all = true;
    auto sym = [&](const char* s) {
      void* p = dlsym(api.handle, s);
      if (!p) all = false;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = all;
  });
  return api;
}

#define AMGB_NCCL(ctx, call)                                                                          \
  do {                                                                                                \
    const int r__ = (call);                                                                           \
    if (r__ != kNcclSuccess)                                                                          \
      return set_error((ctx), AMGB_ERR_COMM, "%s failed: %s", #call, nccl_api().GetErrorString(r__)); \
  } while (0)

struct NcclComm : amgb_comm {
  ncclComm_t comm = nullptr;
  DevBuf<char> stage;  // allgather_host staging
  ~NcclComm() override {
    stage.release();
    if (comm) nccl_api().CommDestroy(comm);
  }
  int alltoallv(amgb_ctx* ctx, const void* send, const size_t* scount, const size_t* sdispl, void* recv,
                const size_t* rcount, const size_t* rdispl) override {
    NcclApi& n = nccl_api();
    // (an inconsistent self count is reported AFTER the exchange with the peers has been posted:
    // leaving early would strand them in their receives)
    const bool self_bad = scount[rank] != rcount[rank];
    if (!self_bad && scount[rank])
      AMGB_CUDA(ctx, cudaMemcpyAsync((char*)recv + rdispl[rank], (const char*)send + sdispl[rank], scount[rank],
                                     cudaMemcpyDeviceToDevice, ctx->stream));
    AMGB_NCCL(ctx, n.GroupStart());
    for (int q = 0; q < size; ++q) {
      if (q == rank) continue;
      if (scount[q]) AMGB_NCCL(ctx, n.Send((const char*)send + sdispl[q], scount[q], kNcclChar, q, comm, ctx->stream));
      if (rcount[q]) AMGB_NCCL(ctx, n.Recv((char*)recv + rdispl[q], rcount[q], kNcclChar, q, comm, ctx->stream));
    }
    AMGB_NCCL(ctx, n.GroupEnd());
    if (self_bad) return set_error(ctx, AMGB_ERR_COMM, "alltoallv: self counts differ");
    return AMGB_OK;
  }
  int allgather_host(amgb_ctx* ctx, const void* mine, size_t bytes, void* all) override {
    NcclApi& n = nccl_api();
    const size_t need = bytes * (size_t)(size + 1);
    if (stage.n < need) AMGB_TRY(stage.alloc(ctx, need < 4096 ? 4096 : need));
    char* in = stage.p;
    char* out = stage.p + bytes;
    AMGB_CUDA(ctx, cudaMemcpyAsync(in, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    AMGB_NCCL(ctx, n.AllGather(in, out, bytes, kNcclChar, comm, ctx->stream));
    AMGB_CUDA(ctx, cudaMemcpyAsync(all, out, bytes * size, cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AMGB_OK;
  }
  int allreduce_sum_f64(amgb_ctx* ctx, double* buf, int count) override {
    AMGB_NCCL(ctx, nccl_api().AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, comm, ctx->stream));
    return AMGB_OK;
  }
  bool capturable() const override { return true; }
  // one process per GPU: CUDA IPC handles; the mapping enables peer access over NVLink
  bool peer_capable() const override { return true; }
  int export_mem(amgb_ctx* ctx, void* base, char* handle) override {
    static_assert(sizeof(cudaIpcMemHandle_t) <= kPeerHandleBytes, "handle size");
    cudaIpcMemHandle_t h;
    AMGB_CUDA(ctx, cudaIpcGetMemHandle(&h, base));
    std::memcpy(handle, &h, sizeof h);
    return AMGB_OK;
  }
  int import_mem(amgb_ctx* ctx, int, int, const char* handle, void** out) override {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    AMGB_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return AMGB_OK;
  }
  void close_mem(void* p) override { cudaIpcCloseMemHandle(p); }
};

// ---------------------------------------------------------------------------
// In-process group: ranks are host threads.
// ---------------------------------------------------------------------------
// A rank that left with an error never arrives: the others give up after kBarrierSeconds and
// the group stays broken (every later wait fails at once) instead of hanging the process.
constexpr int kBarrierSeconds = 120;

struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, generation = 0, parties = 1;
  bool broken = false;
  bool wait() {
    std::unique_lock<std::mutex> lk(m);
    if (broken) return false;
    const int gen = generation;
    if (++count == parties) {
      count = 0;
      ++generation;
      cv.notify_all();
      return true;
    }
    if (!cv.wait_for(lk, std::chrono::seconds(kBarrierSeconds), [&] { return gen != generation || broken; }) ||
        broken) {
      broken = true;
      cv.notify_all();
      return false;
    }
    return true;
  }
};

#define AMGB_BARRIER(ctx, g)                                                                              \
  do {                                                                                                    \
    if (!(g)->barrier.wait())                                                                             \
      return set_error((ctx), AMGB_ERR_COMM, "in-process group: a rank did not arrive (it failed earlier)"); \
  } while (0)

struct Slot {
  const void* send = nullptr;
  const size_t* scount = nullptr;
  const size_t* sdispl = nullptr;
  const void* host = nullptr;
  std::vector<double> red;
  int device = 0;  // device of the rank's buffers
};

}  // namespace amgb

struct amgb_local_group {
  int size = 1;
  amgb::Barrier barrier;
  std::vector<amgb::Slot> slots;
};

namespace amgb {

struct LocalComm : amgb_comm {
  amgb_local_group* g = nullptr;
  int alltoallv(amgb_ctx* ctx, const void* send, const size_t* scount, const size_t* sdispl, void* recv,
                const size_t* rcount, const size_t* rdispl) override {
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // my send buffer is complete
    Slot& me = g->slots[rank];
    me.send = send;
    me.scount = scount;
    me.sdispl = sdispl;
    me.device = ctx->device;
    AMGB_BARRIER(ctx, g);
    int rc = AMGB_OK;
    for (int q = 0; q < size && rc == AMGB_OK; ++q) {
      const Slot& s = g->slots[q];
      if (s.scount[rank] != rcount[q]) {
        rc = set_error(ctx, AMGB_ERR_COMM, "alltoallv: rank %d sends %zu bytes, rank %d expects %zu", q,
                       s.scount[rank], rank, rcount[q]);
        break;
      }
      // (ranks on different devices: the sender's buffer comes from ITS context's memory pool, which a
      // plain unified-addressing copy may not reach; the peer copy works with or without peer access)
      if (rcount[q]) {
        const cudaError_t ce =
            s.device == ctx->device
                ? cudaMemcpyAsync((char*)recv + rdispl[q], (const char*)s.send + s.sdispl[rank], rcount[q],
                                  cudaMemcpyDefault, ctx->stream)
                : cudaMemcpyPeerAsync((char*)recv + rdispl[q], ctx->device, (const char*)s.send + s.sdispl[rank],
                                      s.device, rcount[q], ctx->stream);
        if (ce != cudaSuccess) rc = set_error(ctx, AMGB_ERR_CUDA, "alltoallv copy failed: %s", cudaGetErrorString(ce));
      }
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == AMGB_OK)
      rc = set_error(ctx, AMGB_ERR_CUDA, "alltoallv sync failed");
    if (!g->barrier.wait() && rc == AMGB_OK)  // everyone has read my send buffer
      rc = set_error(ctx, AMGB_ERR_COMM, "in-process group: a rank did not arrive (it failed earlier)");
    return rc;
  }
  int allgather_host(amgb_ctx* ctx, const void* mine, size_t bytes, void* all) override {
    g->slots[rank].host = mine;
    AMGB_BARRIER(ctx, g);
    for (int q = 0; q < size; ++q) std::memcpy((char*)all + (size_t)q * bytes, g->slots[q].host, bytes);
    AMGB_BARRIER(ctx, g);
    return AMGB_OK;
  }
  int allreduce_sum_f64(amgb_ctx* ctx, double* buf, int count) override {
    Slot& me = g->slots[rank];
    me.red.resize(count);
    AMGB_CUDA(ctx, cudaMemcpyAsync(me.red.data(), buf, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    AMGB_BARRIER(ctx, g);
    std::vector<double> sum(count, 0.0);
    for (int q = 0; q < size; ++q)
      for (int i = 0; i < count; ++i) sum[i] += g->slots[q].red[i];
    AMGB_BARRIER(ctx, g);  // everyone has read every slot
    AMGB_CUDA(ctx, cudaMemcpyAsync(buf, sum.data(), count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `sum` goes out of scope
    return AMGB_OK;
  }
  int launch_fence() override { return g->barrier.wait() ? AMGB_OK : AMGB_ERR_COMM; }
  // ranks share one address space: the "handle" is the pointer itself
  bool peer_capable() const override { return true; }
  int export_mem(amgb_ctx*, void* base, char* handle) override {
    std::memset(handle, 0, kPeerHandleBytes);
    std::memcpy(handle, &base, sizeof base);
    return AMGB_OK;
  }
  int import_mem(amgb_ctx* ctx, int, int owner_device, const char* handle, void** out) override {
    std::memcpy(out, handle, sizeof(void*));
    if (owner_device != ctx->device) {
      int can = 0;
      AMGB_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, owner_device));
      if (!can) return set_error(ctx, AMGB_ERR_COMM, "device %d cannot access device %d", ctx->device, owner_device);
      const cudaError_t e = cudaDeviceEnablePeerAccess(owner_device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(ctx, e, "enable peer access", __FILE__, __LINE__);
      (void)cudaGetLastError();
    }
    return AMGB_OK;
  }
};

}  // namespace amgb

// ---------------------------------------------------------------------------
// Peer windows (see amgb_comm.cuh).  Grow-only per slot, kept for the life of the
// communicator so that a theta sweep re-uses the same allocation and mappings.
// ---------------------------------------------------------------------------
int amgb_comm::window_acquire(amgb_ctx* ctx, size_t bytes, const void* mine, size_t extra_bytes, void* all,
                              int* slot_out) {
  using namespace amgb;
  *slot_out = -1;
  window_device = ctx->device;
  const char* env = std::getenv("AMGB_PEER");
  const bool want = peer_capable() && size > 1 && !(env && env[0] == '0');
  int slot = 0;
  while (slot < (int)windows.size() && windows[slot].in_use) ++slot;
  if (slot == (int)windows.size()) windows.emplace_back();
  PeerWindow& w = windows[slot];
  w.peer_base.resize(size, nullptr);
  w.peer_gen.resize(size, 0);
  int ok = want ? 1 : 0;
  if (ok && w.cap < bytes) {
    if (w.base) retired.push_back(w.base);
    w.base = nullptr;
    w.cap = 0;
    size_t cap = bytes + bytes / 4;
    if (cap < (size_t(1) << 20)) cap = size_t(1) << 20;
    if (cudaMalloc((void**)&w.base, cap) != cudaSuccess) {
      (void)cudaGetLastError();
      w.base = nullptr;
      ok = 0;
    } else {
      w.cap = cap;
      ++w.gen;
    }
  }
  // payload: [ok, device, generation][handle][caller's table]
  const size_t head = 3 * sizeof(int64_t) + kPeerHandleBytes, per = head + extra_bytes;
  std::vector<char> snd(per, 0), rcv(per * (size_t)size);
  if (ok) {
    if (cudaMemsetAsync(w.base, 0, bytes, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
        export_mem(ctx, w.base, snd.data() + 3 * sizeof(int64_t)) != AMGB_OK) {
      (void)cudaGetLastError();
      ok = 0;
    }
  }
  const int64_t hd[3] = {ok, ctx->device, (int64_t)w.gen};
  std::memcpy(snd.data(), hd, sizeof hd);
  if (extra_bytes) std::memcpy(snd.data() + head, mine, extra_bytes);
  AMGB_TRY(allgather_host(ctx, snd.data(), per, rcv.data()));
  bool all_ok = true, any_new = false;
  for (int q = 0; q < size; ++q) {
    int64_t h[3];
    std::memcpy(h, rcv.data() + per * q, sizeof h);
    if (!h[0]) all_ok = false;
    if (q != rank && (uint64_t)h[2] != w.peer_gen[q]) any_new = true;
    if (extra_bytes) std::memcpy((char*)all + extra_bytes * q, rcv.data() + per * q + head, extra_bytes);
  }
  if (!all_ok) return AMGB_OK;  // the same verdict on every rank
  if (any_new) {
    int64_t good = 1;
    for (int q = 0; q < size; ++q) {
      if (q == rank) continue;
      int64_t h[3];
      std::memcpy(h, rcv.data() + per * q, sizeof h);
      if ((uint64_t)h[2] == w.peer_gen[q]) continue;
      if (w.peer_base[q]) close_mem(w.peer_base[q]);
      w.peer_base[q] = nullptr;
      w.peer_gen[q] = 0;
      void* mapped = nullptr;
      if (import_mem(ctx, q, (int)h[1], rcv.data() + per * q + 3 * sizeof(int64_t), &mapped) != AMGB_OK) {
        (void)cudaGetLastError();
        good = 0;
        continue;
      }
      w.peer_base[q] = (char*)mapped;
      w.peer_gen[q] = (uint64_t)h[2];
    }
    AMGB_TRY(allreduce_min_i64_host(ctx, this, &good));  // every rank maps every window, or nobody uses them
    if (!good) return AMGB_OK;
  }
  w.in_use = true;
  *slot_out = slot;
  return AMGB_OK;
}

void amgb_comm::window_release(int slot) {
  if (slot >= 0 && slot < (int)windows.size()) windows[slot].in_use = false;
}

void amgb_comm::windows_destroy() {
  if (window_device >= 0) cudaSetDevice(window_device);
  for (amgb::PeerWindow& w : windows) {
    for (char* p : w.peer_base)
      if (p) close_mem(p);
    if (w.base) cudaFree(w.base);
  }
  windows.clear();
  for (void* p : retired) cudaFree(p);
  retired.clear();
}

namespace amgb {

static int gather_i64(amgb_ctx* ctx, amgb_comm* comm, int64_t v, std::vector<int64_t>& all) {
  all.resize(comm->size);
  return comm->allgather_host(ctx, &v, sizeof v, all.data());
}

int allreduce_sum_i64_host(amgb_ctx* ctx, amgb_comm* comm, int64_t* value) {
  std::vector<int64_t> all;
  AMGB_TRY(gather_i64(ctx, comm, *value, all));
  int64_t s = 0;
  for (int64_t v : all) s += v;
  *value = s;
  return AMGB_OK;
}

int allreduce_max_i64_host(amgb_ctx* ctx, amgb_comm* comm, int64_t* value) {
  std::vector<int64_t> all;
  AMGB_TRY(gather_i64(ctx, comm, *value, all));
  int64_t s = all[0];
  for (int64_t v : all) s = v > s ? v : s;
  *value = s;
  return AMGB_OK;
}

int allreduce_min_i64_host(amgb_ctx* ctx, amgb_comm* comm, int64_t* value) {
  std::vector<int64_t> all;
  AMGB_TRY(gather_i64(ctx, comm, *value, all));
  int64_t s = all[0];
  for (int64_t v : all) s = v < s ? v : s;
  *value = s;
  return AMGB_OK;
}

}  // namespace amgb

using namespace amgb;

extern "C" {

int amgb_nccl_unique_id(void* out, int capacity) {
  if (!out || capacity < (int)sizeof(ncclUniqueId)) return AMGB_ERR_BAD_ARG;
  NcclApi& n = nccl_api();
  if (!n.ok) return AMGB_ERR_COMM;
  ncclUniqueId id;
  if (n.GetUniqueId(&id) != kNcclSuccess) return AMGB_ERR_COMM;
  std::memcpy(out, &id, sizeof id);
  return AMGB_OK;
}

int amgb_comm_create_nccl(amgb_ctx* ctx, int nranks, int rank, const void* unique_id, amgb_comm** out) {
  if (!ctx || !unique_id || !out || nranks < 1 || rank < 0 || rank >= nranks) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  NcclApi& n = nccl_api();
  if (!n.ok) return set_error(ctx, AMGB_ERR_COMM, "libnccl.so.2 could not be loaded");
  cudaSetDevice(ctx->device);
  NcclComm* c = new NcclComm;
  c->rank = rank;
  c->size = nranks;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof id);
  const int r = n.CommInitRank(&c->comm, nranks, id, rank);
  if (r != kNcclSuccess) {
    c->comm = nullptr;
    delete c;
    return set_error(ctx, AMGB_ERR_COMM, "ncclCommInitRank failed: %s", n.GetErrorString(r));
  }
  *out = c;
  return AMGB_OK;
}

int amgb_local_group_create(int nranks, amgb_local_group** out) {
  if (!out || nranks < 1) return AMGB_ERR_BAD_ARG;
  amgb_local_group* g = new amgb_local_group;
  g->size = nranks;
  g->barrier.parties = nranks;
  g->slots.resize(nranks);
  *out = g;
  return AMGB_OK;
}

int amgb_local_group_abort(amgb_local_group* g) {
  if (!g) return AMGB_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(g->barrier.m);
  g->barrier.broken = true;
  g->barrier.cv.notify_all();
  return AMGB_OK;
}

int amgb_local_group_destroy(amgb_local_group* g) {
  delete g;
  return AMGB_OK;
}

int amgb_comm_create_local(amgb_local_group* g, int rank, amgb_comm** out) {
  if (!g || !out || rank < 0 || rank >= g->size) return AMGB_ERR_BAD_ARG;
  LocalComm* c = new LocalComm;
  c->rank = rank;
  c->size = g->size;
  c->g = g;
  *out = c;
  return AMGB_OK;
}

int amgb_comm_destroy(amgb_comm* c) {
  if (c) c->windows_destroy();
  delete c;
  return AMGB_OK;
}

int amgb_comm_rank(const amgb_comm* c) { return c ? c->rank : -1; }
int amgb_comm_size(const amgb_comm* c) { return c ? c->size : -1; }

}  // extern "C"
