// Communicator back ends of the row-partitioned path: NCCL (one process per GPU) and an
// in-process thread group (parity tests on one GPU, or several GPUs driven by one process).
// Interface and semantics: amgb_comm.cuh.
#include <dlfcn.h>

#include <condition_variable>
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
    if (scount[rank] != rcount[rank]) return set_error(ctx, AMGB_ERR_COMM, "alltoallv: self counts differ");
    if (scount[rank])
      AMGB_CUDA(ctx, cudaMemcpyAsync((char*)recv + rdispl[rank], (const char*)send + sdispl[rank], scount[rank],
                                     cudaMemcpyDeviceToDevice, ctx->stream));
    AMGB_NCCL(ctx, n.GroupStart());
    for (int q = 0; q < size; ++q) {
      if (q == rank) continue;
      if (scount[q]) AMGB_NCCL(ctx, n.Send((const char*)send + sdispl[q], scount[q], kNcclChar, q, comm, ctx->stream));
      if (rcount[q]) AMGB_NCCL(ctx, n.Recv((char*)recv + rdispl[q], rcount[q], kNcclChar, q, comm, ctx->stream));
    }
    AMGB_NCCL(ctx, n.GroupEnd());
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
};

// ---------------------------------------------------------------------------
// In-process group: ranks are host threads.
// ---------------------------------------------------------------------------
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, generation = 0, parties = 1;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const int gen = generation;
    if (++count == parties) {
      count = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != generation; });
    }
  }
};

struct Slot {
  const void* send = nullptr;
  const size_t* scount = nullptr;
  const size_t* sdispl = nullptr;
  const void* host = nullptr;
  std::vector<double> red;
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
    g->barrier.wait();
    int rc = AMGB_OK;
    for (int q = 0; q < size && rc == AMGB_OK; ++q) {
      const Slot& s = g->slots[q];
      if (s.scount[rank] != rcount[q]) {
        rc = set_error(ctx, AMGB_ERR_COMM, "alltoallv: rank %d sends %zu bytes, rank %d expects %zu", q,
                       s.scount[rank], rank, rcount[q]);
        break;
      }
      if (rcount[q] &&
          cudaMemcpyAsync((char*)recv + rdispl[q], (const char*)s.send + s.sdispl[rank], rcount[q],
                          cudaMemcpyDefault, ctx->stream) != cudaSuccess)
        rc = set_error(ctx, AMGB_ERR_CUDA, "alltoallv copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == AMGB_OK)
      rc = set_error(ctx, AMGB_ERR_CUDA, "alltoallv sync failed");
    g->barrier.wait();  // everyone has read my send buffer
    return rc;
  }
  int allgather_host(amgb_ctx*, const void* mine, size_t bytes, void* all) override {
    g->slots[rank].host = mine;
    g->barrier.wait();
    for (int q = 0; q < size; ++q) std::memcpy((char*)all + (size_t)q * bytes, g->slots[q].host, bytes);
    g->barrier.wait();
    return AMGB_OK;
  }
  int allreduce_sum_f64(amgb_ctx* ctx, double* buf, int count) override {
    Slot& me = g->slots[rank];
    me.red.resize(count);
    AMGB_CUDA(ctx, cudaMemcpyAsync(me.red.data(), buf, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->barrier.wait();
    std::vector<double> sum(count, 0.0);
    for (int q = 0; q < size; ++q)
      for (int i = 0; i < count; ++i) sum[i] += g->slots[q].red[i];
    g->barrier.wait();  // everyone has read every slot
    AMGB_CUDA(ctx, cudaMemcpyAsync(buf, sum.data(), count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `sum` goes out of scope
    return AMGB_OK;
  }
};

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
  delete c;
  return AMGB_OK;
}

int amgb_comm_rank(const amgb_comm* c) { return c ? c->rank : -1; }
int amgb_comm_size(const amgb_comm* c) { return c ? c->size : -1; }

}  // extern "C"
