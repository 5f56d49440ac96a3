// Communicator used by the row-partitioned path.  Two back ends behind one
// interface:
//   NcclComm   one process per GPU; NCCL over NVLink/NVSwitch (ncclSend/ncclRecv groups
//              for halo and ghost-row exchanges, ncclAllReduce for the PCG scalars).
//              libnccl.so.2 is dlopen'ed on first use, so a process that already loaded
//              torch's bundled NCCL shares it and single-GPU users never need it.
//   LocalComm  all ranks are host threads of one process (one amgb_ctx each, on the same
//              or on different devices); exchanges are device-to-device copies between
//              the ranks' buffers.  This is what the parity tests use on a single GPU.
// All counts and displacements are in BYTES.  Collectives must be called by every rank of
// the communicator in the same order.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "amgb_internal.cuh"

namespace amgb {

// A window of this rank's device memory that every other rank of the communicator can
// store into directly (NVLink peer memory).  The kernels of the solve phase put halo
// values and reduction partials straight into the consumer's window and signal with a
// flag; no library call sits on that path (DESIGN.md "Peer windows").
struct PeerWindow {
  char* base = nullptr;            // my allocation (cudaMalloc: exportable)
  size_t cap = 0;
  uint64_t gen = 0;                // bumped on every re-allocation
  bool in_use = false;
  std::vector<char*> peer_base;    // my mapping of rank q's window (null for myself)
  std::vector<uint64_t> peer_gen;  // generation of that mapping
};

constexpr int kPeerHandleBytes = 64;

}  // namespace amgb

struct amgb_comm {
  int rank = 0, size = 1;
  virtual ~amgb_comm() {}
  // ---- peer windows -------------------------------------------------------------------
  // COLLECTIVE.  A zeroed window of at least `bytes` on every rank (sizes may differ per
  // rank) with the peers' windows mapped; `mine` (extra_bytes, the same size on every rank)
  // is gathered into `all` (size * extra_bytes) on the way, after every rank's window has
  // been zeroed.  *slot < 0 on return means that peer memory is not available on some rank
  // and every rank must use the library exchanges instead.
  int window_acquire(amgb_ctx* ctx, size_t bytes, const void* mine, size_t extra_bytes, void* all, int* slot);
  void window_release(int slot);
  void windows_destroy();
  std::vector<amgb::PeerWindow> windows;
  int window_device = -1;
  std::vector<void*> retired;  // outgrown windows; peers may still hold a mapping until they re-import
  // back-end primitives of the above
  // Called before a kernel that spins on the peers' flags is launched.  Ranks that share a
  // device (threads of one process) meet here on the host, so that every rank has LAUNCHED
  // its puts before anyone spins: a launch may have to load its kernel first (lazy module
  // loading), and that waits for the kernels already running in the context.
  virtual int launch_fence() { return AMGB_OK; }
  virtual bool peer_capable() const { return false; }
  virtual int export_mem(amgb_ctx*, void*, char*) { return AMGB_ERR_UNSUPPORTED; }
  virtual int import_mem(amgb_ctx*, int /*owner*/, int /*owner_device*/, const char*, void**) {
    return AMGB_ERR_UNSUPPORTED;
  }
  virtual void close_mem(void*) {}
  // recv[rdispl[q] .. +rcount[q]) <- rank q's send[sdispl_q[me] .. +scount_q[me]).
  // Device buffers; ordered on ctx->stream; the call returns after the data has landed
  // (LocalComm) or has been enqueued (NcclComm).
  virtual int alltoallv(amgb_ctx* ctx, const void* send, const size_t* scount, const size_t* sdispl, void* recv,
                        const size_t* rcount, const size_t* rdispl) = 0;
  // Host values, blocking: all[q*bytes .. ) <- rank q's `mine`.
  virtual int allgather_host(amgb_ctx* ctx, const void* mine, size_t bytes, void* all) = 0;
  // In place, on ctx->stream: buf[i] <- sum over ranks (rank order, same bits on every rank).
  virtual int allreduce_sum_f64(amgb_ctx* ctx, double* buf_device, int count) = 0;
  // true if alltoallv / allreduce_sum_f64 only enqueue stream work (no host synchronisation),
  // so a sequence of kernels and exchanges can be captured in a CUDA graph
  virtual bool capturable() const { return false; }
};

namespace amgb {

// convenience on top of allgather_host
int allreduce_sum_i64_host(amgb_ctx* ctx, amgb_comm* comm, int64_t* value);
int allreduce_max_i64_host(amgb_ctx* ctx, amgb_comm* comm, int64_t* value);
int allreduce_min_i64_host(amgb_ctx* ctx, amgb_comm* comm, int64_t* value);

}  // namespace amgb
