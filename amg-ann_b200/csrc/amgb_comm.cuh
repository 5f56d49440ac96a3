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

struct amgb_comm {
  int rank = 0, size = 1;
  virtual ~amgb_comm() {}
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
