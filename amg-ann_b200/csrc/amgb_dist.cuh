// Row-partitioned (multi-GPU) path: data structures shared by amgb_dist.cu (index
// spaces, exchange plans, ghost rows, the partitioned setup driver) and amgb_solve.cu
// (partitioned V-cycle and PCG).  Design: DESIGN.md "Row partition".
//
// Every rank owns a contiguous range of global rows on every level.  The setup runs the
// SAME kernels as the single-device path on the rank's *extended* index space
//     E = owned points  U  H1 (non-owned columns of owned rows, rows fetched in full)
//                       U  H2 (non-owned columns of those ghost rows; points only)
// numbered by ascending global id, so every ordered sum (row sums, interpolation sums,
// SpGEMM accumulators) sees its terms in the global order and the owned outputs are
// bit-identical to the single-device ones for any number of ranks.  Per-point state of
// non-owned points (measures, C/F markers, coarse ids) is refreshed from the owners after
// every step that changes it.
#pragma once

#include <vector>

#include "amgb_comm.cuh"
#include "amgb_internal.cuh"

namespace amgb {

// Who sends what for per-point arrays over one index space of this rank.
struct HaloPlan {
  int size = 1, rank = 0;
  int64_t n_space = 0, own_begin = 0, n_own = 0;
  std::vector<int64_t> recv_off, recv_cnt;  // per owner: contiguous run of my index space
  std::vector<int64_t> send_off, send_cnt;  // per peer: run of send_idx
  DevBuf<int32_t> send_idx;                 // my indices whose values the peers need
  int64_t send_total = 0, recv_total = 0;
  DevBuf<char> sendbuf;                     // 8 bytes per entry
};

// Rows owned by this rank with GLOBAL column ids: the canonical form a level is handed
// over in (level 0: the user's slab; level l+1: the owned rows of the Galerkin product).
struct OwnedCsr {
  int64_t n_global = 0, g0 = 0;  // global size, first owned global row
  DeviceCsr M;                   // M.n = owned rows, M.col = global ids, ascending
  std::vector<int64_t> starts;   // size+1 range starts of all ranks
};

struct DistLevel {
  OwnedCsr own;              // owned rows, global columns (kept for the accessors)
  int64_t next = 0, o0 = 0;  // extended space size, first owned extended index
  int64_t nloc = 0;
  DevBuf<int32_t> gid;       // extended -> global id (ascending)
  HaloPlan plan;             // per-point state over E
  DeviceCsr Pown;            // owned rows of P (rows over E), GLOBAL coarse column ids
  int64_t nc_global = 0, nc_own = 0, c0 = 0;  // coarse sizes, my first global coarse id
  std::vector<int64_t> cstarts;
  DevBuf<int32_t> tc_gid;    // compact coarse column space of Phat / That: index -> global coarse id
  int64_t nct = 0, tc0 = 0;  // its size, and the compact index of my first owned coarse point
  // solve phase (amgb_solve.cu)
  int64_t nhalo = 0;
  HaloPlan vplan;            // vectors in the solve numbering: [owned C | owned F | halo]
  DevBuf<int32_t> colmap;    // extended -> solve index (-1: not referenced)
};

// ---- peer-memory exchanges of the solve phase (amgb_peer.cu) ----
// One PeerPlan = one recurring exchange among a fixed set of peers: every time, each rank
// PUTS its values straight into the peers' windows (NVLink stores from the pack kernel),
// signals a per-peer flag with the exchange's epoch, and later WAITS for the peers' flags
// and unpacks its own window.  Two staging slots alternate by epoch; a sender cannot be two
// exchanges ahead of a receiver because it has to wait for that receiver's flag of the
// exchange in between (peer sets are symmetric).  Everything is stream work on ctx->stream:
// no host synchronisation, no library call, capturable in a CUDA graph.
constexpr int kMaxPeers = 7;

struct PeerTab {  // kernel parameter
  int npeers, me_pos;                        // me_pos: peers with a smaller rank than mine
  int send_end[kMaxPeers];                   // prefix ends of the peers' segments in my send list
  int recv_end[kMaxPeers];                   // prefix ends of the peers' runs in my staging slot
  long long dst_off[kMaxPeers];              // where peer j's run goes in the destination vector
  double* r_stage[kMaxPeers];                // my run in peer j's staging, slot 0
  long long r_stride[kMaxPeers];             // peer j's slot stride (elements)
  unsigned long long* r_flag[kMaxPeers];     // my flag at peer j
  const unsigned long long* l_flag[kMaxPeers];  // peer j's flag here
  double* staging;                           // my two slots
  long long stride;
  unsigned long long* ctr;                   // completed exchanges of this plan (device)
  unsigned* ticket;
  int* err;                                  // set when a wait timed out
};

struct PeerPlan {
  bool on = false;
  amgb_comm* comm = nullptr;
  PeerTab tab;
  int64_t send_total = 0, recv_total = 0;
  const int32_t* send_idx = nullptr;  // source index of every send-list entry; null: position within the peer's segment
};

// host description of a plan while the windows are laid out (all per rank, in elements)
struct PeerSpec {
  std::vector<int64_t> send_cnt, recv_cnt, dst_off;
  const int32_t* send_idx = nullptr;
};

}  // namespace amgb

struct amgb_dist_matrix {
  amgb_ctx* ctx = nullptr;
  amgb_comm* comm = nullptr;
  amgb::OwnedCsr own;
};

// Distributed state hanging off an amgb_precond (null on a single device).
struct amgb_dist_state {
  amgb_comm* comm = nullptr;
  const amgb_dist_matrix* mat = nullptr;
  std::vector<amgb::DistLevel> dl;
  // levels >= replicated_from hold the WHOLE operator on every rank and run the
  // single-device code without exchanges (coarse-level replication)
  int replicated_from = 1 << 30;
  amgb::DevBuf<double> repl_own, repl_full;  // restriction onto the first replicated level
  // coarsest level, replicated
  int64_t coarse_n = 0;
  std::vector<int64_t> coarse_starts;
  amgb::DevBuf<double> coarse_f, coarse_x;  // full-length staging
  amgb::DevBuf<double> red;                 // PCG reduction staging (device)
  // peer-memory exchanges (window_slot < 0: library exchanges through `comm`)
  int window_slot = -1;
  std::vector<amgb::PeerPlan> vpeer;  // halo plans of the partitioned levels
  amgb::PeerPlan gather_peer;         // all-gather at the replication cut / of the coarsest right-hand side
  amgb::PeerPlan red_peer;            // PCG scalars
  int* peer_err = nullptr;            // device word in my window
  // operators with at least this many rows overlap their halo exchange with the rows that
  // do not need it (amgb_solve.cu launch_sell_halo); smaller ones exchange first
  // (measured on 2 x B200 at 4 M rows per rank the extra launches cost more than the hidden
  // latency, so only very large operators take the overlapped route by default;
  // AMGB_OVERLAP_MIN_ROWS overrides)
  int64_t overlap_min_rows = int64_t(8) << 20;
  ~amgb_dist_state() {
    if (comm && window_slot >= 0) comm->window_release(window_slot);
  }
};

namespace amgb {

// amgb_dist.cu
int plan_sync(amgb_ctx* ctx, amgb_comm* comm, HaloPlan& pl, void* per_point, int elem_bytes);
// values of my owned points, taken from lo (index < split) or hi, sent to the peers' halo
// regions of dst (dst's non-owned runs); lo/hi/dst are over the plan's index space
int plan_sync_split(amgb_ctx* ctx, amgb_comm* comm, HaloPlan& pl, const double* lo, const double* hi, int split,
                    double* dst);
int build_hierarchy_dist(amgb_precond* P);
// amgb_solve.cu
int finish_solve_setup_dist(amgb_precond* P);
// plan for vectors whose non-owned part is the sorted id list `halo_gid` stored at
// [n_own, n_own + n_halo); owners translate a requested id g to own_index[g - g0]
int build_vector_plan(amgb_ctx* ctx, amgb_comm* comm, const int32_t* halo_gid, int64_t n_halo, int64_t n_own,
                      const std::vector<int64_t>& starts, const int32_t* own_index, HaloPlan& pl);
// every rank's owned rows, on every rank (coarsest level)
int allgather_rows(amgb_ctx* ctx, amgb_comm* comm, const OwnedCsr& own, DeviceCsr& full);
int allgather_f64(amgb_ctx* ctx, amgb_comm* comm, const std::vector<int64_t>& starts, const double* mine,
                  double* full);
// amgb_peer.cu
// COLLECTIVE: lays the plans out in a peer window and fills their device tables; leaves
// every plan off (and returns OK) when peer memory is not available on some rank
int build_peer_plans(amgb_ctx* ctx, amgb_comm* comm, const std::vector<PeerSpec>& specs, std::vector<PeerPlan>& plans,
                     int* window_slot, int** err_word);
// put my values (source index p -> p < split ? lo[p] : hi[p]) into the peers' windows and signal
int peer_put(amgb_ctx* ctx, const PeerPlan& pl, const double* lo, const double* hi, int split);
// wait for the peers' puts of the same exchange and unpack them into dst
int peer_get(amgb_ctx* ctx, const PeerPlan& pl, double* dst);
// red[0..count) <- sum over ranks, in rank order (the same bits on every rank); count <= 8
int peer_allreduce(amgb_ctx* ctx, const PeerPlan& pl, double* red, int count);
int peer_check(amgb_ctx* ctx, const int* err_word);

}  // namespace amgb
