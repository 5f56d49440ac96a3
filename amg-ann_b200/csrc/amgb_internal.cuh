// Internal structures shared by the translation units of libamgb.so.
// Nothing here is part of the ABI (include/amgb.h is).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "amgb.h"

namespace amgb {

// Kernel families for the measurement hooks (amgb_timer_name).
enum Family : int {
  F_SPMV = 0,      // y = A x (+ fused dot) on the system matrix
  F_SMOOTH,        // Jacobi-type half sweeps
  F_RESIDUAL,      // r = f - A u
  F_RESTRICT,      // f_c = P^T r
  F_PROLONG,       // u += P e
  F_VEC,           // PCG vector kernels (axpy, dots, finalize)
  F_COARSE,        // dense coarsest-grid solve
  F_STRENGTH,      // setup: strength of connection
  F_COARSEN,       // setup: PMIS rounds
  F_INTERP,        // setup: interpolation
  F_TRANSPOSE,     // setup: P -> P^T
  F_SPGEMM,        // setup: Galerkin products
  F_SCAN,          // setup: prefix sums / compaction helpers
  F_AUX,           // setup: l1 norms, diag, SELL conversion
  F_POOL,          // pooling
  F_SMOOTH_L0,     // level-0 share of F_SMOOTH etc. (fine-grid roofline numbers)
  F_RESIDUAL_L0,
  F_RESTRICT_L0,
  F_PROLONG_L0,
  F_EXCHANGE,      // row-partitioned path: halo put / wait+unpack / scalar all-reduce kernels
  F_TAIL,          // coarse tail of the cycle in one kernel (amgb_tail.cu)
  F_COUNT
};

struct TimerRec {
  cudaEvent_t a, b;
  int family;
  int level;
  double bytes;
};

constexpr int kTimerLevels = 32;

// Code routes taken on the host side of a context (amgb_route_name): which SELL layout,
// which SpGEMM table tier / overflow stage, captured cycle ... so that parity tests at the
// benchmarked sizes can assert that the routes carrying the benchmark were exercised.
enum Route : int {
  R_SELL_T1_STREAM = 0,  // operator stored with one lane per row (streaming loads)
  R_SELL_T_MULTI,        // operator stored with several lanes per row
  R_SPGEMM_G8_T128,      // A*P-type product, 8 lanes per row, numeric table 128 slots
  R_SPGEMM_G8_T256,
  R_SPGEMM_G8_T512,
  R_SPGEMM_G32,          // long B rows: one warp per row
  R_SPGEMM_SYM_BIG,      // symbolic overflow: big shared-memory table
  R_SPGEMM_SYM_GLOBAL,   // symbolic overflow: global-memory table
  R_SPGEMM_NUM_BIG,
  R_SPGEMM_NUM_GLOBAL,
  R_SPGEMM_ROWREG,       // short product rows accumulated in registers (no hash table)
  R_CYCLE_GRAPH,         // V-cycle replayed from a captured CUDA graph
  R_CYCLE_TAIL_FUSED,    // coarse levels of the cycle run inside one persistent kernel
  R_PCG_DEVICE_LOOP,     // PCG iterations issued without a per-iteration host read
  R_DENSE_STEPWISE,      // coarsest grid factorised by the grid-wide step kernels (n > 96)
  R_SPGEMM_FLAT,         // A*P-type product through the flattened first-stage kernels
  R_COUNT
};

}  // namespace amgb

struct amgb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // body of the PCG while-graph is captured here (created on demand)
  bool own_stream = false;
  int sm_count = 148;
  size_t l2_bytes = 0;
  std::string err;
  int64_t launches = 0;
  bool timers_on = false;
  std::vector<amgb::TimerRec> recs;
  std::vector<cudaEvent_t> free_events;
  double fam_ms[amgb::F_COUNT] = {0};
  int64_t fam_launches[amgb::F_COUNT] = {0};
  double fam_bytes[amgb::F_COUNT] = {0};
  // per (family, level) device time while timers are on; cur_level tags launches
  int cur_level = 0;
  double lvl_ms[amgb::F_COUNT][amgb::kTimerLevels] = {{0}};
  double lvl_bytes[amgb::F_COUNT][amgb::kTimerLevels] = {{0}};
  int64_t lvl_launches[amgb::F_COUNT][amgb::kTimerLevels] = {{0}};
  int64_t routes[amgb::R_COUNT] = {0};
  // Private stream-ordered memory pool.  With the device's default pool, a block freed on
  // one context's stream can be handed to another context with a hidden dependency on the
  // first stream; contexts that run independent systems side by side must not couple.
  cudaMemPool_t pool = nullptr;
  // small pinned staging area for device->host scalars
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
};

namespace amgb {

int set_error(amgb_ctx* ctx, int status, const char* fmt, ...);
int cuda_fail(amgb_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);

#define AMGB_CUDA(ctx, call)                                                  \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return amgb::cuda_fail((ctx), e__, #call, __FILE__, __LINE__); \
  } while (0)

#define AMGB_TRY(expr)            \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != AMGB_OK) return rc__; \
  } while (0)

// Stream-ordered device buffer.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  amgb_ctx* ctx = nullptr;
  bool owns = true;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; ctx = o.ctx; owns = o.owns;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  int alloc(amgb_ctx* c, size_t count) {
    release();
    ctx = c;
    owns = true;
    n = count;
    if (count == 0) return AMGB_OK;
    cudaError_t e = c->pool ? cudaMallocFromPoolAsync((void**)&p, count * sizeof(T), c->pool, c->stream)
                            : cudaMallocAsync((void**)&p, count * sizeof(T), c->stream);
    if (e != cudaSuccess) {
      p = nullptr;
      n = 0;
      (void)cudaGetLastError();
      return set_error(c, e == cudaErrorMemoryAllocation ? AMGB_ERR_OOM : AMGB_ERR_CUDA,
                       "cudaMallocAsync(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    }
    return AMGB_OK;
  }
  int alloc_zero(amgb_ctx* c, size_t count) {
    AMGB_TRY(alloc(c, count));
    if (count) AMGB_CUDA(c, cudaMemsetAsync(p, 0, count * sizeof(T), c->stream));
    return AMGB_OK;
  }
  void wrap(amgb_ctx* c, T* ptr, size_t count) {
    release();
    ctx = c; p = ptr; n = count; owns = false;
  }
  void release() {
    if (p && owns && ctx) cudaFreeAsync(p, ctx->stream);
    p = nullptr;
    n = 0;
  }
};

struct DeviceCsr {
  int64_t n = 0, ncols = 0, nnz = 0;
  DevBuf<int32_t> rp, col;
  DevBuf<double> val;
};

// SELL-32 (sliced ELLPACK, slice height = one warp) operator of the solve phase.
// With T lanes per row a slice (= one warp) holds 32/T rows; entry j of local row
// q lives at 32*slice_ptr[slice] + 32*(j/T) + q*T + j%T, so one warp-wide load is
// always 32 consecutive elements.  Padding entries have val = 0 and a valid col.  Rows/columns are in the C/F-permuted
// ("new") numbering of the level, see DESIGN.md "Solve-phase layout".
struct Sell {
  int64_t n = 0, ncols = 0, nslices = 0, nnz = 0, padded = 0;
  int T = 1;  // lanes cooperating on one row (1,2,...,32); a slice holds 32/T rows
  DevBuf<int32_t> slice_ptr;  // nslices+1, in units of 32 elements
  DevBuf<int32_t> col;
  DevBuf<double> val;
  // row-partitioned path: rows [ilo[b], ihi[b]) of block b (0: rows below bsplit, 1: the
  // rest) reference no halo column, so they can run while the halo is still in flight
  bool has_interior = false;
  int bsplit = 0, ilo[2] = {0, 0}, ihi[2] = {0, 0};
  double csr_bytes() const { return 12.0 * nnz + 4.0 * (n + 1); }
};

struct Level {
  // ---- setup representation (CSR, original numbering of the level) ----
  DeviceCsr A;          // level 0 aliases the user's matrix (no copy)
  DeviceCsr P, R;       // prolongator and its explicit transpose
  DevBuf<uint8_t> mask; // strength mask aligned with A (kept for parity accessors)
  DevBuf<int32_t> cf;   // +1 / -1 / -3 as returned by the coarsening
  DevBuf<int32_t> f2c;  // exclusive scan of the C flags (n+1): coarse index of a C point
  int64_t n_coarse = 0;
  // ---- solve representation (SELL-32, C points first) ----
  DevBuf<int32_t> perm;      // new -> old
  DevBuf<int32_t> inv_perm;  // old -> new
  Sell As, Ps, Rs;
  // F rows x C columns block of As: all a pre-smoothing F half sweep from the zero guess needs
  // (the F columns multiply zeros), ~1/6 of the entries of the F rows on a 27-point operator
  Sell Afc;
  DevBuf<double> inv_relax;  // new order: 1/l1 (type 18) or 1/diag (type 0); 0 = skip row
  // multicolour Gauss-Seidel (relax types 103 / 104 / 106): colour of every point (original numbering);
  // the solve numbering is then by colour, colour c = rows [color_ptr[c], color_ptr[c + 1])
  DevBuf<int32_t> color;
  std::vector<int> color_ptr;
  DevBuf<double> u, f, tmp;  // new order
  // Chebyshev smoother (relax type 16; amgb_cheby.cu): 1/sqrt(diag) in the solve numbering,
  // spectrum estimates of D^-1/2 A D^-1/2, coefficients of p in u += p(A) r, work vectors
  DevBuf<double> cheby_ds, cheby_r, cheby_t[2];
  double cheby_max_eig = 0.0, cheby_min_eig = 0.0, cheby_coefs[5] = {0, 0, 0, 0, 0};
  int cheby_degree = 0;
  // solve-phase sizes: rows relaxed here and vector length (= rows + halo in the
  // row-partitioned path; both equal A.n on a single device)
  int64_t n_solve = 0, n_vec = 0;
};

int64_t div_up(int64_t a, int64_t b);

// Coarse tail of the V-cycle as one kernel (amgb_tail.cu): byte offsets into the packed blob /
// the kernel's shared memory.  POD: passed to the kernel by value.
constexpr int kTailMaxLevels = 10;
struct TailOpDesc {
  int rows = 0, nnz = 0, off_rp = 0, off_col = 0, off_val = 0;
};
struct TailLevelDesc {
  int n = 0, nC = 0;
  TailOpDesc A, P, R;   // P: this level <- next; R: next <- this level
  int off_inv = 0, off_u = 0, off_f = 0, off_t = 0;
};
struct TailDesc {
  int nlev = 0, blob_bytes = 0, dense_n = 0, off_dense = 0;
  double w = 1.0;
  TailLevelDesc lv[kTailMaxLevels];
};

// Launch bookkeeping: counts the launch, optionally brackets it with events.
struct LaunchScope {
  amgb_ctx* ctx;
  bool timed;
  TimerRec rec;
  LaunchScope(amgb_ctx* c, int family, double bytes);
  ~LaunchScope();
};

// Kernel names with template commas must be parenthesised at the call site.
template <class... KArgs, class... Args>
inline void launch_kernel(amgb_ctx* ctx, int family, double bytes, void (*kernel)(KArgs...),
                          dim3 grid, dim3 block, size_t smem, Args... args) {
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;  // empty index range (e.g. a rank without rows)
  LaunchScope scope(ctx, family, bytes);
  kernel<<<grid, block, smem, ctx->stream>>>(args...);
}

#define AMGB_LAUNCH(ctx, family, bytes, kernel, grid, block, smem, ...)                       \
  amgb::launch_kernel((ctx), (family), (double)(bytes), kernel, dim3(grid), dim3(block), (smem), \
                      __VA_ARGS__)

#define AMGB_CHECK_LAUNCH(ctx)                                          \
  do {                                                                  \
    cudaError_t e__ = cudaGetLastError();                               \
    if (e__ != cudaSuccess)                                             \
      return amgb::cuda_fail((ctx), e__, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// ---- shared device-side helpers implemented in amgb_core.cu ----
// exclusive scan of int32 counts (n entries) into out (n+1 entries, out[n] = total).
int exclusive_scan_i32(amgb_ctx* ctx, const int32_t* in, int32_t* out, int64_t n);
// total on host (synchronises the stream)
int read_i32(amgb_ctx* ctx, const int32_t* dptr, int32_t* host);
int read_i64(amgb_ctx* ctx, const int64_t* dptr, int64_t* host);

}  // namespace amgb

struct amgb_matrix {
  amgb_ctx* ctx = nullptr;
  amgb::DeviceCsr A;
};

struct amgb_dist_state;  // amgb_dist.cuh: row-partitioned state (null on a single device)
void amgb_dist_state_destroy(amgb_dist_state* s);

struct amgb_precond {
  amgb_ctx* ctx = nullptr;
  amgb_dist_state* dist = nullptr;
  const amgb_matrix* mat = nullptr;
  amgb_boomeramg_data data;
  double theta_eff = 0.25, mrs_eff = 0.9;
  int relax_down = 18, relax_up = 18, relax_coarse = 9;
  std::vector<amgb::Level> lv;
  amgb::DevBuf<double> dense;  // coarsest operator, LU in place (row-major)
  bool dense_ok = false;
  // coarse tail: levels [tail_from, nl) run in one kernel from a packed blob (-1: none)
  int tail_from = -1;
  amgb::TailDesc tail_desc;
  amgb::DevBuf<unsigned char> tail_blob;
  size_t tail_smem = 0;
  // statistics
  std::vector<int64_t> st_rows, st_nnz, st_nnzP;
  // captured V-cycle (valid for the (z, r) pointer pair below)
  cudaGraphExec_t vcycle_graph = nullptr;
  double* graph_z = nullptr;
  const double* graph_r = nullptr;
  bool use_graph = true;
  bool graph_loop = true;   // PCG iterations as a WHILE node of one graph (needs use_graph)
  bool capturing = false;   // a caller's capture is open: the cycle is recorded inline
  int64_t loop_kernels = 0; // launch accounting of one captured PCG step
  int64_t loop_fam_launches[amgb::F_COUNT] = {0};
  double loop_fam_bytes[amgb::F_COUNT] = {0};
  int64_t graph_kernels = 0;  // kernels inside the captured graph (launch accounting)
  int64_t graph_fam_launches[amgb::F_COUNT] = {0};
  double graph_fam_bytes[amgb::F_COUNT] = {0};
};

namespace amgb {
// Hooks of the row-partitioned path into the setup stages (null on a single device).
// The stages run unchanged on the rank's extended index space (owned points plus ghost
// layers, ordered by global id); the hooks refresh the entries of non-owned points from
// their owners and make the global decisions global.
struct DistHooks {
  virtual ~DistHooks() {}
  virtual int sync_i32(int32_t* per_point) = 0;
  virtual int sync_f64(double* per_point) = 0;
  virtual int allreduce_sum(int64_t* v) = 0;
  int64_t own_begin = 0, own_end = 0;  // owned points in the extended numbering
  const int32_t* gid = nullptr;        // extended -> global id
};

// amgb_setup.cu: the setup stages, shared by the single-device and the partitioned driver
int run_strength(amgb_ctx* ctx, const DeviceCsr& A, double theta, double max_row_sum, uint8_t* mask,
                 int32_t* has_strong, double* diagv);
int coarsen_pmis(amgb_ctx* ctx, const DeviceCsr& A, const uint8_t* mask, const int32_t* has_strong, int32_t* cf,
                 DistHooks* hooks);
// f2c <- exclusive scan of the C flags (n+1 entries); returns the number of C points
int number_coarse_points(amgb_ctx* ctx, int64_t n, const int32_t* cf, int32_t* f2c, int32_t* n_coarse);
// classical modified interpolation for rows [row_begin,row_end) (other rows stay empty);
// col_id[fine] is the coarse column id written for a C point (f2c, or global coarse ids)
int build_interp(amgb_ctx* ctx, const DeviceCsr& A, const uint8_t* mask, const int32_t* cf, const int32_t* col_id,
                 const double* diagv, int64_t row_begin, int64_t row_end, int64_t n_coarse_cols, DeviceCsr& P);
int transpose_csr(amgb_ctx* ctx, const DeviceCsr& P, DeviceCsr& R);
int spgemm(amgb_ctx* ctx, const DeviceCsr& A, const DeviceCsr& B, DeviceCsr& C, bool sorted);
int resolve_options(amgb_precond* P);
int build_levels_from(amgb_precond* P, int level0);
int build_hierarchy(amgb_precond* P);
// amgb_solve.cu
int finish_solve_setup(amgb_precond* P);
// the same for levels [l0, nl) only (replicated coarse levels of the row-partitioned path)
int finish_solve_setup_range(amgb_precond* P, int l0);
// z = M^{-1} r with z, r in the level-0 permuted numbering
int vcycle_apply(amgb_precond* P, double* z_dev, const double* r_dev);
int spmv(amgb_ctx* ctx, const DeviceCsr& A, const double* x, double* y, int family);
// amgb_tail.cu
int tail_plan(const amgb_precond* P, int l0);
int tail_pack(amgb_precond* P);
int tail_cycle(amgb_precond* P, const double* f, double* u);
// amgb_cheby.cu: Chebyshev smoother data of level l (needs the level's CSR operator and perm)
int cheby_setup_level(amgb_precond* P, int l);
void destroy_solve_state(amgb_precond* P);
}  // namespace amgb
