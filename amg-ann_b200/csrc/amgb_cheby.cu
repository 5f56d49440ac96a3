// Chebyshev smoother, hypre relax type 16 (deal.II RelaxationType::Chebyshev; hypre
// par_cheby.c with the PCHYPRE defaults: order 2, spectrum estimate by 10 CG steps,
// fraction 0.3, variant 0, diagonal scaling).  This file holds the setup: per level the
// scaling 1/sqrt(diag), the Lanczos estimate of the extreme eigenvalues of
// D^{-1/2} A D^{-1/2} (hypre_ParCSRMaxEigEstimateCG) and the polynomial coefficients
// (hypre_ParCSRRelax_Cheby_Setup).  The sweep itself runs on the SELL operators in
// amgb_solve.cu (relax_cheby).
//
// The estimate feeds coefficients that every later sweep multiplies with, so it is computed
// in a fixed arithmetic order (sequential row sums with separate multiply and add, inner
// products by a fixed 256-wide tree + left-to-right block sums on the host): the same
// numbers for every launch configuration, and the ones the CPU oracle reproduces bit for bit.
#include <algorithm>
#include <cmath>
#include <vector>

#include "amgb_dist.cuh"
#include "amgb_internal.cuh"

namespace amgb {

namespace {

constexpr int kBlock = 256;

// rows [row0, row0 + n) of a square matrix (row0 > 0: the owned rows of a rank's extended matrix)
__global__ void __launch_bounds__(kBlock)
cheby_scaling_kernel(int64_t n, int64_t row0, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                     const double* __restrict__ val, double* __restrict__ ds) {
  const int64_t i = row0 + (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= row0 + n) return;
  double d = 0.0;
  for (int k = rp[i]; k < rp[i + 1]; ++k)
    if (col[k] == i) d = val[k];
  ds[i] = 1.0 / sqrt(d);
}

// hypre_ParVectorSetRandomValues(r, 1): r_i = 2 * hypre_Rand() - 1 with the multiplicative
// generator 16807 mod 2^31-1 seeded with 1, i.e. value i comes from 16807^(i+1) mod m
// (gid: global index of entry i on the row-partitioned path, so that every rank count draws
// the same vector)
__global__ void __launch_bounds__(kBlock)
cheby_random_kernel(int64_t n, const int32_t* __restrict__ gid, double* __restrict__ r) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const unsigned long long m = 2147483647ull;
  unsigned long long e = (unsigned long long)(gid ? (int64_t)gid[i] : i) + 1ull, base = 16807ull, s = 1ull;
  while (e) {
    if (e & 1ull) s = (s * base) % m;
    base = (base * base) % m;
    e >>= 1;
  }
  r[i] = __dadd_rn(__dmul_rn(2.0, (double)s / 2147483647.0), -1.0);
}

// one partial per block of 256 products: shuffle-down tree inside each warp, the eight warp
// sums left to right
__global__ void __launch_bounds__(kBlock)
cheby_dot_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                 double* __restrict__ partial) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  double v = i < n ? __dmul_rn(x[i], y[i]) : 0.0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
  __shared__ double ws[kBlock / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) t = __dadd_rn(t, ws[w]);
    partial[blockIdx.x] = t;
  }
}

// y = ds .* (A (ds .* x)): sequential row sums, separate multiply and add.  The 32 rows of
// a warp are one contiguous run of entries, staged through shared memory with coalesced
// loads when it fits (cf. strength_kernel).
constexpr int kSpBlock = 128;

__global__ void __launch_bounds__(kSpBlock)
cheby_scaled_spmv_kernel(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ col,
                         const double* __restrict__ val, const double* __restrict__ ds_row,
                         const double* __restrict__ ds, const double* __restrict__ x, double* __restrict__ y,
                         int stage) {
  extern __shared__ double cheby_smem[];
  constexpr int kW = kSpBlock / 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* my_val = cheby_smem + (size_t)w * stage;
  int32_t* my_col = reinterpret_cast<int32_t*>(cheby_smem + (size_t)kW * stage) + (size_t)w * stage;
  const int64_t r0 = ((int64_t)blockIdx.x * kW + w) * 32;
  if (r0 >= n) return;
  const int64_t r1 = r0 + 32 < n ? r0 + 32 : n;
  const int eb = rp[r0], cnt = rp[r1] - eb;
  const bool staged = cnt <= stage;
  if (staged) {
    for (int t = lane; t < cnt; t += 32) {
      my_col[t] = col[eb + t];
      my_val[t] = val[eb + t];
    }
    __syncwarp();
  }
  const int64_t i = r0 + lane;
  if (i >= n) return;
  const int32_t* c = staged ? my_col - eb : col;
  const double* v = staged ? my_val - eb : val;
  double s = 0.0;
  for (int k = rp[i]; k < rp[i + 1]; ++k) {
    const int j = c[k];
    s = __dadd_rn(s, __dmul_rn(v[k], __dmul_rn(ds[j], x[j])));
  }
  y[i] = __dmul_rn(ds_row[i], s);
}

// p = r + beta p  (first step: p = r)
__global__ void __launch_bounds__(kBlock)
cheby_direction_kernel(int64_t n, const double* __restrict__ r, double beta, int first, double* __restrict__ p) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) p[i] = first ? r[i] : __dadd_rn(r[i], __dmul_rn(beta, p[i]));
}

// r = r - alpha s
__global__ void __launch_bounds__(kBlock)
cheby_residual_kernel(int64_t n, double alpha, const double* __restrict__ s, double* __restrict__ r) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) r[i] = __dadd_rn(r[i], -__dmul_rn(alpha, s[i]));
}

__global__ void __launch_bounds__(kBlock)
cheby_gather_kernel(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ in,
                    double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) out[i] = in[perm[i]];
}

// ---- host side: EISPACK tql1 (what hypre_LINPACKcgtql1 is) and the coefficient formulas ----
double pythag(double a, double b) {
  const double p = std::max(std::fabs(a), std::fabs(b));
  if (p == 0.0) return 0.0;
  const double q = std::min(std::fabs(a), std::fabs(b)) / p;
  double r = q * q, pp = p;
  for (;;) {
    const double t = 4.0 + r;
    if (t == 4.0) break;
    const double sq = r / t, u = 1.0 + 2.0 * sq;
    pp = u * pp;
    r = (sq / u) * (sq / u) * r;
  }
  return pp;
}

// eigenvalues of the symmetric tridiagonal matrix (d, e[1..n)), ascending in d
int tql1(int n, double* d, double* e) {
  if (n <= 1) return 0;
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  double f = 0.0, tst1 = 0.0;
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int j = 0;
    const double h0 = std::fabs(d[l]) + std::fabs(e[l]);
    if (tst1 < h0) tst1 = h0;
    int m = l;
    for (; m < n; ++m)
      if (tst1 + std::fabs(e[m]) == tst1) break;
    if (m != l) {
      double tst2;
      do {
        if (j == 30) return l + 1;
        ++j;
        const int l1 = l + 1, l2 = l1 + 1;
        double g = d[l];
        double p = (d[l1] - g) / (2.0 * e[l]);
        double r = pythag(p, 1.0);
        const double sr = p >= 0.0 ? std::fabs(r) : -std::fabs(r);
        d[l] = e[l] / (p + sr);
        d[l1] = e[l] * (p + sr);
        const double dl1 = d[l1];
        double h = g - d[l];
        for (int i = l2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
        const double el1 = e[l1];
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = pythag(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
        tst2 = tst1 + std::fabs(e[l]);
      } while (tst2 > tst1);
    }
    const double p = d[l] + f;
    int i = l;
    for (; i >= 1; --i) {
      if (p >= d[i - 1]) break;
      d[i] = d[i - 1];
    }
    d[i] = p;
  }
  return 0;
}

// hypre_ParCSRRelax_Cheby_Setup, variant 0
void cheby_coefficients(double max_eig, double min_eig, double fraction, int order, double* coefs, int* degree) {
  order = std::min(std::max(order, 1), 4);
  const int k = order - 1;
  const double upper = max_eig * 1.1;
  const double lower = (upper - min_eig) * fraction + min_eig;
  const double theta = (upper + lower) / 2, delta = (upper - lower) / 2;
  double den;
  switch (k) {
    case 0:
      coefs[0] = 1.0 / theta;
      break;
    case 1:
      den = 2 * theta * theta - delta * delta;
      coefs[0] = 4 * theta / den;
      coefs[1] = -2 / den;
      break;
    case 2:
      den = 4 * (theta * theta * theta) - 3 * (delta * delta) * theta;
      coefs[0] = (12 * (theta * theta) - 3 * (delta * delta)) / den;
      coefs[1] = -12 * theta / den;
      coefs[2] = 4 / den;
      break;
    default:
      den = std::pow(delta, 4) - 8 * (delta * delta) * (theta * theta) + 8 * std::pow(theta, 4);
      coefs[0] = (32 * std::pow(theta, 3) - 16 * (delta * delta) * theta) / den;
      coefs[1] = (8 * (delta * delta) - 48 * (theta * theta)) / den;
      coefs[2] = 32 * theta / den;
      coefs[3] = -8 / den;
      break;
  }
  *degree = k;
}

}  // namespace

// Level l of P: scaling (solve numbering), spectrum estimate, coefficients.  On the
// row-partitioned path the level's matrix is the rank's extended matrix (owned rows
// [o0, o0 + nloc) with columns over the extended space): the vectors live on the extended
// space, the direction's ghost entries are refreshed before every product, inner products
// are the owned parts summed over the ranks in rank order.
int cheby_setup_level(amgb_precond* P, int l) {
  amgb_ctx* ctx = P->ctx;
  Level& L = P->lv[l];
  const DeviceCsr& A = L.A;
  if (!A.rp.p) return set_error(ctx, AMGB_ERR_UNSUPPORTED, "Chebyshev smoother: level %d has no CSR operator", l);
  amgb_dist_state* dist = P->dist && l < P->dist->replicated_from ? P->dist : nullptr;
  DistLevel* D = dist ? &dist->dl[l] : nullptr;
  amgb_comm* comm = dist ? dist->comm : nullptr;
  const int64_t o0 = D ? D->o0 : 0, n = D ? D->nloc : A.n, next = D ? D->next : A.n;
  const int64_t n_global = D ? D->own.n_global : A.n;
  const int64_t own_nnz = D ? D->own.M.nnz : A.nnz;
  const unsigned vgrid = (unsigned)div_up(n, kBlock);
  const int64_t nblocks = div_up(n, kBlock);
  DevBuf<double> ds, r, p, s, partial;
  AMGB_TRY(ds.alloc_zero(ctx, next));
  AMGB_TRY(r.alloc(ctx, next));
  AMGB_TRY(p.alloc_zero(ctx, next));
  AMGB_TRY(s.alloc(ctx, next));
  AMGB_TRY(partial.alloc(ctx, nblocks));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * own_nnz + 8.0 * n, cheby_scaling_kernel, vgrid, kBlock, 0, n, o0, A.rp.p, A.col.p,
              A.val.p, ds.p);
  AMGB_LAUNCH(ctx, F_AUX, 8.0 * next, cheby_random_kernel, (unsigned)div_up(next, kBlock), kBlock, 0, next,
              D ? (const int32_t*)D->gid.p : (const int32_t*)nullptr, r.p);
  AMGB_CHECK_LAUNCH(ctx);
  if (D) AMGB_TRY(plan_sync(ctx, comm, D->plan, ds.p, 8));
  std::vector<double> hp((size_t)nblocks);
  auto dot = [&](const double* x, const double* y, double* out) -> int {
    AMGB_LAUNCH(ctx, F_AUX, 16.0 * n, cheby_dot_kernel, (unsigned)nblocks, kBlock, 0, n, x + o0, y + o0, partial.p);
    AMGB_CHECK_LAUNCH(ctx);
    if (nblocks)
      AMGB_CUDA(ctx, cudaMemcpyAsync(hp.data(), partial.p, (size_t)nblocks * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double t = 0.0;
    for (double v : hp) t = t + v;
    if (comm) {
      std::vector<double> all((size_t)comm->size);
      AMGB_TRY(comm->allgather_host(ctx, &t, sizeof t, all.data()));
      t = 0.0;
      for (double v : all) t = t + v;
    }
    *out = t;
    return AMGB_OK;
  };
  const double avg = n > 0 ? double(own_nnz) / double(n) : 1.0;
  int stage = (int)(40.0 * avg) / 128 * 128 + 128;
  stage = std::min(std::max(stage, 256), 2048);
  const size_t smem = (size_t)(kSpBlock / 32) * stage * 12;
  if (smem > 48 * 1024)
    AMGB_CUDA(ctx, cudaFuncSetAttribute(cheby_scaled_spmv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        4 * 2048 * 12));
  int max_iter = 10;  // cheby_eig_est
  if (n_global < max_iter) max_iter = (int)n_global;
  std::vector<double> tri(max_iter + 2, 0.0), off(max_iter + 2, 0.0);
  double gamma = 1.0;
  int it = 0;
  while (it < max_iter) {
    const double gamma_old = gamma;
    AMGB_TRY(dot(r.p, r.p, &gamma));
    if (!(gamma > 0.0)) break;
    const double beta = it == 0 ? 1.0 : gamma / gamma_old;
    AMGB_LAUNCH(ctx, F_AUX, 24.0 * n, cheby_direction_kernel, vgrid, kBlock, 0, n, (const double*)(r.p + o0), beta,
                it == 0 ? 1 : 0, p.p + o0);
    AMGB_CHECK_LAUNCH(ctx);
    if (D) AMGB_TRY(plan_sync(ctx, comm, D->plan, p.p, 8));
    AMGB_LAUNCH(ctx, F_AUX, 12.0 * own_nnz + 32.0 * n, cheby_scaled_spmv_kernel, (unsigned)div_up(n, kSpBlock), kSpBlock,
                smem, n, A.rp.p + o0, A.col.p, A.val.p, (const double*)(ds.p + o0), (const double*)ds.p,
                (const double*)p.p, s.p + o0, stage);
    AMGB_CHECK_LAUNCH(ctx);
    double sdotp = 0.0;
    AMGB_TRY(dot(s.p, p.p, &sdotp));
    if (!(sdotp > 0.0)) break;
    const double alpha = gamma / sdotp;
    const double alphainv = 1.0 / alpha;
    tri[it + 1] = alphainv;
    tri[it] *= beta;
    tri[it] += alphainv;
    off[it + 1] = alphainv;
    off[it] *= std::sqrt(beta);
    AMGB_LAUNCH(ctx, F_AUX, 24.0 * n, cheby_residual_kernel, vgrid, kBlock, 0, n, alpha, (const double*)(s.p + o0),
                r.p + o0);
    AMGB_CHECK_LAUNCH(ctx);
    ++it;
  }
  if (it == 0) {
    L.cheby_max_eig = L.cheby_min_eig = 1.0;
  } else {
    tql1(it, tri.data(), off.data());
    L.cheby_max_eig = tri[it - 1];
    L.cheby_min_eig = tri[0];
  }
  cheby_coefficients(L.cheby_max_eig, L.cheby_min_eig, 0.3, 2, L.cheby_coefs, &L.cheby_degree);
  // the sweeps run in the C/F-permuted numbering of the level (owned rows; the work vectors
  // that are gathered from carry the halo as well)
  AMGB_TRY(L.cheby_ds.alloc(ctx, n));
  AMGB_LAUNCH(ctx, F_AUX, 20.0 * n, cheby_gather_kernel, vgrid, kBlock, 0, n, (const int32_t*)L.perm.p,
              (const double*)(ds.p + o0), L.cheby_ds.p);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_TRY(L.cheby_r.alloc(ctx, n));
  AMGB_TRY(L.cheby_t[0].alloc(ctx, L.n_vec));
  if (L.cheby_degree > 1) AMGB_TRY(L.cheby_t[1].alloc(ctx, L.n_vec));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the temporaries go out of scope
  return AMGB_OK;
}

}  // namespace amgb

using namespace amgb;

extern "C" int amgb_precond_level_cheby(const amgb_precond* P, int32_t level, double* max_eig, double* min_eig,
                                        double* coefs, int32_t* n_coefs) {
  if (!P) return AMGB_ERR_BAD_ARG;
  if (level < 0 || level >= (int)P->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = P->lv[level];
  if (!L.cheby_ds.p) return set_error(P->ctx, AMGB_ERR_RANGE, "level %d has no Chebyshev smoother", level);
  if (max_eig) *max_eig = L.cheby_max_eig;
  if (min_eig) *min_eig = L.cheby_min_eig;
  if (coefs)
    for (int i = 0; i <= L.cheby_degree; ++i) coefs[i] = L.cheby_coefs[i];
  if (n_coefs) *n_coefs = L.cheby_degree + 1;
  return AMGB_OK;
}
