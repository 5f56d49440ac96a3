// On-device assembly of the Q1 diffusion system (the step before
// the path).  Same discretisation and conventions as the host generator
// csrc/gen_q1.cpp (ref testcase2-diffusion-structured/src/main.cpp:255-320 assembly,
// :101-113 piecewise-constant mu, :312-318 Dirichlet rows, :239-249 full pattern), and
// the same arithmetic bit for bit: one thread per row walks the <= 8 adjacent cells in
// the host's (z,y,x) order with separately rounded multiplies and adds; everything that
// needs libm (10^eps, the 1-D factors of the manufactured solution, the pattern digit of a
// cell) is tabulated on the host in O(m) and uploaded.  The matrix is born in HBM: no
// 2.6 GB host->device copy per matrix at config-2 size, and a z-slab of the 100 M-DoF
// system of config 5 is assembled in milliseconds instead of ~16 s on the host.
#include <cmath>
#include <vector>

#include "amgb_dist.cuh"
#include "amgb_internal.cuh"

namespace amgb {

constexpr int kAsmBlock = 128;

struct AsmParams {
  int m, mode, ps;
  long long n_epsv;
  double h, hw, dirichlet_diag;  // hw = ((h*h)*h)*w_q
  const double* Kq;       // [8][8]
  const double* phi;      // [8][8]  phi[q][l]
  const double* diffv;    // [n_epsv]
  const int32_t* pdigit;  // [m]   pattern digit of a cell coordinate
  const double* v0q;      // [2m]  1-D factor at the quadrature coordinates
  const double* v2q;      // [2m]  its second derivative
  const double* v0n;      // [m+1] 1-D factor at the nodes
};

__device__ __forceinline__ int valid_per_dim(int m, int i) { return (i == 0 || i == m) ? 2 : 3; }

__global__ void __launch_bounds__(kAsmBlock)
asm_row_len_kernel(int m, long long row_begin, long long nloc, int32_t* __restrict__ len) {
  const long long r = (long long)blockIdx.x * kAsmBlock + threadIdx.x;
  if (r >= nloc) return;
  const long long N = m + 1, gr = row_begin + r;
  const int ix = (int)(gr % N), iy = (int)((gr / N) % N), iz = (int)(gr / (N * N));
  len[r] = valid_per_dim(m, ix) * valid_per_dim(m, iy) * valid_per_dim(m, iz);
}

__device__ __forceinline__ double cell_mu(const AsmParams& p, int cx, int cy, int cz) {
  long long ind = 0, pw = 1;
  const int c[3] = {cx, cy, cz};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < p.mode) {
      ind += (long long)p.pdigit[c[i]] * pw;
      pw *= p.ps;
    }
  }
  if (ind < 0) ind = 0;
  if (ind >= p.n_epsv) ind = p.n_epsv - 1;
  return p.diffv[ind];
}

__global__ void __launch_bounds__(kAsmBlock)
asm_poisson_kernel(AsmParams p, long long row_begin, long long nloc, const int32_t* __restrict__ rowptr,
                   int32_t* __restrict__ col, double* __restrict__ val, double* __restrict__ rhs,
                   double* __restrict__ x0) {
  const long long r = (long long)blockIdx.x * kAsmBlock + threadIdx.x;
  if (r >= nloc) return;
  const int m = p.m;
  const long long N = m + 1, gr = row_begin + r;
  const int ix = (int)(gr % N), iy = (int)((gr / N) % N), iz = (int)(gr / (N * N));
  const bool bnd = ix == 0 || iy == 0 || iz == 0 || ix == m || iy == m || iz == m;
  double acc[27];
#pragma unroll
  for (int s = 0; s < 27; ++s) acc[s] = 0.0;
  double b = 0.0;
#pragma unroll
  for (int dz = -1; dz <= 0; ++dz)
#pragma unroll
    for (int dy = -1; dy <= 0; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 0; ++dx) {
        const int cx = ix + dx, cy = iy + dy, cz = iz + dz;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= m || cy >= m || cz >= m) continue;
        const double mu = cell_mu(p, cx, cy, cz);
        const double muh = __dmul_rn(mu, p.h);
        const int li = (-dx) + 2 * (-dy) + 4 * (-dz);
#pragma unroll
        for (int lj = 0; lj < 8; ++lj) {
          const int ox = (lj & 1) + dx, oy = ((lj >> 1) & 1) + dy, oz = ((lj >> 2) & 1) + dz;
          const int slot = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
          acc[slot] = __dadd_rn(acc[slot], __dmul_rn(muh, p.Kq[li * 8 + lj]));
        }
        if (rhs && !bnd) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            // 1-D factors at the quadrature point of this cell, per direction
            const int qi[3] = {2 * cx + (q & 1), 2 * cy + ((q >> 1) & 1), 2 * cz + ((q >> 2) & 1)};
            double f = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              double dd = 1.0;
#pragma unroll
              for (int j = 0; j < 3; ++j) dd = __dmul_rn(dd, i == j ? p.v2q[qi[j]] : p.v0q[qi[j]]);
              f = __dadd_rn(f, dd);
            }
            const double t = __dmul_rn(__dmul_rn(__dmul_rn(mu, p.phi[q * 8 + li]), -f), p.hw);
            b = __dadd_rn(b, t);
          }
        }
      }
  int k = rowptr[r];
#pragma unroll
  for (int oz = -1; oz <= 1; ++oz) {
    if (iz + oz < 0 || iz + oz > m) continue;
#pragma unroll
    for (int oy = -1; oy <= 1; ++oy) {
      if (iy + oy < 0 || iy + oy > m) continue;
#pragma unroll
      for (int ox = -1; ox <= 1; ++ox) {
        if (ix + ox < 0 || ix + ox > m) continue;
        const bool diag = ox == 0 && oy == 0 && oz == 0;
        col[k] = (int32_t)((ix + ox) + N * ((iy + oy) + N * (long long)(iz + oz)));
        val[k] = bnd ? (diag ? p.dirichlet_diag : 0.0) : acc[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)];
        ++k;
      }
    }
  }
  double bv = 0.0;
  if (bnd) bv = __dmul_rn(__dmul_rn(p.v0n[ix], p.v0n[iy]), p.v0n[iz]);
  if (rhs) rhs[r] = bnd ? __dmul_rn(bv, p.dirichlet_diag) : b;
  if (x0) x0[r] = bnd ? bv : 0.0;
}

// ---- host tables (the libm part; identical expressions to gen_q1.cpp) ----
struct AsmTables {
  double Kq[64], phi[64];
  std::vector<double> diffv, v0q, v2q, v0n;
  std::vector<int32_t> pdigit;
  double h = 0, hw = 0, dirichlet_diag = 0;
};

static double sol1d(int id, double t, double f, int der) {  // ref common/cube_solution.h:6-28
  if (id == 0) {
    if (der == 0) return std::sin(f * t);
    if (der == 1) return f * std::cos(f * t);
    return -f * f * std::sin(f * t);
  }
  if (der == 0) return std::cos(f * t);
  if (der == 1) return -f * std::sin(f * t);
  return -f * f * std::cos(f * t);
}

static void make_tables(int m, int ps, int mode, const double* epsv, long long n_epsv, AsmTables& T) {
  const double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  double grad[8][8][3], w[8];
  for (int q = 0; q < 8; ++q) {
    const double xq[3] = {g[q & 1], g[(q >> 1) & 1], g[(q >> 2) & 1]};
    w[q] = 0.125;
    for (int l = 0; l < 8; ++l) {
      const int s[3] = {l & 1, (l >> 1) & 1, (l >> 2) & 1};
      double f[3], df[3];
      for (int d = 0; d < 3; ++d) {
        f[d] = s[d] ? xq[d] : 1.0 - xq[d];
        df[d] = s[d] ? 1.0 : -1.0;
      }
      T.phi[q * 8 + l] = f[0] * f[1] * f[2];
      grad[q][l][0] = df[0] * f[1] * f[2];
      grad[q][l][1] = f[0] * df[1] * f[2];
      grad[q][l][2] = f[0] * f[1] * df[2];
    }
  }
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double s = 0.0;
      for (int q = 0; q < 8; ++q)
        s += (grad[q][i][0] * grad[q][j][0] + grad[q][i][1] * grad[q][j][1] + grad[q][i][2] * grad[q][j][2]) * w[q];
      T.Kq[i * 8 + j] = s;
    }
  T.diffv.resize(n_epsv);
  for (long long i = 0; i < n_epsv; ++i) T.diffv[i] = std::pow(10.0, epsv[i]);
  const double h = 2.0 / m;
  T.h = h;
  T.hw = h * h * h * w[0];
  const int sol_id = 3 * (1 - (ps % 2));               // ref t2 main.cpp:70-72
  const double freq = M_PI / (2.0 / double(ps));       // ref t2 main.cpp:80-81
  const double hp = 2.0 / double(ps);
  T.pdigit.resize(m);
  T.v0q.resize(2 * (size_t)m);
  T.v2q.resize(2 * (size_t)m);
  T.v0n.resize((size_t)m + 1);
  for (long long c = 0; c < m; ++c) {
    const double centre = -1.0 + (c + 0.5) * h;
    T.pdigit[c] = (int32_t)(long)std::trunc((centre + 1.0) / (hp + 1e-15));  // ref t2 main.cpp:101-113
    for (int s = 0; s < 2; ++s) {
      const double pq = -1.0 + (c + g[s]) * h;
      T.v0q[2 * c + s] = sol1d(sol_id, pq, freq, 0);
      T.v2q[2 * c + s] = sol1d(sol_id, pq, freq, 2);
    }
  }
  for (long long i = 0; i <= m; ++i) T.v0n[i] = sol1d(sol_id, -1.0 + i * h, freq, 0);
  long ind = 0, pw = 1;
  for (int i = 0; i < mode; ++i) {
    ind += (long)T.pdigit[0] * pw;
    pw *= ps;
  }
  if (ind < 0) ind = 0;
  if (ind >= n_epsv) ind = (long)n_epsv - 1;
  T.dirichlet_diag = std::fabs(T.diffv[ind] * h * T.Kq[0]);
}

// Assembles rows [row_begin,row_end) into M (rowptr local, columns GLOBAL ids).
static int assemble_rows(amgb_ctx* ctx, int m, int ps, int mode, const double* epsv, int64_t n_epsv, int64_t row_begin,
                         int64_t row_end, DeviceCsr& M, double* rhs_device, double* x0_device) {
  if (m < 1 || ps < 1 || mode < 1 || mode > 3 || !epsv) return AMGB_ERR_BAD_ARG;
  int64_t want = 1;
  for (int i = 0; i < mode; ++i) want *= ps;
  const int64_t N = (int64_t)m + 1, n = N * N * N;
  if (n_epsv != want || row_begin < 0 || row_end > n || row_begin > row_end) return AMGB_ERR_BAD_ARG;
  if (n >= (int64_t(1) << 31)) return set_error(ctx, AMGB_ERR_RANGE, "n=%lld: global ids are 32-bit", (long long)n);
  const int64_t nloc = row_end - row_begin;
  AsmTables T;
  make_tables(m, ps, mode, epsv, n_epsv, T);
  DevBuf<double> tab;
  DevBuf<int32_t> pd, len;
  const size_t nd = 64 + 64 + (size_t)n_epsv + 2 * (size_t)m + 2 * (size_t)m + (size_t)m + 1;
  AMGB_TRY(tab.alloc(ctx, nd));
  AMGB_TRY(pd.alloc(ctx, m));
  std::vector<double> flat;
  flat.reserve(nd);
  flat.insert(flat.end(), T.Kq, T.Kq + 64);
  flat.insert(flat.end(), T.phi, T.phi + 64);
  flat.insert(flat.end(), T.diffv.begin(), T.diffv.end());
  flat.insert(flat.end(), T.v0q.begin(), T.v0q.end());
  flat.insert(flat.end(), T.v2q.begin(), T.v2q.end());
  flat.insert(flat.end(), T.v0n.begin(), T.v0n.end());
  AMGB_CUDA(ctx, cudaMemcpyAsync(tab.p, flat.data(), nd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(pd.p, T.pdigit.data(), (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
  AsmParams p;
  p.m = m;
  p.mode = mode;
  p.ps = ps;
  p.n_epsv = n_epsv;
  p.h = T.h;
  p.hw = T.hw;
  p.dirichlet_diag = T.dirichlet_diag;
  p.Kq = tab.p;
  p.phi = tab.p + 64;
  p.diffv = tab.p + 128;
  p.v0q = p.diffv + n_epsv;
  p.v2q = p.v0q + 2 * (size_t)m;
  p.v0n = p.v2q + 2 * (size_t)m;
  p.pdigit = pd.p;
  M.n = nloc;
  M.ncols = n;
  AMGB_TRY(len.alloc(ctx, nloc));
  AMGB_TRY(M.rp.alloc(ctx, nloc + 1));
  const unsigned grid = (unsigned)div_up(nloc, kAsmBlock);
  AMGB_LAUNCH(ctx, F_AUX, 4.0 * nloc, asm_row_len_kernel, grid, kAsmBlock, 0, m, (long long)row_begin, (long long)nloc,
              len.p);
  AMGB_TRY(exclusive_scan_i32(ctx, len.p, M.rp.p, nloc));
  // the local nnz of a slab must fit the 32-bit row pointers (checked against the closed form)
  {
    double est = 27.0 * (double)nloc;
    if (est >= 2147483648.0 * 1.05)
      return set_error(ctx, AMGB_ERR_RANGE, "slab of %lld rows has too many entries for 32-bit row pointers",
                       (long long)nloc);
  }
  int32_t nnz = 0;
  AMGB_TRY(read_i32(ctx, M.rp.p + nloc, &nnz));
  if (nnz < 0) return set_error(ctx, AMGB_ERR_RANGE, "slab nnz overflows 32-bit row pointers");
  M.nnz = nnz;
  AMGB_TRY(M.col.alloc(ctx, nnz));
  AMGB_TRY(M.val.alloc(ctx, nnz));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * nnz + 16.0 * nloc, asm_poisson_kernel, grid, kAsmBlock, 0, p, (long long)row_begin,
              (long long)nloc, (const int32_t*)M.rp.p, M.col.p, M.val.p, rhs_device, x0_device);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the tables go out of scope
  return AMGB_OK;
}

// ---------------------------------------------------------------------------
// Q1 vector elasticity (3 DoFs per node, interleaved; ref testcase3-elasticity-structured/src/
// main.cpp:320-342 cell matrix, :88-99 piecewise-constant Lame parameters, :264-268 constraints
// condensed out).  One thread per NODE accumulates the 3 x 27 x 3 block row of its three DoFs
// over the <= 8 adjacent cells in the host generator's order with the host's roundings
// (csrc/gen_q1.cpp amgb_gen_elasticity_q1), so the matrix and the initial guess are the host's
// bit for bit.  The body force needs sines and cosines of non-separable arguments at the
// quadrature points: it is evaluated with the device's libm, so the right-hand side agrees
// with the host's to rounding, not to the bit.
// ---------------------------------------------------------------------------
struct ElasParams {
  int m, mode, ps;
  long long n_young;
  double h, beta, pi;
  const double* G1;       // [24][24]
  const double* G2;       // [24][24]
  const double* muv;      // [n_young]  e_min * young / (1 + nu)
  const int32_t* pdigit;  // [m]
  const double* sn;       // [m+1] sin(pi * x_i) at the nodes
  const double* phi;      // [8][8]
};

__device__ __forceinline__ int elas_cnt(int m, int i) { return 3 - (i == 1) - (i == m - 1); }

__global__ void __launch_bounds__(kAsmBlock)
elas_row_len_kernel(int m, long long nn, int32_t* __restrict__ len) {
  const long long nd = (long long)blockIdx.x * kAsmBlock + threadIdx.x;
  if (nd >= nn) return;
  const long long N = m + 1;
  const int ix = (int)(nd % N), iy = (int)((nd / N) % N), iz = (int)(nd / (N * N));
  const bool bnd = ix == 0 || iy == 0 || iz == 0 || ix == m || iy == m || iz == m;
  int l = 1;
  if (!bnd) {
    l = 3 * elas_cnt(m, ix) * elas_cnt(m, iy) * elas_cnt(m, iz);
    if (m == 2) l = 3;
  }
  len[3 * nd] = len[3 * nd + 1] = len[3 * nd + 2] = l;
}

__global__ void __launch_bounds__(kAsmBlock)
asm_elasticity_kernel(ElasParams p, long long nn, const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                      double* __restrict__ val, double* __restrict__ rhs, double* __restrict__ x0) {
  const long long nd = (long long)blockIdx.x * kAsmBlock + threadIdx.x;
  if (nd >= nn) return;
  const int m = p.m;
  const long long N = m + 1;
  const int ix = (int)(nd % N), iy = (int)((nd / N) % N), iz = (int)(nd / (N * N));
  const bool bnd = ix == 0 || iy == 0 || iz == 0 || ix == m || iy == m || iz == m;
  double acc[3][27][3];
  double b[3] = {0.0, 0.0, 0.0};
  for (int c = 0; c < 3; ++c)
    for (int s = 0; s < 27; ++s) acc[c][s][0] = acc[c][s][1] = acc[c][s][2] = 0.0;
  const double h = p.h;
  for (int dz = -1; dz <= 0; ++dz)
    for (int dy = -1; dy <= 0; ++dy)
      for (int dx = -1; dx <= 0; ++dx) {
        const int cx = ix + dx, cy = iy + dy, cz = iz + dz;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= m || cy >= m || cz >= m) continue;
        long long ind = 0, pw = 1;
        const int cc[3] = {cx, cy, cz};
        for (int i = 0; i < 3; ++i)
          if (i < p.mode) {
            ind += (long long)p.pdigit[cc[i]] * pw;
            pw *= p.ps;
          }
        if (ind < 0) ind = 0;
        if (ind >= p.n_young) ind = p.n_young - 1;
        const double mu = p.muv[ind], lam = __dmul_rn(mu, p.beta);
        const int li = (-dx) + 2 * (-dy) + 4 * (-dz);
        for (int lj = 0; lj < 8; ++lj) {
          const int ox = (lj & 1) + dx, oy = ((lj >> 1) & 1) + dy, oz = ((lj >> 2) & 1) + dz;
          const int slot = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
          for (int ci = 0; ci < 3; ++ci)
            for (int cj = 0; cj < 3; ++cj) {
              const int e = (3 * li + ci) * 24 + 3 * lj + cj;
              const double t = __dadd_rn(__dmul_rn(lam, p.G1[e]), __dmul_rn(mu, p.G2[e]));
              acc[ci][slot][cj] = __dadd_rn(acc[ci][slot][cj], __dmul_rn(h, t));
            }
        }
        if (rhs && !bnd) {
          const double g1 = 0.5 - 0.5 / sqrt(3.0), g2 = 0.5 + 0.5 / sqrt(3.0);
          const double pi = p.pi, pi2 = pi * pi, hw = h * h * h * 0.125;
          for (int q = 0; q < 8; ++q) {
            const double pt[3] = {-1.0 + (cx + ((q & 1) ? g2 : g1)) * h, -1.0 + (cy + (((q >> 1) & 1) ? g2 : g1)) * h,
                                  -1.0 + (cz + (((q >> 2) & 1) ? g2 : g1)) * h};
            for (int comp = 0; comp < 3; ++comp) {  // ref t3 main.cpp:51-86
              const double x = pt[(0 + comp) % 3], y = pt[(1 + comp) % 3], z = pt[(2 + comp) % 3];
              const double siny = sin(pi * y), sinz = sin(pi * z);
              const double f =
                  2 * pi2 *
                  (-0.25 * lam *
                       (cos(pi * (-2 * x + y + z)) + cos(pi * (2 * x - y + z)) + cos(pi * (2 * x + y - z)) -
                        3 * cos(pi * (2 * x + y + z))) *
                       siny * sinz -
                   mu * (sin(pi * x) * siny * siny * sin(pi * (x + 2 * z)) +
                         sin(pi * x) * sinz * sinz * sin(pi * (x + 2 * y)) + 2 * siny * siny * sinz * sinz * cos(2 * pi * x)));
              b[comp] += p.phi[q * 8 + li] * f * hw;
            }
          }
        }
      }
  auto exact = [&](int jx, int jy, int jz) {  // ref t3 main.cpp:124-132
    const double sv = __dmul_rn(__dmul_rn(p.sn[jx], p.sn[jy]), p.sn[jz]);
    return __dmul_rn(sv, sv);
  };
  for (int ci = 0; ci < 3; ++ci) {
    const long long row = 3 * nd + ci;
    int k = rowptr[row];
    if (bnd) {  // constrained DoF: diagonal only
      col[k] = (int32_t)row;
      const double d = acc[ci][13][ci];
      val[k] = d;
      const double gv = exact(ix, iy, iz);
      if (rhs) rhs[row] = __dmul_rn(gv, d);
      if (x0) x0[row] = gv;
      continue;
    }
    double bi = b[ci];
    for (int oz = -1; oz <= 1; ++oz)
      for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
          const int jx = ix + ox, jy = iy + oy, jz = iz + oz;
          const int slot = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
          if (jx == 0 || jy == 0 || jz == 0 || jx == m || jy == m || jz == m) {
            const double gv = exact(jx, jy, jz);  // condensed inhomogeneous Dirichlet contribution
            for (int cj = 0; cj < 3; ++cj) bi = __dsub_rn(bi, __dmul_rn(acc[ci][slot][cj], gv));
            continue;
          }
          const long long nj = jx + N * (jy + N * (long long)jz);
          for (int cj = 0; cj < 3; ++cj) {
            col[k] = (int32_t)(3 * nj + cj);
            val[k] = acc[ci][slot][cj];
            ++k;
          }
        }
    if (rhs) rhs[row] = bi;
    if (x0) x0[row] = 0.0;
  }
}

static int assemble_elasticity(amgb_ctx* ctx, int m, int ps, int mode, const double* young, int64_t n_young,
                               DeviceCsr& M, double* rhs_device, double* x0_device) {
  if (m < 2 || ps < 1 || mode < 1 || mode > 3 || !young) return AMGB_ERR_BAD_ARG;
  int64_t want = 1;
  for (int i = 0; i < mode; ++i) want *= ps;
  if (n_young != want) return AMGB_ERR_BAD_ARG;
  const int64_t N = (int64_t)m + 1, nn = N * N * N, n = 3 * nn;
  if (n >= (int64_t(1) << 31)) return set_error(ctx, AMGB_ERR_RANGE, "n=%lld: ids are 32-bit", (long long)n);
  // host tables: the same expressions as csrc/gen_q1.cpp
  AsmTables T;
  std::vector<double> zero_eps(n_young, 0.0);
  make_tables(m, ps, mode, zero_eps.data(), n_young, T);  // phi, pattern digits
  const double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  double grad[8][8][3];
  for (int q = 0; q < 8; ++q) {
    const double xq[3] = {g[q & 1], g[(q >> 1) & 1], g[(q >> 2) & 1]};
    for (int l = 0; l < 8; ++l) {
      const int sg[3] = {l & 1, (l >> 1) & 1, (l >> 2) & 1};
      double f[3], df[3];
      for (int d = 0; d < 3; ++d) {
        f[d] = sg[d] ? xq[d] : 1.0 - xq[d];
        df[d] = sg[d] ? 1.0 : -1.0;
      }
      grad[q][l][0] = df[0] * f[1] * f[2];
      grad[q][l][1] = f[0] * df[1] * f[2];
      grad[q][l][2] = f[0] * f[1] * df[2];
    }
  }
  std::vector<double> flat(2 * 576 + (size_t)n_young + (size_t)m + 1 + 64);
  double* G1 = flat.data();
  double* G2 = G1 + 576;
  double* muv = G2 + 576;
  double* sn = muv + n_young;
  double* phi = sn + m + 1;
  for (int i = 0; i < 8; ++i)
    for (int ci = 0; ci < 3; ++ci)
      for (int j = 0; j < 8; ++j)
        for (int cj = 0; cj < 3; ++cj) {
          double s1 = 0.0, s2 = 0.0;
          for (int q = 0; q < 8; ++q) {
            s1 += grad[q][i][ci] * grad[q][j][cj] * 0.125;
            double t = grad[q][i][cj] * grad[q][j][ci];
            if (ci == cj) t += grad[q][i][0] * grad[q][j][0] + grad[q][i][1] * grad[q][j][1] + grad[q][i][2] * grad[q][j][2];
            s2 += t * 0.125;
          }
          G1[(3 * i + ci) * 24 + 3 * j + cj] = s1;
          G2[(3 * i + ci) * 24 + 3 * j + cj] = s2;
        }
  const double nu = 0.29, e_min = 1000.0;  // ref t3 main.cpp:48-49
  for (int64_t i = 0; i < n_young; ++i) muv[i] = e_min * young[i] / (1.0 + nu);
  const double h = 2.0 / m, pi = M_PI * ps / 2.0;
  for (int64_t i = 0; i <= m; ++i) sn[i] = std::sin(pi * (-1.0 + i * h));
  for (int i = 0; i < 64; ++i) phi[i] = T.phi[i];
  DevBuf<double> tab;
  DevBuf<int32_t> pd, len;
  AMGB_TRY(tab.alloc(ctx, flat.size()));
  AMGB_TRY(pd.alloc(ctx, m));
  AMGB_CUDA(ctx, cudaMemcpyAsync(tab.p, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AMGB_CUDA(ctx, cudaMemcpyAsync(pd.p, T.pdigit.data(), (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  ElasParams p;
  p.m = m;
  p.mode = mode;
  p.ps = ps;
  p.n_young = n_young;
  p.h = h;
  p.beta = nu / (1.0 - 2.0 * nu);
  p.pi = pi;
  p.G1 = tab.p;
  p.G2 = tab.p + 576;
  p.muv = tab.p + 1152;
  p.sn = p.muv + n_young;
  p.phi = p.sn + m + 1;
  p.pdigit = pd.p;
  M.n = M.ncols = n;
  AMGB_TRY(len.alloc(ctx, n));
  AMGB_TRY(M.rp.alloc(ctx, n + 1));
  const unsigned grid = (unsigned)div_up(nn, kAsmBlock);
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * nn, elas_row_len_kernel, grid, kAsmBlock, 0, m, (long long)nn, len.p);
  // 64-bit total first: the 32-bit scan would wrap silently (config 3 holds 1.55e9 entries)
  const double est = 81.0 * 3.0 * (double)nn;
  if (est >= 2147483648.0 * 1.2) return set_error(ctx, AMGB_ERR_RANGE, "elasticity system too large for 32-bit row pointers");
  AMGB_TRY(exclusive_scan_i32(ctx, len.p, M.rp.p, n));
  int32_t nnz = 0;
  AMGB_TRY(read_i32(ctx, M.rp.p + n, &nnz));
  if (nnz < 0) return set_error(ctx, AMGB_ERR_RANGE, "nnz overflows 32-bit row pointers");
  M.nnz = nnz;
  AMGB_TRY(M.col.alloc(ctx, nnz));
  AMGB_TRY(M.val.alloc(ctx, nnz));
  AMGB_LAUNCH(ctx, F_AUX, 12.0 * nnz + 16.0 * n, asm_elasticity_kernel, grid, kAsmBlock, 0, p, (long long)nn,
              (const int32_t*)M.rp.p, M.col.p, M.val.p, rhs_device, x0_device);
  AMGB_CHECK_LAUNCH(ctx);
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the tables go out of scope
  return AMGB_OK;
}

}  // namespace amgb

using namespace amgb;

extern "C" {

int amgb_matrix_assemble_poisson_q1(amgb_ctx* ctx, int32_t m, int32_t pattern_size, int32_t mode, const double* epsv,
                                    int64_t n_epsv, amgb_matrix** out, double* rhs_device, double* x0_device) {
  if (!ctx || !out) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  const int64_t N = (int64_t)m + 1;
  amgb_matrix* M = new amgb_matrix;
  M->ctx = ctx;
  const int rc = assemble_rows(ctx, m, pattern_size, mode, epsv, n_epsv, 0, N * N * N, M->A, rhs_device, x0_device);
  if (rc != AMGB_OK) {
    delete M;
    return rc;
  }
  M->A.ncols = M->A.n;
  *out = M;
  return AMGB_OK;
}

int amgb_dist_matrix_assemble_poisson_q1(amgb_ctx* ctx, amgb_comm* comm, int32_t m, int32_t pattern_size,
                                         int32_t mode, const double* epsv, int64_t n_epsv, int64_t row_begin,
                                         int64_t row_end, amgb_dist_matrix** out, double* rhs_device,
                                         double* x0_device) {
  if (!ctx || !comm || !out) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  const int64_t N = (int64_t)m + 1, n = N * N * N;
  amgb_dist_matrix* M = new amgb_dist_matrix;
  M->ctx = ctx;
  M->comm = comm;
  OwnedCsr& o = M->own;
  o.n_global = n;
  o.g0 = row_begin;
  std::vector<int64_t> begins(comm->size);
  int rc = comm->allgather_host(ctx, &row_begin, sizeof(int64_t), begins.data());
  o.starts.assign(comm->size + 1, n);
  for (int q = 0; q < comm->size; ++q) o.starts[q] = begins[q];
  for (int q = 0; q < comm->size && rc == AMGB_OK; ++q)
    if (o.starts[q] > o.starts[q + 1]) rc = set_error(ctx, AMGB_ERR_BAD_ARG, "row ranges must ascend with the rank");
  if (rc == AMGB_OK)
    rc = assemble_rows(ctx, m, pattern_size, mode, epsv, n_epsv, row_begin, row_end, o.M, rhs_device, x0_device);
  // collective verdict: a rank that failed locally must not leave its peers waiting later on
  int64_t good = rc == AMGB_OK ? 1 : 0;
  const int arc = allreduce_min_i64_host(ctx, comm, &good);
  if (rc == AMGB_OK && arc != AMGB_OK) rc = arc;
  if (rc == AMGB_OK && !good) rc = set_error(ctx, AMGB_ERR_COMM, "another rank could not assemble its slab");
  if (rc != AMGB_OK) {
    delete M;
    return rc;
  }
  *out = M;
  return AMGB_OK;
}

int amgb_matrix_assemble_elasticity_q1(amgb_ctx* ctx, int32_t m, int32_t pattern_size, int32_t mode, const double* young,
                                       int64_t n_young, amgb_matrix** out, double* rhs_device, double* x0_device) {
  if (!ctx || !out) return AMGB_ERR_BAD_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  amgb_matrix* M = new amgb_matrix;
  M->ctx = ctx;
  const int rc = assemble_elasticity(ctx, m, pattern_size, mode, young, n_young, M->A, rhs_device, x0_device);
  if (rc != AMGB_OK) {
    delete M;
    return rc;
  }
  *out = M;
  return AMGB_OK;
}

/* Same, with the right-hand side and the initial guess returned in HOST arrays (n doubles
 * each, or NULL): for host code that keeps its vectors on the host (dealii_compat). */
int amgb_matrix_assemble_poisson_q1_hostvec(amgb_ctx* ctx, int32_t m, int32_t pattern_size, int32_t mode,
                                            const double* epsv, int64_t n_epsv, amgb_matrix** out, double* rhs_host,
                                            double* x0_host) {
  if (!ctx || !out || m < 1) return AMGB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  const int64_t N = (int64_t)m + 1, n = N * N * N;
  DevBuf<double> drhs, dx0;
  if (rhs_host) AMGB_TRY(drhs.alloc(ctx, n));
  if (x0_host) AMGB_TRY(dx0.alloc(ctx, n));
  AMGB_TRY(amgb_matrix_assemble_poisson_q1(ctx, m, pattern_size, mode, epsv, n_epsv, out, drhs.p, dx0.p));
  if (rhs_host) AMGB_CUDA(ctx, cudaMemcpyAsync(rhs_host, drhs.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (x0_host) AMGB_CUDA(ctx, cudaMemcpyAsync(x0_host, dx0.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

/* CSR of a resident matrix back to the host (tests; rowptr may be NULL etc.). */
int amgb_matrix_download_csr(const amgb_matrix* A, int32_t* rowptr, int32_t* col, double* val) {
  if (!A) return AMGB_ERR_BAD_ARG;
  amgb_ctx* ctx = A->ctx;
  cudaSetDevice(ctx->device);
  if (rowptr)
    AMGB_CUDA(ctx, cudaMemcpyAsync(rowptr, A->A.rp.p, (A->A.n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                   ctx->stream));
  if (col)
    AMGB_CUDA(ctx, cudaMemcpyAsync(col, A->A.col.p, A->A.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (val)
    AMGB_CUDA(ctx, cudaMemcpyAsync(val, A->A.val.p, A->A.nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMGB_OK;
}

}  // extern "C"
