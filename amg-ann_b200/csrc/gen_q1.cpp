// Synthetic Q1 finite-element system generators (host C++, no CUDA).
// Interface and reference citations: include/amgb_gen.h.
//
// Assembly is row-wise: for every node the <=8 adjacent cells are visited in a
// fixed (z,y,x) order and their element-matrix contributions are accumulated
// into a 27-slot stencil, so the result is deterministic and independent of the
// number of OpenMP threads.
#include "amgb_gen.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <random>
#include <vector>

namespace {

constexpr int kOk = 0;
constexpr int kBadArg = -1;

struct RefElement {
  // 2x2x2 Gauss quadrature on the unit cube, trilinear shape functions.
  // local node l = a + 2*b + 4*c,  (a,b,c) in {0,1}^3.
  double w[8];
  double phi[8][8];      // phi[q][l]
  double grad[8][8][3];  // grad[q][l][d] on the unit cube
  RefElement() {
    const double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
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
        phi[q][l] = f[0] * f[1] * f[2];
        grad[q][l][0] = df[0] * f[1] * f[2];
        grad[q][l][1] = f[0] * df[1] * f[2];
        grad[q][l][2] = f[0] * f[1] * df[2];
      }
    }
  }
};

// 1-D factors of the manufactured solutions, ref common/cube_solution.h:6-28.
struct Sol1D {
  int id;  // 0: sin, 3: cos
  double v(double t, double f, int der) const {
    if (id == 0) {
      if (der == 0) return std::sin(f * t);
      if (der == 1) return f * std::cos(f * t);
      return -f * f * std::sin(f * t);
    }
    if (der == 0) return std::cos(f * t);
    if (der == 1) return -f * std::sin(f * t);
    return -f * f * std::cos(f * t);
  }
};

inline long ipow(long b, int e) {
  long r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

// Pattern index of a point, ref t2 main.cpp:101-113.
inline long pattern_index(const double p[3], int ps, int mode) {
  const double hp = 2.0 / double(ps);
  long ind = 0;
  for (int i = 0; i < mode; ++i)
    ind += long(std::trunc((p[i] + 1.0) / (hp + 1e-15))) * ipow(ps, i);
  return ind;
}

struct Grid {
  int m;        // cells per direction
  int64_t N;    // nodes per direction
  double h;
  int64_t node(int64_t ix, int64_t iy, int64_t iz) const { return ix + N * (iy + N * iz); }
  bool boundary(int64_t ix, int64_t iy, int64_t iz) const {
    return ix == 0 || iy == 0 || iz == 0 || ix == m || iy == m || iz == m;
  }
};

inline int valid_per_dim(const Grid& g, int64_t i) { return (i == 0 || i == g.m) ? 2 : 3; }

}  // namespace

extern "C" {

int amgb_gen_sizes(int kind, int m, int64_t* n, int64_t* nnz) {
  if (m < 1 || !n || !nnz) return kBadArg;
  const int64_t N = int64_t(m) + 1;
  if (kind == 0) {
    *n = N * N * N;
    const int64_t t = 3 * int64_t(m) + 1;
    *nnz = t * t * t;
    return kOk;
  }
  if (kind == 1) {
    if (m < 2) return kBadArg;
    *n = 3 * N * N * N;
    const int64_t ni = int64_t(m) - 1;             // interior nodes per dir
    const int64_t pairs = 3 * ni - 2;              // interior-interior pairs per dir
    const int64_t nb = N * N * N - ni * ni * ni;   // boundary nodes
    *nnz = 9 * pairs * pairs * pairs + 3 * nb;
    return kOk;
  }
  return kBadArg;
}

int amgb_gen_poisson_q1_range_sizes(int m, int64_t row_begin, int64_t row_end,
                                    int64_t* nnz) {
  if (m < 1 || !nnz) return kBadArg;
  Grid g{m, int64_t(m) + 1, 2.0 / m};
  const int64_t n = g.N * g.N * g.N;
  if (row_begin < 0 || row_end > n || row_begin > row_end) return kBadArg;
  int64_t total = 0;
  for (int64_t r = row_begin; r < row_end; ++r) {
    const int64_t ix = r % g.N, iy = (r / g.N) % g.N, iz = r / (g.N * g.N);
    total += int64_t(valid_per_dim(g, ix)) * valid_per_dim(g, iy) * valid_per_dim(g, iz);
  }
  *nnz = total;
  return kOk;
}

int amgb_gen_poisson_q1(int m, int pattern_size, int mode, const double* epsv,
                        int64_t n_epsv, int64_t row_begin, int64_t row_end,
                        int64_t* rowptr, int32_t* col, double* val, double* rhs,
                        double* x0) {
  if (m < 1 || pattern_size < 1 || mode < 1 || mode > 3 || !epsv || !rowptr || !col || !val)
    return kBadArg;
  if (n_epsv != ipow(pattern_size, mode)) return kBadArg;
  Grid g{m, int64_t(m) + 1, 2.0 / m};
  const int64_t n = g.N * g.N * g.N;
  if (n >= (int64_t(1) << 31)) return kBadArg;  // int32 column ids
  if (row_begin < 0 || row_end > n || row_begin > row_end) return kBadArg;
  const int64_t nloc = row_end - row_begin;

  static const RefElement re;
  // Unit-mu element matrix and load-shape table; the physical element matrix is
  // mu*h*Kq (gradients scale 1/h, JxW = h^3 w).
  double Kq[8][8];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double s = 0.0;
      for (int q = 0; q < 8; ++q)
        s += (re.grad[q][i][0] * re.grad[q][j][0] + re.grad[q][i][1] * re.grad[q][j][1] +
              re.grad[q][i][2] * re.grad[q][j][2]) *
             re.w[q];
      Kq[i][j] = s;
    }
  std::vector<double> diffv(n_epsv);
  for (int64_t i = 0; i < n_epsv; ++i) diffv[i] = std::pow(10.0, epsv[i]);

  const Sol1D sol{3 * (1 - (pattern_size % 2))};   // ref t2 main.cpp:70-72
  const double freq = M_PI / (2.0 / double(pattern_size));  // ref t2 main.cpp:80-81
  const double h = g.h;

  auto cell_mu = [&](int64_t cx, int64_t cy, int64_t cz) {
    const double c[3] = {-1.0 + (cx + 0.5) * h, -1.0 + (cy + 0.5) * h, -1.0 + (cz + 0.5) * h};
    long ind = pattern_index(c, pattern_size, mode);
    if (ind < 0) ind = 0;
    if (ind >= n_epsv) ind = n_epsv - 1;
    return diffv[ind];
  };
  // value the reference puts on Dirichlet diagonals: |first non-zero diagonal|
  // = |A_00| before boundary treatment (one cell touches node 0).
  const double dirichlet_diag = std::fabs(cell_mu(0, 0, 0) * h * Kq[0][0]);

  rowptr[0] = 0;
  for (int64_t r = 0; r < nloc; ++r) {
    const int64_t gr = row_begin + r;
    const int64_t ix = gr % g.N, iy = (gr / g.N) % g.N, iz = gr / (g.N * g.N);
    rowptr[r + 1] = rowptr[r] + int64_t(valid_per_dim(g, ix)) * valid_per_dim(g, iy) *
                                    valid_per_dim(g, iz);
  }

#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nloc; ++r) {
    const int64_t gr = row_begin + r;
    const int64_t ix = gr % g.N, iy = (gr / g.N) % g.N, iz = gr / (g.N * g.N);
    double acc[27];
    for (int s = 0; s < 27; ++s) acc[s] = 0.0;
    double b = 0.0;
    const bool bnd = g.boundary(ix, iy, iz);
    for (int dz = -1; dz <= 0; ++dz)
      for (int dy = -1; dy <= 0; ++dy)
        for (int dx = -1; dx <= 0; ++dx) {
          const int64_t cx = ix + dx, cy = iy + dy, cz = iz + dz;
          if (cx < 0 || cy < 0 || cz < 0 || cx >= m || cy >= m || cz >= m) continue;
          const double mu = cell_mu(cx, cy, cz);
          const int li = (-dx) + 2 * (-dy) + 4 * (-dz);
          for (int lj = 0; lj < 8; ++lj) {
            const int ox = (lj & 1) + dx, oy = ((lj >> 1) & 1) + dy, oz = ((lj >> 2) & 1) + dz;
            acc[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)] += mu * h * Kq[li][lj];
          }
          if (rhs && !bnd) {
            // cell_rhs(i) += mu * phi_i * f(x_q) * JxW, ref t2 main.cpp:296-299
            for (int q = 0; q < 8; ++q) {
              const double g1 = 0.5 - 0.5 / std::sqrt(3.0), g2 = 0.5 + 0.5 / std::sqrt(3.0);
              const double p[3] = {-1.0 + (cx + ((q & 1) ? g2 : g1)) * h,
                                   -1.0 + (cy + (((q >> 1) & 1) ? g2 : g1)) * h,
                                   -1.0 + (cz + (((q >> 2) & 1) ? g2 : g1)) * h};
              double f = 0.0;  // ref t2 main.cpp:156-169
              for (int i = 0; i < 3; ++i) {
                double dd = 1.0;
                for (int j = 0; j < 3; ++j) dd *= sol.v(p[j], freq, 2 * (i == j));
                f += dd;
              }
              b += mu * re.phi[q][li] * (-f) * (h * h * h * re.w[q]);
            }
          }
        }
    int64_t k = rowptr[r];
    for (int oz = -1; oz <= 1; ++oz) {
      if (iz + oz < 0 || iz + oz > m) continue;
      for (int oy = -1; oy <= 1; ++oy) {
        if (iy + oy < 0 || iy + oy > m) continue;
        for (int ox = -1; ox <= 1; ++ox) {
          if (ix + ox < 0 || ix + ox > m) continue;
          const bool diag = (ox == 0 && oy == 0 && oz == 0);
          col[k] = int32_t(g.node(ix + ox, iy + oy, iz + oz));
          val[k] = bnd ? (diag ? dirichlet_diag : 0.0) : acc[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)];
          ++k;
        }
      }
    }
    if (rhs || x0) {
      double bv = 0.0;
      if (bnd) {
        const double p[3] = {-1.0 + ix * h, -1.0 + iy * h, -1.0 + iz * h};
        bv = sol.v(p[0], freq, 0) * sol.v(p[1], freq, 0) * sol.v(p[2], freq, 0);
      }
      if (rhs) rhs[r] = bnd ? bv * dirichlet_diag : b;
      if (x0) x0[r] = bnd ? bv : 0.0;
    }
  }
  return kOk;
}

int amgb_gen_elasticity_q1(int m, int pattern_size, int mode, const double* young,
                           int64_t n_young, int64_t* rowptr, int32_t* col,
                           double* val, double* rhs, double* x0) {
  if (m < 2 || pattern_size < 1 || mode < 1 || mode > 3 || !young || !rowptr || !col || !val)
    return kBadArg;
  if (n_young != ipow(pattern_size, mode)) return kBadArg;
  Grid g{m, int64_t(m) + 1, 2.0 / m};
  const int64_t nn = g.N * g.N * g.N;
  const int64_t n = 3 * nn;
  int64_t n_chk, nnz_chk;
  amgb_gen_sizes(1, m, &n_chk, &nnz_chk);
  if (n >= (int64_t(1) << 31) || nnz_chk >= (int64_t(1) << 31)) return kBadArg;

  static const RefElement re;
  const double nu = 0.29, e_min = 1000.0;  // ref t3 main.cpp:48-49
  const double beta = nu / (1.0 - 2.0 * nu);
  const double h = g.h;
  // G1[(i,ci),(j,cj)] = sum_q d_ci phi_i d_cj phi_j w          (lambda term)
  // G2[(i,ci),(j,cj)] = sum_q (d_cj phi_i d_ci phi_j + delta grad.grad) w   (mu terms)
  // ref t3 main.cpp:320-342; physical scaling h.
  static double G1[24][24], G2[24][24];
  for (int i = 0; i < 8; ++i)
    for (int ci = 0; ci < 3; ++ci)
      for (int j = 0; j < 8; ++j)
        for (int cj = 0; cj < 3; ++cj) {
          double s1 = 0.0, s2 = 0.0;
          for (int q = 0; q < 8; ++q) {
            s1 += re.grad[q][i][ci] * re.grad[q][j][cj] * re.w[q];
            double t = re.grad[q][i][cj] * re.grad[q][j][ci];
            if (ci == cj)
              t += re.grad[q][i][0] * re.grad[q][j][0] + re.grad[q][i][1] * re.grad[q][j][1] +
                   re.grad[q][i][2] * re.grad[q][j][2];
            s2 += t * re.w[q];
          }
          G1[3 * i + ci][3 * j + cj] = s1;
          G2[3 * i + ci][3 * j + cj] = s2;
        }

  auto cell_mu = [&](int64_t cx, int64_t cy, int64_t cz) {
    const double c[3] = {-1.0 + (cx + 0.5) * h, -1.0 + (cy + 0.5) * h, -1.0 + (cz + 0.5) * h};
    long ind = pattern_index(c, pattern_size, mode);
    if (ind < 0) ind = 0;
    if (ind >= n_young) ind = n_young - 1;
    return e_min * young[ind] / (1.0 + nu);  // ref t3 main.cpp:88-99
  };
  const double pi = M_PI * pattern_size / 2.0;
  auto exact = [&](int64_t ix, int64_t iy, int64_t iz) {  // ref t3 main.cpp:124-132
    const double s = std::sin(pi * (-1.0 + ix * h)) * std::sin(pi * (-1.0 + iy * h)) *
                     std::sin(pi * (-1.0 + iz * h));
    return s * s;
  };

  // row lengths
  rowptr[0] = 0;
  for (int64_t nd = 0; nd < nn; ++nd) {
    const int64_t ix = nd % g.N, iy = (nd / g.N) % g.N, iz = nd / (g.N * g.N);
    int64_t len = 1;
    if (!g.boundary(ix, iy, iz)) {
      auto cnt = [&](int64_t i) { return 3 - (i == 1) - (i == m - 1); };
      len = 3 * cnt(ix) * cnt(iy) * cnt(iz);
      if (m == 2) len = 3;  // single interior node
    }
    for (int c = 0; c < 3; ++c) rowptr[3 * nd + c + 1] = rowptr[3 * nd + c] + len;
  }

#pragma omp parallel for schedule(static)
  for (int64_t nd = 0; nd < nn; ++nd) {
    const int64_t ix = nd % g.N, iy = (nd / g.N) % g.N, iz = nd / (g.N * g.N);
    const bool bnd = g.boundary(ix, iy, iz);
    double acc[3][27][3];
    double b[3] = {0, 0, 0};
    for (int c = 0; c < 3; ++c)
      for (int s = 0; s < 27; ++s) acc[c][s][0] = acc[c][s][1] = acc[c][s][2] = 0.0;
    for (int dz = -1; dz <= 0; ++dz)
      for (int dy = -1; dy <= 0; ++dy)
        for (int dx = -1; dx <= 0; ++dx) {
          const int64_t cx = ix + dx, cy = iy + dy, cz = iz + dz;
          if (cx < 0 || cy < 0 || cz < 0 || cx >= m || cy >= m || cz >= m) continue;
          const double mu = cell_mu(cx, cy, cz), lam = mu * beta;
          const int li = (-dx) + 2 * (-dy) + 4 * (-dz);
          for (int lj = 0; lj < 8; ++lj) {
            const int ox = (lj & 1) + dx, oy = ((lj >> 1) & 1) + dy, oz = ((lj >> 2) & 1) + dz;
            const int slot = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
            for (int ci = 0; ci < 3; ++ci)
              for (int cj = 0; cj < 3; ++cj)
                acc[ci][slot][cj] +=
                    h * (lam * G1[3 * li + ci][3 * lj + cj] + mu * G2[3 * li + ci][3 * lj + cj]);
          }
          if (rhs && !bnd) {
            for (int q = 0; q < 8; ++q) {
              const double g1 = 0.5 - 0.5 / std::sqrt(3.0), g2 = 0.5 + 0.5 / std::sqrt(3.0);
              const double p[3] = {-1.0 + (cx + ((q & 1) ? g2 : g1)) * h,
                                   -1.0 + (cy + (((q >> 1) & 1) ? g2 : g1)) * h,
                                   -1.0 + (cz + (((q >> 2) & 1) ? g2 : g1)) * h};
              const double pi2 = pi * pi;
              for (int comp = 0; comp < 3; ++comp) {  // ref t3 main.cpp:51-86
                const double x = p[(0 + comp) % 3], y = p[(1 + comp) % 3], z = p[(2 + comp) % 3];
                const double siny = std::sin(pi * y), sinz = std::sin(pi * z);
                const double f =
                    2 * pi2 *
                    (-0.25 * lam *
                         (std::cos(pi * (-2 * x + y + z)) + std::cos(pi * (2 * x - y + z)) +
                          std::cos(pi * (2 * x + y - z)) - 3 * std::cos(pi * (2 * x + y + z))) *
                         siny * sinz -
                     mu * (std::sin(pi * x) * siny * siny * std::sin(pi * (x + 2 * z)) +
                           std::sin(pi * x) * sinz * sinz * std::sin(pi * (x + 2 * y)) +
                           2 * siny * siny * sinz * sinz * std::cos(2 * pi * x)));
                b[comp] += re.phi[q][li] * f * (h * h * h * re.w[q]);
              }
            }
          }
        }
    for (int ci = 0; ci < 3; ++ci) {
      const int64_t row = 3 * nd + ci;
      int64_t k = rowptr[row];
      if (bnd) {
        // constrained DoF: diagonal only (keep_constrained_dofs=false)
        col[k] = int32_t(row);
        val[k] = acc[ci][13][ci];
        const double gv = exact(ix, iy, iz);
        if (rhs) rhs[row] = gv * val[k];
        if (x0) x0[row] = gv;
        continue;
      }
      double bi = b[ci];
      for (int oz = -1; oz <= 1; ++oz)
        for (int oy = -1; oy <= 1; ++oy)
          for (int ox = -1; ox <= 1; ++ox) {
            const int64_t jx = ix + ox, jy = iy + oy, jz = iz + oz;
            const int slot = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
            if (g.boundary(jx, jy, jz)) {
              // condensed inhomogeneous Dirichlet contribution
              const double gv = exact(jx, jy, jz);
              for (int cj = 0; cj < 3; ++cj) bi -= acc[ci][slot][cj] * gv;
              continue;
            }
            const int64_t nj = g.node(jx, jy, jz);
            for (int cj = 0; cj < 3; ++cj) {
              col[k] = int32_t(3 * nj + cj);
              val[k] = acc[ci][slot][cj];
              ++k;
            }
          }
      if (rhs) rhs[row] = bi;
      if (x0) x0[row] = 0.0;
    }
  }
  return kOk;
}

int amgb_gen_cuthill_mckee(int64_t n, const int64_t* rowptr, const int32_t* col, int reversed,
                           int32_t* new_to_old) {
  if (n < 0 || !rowptr || (n > 0 && !new_to_old) || n >= (int64_t(1) << 31)) return kBadArg;
  std::vector<uint8_t> done(n, 0);
  std::vector<int32_t> front, next;
  int64_t numbered = 0, scan_from = 0;
  auto degree = [&](int32_t i) { return rowptr[i + 1] - rowptr[i]; };
  while (numbered < n) {
    // starting point of the next component: smallest coordination number, lowest index
    int32_t start = -1;
    for (int64_t i = scan_from; i < n; ++i)
      if (!done[i] && (start < 0 || degree((int32_t)i) < degree(start))) start = (int32_t)i;
    while (scan_from < n && done[scan_from]) ++scan_from;
    done[start] = 1;
    new_to_old[numbered++] = start;
    front.assign(1, start);
    while (!front.empty()) {
      next.clear();
      for (int32_t i : front)
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
          const int32_t j = col[k];
          if (j < 0 || j >= n) return kBadArg;
          if (!done[j]) {
            done[j] = 1;
            next.push_back(j);
          }
        }
      // by coordination number, ties by index (deal.II: sorted list fed into a multimap)
      std::sort(next.begin(), next.end(), [&](int32_t a, int32_t b) {
        const int64_t da = degree(a), db = degree(b);
        return da != db ? da < db : a < b;
      });
      for (int32_t j : next) new_to_old[numbered++] = j;
      front.swap(next);
    }
  }
  if (reversed) std::reverse(new_to_old, new_to_old + n);
  return kOk;
}

int amgb_gen_random_vec(int64_t seed, int64_t len, double max, double* out) {
  if (len < 0 || (len > 0 && !out)) return kBadArg;
  std::uniform_real_distribution<double> distribution(0.0, max);
  std::default_random_engine generator(seed);
  for (int64_t i = 0; i < len; ++i) out[i] = distribution(generator);
  return kOk;
}

int amgb_gen_checkerboard_epsv(int pattern_size, int mode, double contrast_exp, double* out) {
  if (pattern_size < 1 || mode < 1 || mode > 3 || !out) return kBadArg;
  const long size = ipow(pattern_size, mode);
  for (long u0 = 0; u0 < size; ++u0) {
    long u = u0, s = 0;
    for (int t = mode - 1; t >= 1; --t) {
      s += u / ipow(pattern_size, t);
      u = u % ipow(pattern_size, t);
    }
    s += u % pattern_size;
    out[u0] = contrast_exp * double(s % 2);
  }
  return kOk;
}

}  // extern "C"
