// CPU oracle: restatement of the reference's AMG-PCG path (hypre BoomerAMG via
// PETSc via deal.II) and a literal port of its pooling operator.
// TEST INFRASTRUCTURE ONLY -- see amg_oracle.h for scope and the
// "parity unpinned" statement.  Section references (A.x) are to SURVEY.md
// Appendix A; "ref" paths are relative to /root/reference/code/data-generation/.
#include "amg_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Csr {
  int64_t n = 0, ncols = 0;
  std::vector<int32_t> rp, col;
  std::vector<double> val;
  int64_t nnz() const { return (int64_t)col.size(); }
};

// ---------------------------------------------------------------------------
// hypre utilities/random.c: multiplicative LCG a=16807, m=2^31-1, seed 2747+rank.
// ---------------------------------------------------------------------------
constexpr int64_t kRandA = 16807, kRandM = 2147483647;
constexpr int64_t kRandSeed = 2747;

inline int64_t rand_next(int64_t seed) { return (kRandA * seed) % kRandM; }

// ---------------------------------------------------------------------------
// A.3 Strength: hypre_BoomerAMGCreateS (par_strength.c), num_functions = 1.
// Row storage here is ascending columns; hypre keeps the diagonal first, so the
// row sum is accumulated diagonal first, then the off-diagonals left to right.
// ---------------------------------------------------------------------------
void strength(const Csr& A, double theta, double max_row_sum, std::vector<uint8_t>& mask) {
  mask.assign(A.nnz(), 0);
  for (int64_t i = 0; i < A.n; ++i) {
    const int32_t b = A.rp[i], e = A.rp[i + 1];
    double diag = 0.0;
    for (int32_t k = b; k < e; ++k)
      if (A.col[k] == i) diag = A.val[k];
    double row_scale = 0.0, row_sum = diag;
    if (diag < 0) {
      for (int32_t k = b; k < e; ++k)
        if (A.col[k] != i) {
          row_scale = std::max(row_scale, A.val[k]);
          row_sum += A.val[k];
        }
    } else {
      for (int32_t k = b; k < e; ++k)
        if (A.col[k] != i) {
          row_scale = std::min(row_scale, A.val[k]);
          row_sum += A.val[k];
        }
    }
    if (std::fabs(row_sum) > std::fabs(diag) * max_row_sum && max_row_sum < 1.0) continue;
    const double thr = theta * row_scale;
    if (diag < 0) {
      for (int32_t k = b; k < e; ++k)
        if (A.col[k] != i && A.val[k] > thr) mask[k] = 1;
    } else {
      for (int32_t k = b; k < e; ++k)
        if (A.col[k] != i && A.val[k] < thr) mask[k] = 1;
    }
  }
}

// ---------------------------------------------------------------------------
// A.3 PMIS: hypre_BoomerAMGCoarsenPMIS (par_coarsen.c), one rank, CF_init = 0.
// measure_i = |S^T_i| + hypre_Rand(); rows without strong connections are
// special F points (-3); points nobody depends on (measure < 1) are F; then
// synchronous independent-set rounds (hypre_BoomerAMGIndepSet) on the
// undecided graph: i joins the set iff its measure beats that of every
// undecided neighbour in S_i and S^T_i; undecided points that strongly depend
// on a C point become F.  Measures are static (no CLJP-style updates).
// ---------------------------------------------------------------------------
void coarsen_pmis(int64_t n, const int32_t* rp, const int32_t* col, const uint8_t* mask,
                  std::vector<int32_t>& cf) {
  std::vector<double> measure(n, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
      if (mask[k]) measure[col[k]] += 1.0;
  int64_t seed = kRandSeed;
  for (int64_t i = 0; i < n; ++i) {
    seed = rand_next(seed);
    measure[i] += double(seed) / double(kRandM);
  }
  cf.assign(n, 0);
  std::vector<int32_t> graph;
  graph.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    bool any = false;
    for (int32_t k = rp[i]; k < rp[i + 1] && !any; ++k) any = mask[k];
    if (!any) {
      cf[i] = -3;
      measure[i] = 0.0;
    } else if (measure[i] < 1.0) {
      cf[i] = -1;
      measure[i] = 0.0;
    } else {
      graph.push_back((int32_t)i);
    }
  }
  std::vector<int32_t> mark(n, 0);  // tentative independent-set membership
  while (!graph.empty()) {
    for (int32_t i : graph) mark[i] = measure[i] > 1.0 ? 1 : 0;
    for (int32_t i : graph) {
      if (!(measure[i] > 1.0)) continue;
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k) {
        if (!mask[k]) continue;
        const int32_t j = col[k];
        if (measure[j] > 1.0) {
          if (measure[i] > measure[j])
            mark[j] = 0;
          else if (measure[j] > measure[i])
            mark[i] = 0;
        }
      }
    }
    for (int32_t i : graph)
      if (mark[i]) cf[i] = 1;
    for (int32_t i : graph) {
      if (cf[i] != 0) continue;
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
        if (mask[k] && cf[col[k]] > 0) {
          cf[i] = -1;
          break;
        }
    }
    size_t w = 0;
    for (int32_t i : graph) {
      if (cf[i] != 0) {
        measure[i] = 0.0;
        mark[i] = 0;
      } else {
        graph[w++] = i;
      }
    }
    graph.resize(w);
  }
}

// ---------------------------------------------------------------------------
// A.3 Falgout (coarsen type 6) on one rank.
//   stage 1: hypre_BoomerAMGCoarsenRuge first pass -- measure = |S^T_i|, bucket
//            lists (amg_linklist.c: insertion at the tail of a bucket, selection
//            of the head of the highest bucket), dynamic measure updates;
//   stage 2: Ruge second pass -- every strong F-F pair must share a C point;
//   stage 3: hypre_BoomerAMGCoarsen (CLJP) with CF_init = 1: C points kept,
//            the rest re-examined with measures |S^T_i| + hypre_Rand().
// ---------------------------------------------------------------------------
struct Buckets {
  // doubly linked lists keyed by integer measure
  std::vector<int32_t> head, tail, next, prev, key;
  int32_t top = 0;
  Buckets(int64_t n, int32_t maxkey)
      : head(maxkey + 2, -1), tail(maxkey + 2, -1), next(n, -1), prev(n, -1), key(n, -1) {}
  void insert(int32_t i, int32_t k) {
    if (k >= (int32_t)head.size()) {
      head.resize(k + 2, -1);
      tail.resize(k + 2, -1);
    }
    key[i] = k;
    next[i] = -1;
    prev[i] = tail[k];
    if (tail[k] >= 0) next[tail[k]] = i; else head[k] = i;
    tail[k] = i;
    if (k > top) top = k;
  }
  void remove(int32_t i) {
    const int32_t k = key[i];
    if (k < 0) return;
    if (prev[i] >= 0) next[prev[i]] = next[i]; else head[k] = next[i];
    if (next[i] >= 0) prev[next[i]] = prev[i]; else tail[k] = prev[i];
    key[i] = -1;
  }
  int32_t pop_max() {
    while (top > 0 && head[top] < 0) --top;
    if (top <= 0) return -1;
    return head[top];
  }
};

void coarsen_falgout(int64_t n, const int32_t* rp, const int32_t* col, const uint8_t* mask,
                     std::vector<int32_t>& cf) {
  // S (rows) and S^T (columns) as compact patterns
  std::vector<int32_t> srp(n + 1, 0), scol, strp(n + 1, 0), stcol;
  for (int64_t i = 0; i < n; ++i) {
    int32_t c = 0;
    for (int32_t k = rp[i]; k < rp[i + 1]; ++k) c += mask[k] ? 1 : 0;
    srp[i + 1] = srp[i] + c;
  }
  scol.resize(srp[n]);
  for (int64_t i = 0; i < n; ++i) {
    int32_t w = srp[i];
    for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
      if (mask[k]) {
        scol[w++] = col[k];
        strp[col[k] + 1]++;
      }
  }
  for (int64_t i = 0; i < n; ++i) strp[i + 1] += strp[i];
  stcol.resize(strp[n]);
  {
    std::vector<int32_t> fill(strp.begin(), strp.end() - 1);
    for (int64_t i = 0; i < n; ++i)
      for (int32_t k = srp[i]; k < srp[i + 1]; ++k) stcol[fill[scol[k]]++] = (int32_t)i;
  }
  constexpr int32_t C_PT = 1, F_PT = -1, SF_PT = -3, UNDEC = 0;
  cf.assign(n, UNDEC);
  std::vector<int32_t> meas(n);
  int32_t maxm = 0;
  for (int64_t i = 0; i < n; ++i) {
    meas[i] = strp[i + 1] - strp[i];
    maxm = std::max(maxm, meas[i]);
  }
  // --- first pass ---
  for (int64_t i = 0; i < n; ++i)
    if (srp[i + 1] == srp[i]) {
      cf[i] = SF_PT;  // no strong connections: special F point
      meas[i] = 0;
    }
  // points with measure 0 become F and bump the measure of the points they depend on
  for (int64_t i = 0; i < n; ++i) {
    if (cf[i] != UNDEC || meas[i] != 0) continue;
    cf[i] = F_PT;
    for (int32_t k = srp[i]; k < srp[i + 1]; ++k) {
      const int32_t j = scol[k];
      if (cf[j] == UNDEC) {
        // only points not yet passed keep dynamic measures in hypre; all
        // undecided neighbours gain one
        meas[j]++;
      }
    }
  }
  Buckets bk(n, 2 * maxm + 2);
  for (int64_t i = 0; i < n; ++i)
    if (cf[i] == UNDEC && meas[i] > 0) bk.insert((int32_t)i, meas[i]);
    else if (cf[i] == UNDEC) cf[i] = F_PT;
  for (;;) {
    const int32_t idx = bk.pop_max();
    if (idx < 0) break;
    bk.remove(idx);
    cf[idx] = C_PT;
    meas[idx] = 0;
    // every undecided point that strongly depends on idx becomes F ...
    for (int32_t k = strp[idx]; k < strp[idx + 1]; ++k) {
      const int32_t j = stcol[k];
      if (cf[j] != UNDEC) continue;
      cf[j] = F_PT;
      bk.remove(j);
      meas[j] = 0;
      // ... and the undecided points j depends on become more attractive
      for (int32_t kk = srp[j]; kk < srp[j + 1]; ++kk) {
        const int32_t l = scol[kk];
        if (cf[l] != UNDEC) continue;
        bk.remove(l);
        meas[l]++;
        bk.insert(l, meas[l]);
      }
    }
    // the points idx depends on lose one potential dependant
    for (int32_t k = srp[idx]; k < srp[idx + 1]; ++k) {
      const int32_t j = scol[k];
      if (cf[j] != UNDEC) continue;
      bk.remove(j);
      meas[j]--;
      if (meas[j] <= 0) {
        cf[j] = F_PT;
        meas[j] = 0;
        for (int32_t kk = srp[j]; kk < srp[j + 1]; ++kk) {
          const int32_t l = scol[kk];
          if (cf[l] != UNDEC) continue;
          bk.remove(l);
          meas[l]++;
          bk.insert(l, meas[l]);
        }
      } else {
        bk.insert(j, meas[j]);
      }
    }
  }
  for (int64_t i = 0; i < n; ++i)
    if (cf[i] == UNDEC) cf[i] = F_PT;
  // --- second pass: strong F-F pairs need a common C point ---
  {
    std::vector<int32_t> stamp(n, -1);
    for (int64_t i = 0; i < n; ++i) {
      if (cf[i] != F_PT) continue;
      for (int32_t k = srp[i]; k < srp[i + 1]; ++k)
        if (cf[scol[k]] > 0) stamp[scol[k]] = (int32_t)i;
      int32_t tentative = -1;
      bool i_becomes_c = false;
      for (int32_t k = srp[i]; k < srp[i + 1]; ++k) {
        const int32_t j = scol[k];
        if (cf[j] != F_PT) continue;
        bool common = false;
        for (int32_t kk = srp[j]; kk < srp[j + 1]; ++kk)
          if (stamp[scol[kk]] == (int32_t)i) {
            common = true;
            break;
          }
        if (common) continue;
        if (tentative < 0) {
          tentative = j;
          cf[j] = C_PT;  // tentative
          stamp[j] = (int32_t)i;
        } else {
          i_becomes_c = true;
          break;
        }
      }
      if (i_becomes_c) {
        cf[tentative] = F_PT;  // undo the tentative change
        cf[i] = C_PT;
      }
    }
  }
  // --- CLJP stage with CF_init = 1 ---
  {
    std::vector<double> measure(n, 0.0);
    for (int64_t i = 0; i < n; ++i) measure[i] = double(strp[i + 1] - strp[i]);
    int64_t seed = kRandSeed;
    for (int64_t i = 0; i < n; ++i) {
      seed = rand_next(seed);
      measure[i] += double(seed) / double(kRandM);
    }
    // CF_init = 1: C points are final; for every C point the measures of the
    // points it depends on are decremented and its dependants that are not C
    // are fixed as F (they already interpolate from it).  What remains
    // undecided are F points without any strong C neighbour.
    std::vector<int32_t> graph;
    std::vector<int32_t> st(n, 0);  // 0 undecided, else decided
    for (int64_t i = 0; i < n; ++i) {
      if (cf[i] > 0 || cf[i] == SF_PT) {
        st[i] = 1;
        measure[i] = 0.0;
      }
    }
    for (int64_t i = 0; i < n; ++i) {
      if (st[i]) continue;
      bool has_c = false;
      for (int32_t k = srp[i]; k < srp[i + 1] && !has_c; ++k) has_c = cf[scol[k]] > 0;
      if (has_c) {
        st[i] = 1;
        measure[i] = 0.0;
      }
    }
    for (int64_t i = 0; i < n; ++i)
      if (!st[i]) {
        if (measure[i] < 1.0) {
          st[i] = 1;  // stays F
        } else {
          graph.push_back((int32_t)i);
        }
      }
    std::vector<int32_t> mark(n, 0);
    while (!graph.empty()) {
      for (int32_t i : graph) mark[i] = 1;
      for (int32_t i : graph)
        for (int32_t k = srp[i]; k < srp[i + 1]; ++k) {
          const int32_t j = scol[k];
          if (st[j]) continue;
          if (measure[i] > measure[j]) mark[j] = 0;
          else if (measure[j] > measure[i]) mark[i] = 0;
        }
      for (int32_t i : graph)
        if (mark[i]) {
          cf[i] = C_PT;
          st[i] = 1;
        }
      for (int32_t i : graph) {
        if (st[i]) continue;
        for (int32_t k = srp[i]; k < srp[i + 1]; ++k)
          if (cf[scol[k]] > 0) {
            st[i] = 1;  // F with a strong C neighbour
            break;
          }
      }
      size_t w = 0;
      for (int32_t i : graph)
        if (!st[i]) graph[w++] = i; else mark[i] = 0;
      graph.resize(w);
    }
  }
}

// ---------------------------------------------------------------------------
// A.3 CLJP (coarsen type 0): hypre_BoomerAMGCoarsen (par_coarsen.c) with CF_init = 0 on one
// rank, restated from memory like the rest of the AMG part (PARITY UNPINNED).  It is the
// parallel coarsening of the Falgout family (Falgout = Ruge first/second pass + this routine
// with CF_init = 1).
//   measure_i = |S^T_i| + hypre_Rand()  = number of PRESENT edges into i, plus the tie breaker.
//   Every round:
//   (1) undecided points with measure < 1 whose own dependency edges are all gone become F
//       ("make sure all dependencies have been accounted for"); decided points leave the graph;
//   (2) independent set among the points with measure > 1 (hypre_BoomerAMGIndepSet: along every
//       strong connection, removed or not, the smaller measure is knocked out) -> new C points;
//   (3) new C point i: every present edge i -> j is removed and undecided j loses one
//       ("C points don't interpolate from neighbours that influence them");
//   (4) undecided point i: edges to C points are removed; a present edge i -> j whose target j
//       depends (any edge, removed or not) on a C point that i depends on too is removed and j
//       loses one (i and j already share an interpolation point).
// Measures only ever change by exact decrements of 1, so the result does not depend on the
// order in which the points of a round are visited: the device kernels run (3) and (4) in
// parallel and give the same splitting.
// ---------------------------------------------------------------------------
void coarsen_cljp(int64_t n, const int32_t* rp, const int32_t* col, const uint8_t* mask,
                  std::vector<int32_t>& cf) {
  std::vector<double> measure(n, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
      if (mask[k]) measure[col[k]] += 1.0;
  int64_t seed = kRandSeed;
  for (int64_t i = 0; i < n; ++i) {
    seed = rand_next(seed);
    measure[i] += double(seed) / double(kRandM);
  }
  std::vector<uint8_t> edge(mask, mask + rp[n]);  // 1: strong connection still in the graph
  cf.assign(n, 0);
  // rows without strong connections are special F points from the start, as in the Ruge first
  // pass that precedes this routine under Falgout (and as PMIS does): they never coarsen, and
  // edges into them are dropped without touching any measure (the SF_PT branch of hypre's loop)
  std::vector<int32_t> graph;
  graph.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    bool any = false;
    for (int32_t k = rp[i]; k < rp[i + 1] && !any; ++k) any = mask[k];
    if (!any) {
      cf[i] = -3;
      measure[i] = 0.0;
    } else {
      graph.push_back((int32_t)i);
    }
  }
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
      if (edge[k] && cf[col[k]] == -3) edge[k] = 0;
  std::vector<int32_t> mark(n, 0);
  std::vector<uint8_t> common(n, 0);
  while (true) {
    // (1) F points, graph update
    size_t w = 0;
    for (int32_t i : graph) {
      if (cf[i] == 0 && measure[i] < 1.0) {
        bool open = false;
        for (int32_t k = rp[i]; k < rp[i + 1] && !open; ++k) open = edge[k] != 0;
        if (!open) cf[i] = -1;
      }
      if (cf[i] != 0) measure[i] = 0.0;
      else graph[w++] = i;
    }
    graph.resize(w);
    if (graph.empty()) break;
    // (2) independent set
    for (int32_t i : graph) mark[i] = measure[i] > 1.0 ? 1 : 0;
    for (int32_t i : graph) {
      if (!(measure[i] > 1.0)) continue;
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k) {
        if (!mask[k]) continue;
        const int32_t j = col[k];
        if (measure[j] > 1.0) {
          if (measure[i] > measure[j]) mark[j] = 0;
          else if (measure[j] > measure[i]) mark[i] = 0;
        }
      }
    }
    for (int32_t i : graph)
      if (mark[i]) cf[i] = 1;
    // (3), (4)
    for (int32_t i : graph) {
      if (cf[i] > 0) {
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
          if (edge[k]) {
            edge[k] = 0;
            if (cf[col[k]] == 0) measure[col[k]] -= 1.0;
          }
        continue;
      }
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
        if (mask[k] && cf[col[k]] > 0) {
          edge[k] = 0;
          common[col[k]] = 1;
        }
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k) {
        if (!edge[k]) continue;
        const int32_t j = col[k];
        for (int32_t k2 = rp[j]; k2 < rp[j + 1]; ++k2)
          if (mask[k2] && common[col[k2]]) {
            edge[k] = 0;
            measure[j] -= 1.0;
            break;
          }
      }
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
        if (mask[k]) common[col[k]] = 0;
    }
    for (int32_t i : graph) mark[i] = 0;
  }
}

// ---------------------------------------------------------------------------
// A.3 aggressive coarsening (levels < agg_nl), restated from memory like the rest of the
// AMG part (PARITY UNPINNED):
//  * hypre_BoomerAMGCreate2ndS with num_paths = 1 (PCHYPRE default agg_num_paths): for a
//    C point i, S2_i = { C points k != i : k in S_i, or k in S_j for some F point j in S_i }
//    (paths of length <= 2), built as the pattern product A1*B below, in coarse numbering;
//  * a second PMIS on S2 (random part restarted: coarse point c gets hypre_Rand #c);
//  * hypre_BoomerAMGCorrectCFMarker: first-stage C points take the second marker.
// ---------------------------------------------------------------------------
void spgemm(const Csr& A, const Csr& B, Csr& C);

void aggressive_second_pass(const Csr& A, const std::vector<uint8_t>& mask, std::vector<int32_t>& cf) {
  const int64_t n = A.n;
  std::vector<int32_t> f2c(n, -1);
  int32_t nc1 = 0;
  for (int64_t i = 0; i < n; ++i)
    if (cf[i] > 0) f2c[i] = nc1++;
  Csr A1, B, M;
  A1.n = nc1;
  A1.ncols = n;
  A1.rp.assign(1, 0);
  B.n = n;
  B.ncols = nc1;
  B.rp.assign(1, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (cf[i] > 0) {
      for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (mask[k]) {
          A1.col.push_back(A.col[k]);
          A1.val.push_back(1.0);
        }
      A1.rp.push_back((int32_t)A1.col.size());
      B.col.push_back(f2c[i]);
      B.val.push_back(1.0);
    } else {
      for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (mask[k] && cf[A.col[k]] > 0) {
          B.col.push_back(f2c[A.col[k]]);
          B.val.push_back(1.0);
        }
    }
    B.rp.push_back((int32_t)B.col.size());
  }
  spgemm(A1, B, M);
  std::vector<uint8_t> mask2(M.nnz());
  for (int64_t c = 0; c < nc1; ++c)
    for (int32_t k = M.rp[c]; k < M.rp[c + 1]; ++k) mask2[k] = M.col[k] != c;
  std::vector<int32_t> cf2;
  coarsen_pmis(nc1, M.rp.data(), M.col.data(), mask2.data(), cf2);
  int32_t cnt = 0;
  for (int64_t i = 0; i < n; ++i)
    if (cf[i] > 0) cf[i] = cf2[cnt++];
}

// C <- rows of X and Y (same shape; for every row at most one of them is non-empty)
void merge_disjoint_rows(const Csr& X, const Csr& Y, Csr& C) {
  C.n = X.n;
  C.ncols = X.ncols;
  C.rp.assign(1, 0);
  C.col.clear();
  C.val.clear();
  for (int64_t i = 0; i < X.n; ++i) {
    const Csr& S = X.rp[i + 1] > X.rp[i] ? X : Y;
    for (int32_t k = S.rp[i]; k < S.rp[i + 1]; ++k) {
      C.col.push_back(S.col[k]);
      C.val.push_back(S.val[k]);
    }
    C.rp.push_back((int32_t)C.col.size());
  }
}

// ---------------------------------------------------------------------------
// A.3 multipass interpolation (agg_interp_type 4, hypre_BoomerAMGBuildMultipass), restated
// from memory (PARITY UNPINNED).  pass(i) = 0 for C points; an F point joins pass k >= 1 if
// it has a strong neighbour of pass k-1.  For a point i of pass k, with interpolatory set
// I_i = { j in S_i : pass(j) = k-1 } and the sums over the negative / positive off-diagonal
// entries of row i (N: all, C: those in I_i):
//     positive (negative) entries are lumped into the diagonal if I_i has none of that sign;
//     alfa = (sum_N_neg / sum_C_neg) / diag,  beta = (sum_N_pos / sum_C_pos) / diag;
//     w_ij = -alfa a_ij (a_ij < 0) or -beta a_ij (a_ij > 0),  j in I_i;
//     P(i,:) = sum_j w_ij P(j,:)   (pass 1: P(j,:) is the unit row of the C point j),
// accumulated over ascending j.  Points no pass reaches keep an empty row.
// ---------------------------------------------------------------------------
void interp_multipass(const Csr& A, const std::vector<uint8_t>& mask, const std::vector<int32_t>& cf, Csr& P,
                      int64_t* n_coarse) {
  const int64_t n = A.n;
  std::vector<int32_t> f2c(n, -1), pass(n, -1);
  int32_t nc = 0;
  for (int64_t i = 0; i < n; ++i)
    if (cf[i] > 0) {
      f2c[i] = nc++;
      pass[i] = 0;
    }
  *n_coarse = nc;
  int npass = 0;
  for (int k = 1;; ++k) {
    int64_t newly = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (pass[i] >= 0) continue;
      for (int32_t e = A.rp[i]; e < A.rp[i + 1]; ++e)
        if (mask[e] && pass[A.col[e]] == k - 1) {
          pass[i] = k;
          ++newly;
          break;
        }
    }
    if (newly == 0) break;
    npass = k;
  }
  Csr cur;
  cur.n = n;
  cur.ncols = nc;
  cur.rp.assign(1, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (cf[i] > 0) {
      cur.col.push_back(f2c[i]);
      cur.val.push_back(1.0);
    }
    cur.rp.push_back((int32_t)cur.col.size());
  }
  for (int k = 1; k <= npass; ++k) {
    Csr W;
    W.n = n;
    W.ncols = k == 1 ? nc : n;
    W.rp.assign(1, 0);
    for (int64_t i = 0; i < n; ++i) {
      if (pass[i] == k) {
        double diag = 0.0, nneg = 0.0, npos = 0.0, cneg = 0.0, cpos = 0.0;
        for (int32_t e = A.rp[i]; e < A.rp[i + 1]; ++e) {
          const int32_t j = A.col[e];
          const double v = A.val[e];
          if (j == i) {
            diag = v;
            continue;
          }
          const bool in = mask[e] && pass[j] == k - 1;
          if (v < 0) {
            nneg += v;
            if (in) cneg += v;
          } else if (v > 0) {
            npos += v;
            if (in) cpos += v;
          }
        }
        if (cpos == 0) diag += npos;
        if (cneg == 0) diag += nneg;
        const double alfa = (cneg != 0 && diag != 0) ? (nneg / cneg) / diag : 0.0;
        const double beta = (cpos != 0 && diag != 0) ? (npos / cpos) / diag : 0.0;
        for (int32_t e = A.rp[i]; e < A.rp[i + 1]; ++e) {
          const int32_t j = A.col[e];
          if (j == i || !(mask[e] && pass[j] == k - 1)) continue;
          const double v = A.val[e];
          W.col.push_back(k == 1 ? f2c[j] : j);
          W.val.push_back(v < 0 ? -alfa * v : -beta * v);
        }
      }
      W.rp.push_back((int32_t)W.col.size());
    }
    Csr Pk, merged;
    if (k == 1) Pk = W;
    else spgemm(W, cur, Pk);
    merge_disjoint_rows(cur, Pk, merged);
    cur = merged;
  }
  P = cur;
}

// ---------------------------------------------------------------------------
// A.3 Interpolation type 0, "modified classical": hypre_BoomerAMGBuildInterp.
// Off-diagonal entries are visited in ascending column order.
// ---------------------------------------------------------------------------
void interp_classical(const Csr& A, const std::vector<uint8_t>& mask,
                      const std::vector<int32_t>& cf, Csr& P, int64_t* n_coarse) {
  const int64_t n = A.n;
  std::vector<int32_t> f2c(n, -1);
  int32_t nc = 0;
  for (int64_t i = 0; i < n; ++i)
    if (cf[i] > 0) f2c[i] = nc++;
  *n_coarse = nc;
  P.n = n;
  P.ncols = nc;
  P.rp.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    int32_t c = 0;
    if (cf[i] > 0) c = 1;
    else
      for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (mask[k] && cf[A.col[k]] > 0) ++c;
    P.rp[i + 1] = P.rp[i] + c;
  }
  P.col.assign(P.rp[n], 0);
  P.val.assign(P.rp[n], 0.0);
  constexpr int32_t kStrongF = -2;
  std::vector<int32_t> marker(n, -1);
  for (int64_t i = 0; i < n; ++i) {
    const int32_t jb = P.rp[i];
    if (cf[i] > 0) {
      P.col[jb] = f2c[i];
      P.val[jb] = 1.0;
      continue;
    }
    int32_t jj = jb;
    for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      if (!mask[k]) continue;
      const int32_t i1 = A.col[k];
      if (cf[i1] > 0) {
        marker[i1] = jj;
        P.col[jj] = f2c[i1];
        P.val[jj] = 0.0;
        ++jj;
      } else if (cf[i1] != -3) {
        marker[i1] = kStrongF;
      }
    }
    double diagonal = 0.0;
    for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
      if (A.col[k] == i) diagonal = A.val[k];
    for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      const int32_t i1 = A.col[k];
      if (i1 == i) continue;
      if (marker[i1] >= jb) {
        P.val[marker[i1]] += A.val[k];
      } else if (marker[i1] == kStrongF) {
        double sum = 0.0;
        double dk = 0.0;
        for (int32_t k1 = A.rp[i1]; k1 < A.rp[i1 + 1]; ++k1)
          if (A.col[k1] == i1) dk = A.val[k1];
        const double sgn = dk < 0 ? -1.0 : 1.0;
        for (int32_t k1 = A.rp[i1]; k1 < A.rp[i1 + 1]; ++k1) {
          const int32_t i2 = A.col[k1];
          if (marker[i2] >= jb && sgn * A.val[k1] < 0) sum += A.val[k1];
        }
        if (sum != 0) {
          const double distribute = A.val[k] / sum;
          for (int32_t k1 = A.rp[i1]; k1 < A.rp[i1 + 1]; ++k1) {
            const int32_t i2 = A.col[k1];
            if (marker[i2] >= jb && sgn * A.val[k1] < 0)
              P.val[marker[i2]] += distribute * A.val[k1];
          }
        } else {
          diagonal += A.val[k];
        }
      } else if (cf[i1] != -3) {
        diagonal += A.val[k];
      }
    }
    if (diagonal == 0.0) {
      for (int32_t k = jb; k < jj; ++k) P.val[k] = 0.0;
    } else {
      for (int32_t k = jb; k < jj; ++k) P.val[k] /= -diagonal;
    }
    // reset markers touched by this row
    for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k) marker[A.col[k]] = -1;
  }
}

void transpose(const Csr& P, Csr& R) {
  R.n = P.ncols;
  R.ncols = P.n;
  R.rp.assign(R.n + 1, 0);
  for (int64_t k = 0; k < P.nnz(); ++k) R.rp[P.col[k] + 1]++;
  for (int64_t i = 0; i < R.n; ++i) R.rp[i + 1] += R.rp[i];
  R.col.resize(P.nnz());
  R.val.resize(P.nnz());
  std::vector<int32_t> fill(R.rp.begin(), R.rp.end() - 1);
  for (int64_t i = 0; i < P.n; ++i)
    for (int32_t k = P.rp[i]; k < P.rp[i + 1]; ++k) {
      const int32_t w = fill[P.col[k]]++;
      R.col[w] = (int32_t)i;
      R.val[w] = P.val[k];
    }
}

// C = A*B, Gustavson, output columns sorted ascending.  Each C(i,c) is the sum
// of a_ik*b_kc in ascending k, starting from 0.0 (two roundings per term).
void spgemm(const Csr& A, const Csr& B, Csr& C) {
  C.n = A.n;
  C.ncols = B.ncols;
  C.rp.assign(A.n + 1, 0);
  std::vector<int32_t> marker(B.ncols, -1);
  std::vector<double> acc(B.ncols, 0.0);
  std::vector<int32_t> cols;
  C.col.clear();
  C.val.clear();
  for (int64_t i = 0; i < A.n; ++i) {
    cols.clear();
    for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      const int32_t kk = A.col[k];
      const double a = A.val[k];
      for (int32_t m = B.rp[kk]; m < B.rp[kk + 1]; ++m) {
        const int32_t c = B.col[m];
        if (marker[c] != (int32_t)i) {
          marker[c] = (int32_t)i;
          acc[c] = 0.0;
          cols.push_back(c);
        }
        acc[c] += a * B.val[m];
      }
    }
    std::sort(cols.begin(), cols.end());
    for (int32_t c : cols) {
      C.col.push_back(c);
      C.val.push_back(acc[c]);
    }
    C.rp[i + 1] = (int32_t)C.col.size();
  }
}

void spmv(const Csr& A, const double* x, double* y) {
  for (int64_t i = 0; i < A.n; ++i) {
    double s = 0.0;
    for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k) s += A.val[k] * x[A.col[k]];
    y[i] = s;
  }
}

// deal.II RelaxationType -> hypre relax type (Appendix A.1)
int hypre_relax_type(int dealii_type, bool symmetric_operator) {
  switch (dealii_type) {
    case AMGB_RELAX_Jacobi: return 0;
    case AMGB_RELAX_sequentialGaussSeidel: return 1;
    case AMGB_RELAX_seqboundaryGaussSeidel: return 2;
    case AMGB_RELAX_SORJacobi: return symmetric_operator ? 6 : 3;
    case AMGB_RELAX_backwardSORJacobi: return 4;
    case AMGB_RELAX_symmetricSORJacobi: return 6;
    case AMGB_RELAX_l1scaledSORJacobi: return 8;
    case AMGB_RELAX_GaussianElimination: return 9;
    case AMGB_RELAX_l1GaussSeidel: return 13;
    case AMGB_RELAX_backwardl1GaussSeidel: return 14;
    case AMGB_RELAX_CG: return 15;
    case AMGB_RELAX_Chebyshev: return 16;
    case AMGB_RELAX_FCFJacobi: return 17;
    case AMGB_RELAX_l1scaledJacobi: return 18;
    default: return -1;
  }
}

struct Level {
  Csr A, P, R;
  std::vector<uint8_t> mask;
  std::vector<int32_t> cf;       // as returned by the coarsening (+1,-1,-3)
  std::vector<int32_t> cf_relax; // -3 folded into -1 (end of BuildInterp)
  std::vector<int32_t> color;    // multicolour Gauss-Seidel: colour of every point
  int32_t ncolors = 0;
  std::vector<double> diag, l1;
  std::vector<double> u, f, tmp; // work vectors
  // Chebyshev smoother (relax type 16): 1/sqrt(diag), spectrum estimates, polynomial
  std::vector<double> cheby_ds;
  double cheby_max_eig = 0.0, cheby_min_eig = 0.0, cheby_coefs[5] = {0, 0, 0, 0, 0};
  int cheby_order = 0;  // degree of p in u += p(A) r (hypre's "order" minus one)
};

}  // namespace

struct orc_hier {
  amgb_boomeramg_data data;
  double theta_eff, mrs_eff;
  int relax_down, relax_up, relax_coarse;
  std::vector<Level> lv;
  std::vector<double> dense;  // coarsest matrix, row-major, LU in place (no pivoting)
  bool dense_ok = false;
};

namespace {

constexpr int64_t kMaxDenseCoarse = 1024;

void level_aux(Level& L) {
  const int64_t n = L.A.n;
  L.diag.assign(n, 0.0);
  L.l1.assign(n, 0.0);
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int32_t k = L.A.rp[i]; k < L.A.rp[i + 1]; ++k) {
      if (L.A.col[k] == i) L.diag[i] = L.A.val[k];
      s += std::fabs(L.A.val[k]);
    }
    L.l1[i] = s;
  }
  L.u.assign(n, 0.0);
  L.f.assign(n, 0.0);
  L.tmp.assign(n, 0.0);
}

// ---------------------------------------------------------------------------
// Chebyshev smoother, hypre relax type 16 (par_cheby.c; PCHYPRE defaults: order 2,
// eigenvalue estimate by 10 CG iterations, fraction 0.3, variant 0, diagonal scaling).
// ---------------------------------------------------------------------------
// EISPACK tql1: eigenvalues of a symmetric tridiagonal matrix, ascending in d.  e holds the
// sub-diagonal in e[1..n) (e[0] arbitrary), as hypre_LINPACKcgtql1 takes it.
double pythag(double a, double b) {
  const double p = std::max(std::fabs(a), std::fabs(b));
  if (p == 0.0) return 0.0;
  const double q = std::min(std::fabs(a), std::fabs(b)) / p;
  double r = q * q;
  double pp = p;
  for (;;) {
    const double t = 4.0 + r;
    if (t == 4.0) break;
    const double sq = r / t;
    const double u = 1.0 + 2.0 * sq;
    pp = u * pp;
    r = (sq / u) * (sq / u) * r;
  }
  return pp;
}

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
      if (tst1 + std::fabs(e[m]) == tst1) break;  // e[n-1] = 0 ends the search
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

// Inner product in the order of the device kernel (so that the spectrum estimates are the
// same bits on both sides): blocks of 256 products, each warp of 32 by the shuffle-down tree
// (16, 8, 4, 2, 1), the 8 warp sums left to right, the block sums left to right.
double tree_dot(const double* x, const double* y, int64_t n) {
  double total = 0.0;
  for (int64_t b0 = 0; b0 < n; b0 += 256) {
    double block = 0.0;
    for (int w = 0; w < 8; ++w) {
      double v[32];
      for (int l = 0; l < 32; ++l) {
        const int64_t i = b0 + 32 * w + l;
        v[l] = i < n ? x[i] * y[i] : 0.0;
      }
      for (int off = 16; off > 0; off >>= 1)
        for (int l = 0; l < off; ++l) v[l] = v[l] + v[l + off];
      block = block + v[0];
    }
    total = total + block;
  }
  return total;
}

// hypre_ParCSRMaxEigEstimateCG with scale = 1: Lanczos through `max_iter` CG steps on
// D^{-1/2} A D^{-1/2} from the random vector of hypre_ParVectorSetRandomValues(r, 1).
void cheby_eig_estimate(const Csr& A, const std::vector<double>& ds, int max_iter, double* max_eig,
                        double* min_eig) {
  const int64_t n = A.n;
  if (n < max_iter) max_iter = (int)n;
  std::vector<double> r(n), p(n), s(n), u(n), tri(max_iter + 2, 0.0), off(max_iter + 2, 0.0);
  int64_t seed = 1;
  for (int64_t i = 0; i < n; ++i) {
    seed = rand_next(seed);
    r[i] = 2.0 * ((double)seed / (double)kRandM) - 1.0;
  }
  double gamma = 1.0;
  int it = 0;
  while (it < max_iter) {
    const double gamma_old = gamma;
    gamma = tree_dot(r.data(), r.data(), n);
    if (!(gamma > 0.0)) break;
    double beta = 1.0;
    if (it == 0) {
      p = r;
    } else {
      beta = gamma / gamma_old;
      for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
    }
    for (int64_t i = 0; i < n; ++i) u[i] = ds[i] * p[i];
    spmv(A, u.data(), s.data());
    for (int64_t i = 0; i < n; ++i) s[i] = ds[i] * s[i];
    const double sdotp = tree_dot(s.data(), p.data(), n);
    if (!(sdotp > 0.0)) break;
    const double alpha = gamma / sdotp;
    const double alphainv = 1.0 / alpha;
    tri[it + 1] = alphainv;
    tri[it] *= beta;
    tri[it] += alphainv;
    off[it + 1] = alphainv;
    off[it] *= std::sqrt(beta);
    for (int64_t i = 0; i < n; ++i) r[i] = r[i] - alpha * s[i];
    ++it;
  }
  if (it == 0) {
    *max_eig = *min_eig = 1.0;
    return;
  }
  tql1(it, tri.data(), off.data());
  *max_eig = tri[it - 1];
  *min_eig = tri[0];
}

// hypre_ParCSRRelax_Cheby_Setup, variant 0: coefficients of p in u += p(A) r.
void cheby_coefficients(double max_eig, double min_eig, double fraction, int order, double* coefs,
                        int* cheby_order) {
  if (order > 4) order = 4;
  if (order < 1) order = 1;
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
  *cheby_order = k;
}

void cheby_setup(Level& L, int order, int eig_est, double fraction) {
  const int64_t n = L.A.n;
  L.cheby_ds.assign(n, 0.0);
  for (int64_t i = 0; i < n; ++i) L.cheby_ds[i] = 1.0 / std::sqrt(L.diag[i]);
  cheby_eig_estimate(L.A, L.cheby_ds, eig_est, &L.cheby_max_eig, &L.cheby_min_eig);
  cheby_coefficients(L.cheby_max_eig, L.cheby_min_eig, fraction, order, L.cheby_coefs, &L.cheby_order);
}

// hypre_ParCSRRelax_Cheby_Solve with scaling: u += D^{-1/2} p(D^{-1/2} A D^{-1/2}) D^{-1/2} (f - A u)
void cheby_relax(const Level& L, const double* f, double* u, double* tmp) {
  const Csr& A = L.A;
  const int64_t n = A.n;
  const double* ds = L.cheby_ds.data();
  std::vector<double> r(n), q(n), t(n);
  spmv(A, u, tmp);
  for (int64_t i = 0; i < n; ++i) r[i] = ds[i] * (f[i] - tmp[i]);
  const int k = L.cheby_order;
  for (int64_t i = 0; i < n; ++i) q[i] = L.cheby_coefs[k] * r[i];
  for (int c = k - 1; c >= 0; --c) {
    for (int64_t i = 0; i < n; ++i) t[i] = ds[i] * q[i];
    spmv(A, t.data(), tmp);
    for (int64_t i = 0; i < n; ++i) q[i] = L.cheby_coefs[c] * r[i] + ds[i] * tmp[i];
  }
  for (int64_t i = 0; i < n; ++i) u[i] = u[i] + ds[i] * q[i];
}

// Greedy multicolouring (AMGB_SMOOTHER_MULTICOLOR; not part of hypre: the device's parallel order for
// the reference's Gauss-Seidel sweeps).  Rounds: every uncoloured point whose priority is the largest
// among its uncoloured neighbours takes the smallest colour none of its (already coloured) neighbours
// has.  The priority is a hash of the index, ties go to the larger index; two neighbours are never
// selected in the same round, so the result does not depend on the order of evaluation.  Neighbours =
// the columns of the row (the operator is assumed structurally symmetric).
inline uint32_t color_priority(int64_t i) {
  uint32_t h = (uint32_t)(i + 1) * 2654435761u;
  h ^= h >> 15;
  h *= 2246822519u;
  h ^= h >> 13;
  return h;
}

int multicolor(const Csr& A, std::vector<int32_t>& color, int32_t* ncolors) {
  const int64_t n = A.n;
  color.assign(n, -1);
  std::vector<int32_t> next(n, -1);
  int64_t left = n;
  int32_t nc = 0;
  while (left > 0) {
    for (int64_t i = 0; i < n; ++i) {
      if (color[i] >= 0) continue;
      const uint32_t pi = color_priority(i);
      bool is_max = true;
      uint64_t used = 0;
      for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k) {
        const int64_t j = A.col[k];
        if (j == i) continue;
        if (color[j] >= 0) {
          used |= 1ull << color[j];
        } else {
          const uint32_t pj = color_priority(j);
          if (pj > pi || (pj == pi && j > i)) is_max = false;
        }
      }
      if (is_max) {
        int32_t c = 0;
        while (c < 64 && ((used >> c) & 1ull)) ++c;
        if (c >= 64) return AMGB_ERR_RANGE;
        next[i] = c;
      }
    }
    for (int64_t i = 0; i < n; ++i)
      if (color[i] < 0 && next[i] >= 0) {
        color[i] = next[i];
        nc = std::max(nc, next[i] + 1);
        --left;
      }
  }
  // a row must not read a point of its own colour: holds by construction for a structurally symmetric pattern
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
      if (A.col[k] != i && color[A.col[k]] == color[i]) return AMGB_ERR_UNSUPPORTED;
  *ncolors = nc;
  return 0;
}

// Gauss-Seidel in multicolour order: colours ascending (forward), descending (backward) or both
// (symmetric); inside a colour the points are independent.
void relax_multicolor(const Level& L, int type, const double* f, double* u) {
  const Csr& A = L.A;
  auto sweep = [&](int32_t c) {
    for (int64_t i = 0; i < A.n; ++i) {
      if (L.color[i] != c || L.diag[i] == 0.0) continue;
      double res = f[i];
      for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (A.col[k] != i) res -= A.val[k] * u[A.col[k]];
      u[i] = res / L.diag[i];
    }
  };
  if (type != AMGB_RELAX_MC_BACKWARD)
    for (int32_t c = 0; c < L.ncolors; ++c) sweep(c);
  if (type != AMGB_RELAX_MC_FORWARD)
    for (int32_t c = L.ncolors - 1; c >= 0; --c) sweep(c);
}

// One hypre_BoomerAMGRelax call: relax_points 0 = all, else only cf == relax_points.
void relax(const Level& L, int type, int relax_points, double w, const double* f, double* u,
           double* tmp) {
  const Csr& A = L.A;
  const int64_t n = A.n;
  auto want = [&](int64_t i) { return relax_points == 0 || L.cf_relax[i] == relax_points; };
  switch (type) {
    case 0: {  // weighted Jacobi (par_relax.c case 0)
      std::copy(u, u + n, tmp);
      for (int64_t i = 0; i < n; ++i) {
        if (!want(i) || L.diag[i] == 0.0) continue;
        double res = f[i];
        for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
          if (A.col[k] != i) res -= A.val[k] * tmp[A.col[k]];
        u[i] = (1.0 - w) * tmp[i] + w * res / L.diag[i];
      }
      break;
    }
    case 18: {  // l1-scaled Jacobi: u += w (f - A u)/||a_i||_1
      std::copy(u, u + n, tmp);
      for (int64_t i = 0; i < n; ++i) {
        if (!want(i) || L.diag[i] == 0.0) continue;
        double res = f[i];
        for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k) res -= A.val[k] * tmp[A.col[k]];
        u[i] += (w * res) / L.l1[i];
      }
      break;
    }
    case 3:    // hybrid forward GS (1 rank, 1 thread: true Gauss-Seidel)
    case 4:    // backward
    case 6: {  // symmetric: forward then backward
      if (type != 4)
        for (int64_t i = 0; i < n; ++i) {
          if (!want(i) || L.diag[i] == 0.0) continue;
          double res = f[i];
          for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
            if (A.col[k] != i) res -= A.val[k] * u[A.col[k]];
          u[i] = res / L.diag[i];
        }
      if (type != 3)
        for (int64_t i = n - 1; i >= 0; --i) {
          if (!want(i) || L.diag[i] == 0.0) continue;
          double res = f[i];
          for (int32_t k = A.rp[i]; k < A.rp[i + 1]; ++k)
            if (A.col[k] != i) res -= A.val[k] * u[A.col[k]];
          u[i] = res / L.diag[i];
        }
      break;
    }
    default:
      break;
  }
}

// hypre_BoomerAMGRelaxIF: C/F ordering when relax_order == 1.
void relax_if(const orc_hier& h, const Level& L, int type, int cycle_param, const double* f,
              double* u, double* tmp) {
  const double w = h.data.relax_weight;
  if (type == 16) {  // par_cycle.c: the polynomial smoother runs on all points, no C/F ordering
    cheby_relax(L, f, u, tmp);
    return;
  }
  if (type >= AMGB_RELAX_MC_FORWARD) {  // the colours are the ordering: no C/F ordering on top
    relax_multicolor(L, type, f, u);
    return;
  }
  if (h.data.relax_order == 1 && cycle_param < 3) {
    const int pts[2] = {cycle_param < 2 ? 1 : -1, cycle_param < 2 ? -1 : 1};
    relax(L, type, pts[0], w, f, u, tmp);
    relax(L, type, pts[1], w, f, u, tmp);
  } else {
    relax(L, type, 0, w, f, u, tmp);
  }
}

// hypre_gselim: Gaussian elimination without pivoting, split into factor/solve.
bool dense_factor(std::vector<double>& M, int64_t n) {
  for (int64_t k = 0; k + 1 < n; ++k) {
    if (M[k * n + k] == 0.0) continue;
    for (int64_t j = k + 1; j < n; ++j) {
      if (M[j * n + k] == 0.0) continue;
      const double factor = M[j * n + k] / M[k * n + k];
      for (int64_t m = k + 1; m < n; ++m) M[j * n + m] -= factor * M[k * n + m];
      M[j * n + k] = factor;  // keep the multiplier
    }
  }
  return true;
}

void dense_solve(const std::vector<double>& M, int64_t n, double* x) {
  for (int64_t k = 0; k + 1 < n; ++k) {
    if (M[k * n + k] == 0.0) continue;
    for (int64_t j = k + 1; j < n; ++j)
      if (M[j * n + k] != 0.0) x[j] -= M[j * n + k] * x[k];
  }
  for (int64_t k = n - 1; k > 0; --k) {
    if (M[k * n + k] != 0.0) {
      x[k] /= M[k * n + k];
      for (int64_t j = 0; j < k; ++j)
        if (M[j * n + k] != 0.0) x[j] -= x[k] * M[j * n + k];
    }
  }
  if (n > 0 && M[0] != 0.0) x[0] /= M[0];
}

void cycle(orc_hier& h, int l) {
  Level& L = h.lv[l];
  const int nl = (int)h.lv.size();
  if (l == nl - 1) {
    if (h.relax_coarse == 9 && h.dense_ok) {
      std::copy(L.f.begin(), L.f.end(), L.u.begin());
      dense_solve(h.dense, L.A.n, L.u.data());
    } else {
      const int t = h.relax_coarse == 9 ? h.relax_down : h.relax_coarse;
      for (unsigned s = 0; s < std::max(1u, h.data.n_sweeps_coarse); ++s)
        relax_if(h, L, t, 3, L.f.data(), L.u.data(), L.tmp.data());
    }
    return;
  }
  for (unsigned s = 0; s < h.data.n_sweeps; ++s)
    relax_if(h, L, h.relax_down, 1, L.f.data(), L.u.data(), L.tmp.data());
  // residual, restriction
  spmv(L.A, L.u.data(), L.tmp.data());
  for (int64_t i = 0; i < L.A.n; ++i) L.tmp[i] = L.f[i] - L.tmp[i];
  Level& C = h.lv[l + 1];
  spmv(L.R, L.tmp.data(), C.f.data());
  std::fill(C.u.begin(), C.u.end(), 0.0);
  cycle(h, l + 1);
  if (h.data.w_cycle && l + 1 < nl - 1) cycle(h, l + 1);
  // prolongation + correction
  for (int64_t i = 0; i < L.A.n; ++i) {
    double s = 0.0;
    for (int32_t k = L.P.rp[i]; k < L.P.rp[i + 1]; ++k) s += L.P.val[k] * C.u[L.P.col[k]];
    L.u[i] += s;
  }
  for (unsigned s = 0; s < h.data.n_sweeps; ++s)
    relax_if(h, L, h.relax_up, 2, L.f.data(), L.u.data(), L.tmp.data());
}

}  // namespace

extern "C" {

double orc_option_roundtrip(double v) { return std::strtod(std::to_string(v).c_str(), nullptr); }

double orc_hypre_rand(int64_t i) {
  int64_t seed = kRandSeed;
  for (int64_t k = 0; k <= i; ++k) seed = rand_next(seed);
  return double(seed) / double(kRandM);
}

int orc_strength(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
                 double theta, double max_row_sum, uint8_t* mask) {
  Csr A;
  A.n = A.ncols = n;
  A.rp.assign(rowptr, rowptr + n + 1);
  A.col.assign(col, col + rowptr[n]);
  A.val.assign(val, val + rowptr[n]);
  std::vector<uint8_t> m;
  strength(A, theta, max_row_sum, m);
  std::memcpy(mask, m.data(), m.size());
  return 0;
}

int orc_coarsen_pmis(int64_t n, const int32_t* rowptr, const int32_t* col, const uint8_t* mask,
                     int32_t* cf) {
  std::vector<int32_t> c;
  coarsen_pmis(n, rowptr, col, mask, c);
  std::memcpy(cf, c.data(), n * sizeof(int32_t));
  return 0;
}

int orc_coarsen_cljp(int64_t n, const int32_t* rowptr, const int32_t* col, const uint8_t* mask,
                     int32_t* cf) {
  std::vector<int32_t> c;
  coarsen_cljp(n, rowptr, col, mask, c);
  std::memcpy(cf, c.data(), n * sizeof(int32_t));
  return 0;
}

int orc_coarsen_falgout(int64_t n, const int32_t* rowptr, const int32_t* col,
                        const uint8_t* mask, int32_t* cf) {
  std::vector<int32_t> c;
  coarsen_falgout(n, rowptr, col, mask, c);
  std::memcpy(cf, c.data(), n * sizeof(int32_t));
  return 0;
}

int orc_setup(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
              const amgb_boomeramg_data* data, orc_hier** out) {
  if (!rowptr || !col || !val || !data || !out || n < 1) return AMGB_ERR_BAD_ARG;
  // aggressive levels: restated for PMIS only (second PMIS on S2 + multipass interpolation)
  if (data->aggressive_coarsening_num_levels != 0 && data->coarsen_type != AMGB_COARSEN_PMIS)
    return AMGB_ERR_UNSUPPORTED;
  if (data->interp_type != AMGB_INTERP_CLASSICAL) return AMGB_ERR_UNSUPPORTED;
  orc_hier* h = new orc_hier;
  h->data = *data;
  h->theta_eff = data->options_via_string ? orc_option_roundtrip(data->strong_threshold)
                                          : data->strong_threshold;
  h->mrs_eff = data->options_via_string ? orc_option_roundtrip(data->max_row_sum)
                                        : data->max_row_sum;
  const bool sym = data->symmetric_operator != 0;
  h->relax_down = hypre_relax_type(data->relaxation_type_down, sym);
  h->relax_up = hypre_relax_type(data->relaxation_type_up, sym);
  h->relax_coarse = hypre_relax_type(data->relaxation_type_coarse, sym);
  if (data->smoother_policy == AMGB_SMOOTHER_MULTICOLOR) {
    for (int* t : {&h->relax_down, &h->relax_up, &h->relax_coarse})
      if (*t == 3 || *t == 4 || *t == 6) *t += 100;
  }
  auto supported = [](int t) { return t == 0 || t == 3 || t == 4 || t == 6 || t == 18 || t == 103 || t == 104 || t == 106; };
  if (h->relax_down == 16 && h->relax_up == 16 && (supported(h->relax_coarse) || h->relax_coarse == 9)) {
    // Chebyshev on the way down and up
  } else if (!supported(h->relax_down) || !supported(h->relax_up) ||
      !(supported(h->relax_coarse) || h->relax_coarse == 9)) {
    delete h;
    return AMGB_ERR_UNSUPPORTED;
  }
  h->lv.emplace_back();
  {
    Csr& A = h->lv[0].A;
    A.n = A.ncols = n;
    A.rp.assign(rowptr, rowptr + n + 1);
    A.col.assign(col, col + rowptr[n]);
    A.val.assign(val, val + rowptr[n]);
  }
  const int max_levels = std::max(1, data->max_levels);
  for (int level = 0;; ++level) {
    Level& L = h->lv[level];
    if (level == max_levels - 1 || L.A.n <= data->max_coarse_size) break;
    strength(L.A, h->theta_eff, h->mrs_eff, L.mask);
    if (data->coarsen_type == AMGB_COARSEN_PMIS)
      coarsen_pmis(L.A.n, L.A.rp.data(), L.A.col.data(), L.mask.data(), L.cf);
    else if (data->coarsen_type == AMGB_COARSEN_FALGOUT)
      coarsen_falgout(L.A.n, L.A.rp.data(), L.A.col.data(), L.mask.data(), L.cf);
    else if (data->coarsen_type == AMGB_COARSEN_CLJP)
      coarsen_cljp(L.A.n, L.A.rp.data(), L.A.col.data(), L.mask.data(), L.cf);
    else {
      delete h;
      return AMGB_ERR_UNSUPPORTED;
    }
    const bool aggressive = (unsigned)level < data->aggressive_coarsening_num_levels;
    int64_t nc = 0;
    for (int32_t c : L.cf) nc += c > 0;
    if (aggressive && nc > 0 && nc < L.A.n) {
      aggressive_second_pass(L.A, L.mask, L.cf);
      nc = 0;
      for (int32_t c : L.cf) nc += c > 0;
    }
    if (nc == 0 || nc == L.A.n) {
      // coarsening stalled: this level is the coarsest
      L.mask.clear();
      L.cf.clear();
      break;
    }
    if (aggressive) interp_multipass(L.A, L.mask, L.cf, L.P, &nc);
    else interp_classical(L.A, L.mask, L.cf, L.P, &nc);
    transpose(L.P, L.R);
    Csr T;
    spgemm(L.A, L.P, T);
    h->lv.emplace_back();
    spgemm(h->lv[level].R, T, h->lv[level + 1].A);
  }
  for (Level& L : h->lv) {
    L.cf_relax = L.cf;
    for (int32_t& c : L.cf_relax)
      if (c == -3) c = -1;
    if (L.cf_relax.empty()) L.cf_relax.assign(L.A.n, 0);
    level_aux(L);
  }
  if (h->relax_down >= AMGB_RELAX_MC_FORWARD || h->relax_up >= AMGB_RELAX_MC_FORWARD ||
      h->relax_coarse >= AMGB_RELAX_MC_FORWARD)
    for (Level& L : h->lv)
      if (const int rc = multicolor(L.A, L.color, &L.ncolors)) {
        delete h;
        return rc;
      }
  if (h->relax_down == 16)  // PCHYPRE / hypre defaults: order 2, 10 CG steps, fraction 0.3
    for (Level& L : h->lv) cheby_setup(L, 2, 10, 0.3);
  Level& C = h->lv.back();
  if (h->relax_coarse == 9 && C.A.n <= kMaxDenseCoarse) {
    const int64_t nc = C.A.n;
    h->dense.assign(nc * nc, 0.0);
    for (int64_t i = 0; i < nc; ++i)
      for (int32_t k = C.A.rp[i]; k < C.A.rp[i + 1]; ++k) h->dense[i * nc + C.A.col[k]] = C.A.val[k];
    h->dense_ok = dense_factor(h->dense, nc);
  }
  *out = h;
  return 0;
}

void orc_destroy(orc_hier* h) { delete h; }

int orc_num_levels(const orc_hier* h) { return h ? (int)h->lv.size() : 0; }

int orc_level_dims(const orc_hier* h, int level, int64_t* n, int64_t* nnz_A, int64_t* n_coarse,
                   int64_t* nnz_P) {
  if (!h || level < 0 || level >= (int)h->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = h->lv[level];
  if (n) *n = L.A.n;
  if (nnz_A) *nnz_A = L.A.nnz();
  if (n_coarse) *n_coarse = L.P.ncols;
  if (nnz_P) *nnz_P = L.P.nnz();
  return 0;
}

int orc_get_strength_mask(const orc_hier* h, int level, uint8_t* mask) {
  if (!h || level < 0 || level >= (int)h->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = h->lv[level];
  if (L.mask.empty()) return AMGB_ERR_RANGE;
  std::memcpy(mask, L.mask.data(), L.mask.size());
  return 0;
}

int orc_get_cf_marker(const orc_hier* h, int level, int32_t* cf) {
  if (!h || level < 0 || level >= (int)h->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = h->lv[level];
  if (L.cf.empty()) return AMGB_ERR_RANGE;
  std::memcpy(cf, L.cf.data(), L.cf.size() * sizeof(int32_t));
  return 0;
}

static int copy_csr(const Csr& M, int32_t* rowptr, int32_t* col, double* val) {
  if (rowptr) std::memcpy(rowptr, M.rp.data(), M.rp.size() * sizeof(int32_t));
  if (col) std::memcpy(col, M.col.data(), M.col.size() * sizeof(int32_t));
  if (val) std::memcpy(val, M.val.data(), M.val.size() * sizeof(double));
  return 0;
}

int orc_get_A_csr(const orc_hier* h, int level, int32_t* rowptr, int32_t* col, double* val) {
  if (!h || level < 0 || level >= (int)h->lv.size()) return AMGB_ERR_RANGE;
  return copy_csr(h->lv[level].A, rowptr, col, val);
}

int orc_get_P_csr(const orc_hier* h, int level, int32_t* rowptr, int32_t* col, double* val) {
  if (!h || level < 0 || level + 1 >= (int)h->lv.size()) return AMGB_ERR_RANGE;
  return copy_csr(h->lv[level].P, rowptr, col, val);
}

int orc_level_stats(const orc_hier* h, int capacity, int32_t* n_levels, int64_t* rows,
                    int64_t* nnz, double* sparsity, double* grid_cx, double* op_cx,
                    double* mem_cx) {
  if (!h) return AMGB_ERR_BAD_ARG;
  const int nl = (int)h->lv.size();
  if (n_levels) *n_levels = nl;
  if (capacity < nl) return AMGB_ERR_RANGE;
  double sr = 0, sa = 0, sp = 0;
  for (int l = 0; l < nl; ++l) {
    const Level& L = h->lv[l];
    if (rows) rows[l] = L.A.n;
    if (nnz) nnz[l] = L.A.nnz();
    if (sparsity) sparsity[l] = double(L.A.nnz()) / (double(L.A.n) * double(L.A.n));
    sr += double(L.A.n);
    sa += double(L.A.nnz());
    sp += double(L.P.nnz());
  }
  if (grid_cx) *grid_cx = sr / double(h->lv[0].A.n);
  if (op_cx) *op_cx = sa / double(h->lv[0].A.nnz());
  if (mem_cx) *mem_cx = (sa + sp) / double(h->lv[0].A.nnz());
  return 0;
}

int orc_get_colors(const orc_hier* h, int level, int32_t* colors, int32_t* n_colors) {
  if (!h || level < 0 || level >= (int)h->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = h->lv[level];
  if (L.color.empty()) return AMGB_ERR_RANGE;
  if (colors) std::copy(L.color.begin(), L.color.end(), colors);
  if (n_colors) *n_colors = L.ncolors;
  return 0;
}

int orc_effective_relax(const orc_hier* h, int32_t* down, int32_t* up, int32_t* coarse) {
  if (!h) return AMGB_ERR_BAD_ARG;
  if (down) *down = h->relax_down;
  if (up) *up = h->relax_up;
  if (coarse) *coarse = h->relax_coarse;
  return 0;
}

int orc_level_cheby(const orc_hier* h, int level, double* max_eig, double* min_eig, double* coefs,
                    int32_t* n_coefs) {
  if (!h || level < 0 || level >= (int)h->lv.size()) return AMGB_ERR_RANGE;
  const Level& L = h->lv[level];
  if (L.cheby_ds.empty()) return AMGB_ERR_RANGE;
  if (max_eig) *max_eig = L.cheby_max_eig;
  if (min_eig) *min_eig = L.cheby_min_eig;
  if (coefs)
    for (int i = 0; i <= L.cheby_order; ++i) coefs[i] = L.cheby_coefs[i];
  if (n_coefs) *n_coefs = L.cheby_order + 1;
  return AMGB_OK;
}

int orc_tql1(int32_t n, double* diag, double* offdiag) { return tql1(n, diag, offdiag); }

int orc_vmult(orc_hier* h, double* z, const double* r) {
  if (!h || !z || !r) return AMGB_ERR_BAD_ARG;
  Level& L = h->lv[0];
  std::copy(r, r + L.A.n, L.f.begin());
  std::fill(L.u.begin(), L.u.end(), 0.0);
  const unsigned iters = std::max(1u, h->data.max_iter);
  for (unsigned it = 0; it < iters; ++it) cycle(*h, 0);
  std::copy(L.u.begin(), L.u.end(), z);
  return 0;
}

int orc_spmv(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
             const double* x, double* y) {
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
    y[i] = s;
  }
  return 0;
}

// A.4: KSPSolve_CG with KSP_NORM_PRECONDITIONED, deal.II SolverControl (absolute tol).
int orc_cg_solve(orc_hier* h, int64_t n, const int32_t* rowptr, const int32_t* col,
                 const double* val, double* x, const double* b, int64_t max_steps,
                 double abs_tol, double* res_hist, int64_t hist_cap, int64_t* n_iters) {
  if (!h || !x || !b || !n_iters) return AMGB_ERR_BAD_ARG;
  std::vector<double> r(n), z(n), p(n), w(n);
  auto dot = [&](const std::vector<double>& a, const std::vector<double>& c) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * c[i];
    return s;
  };
  orc_spmv(n, rowptr, col, val, x, w.data());
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - w[i];
  orc_vmult(h, z.data(), r.data());
  double dp = std::sqrt(dot(z, z));
  int64_t it = 0;
  if (res_hist && hist_cap > 0) res_hist[0] = dp;
  *n_iters = 0;
  if (!(dp == dp)) return AMGB_ERR_BREAKDOWN;
  if (dp <= abs_tol) return 0;
  double beta = dot(z, r), beta_old = 0.0;
  for (;;) {
    if (it >= max_steps) {
      *n_iters = it;
      return AMGB_ERR_NO_CONVERGENCE;
    }
    if (it == 0) {
      p = z;
    } else {
      const double bb = beta / beta_old;
      for (int64_t i = 0; i < n; ++i) p[i] = z[i] + bb * p[i];
    }
    orc_spmv(n, rowptr, col, val, p.data(), w.data());
    const double pw = dot(p, w);
    if (pw == 0.0 || !(pw == pw)) {
      *n_iters = it;
      return AMGB_ERR_BREAKDOWN;
    }
    const double alpha = beta / pw;
    for (int64_t i = 0; i < n; ++i) x[i] += alpha * p[i];
    for (int64_t i = 0; i < n; ++i) r[i] -= alpha * w[i];
    orc_vmult(h, z.data(), r.data());
    dp = std::sqrt(dot(z, z));
    beta_old = beta;
    beta = dot(z, r);
    ++it;
    if (res_hist && it < hist_cap) res_hist[it] = dp;
    if (!(dp == dp)) {
      *n_iters = it;
      return AMGB_ERR_BREAKDOWN;
    }
    if (dp <= abs_tol) break;
  }
  *n_iters = it;
  return 0;
}

// ref common/view_maker.h:26-74 (literal restatement; PetscInt arithmetic).
int orc_make_view(int64_t n64, const int32_t* rowptr, const int32_t* col, const double* val,
                  int32_t view_size, double* sum, int64_t* count, double* max_pp,
                  double* max_np) {
  if (view_size < 1 || n64 < 1) return AMGB_ERR_BAD_ARG;
  const int64_t vv = int64_t(view_size) * view_size;
  for (int64_t k = 0; k < vv; ++k) {
    sum[k] = 0.0;
    max_pp[k] = 0.0;
    max_np[k] = 0.0;
    count[k] = 0;
  }
  const int32_t n = (int32_t)n64;
  const int32_t q = n / view_size;
  const int32_t q1 = q + 1;
  const int32_t p = n % view_size;
  const int32_t t = q1 * p;
  for (int32_t i = 0; i < n; ++i) {
    const int32_t bin_row = (i < t) ? (i / q1) : ((i - t) / q + p);
    for (int32_t j = rowptr[i]; j < rowptr[i + 1]; ++j) {
      const int32_t c = col[j];
      const int32_t bin_col = (c < t) ? (c / q1) : ((c - t) / q + p);
      const int64_t flat = int64_t(view_size) * bin_row + bin_col;
      sum[flat] += val[j];
      count[flat] += 1;
      max_pp[flat] = std::max(std::max(val[j], 0.0), max_pp[flat]);
      max_np[flat] = std::max(std::max(-val[j], 0.0), max_np[flat]);
    }
  }
  return 0;
}

}  // extern "C"
