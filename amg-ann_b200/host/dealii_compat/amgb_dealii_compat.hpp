// amgb_dealii_compat.hpp -- the slice of deal.II's PETScWrappers API that the
// reference's hot path touches, implemented on the C ABI of libamgb.so.
//
// Purpose: the reference's own harness headers
//     common/amg_solver.h   (amg_solver::amg_solve, :22-92)
//     common/view_maker.h   (ViewMaker, :17-92)
// compile UNMODIFIED against this directory (put it on the include path in place
// of deal.II's; the <deal.II/...> headers here all include this file), so the
// path "initialize(matrix, data) then vmult inside cg.solve" runs on the B200
// without touching the reference's source.  Only what those two headers and the
// theta-sweep loop (ref testcase2-diffusion-structured/src/main.cpp:440-467) use
// is provided; meshes, DoF handlers and assembly are out of scope.
//
// Behaviour mirrored (deal.II petsc_precondition.cc,
// petsc_solver.cc):
//  * AdditionalData: deal.II's 12-argument constructor, same order and defaults.
//  * initialize(): theta and max_row_sum are forwarded through std::to_string
//    (options_via_string = 1); with output_details the hypre setup statistics are
//    printed on stdout in hypre's layout so that the reference's BoomerAMGParser
//    (common/parser.h:181-266) scrapes them as before.
//  * SolverCG::solve(): KSPCG, initial guess honoured, convergence on the
//    preconditioned residual against SolverControl's ABSOLUTE tolerance, failure
//    -> SolverControl::NoConvergence; with the "-ksp_monitor" option set
//    (amg_solver.h:36) every residual is printed as PETSc does
//    ("%3d KSP Residual norm %14.12e"), which parser.h:149-155 scrapes.
//  * MatGetRow/MatRestoreRow serve the host copy of the CSR (view_maker.h:48,67).
// Device-only knobs (coarsening type, smoother policy) that deal.II leaves to
// PETSc's option database are members of AdditionalData with the library defaults.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../amgb.hpp"

// ---- the PETSc C names the reference uses directly -------------------------
typedef int32_t PetscInt;  // default PETSc build: 32-bit indices
typedef double PetscScalar;
typedef int PetscErrorCode;
typedef int MPI_Comm;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif

namespace amgb {
namespace compat {

// One context per host thread (include/amgb.h: "one ctx per host thread").
inline Context& default_context() {
  static thread_local std::unique_ptr<Context> ctx;
  if (!ctx) {
    const char* dev = std::getenv("AMGB_DEVICE");
    ctx.reset(new Context(dev ? std::atoi(dev) : 0));
  }
  return *ctx;
}

// PETSc's global options database, reduced to the keys this path reads.
inline std::map<std::string, std::string>& options() {
  static std::map<std::string, std::string> db;
  return db;
}

struct CsrHost {
  int64_t n = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<double> val;
};

// hypre_BoomerAMGSetupStats (par_stats.c) layout, as the reference's regex expects it
// (ref common/parser.h:185-225).  Columns the parser does not read (entries per row,
// row sums, interpolation table) carry the values the library reports.
struct RowStats {
  int32_t min_entries = 0, max_entries = 0;
  double min_row_sum = 0.0, max_row_sum = 0.0;
};

inline std::string format_hypre_setup_stats(const LevelStats& st, double theta, double max_row_sum,
                                            int max_levels, const char* coarsening, const char* interpolation,
                                            const std::vector<int64_t>* nnz_P = nullptr,
                                            const std::vector<RowStats>* rows = nullptr) {
  std::string out;
  char buf[512];
  auto add = [&](const char* fmt, auto... a) {
    std::snprintf(buf, sizeof buf, fmt, a...);
    out += buf;
  };
  const int nl = (int)st.rows.size();
  out += "\n\n Num MPI tasks = 1\n\n Num OpenMP threads = 1\n\n\nBoomerAMG SETUP PARAMETERS:\n\n";
  add(" Max levels = %d\n Num levels = %d\n\n", max_levels, nl);
  add(" Strength Threshold = %f\n", theta);
  add(" Interpolation Truncation Factor = %f\n", 0.0);
  add(" Maximum Row Sum Threshold for Dependency Weakening = %f\n\n", max_row_sum);
  add(" Coarsening Type = %s\n", coarsening);
  out += " measures are determined locally\n\n\n No global partition option chosen.\n\n";
  add(" Interpolation = %s\n\n", interpolation);
  out += "Operator Matrix Information:\n\n";
  out += "             nonzero            entries/row          row sums\n";
  out += "lev    rows  entries sparse   min  max     avg      min         max\n";
  out += "======================================================================\n";
  for (int l = 0; l < nl; ++l) {
    const double avg = st.rows[l] ? double(st.nnz[l]) / double(st.rows[l]) : 0.0;
    const RowStats rs = rows ? (*rows)[l] : RowStats();
    add("%2d %7lld %8lld  %0.3f  %4d %4d  %6.1f  %10.3e  %10.3e\n", l, (long long)st.rows[l],
        (long long)st.nnz[l], st.sparsity[l], (int)rs.min_entries, (int)rs.max_entries, avg, rs.min_row_sum,
        rs.max_row_sum);
  }
  out += "\n\nInterpolation Matrix Information:\n";
  out += "                    entries/row        min        max            row sums\n";
  out += "lev  rows x cols   min  max  avgW     weight      weight       min         max\n";
  out += "================================================================================\n";
  for (int l = 0; l + 1 < nl; ++l) {
    const double avgw = (nnz_P && st.rows[l]) ? double((*nnz_P)[l]) / double(st.rows[l]) : 0.0;
    add("%2d %7lld x %-7lld %3d %4d  %4.1f  %10.3e  %10.3e  %10.3e  %10.3e\n", l, (long long)st.rows[l],
        (long long)st.rows[l + 1], 0, 0, avgw, 0.0, 0.0, 0.0, 0.0);
  }
  add("\n\n     Complexity:    grid = %f\n", st.grid);
  add("                operator = %f\n", st.op);
  add("                memory = %f\n\n", st.memory);
  return out;
}

// "%3D KSP Residual norm %14.12e \n" (PETSc KSPMonitorResidual; offsets 18 / 38 at parser.h:153-154)
inline std::string format_ksp_monitor_line(int it, double rnorm) {
  char buf[96];
  std::snprintf(buf, sizeof buf, "%3d KSP Residual norm %14.12e \n", it, rnorm);
  return buf;
}

}  // namespace compat
}  // namespace amgb

// Mat is an opaque handle to the host CSR copy kept by the compat SparseMatrix.
typedef const amgb::compat::CsrHost* Mat;

inline PetscErrorCode MatGetRow(Mat A, PetscInt row, PetscInt* ncols, const PetscInt** cols,
                                const PetscScalar** vals) {
  if (!A || row < 0 || row >= A->n || (int64_t)A->rowptr.size() != A->n + 1) return 63;  // PETSC_ERR_ARG_OUTOFRANGE
  const int64_t b = A->rowptr[row];
  if (ncols) *ncols = (PetscInt)(A->rowptr[row + 1] - b);
  if (cols) *cols = A->col.data() + b;
  if (vals) *vals = A->val.data() + b;
  return 0;
}

inline PetscErrorCode MatRestoreRow(Mat, PetscInt, PetscInt*, const PetscInt**, const PetscScalar**) { return 0; }

namespace dealii {

class ConditionalOStream {
 public:
  ConditionalOStream(std::ostream& s, bool active = true) : s_(s), active_(active) {}
  template <class T>
  const ConditionalOStream& operator<<(const T& t) const {
    if (active_) s_ << t;
    return *this;
  }
  const ConditionalOStream& operator<<(std::ostream& (*p)(std::ostream&)) const {
    if (active_) s_ << p;
    return *this;
  }

 private:
  std::ostream& s_;
  bool active_;
};

class SolverControl {
 public:
  enum State { iterate = 0, success, failure };
  class NoConvergence : public std::runtime_error {
   public:
    NoConvergence(unsigned int step, double residual)
        : std::runtime_error("Iterative method reported convergence failure in step " + std::to_string(step) +
                             ". The residual in the last step was " + std::to_string(residual) + "."),
          last_step(step),
          last_residual(residual) {}
    const unsigned int last_step;
    const double last_residual;
  };
  explicit SolverControl(unsigned int n = 100, double tol = 1.e-10, bool /*log_history*/ = false,
                         bool /*log_result*/ = true)
      : maxsteps_(n), tol_(tol) {}
  unsigned int last_step() const { return lstep_; }
  double last_value() const { return lvalue_; }
  unsigned int max_steps() const { return maxsteps_; }
  double tolerance() const { return tol_; }
  State last_check() const { return lcheck_; }
  // solvers report through this
  void set_result(unsigned int step, double value, State s) {
    lstep_ = step;
    lvalue_ = value;
    lcheck_ = s;
  }

 private:
  unsigned int maxsteps_;
  double tol_;
  unsigned int lstep_ = 0;
  double lvalue_ = 0.0;
  State lcheck_ = iterate;
};

template <class Number>
class Vector;

namespace PETScWrappers {

inline void set_option_value(const std::string& name, const std::string& value) {
  amgb::compat::options()[name] = value;
}

namespace MPI {

class Vector {
 public:
  Vector() = default;
  explicit Vector(std::size_t n) : v_(n, 0.0) {}
  void reinit(std::size_t n) { v_.assign(n, 0.0); }
  std::size_t size() const { return v_.size(); }
  double& operator[](std::size_t i) { return v_[i]; }
  double operator[](std::size_t i) const { return v_[i]; }
  double& operator()(std::size_t i) { return v_[i]; }
  double operator()(std::size_t i) const { return v_[i]; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  Vector& operator=(double s) {
    std::fill(v_.begin(), v_.end(), s);
    return *this;
  }
  template <class Number>
  Vector& operator=(const dealii::Vector<Number>& o);
  double l2_norm() const {
    double s = 0;
    for (double x : v_) s += x * x;
    return std::sqrt(s);
  }
  void compress(int = 0) {}

 private:
  std::vector<double> v_;
};

// System matrix: a host CSR copy (for MatGetRow) plus the device-resident matrix,
// uploaded once and kept across the whole theta sweep.
class SparseMatrix {
 public:
  SparseMatrix() = default;
  // rowptr: n+1 entries; columns ascending per row, diagonal stored (PETSc AIJ as deal.II builds it)
  template <class Index>
  void reinit_csr(int64_t n, const Index* rowptr, const int32_t* col, const double* val) {
    host_.n = n;
    host_.rowptr.assign(rowptr, rowptr + n + 1);
    const int64_t nnz = host_.rowptr[n];
    host_.col.assign(col, col + nnz);
    host_.val.assign(val, val + nnz);
    dev_.reset();
  }
  // The system assembled directly in device memory (no host CSR: MatGetRow is not served;
  // pooling and the solve run on the device copy).  rhs / x0: n = (m+1)^3 doubles.
  void reinit_device_poisson_q1(int m, int pattern_size, int mode, const double* epsv, int64_t n_epsv, Vector& rhs,
                                Vector& x0) {
    const int64_t N = (int64_t)m + 1, n = N * N * N;
    host_ = amgb::compat::CsrHost();
    host_.n = n;
    host_.rowptr.assign(1, 0);  // no host entries
    rhs.reinit((size_t)n);
    x0.reinit((size_t)n);
    dev_ = amgb::Matrix::assemble_poisson_q1(amgb::compat::default_context(), m, pattern_size, mode, epsv, n_epsv,
                                             rhs.data(), x0.data());
  }
  PetscInt m() const { return (PetscInt)host_.n; }
  PetscInt n() const { return (PetscInt)host_.n; }
  int64_t n_nonzero_elements() const { return host_.rowptr.size() > 1 ? host_.rowptr.back() : 0; }
  Mat petsc_matrix() const { return &host_; }
  operator Mat() const { return &host_; }
  const amgb::Matrix& device() const {
    if (!dev_)
      dev_ = amgb::Matrix(amgb::compat::default_context(), host_.n, host_.rowptr.data(), host_.col.data(),
                          host_.val.data());
    return dev_;
  }
  void vmult(Vector& dst, const Vector& src) const {
    const auto& A = device();
    A.context().check(amgb_matrix_vmult(A.context().get(), A.get(), dst.data(), src.data()), "amgb_matrix_vmult");
  }
  void compress(int = 0) {}

 private:
  amgb::compat::CsrHost host_;
  mutable amgb::Matrix dev_;
};

}  // namespace MPI

class PreconditionBoomerAMG {
 public:
  struct AdditionalData {
    enum class RelaxationType {
      Jacobi,
      sequentialGaussSeidel,
      seqboundaryGaussSeidel,
      SORJacobi,
      backwardSORJacobi,
      symmetricSORJacobi,
      l1scaledSORJacobi,
      GaussianElimination,
      l1GaussSeidel,
      backwardl1GaussSeidel,
      CG,
      Chebyshev,
      FCFJacobi,
      l1scaledJacobi,
      None
    };
    AdditionalData(const bool symmetric_operator = false, const double strong_threshold = 0.25,
                   const double max_row_sum = 0.9, const unsigned int aggressive_coarsening_num_levels = 0,
                   const bool output_details = false,
                   const RelaxationType relaxation_type_up = RelaxationType::SORJacobi,
                   const RelaxationType relaxation_type_down = RelaxationType::SORJacobi,
                   const RelaxationType relaxation_type_coarse = RelaxationType::GaussianElimination,
                   const unsigned int n_sweeps_coarse = 1, const double tol = 0.0, const unsigned int max_iter = 1,
                   const bool w_cycle = false)
        : symmetric_operator(symmetric_operator),
          strong_threshold(strong_threshold),
          max_row_sum(max_row_sum),
          aggressive_coarsening_num_levels(aggressive_coarsening_num_levels),
          output_details(output_details),
          relaxation_type_up(relaxation_type_up),
          relaxation_type_down(relaxation_type_down),
          relaxation_type_coarse(relaxation_type_coarse),
          n_sweeps_coarse(n_sweeps_coarse),
          tol(tol),
          max_iter(max_iter),
          w_cycle(w_cycle) {}
    bool symmetric_operator;
    double strong_threshold;
    double max_row_sum;
    unsigned int aggressive_coarsening_num_levels;
    bool output_details;
    RelaxationType relaxation_type_up, relaxation_type_down, relaxation_type_coarse;
    unsigned int n_sweeps_coarse;
    double tol;
    unsigned int max_iter;
    bool w_cycle;
    // knobs deal.II leaves to PETSc's options database (include/amgb.h), library defaults
    int coarsen_type = AMGB_COARSEN_PMIS;
    int interp_type = AMGB_INTERP_CLASSICAL;
    int smoother_policy = AMGB_SMOOTHER_SUBSTITUTE;

    amgb_boomeramg_data to_c() const {
      amgb_boomeramg_data d;
      amgb_boomeramg_data_default(&d);
      d.symmetric_operator = symmetric_operator;
      d.strong_threshold = strong_threshold;
      d.max_row_sum = max_row_sum;
      d.aggressive_coarsening_num_levels = aggressive_coarsening_num_levels;
      d.output_details = output_details;
      d.relaxation_type_up = (int)relaxation_type_up;
      d.relaxation_type_down = (int)relaxation_type_down;
      d.relaxation_type_coarse = (int)relaxation_type_coarse;
      d.n_sweeps_coarse = n_sweeps_coarse;
      d.tol = tol;
      d.max_iter = max_iter;
      d.w_cycle = w_cycle;
      d.coarsen_type = coarsen_type;
      d.interp_type = interp_type;
      d.smoother_policy = smoother_policy;
      d.options_via_string = 1;
      return d;
    }
  };

  PreconditionBoomerAMG() = default;
  PreconditionBoomerAMG(const MPI::SparseMatrix& matrix, const AdditionalData& data = AdditionalData()) {
    initialize(matrix, data);
  }

  // ref common/amg_solver.h:48
  void initialize(const MPI::SparseMatrix& matrix, const AdditionalData& data = AdditionalData()) {
    data_ = data;
    matrix_ = &matrix;
    amgb_boomeramg_data d = data.to_c();
    // Knobs the reference's AdditionalData has no field for, PETSc style: the options database
    // (PETScWrappers::set_option_value) or the environment, so that the UNMODIFIED reference harness can
    // ask for the flavour closest to its CPU defaults (Falgout's parallel stage + Gauss-Seidel sweeps):
    //   -amgb_coarsen_type  pmis | cljp                      (AMGB_COARSEN_TYPE)
    //   -amgb_smoother_policy  substitute | strict | multicolor   (AMGB_SMOOTHER_POLICY)
    {
      auto knob = [](const char* opt, const char* env) -> std::string {
        auto it = amgb::compat::options().find(opt);
        if (it != amgb::compat::options().end()) return it->second;
        const char* e = std::getenv(env);
        return e ? std::string(e) : std::string();
      };
      const std::string ct = knob("-amgb_coarsen_type", "AMGB_COARSEN_TYPE");
      if (ct == "cljp") d.coarsen_type = AMGB_COARSEN_CLJP;
      else if (ct == "pmis") d.coarsen_type = AMGB_COARSEN_PMIS;
      const std::string sp = knob("-amgb_smoother_policy", "AMGB_SMOOTHER_POLICY");
      if (sp == "multicolor") d.smoother_policy = AMGB_SMOOTHER_MULTICOLOR;
      else if (sp == "strict") d.smoother_policy = AMGB_SMOOTHER_STRICT;
      else if (sp == "substitute") d.smoother_policy = AMGB_SMOOTHER_SUBSTITUTE;
    }
    // on the calling thread's context: threads that sweep theta over one matrix side by side
    // share the (read-only) device copy and work on their own streams
    prec_.initialize(amgb::compat::default_context(), matrix.device(), d);
    if (data.output_details) {
      const amgb::LevelStats st = prec_.level_stats();
      std::vector<int64_t> nnzP(st.rows.size(), 0);
      for (int l = 0; l + 1 < (int)st.rows.size(); ++l) {
        int64_t n, nnzA, nc, np;
        if (amgb_precond_level_dims(prec_.get(), l, &n, &nnzA, &nc, &np) == AMGB_OK) nnzP[l] = np;
      }
      const double theta = std::strtod(std::to_string(data.strong_threshold).c_str(), nullptr);
      const double mrs = std::strtod(std::to_string(data.max_row_sum).c_str(), nullptr);
      std::vector<amgb::compat::RowStats> rows(st.rows.size());
      for (int l = 0; l < (int)st.rows.size(); ++l)
        amgb_precond_level_row_stats(prec_.get(), l, &rows[l].min_entries, &rows[l].max_entries,
                                     &rows[l].min_row_sum, &rows[l].max_row_sum);
      const std::string text = amgb::compat::format_hypre_setup_stats(
          st, theta, mrs, d.max_levels, d.coarsen_type == AMGB_COARSEN_PMIS ? "PMIS" : "Falgout-CLJP",
          "modified classical interpolation", &nnzP, &rows);
      std::fputs(text.c_str(), stdout);  // C stdio: the reference redirects fd 1 (redirector.h:101)
    }
  }

  // PCApply: dst = one V-cycle applied to src
  void vmult(MPI::Vector& dst, const MPI::Vector& src) const { prec_.vmult(dst.data(), src.data()); }

  const amgb::Preconditioner& backend() const { return prec_; }
  const MPI::SparseMatrix* matrix() const { return matrix_; }

 private:
  AdditionalData data_;
  const MPI::SparseMatrix* matrix_ = nullptr;
  amgb::Preconditioner prec_;
};

class SolverCG {
 public:
  struct AdditionalData {};
  explicit SolverCG(SolverControl& cn, const MPI_Comm& = MPI_COMM_WORLD, const AdditionalData& = AdditionalData())
      : control_(cn) {}

  // ref common/amg_solver.h:54
  void solve(const MPI::SparseMatrix& A, MPI::Vector& x, const MPI::Vector& b,
             const PreconditionBoomerAMG& preconditioner) {
    const amgb::Matrix& Ad = A.device();
    const amgb::Context& ctx = preconditioner.backend().context();
    const int64_t max_steps = control_.max_steps();
    // (the library keeps at most 65 536 history entries; sized outside of nothing bigger than that, so
    // the timed solve does not pay for zero-filling max_steps = n doubles)
    std::vector<double> hist((size_t)std::min<int64_t>(max_steps, (int64_t)65535) + 1);
    int64_t nit = 0;
    const int rc = amgb_cg_solve(ctx.get(), Ad.get(), x.data(), b.data(), preconditioner.backend().get(), max_steps,
                                 control_.tolerance(), hist.data(), (int64_t)hist.size(), &nit);
    const size_t k = std::min<size_t>(hist.size(), (size_t)nit + 1);
    if (amgb::compat::options().count("-ksp_monitor"))
      for (size_t i = 0; i < k; ++i) std::fputs(amgb::compat::format_ksp_monitor_line((int)i, hist[i]).c_str(), stdout);
    const double last = k ? hist[k - 1] : 0.0;
    history_.assign(hist.begin(), hist.begin() + k);
    if (rc == AMGB_ERR_NO_CONVERGENCE) {
      control_.set_result((unsigned)nit, last, SolverControl::failure);
      throw SolverControl::NoConvergence((unsigned)nit, last);
    }
    ctx.check(rc, "amgb_cg_solve");
    control_.set_result((unsigned)nit, last, SolverControl::success);
  }
  SolverControl& control() const { return control_; }
  const std::vector<double>& residual_history() const { return history_; }

 private:
  SolverControl& control_;
  std::vector<double> history_;
};

}  // namespace PETScWrappers

// dealii::Vector<double>: the serial copy used for constraints.distribute (amg_solver.h:88-90)
template <class Number>
class Vector {
 public:
  Vector() = default;
  explicit Vector(std::size_t n) : v_(n, Number(0)) {}
  Vector(const PETScWrappers::MPI::Vector& o) : v_(o.data(), o.data() + o.size()) {}
  std::size_t size() const { return v_.size(); }
  Number& operator[](std::size_t i) { return v_[i]; }
  Number operator[](std::size_t i) const { return v_[i]; }
  Number& operator()(std::size_t i) { return v_[i]; }
  Number operator()(std::size_t i) const { return v_[i]; }
  const Number* data() const { return v_.data(); }

 private:
  std::vector<Number> v_;
};

template <class Number>
PETScWrappers::MPI::Vector& PETScWrappers::MPI::Vector::operator=(const dealii::Vector<Number>& o) {
  v_.assign(o.data(), o.data() + o.size());
  return *this;
}

// Only constraints of the form x_i = value (Dirichlet) exist on the uniformly refined
// meshes of this path; distribute() re-imposes them (ref t3 main.cpp:264-268), and is a
// no-op for an empty object (ref t2: no hanging nodes).
template <class Number = double>
class AffineConstraints {
 public:
  void add_line(std::size_t i) { lines_[i]; }
  void set_inhomogeneity(std::size_t i, Number v) { lines_[i] = v; }
  void close() {}
  void clear() { lines_.clear(); }
  std::size_t n_constraints() const { return lines_.size(); }
  template <class VectorType>
  void distribute(VectorType& v) const {
    for (const auto& kv : lines_) v[kv.first] = kv.second;
  }

 private:
  std::map<std::size_t, Number> lines_;
};

}  // namespace dealii
