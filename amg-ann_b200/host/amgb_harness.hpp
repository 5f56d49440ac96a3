// amgb_harness.hpp -- host-side mirror of the reference's solve harness and pooling
// class on top of the deal.II-compat layer:
//   amgb::harness::amg_solve   ref common/amg_solver.h:22-92
//   amgb::harness::ViewMaker   ref common/view_maker.h:17-92
// Same call signatures, same CSV fields in the same order and number format
// (ref common/myutils.h:69-83 quoted arrays; the caller sets
// std::scientific << std::setprecision(17) as ref t2 main.cpp:503 does).  Differences:
// level statistics and the residual history come back as structs from the library
// instead of being scraped from redirected stdout (ref amg_solver.h:41-44,58-86), and the
// pooled image is computed on the device.
#pragma once

#include <chrono>
#include <fstream>
#include <vector>

#include "dealii_compat/amgb_dealii_compat.hpp"

namespace amgb {
namespace harness {

using BoomerAMGData = dealii::PETScWrappers::PreconditionBoomerAMG::AdditionalData;

template <class Iterable>
inline void print_quoted(const Iterable& v, std::ostream& out) {  // ref myutils.h:69-83
  out << '"';
  size_t i = 1;
  for (const auto& e : v) {
    out << e;
    if (i != v.size()) out << ',';
    ++i;
  }
  out << '"';
}

struct SolveRecord {
  long long t_setup_us = 0, t_solve_us = 0;
  unsigned int niters = 0;
  bool converged = false;
  LevelStats stats;
  std::vector<double> p_res;
};

// ref common/amg_solver.h:22-92.  Returns the number of CG iterations; `rec` (optional)
// receives the same numbers that go to the CSV.
inline unsigned int amg_solve(const BoomerAMGData& data, double rtol, std::ostream& filestream,
                              dealii::PETScWrappers::MPI::SparseMatrix& system_matrix,
                              dealii::PETScWrappers::MPI::Vector& system_rhs,
                              dealii::PETScWrappers::MPI::Vector& solution, SolveRecord* rec = nullptr) {
  using namespace std::chrono;
  filestream << data.strong_threshold << "," << data.max_row_sum << "," << data.symmetric_operator << ","
             << data.aggressive_coarsening_num_levels << "," << rtol << ",";
  dealii::SolverControl solver_control((unsigned)solution.size(), rtol);  // absolute tolerance (A.4)
  dealii::PETScWrappers::SolverCG cg(solver_control);
  dealii::PETScWrappers::PreconditionBoomerAMG preconditioner;
  SolveRecord local;
  SolveRecord& r = rec ? *rec : local;
  BoomerAMGData quiet = data;
  quiet.output_details = false;  // statistics are read from the struct, nothing is printed
  system_matrix.device();        // upload (once per matrix) stays outside the setup timer, like assembly
  {
    const auto t1 = high_resolution_clock::now();
    preconditioner.initialize(system_matrix, quiet);
    const auto t2 = high_resolution_clock::now();
    r.t_setup_us = duration_cast<microseconds>(t2 - t1).count();
    filestream << r.t_setup_us << ",";
  }
  const auto t3 = high_resolution_clock::now();
  bool failed = false;
  try {
    cg.solve(system_matrix, solution, system_rhs, preconditioner);
  } catch (const dealii::SolverControl::NoConvergence&) {
    failed = true;  // the reference lets the exception end the run (t2 main.cpp:540-544)
  }
  const auto t4 = high_resolution_clock::now();
  r.t_solve_us = duration_cast<microseconds>(t4 - t3).count();
  filestream << r.t_solve_us << ",";
  if (data.output_details) {
    r.stats = preconditioner.backend().level_stats();
    print_quoted(r.stats.rows, filestream);
    filestream << ",";
    print_quoted(r.stats.nnz, filestream);
    filestream << ",";
    print_quoted(r.stats.sparsity, filestream);
    filestream << "," << r.stats.grid << "," << r.stats.op << "," << r.stats.memory << ",";
  }
  r.niters = solver_control.last_step();
  r.converged = !failed;
  r.p_res = cg.residual_history();
  filestream << r.niters << ",";
  print_quoted(r.p_res, filestream);
  filestream << "\n";
  if (failed) throw dealii::SolverControl::NoConvergence(r.niters, solver_control.last_value());
  return r.niters;
}

// ref common/view_maker.h:17-92, computed by amgb_make_view on the device.
class ViewMaker {
 public:
  explicit ViewMaker(PetscInt vs)
      : m_view_size(vs), view((size_t)vs * vs), max_pp((size_t)vs * vs), max_np((size_t)vs * vs),
        count((size_t)vs * vs) {}
  void make_view(std::ostream& filestream, dealii::PETScWrappers::MPI::SparseMatrix& system_matrix) {
    using namespace std::chrono;
    const Matrix& A = system_matrix.device();
    const auto t1 = high_resolution_clock::now();
    A.context().check(amgb_make_view(A.context().get(), A.get(), m_view_size, view.data(), count.data(),
                                     max_pp.data(), max_np.data(), &device_us),
                      "amgb_make_view");
    const auto t2 = high_resolution_clock::now();
    filestream << duration_cast<microseconds>(t2 - t1).count() << ",";
  }
  void print_view(std::ostream& filestream) const {
    filestream << m_view_size << ",";
    print_quoted(view, filestream);
    filestream << ",";
    print_quoted(count, filestream);
    filestream << ",";
    print_quoted(max_pp, filestream);
    filestream << ",";
    print_quoted(max_np, filestream);
    filestream << "\n";
  }
  const PetscInt m_view_size;
  std::vector<double> view, max_pp, max_np;
  std::vector<int64_t> count;
  double device_us = 0.0;
};

}  // namespace harness
}  // namespace amgb
