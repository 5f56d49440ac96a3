// The REFERENCE's own harness -- amg_solver::amg_solve (ref common/amg_solver.h:22-92)
// and ViewMaker (ref common/view_maker.h) -- compiled UNMODIFIED from /root/reference
// against amg-ann_b200/host/dealii_compat and linked to libamgb.so, driven by the
// theta-sweep loop of ref testcase2-diffusion-structured/src/main.cpp:440-467 on a
// system from libamgb_gen.so.  It shows the drop-in boundary end to end: the reference
// code scrapes the hypre-format statistics and the -ksp_monitor lines the compat layer
// prints, and writes its own stats.csv rows.  TEST INFRASTRUCTURE (oracle/_ref):
// tests/test_gpu_reference_harness.py compares the CSV with the library's direct
// results.  Needs a GPU at run time.
//
//   ref_harness_gpu <m> <pattern_size> <mode> <contrast_exp> <theta0,theta1,dtheta> <view_size> <out.csv>
#include <cstdio>
#include <iomanip>

#include "amg_solver.h"  // the reference's files
#include "view_maker.h"
#include "amgb_gen.h"

int main(int argc, char** argv) {
  if (argc != 8) return 2;
  const int m = std::atoi(argv[1]), ps = std::atoi(argv[2]), mode = std::atoi(argv[3]);
  const double contrast = std::atof(argv[4]);
  double t0, t1, dt;
  if (std::sscanf(argv[5], "%lf,%lf,%lf", &t0, &t1, &dt) != 3) return 2;
  const int vs = std::atoi(argv[6]);
  int64_t n = 0, nnz = 0;
  if (amgb_gen_sizes(0, m, &n, &nnz)) return 3;
  int64_t ne = 1;
  for (int i = 0; i < mode; ++i) ne *= ps;
  std::vector<double> epsv(ne);
  amgb_gen_checkerboard_epsv(ps, mode, contrast, epsv.data());
  std::vector<int64_t> rp(n + 1);
  std::vector<int32_t> col(nnz);
  std::vector<double> val(nnz), rhs(n), x0(n);
  if (amgb_gen_poisson_q1(m, ps, mode, epsv.data(), ne, 0, n, rp.data(), col.data(), val.data(), rhs.data(),
                          x0.data()))
    return 3;
  using namespace dealii;
  PETScWrappers::MPI::SparseMatrix system_matrix;
  system_matrix.reinit_csr(n, rp.data(), col.data(), val.data());
  PETScWrappers::MPI::Vector system_rhs(n), solution(n), zero_solution(n);
  for (int64_t i = 0; i < n; ++i) {
    system_rhs[i] = rhs[i];
    zero_solution[i] = x0[i];
  }
  AffineConstraints<double> hanging_node_constraints;
  MPI_Comm comm = MPI_COMM_WORLD;
  const std::string stats_filename = argv[7];
  std::fstream filestream(stats_filename, std::fstream::out | std::fstream::trunc);
  filestream << std::scientific << std::setprecision(17);
  {
    ViewMaker vm(vs);
    filestream << "view,";
    vm.make_view(filestream, system_matrix);
    vm.print_view(filestream);
  }
  for (double t = t0; t <= t1; t += dt) {  // ref t2 main.cpp:443
    solution = zero_solution;
    const amg_solver::BoomerAMGData data(1, t, 0.9, 0, true);  // ref t2 main.cpp:447-453
    filestream << "solve,";
    amg_solver::amg_solve(data, 1e-8, filestream, system_matrix, system_rhs, solution, comm, stats_filename,
                          hanging_node_constraints);
  }
  filestream.close();
  return 0;
}
