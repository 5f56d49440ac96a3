// amgb_datagen -- C++ host driver of the hot path: the theta-sweep data-generation loop of
// ref testcase2-diffusion-structured/src/main.cpp:400-467 (Problem::run) on synthetic Q1
// systems from libamgb_gen.so, writing the reference's stats.csv schema (:403-416).
//
//   amgb_datagen [--m 100] [--pattern-size 4] [--mode 3] [--contrast 6 | --seed S --eps-max 6]
//                [--theta 0.05,0.96,0.05] [--max-row-sum 0.9] [--tol 1e-8] [--details 1]
//                [--make-view 0|1] [--view-size 75] [--systems 1] [--threads 1]
//                [--theta-lanes 1] [--device-assembly 0|1] [--setting NAME] --out stats.csv
//
// --systems N --threads T: N independent systems (seeds S..S+N-1, ref 00_data-generation.py
// :105-116 fans them out over processes) are processed by T host threads, each with its own
// amgb context/stream on the same GPU, so small systems overlap on the device.  Rows of one
// system are written contiguously; systems appear in completion order.
// --theta-lanes L: the theta values of ONE system are independent too; L host threads (one
// context/stream each) sweep them side by side on the shared, read-only device matrix.  Rows
// keep the order of the sweep.
#include <atomic>
#include <cstring>
#include <ctime>
#include <iomanip>
#include <mutex>
#include <sstream>
#include <thread>

#include "amgb_gen.h"
#include "amgb_harness.hpp"

namespace {

struct Args {
  int m = 100, ps = 4, mode = 3, details = 1, make_view = 0, view_size = 75, systems = 1, threads = 1;
  int device_assembly = 0, theta_lanes = 1;
  double contrast = 6.0, eps_max = 6.0, t0 = 0.05, t1 = 0.96, dt = 0.05, mrs = 0.9, tol = 1e-8;
  long seed = -1;
  std::string out, setting = "synthetic";
};

bool parse(int argc, char** argv, Args& a) {
  for (int i = 1; i < argc; ++i) {
    const std::string k = argv[i];
    auto val = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (k == "--m") a.m = std::atoi(val());
    else if (k == "--pattern-size") a.ps = std::atoi(val());
    else if (k == "--mode") a.mode = std::atoi(val());
    else if (k == "--contrast") a.contrast = std::atof(val());
    else if (k == "--seed") a.seed = std::atol(val());
    else if (k == "--eps-max") a.eps_max = std::atof(val());
    else if (k == "--theta") { if (std::sscanf(val(), "%lf,%lf,%lf", &a.t0, &a.t1, &a.dt) != 3) return false; }
    else if (k == "--max-row-sum") a.mrs = std::atof(val());
    else if (k == "--tol") a.tol = std::atof(val());
    else if (k == "--details") a.details = std::atoi(val());
    else if (k == "--make-view") a.make_view = std::atoi(val());
    else if (k == "--view-size") a.view_size = std::atoi(val());
    else if (k == "--systems") a.systems = std::atoi(val());
    else if (k == "--threads") a.threads = std::atoi(val());
    else if (k == "--device-assembly") a.device_assembly = std::atoi(val());
    else if (k == "--theta-lanes") a.theta_lanes = std::atoi(val());
    else if (k == "--setting") a.setting = val();
    else if (k == "--out") a.out = val();
    else return false;
  }
  return !a.out.empty() && a.m > 0 && a.mode >= 1 && a.mode <= 3 && a.dt > 0;
}

struct Totals {
  std::atomic<long long> setup_us{0}, solve_us{0}, view_us{0};
  std::atomic<int> solves{0}, views{0};
};

// one system = one matrix: optional pooled image, then the theta sweep (ref t2 main.cpp:431-467)
void run_system(const Args& a, long seed, std::ostream& out, Totals& tot) {
  int64_t n = 0, nnz = 0;
  if (amgb_gen_sizes(0, a.m, &n, &nnz)) throw std::runtime_error("amgb_gen_sizes");
  int64_t ne = 1;
  for (int i = 0; i < a.mode; ++i) ne *= a.ps;
  std::vector<double> epsv(ne);
  if (seed >= 0) amgb_gen_random_vec(seed, ne, a.eps_max, epsv.data());  // ref myutils.h:47-54
  else amgb_gen_checkerboard_epsv(a.ps, a.mode, a.contrast, epsv.data());
  using namespace dealii;
  PETScWrappers::MPI::SparseMatrix system_matrix;
  PETScWrappers::MPI::Vector system_rhs(n), solution(n), zero_solution(n);
  if (a.device_assembly) {  // the matrix is born in HBM (amgb_matrix_assemble_poisson_q1)
    system_matrix.reinit_device_poisson_q1(a.m, a.ps, a.mode, epsv.data(), ne, system_rhs, zero_solution);
  } else {
    std::vector<int64_t> rp(n + 1);
    std::vector<int32_t> col(nnz);
    std::vector<double> val(nnz), rhs(n), x0(n);
    if (amgb_gen_poisson_q1(a.m, a.ps, a.mode, epsv.data(), ne, 0, n, rp.data(), col.data(), val.data(), rhs.data(),
                            x0.data()))
      throw std::runtime_error("amgb_gen_poisson_q1");
    system_matrix.reinit_csr(n, rp.data(), col.data(), val.data());
    for (int64_t i = 0; i < n; ++i) {
      system_rhs[i] = rhs[i];
      zero_solution[i] = x0[i];
    }
  }
  auto print_stats = [&]() {  // ref t2 main.cpp:498-512
    out << std::scientific << std::setprecision(17);
    out << a.setting << "," << 3 << "," << n << "," << a.m << "," << 1 << "," << 0 << "," << a.ps << ",";
    amgb::harness::print_quoted(epsv, out);
    out << "," << a.mode << "," << std::time(nullptr) << ",";
  };
  if (a.make_view) {
    amgb::harness::ViewMaker vm(a.view_size);
    print_stats();
    std::ostringstream t;
    vm.make_view(t, system_matrix);
    out << t.str();
    vm.print_view(out);
    tot.view_us += (long long)vm.device_us;
    tot.views++;
    return;
  }
  std::vector<double> thetas;
  for (double t = a.t0; t <= a.t1; t += a.dt) thetas.push_back(t);  // accumulation, as ref t2 main.cpp:443
  if (a.theta_lanes <= 1) {
    for (double t : thetas) {
      solution = zero_solution;
      const amgb::harness::BoomerAMGData data(true, t, a.mrs, 0, a.details != 0);
      print_stats();
      amgb::harness::SolveRecord rec;
      amgb::harness::amg_solve(data, a.tol, out, system_matrix, system_rhs, solution, &rec);
      tot.setup_us += rec.t_setup_us;
      tot.solve_us += rec.t_solve_us;
      tot.solves++;
    }
    return;
  }
  // theta lanes: the matrix is uploaded once here, then only read
  system_matrix.device();
  std::vector<std::string> rows(thetas.size());
  std::atomic<size_t> next{0};
  std::mutex err_mutex;
  std::string error;
  auto lane = [&]() {
    try {
      amgb::compat::default_context().reserve(110 * nnz + (int64_t(64) << 20));  // this lane's pool, grown once
      PETScWrappers::MPI::Vector x(n);
      for (;;) {
        // largest theta (most iterations) first; the rows are emitted in sweep order below
        const size_t k = next++;
        if (k >= thetas.size()) break;
        const size_t idx = thetas.size() - 1 - k;
        x = zero_solution;
        const amgb::harness::BoomerAMGData data(true, thetas[idx], a.mrs, 0, a.details != 0);
        std::ostringstream row;
        row << std::scientific << std::setprecision(17);
        amgb::harness::SolveRecord rec;
        amgb::harness::amg_solve(data, a.tol, row, system_matrix, system_rhs, x, &rec);
        rows[idx] = row.str();
        tot.setup_us += rec.t_setup_us;
        tot.solve_us += rec.t_solve_us;
        tot.solves++;
      }
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> g(err_mutex);
      error = e.what();
    }
  };
  std::vector<std::thread> lanes;
  for (int t = 0; t < a.theta_lanes; ++t) lanes.emplace_back(lane);
  for (auto& t : lanes) t.join();
  if (!error.empty()) throw std::runtime_error(error);
  for (const std::string& r : rows) {
    print_stats();
    out << r;
  }
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  if (!parse(argc, argv, a)) {
    std::fprintf(stderr, "usage: see the header of amgb_datagen.cpp\n");
    return 2;
  }
  try {
    std::fstream file(a.out, std::fstream::out | std::fstream::app);
    if (a.make_view) {
      file << "setting,dim,ndof,mesh_ref,degree,sol_id,sol_pattern_size,epsv,mode,timestamp,t_view,view_size,view,"
              "view_count,view_max_pp,view_max_np\n";
    } else {  // ref t2 main.cpp:410-415 (column names as the reference writes them)
      file << "setting,dim,ndof,mesh_ref,degree,sol_id,sol_pattern_size,epsv,mode,timestamp,theta,maxrowsum,symop,"
              "tol,t_amg_setup,";
      if (a.details) file << "nrows,nze,sparsity,grid,operator,memory,";
      file << "t_solve,niters,p_res\n";
    }
    Totals tot;
    std::mutex file_mutex;
    std::atomic<int> next{0};
    std::string error;
    const auto w0 = std::chrono::high_resolution_clock::now();
    int warm_threads = 0, warm_solves = 0;  // (under file_mutex)
    double warm_wall = 0.0;
    auto worker = [&]() {
      try {
        {  // this thread's pool grown once (matrix + hierarchy + work vectors of one system) instead of
           // allocation by allocation: pool growth stalls every lane of the device
          int64_t n = 0, nnz = 0;
          if (!a.make_view && amgb_gen_sizes(0, a.m, &n, &nnz) == 0)  // (pooling allocates a few KB: nothing to reserve)
            amgb::compat::default_context().reserve(130 * nnz + (int64_t(64) << 20));
        }
        bool first = true;
        for (;;) {
          const int s = next++;
          if (s >= a.systems) break;
          std::ostringstream rows;
          run_system(a, a.seed >= 0 ? a.seed + s : (a.systems > 1 ? s : -1), rows, tot);
          std::lock_guard<std::mutex> g(file_mutex);
          file << rows.str();
          if (first) {  // this thread is warm (context, pool, kernels loaded): steady state starts when all are
            first = false;
            if (++warm_threads == std::min(a.threads, a.systems)) {
              warm_wall = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - w0).count();
              warm_solves = tot.solves.load();
            }
          }
        }
      } catch (const std::exception& e) {
        std::lock_guard<std::mutex> g(file_mutex);
        error = e.what();
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < a.threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    const double wall = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - w0).count();
    if (!error.empty()) throw std::runtime_error(error);
    // steady state: from the moment every thread has finished its first system (process start-up, context
    // creation and the first launch of every kernel are behind it) to the end
    const double steady_s = warm_wall > 0 ? wall - warm_wall : 0.0;
    const int steady_solves = warm_wall > 0 ? tot.solves.load() - warm_solves : 0;
    std::printf("{\"systems\": %d, \"solves\": %d, \"views\": %d, \"threads\": %d, \"wall_s\": %.6f, "
                "\"setup_s\": %.6f, \"solve_s\": %.6f, \"view_device_s\": %.6f, \"steady_s\": %.6f, "
                "\"steady_solves\": %d}\n",
                a.systems, tot.solves.load(), tot.views.load(), a.threads, wall, tot.setup_us / 1e6,
                tot.solve_us / 1e6, tot.view_us / 1e6, steady_s, steady_solves);
  } catch (const std::exception& e) {  // ref t2 main.cpp:540-544
    std::cerr << "Exception on processing: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
