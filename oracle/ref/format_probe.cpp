// Prints the text the dealii_compat layer emits (hypre-format setup statistics, PETSc
// -ksp_monitor lines) for stats given on stdin, so tests can pipe it into the
// reference's scrapers (ref_parse).  No device needed.  TEST INFRASTRUCTURE.
//   format_probe boomeramg theta mrs max_levels  < "rows nnz\n..." lines
//   format_probe ksp < "residual\n..." lines
#include <cstring>

#include "amgb_dealii_compat.hpp"

int main(int argc, char** argv) {
  if (argc >= 2 && !std::strcmp(argv[1], "ksp")) {
    double r;
    for (int it = 0; std::cin >> r; ++it) std::fputs(amgb::compat::format_ksp_monitor_line(it, r).c_str(), stdout);
    return 0;
  }
  if (argc != 5) return 2;
  amgb::LevelStats st;
  long long rows, nnz;
  double sr = 0, sa = 0;
  while (std::cin >> rows >> nnz) {
    st.rows.push_back(rows);
    st.nnz.push_back(nnz);
    st.sparsity.push_back(double(nnz) / (double(rows) * double(rows)));
    sr += rows;
    sa += nnz;
  }
  st.grid = sr / st.rows[0];
  st.op = sa / st.nnz[0];
  st.memory = st.op * 1.25;
  std::fputs(amgb::compat::format_hypre_setup_stats(st, std::atof(argv[2]), std::atof(argv[3]), std::atoi(argv[4]),
                                                    "PMIS", "modified classical interpolation")
                 .c_str(),
             stdout);
  return 0;
}
