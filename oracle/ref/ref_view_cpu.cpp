// Runs the REFERENCE's own pooling code (ViewMaker, ref common/view_maker.h:17-92),
// compiled from where it lies under /root/reference, on a CSR read from a file.
// The deal.II/PETSc names it needs (SparseMatrix::m(), petsc_matrix(), MatGetRow,
// MatRestoreRow, PetscInt, PetscScalar) come from amg-ann_b200/host/dealii_compat,
// which serves rows from a host CSR copy; no device is touched.
// TEST INFRASTRUCTURE (oracle/): pins oracle/amg_oracle.cpp:orc_make_view and
// generates tests/golden/view_*.npz.
//
//   ref_view_cpu <csr.bin> <view_size> <out.csv>
// csr.bin: int64 n, int64 nnz, int64 rowptr[n+1], int32 col[nnz], double val[nnz]
// out.csv: the reference's own CSV fields  t_view,view_size,"view","count","max_pp","max_np"
#include <iomanip>

#include "view_maker.h"  // the reference's file, -I/root/reference/code/data-generation/common

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int64_t n = 0, nnz = 0;
  in.read((char*)&n, 8);
  in.read((char*)&nnz, 8);
  std::vector<int64_t> rp(n + 1);
  std::vector<int32_t> col(nnz);
  std::vector<double> val(nnz);
  in.read((char*)rp.data(), 8 * (n + 1));
  in.read((char*)col.data(), 4 * nnz);
  in.read((char*)val.data(), 8 * nnz);
  if (!in) return 3;
  dealii::PETScWrappers::MPI::SparseMatrix A;
  A.reinit_csr(n, rp.data(), col.data(), val.data());
  std::fstream out(argv[3], std::fstream::out | std::fstream::trunc);
  out << std::scientific << std::setprecision(17);  // ref t2 main.cpp:503
  ViewMaker vm(std::atoi(argv[2]));
  vm.make_view(out, A);
  vm.print_view(out);
  return 0;
}
