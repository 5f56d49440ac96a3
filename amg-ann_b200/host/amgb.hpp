// amgb.hpp -- C++ RAII layer over the C ABI (include/amgb.h) for host code.
//
// This is the host side of the drop-in boundary for C++ callers: exceptions
// instead of status codes, owning handles, nothing else.  The deal.II-shaped
// classes the reference's drivers use (PreconditionBoomerAMG, SolverCG,
// SolverControl, MPI::SparseMatrix ...; ref common/amg_solver.h:22-92) are in
// dealii_compat/ and are built on these.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "amgb.h"

namespace amgb {

class Error : public std::runtime_error {
 public:
  Error(int status, const std::string& where, const std::string& detail)
      : std::runtime_error(where + ": " + amgb_status_string(status) + " " + detail), status_(status) {}
  int status() const { return status_; }

 private:
  int status_;
};

class Context {
 public:
  explicit Context(int device = 0, void* stream = nullptr) {
    const int rc = amgb_ctx_create(&h_, device, stream);
    if (rc != AMGB_OK)
      throw Error(rc, "amgb_ctx_create", "(a CUDA device is required: this library has no CPU path)");
  }
  ~Context() { amgb_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  amgb_ctx* get() const { return h_; }
  void check(int rc, const char* where) const {
    if (rc != AMGB_OK) throw Error(rc, where, amgb_last_error(h_));
  }
  void synchronize() const { check(amgb_ctx_synchronize(h_), "amgb_ctx_synchronize"); }
  // grow the context's memory pool once, up front (amgb_ctx_reserve)
  void reserve(int64_t bytes) const { check(amgb_ctx_reserve(h_, bytes), "amgb_ctx_reserve"); }
  int64_t kernel_launches() const {
    int64_t v = 0;
    check(amgb_ctx_kernel_launches(h_, &v), "amgb_ctx_kernel_launches");
    return v;
  }

 private:
  amgb_ctx* h_ = nullptr;
};

// Device-resident CSR system matrix (stays resident across the theta sweep).
class Matrix {
 public:
  Matrix() = default;
  Matrix(const Context& ctx, int64_t n, const int64_t* rowptr, const int32_t* col, const double* val)
      : ctx_(&ctx) {
    ctx.check(amgb_matrix_upload_csr64(ctx.get(), n, rowptr, col, val, &h_), "amgb_matrix_upload_csr64");
  }
  Matrix(const Context& ctx, int64_t n, const int32_t* rowptr, const int32_t* col, const double* val)
      : ctx_(&ctx) {
    ctx.check(amgb_matrix_upload_csr(ctx.get(), n, rowptr, col, val, &h_), "amgb_matrix_upload_csr");
  }
  // on-device assembly of the Q1 diffusion system (ref t2 main.cpp:255-320); rhs / x0 on the host
  static Matrix assemble_poisson_q1(const Context& ctx, int m, int pattern_size, int mode, const double* epsv,
                                    int64_t n_epsv, double* rhs_host, double* x0_host) {
    Matrix M;
    M.ctx_ = &ctx;
    ctx.check(amgb_matrix_assemble_poisson_q1_hostvec(ctx.get(), m, pattern_size, mode, epsv, n_epsv, &M.h_,
                                                      rhs_host, x0_host),
              "amgb_matrix_assemble_poisson_q1_hostvec");
    return M;
  }
  ~Matrix() { reset(); }
  Matrix(Matrix&& o) noexcept : ctx_(o.ctx_), h_(o.h_) { o.h_ = nullptr; }
  Matrix& operator=(Matrix&& o) noexcept {
    if (this != &o) {
      reset();
      ctx_ = o.ctx_;
      h_ = o.h_;
      o.h_ = nullptr;
    }
    return *this;
  }
  void reset() {
    if (h_) amgb_matrix_destroy(h_);
    h_ = nullptr;
  }
  explicit operator bool() const { return h_ != nullptr; }
  amgb_matrix* get() const { return h_; }
  const Context& context() const { return *ctx_; }

 private:
  const Context* ctx_ = nullptr;
  amgb_matrix* h_ = nullptr;
};

struct LevelStats {
  std::vector<int64_t> rows, nnz;
  std::vector<double> sparsity;
  double grid = 0, op = 0, memory = 0;
};

class Preconditioner {
 public:
  Preconditioner() = default;
  ~Preconditioner() { reset(); }
  Preconditioner(const Preconditioner&) = delete;
  Preconditioner& operator=(const Preconditioner&) = delete;
  void initialize(const Matrix& A, const amgb_boomeramg_data& data) { initialize(A.context(), A, data); }
  // On another context (stream, host thread) of the same device than the one that uploaded
  // the matrix: an uploaded matrix is read-only, so the independent systems of a theta sweep
  // can be in flight side by side (include/amgb.h, amgb_matrix_upload_csr).
  void initialize(const Context& ctx, const Matrix& A, const amgb_boomeramg_data& data) {
    reset();
    ctx_ = &ctx;
    ctx_->check(amgb_precond_initialize(ctx_->get(), A.get(), &data, &h_), "amgb_precond_initialize");
  }
  void reset() {
    if (h_) amgb_precond_destroy(h_);
    h_ = nullptr;
  }
  void vmult(double* dst, const double* src) const {
    ctx_->check(amgb_precond_vmult(h_, dst, src), "amgb_precond_vmult");
  }
  LevelStats level_stats() const {
    LevelStats s;
    int32_t nl = 0;
    ctx_->check(amgb_precond_num_levels(h_, &nl), "amgb_precond_num_levels");
    s.rows.resize(nl);
    s.nnz.resize(nl);
    s.sparsity.resize(nl);
    ctx_->check(amgb_precond_level_stats(h_, nl, &nl, s.rows.data(), s.nnz.data(), s.sparsity.data(), &s.grid,
                                         &s.op, &s.memory),
                "amgb_precond_level_stats");
    return s;
  }
  amgb_precond* get() const { return h_; }
  const Context& context() const { return *ctx_; }

 private:
  const Context* ctx_ = nullptr;
  amgb_precond* h_ = nullptr;
};

}  // namespace amgb
