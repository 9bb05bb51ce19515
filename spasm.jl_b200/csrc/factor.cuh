// factor.cuh — device-resident echelon factor (U, qinv) and helpers shared by the entry points.
#pragma once
#include "common.cuh"
#include "solve_sparse.cuh"

namespace sb {

void transpose_csr(const DCsr &A, DCsr &T);  // runtime.cu

struct DevFactor {
  DCsr U;            // r x m, unit pivot stored first in each row
  DBuf<int> qinv;    // [m]
  int64_t prime = 0;
  Fp F;
  void upload(const spasm_lu *fact);
};

// pdesc of the system x.U = b: column c is eliminated by row qinv[c], prio = that row index
void build_pdesc_U(const DCsr &U, const int *qinv, DBuf<PDesc> &pdesc);
// capacity helpers for a growing U on the device
void csr_reserve(DCsr &U, int64_t nnz_needed, int rows_needed);

// the kernel-side system: Ut relabelled so that the solution vector is indexed by pivot COLUMN
struct KernelSystem {
  DCsr Ut;              // m x r, column indices replaced by pivot columns
  DBuf<PDesc> pdesc;    // [m]
  DBuf<int> freecols;   // [m - r] increasing
  int nfree = 0;
};
void build_kernel_system(const DevFactor &f, KernelSystem &K);

spasm_csr *result_to_host_csr(const SolveResult &R, int nrows, int m, int64_t prime, const Fp &F);

}  // namespace sb
