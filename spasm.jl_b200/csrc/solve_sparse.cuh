// solve_sparse.cuh — interface of the sparse row-solve engine (solve_sparse.cu).
//
// One engine serves every "eliminate a sparse row against a triangular system" step of the hot
// path: spasm_sparse_triangular_solve / spasm_schur (src/SpaSM.jl:694-722, :761-762), the density
// estimate (:763-764), GPLU batches (README.md:34-36), spasm_rref (:871) and the back-substitution
// of spasm_kernel (:876-882, on Ut).
#pragma once
#include "common.cuh"

namespace sb {

// per-column descriptor of the triangular system G: which row eliminates column c, and when.
struct __align__(16) PDesc {
  long long start;  // offset of that row in Gj/Gx
  int len;          // its length; < 0: column c is not pivotal
  int prio;         // elimination order: pending columns are processed by increasing prio
};

struct SolveSystem {
  const int *Gj = nullptr;
  const uint32_t *Gx = nullptr;
  const PDesc *pdesc = nullptr;
  int width = 0;  // number of columns of the solution vector x
  // when the system is x.U = b, rows that overflow the warp tiers are eliminated all at once by the
  // column-major SpTRSM engine (dense_engine.cu) and converted back to sorted sparse rows
  const DCsr *U_dense = nullptr;
  const int *qinv_dense = nullptr;
  // optional, [nprio/32 + 1] words: bit set at the first prio of every class of mutually independent pivots
  // (consecutive prios; e.g. the structural pivots of one round that have the same height in the pivot DAG).
  // The global tier eliminates all pending pivots of a class in one step.  nullptr: one pivot per step.
  const unsigned *classbit = nullptr;
};

struct SolveRows {
  const long long *Bp = nullptr;
  const int *Bj = nullptr;
  const uint32_t *Bx = nullptr;
  const int *rows = nullptr;  // [nrows] row indices into B (nullptr: 0..nrows-1)
  int nrows = 0;
  const int *mask = nullptr;  // [nrows] optional: column treated as non-pivotal for row k (rref)
  bool few_pivots = false;    // hint: these rows reach only a handful of pivots (GPLU rows carried in reduced form): row-at-a-time tiers first
  // every rank of the communicator makes this call with the same system and the same rows (Schur complement inside
  // spasm_echelonize; kernel / rref / gesv when the host opted in with spasm_b200_dist_shard_rows): the rows are split
  // into contiguous shares, one per rank, and the results all-gathered (counts first, then the payload)
  bool collective = false;
};

// what to emit for each solved row
struct SolveEmit {
  bool count_only = false;         // only cnt[k] (density estimate)
  bool all_columns = false;        // emit every nonzero entry (kernel) instead of the non-pivotal ones
  bool structural = false;         // follow zero multipliers too; emit pattern incl. zeros (triangular_solve ABI)
  const int *colmap = nullptr;     // emitted column = colmap[c]
  const int *prefix_col = nullptr; // [nrows] optional leading entry (column, prefix_val)
  uint32_t prefix_val = 0;
  bool want_L = false;             // also emit the multipliers (prio, value) by increasing prio
};

// result: rows in input order, entries by increasing (mapped) column (normalisation N1)
struct SolveResult {
  DBuf<int> cnt;        // [nrows+1]
  DBuf<long long> p;    // [nrows+1] row pointers (exclusive scan of cnt)
  DBuf<int> j;
  DBuf<uint32_t> x;
  long long nnz = 0;
  // L stream
  DBuf<int> lcnt;
  DBuf<long long> lp;
  DBuf<int> lj;
  DBuf<uint32_t> lx;
  long long lnnz = 0;
  WorkStats stats;
};

// heavy rows (more distinct columns than the shared-memory tiers hold) are delegated to this
// callback: it must fill cnt/j/x for the listed k (dense engine, solve_dense.cu)
void solve_rows(const SolveSystem &G, const SolveRows &B, const SolveEmit &E, const Fp &F, SolveResult &R);
// the same on this rank alone, whatever B.collective says
void solve_rows_local(const SolveSystem &G, const SolveRows &B, const SolveEmit &E, const Fp &F, SolveResult &R);
// {calls that were split over the ranks, rows this rank solved in them} since the last reset
extern long long g_shard_stats[2];
// every row of a device CSR sorted by column index
void sort_csr_rows(const long long *p, int n, int *j, uint32_t *x);

}  // namespace sb
