// dense.cuh — dense tail (dense.cu): replaces spasm_schur_dense + spasm_ffpack_rref
// (prototypes src/SpaSM.jl:765-766, :805; algorithm SURVEY.md A.7)
#pragma once
#include "factor.cuh"

namespace sb {

// in-place reduced row echelon form of the row-major n x m matrix S (u32 residues, leading
// dimension ld).  Step s finds pivot column pivcol[s] (increasing: column rank profile) held by
// row pivrow[s].  Returns the rank.
int dense_rref_device(uint32_t *S, int n, int m, long long ld, const Fp &F, DBuf<int> &pivcol, DBuf<int> &pivrow);

// dense Schur complement of the rows `rows` of A w.r.t. U, stored transposed (dense_engine.cu):
// Dt[c][k] = entry of remaining row k on free column q0[c]; leading dimension ld (k contiguous)
struct DenseSchur {
  int n_rem = 0, Sm0 = 0, levels = 0;
  long long ld = 0;
  DBuf<int> q0;
  DBuf<uint32_t> Dt;
  DBuf<uint32_t> Vp;  // optional: multipliers Vp[i][k] of U row i (kept when asked), leading dimension ldv
  long long ldv = 0;
};
// work (optional, device): [0] += algorithmic bytes, [1] += multiply-adds of the row eliminations this stands for
void build_dense_schur_raw(const long long *Ap, const int *Aj, const uint32_t *Ax, int m, const int *rows, int nrows, const DCsr &U,
                           const int *Uqinv, const Fp &F, DenseSchur &D, bool keep_pivot_part, unsigned long long *work = nullptr);
// columns of M (nvec vectors of length nrows, leading dimension ld) -> sorted sparse rows for the rows todo[off..off+nrows)
void dense_rows_to_sparse(const uint32_t *M, long long ld, int nvec, const int *label, int nrows, const int *todo, int off, int *cnt,
                          unsigned long long *offs, unsigned long long slab_tag, DBuf<int> &oj, DBuf<uint32_t> &ox);
void build_dense_schur(const DCsr &A, const int *rows, int nrows, const DCsr &U, const int *Uqinv, const Fp &F, DenseSchur &D);

// C (M x N, row-major, ldc) = [C -] A (M x K, row-major lda) . B^T (B is N x K, row-major ldb)   (mod p)
// rowmap (device, M entries, optional): row r of the product reads row rowmap[r] of A and updates row rowmap[r] of C
void gemm_nt(uint32_t *C, long long ldc, int M, int N, const uint32_t *A, long long lda, const uint32_t *B, long long ldb, int K,
             bool subtract, const Fp &F, const int *rowmap = nullptr);

// largest K one tcgen05 launch takes for this prime (int32 accumulator bound of the 2 / 3 / 4 limb kernels)
int gemm_max_k(const Fp &F);

// the knobs of EchelonizeOpts (src/SpaSM.jl:325-343) the dense loop looks at: a block of Sn rows that yields fewer than
// low_rank_ratio * Sn pivots hands the remaining rows to the low-rank mode
// with-L mode of the dense tail (replaces spasm_ffpack_LU, src/SpaSM.jl:806): receives the multipliers and the row permutation
struct LSink {
  virtual ~LSink() {}
  // sparse rows of L for the tail rows [row0, row0 + nrows) (positions in the tail's row list): row i has cnt[i] entries
  // (label, value) starting at offs[i] in (oj, ox); the entry belongs to column ubase + label of L.  Device pointers.
  virtual void rows(int row0, int nrows, int ubase, const int *cnt, const unsigned long long *offs, const int *oj, const uint32_t *ox) = 0;
  // U rows [ubase, ubase + rr) were produced by the tail rows row0 + pivrow[s] (host array)
  virtual void pivots(int ubase, const int *pivrow, int rr, int row0) = 0;
};
struct TailOpts {
  bool tall_skinny = false;
  double low_rank_ratio = 0.5;
  double start_weight = -1;
  LSink *lsink = nullptr;  // not null: row echelon form + L instead of the reduced form (every rank computes everything)
};
// deferred trailing updates of the dense tail: my panels per flush, capacity of the accumulators (dense.cu: plan_tail)
struct TailPlan {
  bool lazy = false;
  int group = 1, kdepth = 0;
  long long LDK = 0;
};
TailPlan plan_tail(int Sm0, int n_local, int block_size, int Bmax, int NR, int kcap, int max_k, size_t free_bytes);
// eliminate the rows `rows` of A against U block by block, RREF each block, append to U
void echelonize_dense_device(const DCsr &A, const int *rows, int nrows, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size,
                             const TailOpts &opts = TailOpts());
// low-rank / tall-and-skinny mode: blocks of random combinations of ALL remaining rows
void echelonize_lowrank_device(const DCsr &A, const int *rows, int nrows, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size,
                               double start_weight);

}  // namespace sb
