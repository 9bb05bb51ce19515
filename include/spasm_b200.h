/*
 * spasm_b200.h — C ABI of the B200-native SpaSM echelonization hot path.
 *
 * This is the drop-in boundary: every symbol below is what SpaSM.jl binds with
 * `@ccall spasm_lib.<sym>(...)` (reference: /root/reference/src/SpaSM.jl:14 is the
 * single binding point, `const spasm_lib = Spasm_jll.spasm`).  Struct layouts are the
 * C mirrors of the Julia structs the reference `unsafe_load`s / writes in place:
 *
 *   spasm_field      <- struct Field            src/SpaSM.jl:51-56     (32 B)
 *   spasm_csr        <- struct _CSR{F}          src/SpaSM.jl:126-134   (72 B)
 *   spasm_triplet    <- struct _Triplet{F}      src/SpaSM.jl:234-243   (80 B)
 *   spasm_lu         <- mutable struct _LU{F}   src/SpaSM.jl:262-270   (48 B)
 *   echelonize_opts  <- mutable struct EchelonizeOpts src/SpaSM.jl:325-343 (64 B)
 *
 * All pointers are plain host (malloc-family) memory; no torch / CUDA types cross
 * this boundary.  The same header is compiled by the CPU oracle (oracle/) which
 * exports the same symbols, so one harness drives both.
 */
#ifndef SPASM_B200_H
#define SPASM_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef int32_t spasm_ZZp; /* balanced representative in [mhalfp, halfp]  (src/SpaSM.jl:79-88) */

/* src/SpaSM.jl:51-56, :73-76 */
struct spasm_field_struct {
  i64 p;
  i64 halfp;
  i64 mhalfp;
  double dinvp;
};
typedef struct spasm_field_struct spasm_field[1];

/* src/SpaSM.jl:126-134 */
struct spasm_csr {
  i64 nzmax;
  int n; /* rows */
  int m; /* columns */
  i64 *p; /* n+1 row starts, 0-based */
  int *j; /* column indices, 0-based */
  spasm_ZZp *x; /* values (may be NULL: pattern only) */
  spasm_field field;
};

/* src/SpaSM.jl:234-243 */
struct spasm_triplet {
  i64 nzmax;
  i64 nz;
  int n;
  int m;
  int *i;
  int *j;
  spasm_ZZp *x;
  spasm_field field;
};

/* src/SpaSM.jl:262-270 */
struct spasm_lu {
  int r;
  bool complete;
  unsigned char partial; /* spasm_b200 only (padding byte 5 of libspasm's struct, which never reads it): non-zero when this
                            process holds only SOME rows of U (multi-GPU run): kernel / rref / solve refuse such a factor */
  struct spasm_csr *L; /* NULL unless opts->L */
  struct spasm_csr *U; /* r x m, unit pivots, pivot entry stored first in each row */
  int *qinv; /* [m] column -> U row, or -1 */
  int *p; /* [>= n] U row -> original row of A (only with L) */
  struct spasm_triplet *Ltmp;
};

/* src/SpaSM.jl:325-343 */
struct echelonize_opts {
  bool enable_greedy_pivot_search;
  bool enable_tall_and_skinny;
  bool enable_dense;
  bool enable_GPLU;
  bool L;
  bool complete;
  double min_pivot_proportion;
  int max_round;
  double sparsity_threshold;
  int dense_block_size; /* Julia declares Int64 at the same offset (low word) */
  double low_rank_ratio;
  double tall_and_skinny_ratio;
  double low_rank_start_weight;
};

/* src/SpaSM.jl:34-46: data symbol poked through cglobal; NULL = print to stderr */
extern int (*logcallback)(const char *);

/* ---- spasm_ZZp.c  (src/SpaSM.jl:49, :383-390) ---- */
void spasm_field_init(i64 p, spasm_field F);
spasm_ZZp spasm_ZZp_init(const spasm_field F, i64 x);
spasm_ZZp spasm_ZZp_add(const spasm_field F, spasm_ZZp a, spasm_ZZp b);
spasm_ZZp spasm_ZZp_sub(const spasm_field F, spasm_ZZp a, spasm_ZZp b);
spasm_ZZp spasm_ZZp_mul(const spasm_field F, spasm_ZZp a, spasm_ZZp b);
spasm_ZZp spasm_ZZp_inverse(const spasm_field F, spasm_ZZp a);
spasm_ZZp spasm_ZZp_axpy(const spasm_field F, spasm_ZZp a, spasm_ZZp x, spasm_ZZp y);

/* ---- spasm_util.c  (src/SpaSM.jl:430-475) ---- */
double spasm_wtime(void);
i64 spasm_nnz(const struct spasm_csr *A);
void *spasm_malloc(i64 size);
void *spasm_calloc(i64 count, i64 size);
void *spasm_realloc(void *ptr, i64 size);
struct spasm_csr *spasm_csr_alloc(int n, int m, i64 nzmax, i64 prime, bool with_values); /* :441 */
void spasm_csr_realloc(struct spasm_csr *A, i64 nzmax); /* :447  (nzmax < 0: shrink to fit) */
void spasm_csr_resize(struct spasm_csr *A, int n, int m); /* :449 */
void spasm_csr_free(struct spasm_csr *A); /* :451 */
struct spasm_triplet *spasm_triplet_alloc(int n, int m, i64 nzmax, i64 prime, bool with_values); /* :453 */
void spasm_triplet_realloc(struct spasm_triplet *A, i64 nzmax); /* :455 */
void spasm_triplet_free(struct spasm_triplet *A); /* :457 */
void spasm_lu_free(struct spasm_lu *N); /* :463 */
int spasm_get_num_threads(void); /* :470 */
int spasm_get_thread_num(void); /* :475 */

/* ---- spasm_triplet.c  (src/SpaSM.jl:482-493) ---- */
void spasm_add_entry(struct spasm_triplet *T, int i, int j, i64 x);
void spasm_triplet_transpose(struct spasm_triplet *T);
struct spasm_csr *spasm_compress(const struct spasm_triplet *T);

/* ---- spasm_io.c  (src/SpaSM.jl:498-529): SMS text format ---- */
struct spasm_triplet *spasm_triplet_load(FILE *f, i64 prime, u8 *hash);
void spasm_triplet_save(const struct spasm_triplet *A, FILE *f);
void spasm_csr_save(const struct spasm_csr *A, FILE *f);

/* ---- spasm_transpose.c  (src/SpaSM.jl:589) — called with ONE argument; values always kept ---- */
struct spasm_csr *spasm_transpose(const struct spasm_csr *A);

/* ---- spasm_scatter.c / spasm_spmv.c  (src/SpaSM.jl:620, :643, :656) ---- */
void spasm_scatter(const struct spasm_csr *A, int i, spasm_ZZp beta, spasm_ZZp *x); /* x += beta*A[i] */
void spasm_xApy(const spasm_ZZp *x, const struct spasm_csr *A, spasm_ZZp *y); /* y += x.A */
void spasm_Axpy(const struct spasm_csr *A, const spasm_ZZp *x, spasm_ZZp *y); /* y += A.x */

/* ---- spasm_reach.c / spasm_triangular.c  (src/SpaSM.jl:627-628, :673-722) ---- */
/* spasm_dfs / spasm_reach are quoted but NOT bound by the reference ("we don't expose these",
 * src/SpaSM.jl:625-629): they are internal to the CPU algorithm and have no GPU counterpart. */
int spasm_sparse_triangular_solve(const struct spasm_csr *U, const struct spasm_csr *B, int k, int *xj,
                                  spasm_ZZp *x, const int *qinv); /* :694-722 */
bool spasm_dense_back_solve(const struct spasm_csr *L, spasm_ZZp *b, spasm_ZZp *x, const int *p); /* :673 */
bool spasm_dense_forward_solve(const struct spasm_csr *U, spasm_ZZp *b, spasm_ZZp *x, const int *q); /* :688 */

/* ---- spasm_pivots.c / spasm_schur.c  (prototypes quoted at src/SpaSM.jl:761-778) ---- */
int spasm_pivots_extract_structural(const struct spasm_csr *A, const int *p_in, struct spasm_lu *fact, int *p,
                                    struct echelonize_opts *opts);
double spasm_schur_estimate_density(const struct spasm_csr *A, const int *p, int n, const struct spasm_csr *U,
                                    const int *qinv, int R);
struct spasm_csr *spasm_schur(const struct spasm_csr *A, const int *p, int n, const struct spasm_lu *fact,
                              double est_density, struct spasm_triplet *L, const int *p_in, int *p_out);

/* ---- spasm_echelonize.c  (src/SpaSM.jl:817, :863) ---- */
void spasm_echelonize_init_opts(struct echelonize_opts *opts);
struct spasm_lu *spasm_echelonize(const struct spasm_csr *A, struct echelonize_opts *opts);

/* ---- spasm_rref.c / spasm_kernel.c / spasm_solve.c  (src/SpaSM.jl:871, :879, :903, :920) ---- */
struct spasm_csr *spasm_rref(const struct spasm_lu *fact, int *Rqinv);
struct spasm_csr *spasm_kernel(const struct spasm_lu *fact);
bool spasm_solve(const struct spasm_lu *fact, const spasm_ZZp *b, spasm_ZZp *x);
struct spasm_csr *spasm_gesv(const struct spasm_lu *fact, const struct spasm_csr *B, bool *ok);

/* ---- dense tail entry point (replaces spasm_ffpack_rref, prototype quoted at src/SpaSM.jl:805) ----
 * In-place RREF of a row-major n x m matrix of balanced int32 residues mod `prime`.
 * On return the first `rank` rows hold the reduced rows (row i has its pivot, equal to 1, in
 * column pivcol[i]; pivcol is increasing = column rank profile).  Returns the rank. */
int spasm_dense_rref(i64 prime, int n, int m, spasm_ZZp *A, i64 ldA, int *pivcol);

/* ---- library identification and instrumentation (not in libspasm) ---- */
const char *spasm_b200_backend(void); /* "cuda-sm_100a" for the product, "cpu-oracle" for oracle/ */
void spasm_b200_seed(u64 seed);       /* re-seed the PRNG of the density estimate (normalisation N4) */

#ifdef __cplusplus
}
#endif
#endif
