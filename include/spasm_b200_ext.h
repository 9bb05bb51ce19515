/*
 * spasm_b200_ext.h — entry points of libspasm_b200.so that libspasm does NOT have: memory policy, multi-GPU
 * bootstrap, device-resident timing and instrumentation.  A SpaSM.jl user never needs them; bench.py, the tests
 * and a multi-GPU host do.  (The drop-in ABI proper — every symbol SpaSM.jl binds on the echelonization path — is
 * include/spasm_b200.h.)  Plain C types only; all functions are safe to call from the thread that calls the library.
 */
#ifndef SPASM_B200_EXT_H
#define SPASM_B200_EXT_H
#include "spasm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- memory policy (csrc/runtime.cu).  Default: the library holds device memory only while one of its entry points
 * runs; everything is returned to the driver when the call returns (a second process / CUDA.jl finds the HBM free).
 * set_cache(1) (or SPASM_B200_KEEP_CACHE=1) keeps the large blocks and the pool between calls: same-shaped calls in a
 * loop allocate nothing.  trim() gives them back at once; cached_bytes() reports what is still held. */
void spasm_b200_set_cache(int keep);
void spasm_b200_trim(void);
long long spasm_b200_cached_bytes(void);

/* ---- multi-GPU: one process per GPU (csrc/dist.cu).  Rank 0 creates the 128-byte NCCL id, the host framework ships it
 * to the other ranks (bench.py uses torch.distributed for that), every rank calls dist_init; spasm_echelonize then
 * shards the dense tail over the ranks (libnccl.so.2 is dlopen'ed, never linked).  shard_factor(0) (default): rank 0
 * returns the complete factor, the other ranks a factor flagged `partial` (r, qinv only).  shard_factor(1): every rank
 * returns the rows of U of the dense panels it owns (all flagged `partial`); the download scales with 1/N. */
int spasm_b200_nccl_unique_id(unsigned char *out128);
int spasm_b200_dist_init(int rank, int nranks, const unsigned char *id128);
void spasm_b200_dist_shard_factor(int on);
/* The rows of the row engine are split over the ranks (contiguous shares, counts all-gathered, entries broadcast into place):
 * always for the sparse Schur complements inside spasm_echelonize (src/SpaSM.jl:761-762; that call is collective already);
 * for spasm_kernel (free columns, :876-882), spasm_rref (rows of U, :871) and spasm_gesv (right-hand sides, :915-923) after
 * shard_rows(1) — those calls then become collective: every rank makes them, with the same complete factor, and every
 * rank receives the complete, single-GPU-identical result. */
void spasm_b200_dist_shard_rows(int on);
void spasm_b200_shard_stats(long long *out2, int reset); /* row-engine calls split over the ranks, rows this rank solved in them */
void spasm_b200_dist_finalize(void);
/* host logic of the block-cyclic panel ownership (exported for the CPU / gloo tests) */
long long spasm_b200_local_positions(long long n_rem, int block, int nranks, int rank, int *out, long long cap);
int spasm_b200_panel_owner(long long b, int nranks);
void spasm_b200_tail_plan(int Sm0, int n_local, int block_size, int nranks, int kcap, int max_k, long long free_bytes,
                          long long *out4); /* deferred updates of the dense tail: {lazy, my panels per flush, accumulator depth, its leading dimension} */
void spasm_b200_row_share(long long nrows, int nranks, int rank, long long *lo, long long *hi); /* share of the split row engine */

/* ---- device-resident timing (bench.py `value`): the CSR is uploaded once; each call echelonizes from HBM, leaves the
 * factor on the device and returns the rank (ms = CUDA-event time of the call on the library's stream) */
void *spasm_b200_upload(const struct spasm_csr *A);
int spasm_b200_echelonize_resident(void *handle, struct echelonize_opts *opts, double *ms);
void spasm_b200_release(void *handle);

/* ---- instrumentation */
void spasm_b200_last_timings(double *out16);         /* phases of the last spasm_echelonize, seconds (bench.TIMING_NAMES) */
void spasm_b200_last_stats(long long *out7);          /* last Schur / kernel solve: bytes, MACs, rows, smem / global / dense rows, us */
void spasm_b200_mma_timing(int on);                   /* CUDA events around every tcgen05 launch (read lazily, no host stall) */
void spasm_b200_mma_stats(double *out4, int reset);   /* ms in k_gemm_i8limb, modular MACs, launches, kernels launched by the library */
void spasm_b200_tail_stats(long long *out4, int reset); /* deferred trailing updates: far flushes, multiplier corrections, far rows x depth, near updates */
long long spasm_b200_lowrank_switches(int reset);      /* how often the dense loop handed its remaining rows to the low-rank mode */
double spasm_b200_gemm_probe(long long prime, int M, int N, int K, int reps); /* ms of k_gemm_i8limb alone on the GPU for C (M x N) -= A . B^T of depth K */
double spasm_b200_utcimma_peak(int iters, int reps);  /* measured back-to-back tcgen05.mma.kind::i8 M128 N256 K32 rate, TOP/s */

/* ---- triplets -> CSR on the device (csrc/compress.cu; the role of spasm_compress, src/SpaSM.jl:479-493, which stays host
 * code as in the reference): rows in the order of the triplet list, duplicates summed into the first occurrence, zero sums
 * dropped — the same arrays as spasm_compress, bit for bit.  NULL (message on stderr) without a GPU. */
struct spasm_csr *spasm_b200_compress(const struct spasm_triplet *T);

/* ---- bench / test hooks */
/* BASELINE configs[3]: n x m matrix mod prime generated on the device (iid, or of planted rank r), through the blocked dense tail */
int spasm_b200_dense_tail_bench(long long prime, int n, int m, int block_size, unsigned long long seed, double *ms);
int spasm_b200_dense_tail_bench_planted(long long prime, int n, int m, int r, int block_size, unsigned long long seed, double *ms);
/* C = [C -] A . B^T mod prime on host arrays of residues; path 0 = as the library chooses, 1 = CUDA cores only.
 * Returns 1 when the tcgen05 kernel ran, 0 when not, < 0 on error */
int spasm_b200_gemm_nt_host(long long prime, int M, int N, int K, const unsigned *A, const unsigned *B, unsigned *C, int subtract, int path,
                            double *ms_out);

#ifdef __cplusplus
}
#endif
#endif
