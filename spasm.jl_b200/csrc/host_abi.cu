// host_abi.cpp — the container half of the C ABI: allocation, triplets, SMS I/O, scalar field
// operations, logging.  These mirror what the Julia wrapper calls around the hot path
// (src/SpaSM.jl:430-529) and are plain host code by nature: SpaSM.jl WRITES the p/j/x arrays of a
// freshly allocated CSR in place (src/SpaSM.jl:949-966), so they must be malloc-family memory.
// No matrix arithmetic happens here — that is all CUDA (solve_*.cu, pivots.cu, dense*.cu).
#include <omp.h>
#include <sys/time.h>

#include <cinttypes>

#include <algorithm>

#include "common.cuh"

extern "C" {
int (*logcallback)(const char *) = nullptr;
const char *spasm_b200_backend(void) { return "cuda-sm_100a"; }
}

namespace sb {
const char *g_phase = "";
void logf(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (logcallback != nullptr)
    logcallback(buf);  // called synchronously from the calling thread only
  else
    fputs(buf, stderr);
}

void errf(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  fputs(buf, stderr);
  if (logcallback != nullptr) logcallback(buf);
}

Fp make_field(int64_t p) {
  Fp F;
  F.p = (uint32_t)p;
  F.half = (uint32_t)(p / 2);
  F.M32 = (uint32_t)((1ULL << 32) / (uint64_t)p);
  F.M64 = (uint64_t)(((unsigned __int128)1 << 64) / (unsigned __int128)p);
  F.small = p < 65536;
  return F;
}

uint32_t host_inv(uint32_t a, uint32_t p) {
  int64_t r0 = a, r1 = p, s0 = 1, s1 = 0;
  while (r1 != 0) {
    int64_t q = r0 / r1, t = r0 - q * r1;
    r0 = r1, r1 = t;
    t = s0 - q * s1;
    s0 = s1, s1 = t;
  }
  s0 %= (int64_t)p;
  if (s0 < 0) s0 += p;
  return (uint32_t)s0;
}
}  // namespace sb

static inline spasm_ZZp norm_(const spasm_field F, i64 x) {
  if (x < F->mhalfp) return (spasm_ZZp)(x + F->p);
  if (x > F->halfp) return (spasm_ZZp)(x - F->p);
  return (spasm_ZZp)x;
}

extern "C" {

// ---- scalar field ops: same results as src/SpaSM.jl:383-390 (exact integer remainder) ----
void spasm_field_init(i64 p, spasm_field F) {
  F->p = p;
  F->halfp = p / 2;
  F->mhalfp = p / 2 - p + 1;
  F->dinvp = 1.0 / (double)p;
}
spasm_ZZp spasm_ZZp_init(const spasm_field F, i64 x) { return norm_(F, x % F->p); }
spasm_ZZp spasm_ZZp_add(const spasm_field F, spasm_ZZp a, spasm_ZZp b) { return norm_(F, (i64)a + b); }
spasm_ZZp spasm_ZZp_sub(const spasm_field F, spasm_ZZp a, spasm_ZZp b) { return norm_(F, (i64)a - b); }
spasm_ZZp spasm_ZZp_mul(const spasm_field F, spasm_ZZp a, spasm_ZZp b) { return norm_(F, ((i64)a * b) % F->p); }
spasm_ZZp spasm_ZZp_axpy(const spasm_field F, spasm_ZZp a, spasm_ZZp x, spasm_ZZp y) {
  return norm_(F, ((i64)a * x + y) % F->p);
}
spasm_ZZp spasm_ZZp_inverse(const spasm_field F, spasm_ZZp a) {
  uint32_t u = a < 0 ? (uint32_t)((i64)a + F->p) : (uint32_t)a;
  return norm_(F, (i64)sb::host_inv(u, (uint32_t)F->p));
}

// ---- util ----
double spasm_wtime(void) {
  struct timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}
int spasm_get_num_threads(void) { return omp_get_max_threads(); }
int spasm_get_thread_num(void) { return omp_get_thread_num(); }

static void *must(void *q) {
  if (q == nullptr) {
    sb::logf("[spasm_b200] host allocation failed\n");
    abort();  // same contract as libspasm's spasm_malloc
  }
  return q;
}
void *spasm_malloc(i64 size) { return must(malloc(size > 0 ? (size_t)size : 1)); }
void *spasm_calloc(i64 count, i64 size) { return must(calloc(count > 0 ? (size_t)count : 1, size > 0 ? (size_t)size : 1)); }
void *spasm_realloc(void *ptr, i64 size) { return must(realloc(ptr, size > 0 ? (size_t)size : 1)); }

i64 spasm_nnz(const struct spasm_csr *A) { return A->p[A->n]; }

struct spasm_csr *spasm_csr_alloc(int n, int m, i64 nzmax, i64 prime, bool with_values) {
  auto *A = (struct spasm_csr *)spasm_malloc(sizeof(struct spasm_csr));
  spasm_field_init(prime, A->field);
  A->n = n, A->m = m, A->nzmax = nzmax;
  A->p = (i64 *)spasm_malloc((i64)(n + 1) * (i64)sizeof(i64));
  A->j = (int *)spasm_malloc(nzmax * (i64)sizeof(int));
  A->x = with_values ? (spasm_ZZp *)spasm_malloc(nzmax * (i64)sizeof(spasm_ZZp)) : nullptr;
  A->p[0] = 0;
  return A;
}
// arrays handed out by host_big_alloc (cache mode) keep their capacity; growing them moves to a malloc'ed array
static void *regrow(void *old, i64 new_bytes, i64 keep_bytes) {
  const size_t cap = sb::host_big_capacity(old);
  if (cap == 0) return spasm_realloc(old, new_bytes);
  if ((size_t)new_bytes <= cap) return old;
  void *q = spasm_malloc(new_bytes);
  memcpy(q, old, (size_t)std::min<i64>(keep_bytes, new_bytes));
  sb::host_big_release(old);
  return q;
}
void spasm_csr_realloc(struct spasm_csr *A, i64 nzmax) {
  if (nzmax < 0) nzmax = spasm_nnz(A);
  if (nzmax == A->nzmax) return;
  if (sb::host_big_capacity(A->j) != 0 || (A->x && sb::host_big_capacity(A->x) != 0)) {
    const i64 keep = std::min(nzmax, A->nzmax) * (i64)sizeof(int);
    A->j = (int *)regrow(A->j, nzmax * (i64)sizeof(int), keep);
    if (A->x) A->x = (spasm_ZZp *)regrow(A->x, nzmax * (i64)sizeof(spasm_ZZp), keep);
    A->nzmax = nzmax;
    return;
  }
  A->j = (int *)spasm_realloc(A->j, nzmax * (i64)sizeof(int));
  if (A->x) A->x = (spasm_ZZp *)spasm_realloc(A->x, nzmax * (i64)sizeof(spasm_ZZp));
  A->nzmax = nzmax;
}
void spasm_csr_resize(struct spasm_csr *A, int n, int m) {
  A->m = m;
  if (n > A->n) {
    A->p = (i64 *)spasm_realloc(A->p, (i64)(n + 1) * (i64)sizeof(i64));
    for (int i = A->n + 1; i <= n; i++) A->p[i] = A->p[A->n];
  }
  A->n = n;
}
void spasm_csr_free(struct spasm_csr *A) {
  if (!A) return;
  free(A->p);
  if (!sb::host_big_release(A->j)) free(A->j);
  if (!sb::host_big_release(A->x)) free(A->x);
  free(A);
}
struct spasm_triplet *spasm_triplet_alloc(int n, int m, i64 nzmax, i64 prime, bool with_values) {
  auto *T = (struct spasm_triplet *)spasm_malloc(sizeof(struct spasm_triplet));
  spasm_field_init(prime, T->field);
  T->n = n, T->m = m, T->nzmax = nzmax, T->nz = 0;
  T->i = (int *)spasm_malloc(nzmax * (i64)sizeof(int));
  T->j = (int *)spasm_malloc(nzmax * (i64)sizeof(int));
  T->x = with_values ? (spasm_ZZp *)spasm_malloc(nzmax * (i64)sizeof(spasm_ZZp)) : nullptr;
  return T;
}
void spasm_triplet_realloc(struct spasm_triplet *T, i64 nzmax) {
  if (nzmax < 0) nzmax = T->nz;
  T->i = (int *)spasm_realloc(T->i, nzmax * (i64)sizeof(int));
  T->j = (int *)spasm_realloc(T->j, nzmax * (i64)sizeof(int));
  if (T->x) T->x = (spasm_ZZp *)spasm_realloc(T->x, nzmax * (i64)sizeof(spasm_ZZp));
  T->nzmax = nzmax;
}
void spasm_triplet_free(struct spasm_triplet *T) {
  if (!T) return;
  free(T->i), free(T->j), free(T->x), free(T);
}

// ---- triplets (src/SpaSM.jl:482-493) ----
void spasm_add_entry(struct spasm_triplet *T, int i, int j, i64 x) {
  spasm_ZZp v = spasm_ZZp_init(T->field, x);
  if (v == 0) return;
  if (T->nz == T->nzmax) spasm_triplet_realloc(T, 2 * T->nzmax + 1);
  T->i[T->nz] = i, T->j[T->nz] = j;
  if (T->x) T->x[T->nz] = v;
  T->nz++;
  if (i >= T->n) T->n = i + 1;
  if (j >= T->m) T->m = j + 1;
}
void spasm_triplet_transpose(struct spasm_triplet *T) {
  std::swap(T->i, T->j);
  std::swap(T->n, T->m);
}
// bucket by row (stable), then merge duplicate columns inside each row, dropping zero sums
struct spasm_csr *spasm_compress(const struct spasm_triplet *T) {
  const int n = T->n, m = T->m;
  const i64 nz = T->nz;
  struct spasm_csr *C = spasm_csr_alloc(n, m, nz, T->field->p, T->x != nullptr);
  std::vector<i64> cursor(n + 1, 0);
  for (i64 e = 0; e < nz; e++) cursor[T->i[e] + 1]++;
  for (int i = 0; i < n; i++) cursor[i + 1] += cursor[i];
  for (int i = 0; i <= n; i++) C->p[i] = cursor[i];
  for (i64 e = 0; e < nz; e++) {
    i64 d = cursor[T->i[e]]++;
    C->j[d] = T->j[e];
    if (C->x) C->x[d] = T->x[e];
  }
  std::vector<i64> where(m, -1);
  i64 w = 0;
  for (int i = 0; i < n; i++) {
    const i64 lo = C->p[i], hi = C->p[i + 1], base = w;
    for (i64 e = lo; e < hi; e++) {
      int c = C->j[e];
      if (where[c] >= base) {
        if (C->x) C->x[where[c]] = spasm_ZZp_add(C->field, C->x[where[c]], C->x[e]);
      } else {
        where[c] = w;
        C->j[w] = c;
        if (C->x) C->x[w] = C->x[e];
        w++;
      }
    }
    i64 keep = base;
    for (i64 e = base; e < w; e++) {
      where[C->j[e]] = -1;
      if (!C->x || C->x[e] != 0) {
        C->j[keep] = C->j[e];
        if (C->x) C->x[keep] = C->x[e];
        keep++;
      }
    }
    w = keep;
    C->p[i] = base;
  }
  C->p[n] = w;
  return C;
}

// ---- SMS text format (src/SpaSM.jl:498-529, :1029-1086): "n m M" / "i j v" 1-based / "0 0 0" ----
struct spasm_triplet *spasm_triplet_load(FILE *f, i64 prime, u8 *hash) {
  int n, m;
  char kind;
  if (fscanf(f, "%d %d %c\n", &n, &m, &kind) != 3) {
    sb::logf("[spasm_triplet_load] bad SMS header\n");
    return nullptr;
  }
  struct spasm_triplet *T = spasm_triplet_alloc(n, m, 16, prime, true);
  long long i, j, v;
  while (fscanf(f, "%lld %lld %lld\n", &i, &j, &v) == 3 && !(i == 0 && j == 0 && v == 0))
    spasm_add_entry(T, (int)(i - 1), (int)(j - 1), v);
  if (hash) memset(hash, 0, 32);  // certificates (SHA-256) are out of scope
  return T;
}
void spasm_triplet_save(const struct spasm_triplet *A, FILE *f) {
  fprintf(f, "%d %d M\n", A->n, A->m);
  for (i64 e = 0; e < A->nz; e++) fprintf(f, "%d %d %d\n", A->i[e] + 1, A->j[e] + 1, A->x ? A->x[e] : 1);
  fprintf(f, "0 0 0\n");
}
void spasm_csr_save(const struct spasm_csr *A, FILE *f) {
  fprintf(f, "%d %d M\n", A->n, A->m);
  for (int i = 0; i < A->n; i++)
    for (i64 e = A->p[i]; e < A->p[i + 1]; e++) fprintf(f, "%d %d %d\n", i + 1, A->j[e] + 1, A->x ? A->x[e] : 1);
  fprintf(f, "0 0 0\n");
}

// ---- x += beta*A[i] on host vectors (src/SpaSM.jl:620): a handful of scalar updates ----
void spasm_scatter(const struct spasm_csr *A, int i, spasm_ZZp beta, spasm_ZZp *x) {
  for (i64 e = A->p[i]; e < A->p[i + 1]; e++) x[A->j[e]] = spasm_ZZp_axpy(A->field, beta, A->x[e], x[A->j[e]]);
}

void spasm_echelonize_init_opts(struct echelonize_opts *o) {  // src/SpaSM.jl:817; defaults SURVEY.md §8a4
  o->enable_greedy_pivot_search = true;
  o->enable_tall_and_skinny = true;
  o->enable_dense = true;
  o->enable_GPLU = true;
  o->L = false;
  o->complete = false;
  o->min_pivot_proportion = 0.1;
  o->max_round = 3;
  o->sparsity_threshold = 0.05;
  o->dense_block_size = 1000;
  o->low_rank_ratio = 0.5;
  o->tall_and_skinny_ratio = 5;
  o->low_rank_start_weight = -1;
}

}  // extern "C"
