/* C test driver: replays SpaSM.jl's call sequences against ANY library that exports the
 * ABI of include/spasm_b200.h — the CPU oracle or the CUDA product — without Python in between.
 *   driver <lib.so> <case>      case = runtests | runtests_t | readme
 * Prints the rank and the kernel basis in a canonical text form (one "i j v" line per entry, 0-based,
 * v in [0,p), rows in order, entries by column).  tests/test_c_driver.py compares it with the reference's
 * golden vectors (test/runtests.jl:7-24, README.md:9-48) and the two libraries with each other. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spasm_b200.h"

#define LOAD(name)                                             \
  __typeof__(name) *name##_ = (__typeof__(name) *)dlsym(h, #name); \
  if (!name##_) {                                              \
    fprintf(stderr, "missing symbol %s\n", #name);             \
    return 3;                                                  \
  }

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <lib.so> <runtests|runtests_t|readme>\n", argv[0]);
    return 2;
  }
  void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "dlopen: %s\n", dlerror());
    return 3;
  }
  LOAD(spasm_csr_alloc)
  LOAD(spasm_csr_free)
  LOAD(spasm_transpose)
  LOAD(spasm_echelonize_init_opts)
  LOAD(spasm_echelonize)
  LOAD(spasm_kernel)
  LOAD(spasm_lu_free)
  LOAD(spasm_nnz)
  const i64 prime = 42013;
  /* Julia sparse(I, J, V) entries; CSR(m) stores the TRANSPOSE: a Julia column is a SpaSM row (src/SpaSM.jl:941-968) */
  int I[4], J[4], V[4], rows, cols;
  if (!strcmp(argv[2], "readme")) {
    int i[4] = {1, 1, 2, 2}, j[4] = {1, 2, 1, 2}, v[4] = {1, 2, 3, 6};
    memcpy(I, i, sizeof I), memcpy(J, j, sizeof J), memcpy(V, v, sizeof V);
    rows = 2, cols = 2;
  } else {
    int i[4] = {1, 1, 3, 3}, j[4] = {1, 2, 3, 4}, v[4] = {1, 2, 3, 4};
    memcpy(I, i, sizeof I), memcpy(J, j, sizeof J), memcpy(V, v, sizeof V);
    rows = 3, cols = 4;
  }
  struct spasm_csr *A = spasm_csr_alloc_(cols, rows, 4, prime, true); /* cols x rows: the transpose */
  i64 nz = 0;
  for (int c = 1; c <= cols; c++) { /* SpaSM row c-1 = Julia column c, entries by increasing Julia row */
    A->p[c - 1] = nz;
    for (int k = 0; k < 4; k++)
      if (J[k] == c) A->j[nz] = I[k] - 1, A->x[nz] = V[k], nz++;
  }
  A->p[cols] = nz;
  struct spasm_csr *B = A;
  if (!strcmp(argv[2], "runtests_t")) B = spasm_transpose_(A); /* kernel(transpose(sm)), test/runtests.jl:23 */
  struct echelonize_opts opts;
  spasm_echelonize_init_opts_(&opts);
  struct spasm_lu *fact = spasm_echelonize_(B, &opts);
  if (!fact) {
    fprintf(stderr, "echelonize returned NULL\n");
    return 4;
  }
  printf("rank %d\n", fact->r);
  printf("nnzU %lld\n", (long long)spasm_nnz_(fact->U));
  struct spasm_csr *K = spasm_kernel_(fact);
  if (!K) {
    fprintf(stderr, "kernel returned NULL\n");
    return 4;
  }
  printf("kernel %d %d %lld\n", K->n, K->m, (long long)spasm_nnz_(K));
  for (int i = 0; i < K->n; i++) {
    int len = (int)(K->p[i + 1] - K->p[i]);
    int *ord = (int *)malloc(sizeof(int) * (len > 0 ? len : 1));
    for (int k = 0; k < len; k++) ord[k] = K->j[K->p[i] + k];
    qsort(ord, len, sizeof(int), cmp_int);
    for (int k = 0; k < len; k++)
      for (i64 e = K->p[i]; e < K->p[i + 1]; e++)
        if (K->j[e] == ord[k]) {
          long long v = K->x[e];
          if (v < 0) v += prime;
          if (v) printf("%d %d %lld\n", i, ord[k], v);
        }
    free(ord);
  }
  spasm_csr_free_(K);
  spasm_lu_free_(fact);
  if (B != A) spasm_csr_free_(B);
  spasm_csr_free_(A);
  /* no dlclose: the library's OpenMP workers (and, for the CUDA library, the runtime's own threads) are still parked in
   * code that unloading would unmap — one run in five died with SIGSEGV at exit.  A Julia host never unloads it either. */
  (void)h;
  return 0;
}
