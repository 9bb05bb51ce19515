// common.cuh — shared device/host helpers of the B200 SpaSM library.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "spasm_b200.h"

namespace sb {

// ------------------------------------------------------------------ errors / logging
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

extern const char *g_phase;  // what the library was doing (for error messages)
void logf(const char *fmt, ...);  // -> logcallback or stderr (host_abi.cu)
void errf(const char *fmt, ...);  // errors: always stderr, and the callback when installed

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      throw sb::Error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                      std::to_string(__LINE__));                                                   \
  } while (0)

// The product has no CPU fallback: every compute entry point calls this first.
void require_gpu();
// RAII scope of one public compute entry point: require_gpu() on entry; when the outermost scope ends the
// cached device memory goes back to the driver (runtime.cu, "Memory policy").  Declare it FIRST in the entry
// point so that every DBuf of the call is released before it.
struct ApiCall {
  ApiCall();
  ~ApiCall();
  ApiCall(const ApiCall &) = delete;
  ApiCall &operator=(const ApiCall &) = delete;
};
cudaStream_t stream();       // the stream the library's launches, copies and stream-ordered allocations go to right now
cudaStream_t main_stream();  // the library stream (highest priority)
cudaStream_t aux_stream();   // second, low-priority stream: the deferred far-row updates of the dense tail run there
// while alive, stream() is `s` (work of a whole code path — kernels, temporaries, frees — moves to that stream)
struct StreamScope {
  cudaStream_t prev;
  explicit StreamScope(cudaStream_t s);
  ~StreamScope();
  StreamScope(const StreamScope &) = delete;
  StreamScope &operator=(const StreamScope &) = delete;
};
int sm_count();

// ------------------------------------------------------------------ field
// Residues live on the device as u32 in [0,p); the ABI's balanced int32 form
// (src/SpaSM.jl:79-88) is converted at load / store.  Any exact reduction gives the same
// bits as the reference's double-estimate reduction (src/SpaSM.jl:385-390) once balanced.
struct Fp {
  uint32_t p;
  uint32_t half;  // p/2: u > half  <=>  balanced value is u - p
  uint32_t M32;   // floor(2^32/p)  (products < 2^32, p < 2^16)
  uint64_t M64;   // floor(2^64/p)
  bool small;     // p < 2^16
};
Fp make_field(int64_t p);

__host__ __device__ inline uint32_t to_u(int32_t v, const Fp &F) { return v < 0 ? (uint32_t)((int64_t)v + F.p) : (uint32_t)v; }
__host__ __device__ inline int32_t to_bal(uint32_t u, const Fp &F) { return u > F.half ? (int32_t)(u - F.p) : (int32_t)u; }

template <bool SMALL>
__device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b, const Fp &F) {
  if (SMALL) {
    uint32_t t = a * b;
    uint32_t q = __umulhi(t, F.M32);
    uint32_t r = t - q * F.p;
    if (r >= F.p) r -= F.p;
    return r;
  } else {
    uint64_t t = (uint64_t)a * b;
    uint64_t q = __umul64hi(t, F.M64);
    uint64_t r = t - q * F.p;
    if (r >= F.p) r -= F.p;
    if (r >= F.p) r -= F.p;
    return (uint32_t)r;
  }
}
__device__ __forceinline__ uint32_t addmod(uint32_t a, uint32_t b, const Fp &F) {
  uint32_t s = a + b;
  if (s < a || s >= F.p) s -= F.p;
  return s;
}
__device__ __forceinline__ uint32_t negmod(uint32_t a, const Fp &F) { return a ? F.p - a : 0u; }
// reduce a u64 (any value) mod p
__device__ __forceinline__ uint32_t red64(uint64_t t, const Fp &F) {
  uint64_t q = __umul64hi(t, F.M64);
  uint64_t r = t - q * F.p;
  if (r >= F.p) r -= F.p;
  if (r >= F.p) r -= F.p;
  return (uint32_t)r;
}
uint32_t host_inv(uint32_t a, uint32_t p);
// a^(p-2) mod p for p < 2^16: 32-bit Barrett multiplications only (no 64-bit divisions)
__device__ inline uint32_t dev_inv_small(uint32_t a, const Fp &F) {
  uint32_t result = 1, base = a, e = F.p - 2;
  while (e) {
    if (e & 1) result = mulmod<true>(result, base, F);
    base = mulmod<true>(base, base, F);
    e >>= 1;
  }
  return result;
}
__device__ inline uint32_t dev_inv(uint32_t a, uint32_t p) {  // extended Euclid, a in (0,p)
  int64_t r0 = a, r1 = p, s0 = 1, s1 = 0;
  while (r1 != 0) {
    int64_t q = r0 / r1, t = r0 - q * r1;
    r0 = r1;
    r1 = t;
    t = s0 - q * s1;
    s0 = s1;
    s1 = t;
  }
  s0 %= (int64_t)p;
  if (s0 < 0) s0 += p;
  return (uint32_t)s0;
}

// ------------------------------------------------------------------ device memory
// stream-ordered allocations from the default pool (no cudaFree synchronisation)
void *dmalloc_bytes(size_t bytes);
// cache mode only: pinned host blocks for large result arrays (runtime.cu); nullptr = use malloc
bool cache_enabled();
void *host_big_alloc(size_t bytes);
size_t host_big_capacity(const void *p);
bool host_big_release(void *p);
void dfree(void *p);
size_t dev_free_bytes();

template <class T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  DBuf() = default;
  explicit DBuf(size_t n_) { alloc(n_); }
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr, o.n = 0; }
  DBuf &operator=(DBuf &&o) noexcept {
    if (this != &o) {
      release();
      p = o.p, n = o.n;
      o.p = nullptr, o.n = 0;
    }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t n_) {
    release();
    n = n_;
    p = (T *)dmalloc_bytes((n_ ? n_ : 1) * sizeof(T));
  }
  void release() {
    if (p) dfree(p);
    p = nullptr, n = 0;
  }
  void zero() { CK(cudaMemsetAsync(p, 0, n * sizeof(T), stream())); }
  void fill_ff() { CK(cudaMemsetAsync(p, 0xff, n * sizeof(T), stream())); }
  void upload(const T *h, size_t cnt) { CK(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, stream())); }
  void download(T *h, size_t cnt) const { CK(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, stream())); }
  void grow(size_t need) {  // keep contents
    if (need <= n) return;
    size_t nn = (need > ((size_t)1 << 28)) ? need + need / 8 : need + need / 2 + 1024;
    T *q = (T *)dmalloc_bytes(nn * sizeof(T));
    if (p) {
      CK(cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, stream()));
      dfree(p);
    }
    p = q, n = nn;
  }
};

void sync();
template <class T>
T fetch(const T *dptr) {
  T h;
  CK(cudaMemcpyAsync(&h, dptr, sizeof(T), cudaMemcpyDeviceToHost, stream()));
  sync();
  return h;
}

// ------------------------------------------------------------------ device CSR (values as u32 residues)
struct DCsr {
  int n = 0, m = 0;
  int64_t nnz = 0;
  DBuf<long long> p;  // n+1
  DBuf<int> j;
  DBuf<uint32_t> x;
};
void upload_csr(const spasm_csr *A, DCsr &D, const Fp &F);          // balanced -> u32
spasm_csr *download_csr(const DCsr &D, int64_t prime, const Fp &F); // u32 -> balanced, malloc'd host CSR
void convert_to_balanced(const uint32_t *in, int *out, long long n, const Fp &F);
// device -> pageable host through pinned bounce buffers, pages touched by all host threads
void download_large(void *dst, const void *src_dev, size_t bytes);
void convert_to_residues(const int *in, uint32_t *out, long long n, const Fp &F);

// exclusive scan helpers (cub underneath; scan.cu)
void exclusive_scan_i64(const long long *in, long long *out, size_t n);  // out[n] NOT written
void exclusive_scan_i32_to_i64(const int *in, long long *out, size_t n_plus_one);  // out[0..n], in[n] ignored
long long reduce_sum_i32(const int *in, size_t n);

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// work counters (algorithmic bytes / MACs per SURVEY.md §8d) of the last solve-type call
extern long long g_launches;  // kernels launched by this library (counted at the call sites of the hot phases)

struct WorkStats {
  long long bytes = 0, macs = 0, rows = 0, light = 0, medium = 0, heavy = 0;
  double ms = 0;
};

}  // namespace sb
