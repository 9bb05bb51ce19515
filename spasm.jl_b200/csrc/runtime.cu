// runtime.cu — device runtime: stream, stream-ordered memory, scans, CSR up/download, transpose.
#include <omp.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <mutex>

#include "common.cuh"

namespace sb {

long long g_launches = 0;
static cudaStream_t g_stream = nullptr;  // main stream
static cudaStream_t g_aux = nullptr;     // second stream, created on first use
static cudaStream_t g_cur = nullptr;     // override set by StreamScope (nullptr: the main stream)
static int g_sms = 0;
static bool g_ready = false;
// Memory policy.  By default the library holds device memory only while one of its entry points runs: a
// second process (CUDA.jl in the same Julia session, another library instance) must find the HBM free after a
// 200k x 200k echelonization.  A host that calls the library in a loop on same-shaped inputs (bench.py) may
// opt in to keeping the large blocks and the pool across calls with spasm_b200_set_cache(1) /
// SPASM_B200_KEEP_CACHE=1 and give them back with spasm_b200_trim().
static bool g_keep_cache = false;
static int g_api_depth = 0;
static const size_t POOL_KEEP_BYTES = (size_t)64 << 20;

void require_gpu() {
  if (g_ready) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error("spasm_b200: no CUDA device — this library has no CPU fallback");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) throw Error("spasm_b200: built for sm_100a (B200) only");
  g_sms = prop.multiProcessorCount;
  int prio_lo = 0, prio_hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CK(cudaStreamCreateWithPriority(&g_stream, cudaStreamNonBlocking, prio_hi));
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, dev));
  // freed blocks stay cached in the pool WHILE a call runs (thousands of stream-ordered allocations per
  // echelonization); ApiCall gives everything back to the driver when the outermost call returns.
  uint64_t thresh = UINT64_MAX;
  CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  // never let an allocation on one stream wait for the other stream's queue in order to reuse a block freed there
  // (the dense tail's main stream would stall behind the deferred updates of the second stream)
  int no = 0;
  CK(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &no));
  if (const char *e = getenv("SPASM_B200_KEEP_CACHE")) g_keep_cache = atoi(e) != 0;
  g_ready = true;
}
cudaStream_t stream() { return g_cur ? g_cur : g_stream; }
cudaStream_t main_stream() { return g_stream; }
cudaStream_t aux_stream() {
  if (!g_aux) {
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&g_aux, cudaStreamNonBlocking, prio_lo));
  }
  return g_aux;
}
StreamScope::StreamScope(cudaStream_t s) : prev(g_cur) { g_cur = s; }
StreamScope::~StreamScope() { g_cur = prev; }
int sm_count() { return g_sms; }
void sync() { CK(cudaStreamSynchronize(stream())); }

// Large blocks (>= 256 MiB: the dense Schur complement, the reserved U, the SpTRSM workspace) are
// kept in a small best-fit cache across calls: an echelonization of the same shape then allocates
// nothing.  (The stream-ordered pool alone re-maps tens of GB per call, seconds of driver time.)
struct BigBlock {
  void *p;
  size_t bytes;
  cudaStream_t last = nullptr;  // stream whose work used the block last (set when it is parked in the cache)
  cudaEvent_t ev = nullptr;     // recorded on `last` at that moment: another stream waits for it before reusing the block
};
long long g_alloc_retries = 0;  // allocations that only succeeded after the caches were emptied (a full stall of both streams)
static std::vector<BigBlock> g_big_free;
static std::vector<BigBlock> g_big_live;
static const size_t BIG = (size_t)256 << 20;

static void big_trim() {
  for (auto &b : g_big_free) {
    cudaFreeAsync(b.p, b.last ? b.last : g_stream);
    if (b.ev) cudaEventDestroy(b.ev);
  }
  g_big_free.clear();
}

static void release_cached(size_t keep) {
  if (!g_stream) return;
  big_trim();
  cudaStreamSynchronize(g_stream);
  if (g_aux) cudaStreamSynchronize(g_aux);
  cudaMemPool_t pool;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, keep);
}
ApiCall::ApiCall() {
  require_gpu();
  g_api_depth++;
}
ApiCall::~ApiCall() {
  if (--g_api_depth == 0 && !g_keep_cache) release_cached(POOL_KEEP_BYTES);
}

// ---- large HOST arrays of results, cache mode only.  A factor of tens of GB written into freshly mmap'ed pageable
// memory costs seconds of page faults per call; when the host keeps the caches (spasm_b200_set_cache(1)) the big
// result arrays are PINNED blocks that the library hands out and takes back in spasm_csr_free (the reference frees
// every object through the library's own free functions, src/SpaSM.jl:148,275,451,463), so the device writes into
// them directly and repeated calls touch no new page.  Without cache mode nothing changes: plain malloc'ed arrays.
static std::mutex g_hb_mu;
static std::vector<BigBlock> g_hb_live, g_hb_free;
bool cache_enabled() { return g_keep_cache; }
void *host_big_alloc(size_t bytes) {
  if (!g_keep_cache || bytes < BIG) return nullptr;
  {
    std::lock_guard<std::mutex> lk(g_hb_mu);
    int best = -1;
    for (int i = 0; i < (int)g_hb_free.size(); i++)
      if (g_hb_free[i].bytes >= bytes && g_hb_free[i].bytes <= bytes + bytes / 4 && (best < 0 || g_hb_free[i].bytes < g_hb_free[best].bytes)) best = i;
    if (best >= 0) {
      BigBlock b = g_hb_free[best];
      g_hb_free.erase(g_hb_free.begin() + best);
      g_hb_live.push_back(b);
      return b.p;
    }
  }
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;  // the caller falls back to malloc
  }
  std::lock_guard<std::mutex> lk(g_hb_mu);
  g_hb_live.push_back({p, bytes});
  return p;
}
size_t host_big_capacity(const void *p) {
  std::lock_guard<std::mutex> lk(g_hb_mu);
  for (auto &b : g_hb_live)
    if (b.p == p) return b.bytes;
  return 0;
}
// true when p was one of ours (it goes back to the cache; no CUDA call: safe from a finalizer thread)
bool host_big_release(void *p) {
  if (p == nullptr) return false;
  std::lock_guard<std::mutex> lk(g_hb_mu);
  for (size_t i = 0; i < g_hb_live.size(); i++)
    if (g_hb_live[i].p == p) {
      g_hb_free.push_back(g_hb_live[i]);
      g_hb_live.erase(g_hb_live.begin() + i);
      return true;
    }
  return false;
}
static void host_big_trim() {
  std::vector<BigBlock> blocks;
  {
    std::lock_guard<std::mutex> lk(g_hb_mu);
    blocks.swap(g_hb_free);
  }
  for (auto &b : blocks) cudaFreeHost(b.p);
}

void *dmalloc_bytes(size_t bytes) {
  if (bytes >= BIG) {
    // best fit among the blocks this stream may take without waiting: its own (stream order), or another stream's
    // whose last use is already over (a block freed by the second stream must not make the main stream wait for
    // everything the second stream still has queued)
    int best = -1;
    for (int i = 0; i < (int)g_big_free.size(); i++) {
      const BigBlock &c = g_big_free[i];
      if (c.bytes < bytes || c.bytes > bytes + bytes / 4) continue;
      if (c.last != nullptr && c.last != stream() && c.ev != nullptr && cudaEventQuery(c.ev) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      if (best < 0 || c.bytes < g_big_free[best].bytes) best = i;
    }
    if (best >= 0) {
      BigBlock b = g_big_free[best];
      g_big_free.erase(g_big_free.begin() + best);
      g_big_live.push_back(b);
      return b.p;
    }
  }
  void *p = nullptr;
  cudaError_t e = cudaMallocAsync(&p, bytes, stream());
  if (e != cudaSuccess) {  // give the cached blocks and the idle part of the pool back to the driver and retry
    g_alloc_retries++;
    cudaGetLastError();
    big_trim();
    cudaStreamSynchronize(g_stream);
    if (g_aux) cudaStreamSynchronize(g_aux);
    cudaMemPool_t pool;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    e = cudaMallocAsync(&p, bytes, stream());
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    uint64_t reserved = 0, used = 0;
    cudaMemPool_t pool;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
    }
    size_t live = 0;
    for (auto &b : g_big_live) live += b.bytes;
    throw Error("spasm_b200: device allocation of " + std::to_string(bytes >> 20) + " MiB failed (" + std::to_string(f >> 20) + " MiB free of " +
                std::to_string(t >> 20) + "; pool reserved " + std::to_string(reserved >> 20) + " / in use " + std::to_string(used >> 20) +
                " MiB; large blocks in use " + std::to_string(live >> 20) + " MiB in " + std::to_string(g_big_live.size()) + "): " +
                cudaGetErrorString(e));
  }
  if (bytes >= BIG) {
    BigBlock b;
    b.p = p, b.bytes = bytes;
    g_big_live.push_back(b);
  }
  return p;
}
void dfree(void *p) {
  for (size_t i = 0; i < g_big_live.size(); i++)
    if (g_big_live[i].p == p) {
      // reuse on the same stream is stream-ordered; a different stream waits for the event recorded here
      BigBlock b = g_big_live[i];
      b.last = stream();
      if (!b.ev && cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess) b.ev = nullptr;
      if (b.ev) cudaEventRecord(b.ev, b.last);
      g_big_free.push_back(b);
      g_big_live.erase(g_big_live.begin() + i);
      return;
    }
  cudaFreeAsync(p, stream());
}
// {allocation retries, bytes parked in the block cache, pool reserved, pool in use} — diagnostics (SPASM_B200_PROFILE)
void alloc_diagnostics(long long out[4]) {
  out[0] = g_alloc_retries;
  size_t c = 0;
  for (auto &b : g_big_free) c += b.bytes;
  out[1] = (long long)c;
  uint64_t reserved = 0, used = 0;
  cudaMemPool_t pool;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
  }
  out[2] = (long long)reserved, out[3] = (long long)used;
}
size_t dev_free_bytes() {  // what a new allocation could get: free memory + our own cached blocks
  size_t f = 0, t = 0;
  cudaMemGetInfo(&f, &t);
  for (auto &b : g_big_free) f += b.bytes;
  return f;
}

// ------------------------------------------------------------------ scans
void exclusive_scan_i64(const long long *in, long long *out, size_t n) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, stream());
  DBuf<char> t(tmp);
  cub::DeviceScan::ExclusiveSum(t.p, tmp, in, out, n, stream());
}
struct I32ToI64 {
  __device__ long long operator()(int v) const { return (long long)v; }
};
void exclusive_scan_i32_to_i64(const int *in, long long *out, size_t n_plus_one) {
  cub::TransformInputIterator<long long, I32ToI64, const int *> it(in, I32ToI64());
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, out, n_plus_one, stream());
  DBuf<char> t(tmp);
  cub::DeviceScan::ExclusiveSum(t.p, tmp, it, out, n_plus_one, stream());
}
long long reduce_sum_i32(const int *in, size_t n) {
  cub::TransformInputIterator<long long, I32ToI64, const int *> it(in, I32ToI64());
  DBuf<long long> out(1);
  size_t tmp = 0;
  cub::DeviceReduce::Sum(nullptr, tmp, it, out.p, n, stream());
  DBuf<char> t(tmp);
  cub::DeviceReduce::Sum(t.p, tmp, it, out.p, n, stream());
  return fetch(out.p);
}

// ------------------------------------------------------------------ large device -> host copies
// malloc'd destinations are pageable and mostly untouched: a plain cudaMemcpy runs at ~2 GB/s
// (page faults on one thread).  Bounce through two pinned buffers and let every host thread fault
// and fill its slice of the destination.
static int g_copy_threads = std::max(1, std::min(16, omp_get_num_procs()));
void download_large(void *dst, const void *src_dev, size_t bytes) {
  const size_t CH = (size_t)128 << 20;
  if (bytes < CH / 2) {
    CK(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, stream()));
    sync();
    return;
  }
  static void *pin[2] = {nullptr, nullptr};
  static cudaEvent_t ev[2];
  if (!pin[0]) {
    for (int i = 0; i < 2; i++) {
      CK(cudaHostAlloc(&pin[i], CH, cudaHostAllocDefault));
      CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
  }
  const size_t nch = (bytes + CH - 1) / CH;
  auto issue = [&](size_t c) {
    size_t off = c * CH, len = std::min(CH, bytes - off);
    CK(cudaMemcpyAsync(pin[c & 1], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, stream()));
    CK(cudaEventRecord(ev[c & 1], stream()));
  };
  issue(0);
  for (size_t c = 0; c < nch; c++) {
    if (c + 1 < nch) issue(c + 1);
    CK(cudaEventSynchronize(ev[c & 1]));
    const size_t off = c * CH, len = std::min(CH, bytes - off);
    char *d = (char *)dst + off;
    const char *srcp = (const char *)pin[c & 1];
    // explicit thread count: launchers such as torchrun export OMP_NUM_THREADS=1
#pragma omp parallel for schedule(static) num_threads(g_copy_threads)
    for (long long b = 0; b < (long long)((len + (1 << 20) - 1) >> 20); b++) {
      size_t o = (size_t)b << 20, l = std::min((size_t)1 << 20, len - o);
      memcpy(d + o, srcp + o, l);
    }
  }
}

// ------------------------------------------------------------------ CSR transfer
__global__ void k_to_u(const int *__restrict__ in, uint32_t *__restrict__ out, long long n, Fp F) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_u(in[i], F);
}
__global__ void k_to_bal(const uint32_t *in, int *out, long long n, Fp F) {  // in == out allowed
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_bal(in[i], F);
}

void convert_to_balanced(const uint32_t *in, int *out, long long n, const Fp &F) {
  if (n) k_to_bal<<<cdiv(n, 256), 256, 0, stream()>>>(in, out, n, F);
  CK(cudaGetLastError());
}
void convert_to_residues(const int *in, uint32_t *out, long long n, const Fp &F) {
  if (n) k_to_u<<<cdiv(n, 256), 256, 0, stream()>>>(in, out, n, F);
  CK(cudaGetLastError());
}

void upload_csr(const spasm_csr *A, DCsr &D, const Fp &F) {
  D.n = A->n, D.m = A->m, D.nnz = A->p[A->n];
  D.p.alloc(A->n + 1);
  D.j.alloc(D.nnz);
  D.x.alloc(D.nnz);
  D.p.upload((const long long *)A->p, A->n + 1);
  if (D.nnz) {
    D.j.upload(A->j, D.nnz);
    DBuf<int> tmp(D.nnz);
    tmp.upload(A->x, D.nnz);
    k_to_u<<<cdiv(D.nnz, 256), 256, 0, stream()>>>(tmp.p, D.x.p, D.nnz, F);
    CK(cudaGetLastError());
  }
}

spasm_csr *download_csr(const DCsr &D, int64_t prime, const Fp &F) {
  spasm_csr *A = spasm_csr_alloc(D.n, D.m, D.nnz, prime, true);
  D.p.download((long long *)A->p, D.n + 1);
  if (D.nnz) {
    D.j.download(A->j, D.nnz);
    DBuf<int> tmp(D.nnz);
    k_to_bal<<<cdiv(D.nnz, 256), 256, 0, stream()>>>(D.x.p, tmp.p, D.nnz, F);
    CK(cudaGetLastError());
    tmp.download(A->x, D.nnz);
  }
  sync();
  return A;
}

// ------------------------------------------------------------------ transpose
// T = A^T with every row of T sorted by original row index (the order a stable counting sort produces — same as
// the oracle's, src/SpaSM.jl:589).  Hand-written counting sort: (1) histogram of the columns, (2) exclusive scan =
// row pointers of T, (3) scatter with one atomic cursor per column (order inside a column is arbitrary),
// (4) every row of T sorted by original row index in shared memory (rows are short: the column weights of A).
// Traffic: read j,x once, write (row, x) once = 16 B per non-zero + 8 (n + m + 2), plus the in-place sort of (4)
// which touches each entry once more; no multi-pass radix sort over the whole matrix.
__global__ void k_count_cols(const int *__restrict__ j, long long nnz, int *__restrict__ cnt) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < nnz) atomicAdd(&cnt[j[i]], 1);
}
__global__ void k_scatter_t(const long long *__restrict__ Ap, const int *__restrict__ Aj, const uint32_t *__restrict__ Ax, int n,
                            const long long *__restrict__ Tp, int *__restrict__ cursor, int *__restrict__ Tj, uint32_t *__restrict__ Tx) {
  // one warp per row of A
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  for (long long e = Ap[w] + lane; e < Ap[w + 1]; e += 32) {
    const int c = Aj[e];
    const long long d = Tp[c] + atomicAdd(&cursor[c], 1);
    Tj[d] = w;
    Tx[d] = Ax[e];
  }
}
// rows of T by increasing original row index.  Short rows (<= 32 entries): one warp per row, rank by counting;
// longer rows: one CTA per row, bitonic network in shared memory (or in place beyond its capacity).
static constexpr int TSORT_CAP = 2048;
__device__ void t_bitonic(int *kk, uint32_t *vv, int n) {
  int N = 1;
  while (N < n) N <<= 1;
  auto cmpx = [&](int lo, int hi) {
    if (hi < n) {
      const int ka = kk[lo], kb = kk[hi];
      if (ka > kb) {
        kk[lo] = kb, kk[hi] = ka;
        const uint32_t tv = vv[lo];
        vv[lo] = vv[hi], vv[hi] = tv;
      }
    }
  };
  for (int size = 2; size <= N; size <<= 1) {
    const int half = size >> 1;
    for (int i = threadIdx.x; i < (N >> 1); i += blockDim.x) {
      const int blk = i / half, o = i - blk * half;
      cmpx(blk * size + o, blk * size + size - 1 - o);
    }
    __syncthreads();
    for (int stride = size >> 2; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (N >> 1); i += blockDim.x) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        cmpx(lo, lo | stride);
      }
      __syncthreads();
    }
  }
}
__global__ void __launch_bounds__(256) k_sort_t_rows_short(const long long *__restrict__ Tp, int m, int *__restrict__ Tj, uint32_t *__restrict__ Tx) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= m) return;
  const long long a = Tp[c];
  const int len = (int)(Tp[c + 1] - a);
  if (len <= 1 || len > 32) return;
  const int key = lane < len ? Tj[a + lane] : 0x7fffffff;
  const uint32_t val = lane < len ? Tx[a + lane] : 0u;
  int rank = 0;  // keys are distinct unless a row of A stores a column twice: ties are broken by position
  for (int u = 0; u < len; u++) {
    const int ku = __shfl_sync(0xffffffffu, key, u);
    rank += (ku < key) || (ku == key && u < lane);
  }
  __syncwarp();
  if (lane < len) Tj[a + rank] = key, Tx[a + rank] = val;
}
__global__ void __launch_bounds__(256) k_sort_t_rows_long(const long long *__restrict__ Tp, const int *__restrict__ list, int nlist,
                                                           int *__restrict__ Tj, uint32_t *__restrict__ Tx) {
  __shared__ int sk[TSORT_CAP];
  __shared__ uint32_t sv[TSORT_CAP];
  for (int t = blockIdx.x; t < nlist; t += gridDim.x) {
    const int c = list[t];
    const long long a = Tp[c];
    const int len = (int)(Tp[c + 1] - a);
    if (len <= TSORT_CAP) {
      for (int i = threadIdx.x; i < len; i += blockDim.x) sk[i] = Tj[a + i], sv[i] = Tx[a + i];
      __syncthreads();
      t_bitonic(sk, sv, len);
      for (int i = threadIdx.x; i < len; i += blockDim.x) Tj[a + i] = sk[i], Tx[a + i] = sv[i];
      __syncthreads();
    } else {
      __syncthreads();
      t_bitonic(Tj + a, Tx + a, len);
    }
  }
}
__global__ void k_flag_long_rows(const long long *__restrict__ Tp, int m, int *__restrict__ flag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < m) flag[c] = (Tp[c + 1] - Tp[c]) > 32;
  if (c == m) flag[c] = 0;
}
__global__ void k_compact_long(const int *__restrict__ flag, const long long *__restrict__ pos, int m, int *__restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < m && flag[c]) out[pos[c]] = c;
}

void transpose_csr(const DCsr &A, DCsr &T) {
  T.n = A.m, T.m = A.n, T.nnz = A.nnz;
  T.p.alloc(A.m + 1);
  T.j.alloc(A.nnz);
  T.x.alloc(A.nnz);
  DBuf<int> cnt(A.m + 1);
  cnt.zero();
  if (A.nnz) k_count_cols<<<cdiv(A.nnz, 256), 256, 0, stream()>>>(A.j.p, A.nnz, cnt.p);
  exclusive_scan_i32_to_i64(cnt.p, T.p.p, A.m + 1);
  if (A.nnz == 0) return;
  cnt.zero();  // now the scatter cursors
  k_scatter_t<<<cdiv((long long)A.n * 32, 256), 256, 0, stream()>>>(A.p.p, A.j.p, A.x.p, A.n, T.p.p, cnt.p, T.j.p, T.x.p);
  k_sort_t_rows_short<<<cdiv((long long)A.m * 32, 256), 256, 0, stream()>>>(T.p.p, A.m, T.j.p, T.x.p);
  DBuf<int> flag(A.m + 1), list(std::max(A.m, 1));
  DBuf<long long> pos(A.m + 1);
  k_flag_long_rows<<<cdiv(A.m + 1, 256), 256, 0, stream()>>>(T.p.p, A.m, flag.p);
  exclusive_scan_i32_to_i64(flag.p, pos.p, A.m + 1);
  const int nlong = (int)fetch(pos.p + A.m);
  if (nlong > 0) {
    k_compact_long<<<cdiv(A.m, 256), 256, 0, stream()>>>(flag.p, pos.p, A.m, list.p);
    k_sort_t_rows_long<<<std::min(nlong, g_sms * 8), 256, 0, stream()>>>(T.p.p, list.p, nlong, T.j.p, T.x.p);
  }
  CK(cudaGetLastError());
  g_launches += 6;
}

}  // namespace sb

// give the cached large device blocks back to the driver
extern "C" void spasm_b200_trim(void) {
  sb::release_cached(0);
  sb::host_big_trim();
}
// keep != 0: device blocks stay cached between calls (same-shaped calls in a loop allocate nothing);
// keep == 0 (default): everything is returned to the driver when an entry point returns
extern "C" void spasm_b200_set_cache(int keep) {
  sb::g_keep_cache = keep != 0;
  if (!keep) sb::release_cached(0), sb::host_big_trim();
}
// bytes of device memory this process still holds in the library's caches (0 after a trim)
extern "C" long long spasm_b200_cached_bytes(void) {
  long long tot = 0;
  for (auto &b : sb::g_big_free) tot += (long long)b.bytes;
  if (sb::stream()) {
    cudaMemPool_t pool;
    int dev = 0;
    cudaGetDevice(&dev);
    uint64_t reserved = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess)
      tot += (long long)reserved;
  }
  return tot;
}

// src/SpaSM.jl:589 — ONE argument; values always kept (test/runtests.jl:12-15)
extern "C" struct spasm_csr *spasm_transpose(const struct spasm_csr *A) {
  try {
    sb::ApiCall api_scope_;
    sb::Fp F = sb::make_field(A->field->p);
    sb::DCsr dA, dT;
    if (A->x == nullptr) throw sb::Error("spasm_transpose: pattern-only matrices are not supported");
    sb::upload_csr(A, dA, F);
    sb::transpose_csr(dA, dT);
    return sb::download_csr(dT, A->field->p, F);
  } catch (const std::exception &e) {
    sb::errf("[spasm_b200] spasm_transpose failed: %s\n", e.what());
    return nullptr;
  }
}
