// sink.cu — streams finished parts of the factor U to host memory WHILE the dense tail is still
// running: a worker thread drains a job queue on its own CUDA stream (device -> pinned bounce
// buffers -> all host threads fill the pageable destination).  The 48 GB factor of configs[1] then
// costs ~0.3 s of exposed download instead of 2.3 s.
#include <omp.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "sink.cuh"

namespace sb {

struct Job {
  char *dst;
  const char *src;
  size_t bytes;
  cudaEvent_t ready;
  bool direct;  // dst is pinned: one device -> host copy, no bounce buffer
};

struct HostSink::Impl {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job> jobs;
  bool stop = false, failed = false;
  size_t inflight = 0;
  int device = 0;
  cudaStream_t cs = nullptr;
  void *pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2];
  static constexpr size_t CH = (size_t)128 << 20;
  int threads = 1;
  bool pin_cached = false;

  void run() {
    cudaSetDevice(device);
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || !jobs.empty(); });
        if (jobs.empty()) return;
        j = jobs.front();
        jobs.pop_front();
      }
      bool ok = cudaStreamWaitEvent(cs, j.ready, 0) == cudaSuccess;
      if (j.direct) {
        ok = ok && cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyDeviceToHost, cs) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(cs) == cudaSuccess;
        j.bytes = 0;
      }
      const size_t nch = (j.bytes + CH - 1) / CH;
      auto issue = [&](size_t c) {
        size_t off = c * CH, len = std::min(CH, j.bytes - off);
        ok = ok && cudaMemcpyAsync(pin[c & 1], j.src + off, len, cudaMemcpyDeviceToHost, cs) == cudaSuccess;
        ok = ok && cudaEventRecord(ev[c & 1], cs) == cudaSuccess;
      };
      if (nch) issue(0);
      for (size_t c = 0; c < nch && ok; c++) {
        if (c + 1 < nch) issue(c + 1);
        ok = ok && cudaEventSynchronize(ev[c & 1]) == cudaSuccess;
        const size_t off = c * CH, len = std::min(CH, j.bytes - off);
        char *d = j.dst + off;
        const char *sp = (const char *)pin[c & 1];
#pragma omp parallel for schedule(static) num_threads(threads)
        for (long long b = 0; b < (long long)((len + (1 << 20) - 1) >> 20); b++) {
          size_t o = (size_t)b << 20, l = std::min((size_t)1 << 20, len - o);
          memcpy(d + o, sp + o, l);
        }
      }
      cudaEventDestroy(j.ready);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) failed = true;
        inflight--;
      }
      cv.notify_all();
    }
  }
};

HostSink *g_sink = nullptr;

// the two pinned bounce buffers cost ~0.1 s to allocate: keep one pair for the life of the process
static std::mutex g_pin_mu;
static void *g_pin_cache[2] = {nullptr, nullptr};
static bool g_pin_busy = false;

HostSink::HostSink() : impl(new Impl) {
  CK(cudaGetDevice(&impl->device));
  CK(cudaStreamCreateWithFlags(&impl->cs, cudaStreamNonBlocking));
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (!g_pin_busy) {
      for (int i = 0; i < 2; i++)
        if (g_pin_cache[i] == nullptr) CK(cudaHostAlloc(&g_pin_cache[i], Impl::CH, cudaHostAllocDefault));
      impl->pin[0] = g_pin_cache[0], impl->pin[1] = g_pin_cache[1];
      g_pin_busy = true;
      impl->pin_cached = true;
    }
  }
  for (int i = 0; i < 2; i++) {
    if (!impl->pin_cached) CK(cudaHostAlloc(&impl->pin[i], Impl::CH, cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&impl->ev[i], cudaEventDisableTiming));
  }
  impl->threads = std::max(1, std::min(12, omp_get_num_procs() - 2));
  impl->th = std::thread([this] { impl->run(); });
}

HostSink::~HostSink() {
  {
    std::lock_guard<std::mutex> lk(impl->mu);
    impl->stop = true;
  }
  impl->cv.notify_all();
  if (impl->th.joinable()) impl->th.join();
  for (int i = 0; i < 2; i++) {
    if (!impl->pin_cached) cudaFreeHost(impl->pin[i]);
    cudaEventDestroy(impl->ev[i]);
  }
  if (impl->pin_cached) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_busy = false;
  }
  cudaStreamDestroy(impl->cs);
  if (!host_big_release(hj)) free(hj);
  if (!host_big_release(hx)) free(hx);
  delete impl;
}

void HostSink::wait_all() {
  std::unique_lock<std::mutex> lk(impl->mu);
  impl->cv.wait(lk, [&] { return impl->inflight == 0; });
  if (impl->failed) throw Error("asynchronous download of the factor failed");
}

void HostSink::ensure(long long entries) {
  if (entries <= cap) return;
  wait_all();  // nobody may be writing into the old arrays
  long long ncap = entries + entries / 8 + 1024;
  int *nj = (int *)host_big_alloc((size_t)ncap * sizeof(int)), *nx = nullptr;
  if (nj != nullptr) nx = (int *)host_big_alloc((size_t)ncap * sizeof(int));
  if (nj != nullptr && nx != nullptr) {
    // cache mode: pinned blocks (the device writes into them directly; they return to the cache in spasm_csr_free)
    if (submitted > 0) memcpy(nj, hj, (size_t)submitted * sizeof(int)), memcpy(nx, hx, (size_t)submitted * sizeof(int));
    if (!host_big_release(hj)) free(hj);
    if (!host_big_release(hx)) free(hx);
    hj = nj, hx = nx;
    pinned = true;
  } else {
    if (nj != nullptr) host_big_release(nj);
    if (pinned) {  // (cannot happen in practice: a pinned pair that has to grow while pinned memory ran out)
      int *mj = (int *)spasm_malloc(ncap * (i64)sizeof(int)), *mx = (int *)spasm_malloc(ncap * (i64)sizeof(int));
      memcpy(mj, hj, (size_t)submitted * sizeof(int)), memcpy(mx, hx, (size_t)submitted * sizeof(int));
      host_big_release(hj), host_big_release(hx);
      hj = mj, hx = mx;
      pinned = false;
    } else {
      hj = (int *)spasm_realloc(hj, ncap * (i64)sizeof(int));
      hx = (int *)spasm_realloc(hx, ncap * (i64)sizeof(int));
    }
  }
  cap = ncap;
}

// entries [first, first+count) of U.j / U.x (x already balanced in place) are final on the compute stream
void HostSink::submit(const int *dev_j, const int *dev_x, long long first, long long count) {
  if (count <= 0) return;
  if (first != submitted) throw Error("HostSink: ranges must be submitted in order");
  ensure(first + count);
  const int *srcs[2] = {dev_j + first, dev_x + first};
  int *dsts[2] = {hj + first, hx + first};
  for (int a = 0; a < 2; a++) {
    Job j;
    CK(cudaEventCreateWithFlags(&j.ready, cudaEventDisableTiming));
    CK(cudaEventRecord(j.ready, stream()));
    j.dst = (char *)dsts[a], j.src = (const char *)srcs[a], j.bytes = (size_t)count * 4;
    j.direct = pinned;
    {
      std::lock_guard<std::mutex> lk(impl->mu);
      impl->jobs.push_back(j);
      impl->inflight++;
    }
    impl->cv.notify_all();
  }
  submitted = first + count;
}

// detach the host arrays (the caller owns them afterwards)
void HostSink::release(int **pj, int **px) {
  wait_all();
  *pj = hj, *px = hx;
  hj = hx = nullptr;
  cap = 0;
}

}  // namespace sb
