// sink.cuh — asynchronous device->host streaming of the factor (sink.cu)
#pragma once
#include "common.cuh"

namespace sb {

struct HostSink {
  struct Impl;
  Impl *impl;
  int *hj = nullptr, *hx = nullptr;  // host arrays being filled (malloc family)
  long long cap = 0;                 // their capacity in entries
  long long submitted = 0;           // entries [0, submitted) are on their way / done
  bool pinned = false;               // hj / hx are pinned blocks of the host cache (cache mode): direct device -> host copies
  HostSink();
  ~HostSink();
  HostSink(const HostSink &) = delete;
  HostSink &operator=(const HostSink &) = delete;
  void ensure(long long entries);
  void submit(const int *dev_j, const int *dev_x, long long first, long long count);
  void wait_all();
  void release(int **pj, int **px);
};
extern HostSink *g_sink;  // set by spasm_echelonize while it runs with host output; nullptr otherwise

}  // namespace sb
