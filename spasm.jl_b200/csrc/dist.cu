// dist.cu — NCCL plumbing (dlopen) + the partition helper.
#include <dlfcn.h>
#include <nccl.h>

#include "dist.cuh"

namespace sb {

static Dist g_dist;
Dist &dist() { return g_dist; }

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi &nccl() {
  static NcclApi api;
  if (api.h == nullptr) {
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // reuses torch's copy when it is already loaded
    if (api.h == nullptr) throw Error(std::string("cannot load libnccl.so.2: ") + dlerror());
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
    api.Broadcast = (decltype(api.Broadcast))dlsym(api.h, "ncclBroadcast");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
    api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.Broadcast || !api.AllReduce || !api.AllGather || !api.CommDestroy) throw Error("libnccl.so.2 lacks the expected symbols");
  }
  return api;
}
#define NCK(call)                                                                                         \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess) throw sb::Error(std::string("NCCL error ") + (nccl().GetErrorString ? nccl().GetErrorString(r_) : "?")); \
  } while (0)

void dist_broadcast(void *dev_buf, size_t bytes, int root) {
  if (g_dist.nranks <= 1 || bytes == 0) return;
  // chunk: ncclBroadcast counts are size_t but keep single calls below 1 GiB
  const size_t CH = (size_t)1 << 30;
  for (size_t off = 0; off < bytes; off += CH) {
    size_t len = std::min(CH, bytes - off);
    NCK(nccl().Broadcast((char *)dev_buf + off, (char *)dev_buf + off, len, ncclUint8, root, (ncclComm_t)g_dist.comm, stream()));
  }
}

void dist_allreduce_sum_u64(unsigned long long *dev_buf, size_t count) {
  if (g_dist.nranks <= 1 || count == 0) return;
  const size_t CH = (size_t)1 << 27;  // elements per call (1 GiB)
  for (size_t off = 0; off < count; off += CH) {
    size_t len = std::min(CH, count - off);
    NCK(nccl().AllReduce(dev_buf + off, dev_buf + off, len, ncclUint64, ncclSum, (ncclComm_t)g_dist.comm, stream()));
  }
}
void dist_allgather(const void *send, void *recv, size_t bytes_per_rank) {
  if (bytes_per_rank == 0) return;
  if (g_dist.nranks <= 1) {
    if (send != recv) CK(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, stream()));
    return;
  }
  NCK(nccl().AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)g_dist.comm, stream()));
}

std::vector<int> local_positions(long long n_rem, int block, int nranks, int rank) {
  std::vector<int> out;
  const long long nb = (n_rem + block - 1) / block;
  for (long long b = rank; b < nb; b += nranks)
    for (long long k = b * block; k < std::min<long long>((b + 1) * block, n_rem); k++) out.push_back((int)k);
  return out;
}

}  // namespace sb

using namespace sb;

extern "C" {

// rank 0 creates the id (128 bytes), the host framework ships it to the other ranks
// (torch.distributed broadcast in bench.py), every rank then calls spasm_b200_dist_init.
int spasm_b200_nccl_unique_id(unsigned char *out128) {
  try {
    ncclUniqueId id;
    NCK(nccl().GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return 0;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_b200_nccl_unique_id failed: %s\n", e.what());
    return -1;
  }
}
int spasm_b200_dist_init(int rank, int nranks, const unsigned char *id128) {
  try {
    ApiCall api_scope_;
    if (nranks <= 1) {
      dist() = Dist();
      return 0;
    }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NCK(nccl().CommInitRank(&comm, nranks, id, rank));
    const bool keep = dist().shard_factor, keep_rows = dist().shard_rows;
    dist().rank = rank, dist().nranks = nranks, dist().comm = comm, dist().shard_factor = keep, dist().shard_rows = keep_rows;
    return 0;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_b200_dist_init failed: %s\n", e.what());
    return -1;
  }
}
// 0 (default): rank 0 returns the complete factor; 1: every rank returns the rows of U it owns (see dist.cuh)
void spasm_b200_dist_shard_factor(int on) { dist().shard_factor = on != 0; }
// 1: spasm_kernel / spasm_rref / spasm_gesv become collective calls (every rank, same complete factor) whose rows
// (free columns, rows of U, right-hand sides) are split over the ranks and all-gathered; 0 (default): each call is local
void spasm_b200_dist_shard_rows(int on) { dist().shard_rows = on != 0; }
void spasm_b200_dist_finalize(void) {
  if (dist().comm) nccl().CommDestroy((ncclComm_t)dist().comm);
  dist() = Dist();
}
// pure host logic of the sharding, exported so that the CPU (gloo) tests can check it:
// writes the positions owned by `rank` into out (capacity cap), returns how many
long long spasm_b200_local_positions(long long n_rem, int block, int nranks, int rank, int *out, long long cap) {
  std::vector<int> v = local_positions(n_rem, block, nranks, rank);
  for (size_t i = 0; i < v.size() && (long long)i < cap; i++) out[i] = v[i];
  return (long long)v.size();
}
int spasm_b200_panel_owner(long long b, int nranks) { return panel_owner(b, nranks); }
void spasm_b200_row_share(long long nrows, int nranks, int rank, long long *lo, long long *hi) { row_share(nrows, nranks, rank, lo, hi); }

}  // extern "C"
