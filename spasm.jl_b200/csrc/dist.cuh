// dist.cuh — one process per GPU; NCCL (loaded with dlopen, never linked) for the one real exchange
// step of the path: the broadcast of each factored dense panel to the ranks that hold later rows.
#pragma once
#include "common.cuh"

namespace sb {

struct Dist {
  int rank = 0, nranks = 1;
  void *comm = nullptr;  // ncclComm_t
  // false: rank 0 materialises every row of U (the other ranks return a factor flagged `partial` that only carries
  // r and qinv).  true: the owner of each dense panel materialises (and downloads) its rows — the factor is
  // distributed over the ranks, every rank's struct is flagged `partial`, and the download scales with 1/N.
  bool shard_factor = false;
  // kernel / rref / gesv split their rows over the ranks (every rank must then make the call, with the same complete
  // factor); the Schur complements inside spasm_echelonize are always split (that call is collective already)
  bool shard_rows = false;
};
Dist &dist();
void dist_broadcast(void *dev_buf, size_t bytes, int root);  // on the library stream
void dist_allreduce_sum_u64(unsigned long long *dev_buf, size_t count);  // in place, on the library stream
// recv = the ranks' `bytes_per_rank`-byte pieces in rank order (send may alias recv + rank * bytes_per_rank)
void dist_allgather(const void *send, void *recv, size_t bytes_per_rank);

// block-cyclic ownership of the dense panels (pure host logic, also exported for the CPU tests)
inline int panel_owner(long long b, int nranks) { return (int)(b % nranks); }
// contiguous share [lo, hi) of `nrows` independent rows that `rank` solves when the row engine is split over the ranks
inline void row_share(long long nrows, int nranks, int rank, long long *lo, long long *hi) {
  const long long per = (nrows + nranks - 1) / nranks;
  *lo = std::min(per * rank, nrows);
  *hi = std::min(*lo + per, nrows);
}
// positions (into the remaining-row list) owned by `rank`: panels b = rank, rank+nranks, ...
std::vector<int> local_positions(long long n_rem, int block, int nranks, int rank);

}  // namespace sb
