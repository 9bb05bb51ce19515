// compress.cu — triplets -> CSR on the device (the role of spasm_compress, src/SpaSM.jl:479-493).
//
// Semantics of the host code (host_abi.cu / oracle): entries are bucketed by row IN THEIR ORIGINAL ORDER, duplicates of
// a (row, column) pair are summed into the first occurrence, entries whose sum is zero are dropped.  On the device:
//   1. histogram of the rows + scan -> row pointers;
//   2. scatter of the ENTRY INDICES with one atomic cursor per row (order inside a row: whatever the atomics gave),
//      then every row sorted by entry index (sort_csr_rows: the CTA-per-row bitonic sort of the row engine) -> the
//      original order, deterministically; gather (column, value) through the sorted indices;
//   3. duplicates: short rows (<= 64 entries) by one warp per row, quadratic in registers; longer rows by a few CTAs
//      with a direct-indexed scratch over the m columns (first position by atomicMin, unreduced 64-bit sum by atomicAdd);
//   4. per-row counts of what survives, scan, stable compaction (warp per row, ballot prefix).
// 8 B per entry read + 8 B written per pass; the passes are HBM-bound streaming except the per-row sort.
#include <climits>

#include "common.cuh"
#include "solve_sparse.cuh"

namespace sb {

static constexpr int SHORT_ROW = 64;
static constexpr int LONG_CTAS = 32;

__global__ void k_trip_count(const int *__restrict__ Ti, long long nz, int *__restrict__ cnt) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nz) atomicAdd(&cnt[Ti[e]], 1);
}
__global__ void k_trip_scatter(const int *__restrict__ Ti, long long nz, const long long *__restrict__ rp, int *__restrict__ cursor,
                               int *__restrict__ key) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= nz) return;
  const int i = Ti[e];
  key[rp[i] + atomicAdd(&cursor[i], 1)] = (int)e;
}
__global__ void k_trip_gather(const int *__restrict__ key, long long nz, const int *__restrict__ Tj, const uint32_t *__restrict__ Tx,
                              int *__restrict__ cj, uint32_t *__restrict__ cx) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nz) return;
  const int e = key[t];
  cj[t] = Tj[e];
  if (Tx != nullptr) cx[t] = Tx[e];
}
// rows of at most SHORT_ROW entries: one warp per row.  keep[t] = 1 for the first occurrence of a column whose sum is
// non-zero (or, without values, for every first occurrence); sx[t] = that sum.  Rows longer than that are flagged.
__global__ void k_dedupe_short(const long long *__restrict__ rp, int n, const int *__restrict__ cj, const uint32_t *__restrict__ cx,
                               unsigned char *__restrict__ keep, uint32_t *__restrict__ sx, int *__restrict__ outcnt,
                               int *__restrict__ islong, Fp F) {
  const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const long long a = rp[row];
  const int k = (int)(rp[row + 1] - a);
  if (k > SHORT_ROW) {
    if (lane == 0) islong[row] = 1;
    return;
  }
  if (lane == 0) islong[row] = 0;
  int kept = 0;
  for (int t = lane; t < k; t += 32) {
    const int c = cj[a + t];
    bool first = true;
    for (int t2 = 0; t2 < t; t2++)
      if (cj[a + t2] == c) {
        first = false;
        break;
      }
    uint32_t s = 0;
    bool kp = first;
    if (first && cx != nullptr) {
      s = cx[a + t];
      for (int t2 = t + 1; t2 < k; t2++)
        if (cj[a + t2] == c) s = addmod(s, cx[a + t2], F);
      kp = (s != 0);
    }
    keep[a + t] = kp ? 1 : 0;
    if (cx != nullptr) sx[a + t] = s;
    kept += kp ? 1 : 0;
  }
  for (int o = 16; o; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
  if (lane == 0) outcnt[row] = kept;
}
// long rows: CTA g owns scratch slice g (first[m] initialised to INT_MAX, sum[m] to 0) and takes the rows g, g + G, ...
// of the list; the scratch is restored after every row.
__global__ void __launch_bounds__(256) k_dedupe_long(const int *__restrict__ list, int nlong, const long long *__restrict__ rp,
                                                      const int *__restrict__ cj, const uint32_t *__restrict__ cx, int m,
                                                      int *__restrict__ first_all, unsigned long long *__restrict__ sum_all,
                                                      unsigned char *__restrict__ keep, uint32_t *__restrict__ sx, int *__restrict__ outcnt,
                                                      Fp F) {
  __shared__ int s_cnt;
  int *first = first_all + (size_t)blockIdx.x * m;
  unsigned long long *sum = sum_all + (size_t)blockIdx.x * m;
  for (int q = blockIdx.x; q < nlong; q += gridDim.x) {
    const int row = list[q];
    const long long a = rp[row];
    const int k = (int)(rp[row + 1] - a);
    if (threadIdx.x == 0) s_cnt = 0;
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
      const int c = cj[a + t];
      atomicMin(&first[c], t);
      if (cx != nullptr) atomicAdd(&sum[c], (unsigned long long)cx[a + t]);
    }
    __syncthreads();
    int kept = 0;
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
      const int c = cj[a + t];
      bool kp = (first[c] == t);
      uint32_t s = 0;
      if (kp && cx != nullptr) {
        s = (uint32_t)(sum[c] % (unsigned long long)F.p);
        kp = (s != 0);
      }
      keep[a + t] = kp ? 1 : 0;
      if (cx != nullptr) sx[a + t] = s;
      kept += kp ? 1 : 0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
      const int c = cj[a + t];
      first[c] = INT_MAX;
      if (cx != nullptr) sum[c] = 0ULL;
    }
    if (kept) atomicAdd(&s_cnt, kept);
    __syncthreads();
    if (threadIdx.x == 0) outcnt[row] = s_cnt;
    __syncthreads();
  }
}
__global__ void k_fill_int_max(int *a, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) a[i] = INT_MAX;
}
// stable compaction of the surviving entries of every row (warp per row)
__global__ void k_trip_compact(const long long *__restrict__ rp, int n, const int *__restrict__ cj, const uint32_t *__restrict__ sx,
                               const unsigned char *__restrict__ keep, const long long *__restrict__ Cp, int *__restrict__ Cj,
                               uint32_t *__restrict__ Cx) {
  const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const long long a = rp[row], b = rp[row + 1];
  long long out = Cp[row];
  for (long long t0 = a; t0 < b; t0 += 32) {
    const long long t = t0 + lane;
    const bool kp = t < b && keep[t] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, kp);
    if (kp) {
      const long long d = out + __popc(bal & ((1u << lane) - 1u));
      Cj[d] = cj[t];
      if (Cx != nullptr) Cx[d] = sx[t];
    }
    out += __popc(bal);
  }
}
__global__ void k_trip_check(const int *__restrict__ Ti, const int *__restrict__ Tj, long long nz, int n, int m, int *__restrict__ bad) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nz && ((unsigned)Ti[e] >= (unsigned)n || (unsigned)Tj[e] >= (unsigned)m)) atomicAdd(bad, 1);
}
__global__ void k_compact_flagged(const int *__restrict__ flag, const long long *__restrict__ pos, int n, int *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[pos[i]] = i;
}

}  // namespace sb

using namespace sb;

// device version of spasm_compress: same result, bit for bit, as the host code (first occurrence keeps its place,
// duplicates summed, zero sums dropped).  Returns NULL (message on stderr) without a GPU or on malformed input.
extern "C" struct spasm_csr *spasm_b200_compress(const struct spasm_triplet *T) {
  try {
    ApiCall api_scope_;
    cudaStream_t s = stream();
    const int n = T->n, m = T->m;
    const long long nz = T->nz;
    const bool with_values = T->x != nullptr;
    const int64_t prime = T->field->p;
    if (nz >= (1LL << 31)) throw Error("spasm_b200_compress: more than 2^31 - 1 triplets");
    Fp F = make_field(prime);
    DBuf<int> Ti(std::max<long long>(nz, 1)), Tj(std::max<long long>(nz, 1)), cnt(n + 1), cursor(std::max(n, 1));
    DBuf<uint32_t> Tx;
    DBuf<long long> rp(n + 1);
    cnt.zero();
    cursor.zero();
    if (nz) {
      Ti.upload(T->i, nz), Tj.upload(T->j, nz);
      if (with_values) {
        DBuf<int> raw(nz);
        raw.upload(T->x, nz);
        Tx.alloc(nz);
        convert_to_residues(raw.p, Tx.p, nz, F);
        sync();  // raw goes away
      }
      DBuf<int> bad(1);
      bad.zero();
      k_trip_check<<<cdiv(nz, 256), 256, 0, s>>>(Ti.p, Tj.p, nz, n, m, bad.p);
      if (fetch(bad.p) != 0) throw Error("spasm_b200_compress: a triplet lies outside the matrix");
      k_trip_count<<<cdiv(nz, 256), 256, 0, s>>>(Ti.p, nz, cnt.p);
    }
    exclusive_scan_i32_to_i64(cnt.p, rp.p, n + 1);
    DBuf<int> key(std::max<long long>(nz, 1)), cj(std::max<long long>(nz, 1)), outcnt(n + 1), islong(n + 1);
    DBuf<uint32_t> dummy(std::max<long long>(nz, 1)), cx, sx;
    DBuf<unsigned char> keep(std::max<long long>(nz, 1));
    outcnt.zero();
    islong.zero();
    if (with_values) cx.alloc(std::max<long long>(nz, 1)), sx.alloc(std::max<long long>(nz, 1));
    if (nz) {
      k_trip_scatter<<<cdiv(nz, 256), 256, 0, s>>>(Ti.p, nz, rp.p, cursor.p, key.p);
      dummy.zero();
      sort_csr_rows(rp.p, n, key.p, dummy.p);  // every row back in the order of the triplet list
      k_trip_gather<<<cdiv(nz, 256), 256, 0, s>>>(key.p, nz, Tj.p, with_values ? Tx.p : nullptr, cj.p, with_values ? cx.p : nullptr);
      k_dedupe_short<<<cdiv((long long)n * 32, 256), 256, 0, s>>>(rp.p, n, cj.p, with_values ? cx.p : nullptr, keep.p,
                                                                   with_values ? sx.p : nullptr, outcnt.p, islong.p, F);
      // rows with more than SHORT_ROW entries
      DBuf<long long> lpos(n + 1);
      exclusive_scan_i32_to_i64(islong.p, lpos.p, n + 1);
      const int nlong = (int)fetch(lpos.p + n);
      if (nlong > 0) {
        DBuf<int> list(nlong);
        k_compact_flagged<<<cdiv(n, 256), 256, 0, s>>>(islong.p, lpos.p, n, list.p);
        const int G = std::min(LONG_CTAS, nlong);
        DBuf<int> first((size_t)G * m);
        DBuf<unsigned long long> sum(with_values ? (size_t)G * m : 1);
        k_fill_int_max<<<cdiv((long long)G * m, 256), 256, 0, s>>>(first.p, (long long)G * m);
        sum.zero();
        k_dedupe_long<<<G, 256, 0, s>>>(list.p, nlong, rp.p, cj.p, with_values ? cx.p : nullptr, m, first.p, sum.p, keep.p,
                                        with_values ? sx.p : nullptr, outcnt.p, F);
        sync();  // the scratch goes away
      }
      CK(cudaGetLastError());
    }
    DBuf<long long> Cp(n + 1);
    exclusive_scan_i32_to_i64(outcnt.p, Cp.p, n + 1);
    const long long nnz = fetch(Cp.p + n);
    spasm_csr *C = spasm_csr_alloc(n, m, std::max<long long>(nnz, 1), prime, with_values);
    if (nnz) {
      DBuf<int> Cj(nnz);
      DBuf<uint32_t> Cx;
      if (with_values) Cx.alloc(nnz);
      k_trip_compact<<<cdiv((long long)n * 32, 256), 256, 0, s>>>(rp.p, n, cj.p, with_values ? sx.p : nullptr, keep.p, Cp.p, Cj.p,
                                                                   with_values ? Cx.p : nullptr);
      CK(cudaGetLastError());
      download_large(C->j, Cj.p, (size_t)nnz * sizeof(int));
      if (with_values) {
        convert_to_balanced(Cx.p, (int *)Cx.p, nnz, F);
        download_large(C->x, Cx.p, (size_t)nnz * sizeof(int));
      }
    }
    CK(cudaMemcpyAsync(C->p, Cp.p, (size_t)(n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, s));
    sync();
    return C;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_b200_compress failed: %s\n", e.what());
    return nullptr;
  }
}
