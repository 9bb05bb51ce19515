// dense_engine.cu — dense Schur complement in one shot (replaces spasm_schur_dense, prototype
// src/SpaSM.jl:765-766): all remaining rows are eliminated against the sparse U simultaneously.
//
// In the dense regime every row reaches most pivots, so row-at-a-time elimination (one pivot per
// step, random scatter) is the wrong shape for the GPU.  Here the unknowns are stored TRANSPOSED,
// one contiguous vector over the remaining rows k per column, and x.U = A[k] becomes a sparse
// triangular solve with n_rem right-hand sides, level-scheduled over the pivot DAG:
//     Y[i][:] = A^T[pc_i][:] - sum_{i' != i, U[i'][pc_i] != 0} U[i'][pc_i] * Y[i'][:]      (pivot rows, by level)
//     D^T[c][:] = A^T[c][:]  - sum_{i'} U[i'][c] * Y[i'][:]                                (free columns)
// Each term is a coalesced axpy of length n_rem: nnz(U) * n_rem MACs, 4 B per MAC from HBM/L2.
// Values are exact in F_p, so the result equals the row-by-row elimination of the oracle bit for bit.
#include <cub/cub.cuh>

#include "dense.cuh"

namespace sb {

__global__ void k_pivcol_of_rows(const long long *__restrict__ Up, const int *__restrict__ Uj, int r, int *__restrict__ pivcol) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < r) pivcol[i] = Uj[Up[i]];
}
static constexpr int RELAX_SWEEPS = 8;
// level[i] = 1 + max level of the rows i' != i that hold column pc_i (they must be final first)
__global__ void k_level_relax(const long long *__restrict__ Tp, const int *__restrict__ Tj, const int *__restrict__ pivcol, int r,
                              int *__restrict__ level, int *__restrict__ changed) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= r) return;
  const int c = pivcol[i];
  // several sweeps per launch: the iteration is monotone (levels only grow towards the longest-path
  // fixpoint), so re-evaluating against whatever the neighbours hold right now is safe, and a launch
  // in which nobody changed anything proves the fixpoint.  Cuts the number of launches of this
  // latency-bound loop (thousands of levels) by the sweep count.
  for (int sweep = 0; sweep < RELAX_SWEEPS; sweep++) {
    int h = 0;
    for (long long e = Tp[c] + lane; e < Tp[c + 1]; e += 32) {
      int i2 = Tj[e];
      if (i2 != i) h = max(h, ((volatile int *)level)[i2] + 1);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, o));
    if (lane == 0 && h > ((volatile int *)level)[i]) {
      ((volatile int *)level)[i] = h;
      *changed = 1;
    }
    __syncwarp();
  }
}
__global__ void k_iota_int(int *a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
__global__ void k_hist(const int *__restrict__ level, int r, int *__restrict__ hist) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < r) atomicAdd(&hist[level[i]], 1);
}
__global__ void k_scatter_rows_T(const long long *__restrict__ Ap, const int *__restrict__ Aj, const uint32_t *__restrict__ Ax,
                                 const int *__restrict__ rows, int k0, int kc, const int *__restrict__ qinv,
                                 const int *__restrict__ qpos, uint32_t *__restrict__ Vp, long long ldv, uint32_t *__restrict__ Dt,
                                 long long ldd) {
  int kk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (kk >= kc) return;
  const int row = rows[k0 + kk];
  for (long long e = Ap[row] + lane; e < Ap[row + 1]; e += 32) {
    const int c = Aj[e];
    const int i = qinv[c];
    if (i >= 0)
      Vp[(long long)i * ldv + kk] = Ax[e];
    else
      Dt[(long long)qpos[c] * ldd + k0 + kk] = Ax[e];
  }
}
// one output vector (pivot row i of this level, or free column) per blockIdx.x, k along threads
template <bool SMALL, bool PIVOTS>
__global__ void __launch_bounds__(256) k_sptrsm(const long long *__restrict__ Tp, const int *__restrict__ Tj, const uint32_t *__restrict__ Tx,
                                                 const int *__restrict__ list, const int *__restrict__ pivcol,
                                                 uint32_t *__restrict__ Vp, long long ldv, uint32_t *__restrict__ Dt, long long ldd,
                                                 int k0, int kc, Fp F) {
  const int item = list[blockIdx.x];  // PIVOTS: U row i ; else: index into q0
  const int kk = blockIdx.y * blockDim.x + threadIdx.x;
  int c, self;
  uint32_t *out;
  if (PIVOTS) {
    c = pivcol[item], self = item;
    out = Vp + (long long)item * ldv;
  } else {
    c = pivcol[item], self = -1;  // pivcol == q0 here
    out = Dt + (long long)item * ldd + k0;
  }
  const long long a = Tp[c], b = Tp[c + 1];
  if (b - a <= (PIVOTS ? 1 : 0)) return;  // nothing to subtract
  if (kk >= kc) return;
  if (SMALL) {
    unsigned long long acc = out[kk];
    for (long long e = a; e < b; e++) {
      const int i2 = Tj[e];
      if (i2 == self) continue;
      acc += (unsigned long long)(F.p - Tx[e]) * Vp[(long long)i2 * ldv + kk];
    }
    out[kk] = red64(acc, F);
  } else {
    uint32_t acc = out[kk];
    for (long long e = a; e < b; e++) {
      const int i2 = Tj[e];
      if (i2 == self) continue;
      acc = addmod(acc, mulmod<false>(negmod(Tx[e], F), Vp[(long long)i2 * ldv + kk], F), F);
    }
    out[kk] = acc;
  }
}

void build_dense_schur(const DCsr &A, const int *rows, int nrows, const DCsr &U, const int *Uqinv, const Fp &F, DenseSchur &D) {
  build_dense_schur_raw(A.p.p, A.j.p, A.x.p, A.m, rows, nrows, U, Uqinv, F, D, false);
}

void build_dense_schur_raw(const long long *Ap, const int *Aj, const uint32_t *Ax, int m_, const int *rows, int nrows, const DCsr &U,
                           const int *Uqinv, const Fp &F, DenseSchur &D, bool keep_pivot_part) {
  cudaStream_t s = stream();
  const int m = m_, r = U.n;
  D.n_rem = nrows;
  D.Sm0 = m - r;
  D.ld = ((long long)nrows + 63) / 64 * 64;
  // free columns
  DBuf<int> flag(m + 1), qpos(m);
  DBuf<long long> pos(m + 1);
  D.q0.alloc(std::max(D.Sm0, 1));
  {
    extern void make_free_columns(const int *qinv, int m, int *flag, long long *pos, int *q, int *qpos);
    make_free_columns(Uqinv, m, flag.p, pos.p, D.q0.p, qpos.p);
  }
  D.Dt.alloc((size_t)std::max(D.Sm0, 1) * D.ld);
  D.Dt.zero();
  if (nrows == 0 || D.Sm0 == 0) return;
  // transpose of U and the level schedule
  DCsr Ut;
  transpose_csr(U, Ut);
  DBuf<int> pivcol(std::max(r, 1)), level(std::max(r, 1)), order(std::max(r, 1)), order2(std::max(r, 1)), lev2(std::max(r, 1)), chg(1);
  int maxlev = 0;
  std::vector<int> hist_h(1, 0);
  if (r > 0) {
    k_pivcol_of_rows<<<cdiv(r, 256), 256, 0, s>>>(U.p.p, U.j.p, r, pivcol.p);
    level.zero();
    for (int it = 0;; it += 4) {
      chg.zero();
      for (int rep = 0; rep < 4; rep++) k_level_relax<<<cdiv((long long)r * 32, 256), 256, 0, s>>>(Ut.p.p, Ut.j.p, pivcol.p, r, level.p, chg.p);
      if (fetch(chg.p) == 0) break;
      if (it > r + 8) throw Error("dense engine: U is not triangular");
    }
    DBuf<int> mx(1);
    size_t tmp = 0;
    cub::DeviceReduce::Max(nullptr, tmp, level.p, mx.p, r, s);
    DBuf<char> t1(tmp);
    cub::DeviceReduce::Max(t1.p, tmp, level.p, mx.p, r, s);
    maxlev = fetch(mx.p);
    k_iota_int<<<cdiv(r, 256), 256, 0, s>>>(order.p, r);
    int bits = 1;
    while ((1LL << bits) <= maxlev) bits++;
    tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, level.p, lev2.p, order.p, order2.p, r, 0, bits, s);
    DBuf<char> t2(tmp);
    cub::DeviceRadixSort::SortPairs(t2.p, tmp, level.p, lev2.p, order.p, order2.p, r, 0, bits, s);
    DBuf<int> hist(maxlev + 1);
    hist.zero();
    k_hist<<<cdiv(r, 256), 256, 0, s>>>(level.p, r, hist.p);
    hist_h.resize(maxlev + 1);
    hist.download(hist_h.data(), maxlev + 1);
    sync();
  }
  // chunk the right-hand sides so that the pivot part fits comfortably
  size_t avail = dev_free_bytes();
  long long kc_max = nrows;
  if (r > 0) {
    long long fit = (long long)(avail / 2 / ((size_t)r * 4));
    fit = fit / 256 * 256;
    if (fit < 256) throw Error("dense engine: not enough device memory for the pivot part");
    kc_max = std::min<long long>(nrows, fit);
    if (keep_pivot_part && kc_max < nrows) throw Error("dense engine: chunk too large to keep the multipliers");
  }
  DBuf<uint32_t> Vp((size_t)std::max(r, 1) * kc_max);
  DBuf<int> freelist(std::max(D.Sm0, 1));
  k_iota_int<<<cdiv(D.Sm0, 256), 256, 0, s>>>(freelist.p, D.Sm0);
  for (long long k0 = 0; k0 < nrows; k0 += kc_max) {
    const int kc = (int)std::min<long long>(kc_max, nrows - k0);
    if (r > 0) CK(cudaMemsetAsync(Vp.p, 0, (size_t)r * kc_max * 4, s));
    k_scatter_rows_T<<<cdiv((long long)kc * 32, 256), 256, 0, s>>>(Ap, Aj, Ax, rows, (int)k0, kc, Uqinv, qpos.p, Vp.p, kc_max, D.Dt.p, D.ld);
    const int ktiles = cdiv(kc, 256);
    int off = r > 0 ? hist_h[0] : 0;
    for (int L = 1; L <= maxlev; L++) {
      const int cnt = hist_h[L];
      if (cnt == 0) continue;
      dim3 grid(cnt, ktiles);
      if (F.small)
        k_sptrsm<true, true><<<grid, 256, 0, s>>>(Ut.p.p, Ut.j.p, Ut.x.p, order2.p + off, pivcol.p, Vp.p, kc_max, D.Dt.p, D.ld, (int)k0, kc, F);
      else
        k_sptrsm<false, true><<<grid, 256, 0, s>>>(Ut.p.p, Ut.j.p, Ut.x.p, order2.p + off, pivcol.p, Vp.p, kc_max, D.Dt.p, D.ld, (int)k0, kc, F);
      off += cnt;
      g_launches += 1;
    }
    if (r > 0) {
      dim3 grid(D.Sm0, ktiles);
      if (F.small)
        k_sptrsm<true, false><<<grid, 256, 0, s>>>(Ut.p.p, Ut.j.p, Ut.x.p, freelist.p, D.q0.p, Vp.p, kc_max, D.Dt.p, D.ld, (int)k0, kc, F);
      else
        k_sptrsm<false, false><<<grid, 256, 0, s>>>(Ut.p.p, Ut.j.p, Ut.x.p, freelist.p, D.q0.p, Vp.p, kc_max, D.Dt.p, D.ld, (int)k0, kc, F);
    }
    CK(cudaGetLastError());
  }
  D.levels = maxlev + 1;
  if (keep_pivot_part) {
    D.ldv = kc_max;
    D.Vp = std::move(Vp);
  }
  sync();
}

// ---- dense rows back to sorted sparse rows (the heavy tier of the row-solve engine)
__global__ void k_dense_count(const uint32_t *__restrict__ M, long long ld, int nvec, int nrows, const int *__restrict__ todo, int off,
                              int *__restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  int c = 0;
  for (int v = 0; v < nvec; v++) c += (M[(long long)v * ld + i] != 0);
  cnt[todo[off + i]] = c;
}
__global__ void k_dense_write(const uint32_t *__restrict__ M, long long ld, int nvec, int nrows, const int *__restrict__ todo, int off,
                              const int *__restrict__ label, const long long *__restrict__ pos, int *__restrict__ oj,
                              uint32_t *__restrict__ ox, unsigned long long *__restrict__ offs, unsigned long long slab_tag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  long long w = pos[i];
  offs[todo[off + i]] = slab_tag | (unsigned long long)w;
  for (int v = 0; v < nvec; v++) {
    const uint32_t x = M[(long long)v * ld + i];
    if (x != 0) {
      oj[w] = label ? label[v] : v;
      ox[w] = x;
      w++;
    }
  }
}
__global__ void k_gather_cnt(const int *__restrict__ cnt, const int *__restrict__ todo, int off, int n, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cnt[todo[off + i]];
  if (i == n) out[i] = 0;
}

void dense_rows_to_sparse(const uint32_t *M, long long ld, int nvec, const int *label, int nrows, const int *todo, int off, int *cnt,
                          unsigned long long *offs, unsigned long long slab_tag, DBuf<int> &oj, DBuf<uint32_t> &ox) {
  cudaStream_t s = stream();
  k_dense_count<<<cdiv(nrows, 256), 256, 0, s>>>(M, ld, nvec, nrows, todo, off, cnt);
  DBuf<int> c2(nrows + 1);
  DBuf<long long> pos(nrows + 1);
  k_gather_cnt<<<cdiv(nrows + 1, 256), 256, 0, s>>>(cnt, todo, off, nrows, c2.p);
  exclusive_scan_i32_to_i64(c2.p, pos.p, nrows + 1);
  const long long tot = fetch(pos.p + nrows);
  oj.alloc(std::max<long long>(tot, 1));
  ox.alloc(std::max<long long>(tot, 1));
  k_dense_write<<<cdiv(nrows, 256), 256, 0, s>>>(M, ld, nvec, nrows, todo, off, label, pos.p, oj.p, ox.p, offs, slab_tag);
  CK(cudaGetLastError());
}

}  // namespace sb
