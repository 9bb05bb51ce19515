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
__global__ void k_iota_int(int *a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
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

// ---- pivot part without any level schedule.  Right-hand sides are independent, and U row i only holds pivot
// columns of LATER rows, so the multiplier of pivot i depends on rows i' < i only: a thread that walks the pivots in
// row order i = 0, 1, ... for ITS right-hand side reads values it wrote itself — no barrier of any kind.  One warp
// owns 32 consecutive right-hand sides (coalesced 128-byte accesses of Vp[i][kk..kk+31]) and streams the
// "program" (for each pivot i: the pairs (i', -U[i'][pc_i])), prefetched 32 entries per coalesced load.
// This replaces the level-scheduled loop (thousands of launches on deep pivot DAGs).
__global__ void k_prog_count(const long long *__restrict__ Tp, const int *__restrict__ pivcol, int r, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < r) {
    const int c = pivcol[i];
    cnt[i] = (int)(Tp[c + 1] - Tp[c]) - 1;  // every entry of column pc_i except the unit pivot of row i itself
  }
  if (i == r) cnt[i] = 0;
}
__global__ void k_prog_fill(const long long *__restrict__ Tp, const int *__restrict__ Tj, const uint32_t *__restrict__ Tx,
                            const int *__restrict__ pivcol, int r, const long long *__restrict__ Pp, int2 *__restrict__ prog, Fp F) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= r) return;
  const int c = pivcol[i];
  const long long a = Tp[c], b = Tp[c + 1];
  long long w = Pp[i];
  for (long long e0 = a; e0 < b; e0 += 32) {
    const long long e = e0 + lane;
    const bool ok = e < b && Tj[e] != i;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (ok) prog[w + __popc(bal & ((1u << lane) - 1u))] = make_int2(Tj[e], (int)negmod(Tx[e], F));
    w += __popc(bal);
  }
}
template <bool SMALL>
__global__ void __launch_bounds__(128) k_sptrsm_seq(const long long *__restrict__ Pp, const int2 *__restrict__ prog, int r,
                                                    uint32_t *__restrict__ Vp, long long ldv, int kc, Fp F) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int kk = wid * 32 + lane;
  if (wid * 32 >= kc) return;
  const bool live = kk < kc;
  const int kq = live ? kk : 0;
  for (int i0 = 0; i0 < r; i0 += 32) {
    const int nb = min(32, r - i0);
    const long long myP0 = (lane < nb) ? Pp[i0 + lane] : 0, myP1 = (lane < nb) ? Pp[i0 + lane + 1] : 0;
    // first chunk of the first pivot of this group
    long long a = __shfl_sync(FULL, myP0, 0), b = __shfl_sync(FULL, myP1, 0);
    int2 ent = (a + lane < b) ? __ldg(&prog[a + lane]) : make_int2(0, 0);
    for (int t = 0; t < nb; t++) {
      const int i = i0 + t;
      // prefetch the first chunk of the next pivot while this one is computed
      long long na = 0, nbb = 0;
      int2 nent = make_int2(0, 0);
      if (t + 1 < nb) {
        na = __shfl_sync(FULL, myP0, t + 1), nbb = __shfl_sync(FULL, myP1, t + 1);
        if (na + lane < nbb) nent = __ldg(&prog[na + lane]);
      }
      if (b > a) {
        uint32_t *out = Vp + (long long)i * ldv + kq;
        unsigned long long acc = *out;
        uint32_t accm = (uint32_t)acc;
        for (long long e0 = a; e0 < b; e0 += 32) {
          if (e0 > a) ent = (e0 + lane < b) ? __ldg(&prog[e0 + lane]) : make_int2(0, 0);
          const int cnt = (int)min((long long)32, b - e0);
          // 8 independent loads in flight per step (an entry beyond the list has coefficient 0 and reads row 0)
          for (int u = 0; u < cnt; u += 8) {
            uint32_t y[8], cf[8];
#pragma unroll
            for (int v = 0; v < 8; v++) {
              const int i2 = __shfl_sync(FULL, ent.x, (u + v) & 31);
              cf[v] = (uint32_t)__shfl_sync(FULL, ent.y, (u + v) & 31);
              y[v] = Vp[(long long)i2 * ldv + kq];
            }
#pragma unroll
            for (int v = 0; v < 8; v++) {
              if (SMALL)
                acc += (unsigned long long)(cf[v] * y[v]);
              else
                accm = addmod(accm, mulmod<false>(cf[v], y[v], F), F);
            }
          }
        }
        if (live) *out = SMALL ? red64(acc, F) : accm;
      }
      a = na, b = nbb, ent = nent;
    }
  }
}

// algorithmic work of the eliminations this engine stands for (the oracle's count): every pivot i with a
// non-zero multiplier on right-hand side k costs one use of U row i: 8 nnz(U_i) + 8 bytes, nnz(U_i) multiply-adds
__global__ void __launch_bounds__(256) k_dense_work(const uint32_t *__restrict__ Vp, long long ldv, int r, int kc, const long long *__restrict__ Up,
                                                     unsigned long long *__restrict__ work) {
  __shared__ int red[8];
  for (int i = blockIdx.x; i < r; i += gridDim.x) {
    int c = 0;
    for (int kk = threadIdx.x; kk < kc; kk += blockDim.x) c += Vp[(long long)i * ldv + kk] != 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; w++) tot += red[w];
      const unsigned long long len = (unsigned long long)(Up[i + 1] - Up[i]);
      if (tot) atomicAdd(&work[0], (unsigned long long)tot * (8ULL * len + 8ULL)), atomicAdd(&work[1], (unsigned long long)tot * len);
    }
    __syncthreads();
  }
}

void build_dense_schur(const DCsr &A, const int *rows, int nrows, const DCsr &U, const int *Uqinv, const Fp &F, DenseSchur &D) {
  build_dense_schur_raw(A.p.p, A.j.p, A.x.p, A.m, rows, nrows, U, Uqinv, F, D, false);
}

void build_dense_schur_raw(const long long *Ap, const int *Aj, const uint32_t *Ax, int m_, const int *rows, int nrows, const DCsr &U,
                           const int *Uqinv, const Fp &F, DenseSchur &D, bool keep_pivot_part, unsigned long long *work) {
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
  DBuf<int> pivcol(std::max(r, 1)), pcnt(std::max(r, 1) + 1);
  DBuf<long long> Pp(std::max(r, 1) + 1);
  DBuf<int2> prog;
  if (r > 0) {
    // the elimination program of the pivot part, in U row order (a topological order of the pivot DAG)
    k_pivcol_of_rows<<<cdiv(r, 256), 256, 0, s>>>(U.p.p, U.j.p, r, pivcol.p);
    k_prog_count<<<cdiv(r + 1, 256), 256, 0, s>>>(Ut.p.p, pivcol.p, r, pcnt.p);
    exclusive_scan_i32_to_i64(pcnt.p, Pp.p, r + 1);
    const long long plen = fetch(Pp.p + r);
    prog.alloc(std::max<long long>(plen, 1));
    k_prog_fill<<<cdiv((long long)r * 32, 256), 256, 0, s>>>(Ut.p.p, Ut.j.p, Ut.x.p, pivcol.p, r, Pp.p, prog.p, F);
    g_launches += 4;
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
    if (r > 0) {
      const int warps = cdiv(kc, 32);
      if (F.small)
        k_sptrsm_seq<true><<<cdiv(warps, 4), 128, 0, s>>>(Pp.p, prog.p, r, Vp.p, kc_max, kc, F);
      else
        k_sptrsm_seq<false><<<cdiv(warps, 4), 128, 0, s>>>(Pp.p, prog.p, r, Vp.p, kc_max, kc, F);
      g_launches += 1;
      if (work != nullptr) k_dense_work<<<std::min(r, sm_count() * 8), 256, 0, s>>>(Vp.p, kc_max, r, kc, U.p.p, work);
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
  D.levels = 0;  // no level schedule any more (k_sptrsm_seq)
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
