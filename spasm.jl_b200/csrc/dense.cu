// dense.cu — dense tail, first version: Gauss-Jordan by pivot steps on CUDA cores.
// (The blocked tcgen05 path that replaces the elimination step lives in dense_mma.cu.)
#include "dense.cuh"

namespace sb {

static constexpr int INF = 0x7fffffff;

// cur[0] = pivot column (or -1 when finished), cur[1] = pivot row, cur[2] = inverse of the pivot,
// cur[3] = rank so far
__global__ void k_lead_init(const uint32_t *__restrict__ S, int n, int m, long long ld, int *__restrict__ lead,
                            int *__restrict__ ispiv, int *__restrict__ cur) {
  __shared__ int red[32];
  int r = blockIdx.x;
  int best = INF;
  for (int k = threadIdx.x; k < m; k += blockDim.x)
    if (S[r * ld + k] != 0) {
      best = k;
      break;
    }
  for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)blockDim.x / 32; i++) best = min(best, red[i]);
    lead[r] = best;
    ispiv[r] = 0;
    if (r == 0) cur[0] = 0, cur[3] = 0;
  }
}
__global__ void k_select(uint32_t *__restrict__ S, int n, long long ld, const int *__restrict__ lead, int *__restrict__ ispiv,
                         int *__restrict__ cur, int *__restrict__ pivcol, int *__restrict__ pivrow, Fp F) {
  __shared__ unsigned long long red[32];
  if (cur[0] < 0) return;
  unsigned long long best = ~0ULL;
  for (int r = threadIdx.x; r < n; r += blockDim.x)
    if (!ispiv[r] && lead[r] != INF) {
      unsigned long long key = ((unsigned long long)(unsigned)lead[r] << 32) | (unsigned)r;
      best = key < best ? key : best;
    }
  for (int o = 16; o; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, best, o);
    best = w < best ? w : best;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)blockDim.x / 32; i++) best = red[i] < best ? red[i] : best;
    if (best == ~0ULL) {
      cur[0] = -1;
    } else {
      int c = (int)(best >> 32), r = (int)(best & 0xffffffffu);
      int s = cur[3];
      pivcol[s] = c, pivrow[s] = r;
      ispiv[r] = 1;
      cur[0] = c, cur[1] = r, cur[2] = (int)dev_inv(S[r * ld + c], F.p), cur[3] = s + 1;
    }
  }
}
template <bool SMALL>
__global__ void k_scale(uint32_t *__restrict__ S, int m, long long ld, const int *__restrict__ cur, Fp F) {
  const int c = cur[0];
  if (c < 0) return;
  const uint32_t alpha = (uint32_t)cur[2];
  if (alpha == 1) return;
  uint32_t *P = S + cur[1] * ld;
  for (int k = c + blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) P[k] = mulmod<SMALL>(alpha, P[k], F);
}
template <bool SMALL>
__global__ void k_eliminate(uint32_t *__restrict__ S, int n, int m, long long ld, int *__restrict__ lead,
                            const int *__restrict__ ispiv, const int *__restrict__ cur, Fp F) {
  __shared__ int red[32];
  const int c = cur[0];
  if (c < 0) return;
  const int r = blockIdx.x, pr = cur[1];
  if (r == pr) return;
  uint32_t *R = S + r * ld;
  const uint32_t f = R[c];
  if (f == 0) return;
  const uint32_t *P = S + pr * ld;
  const uint32_t nf = negmod(f, F);
  int best = INF;
  for (int k = c + threadIdx.x; k < m; k += blockDim.x) {
    uint32_t v = addmod(R[k], mulmod<SMALL>(nf, P[k], F), F);
    R[k] = v;
    if (v != 0 && k < best) best = k;
  }
  if (ispiv[r]) return;
  for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)blockDim.x / 32; i++) best = min(best, red[i]);
    lead[r] = best;
  }
}

int dense_rref_device(uint32_t *S, int n, int m, long long ld, const Fp &F, DBuf<int> &pivcol, DBuf<int> &pivrow) {
  cudaStream_t s = stream();
  const int maxr = std::min(n, m);
  pivcol.alloc(std::max(maxr, 1));
  pivrow.alloc(std::max(maxr, 1));
  if (n == 0 || m == 0) return 0;
  DBuf<int> lead(n), ispiv(n), cur(4);
  k_lead_init<<<n, 256, 0, s>>>(S, n, m, ld, lead.p, ispiv.p, cur.p);
  const int sblocks = std::max(1, std::min(cdiv(m, 256), sm_count() * 2));
  for (int step = 0; step < maxr; step++) {
    k_select<<<1, 1024, 0, s>>>(S, n, ld, lead.p, ispiv.p, cur.p, pivcol.p, pivrow.p, F);
    if (F.small) {
      k_scale<true><<<sblocks, 256, 0, s>>>(S, m, ld, cur.p, F);
      k_eliminate<true><<<n, 256, 0, s>>>(S, n, m, ld, lead.p, ispiv.p, cur.p, F);
    } else {
      k_scale<false><<<sblocks, 256, 0, s>>>(S, m, ld, cur.p, F);
      k_eliminate<false><<<n, 256, 0, s>>>(S, n, m, ld, lead.p, ispiv.p, cur.p, F);
    }
    if ((step & 63) == 63) {
      CK(cudaGetLastError());
      if (fetch(cur.p) < 0) break;
    }
  }
  CK(cudaGetLastError());
  int h[4];
  cur.download(h, 4);
  sync();
  return h[3];
}

// ------------------------------------------------------------------ block assembly / emission
__global__ void k_free_cols(const int *__restrict__ qinv, int m, int *__restrict__ flag) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) flag[j] = qinv[j] < 0;
  if (j == m) flag[j] = 0;
}
__global__ void k_make_q(const int *__restrict__ flag, const long long *__restrict__ pos, int m, int *__restrict__ q, int *__restrict__ qpos) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  if (flag[j]) {
    q[pos[j]] = j;
    qpos[j] = (int)pos[j];
  } else
    qpos[j] = -1;
}
__global__ void k_scatter_dense(const long long *__restrict__ Rp, const int *__restrict__ Rj, const uint32_t *__restrict__ Rx, int nrows,
                                const int *__restrict__ qpos, uint32_t *__restrict__ S, long long ld) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= nrows) return;
  for (long long e = Rp[k] + lane; e < Rp[k + 1]; e += 32) S[k * ld + qpos[Rj[e]]] = Rx[e];
}
// count / write the reduced rows as U rows: (q[pivcol], 1) first, then the other nonzeros by column
__global__ void k_count_rows(const uint32_t *__restrict__ S, long long ld, int m, const int *__restrict__ pivrow, int rr, int *__restrict__ cnt) {
  __shared__ int red[32];
  int s = blockIdx.x;
  const uint32_t *R = S + pivrow[s] * ld;
  int c = 0;
  for (int k = threadIdx.x; k < m; k += blockDim.x) c += (R[k] != 0);
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)blockDim.x / 32; i++) c += red[i];
    cnt[s] = c;  // includes the pivot entry itself
    if (s == 0) cnt[rr] = 0;
  }
}
__global__ void k_write_rows(const uint32_t *__restrict__ S, long long ld, int m, const int *__restrict__ pivrow,
                             const int *__restrict__ pivcol, const int *__restrict__ q, const long long *__restrict__ pos,
                             long long ubase, int urow0, long long *__restrict__ Up, int *__restrict__ Uj, uint32_t *__restrict__ Ux,
                             int *__restrict__ Uqinv) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int s = blockIdx.x, pc = pivcol[s];
  const uint32_t *R = S + pivrow[s] * ld;
  const long long dst = ubase + pos[s];
  if (threadIdx.x == 0) {
    Uj[dst] = q[pc];
    Ux[dst] = 1;
    Uqinv[q[pc]] = urow0 + s;
    Up[urow0 + s + 1] = ubase + pos[s + 1];
    carry = 1;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k0 = 0; k0 < m; k0 += blockDim.x) {
    int k = k0 + threadIdx.x;
    uint32_t v = (k < m && k != pc) ? R[k] : 0u;
    unsigned b = __ballot_sync(0xffffffffu, v != 0);
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    int base = carry;
    for (int i = 0; i < w; i++) base += wsum[i];
    if (v != 0) {
      long long d = dst + base + __popc(b & ((1u << lane) - 1u));
      Uj[d] = q[k];
      Ux[d] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int i = 0; i < nw; i++) t += wsum[i];
      carry += t;
    }
    __syncthreads();
  }
}

void echelonize_dense_device(const DCsr &A, const int *rows, int nrows, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size) {
  cudaStream_t s = stream();
  const int m = A.m;
  if (block_size <= 0) block_size = 1000;
  DBuf<int> flag(m + 1), q(m), qpos(m), pivcol, pivrow;
  DBuf<long long> pos(m + 1);
  DBuf<PDesc> pdesc;
  for (int processed = 0; processed < nrows;) {
    const int Sm = m - U.n;
    if (Sm == 0) break;
    const int Sn = std::min(block_size, nrows - processed);
    logf("[echelonize/dense] processing dense schur complement of dimension %d x %d; block size=%d\n", nrows - processed, Sm, block_size);
    k_free_cols<<<cdiv(m + 1, 256), 256, 0, s>>>(Uqinv.p, m, flag.p);
    exclusive_scan_i32_to_i64(flag.p, pos.p, m + 1);
    k_make_q<<<cdiv(m, 256), 256, 0, s>>>(flag.p, pos.p, m, q.p, qpos.p);
    build_pdesc_U(U, Uqinv.p, pdesc);
    SolveSystem G{U.j.p, U.x.p, pdesc.p, m};
    SolveRows B{A.p.p, A.j.p, A.x.p, rows + processed, Sn, nullptr};
    SolveEmit E;
    SolveResult R;
    solve_rows(G, B, E, F, R);
    DBuf<uint32_t> S((size_t)Sn * Sm);
    S.zero();
    k_scatter_dense<<<cdiv((long long)Sn * 32, 256), 256, 0, s>>>(R.p.p, R.j.p, R.x.p, Sn, qpos.p, S.p, Sm);
    CK(cudaGetLastError());
    const int rr = dense_rref_device(S.p, Sn, Sm, Sm, F, pivcol, pivrow);
    if (rr > 0) {
      DBuf<int> cnt(rr + 1);
      DBuf<long long> rpos(rr + 1);
      k_count_rows<<<rr, 256, 0, s>>>(S.p, Sm, Sm, pivrow.p, rr, cnt.p);
      exclusive_scan_i32_to_i64(cnt.p, rpos.p, rr + 1);
      const long long add = fetch(rpos.p + rr);
      csr_reserve(U, U.nnz + add, U.n + rr);
      k_write_rows<<<rr, 256, 0, s>>>(S.p, Sm, Sm, pivrow.p, pivcol.p, q.p, rpos.p, U.nnz, U.n, U.p.p, U.j.p, U.x.p, Uqinv.p);
      CK(cudaGetLastError());
      U.nnz += add;
      U.n += rr;
    }
    processed += Sn;
    logf("[echelonize/dense] block done: %d new pivots, rank %d\n", rr, U.n);
  }
}

}  // namespace sb

using namespace sb;

// the dense-tail entry point of the ABI (replaces spasm_ffpack_rref, src/SpaSM.jl:805)
extern "C" int spasm_dense_rref(i64 prime, int n, int m, spasm_ZZp *A, i64 ldA, int *pivcol_out) {
  try {
    require_gpu();
    Fp F = make_field(prime);
    DBuf<int> tmp((size_t)n * m);
    DBuf<uint32_t> S((size_t)n * m);
    std::vector<int> host((size_t)n * m);
    for (int r = 0; r < n; r++) memcpy(host.data() + (size_t)r * m, A + (size_t)r * ldA, (size_t)m * sizeof(int));
    std::vector<uint32_t> hu((size_t)n * m);
    for (size_t i = 0; i < hu.size(); i++) hu[i] = to_u(host[i], F);
    S.upload(hu.data(), hu.size());
    DBuf<int> pivcol, pivrow;
    const int rr = dense_rref_device(S.p, n, m, m, F, pivcol, pivrow);
    S.download(hu.data(), hu.size());
    std::vector<int> pc(std::max(rr, 1)), pr(std::max(rr, 1));
    if (rr) pivcol.download(pc.data(), rr), pivrow.download(pr.data(), rr);
    sync();
    // reduced rows first (by increasing pivot column), the rest are zero
    for (int r = 0; r < n; r++)
      for (int k = 0; k < m; k++) A[(size_t)r * ldA + k] = 0;
    for (int t = 0; t < rr; t++) {
      for (int k = 0; k < m; k++) A[(size_t)t * ldA + k] = to_bal(hu[(size_t)pr[t] * m + k], F);
      pivcol_out[t] = pc[t];
    }
    return rr;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_dense_rref failed: %s\n", e.what());
    return -1;
  }
}
