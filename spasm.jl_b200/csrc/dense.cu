// dense.cu — dense tail, first version: Gauss-Jordan by pivot steps on CUDA cores.
// (The blocked tcgen05 path that replaces the elimination step lives in dense_mma.cu.)
#include <cooperative_groups.h>

#include "dense.cuh"
#include "dist.cuh"
#include "sink.cuh"

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

void dense_tail_core(DenseSchur &D, int nrows, int n_local, int m_total, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size,
                     const TailOpts &opts);
// The arithmetic of the deferred trailing updates (pure host logic, exported as spasm_b200_tail_plan for the CPU tests):
// kcap = wanted flush depth, max_k = deepest product one tensor-core launch takes for this prime.
TailPlan plan_tail(int Sm0, int n_local, int block_size, int Bmax, int NR, int kcap, int max_k, size_t free_bytes) {
  TailPlan pl;
  const int B16 = (Bmax + 15) / 16 * 16;
  kcap = std::max(0, std::min(kcap, max_k - B16));
  while (kcap >= 2 * B16 && (size_t)(Sm0 + n_local + B16) * (size_t)(kcap + B16) * 8 > free_bytes / 4) kcap /= 2;  // keep the (two sets of) factor buffers small
  pl.lazy = kcap >= 2 * B16 && n_local > 2 * block_size;
  pl.group = pl.lazy ? std::max(1, kcap / std::max(block_size * NR, 1)) : 1;
  pl.kdepth = pl.lazy ? (int)std::min<long long>(std::max<long long>(kcap, ((long long)pl.group * NR + NR - 1) * B16), max_k - B16) : 0;
  pl.LDK = pl.lazy ? (long long)pl.kdepth + B16 : 0;
  return pl;
}
// how the rows of the low-rank mode are spread when the dense loop switches to it on several ranks
struct LowRankShard {
  const int *gidx = nullptr;         // device: global index (in the remaining-row list) of local row i; nullptr = identity
  const int *loc_of_glob = nullptr;  // device: local row of global row k, -1 when another rank holds it; nullptr = identity
  bool reduce = false;               // a combined block is the sum (mod p) of the ranks' parts
  long long panel_base = 0;          // panel counter for the who-materialises rule of the sharded factor
};
void lowrank_core(DenseSchur &D, long long row0, int nloc, int nglob, const LowRankShard &sh, unsigned char *colpiv, DCsr &U,
                  DBuf<int> &Uqinv, const Fp &F, int block_size, double start_weight, int m_total);
// counters of the deferred trailing updates since the last reset (tests assert that the far-row path ran):
// [0] far flushes with far rows  [1] multiplier-correction GEMMs  [2] far rows x depth flushed  [3] near updates
long long g_tail_stats[5] = {0, 0, 0, 0, 0};

void make_free_columns(const int *qinv, int m, int *flag, long long *pos, int *q, int *qpos) {
  cudaStream_t s = stream();
  k_free_cols<<<cdiv(m + 1, 256), 256, 0, s>>>(qinv, m, flag);
  exclusive_scan_i32_to_i64(flag, pos, m + 1);
  k_make_q<<<cdiv(m, 256), 256, 0, s>>>(flag, pos, m, q, qpos);
  CK(cudaGetLastError());
}

// ------------------------------------------------------------------ GEMM mod p on CUDA cores
// 64x64 output tile, K chunks of 32 through shared memory, 4x4 outputs per thread.  Products of
// residues below 2^16 fit 32 bits and are summed in 64-bit accumulators with ONE reduction at the
// end; larger primes reduce every product.  (The tcgen05 int8-limb kernel in dense_mma.cu takes
// over for p < 2^16 and large shapes.)
template <bool SMALL>
__global__ void __launch_bounds__(256) k_gemm_nt(uint32_t *__restrict__ C, long long ldc, int M, int N, const uint32_t *__restrict__ A,
                                                  long long lda, const uint32_t *__restrict__ B, long long ldb, int K, int subtract,
                                                  const int *__restrict__ rowmap, Fp F) {
  __shared__ uint32_t As[64][33], Bs[64][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  unsigned long long acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += 256) {
      const int rr = idx >> 5, kk = idx & 31;
      const int gi = i0 + rr, gj = j0 + rr, gk = k0 + kk;
      As[rr][kk] = (gi < M && gk < K) ? A[(long long)(rowmap ? rowmap[gi] : gi) * lda + gk] : 0u;
      Bs[rr][kk] = (gj < N && gk < K) ? B[(long long)gj * ldb + gk] : 0u;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; kk++) {
      uint32_t av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; a++) av[a] = As[ty + 16 * a][kk];
#pragma unroll
      for (int b = 0; b < 4; b++) bv[b] = Bs[tx + 16 * b][kk];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          if (SMALL)
            acc[a][b] += (unsigned long long)(av[a] * bv[b]);
          else
            acc[a][b] += mulmod<false>(av[a], bv[b], F);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int gi = i0 + ty + 16 * a, gj = j0 + tx + 16 * b;
      if (gi < M && gj < N) {
        uint32_t v = red64(acc[a][b], F);
        uint32_t *c = C + (long long)(rowmap ? rowmap[gi] : gi) * ldc + gj;
        *c = subtract ? addmod(*c, negmod(v, F), F) : v;
      }
    }
}

bool gemm_nt_mma(uint32_t *C, long long ldc, int M, int N, const uint32_t *A, long long lda, const uint32_t *B, long long ldb, int K,
                 bool subtract, const Fp &F, const int *rowmap);  // dense_mma.cu: returns false when the shape / prime is not handled
extern int g_gemm_cta_limit;  // dense_mma.cu: > 0 caps the CTAs (= SMs) of the persistent tensor-core kernel
// while alive, the tensor-core launches use at most `n` SMs (0: no change) so that another stream's kernels find room
struct GemmCtaLimit {
  int prev;
  explicit GemmCtaLimit(int n) : prev(g_gemm_cta_limit) {
    if (n > 0) g_gemm_cta_limit = n;
  }
  ~GemmCtaLimit() { g_gemm_cta_limit = prev; }
};

void gemm_nt(uint32_t *C, long long ldc, int M, int N, const uint32_t *A, long long lda, const uint32_t *B, long long ldb, int K,
             bool subtract, const Fp &F, const int *rowmap) {
  if (M <= 0 || N <= 0) return;
  if (gemm_nt_mma(C, ldc, M, N, A, lda, B, ldb, K, subtract, F, rowmap)) {
    g_launches += 3;  // two limb splits + the tcgen05 kernel
    return;
  }
  g_launches += 1;
  dim3 grid(cdiv(N, 64), cdiv(M, 64));
  if (F.small)
    k_gemm_nt<true><<<grid, 256, 0, stream()>>>(C, ldc, M, N, A, lda, B, ldb, K, subtract, rowmap, F);
  else
    k_gemm_nt<false><<<grid, 256, 0, stream()>>>(C, ldc, M, N, A, lda, B, ldb, K, subtract, rowmap, F);
  CK(cudaGetLastError());
}

// ------------------------------------------------------------------ panel factorisation
// The panel is rows [k0, k0+Sn) of D, i.e. Dt[c][k0 + t].  T (Sn x Sn) accumulates the row
// operations: T.Panel is in reduced row echelon form on the columns scanned so far.  Columns are
// scanned left to right in tiles; for each tile  W = T.Panel_tile  (GEMM), one CTA runs
// Gauss-Jordan on W among the rows that are not pivots yet (at most 32 new pivots per tile) while
// recording the operations as the columns Gc of the update  T <- G.T.
static constexpr int PB = 32;   // pivots per tile call (number of recorded operation columns Gc)
static constexpr int WKS = 4;   // K slices of the narrow W product (k_wtile): 4x more CTAs, summed when the tile is loaded
static constexpr int WMAX = 2048;
struct PanelCtl {
  int npiv;      // pivots found so far in this panel
  int found;     // pivots found in the last tile
  int consumed;  // columns consumed by the last tile
  int c0;        // first column not scanned yet (advanced on the device: no host round trip per tile)
  int low;       // tiles that found fewer than PB/2 pivots (time to widen the tiles)
  int pad[3];
};
// Gauss-Jordan on one tile.  The tile is stored COLUMN-major (Wt[c][r], r contiguous) and every
// thread owns rows r = tid, tid+1024, ...: finding the pivot is one coalesced pass over a column,
// eliminating it is wc-cc independent coalesced read-modify-writes per thread (no dependent chains).
// The recorded operations Gc[s][r] (column s = s-th pivot of this call) define T <- G.T.
// colflag[c] = 1 iff column c of the tile is nonzero on some row that is not a pivot yet.  A column
// that is zero there stays zero under the eliminations of the tile (the pivot row is one of those rows).
__global__ void k_col_flags(const uint32_t *__restrict__ Wt, long long ldw, int Sn, const int *__restrict__ ispiv, unsigned char *__restrict__ flag) {
  const uint32_t *col = Wt + (long long)blockIdx.x * ldw;
  int any = 0;
  for (int r = threadIdx.x; r < Sn; r += blockDim.x) any |= (!ispiv[r] && col[r] != 0);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flag[blockIdx.x] = any ? 1 : 0;
}
template <bool SMALL>
__global__ void __launch_bounds__(1024) k_tile_gauss(uint32_t *__restrict__ Wt, int Sn, int wc, long long ldw, int c0,
                                                      int *__restrict__ ispiv, int *__restrict__ pivrow, int *__restrict__ pivcol,
                                                      uint32_t *__restrict__ Gc, int *__restrict__ tilepiv, PanelCtl *__restrict__ ctl,
                                                      const unsigned char *__restrict__ colflag, const int *__restrict__ cand, Fp F,
                                                      uint32_t *__restrict__ GjT, long long ldg, int *__restrict__ pividx) {
  __shared__ int red[32];
  __shared__ int s_piv;
  __shared__ uint32_t prow[WMAX], gprow[PB];
  __shared__ unsigned char s_flag[WMAX];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int npiv = ctl->npiv, found = 0, cc = 0;
  c0 = ctl->c0;
  for (long long idx = tid; idx < (long long)PB * ldw; idx += 1024) Gc[idx] = 0;
  for (int c = tid; c < wc; c += 1024) s_flag[c] = colflag[c];
  __syncthreads();
  for (; cc < wc && found < PB && npiv < Sn; cc++) {
    if (!s_flag[cc]) continue;
    int best = 0x7fffffff;
    const uint32_t *col = Wt + (long long)cc * ldw;
    for (int r = tid; r < Sn; r += 1024)
      if (!ispiv[r] && col[r] != 0) {
        best = r;
        break;
      }
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) red[wid] = best;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < 32; i++) best = min(best, red[i]);
      s_piv = best;
      if (best != 0x7fffffff) {
        ispiv[best] = 1;
        pivrow[npiv] = best;
        pivcol[npiv] = cand[c0 + cc];
        tilepiv[found] = best;
        Gc[(long long)found * ldw + best] = 1;
        if (pividx) pividx[best] = npiv;
      }
    }
    __syncthreads();
    const int pr = s_piv;
    if (pr == 0x7fffffff) continue;
    // the scaled pivot row goes to shared memory (and back to the tile)
    const uint32_t alpha = dev_inv(Wt[(long long)cc * ldw + pr], F.p);
    for (int k = cc + tid; k < wc; k += 1024) {
      uint32_t v = mulmod<SMALL>(alpha, Wt[(long long)k * ldw + pr], F);
      prow[k] = v;
      Wt[(long long)k * ldw + pr] = v;
    }
    if (tid <= found) {
      uint32_t v = mulmod<SMALL>(alpha, Gc[(long long)tid * ldw + pr], F);
      gprow[tid] = v;
      Gc[(long long)tid * ldw + pr] = v;
    }
    __syncthreads();
    for (int r = tid; r < Sn; r += 1024) {
      if (r == pr) continue;
      const uint32_t f = Wt[(long long)cc * ldw + r];
      if (f == 0) continue;
      if (GjT != nullptr && pividx[r] >= 0) GjT[(long long)npiv * ldg + pividx[r]] = f;  // what this pivot removes from an earlier pivot row
      const uint32_t nf = negmod(f, F);
      for (int k = cc; k < wc; k++) {
        uint32_t *w = Wt + (long long)k * ldw + r;
        *w = addmod(*w, mulmod<SMALL>(nf, prow[k], F), F);
      }
      for (int c = 0; c <= found; c++) {
        uint32_t *gp = Gc + (long long)c * ldw + r;
        *gp = addmod(*gp, mulmod<SMALL>(nf, gprow[c], F), F);
      }
    }
    found++;
    npiv++;
    __syncthreads();
  }
  if (tid == 0) ctl->npiv = npiv, ctl->found = found, ctl->consumed = cc, ctl->c0 = c0 + cc;
}
// shared-memory version for p < 2^16 and panels of at most 1024 rows (the default block size is
// 1000): the whole tile [W | Gc] lives in smem as u16, column-major, one thread per row.
__global__ void __launch_bounds__(1024) k_tile_gauss_smem(const uint32_t *__restrict__ Wt, int Sn, int Sm0, long long ldw,
                                                           int *__restrict__ ispiv, int *__restrict__ pivrow, int *__restrict__ pivcol,
                                                           uint32_t *__restrict__ Gc, int *__restrict__ tilepiv, PanelCtl *__restrict__ ctl,
                                                           const int *__restrict__ cand, Fp F, uint32_t *__restrict__ GjT, long long ldg,
                                                           int *__restrict__ pividx) {
  extern __shared__ unsigned short tile[];  // [32 + PB][SP]
  __shared__ int red[32];
  __shared__ int s_piv;
  __shared__ uint32_t prow[32 + PB];
  const int SP = 1024;
  const int r = threadIdx.x, lane = r & 31, wid = r >> 5;
  const bool live = r < Sn;
  const int c0 = ctl->c0;
  if (ctl->npiv >= Sn || c0 >= Sm0) {  // panel finished: this launch of the group is a no-op
    __syncthreads();
    if (r == 0) ctl->found = 0, ctl->consumed = 0;
    return;
  }
  const int wc = min(32, Sm0 - c0);
  for (int c = 0; c < 32; c++) {
    uint32_t v = 0;
    if (live && c < wc) {
      for (int z = 0; z < WKS; z++) v += Wt[((long long)z * 32 + c) * ldw + r];  // KS partial products, each < p < 2^16
      v %= F.p;
    }
    tile[c * SP + r] = (unsigned short)v;
  }
  for (int s = 0; s < PB; s++) tile[(32 + s) * SP + r] = 0;
  int my_ispiv = live ? ispiv[r] : 1;
  int my_pividx = (live && pividx != nullptr) ? pividx[r] : -1;
  int npiv = ctl->npiv, found = 0, cc = 0;
  __syncthreads();
  for (; cc < wc && found < PB && npiv < Sn; cc++) {
    int best = (!my_ispiv && tile[cc * SP + r] != 0) ? r : 0x7fffffff;
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) red[wid] = best;
    __syncthreads();
    if (r < 32) {
      int b2 = red[r];
      for (int o = 16; o; o >>= 1) b2 = min(b2, __shfl_xor_sync(0xffffffffu, b2, o));
      if (r == 0) s_piv = b2;
    }
    __syncthreads();
    const int pr = s_piv;
    if (pr == 0x7fffffff) continue;
    if (r == pr) {
      my_ispiv = 1;
      ispiv[r] = 1;
      pivrow[npiv] = r;
      pivcol[npiv] = cand[c0 + cc];
      tilepiv[found] = r;
      tile[(32 + found) * SP + r] = 1;
      if (pividx != nullptr) pividx[r] = npiv, my_pividx = npiv;
    }
    __syncthreads();
    if (r < 32 + PB) {
      // the scaled pivot row (columns cc..wc of W and 0..found of Gc)
      const bool used = (r < 32) ? (r >= cc && r < wc) : (r - 32 <= found);
      uint32_t v = 0;
      if (used) {
        const uint32_t alpha = dev_inv_small(tile[cc * SP + pr], F);
        v = mulmod<true>(alpha, tile[r * SP + pr], F);
      }
      prow[r] = v;
    }
    __syncthreads();
    if (live) {
      if (r == pr) {
        for (int k = cc; k < wc; k++) tile[k * SP + r] = (unsigned short)prow[k];
        for (int s = 0; s <= found; s++) tile[(32 + s) * SP + r] = (unsigned short)prow[32 + s];
      } else {
        const uint32_t f = tile[cc * SP + r];
        if (f != 0) {
          if (GjT != nullptr && my_pividx >= 0) GjT[(long long)npiv * ldg + my_pividx] = f;
          const uint32_t nf = F.p - f;
          for (int k = cc; k < wc; k++) {
            uint32_t t = (uint32_t)tile[k * SP + r] + mulmod<true>(nf, prow[k], F);
            tile[k * SP + r] = (unsigned short)(t >= F.p ? t - F.p : t);
          }
          for (int s = 0; s <= found; s++) {
            uint32_t t = (uint32_t)tile[(32 + s) * SP + r] + mulmod<true>(nf, prow[32 + s], F);
            tile[(32 + s) * SP + r] = (unsigned short)(t >= F.p ? t - F.p : t);
          }
        }
      }
    }
    found++;
    npiv++;
    __syncthreads();
  }
  if (live)
    for (int s = 0; s < found; s++) Gc[(long long)s * ldw + r] = tile[(32 + s) * SP + r];
  if (r == 0) {
    ctl->npiv = npiv, ctl->found = found, ctl->consumed = cc, ctl->c0 = c0 + cc;
    if (found < PB / 2 && npiv < Sn) ctl->low += 1;
  }
}

// CLUSTER version of the tile Gauss-Jordan: the (<= 1024) rows are split over a thread-block cluster of 8 CTAs, the
// tile slice of 128 rows in each CTA's shared memory, FOUR threads per row (each owns every fourth column).
// ONE cluster barrier per pivot: every CTA finds its own first candidate row, scales it (one modular inverse, computed
// while the other CTAs do the same) and publishes candidate index + scaled row into all eight CTAs' shared memory
// (DSMEM); after the barrier everybody knows the winner (smallest row index) and already holds its scaled row, and
// eliminates its own rows.  (The version of round 1 needed two barriers per pivot — candidates, then the winner's row —
// with the inverse on the path between them, and one thread per row: 4.7 us per pivot.)
static constexpr int GC = 8;     // CTAs per cluster
static constexpr int GRP = 128;  // rows per CTA
static constexpr int TPR = 4;    // threads per row
__global__ void __cluster_dims__(GC, 1, 1) __launch_bounds__(GRP * TPR)
k_tile_gauss_cluster(const uint32_t *__restrict__ Wt, int Sn, int Sm0, long long ldw, int *__restrict__ ispiv, int *__restrict__ pivrow,
                     int *__restrict__ pivcol, uint32_t *__restrict__ Gc, int *__restrict__ tilepiv, PanelCtl *__restrict__ ctl,
                     const int *__restrict__ cand, Fp F, uint32_t *__restrict__ GjT, long long ldg, int *__restrict__ pividx) {
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  __shared__ unsigned short tile[32 + PB][GRP];
  __shared__ int s_best[2][GC];                  // candidate row of every CTA (double buffered by pivot parity)
  __shared__ uint32_t s_rows[2][GC][32 + PB];    // ... and that row, scaled: columns of W, then the recorded operations
  __shared__ int red[GRP / 32];
  const int cta = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int rl = tid & (GRP - 1), q = tid / GRP;  // my row (local) and which columns of it I own
  const int r = cta * GRP + rl;
  const bool live = r < Sn;
  const int c0 = ctl->c0;
  const int npiv0 = ctl->npiv;
  if (npiv0 >= Sn || c0 >= Sm0) {  // panel finished: no-op launch (uniform over the cluster)
    if (cta == 0 && tid == 0) ctl->found = 0, ctl->consumed = 0;
    return;
  }
  const int wc = min(32, Sm0 - c0);
  for (int c = q; c < 32; c += TPR) {
    uint32_t v = 0;
    if (live && c < wc) {
      for (int z = 0; z < WKS; z++) v += Wt[((long long)z * 32 + c) * ldw + r];  // KS partial products, each < p < 2^16
      v %= F.p;
    }
    tile[c][rl] = (unsigned short)v;
  }
  for (int sx = q; sx < PB; sx += TPR) tile[32 + sx][rl] = 0;
  int my_ispiv = live ? ispiv[r] : 1;
  int my_pividx = (live && pividx != nullptr) ? pividx[r] : -1;
  int npiv = npiv0, found = 0, cc = 0, step = 0;
  cluster.sync();
  for (; cc < wc && found < PB && npiv < Sn; cc++) {
    const int par = step & 1;
    // ---- this CTA's candidate: its first row that is not a pivot yet and is non-zero on this column
    if (tid < GRP) {
      int best = (!my_ispiv && tile[cc][rl] != 0) ? rl : 0x7fffffff;
      for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (lane == 0) red[wid] = best;
    }
    __syncthreads();
    int b2 = red[0];
#pragma unroll
    for (int i = 1; i < GRP / 32; i++) b2 = min(b2, red[i]);
    // ---- publish it (index + scaled row) to every CTA of the cluster
    if (tid < 32 + PB) {
      uint32_t v = 0;
      if (b2 != 0x7fffffff) {
        const bool used = (tid < 32) ? (tid > cc && tid < wc) : (tid - 32 <= found);
        if (used) {
          const uint32_t alpha = dev_inv_small(tile[cc][b2], F);
          const uint32_t e = (tid - 32 == found) ? 1u : (uint32_t)tile[tid][b2];  // (the new operation column starts as the unit vector of the pivot row)
          v = mulmod<true>(alpha, e, F);
        }
      }
#pragma unroll
      for (int dst = 0; dst < GC; dst++) cluster.map_shared_rank(&s_rows[par][cta][0], dst)[tid] = v;
    } else if (tid >= 96 && tid < 96 + GC) {
      cluster.map_shared_rank(&s_best[par][0], tid - 96)[cta] = (b2 == 0x7fffffff) ? 0x7fffffff : cta * GRP + b2;
    }
    cluster.sync();
    int pr = s_best[par][0];
#pragma unroll
    for (int i = 1; i < GC; i++) pr = min(pr, s_best[par][i]);
    step++;
    if (pr == 0x7fffffff) continue;  // no pivot on this column (uniform decision)
    const uint32_t *prow = s_rows[par][pr / GRP];
    if (live) {
      if (r == pr) {
        my_ispiv = 1;
        my_pividx = npiv;
        if (q == 0) {
          ispiv[r] = 1;
          pivrow[npiv] = r;
          pivcol[npiv] = cand[c0 + cc];
          tilepiv[found] = r;
          if (pividx != nullptr) pividx[r] = npiv;
        }
        for (int k = cc + 1 + q; k < wc; k += TPR) tile[k][rl] = (unsigned short)prow[k];
        for (int sx = q; sx <= found; sx += TPR) tile[32 + sx][rl] = (unsigned short)prow[32 + sx];
      } else {
        const uint32_t f = tile[cc][rl];  // (column cc itself is never read again: it is not updated)
        if (f != 0) {
          if (q == 0 && GjT != nullptr && my_pividx >= 0) GjT[(long long)npiv * ldg + my_pividx] = f;
          const uint32_t nf = F.p - f;
          for (int k = cc + 1 + q; k < wc; k += TPR) {
            uint32_t t = (uint32_t)tile[k][rl] + mulmod<true>(nf, prow[k], F);
            tile[k][rl] = (unsigned short)(t >= F.p ? t - F.p : t);
          }
          for (int sx = q; sx <= found; sx += TPR) {
            uint32_t t = (uint32_t)tile[32 + sx][rl] + mulmod<true>(nf, prow[32 + sx], F);
            tile[32 + sx][rl] = (unsigned short)(t >= F.p ? t - F.p : t);
          }
        }
      }
    }
    found++;
    npiv++;
    __syncthreads();  // the next column is read by threads that did not write it
  }
  __syncthreads();
  if (live)
    for (int sx = q; sx < found; sx += TPR) Gc[(long long)sx * ldw + r] = tile[32 + sx][rl];
  if (cta == 0 && tid == 0) {
    ctl->npiv = npiv, ctl->found = found, ctl->consumed = cc, ctl->c0 = c0 + cc;
    if (found < PB / 2 && npiv < Sn) ctl->low += 1;
  }
  cluster.sync();  // nobody leaves while its shared memory may still be written remotely
}

template <bool SMALL>
__global__ void __launch_bounds__(256) k_wtile(const uint32_t *__restrict__ Dt_panel, long long ld, const uint32_t *__restrict__ T, int Sn, int Sm0,
                                                const PanelCtl *__restrict__ ctl, const int *__restrict__ cand, uint32_t *__restrict__ Wt,
                                                long long ldw, Fp F) {
  const int c0 = ctl->c0;
  if (ctl->npiv >= Sn || c0 >= Sm0) return;
  const int wc = min(32, Sm0 - c0);
  if ((int)blockIdx.y * 8 >= wc) return;
  // blockIdx.y selects 8 of the <=32 columns: 4x more CTAs in flight for this latency-bound product
  __shared__ uint32_t Ts[32][33], Ds[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // tx: r within the tile, ty: column within the group
  const int r0 = blockIdx.x * 32, cg = blockIdx.y * 8;
  // blockIdx.z takes one of the KS slices of the K range; the partial sums land in separate planes of Wt
  const int kslice = ((Sn + 31) / 32 + gridDim.z - 1) / gridDim.z * 32;
  const int tbeg = blockIdx.z * kslice, tend = min(Sn, tbeg + kslice);
  Wt += (long long)blockIdx.z * 32 * ldw;
  unsigned long long acc = 0;
  for (int t0 = tbeg; t0 < tend; t0 += 32) {
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 256) {
      const int a = idx >> 5, t = idx & 31;
      Ts[a][t] = (r0 + a < Sn && t0 + t < tend) ? T[(long long)(r0 + a) * Sn + t0 + t] : 0u;
    }
    {
      const int a = threadIdx.x >> 5, t = threadIdx.x & 31;
      Ds[a][t] = (cg + a < wc && t0 + t < tend) ? Dt_panel[(long long)cand[c0 + cg + a] * ld + t0 + t] : 0u;
    }
    __syncthreads();
#pragma unroll 8
    for (int t = 0; t < 32; t++) {
      if (SMALL)
        acc += (unsigned long long)(Ts[tx][t] * Ds[ty][t]);
      else
        acc += mulmod<false>(Ts[tx][t], Ds[ty][t], F);
    }
    __syncthreads();
  }
  if (r0 + tx < Sn && cg + ty < wc) Wt[(long long)(cg + ty) * ldw + r0 + tx] = red64(acc, F);
}

__global__ void k_set_identity(uint32_t *T, int n) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx < (long long)n * n) T[idx] = (idx / n == idx % n) ? 1u : 0u;
}
__global__ void k_gather_T_rows(const uint32_t *__restrict__ T, int Sn, const int *__restrict__ rowsel, const int *__restrict__ count_ptr,
                                int count_fixed, uint32_t *__restrict__ out) {
  const int cnt = count_ptr ? *count_ptr : count_fixed;
  const int s = blockIdx.y;
  if (s >= cnt) return;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < Sn) out[(long long)s * Sn + col] = T[(long long)rowsel[s] * Sn + col];
}
// T[r][:] = (r is a pivot of this tile ? 0 : T[r][:]) + sum_s Gc[r][s] * Tp[s][:]
template <bool SMALL>
__global__ void k_update_T(uint32_t *__restrict__ T, int Sn, const uint32_t *__restrict__ Gc, long long ldg, const uint32_t *__restrict__ Tp,
                           const int *__restrict__ tilepiv, const PanelCtl *__restrict__ ctl, Fp F) {
  const int found = ctl->found;
  if (found == 0) return;
  const int r = blockIdx.y, col = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ int s_isp;
  if (threadIdx.x == 0) {
    int f = 0;
    for (int s = 0; s < found; s++) f |= (tilepiv[s] == r);
    s_isp = f;
  }
  __syncthreads();
  if (col >= Sn) return;
  unsigned long long acc = s_isp ? 0ull : T[(long long)r * Sn + col];
  for (int s = 0; s < found; s++) {
    const uint32_t g = Gc[(long long)s * ldg + r];
    if (g == 0) continue;
    if (SMALL)
      acc += (unsigned long long)(g * Tp[(long long)s * Sn + col]);
    else
      acc += mulmod<false>(g, Tp[(long long)s * Sn + col], F);
  }
  T[(long long)r * Sn + col] = red64(acc, F);
}
__global__ void k_transpose_u32(const uint32_t *__restrict__ in, long long ldi, int rows, int cols, uint32_t *__restrict__ out, long long ldo) {
  // out[c][r] = in[r][c]
  __shared__ uint32_t tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int y = threadIdx.y; y < 32; y += 8) {
    int r = r0 + y, c = c0 + threadIdx.x;
    tile[y][threadIdx.x] = (r < rows && c < cols) ? in[(long long)r * ldi + c] : 0u;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    int c = c0 + y, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(long long)c * ldo + r] = tile[threadIdx.x][y];
  }
}
// Pt[k][s] = Dt[pivcol[s]][kbase + k]
__global__ void k_gather_pivot_cols_T(const uint32_t *__restrict__ Dt, long long ld, const int *__restrict__ pivcol, int rr, long long kbase,
                                      int nk, uint32_t *__restrict__ Pt, long long ldp) {
  __shared__ uint32_t tile[32][33];
  const int k0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  for (int y = threadIdx.y; y < 32; y += 8) {
    int s = s0 + y, k = k0 + threadIdx.x;
    tile[y][threadIdx.x] = (s < rr && k < nk) ? Dt[(long long)pivcol[s] * ld + kbase + k] : 0u;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    int k = k0 + y, s = s0 + threadIdx.x;
    if (k < nk && s < rr) Pt[(long long)k * ldp + s] = tile[threadIdx.x][y];
  }
}
// Dt[pivcol[s]][kbase + k] = 0: the columns pivoted by this panel are dead on every remaining row (the trailing
// products skip them), and later panels read R = T . Dt over all columns
__global__ void k_zero_pivot_cols(uint32_t *__restrict__ Dt, long long ld, const int *__restrict__ pivcol, int rr, long long kbase, int nk) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (k < nk && s < rr) Dt[(long long)pivcol[s] * ld + kbase + k] = 0u;
}
// the factored panel on the wire: live columns only (dead columns of R are zero), 16 bits per entry when p < 2^16
template <class T>
__global__ void k_pack_R(const uint32_t *__restrict__ R, long long ldr, const int *__restrict__ cand, int nlive, int rr, T *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (t < nlive && s < rr) out[(long long)s * nlive + t] = (T)R[(long long)s * ldr + cand[t]];
}
template <class T>
__global__ void k_unpack_R(const T *__restrict__ in, const int *__restrict__ cand, int nlive, int rr, uint32_t *__restrict__ R, long long ldr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (t < nlive && s < rr) R[(long long)s * ldr + cand[t]] = (uint32_t)in[(long long)s * nlive + t];
}
// out[s][u] = in[idx[s]][u], u < cols
__global__ void k_gather_rows_ld(const uint32_t *__restrict__ in, long long ldi, const int *__restrict__ idx, int rows, int cols,
                                 uint32_t *__restrict__ out, long long ldo) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, sidx = blockIdx.y;
  if (u < cols && sidx < rows) out[(long long)sidx * ldo + u] = in[(long long)idx[sidx] * ldi + u];
}
__global__ void k_iota2(int *a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// returns rr; T final, pivrow/pivcol filled (pivcol increasing)
__global__ void k_gather_rows_u32(const uint32_t *__restrict__ Dt_panel, long long ld, const int *__restrict__ cand, int wc, int Sn,
                                  uint32_t *__restrict__ out) {
  const int c = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < wc && t < Sn) out[(long long)c * Sn + t] = Dt_panel[(long long)cand[c] * ld + t];
}
__global__ void k_not_flag(const unsigned char *__restrict__ colpiv, int n, int *__restrict__ flag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) flag[c] = !colpiv[c];
  if (c == n) flag[c] = 0;
}
__global__ void k_mark_cols(const int *__restrict__ pivcol, int rr, unsigned char *__restrict__ colpiv) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < rr) colpiv[pivcol[s]] = 1;
}

// cand[0:Sm0) = the columns that are not pivots of an earlier panel (those are zero on this panel), increasing
// ---- with-L mode.  The panel is factored in REDUCED form as always (R = reduced rows); the row echelon form the
// factorisation A = L.U needs is recovered from the Jordan multipliers G (strictly upper: G[s][s'] = what pivot s' removed
// from pivot row s):  R = (I - G).Ub  =>  Ub = U_M.R  with  U_M = (I - G)^-1  (= the pivot columns of Ub, unit upper
// triangular), and the multipliers of ANY row on the panel's pivots are  (its entries on the pivot columns).(I - G).
// U_M[s][j], one thread per column j, by back substitution:  U_M[s][j] = sum_{s < t <= j} G[s][t] . U_M[t][j]
template <bool SMALL>
__global__ void k_um_from_gT(const uint32_t *__restrict__ GjT, long long ldg, int rr, uint32_t *__restrict__ UM, long long ldu, Fp F) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= rr) return;
  for (int s = rr - 1; s > j; s--) UM[(long long)s * ldu + j] = 0;
  UM[(long long)j * ldu + j] = 1;
  for (int s = j - 1; s >= 0; s--) {
    uint32_t acc = 0;
    for (int t = s + 1; t <= j; t++) {
      const uint32_t g = GjT[(long long)t * ldg + s];
      if (g) acc = addmod(acc, mulmod<SMALL>(g, UM[(long long)t * ldu + j], F), F);
    }
    UM[(long long)s * ldu + j] = acc;
  }
}
// in place:  GjT -> (I - G)^T   (row s', column s:  1 on the diagonal, -G[s][s'] below it, 0 above)
__global__ void k_make_uinvT(uint32_t *__restrict__ GjT, long long ldg, int rr, uint32_t p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, s2 = blockIdx.y;
  if (s >= rr || s2 >= rr) return;
  uint32_t *q = GjT + (long long)s2 * ldg + s;
  const uint32_t g = *q;
  *q = (s == s2) ? 1u : (s < s2 ? (g ? p - g : 0u) : 0u);
}
__global__ void k_fill_int(int *a, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// GjT / pividx (optional, with-L mode): GjT[s'][s] = what pivot s' removed from the earlier pivot row s (Gauss-JORDAN
// multipliers; zero-initialised by the caller), pividx[row] = index of the pivot that row holds (-1-initialised)
static int panel_factor(const uint32_t *Dt, long long ld, const int *cand, int Sm0, long long k0, int Sn, uint32_t *T, DBuf<int> &ispiv,
                        DBuf<int> &pivrow, DBuf<int> &pivcol, const Fp &F, uint32_t *GjT = nullptr, long long ldg = 0, int *pividx = nullptr) {
  cudaStream_t s = stream();
  const long long ldw = ((long long)Sn + 31) / 32 * 32;
  DBuf<uint32_t> Wt((size_t)WMAX * ldw), Gc((size_t)PB * ldw), Tp((size_t)PB * Sn), Agather((size_t)WMAX * Sn);
  DBuf<int> tilepiv(PB);
  DBuf<unsigned char> colflag(WMAX);
  DBuf<PanelCtl> ctl(1);
  ctl.zero();
  ispiv.alloc(Sn);
  ispiv.zero();
  pivrow.alloc(Sn);
  pivcol.alloc(Sn);
  k_set_identity<<<cdiv((long long)Sn * Sn, 256), 256, 0, s>>>(T, Sn);
  const bool fast = F.small && Sn <= 1024;
  static const bool use_cluster = getenv("SPASM_B200_NO_CLUSTER") == nullptr;
  static bool attr = false;
  const size_t gsm = (size_t)(32 + PB) * 1024 * sizeof(unsigned short);
  if (fast && !attr) {
    CK(cudaFuncSetAttribute(k_tile_gauss_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    attr = true;
  }
  auto apply_T = [&]() {
    k_gather_T_rows<<<dim3(cdiv(Sn, 256), PB), 256, 0, s>>>(T, Sn, tilepiv.p, &ctl.p->found, 0, Tp.p);
    if (F.small)
      k_update_T<true><<<dim3(cdiv(Sn, 256), Sn), 256, 0, s>>>(T, Sn, Gc.p, ldw, Tp.p, tilepiv.p, ctl.p, F);
    else
      k_update_T<false><<<dim3(cdiv(Sn, 256), Sn), 256, 0, s>>>(T, Sn, Gc.p, ldw, Tp.p, tilepiv.p, ctl.p, F);
  };
  PanelCtl h{};
  // ---- dense phase: 32-column tiles, 8 tiles per host round trip, cursor advanced on the device
  while (fast && h.npiv < Sn && h.c0 < Sm0 && h.low == 0) {
    for (int g = 0; g < 8; g++) {
      k_wtile<true><<<dim3(cdiv(Sn, 32), 4, WKS), 256, 0, s>>>(Dt + k0, ld, T, Sn, Sm0, ctl.p, cand, Wt.p, ldw, F);
      if (use_cluster)
        k_tile_gauss_cluster<<<GC, GRP * TPR, 0, s>>>(Wt.p, Sn, Sm0, ldw, ispiv.p, pivrow.p, pivcol.p, Gc.p, tilepiv.p, ctl.p, cand, F, GjT, ldg, pividx);
      else
        k_tile_gauss_smem<<<1, 1024, gsm, s>>>(Wt.p, Sn, Sm0, ldw, ispiv.p, pivrow.p, pivcol.p, Gc.p, tilepiv.p, ctl.p, cand, F, GjT, ldg, pividx);
      apply_T();
    }
    CK(cudaGetLastError());
    g_launches += 32;
    h = fetch(ctl.p);
  }
  // ---- sparse phase (few pivots per tile, or a prime / panel the fast path does not take): adaptive width
  int w = 32;
  while (h.npiv < Sn && h.c0 < Sm0) {
    const int c0 = h.c0;
    const int wc = std::min(w, Sm0 - c0);
    // Wt[c][r] = sum_t Dt[cand[c0+c]][k0+t] * T[r][t]     (the tile of T.Panel, transposed)
    k_gather_rows_u32<<<dim3(cdiv(Sn, 256), wc), 256, 0, s>>>(Dt + k0, ld, cand + c0, wc, Sn, Agather.p);
    gemm_nt(Wt.p, ldw, wc, Sn, Agather.p, Sn, T, Sn, Sn, false, F);
    k_col_flags<<<wc, 256, 0, s>>>(Wt.p, ldw, Sn, ispiv.p, colflag.p);
    if (F.small)
      k_tile_gauss<true><<<1, 1024, 0, s>>>(Wt.p, Sn, wc, ldw, c0, ispiv.p, pivrow.p, pivcol.p, Gc.p, tilepiv.p, ctl.p, colflag.p, cand, F, GjT, ldg, pividx);
    else
      k_tile_gauss<false><<<1, 1024, 0, s>>>(Wt.p, Sn, wc, ldw, c0, ispiv.p, pivrow.p, pivcol.p, Gc.p, tilepiv.p, ctl.p, colflag.p, cand, F, GjT, ldg, pividx);
    apply_T();
    CK(cudaGetLastError());
    g_launches += 5;
    h = fetch(ctl.p);
    if (h.found >= PB / 2)
      w = 32;
    else if (h.found == 0)
      w = std::min(WMAX, w * 4);
    else
      w = std::min(WMAX, std::max(32, (int)((double)PB * h.consumed / h.found)));
  }
  const int npiv = h.npiv;
  return npiv;
}

__global__ void k_compact_idx(const int *__restrict__ flag, const long long *__restrict__ pos, int n, int *__restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n && flag[c]) out[pos[c]] = c;
}
static void k_compact_flags_i32(const int *flag, const long long *pos, int n, int *out) {
  if (n) k_compact_idx<<<cdiv(n, 256), 256, 0, stream()>>>(flag, pos, n, out);
}
__global__ void k_gather_int(const int *__restrict__ src, const int *__restrict__ idx, int n, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}
// ranks that do not materialise the factor still keep U.n / qinv consistent: empty rows
__global__ void k_register_pivots_only(const int *__restrict__ pivcol, const int *__restrict__ q, int rr, int urow0, long long unz,
                                       long long *__restrict__ Up, int *__restrict__ Uqinv) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= rr) return;
  Uqinv[q[pivcol[s]]] = urow0 + s;
  Up[urow0 + s + 1] = unz;
}

// Dense tail.  With N ranks the remaining rows are dealt to the ranks panel by panel
// (block-cyclic, panel_owner): every rank builds and keeps only ITS rows of the dense Schur
// complement; the owner of panel b factors it and broadcasts the reduced rows R_b (NCCL over
// NVLink); every rank then updates its own later rows with the tensor-core GEMM.  Rank 0
// materialises the rows of U; the other ranks only track the pivots.
void echelonize_dense_device(const DCsr &A, const int *rows, int nrows, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size,
                             const TailOpts &opts) {
  cudaStream_t s = stream();
  if (block_size <= 0) block_size = 1000;
  if (nrows == 0 || A.m == U.n) return;
  const Dist &dd = dist();
  const int me = opts.lsink ? 0 : dd.rank, NR = opts.lsink ? 1 : dd.nranks;
  double t0 = spasm_wtime();
  // ---- my rows
  DBuf<int> rows_local;
  int n_local = nrows;
  const int *my_rows = rows;
  if (NR > 1) {
    std::vector<int> pos = local_positions(nrows, block_size, NR, me);
    n_local = (int)pos.size();
    DBuf<int> dpos(std::max(n_local, 1));
    rows_local.alloc(std::max(n_local, 1));
    if (n_local) {
      dpos.upload(pos.data(), n_local);
      k_gather_int<<<cdiv(n_local, 256), 256, 0, s>>>(rows, dpos.p, n_local, rows_local.p);
      sync();
    }
    my_rows = rows_local.p;
  }
  DenseSchur D;
  if (opts.lsink) {
    // with L: the multipliers on the structural pivots are part of the result (as in spasm_schur)
    build_dense_schur_raw(A.p.p, A.j.p, A.x.p, A.m, my_rows, n_local, U, Uqinv.p, F, D, true);
    if (U.n > 0 && n_local > 0) {
      DBuf<int> iota(n_local), lcnt(n_local), oj;
      DBuf<unsigned long long> loffs(n_local);
      DBuf<uint32_t> ox;
      k_iota2<<<cdiv(n_local, 256), 256, 0, s>>>(iota.p, n_local);
      dense_rows_to_sparse(D.Vp.p, D.ldv, U.n, nullptr, n_local, iota.p, 0, lcnt.p, loffs.p, 0ULL, oj, ox);
      opts.lsink->rows(0, n_local, 0, lcnt.p, loffs.p, oj.p, ox.p);
    }
    D.Vp = DBuf<uint32_t>();
  } else
    build_dense_schur(A, my_rows, n_local, U, Uqinv.p, F, D);
  logf("[echelonize/dense] dense schur complement %d x %d built in %.2fs (%d levels)%s\n", nrows, D.Sm0, spasm_wtime() - t0, D.levels,
       NR > 1 ? " [sharded]" : (opts.lsink ? " (with L)" : ""));
  dense_tail_core(D, nrows, n_local, A.m, U, Uqinv, F, block_size, opts);
}

// the blocked elimination itself, on a dense Schur complement that is already in HBM (D holds MY rows)
void dense_tail_core(DenseSchur &D, int nrows, int n_local, int m_total, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size,
                     const TailOpts &opts) {
  cudaStream_t s = stream();
  const Dist &dd = dist();
  LSink *const ls = opts.lsink;  // with-L mode: row echelon form + multipliers; every rank computes the whole tail
  const int me = ls ? 0 : dd.rank, NR = ls ? 1 : dd.nranks;
  // who materialises the rows of U of panel b: rank 0 (complete factor on rank 0) or the panel's owner (sharded factor)
  auto emits = [&](long long b) { return NR == 1 || (dd.shard_factor ? panel_owner(b, NR) == me : me == 0); };
  const bool emit_rows = NR == 1 || dd.shard_factor || me == 0;  // this rank materialises at least some panels
  const int Sm0 = D.Sm0;
  const long long ld = D.ld;
  // SPASM_B200_PROFILE=1: host clock around synchronised phases (perturbs multi-rank runs: every tick is a barrier for
  // the stream); =2: CUDA events recorded on the stream and read once at the end (device time per phase, no stalls)
  const char *prof_env = getenv("SPASM_B200_PROFILE");
  const bool prof = prof_env != nullptr;
  const bool prof_ev = prof && atoi(prof_env) == 2;
  double tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // panel, R gemm, emit, gather/transposes, trailing gemm, blocks, broadcast, handover (wait for the second stream)
  std::vector<std::pair<int, cudaEvent_t>> evs;  // (slot that ENDS at this event, event); slot -1 = start marker
  auto mark = [&](int slot) {
    cudaEvent_t ev;
    CK(cudaEventCreate(&ev));
    CK(cudaEventRecord(ev, s));
    evs.push_back({slot, ev});
  };
  auto tick = [&](int slot, double &t1) {
    if (!prof) return;
    if (prof_ev) {
      mark(slot);
      return;
    }
    sync();
    double now = spasm_wtime();
    tp[slot] += now - t1;
    t1 = now;
  };
  if (emit_rows) {
    // upper bound of what the dense rows add to U (every block full rank): reserve once, no regrowth copies
    long long ub = 0, left = Sm0;
    for (long long k0 = 0; k0 < nrows && left > 0; k0 += block_size) {
      long long sn = std::min<long long>(block_size, nrows - k0), rr = std::min(sn, left);
      if (emits(k0 / block_size)) ub += rr * left;
      left -= rr;
    }
    if ((size_t)ub * 8 < dev_free_bytes() / 2) csr_reserve(U, U.nnz + ub, U.n + std::min(nrows, Sm0));
  } else {
    csr_reserve(U, U.nnz, U.n + std::min(nrows, Sm0));
  }
  HostSink *sink = emit_rows ? g_sink : nullptr;
  if (sink) {
    // nothing on the device reads U's rows any more (the tail only appends): stream them out now
    sink->ensure((long long)U.j.n);
    convert_to_balanced(U.x.p + sink->submitted, (int *)U.x.p + sink->submitted, U.nnz - sink->submitted, F);
    sink->submit(U.j.p, (const int *)U.x.p, sink->submitted, U.nnz - sink->submitted);
  }
  const int Bmax = std::min(block_size, nrows);
  DBuf<uint32_t> T((size_t)Bmax * Bmax), Tsel((size_t)Bmax * Bmax), R((size_t)Bmax * Sm0), Rt, Pt;
  DBuf<int> ispiv, pivrow, ident(Bmax), hdr(1), cflag(Sm0 + 1), cand(std::max(Sm0, 1));
  DBuf<long long> cpos(Sm0 + 1);
  DBuf<unsigned char> colpiv(std::max(Sm0, 1));
  colpiv.zero();
  k_iota2<<<cdiv(Bmax, 256), 256, 0, s>>>(ident.p, Bmax);
  const long long nb = ((long long)nrows + block_size - 1) / block_size;
  long long lb = 0;  // my panels already factored
  // ---- deferred trailing updates.  A panel's update has K = its rank (<= block_size), and at K = 1000 the
  // tensor-core kernel spends 1/3 of its time in the read-modify-write epilogue of C.  So only the NEAR rows
  // (my next `group` panels) are updated at once; for the FAR rows [fe, n_local) the factors are appended to
  //     Rt_acc (Sm0 x Kacc)  and  Pt_acc (n_local x Kacc)
  // and applied in ONE product of depth Kacc <= kcap when the near rows are used up (or kcap is reached).
  // The multipliers of a new panel on far rows are read from columns that still miss the pending updates and
  // are corrected first:  Pt_b[k][s] = Dt[pc_b[s]][k] - sum_u Pt_acc[k][u] . Rt_acc[pc_b[s]][u]  (exact mod p,
  // so the result is bit-identical to the eager order).
  const int B16 = (Bmax + 15) / 16 * 16;
  int kcap = 4096;
  if (const char *e = getenv("SPASM_B200_LAZY_K")) kcap = atoi(e);
  const TailPlan plan = plan_tail(Sm0, n_local, block_size, Bmax, NR, kcap, gemm_max_k(F), dev_free_bytes());
  const bool lazy = plan.lazy;
  // my panels per flush: the factors of group * NR panels (mine and the other ranks') accumulate between two flushes,
  // and a flush happens exactly when my near rows are used up — the moment the next `group` panels change hands
  // (rank r's FIRST interval is r panels longer — the ranks' groups are staggered by one panel each — hence the NR - 1)
  const int group = plan.group;
  const int kdepth = plan.kdepth;
  const long long LDK = plan.LDK;
  // ---- look-ahead on two streams.  Everything the NEXT panel waits for — this panel's factorisation, its
  // broadcast, the update of the near rows — stays on the main stream (A, high priority); the far rows are only
  // touched by the second stream (B): gathering a panel's multipliers there, correcting them, and the deep flushes.
  // A hands rows to B never; B hands the next `group` panels to A at a flush (one event).  Two sets of accumulators:
  // while B applies one, A already fills the other.  B's tensor-core launches leave some SMs to A's (small, latency
  // bound) kernels, which would otherwise queue behind a persistent kernel that owns every SM's shared memory.
  // Default: two streams when the tail is spread over several ranks (the second stream then has 1/N of the work and the
  // serial chain panel -> broadcast -> near rows is what bounds the step: 4 GPUs 5.1 -> 3.6 s).  On ONE GPU the second
  // stream carries as much work as the first, the two share the SMs, and the measured gain at 200 000 rows is within the
  // run-to-run spread (+-0.4 s under the power cap) — one stream stays the default there; SPASM_B200_TWO_STREAMS=1 /
  // SPASM_B200_ONE_STREAM=1 override either way.
  const bool want_two = getenv("SPASM_B200_TWO_STREAMS") != nullptr || (NR > 1 && getenv("SPASM_B200_ONE_STREAM") == nullptr);
  const bool two_streams = lazy && ls == nullptr && want_two;
  const int nsets = two_streams ? 2 : 1;
  cudaStream_t sA = s, sB = two_streams ? aux_stream() : s;
  int aux_ctas = sm_count() - (NR > 1 ? 48 : 32);  // (with several ranks the second stream has 1/N of the work: the critical path gets more room)
  if (const char *e = getenv("SPASM_B200_AUX_SMS")) aux_ctas = atoi(e);
  aux_ctas = std::max(8, std::min(aux_ctas, sm_count()));
  DBuf<uint32_t> Rt_accs[2], Pt_accs[2], Rsel, wire;
  DBuf<int> cand_snap[2];
  if (lazy) {
    for (int i = 0; i < nsets; i++)
      Rt_accs[i].alloc((size_t)Sm0 * LDK), Pt_accs[i].alloc((size_t)n_local * LDK), cand_snap[i].alloc(std::max(Sm0, 1));
    Rsel.alloc((size_t)B16 * LDK);
  }
  // per-panel pivot columns: the second stream reads a panel's list after the main stream has moved on
  DBuf<int> pc_hist((size_t)std::min(nrows, Sm0) + Bmax + 1);
  std::vector<cudaEvent_t> ev_pool;
  auto new_event = [&]() {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ev_pool.push_back(e);
    return e;
  };
  // `to` will not start anything enqueued after this call before `from` has finished everything enqueued before it
  auto order = [&](cudaStream_t from, cudaStream_t to) {
    if (from == to) return;
    cudaEvent_t e = new_event();
    CK(cudaEventRecord(e, from));
    CK(cudaStreamWaitEvent(to, e, 0));
  };
  struct AuxGuard {  // an exception must not free buffers the second stream still works on
    cudaStream_t b;
    std::vector<cudaEvent_t> *pool;
    ~AuxGuard() {
      if (b) cudaStreamSynchronize(b);
      for (cudaEvent_t e : *pool) cudaEventDestroy(e);
    }
  } aux_guard{two_streams ? sB : nullptr, &ev_pool};
  cudaEvent_t ev_flush[2] = {nullptr, nullptr};  // end of the last flush that read accumulator set i
  int cur = 0;                                           // accumulator set being filled
  int Kacc = 0;                                          // depth of the pending product
  long long gend = (long long)group * block_size;        // my local rows [.., gend) are always up to date
  // live columns: cand[0:nlive) = the columns that are not pivots of a panel factored so far, increasing.  Pivoted
  // columns are dead (zero on every remaining row): the panel factorisation scans, and the trailing products
  // update, the live columns only — the work shrinks with the elimination (n^3/3 instead of n^3/2).
  int nlive = Sm0;
  auto refresh_live = [&]() {
    k_not_flag<<<cdiv(Sm0 + 1, 256), 256, 0, s>>>(colpiv.p, Sm0, cflag.p);
    exclusive_scan_i32_to_i64(cflag.p, cpos.p, Sm0 + 1);
    k_compact_flags_i32(cflag.p, cpos.p, Sm0, cand.p);
  };
  refresh_live();
  // apply the pending product to the far rows [min(gend, n_local), n_local); the rows below next_gend first: they are
  // the ones the main stream is about to need (it waits for that part only)
  auto flush_far = [&](long long next_gend) {
    const long long fe = std::min<long long>(gend, n_local);
    bool flushed = false;
    if (Kacc > 0 && fe < n_local && nlive > 0) {
      uint32_t *Rt_acc = Rt_accs[cur].p, *Pt_acc = Pt_accs[cur].p;
      const int nl = nlive;
      const int *cmap = cand.p;
      if (two_streams) {  // A refreshes cand with every panel: B works on a snapshot
        CK(cudaMemcpyAsync(cand_snap[cur].p, cand.p, (size_t)nl * sizeof(int), cudaMemcpyDeviceToDevice, sA));
        cmap = cand_snap[cur].p;
      }
      order(sA, sB);
      const long long mid = std::min<long long>(std::max(next_gend, fe), n_local);
      {
        StreamScope on_b(sB);
        GemmCtaLimit lim(two_streams ? aux_ctas : 0);
        if (mid > fe) {
          gemm_nt(D.Dt.p + fe, ld, nl, (int)(mid - fe), Rt_acc, LDK, Pt_acc + fe * LDK, LDK, Kacc, true, F, cmap);
          order(sB, sA);  // rows [fe, mid) change hands: A goes on as soon as THEY are up to date
        }
        if (n_local > mid) gemm_nt(D.Dt.p + mid, ld, nl, (int)(n_local - mid), Rt_acc, LDK, Pt_acc + mid * LDK, LDK, Kacc, true, F, cmap);
      }
      flushed = true;
      g_tail_stats[0] += 1, g_tail_stats[2] += (n_local - fe) * (long long)Kacc;
      if (two_streams) {
        if (!ev_flush[cur]) ev_flush[cur] = new_event();
        CK(cudaEventRecord(ev_flush[cur], sB));
        cur ^= 1;  // A fills the other set meanwhile — once the flush that read it (two flushes ago) is over
        if (ev_flush[cur]) CK(cudaStreamWaitEvent(sA, ev_flush[cur], 0));
      }
    }
    if (!flushed && next_gend > gend) order(sB, sA);  // rows change hands without a flush: B must be done with them
    Kacc = 0;
  };
  long long pc_off = 0;
  // with-L mode buffers
  DBuf<uint32_t> GjT, UM, PanelPt, Ub, RtL, LT;
  DBuf<int> pividx, iota_rows, lcnt;
  DBuf<unsigned long long> loffs;
  const long long ldlt = ((long long)std::max(Bmax, n_local) + 63) / 64 * 64;
  if (ls) {
    GjT.alloc((size_t)Bmax * Bmax), UM.alloc((size_t)Bmax * Bmax), PanelPt.alloc((size_t)Bmax * B16), Ub.alloc((size_t)Bmax * Sm0);
    RtL.alloc((size_t)Sm0 * B16), LT.alloc((size_t)B16 * ldlt), pividx.alloc(Bmax), iota_rows.alloc(std::max(n_local, Bmax));
    lcnt.alloc(std::max(n_local, Bmax)), loffs.alloc(std::max(n_local, Bmax));
    k_iota2<<<cdiv(std::max(n_local, Bmax), 256), 256, 0, s>>>(iota_rows.p, std::max(n_local, Bmax));
  }
  // multipliers of the tail rows [row0, row0 + nr) on the rr pivots of this panel:  LT[s][k] = sum_s' (I - G)^T[s][s'] . cols[k][s']
  auto emit_L = [&](const uint32_t *cols, long long ldcols, int row0, int nr, int rr_, int ubase) {
    if (nr <= 0 || rr_ <= 0) return;
    gemm_nt(LT.p, ldlt, rr_, nr, GjT.p, Bmax, cols, ldcols, rr_, false, F);
    DBuf<int> oj;
    DBuf<uint32_t> ox;
    dense_rows_to_sparse(LT.p, ldlt, rr_, nullptr, nr, iota_rows.p, 0, lcnt.p, loffs.p, 0ULL, oj, ox);
    ls->rows(row0, nr, ubase, lcnt.p, loffs.p, oj.p, ox.p);
  };
  for (long long b = 0; b < nb; b++) {
    const long long kg = b * block_size;
    const int Sn = (int)std::min<long long>(block_size, nrows - kg);
    const int owner = panel_owner(b, NR);
    if (lazy && owner == me && lb * block_size >= gend) {  // the near rows are used up: bring the far rows up to date
      double tf = spasm_wtime();
      if (prof_ev) mark(-1);
      flush_far(lb * block_size + (long long)group * block_size);
      gend = lb * block_size + (long long)group * block_size;
      if (prof_ev)
        mark(7);
      else if (prof) {
        sync();
        tp[7] += spasm_wtime() - tf;
      }
    }
    logf("[echelonize/dense] processing dense schur complement of dimension %lld x %d; block size=%d\n", (long long)nrows - kg, m_total - U.n,
         block_size);
    double t1 = spasm_wtime();
    if (prof_ev) mark(-1);
    int rr = 0;
    int *pc_b = pc_hist.p + pc_off;  // this panel's pivot columns (kept: the second stream reads them later)
    if (owner == me) {
      const long long k0 = lb * block_size;
      DBuf<int> pc_tmp;
      // candidate columns of this panel: everything that is not a pivot of an earlier panel
      if (ls) {
        CK(cudaMemsetAsync(GjT.p, 0, (size_t)Bmax * Bmax * sizeof(uint32_t), s));
        k_fill_int<<<cdiv(Bmax, 256), 256, 0, s>>>(pividx.p, Bmax, -1);
      }
      rr = panel_factor(D.Dt.p, ld, cand.p, nlive, k0, Sn, T.p, ispiv, pivrow, pc_tmp, F, ls ? GjT.p : nullptr, Bmax, ls ? pividx.p : nullptr);
      if (rr > 0) CK(cudaMemcpyAsync(pc_b, pc_tmp.p, (size_t)rr * sizeof(int), cudaMemcpyDeviceToDevice, s));
      tick(0, t1);
      if (rr > 0) {
        // reduced rows  R[s][c] = sum_t T[pivrow[s]][t] * Dt[c][k0+t]
        k_gather_T_rows<<<dim3(cdiv(Sn, 256), rr), 256, 0, s>>>(T.p, Sn, pivrow.p, nullptr, rr, Tsel.p);
        gemm_nt(R.p, Sm0, rr, Sm0, Tsel.p, Sn, D.Dt.p + k0, ld, Sn, false, F);
      }
      if (ls && rr > 0) {
        // ---- with L: U_M = (I - G)^-1, then GjT <- (I - G)^T; the panel's own multipliers; the row echelon rows Ub = U_M . R
        if (F.small)
          k_um_from_gT<true><<<cdiv(rr, 64), 64, 0, s>>>(GjT.p, Bmax, rr, UM.p, Bmax, F);
        else
          k_um_from_gT<false><<<cdiv(rr, 64), 64, 0, s>>>(GjT.p, Bmax, rr, UM.p, Bmax, F);
        k_make_uinvT<<<dim3(cdiv(rr, 256), rr), 256, 0, s>>>(GjT.p, Bmax, rr, F.p);
        k_gather_pivot_cols_T<<<dim3(cdiv(Sn, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(D.Dt.p, ld, pc_b, rr, k0, Sn, PanelPt.p, B16);
        emit_L(PanelPt.p, B16, (int)kg, Sn, rr, U.n);
        k_transpose_u32<<<dim3(cdiv(Sm0, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(R.p, Sm0, rr, Sm0, RtL.p, B16);
        gemm_nt(Ub.p, Sm0, rr, Sm0, UM.p, Bmax, RtL.p, B16, rr, false, F);
        std::vector<int> hpr(rr);
        pivrow.download(hpr.data(), rr);
        sync();
        ls->pivots(U.n, hpr.data(), rr, (int)kg);
      }
      tick(1, t1);
      lb++;
    }
    tp[5] += 1;
    if (NR > 1) {
      if (owner == me) CK(cudaMemcpyAsync(hdr.p, &rr, sizeof(int), cudaMemcpyHostToDevice, s));
      dist_broadcast(hdr.p, sizeof(int), owner);
      rr = fetch(hdr.p);
      if (rr > 0) {
        dist_broadcast(pc_b, (size_t)rr * sizeof(int), owner);
        // R_b travels packed: the nlive columns that were live before this panel (all others are zero), u16 when p < 2^16
        const size_t esz = F.small ? 2 : 4;
        wire.alloc(((size_t)rr * nlive * esz + 3) / 4);
        const dim3 pg(cdiv(nlive, 256), rr);
        if (owner == me) {
          if (F.small)
            k_pack_R<unsigned short><<<pg, 256, 0, s>>>(R.p, Sm0, cand.p, nlive, rr, (unsigned short *)wire.p);
          else
            k_pack_R<uint32_t><<<pg, 256, 0, s>>>(R.p, Sm0, cand.p, nlive, rr, wire.p);
        }
        dist_broadcast(wire.p, (size_t)rr * nlive * esz, owner);
        if (owner != me) {
          CK(cudaMemsetAsync(R.p, 0, (size_t)rr * Sm0 * sizeof(uint32_t), s));
          if (F.small)
            k_unpack_R<unsigned short><<<pg, 256, 0, s>>>((const unsigned short *)wire.p, cand.p, nlive, rr, R.p, Sm0);
          else
            k_unpack_R<uint32_t><<<pg, 256, 0, s>>>(wire.p, cand.p, nlive, rr, R.p, Sm0);
        }
        CK(cudaGetLastError());
      }
      tick(6, t1);
    }
    if (rr > 0) {
      pc_off += rr;
      k_mark_cols<<<cdiv(rr, 256), 256, 0, s>>>(pc_b, rr, colpiv.p);
      refresh_live();
      nlive -= rr;
      if (emits(b)) {
        // append to U: (q0[pivcol[s]], 1) then the other nonzeros by increasing column
        DBuf<int> cnt(rr + 1);
        DBuf<long long> rpos(rr + 1);
        const uint32_t *Urows = ls ? Ub.p : R.p;  // with L: the rows as they were when they became pivots (row echelon form)
        k_count_rows<<<rr, 256, 0, s>>>(Urows, Sm0, Sm0, ident.p, rr, cnt.p);
        exclusive_scan_i32_to_i64(cnt.p, rpos.p, rr + 1);
        const long long add = fetch(rpos.p + rr);
        if (sink && (size_t)(U.nnz + add) > U.j.n) sink->wait_all();  // the arrays are about to move
        csr_reserve(U, U.nnz + add, U.n + rr);
        k_write_rows<<<rr, 256, 0, s>>>(Urows, Sm0, Sm0, ident.p, pc_b, D.q0.p, rpos.p, U.nnz, U.n, U.p.p, U.j.p, U.x.p, Uqinv.p);
        CK(cudaGetLastError());
        if (sink) {
          convert_to_balanced(U.x.p + U.nnz, (int *)U.x.p + U.nnz, add, F);
          sink->submit(U.j.p, (const int *)U.x.p, U.nnz, add);
        }
        U.nnz += add;
      } else {
        k_register_pivots_only<<<cdiv(rr, 256), 256, 0, s>>>(pc_b, D.q0.p, rr, U.n, U.nnz, U.p.p, Uqinv.p);
      }
      U.n += rr;
      g_launches += 6;
      tick(2, t1);
      // trailing update of MY later rows:  Dt[c][k] -= sum_s R[s][c] * Dt[pivcol[s]][k]
      const long long kb = lb * block_size;
      const int nk = (int)std::max<long long>(0, n_local - kb);
      if (nk > 0 && !lazy) {
        const long long ldk = ((long long)rr + 15) / 16 * 16;
        Rt.alloc((size_t)Sm0 * ldk);
        Pt.alloc((size_t)nk * ldk);
        k_transpose_u32<<<dim3(cdiv(Sm0, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(R.p, Sm0, rr, Sm0, Rt.p, ldk);
        k_gather_pivot_cols_T<<<dim3(cdiv(nk, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(D.Dt.p, ld, pc_b, rr, kb, nk, Pt.p, ldk);
        k_zero_pivot_cols<<<dim3(cdiv(nk, 256), rr), 256, 0, s>>>(D.Dt.p, ld, pc_b, rr, kb, nk);
        tick(3, t1);
        if (ls) emit_L(Pt.p, ldk, (int)kb, nk, rr, U.n - rr);
        if (nlive > 0) gemm_nt(D.Dt.p + kb, ld, nlive, nk, Rt.p, ldk, Pt.p, ldk, rr, true, F, cand.p);
        tick(4, t1);
      } else if (nk > 0) {
        const int rr16 = (rr + 15) / 16 * 16;
        const long long fe = std::min<long long>(std::max(gend, kb), n_local);  // near rows [kb, fe) (stream A), far rows [fe, n_local) (stream B)
        const int nnear = (int)(fe - kb), nfar = (int)(n_local - fe);
        uint32_t *Rt_acc = Rt_accs[cur].p, *Pt_acc = Pt_accs[cur].p;
        uint32_t *Rt_b = Rt_acc + Kacc, *Pt_b = Pt_acc + kb * LDK + Kacc;
        k_transpose_u32<<<dim3(cdiv(Sm0, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(R.p, Sm0, rr, Sm0, Rt_b, LDK);
        if (rr16 > rr) CK(cudaMemset2DAsync(Rt_b + rr, (size_t)LDK * 4, 0, (size_t)(rr16 - rr) * 4, Sm0, s));  // zero pad columns: the pad of Pt may hold anything
        if (nnear > 0) {
          k_gather_pivot_cols_T<<<dim3(cdiv(nnear, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(D.Dt.p, ld, pc_b, rr, kb, nnear, Pt_b, LDK);
          k_zero_pivot_cols<<<dim3(cdiv(nnear, 256), rr), 256, 0, s>>>(D.Dt.p, ld, pc_b, rr, kb, nnear);
          if (rr16 > rr) CK(cudaMemset2DAsync(Pt_b + rr, (size_t)LDK * 4, 0, (size_t)(rr16 - rr) * 4, nnear, s));  // and the pad of Pt: no garbage x 0
        }
        if (nfar > 0) {
          // the far rows of these pivot columns (they still miss the pending updates: corrected with what is accumulated)
          order(sA, sB);
          StreamScope on_b(sB);
          GemmCtaLimit lim(two_streams ? aux_ctas : 0);
          uint32_t *Pt_f = Pt_acc + fe * LDK + Kacc;
          k_gather_pivot_cols_T<<<dim3(cdiv(nfar, 32), cdiv(rr, 32)), dim3(32, 8), 0, sB>>>(D.Dt.p, ld, pc_b, rr, fe, nfar, Pt_f, LDK);
          k_zero_pivot_cols<<<dim3(cdiv(nfar, 256), rr), 256, 0, sB>>>(D.Dt.p, ld, pc_b, rr, fe, nfar);
          if (rr16 > rr) CK(cudaMemset2DAsync(Pt_f + rr, (size_t)LDK * 4, 0, (size_t)(rr16 - rr) * 4, nfar, sB));
          if (Kacc > 0) {
            k_gather_rows_ld<<<dim3(cdiv(Kacc, 256), rr), 256, 0, sB>>>(Rt_acc, LDK, pc_b, rr, Kacc, Rsel.p, LDK);
            gemm_nt(Pt_f, LDK, nfar, rr, Pt_acc + fe * LDK, LDK, Rsel.p, LDK, Kacc, true, F);
            g_launches += 1;
            g_tail_stats[1] += 1;
          }
        }
        tick(3, t1);
        if (ls) emit_L(Pt_b, LDK, (int)kb, nk, rr, U.n - rr);  // (one stream in this mode: the far rows' multipliers are corrected by now)
        if (nnear > 0 && nlive > 0) gemm_nt(D.Dt.p + kb, ld, nlive, nnear, Rt_b, LDK, Pt_b, LDK, rr, true, F, cand.p), g_tail_stats[3] += 1;
        Kacc += rr16;
        if (Kacc > kdepth || fe >= n_local) flush_far(gend);  // (accumulators full before my near rows are used up: rare)
        tick(4, t1);
      }
      g_launches += 2;
    }
    logf("[echelonize/dense] block done: %d new pivots, rank %d\n", rr, U.n);
    if (U.n == m_total) break;
    // ---- a block far below full rank hands the rows still to come to the low-rank mode
    const long long rows_left = (long long)nrows - (kg + Sn);
    if (opts.tall_skinny && ls == nullptr && rows_left > 0 && (double)rr < opts.low_rank_ratio * (double)Sn) {
      logf("[echelonize/dense] %d pivots in a block of %d rows: switching to low-rank mode\n", rr, Sn);
      if (lazy) flush_far(n_local);  // every remaining row up to date (the main stream waits for all of it)
      const long long kb = lb * block_size;  // my remaining rows are [kb, n_local)
      const int nloc = (int)std::max<long long>(0, n_local - kb);
      LowRankShard sh;
      DBuf<int> gidx, locg;
      if (NR > 1) {
        // global index (among the rows_left remaining rows) of my local rows, and the inverse map
        std::vector<int> pos = local_positions(nrows, block_size, NR, me), g(std::max(nloc, 1)), inv((size_t)rows_left, -1);
        for (int i = 0; i < nloc; i++) {
          g[i] = (int)(pos[kb + i] - (kg + Sn));
          inv[g[i]] = i;
        }
        gidx.alloc(std::max(nloc, 1)), locg.alloc((size_t)rows_left);
        if (nloc) gidx.upload(g.data(), nloc);
        locg.upload(inv.data(), (size_t)rows_left);
        sync();
        sh.gidx = gidx.p, sh.loc_of_glob = locg.p, sh.reduce = true, sh.panel_base = b + 1;
      }
      g_tail_stats[4] += 1;
      lowrank_core(D, kb, nloc, (int)rows_left, sh, colpiv.p, U, Uqinv, F, block_size, opts.start_weight, m_total);
      break;
    }
  }
  order(sB, sA);  // whatever the second stream still does is part of this call
  if (prof_ev) {
    sync();
    for (size_t i = 1; i < evs.size(); i++) {
      if (evs[i].first >= 0) {
        float ms = 0;
        cudaEventElapsedTime(&ms, evs[i - 1].second, evs[i].second);
        tp[evs[i].first] += ms * 1e-3;
      }
    }
    for (auto &e_ : evs) cudaEventDestroy(e_.second);
  }
  if (prof) {
    extern void alloc_diagnostics(long long out[4]);
    long long ad[4];
    alloc_diagnostics(ad);
    size_t fr = 0, to = 0;
    cudaMemGetInfo(&fr, &to);
    fprintf(stderr, "[dense] rank %d allocator: %lld retries so far, %.1f GB in the block cache, pool %.1f GB reserved / %.1f GB in use, %.1f GB free on the device\n",
            me, ad[0], ad[1] / 1e9, ad[2] / 1e9, ad[3] / 1e9, fr / 1e9);
  }
  if (prof)
    fprintf(stderr, "[dense] rank %d/%d blocks=%d panel=%.3fs Rgemm=%.3fs emit=%.3fs gather=%.3fs near=%.3fs bcast=%.3fs handover=%.3fs (%s)\n", me, NR,
            (int)tp[5], tp[0], tp[1], tp[2], tp[3], tp[4], tp[6], tp[7], two_streams ? "two streams" : "one stream");
}

}  // namespace sb

using namespace sb;

namespace sb {

// ------------------------------------------------------------------ low-rank / tall-and-skinny mode
// (prototype spasm_schur_dense_randomized, src/SpaSM.jl:767-769).  Same counter-based
// random numbers as the oracle.  The dense Schur complement D of ALL remaining rows is built once;
// a block of random combinations is then simply  Blk = D^T-layout GEMM  (Sm0 x B) = Dt . Coef^T, which is
// what eliminating the combined rows against U would give (elimination is linear).  The block is
// factored like a panel; the trailing update touches every row of D (rows are never consumed).
__host__ __device__ inline unsigned long long lowrank_hash(unsigned long long blk, unsigned long long t, unsigned long long k) {
  unsigned long long z = 0x5a5a5a5a2e6306e0ULL ^ (blk * 0x9e3779b97f4a7c15ULL) ^ (t * 0xbf58476d1ce4e5b9ULL + 0x1234567ULL) ^
                         (k * 0x94d049bb133111ebULL + 0x89abcdefULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
// gidx (optional): global index, in the remaining-row list, of local row k (rows spread over several ranks)
__global__ void k_coef_full(uint32_t *__restrict__ coef, long long ldc, int B, int n, unsigned long long blk, uint32_t p, int negate,
                            const int *__restrict__ gidx) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (k >= n || t >= B) return;
  uint32_t c = (uint32_t)(lowrank_hash(blk, t, (unsigned long long)(gidx ? gidx[k] : k)) % p);
  coef[(long long)t * ldc + k] = (negate && c) ? p - c : c;
}
// n = number of remaining rows over all ranks; loc_of_glob (optional): local row of global row k, -1 = another rank's
__global__ void k_coef_sparse(uint32_t *__restrict__ coef, long long ldc, int B, int n, int w, unsigned long long blk, Fp F,
                              const int *__restrict__ loc_of_glob) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // one combination per thread: its w terms may collide
  if (t >= B) return;
  for (int s = 0; s < w; s++) {
    int k = (int)(lowrank_hash(blk, t, 2ULL * s) % (unsigned long long)n);
    const uint32_t c = (uint32_t)(1 + lowrank_hash(blk, t, 2ULL * s + 1) % (unsigned long long)(F.p - 1));
    if (loc_of_glob) k = loc_of_glob[k];
    if (k < 0) continue;
    uint32_t *q = coef + (long long)t * ldc + k;
    *q = addmod(*q, c, F);
  }
}
__global__ void k_negate_rows(uint32_t *a, long long ld, int cols, uint32_t p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cols) return;
  uint32_t *q = a + (long long)blockIdx.y * ld + k;
  if (*q) *q = p - *q;
}
__global__ void k_widen_u64(const uint32_t *__restrict__ in, unsigned long long *__restrict__ out, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_narrow_mod(const unsigned long long *__restrict__ in, uint32_t *__restrict__ out, long long n, unsigned long long p) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)(in[i] % p);
}

// The loop of the low-rank mode on rows [row0, row0 + nloc) of a dense Schur complement that is up to date with
// respect to U (dead columns flagged in colpiv and zero on these rows).  nglob = remaining rows over all ranks.
void lowrank_core(DenseSchur &D, long long row0, int nloc, int nglob, const LowRankShard &sh, unsigned char *colpiv, DCsr &U,
                  DBuf<int> &Uqinv, const Fp &F, int block_size, double start_weight, int m_total) {
  cudaStream_t s = stream();
  if (block_size <= 0) block_size = 1000;
  if (nglob == 0 || m_total == U.n) return;
  const Dist &dd = dist();
  const int B = block_size;
  int w = (start_weight >= 1) ? (int)start_weight : nglob;
  if (w > nglob) w = nglob;
  const int Sm0 = D.Sm0;
  const long long ld = D.ld;
  logf("[echelonize/low-rank] %d rows, %d columns left, block size %d, starting weight %d\n", nglob, m_total - U.n, B, w);
  const long long ldb = ((long long)B + 63) / 64 * 64;
  const long long ldc = ((long long)std::max(nloc, 1) + 63) / 64 * 64;
  DBuf<uint32_t> Coef((size_t)B * ldc), Blk((size_t)Sm0 * ldb), T((size_t)B * B), Tsel((size_t)B * B), R((size_t)B * Sm0), Rt, Pt;
  DBuf<unsigned long long> wide;
  DBuf<int> ispiv, pivrow, pc_tmp, ident(B), cflag(Sm0 + 1), cand(std::max(Sm0, 1));
  DBuf<long long> cpos(Sm0 + 1);
  k_iota2<<<cdiv(B, 256), 256, 0, s>>>(ident.p, B);
  const int KCH = gemm_max_k(F);  // depth one tensor-core launch takes (int32 accumulator bound of the limb kernel)
  for (unsigned long long blk = 0;; blk++) {
    if (U.n == m_total) break;
    // ---- coefficients and the combined block  Blk[c][t] = sum_k Dt[c][k] * Coef[t][k]
    if (nloc > 0) {
      if (w >= nglob) {
        k_coef_full<<<dim3(cdiv(nloc, 256), B), 256, 0, s>>>(Coef.p, ldc, B, nloc, blk, F.p, 0, sh.gidx);
      } else {
        Coef.zero();
        k_coef_sparse<<<cdiv(B, 128), 128, 0, s>>>(Coef.p, ldc, B, nglob, w, blk, F, sh.loc_of_glob);
      }
      for (int kc = 0; kc < nloc; kc += KCH) {
        const int kk = std::min(KCH, nloc - kc);
        if (kc > 0) {  // C += A.B as C -= A.(-B)
          k_negate_rows<<<dim3(cdiv(kk, 256), B), 256, 0, s>>>(Coef.p + kc, ldc, kk, F.p);
        }
        gemm_nt(Blk.p, ldb, Sm0, B, D.Dt.p + row0 + kc, ld, Coef.p + kc, ldc, kk, kc > 0, F);
      }
    } else
      Blk.zero();
    if (sh.reduce && dd.nranks > 1) {  // my rows' part of every combination -> the combination
      const long long cnt = (long long)Sm0 * ldb;
      wide.alloc((size_t)cnt);
      k_widen_u64<<<cdiv(cnt, 256), 256, 0, s>>>(Blk.p, wide.p, cnt);
      dist_allreduce_sum_u64(wide.p, (size_t)cnt);
      k_narrow_mod<<<cdiv(cnt, 256), 256, 0, s>>>(wide.p, Blk.p, cnt, (unsigned long long)F.p);
    }
    // ---- factor the block like a panel
    k_not_flag<<<cdiv(Sm0 + 1, 256), 256, 0, s>>>(colpiv, Sm0, cflag.p);
    exclusive_scan_i32_to_i64(cflag.p, cpos.p, Sm0 + 1);
    k_compact_flags_i32(cflag.p, cpos.p, Sm0, cand.p);
    const int ncand = (int)fetch(cpos.p + Sm0);
    const int rr = panel_factor(Blk.p, ldb, cand.p, ncand, 0, B, T.p, ispiv, pivrow, pc_tmp, F);
    if (rr == 0) {
      if (w >= nglob) break;
      w = (2 * w < nglob) ? 2 * w : nglob;
      continue;
    }
    k_gather_T_rows<<<dim3(cdiv(B, 256), rr), 256, 0, s>>>(T.p, B, pivrow.p, nullptr, rr, Tsel.p);
    gemm_nt(R.p, Sm0, rr, Sm0, Tsel.p, B, Blk.p, ldb, B, false, F);
    k_mark_cols<<<cdiv(rr, 256), 256, 0, s>>>(pc_tmp.p, rr, colpiv);
    // who materialises these rows: everybody when every rank holds all rows, else as the dense panels do
    const bool emit = !sh.reduce || dd.nranks == 1 || (dd.shard_factor ? panel_owner(sh.panel_base + (long long)blk, dd.nranks) == dd.rank : dd.rank == 0);
    if (emit) {
      DBuf<int> cnt(rr + 1);
      DBuf<long long> rpos(rr + 1);
      k_count_rows<<<rr, 256, 0, s>>>(R.p, Sm0, Sm0, ident.p, rr, cnt.p);
      exclusive_scan_i32_to_i64(cnt.p, rpos.p, rr + 1);
      const long long add = fetch(rpos.p + rr);
      if (g_sink && (size_t)(U.nnz + add) > U.j.n) g_sink->wait_all();  // the arrays are about to move
      csr_reserve(U, U.nnz + add, U.n + rr);
      k_write_rows<<<rr, 256, 0, s>>>(R.p, Sm0, Sm0, ident.p, pc_tmp.p, D.q0.p, rpos.p, U.nnz, U.n, U.p.p, U.j.p, U.x.p, Uqinv.p);
      CK(cudaGetLastError());
      U.nnz += add;
    } else {
      csr_reserve(U, U.nnz, U.n + rr);
      k_register_pivots_only<<<cdiv(rr, 256), 256, 0, s>>>(pc_tmp.p, D.q0.p, rr, U.n, U.nnz, U.p.p, Uqinv.p);
    }
    U.n += rr;
    // ---- trailing update of EVERY row:  Dt[c][k] -= sum_s R[s][c] * Dt[pivcol[s]][k]
    if (nloc > 0) {
      const long long ldk = ((long long)rr + 15) / 16 * 16;
      Rt.alloc((size_t)Sm0 * ldk);
      Pt.alloc((size_t)nloc * ldk);
      k_transpose_u32<<<dim3(cdiv(Sm0, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(R.p, Sm0, rr, Sm0, Rt.p, ldk);
      k_gather_pivot_cols_T<<<dim3(cdiv(nloc, 32), cdiv(rr, 32)), dim3(32, 8), 0, s>>>(D.Dt.p, ld, pc_tmp.p, rr, row0, nloc, Pt.p, ldk);
      gemm_nt(D.Dt.p + row0, ld, Sm0, nloc, Rt.p, ldk, Pt.p, ldk, rr, true, F);
    }
    logf("[echelonize/low-rank] block %d: %d new pivots (weight %d), rank %d\n", (int)blk, rr, w, U.n);
  }
}

void echelonize_lowrank_device(const DCsr &A, const int *rows, int nrows, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size,
                               double start_weight) {
  if (nrows == 0 || A.m == U.n) return;
  DenseSchur D;
  build_dense_schur(A, rows, nrows, U, Uqinv.p, F, D);  // every rank holds all rows: no exchange in this mode
  DBuf<unsigned char> colpiv(std::max(D.Sm0, 1));
  colpiv.zero();
  lowrank_core(D, 0, nrows, nrows, LowRankShard(), colpiv.p, U, Uqinv, F, block_size, start_weight, A.m);
}

__global__ void k_fill_random(uint32_t *__restrict__ a, long long n, uint32_t p, unsigned long long seed) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + (unsigned long long)i * 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  z ^= z >> 31;
  a[i] = (uint32_t)(z % p);
}
}  // namespace sb

// BASELINE configs[3]: dense n x m Schur complement mod `prime`, generated on the device, through the
// blocked elimination of the dense tail (panel factorisation + tcgen05 trailing updates).
// Returns the rank; ms = CUDA-event time of the elimination (generation excluded).
static int dense_tail_bench_impl(long long prime, int n, int m, int planted_rank, int block_size, unsigned long long seed, double *ms);
extern "C" int spasm_b200_dense_tail_bench(long long prime, int n, int m, int block_size, unsigned long long seed, double *ms) {
  return dense_tail_bench_impl(prime, n, m, 0, block_size, seed, ms);
}
// same with a PLANTED rank r < min(n, m): the matrix is the product of a random n x r and a random r x m
// matrix (rank r unless a random r x r minor is singular, probability about 1/p), SURVEY.md configs[3] variant
extern "C" int spasm_b200_dense_tail_bench_planted(long long prime, int n, int m, int r, int block_size, unsigned long long seed,
                                                   double *ms) {
  return dense_tail_bench_impl(prime, n, m, r, block_size, seed, ms);
}
static int dense_tail_bench_impl(long long prime, int n, int m, int planted_rank, int block_size, unsigned long long seed, double *ms) {
  try {
    ApiCall api_scope_;
    Fp F = make_field(prime);
    DenseSchur D;
    D.n_rem = n, D.Sm0 = m, D.levels = 0;
    D.ld = ((long long)n + 63) / 64 * 64;
    D.Dt.alloc((size_t)m * D.ld);
    if (planted_rank > 0) {
      const long long ldr = ((long long)planted_rank + 15) / 16 * 16;
      DBuf<uint32_t> Yt((size_t)m * ldr), X((size_t)n * ldr);
      k_fill_random<<<cdiv((long long)m * ldr, 256), 256, 0, stream()>>>(Yt.p, (long long)m * ldr, F.p, seed ^ 0x1111);
      k_fill_random<<<cdiv((long long)n * ldr, 256), 256, 0, stream()>>>(X.p, (long long)n * ldr, F.p, seed ^ 0x2222);
      D.Dt.zero();
      for (int k0 = 0; k0 < planted_rank; k0 += 8192)  // Dt[c][k] = - sum_s Y[s][c] X[k][s], in K chunks the tensor-core path takes
        gemm_nt(D.Dt.p, D.ld, m, n, Yt.p + k0, ldr, X.p + k0, ldr, std::min(8192, planted_rank - k0), true, F);
    } else
      k_fill_random<<<cdiv((long long)m * D.ld, 256), 256, 0, stream()>>>(D.Dt.p, (long long)m * D.ld, F.p, seed);
    D.q0.alloc(m);
    k_iota2<<<cdiv(m, 256), 256, 0, stream()>>>(D.q0.p, m);
    DCsr U;
    U.n = 0, U.m = m, U.nnz = 0;
    U.p.alloc(n + 2);
    U.p.zero();
    U.j.alloc(16), U.x.alloc(16);
    DBuf<int> Uqinv(m);
    Uqinv.fill_ff();
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, stream()));
    dense_tail_core(D, n, n, m, U, Uqinv, F, block_size, TailOpts());
    CK(cudaEventRecord(e1, stream()));
    CK(cudaEventSynchronize(e1));
    float t = 0;
    CK(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    if (ms) *ms = t;
    return U.n;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_b200_dense_tail_bench failed: %s\n", e.what());
    return -1;
  }
}

// the tensor-core kernel ALONE on the GPU at one shape of the dense tail's deferred updates:  C (M x N) -= A (M x K) . B^T
// with device-generated operands.  Returns the mean duration of k_gemm_i8limb in ms over `reps` launches (limb split
// excluded, as in spasm_b200_mma_stats), < 0 on error or when the shape does not take the tensor-core path.
extern "C" void spasm_b200_mma_timing(int on);
extern "C" void spasm_b200_mma_stats(double *out, int reset);
extern "C" double spasm_b200_gemm_probe(long long prime, int M, int N, int K, int reps) {
  using namespace sb;
  try {
    ApiCall api_scope_;
    Fp F = make_field(prime);
    const long long ldk = ((long long)K + 15) / 16 * 16, ldc = ((long long)N + 63) / 64 * 64;
    DBuf<uint32_t> A((size_t)M * ldk), B((size_t)N * ldk), Cm((size_t)M * ldc);
    k_fill_random<<<cdiv((long long)M * ldk, 256), 256, 0, stream()>>>(A.p, (long long)M * ldk, F.p, 0x1111);
    k_fill_random<<<cdiv((long long)N * ldk, 256), 256, 0, stream()>>>(B.p, (long long)N * ldk, F.p, 0x2222);
    k_fill_random<<<cdiv((long long)M * ldc, 256), 256, 0, stream()>>>(Cm.p, (long long)M * ldc, F.p, 0x3333);
    gemm_nt(Cm.p, ldc, M, N, A.p, ldk, B.p, ldk, K, true, F);  // warm-up
    sb::sync();
    double st0[4], st1[4];
    spasm_b200_mma_timing(1);
    spasm_b200_mma_stats(st0, 0);
    for (int i = 0; i < reps; i++) gemm_nt(Cm.p, ldc, M, N, A.p, ldk, B.p, ldk, K, true, F);
    sb::sync();
    spasm_b200_mma_stats(st1, 0);
    const double calls = st1[2] - st0[2];
    return calls > 0 ? (st1[0] - st0[0]) / calls : -1.0;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_b200_gemm_probe failed: %s\n", e.what());
    return -1.0;
  }
}

// host logic of the deferred updates for the CPU tests: out = {lazy, group, kdepth, LDK}
extern "C" void spasm_b200_tail_plan(int Sm0, int n_local, int block_size, int nranks, int kcap, int max_k, long long free_bytes, long long *out) {
  const int Bmax = std::min(block_size, std::max(n_local, 1) * std::max(nranks, 1));
  const sb::TailPlan pl = sb::plan_tail(Sm0, n_local, block_size, Bmax, nranks, kcap, max_k, (size_t)free_bytes);
  out[0] = pl.lazy, out[1] = pl.group, out[2] = pl.kdepth, out[3] = pl.LDK;
}
extern "C" void spasm_b200_tail_stats(long long *out, int reset) {
  for (int i = 0; i < 4; i++) out[i] = sb::g_tail_stats[i];
  if (reset)
    for (int i = 0; i < 4; i++) sb::g_tail_stats[i] = 0;
}
// how often the dense loop switched to the low-rank mode since the last reset
extern "C" long long spasm_b200_lowrank_switches(int reset) {
  const long long v = sb::g_tail_stats[4];
  if (reset) sb::g_tail_stats[4] = 0;
  return v;
}

// the dense-tail entry point of the ABI (replaces spasm_ffpack_rref, src/SpaSM.jl:805)
extern "C" int spasm_dense_rref(i64 prime, int n, int m, spasm_ZZp *A, i64 ldA, int *pivcol_out) {
  try {
    ApiCall api_scope_;
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
    sb::sync();
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
