// solve_sparse.cu — the sparse row-solve engine: x.G = B[k] for many rows k at once.
//
// Replaces spasm_sparse_triangular_solve + spasm_reach/spasm_dfs/spasm_scatter and the OpenMP row
// loops of spasm_schur / spasm_kernel / spasm_rref (reference bindings src/SpaSM.jl:619-629,
// :694-722, :761-762, :871, :876-882; algorithm SURVEY.md A.3/A.6).
//
// B200 design.  One warp per row (light tier) or one CTA per row (medium tier).  The row under
// elimination lives in a SHARED-MEMORY open-addressing hash accumulator keyed by column; pivot
// rows of G stream through L2 with coalesced 32-lane loads of (j,x).  Instead of the reference's
// DFS (inherently sequential) the elimination order is "pending pivotal columns by increasing
// prio": every row of G only references pivots of larger prio (U rows are appended in
// topological order), so popping the minimum is a valid topological order, and since arithmetic
// in F_p is exact every valid order gives bit-identical values.  Multipliers that are zero are
// skipped (their sub-tree contributes nothing).  Output entries are emitted in increasing column
// order (normalisation N1) after an in-place compaction + bitonic sort of the accumulator.
// Reduction mod p: 32-bit Barrett for p < 2^16 (products < 2^32), 64-bit Barrett otherwise.
//
// Rows whose accumulator overflows the tier are re-run in the next tier (status 1); rows that do
// not fit the output slab are re-run with an exactly sized slab (status 2).  The heavy tier keeps
// the accumulator as a direct-indexed array in global memory with a bitmap priority queue.
#include "solve_sparse.cuh"
#include "dist.cuh"

#include "dense.cuh"

namespace sb {

static constexpr int EMPTY = -1;
static constexpr int ST_OK = 0, ST_TABLE = 1, ST_SLAB = 2;

struct __align__(16) PRow {
  long long start;
  int len;
  int pad;
};

struct KArgs {
  SolveSystem G;
  SolveRows B;
  const int *todo;  // [ntodo] indices k into B.rows
  int ntodo;
  int *work_counter;
  // emit
  int count_only, all_columns, structural, want_L;
  const int *prefix_col;
  uint32_t prefix_val;
  int *cnt;             // [nrows]
  unsigned long long *off;  // [nrows]  (slab id << 56 | offset)
  int *oj;
  uint32_t *ox;
  unsigned long long cap, slab_id;
  unsigned long long *cursor;
  int *lcnt;
  unsigned long long *loff;
  int *lj;
  uint32_t *lx;
  unsigned long long lcap;
  unsigned long long *lcursor;
  int *status;  // [nrows]
  unsigned long long *stats;  // [0] bytes, [1] macs, [2] #table overflow, [3] #slab overflow, [4] slab need, [5] L slab need
  Fp F;
  // heavy tier
  int *gkeys;            // [groups][width]
  uint32_t *gvals;       // [groups][width]
  unsigned *gbitmap;     // [groups][nprio/32+1]
  int nprio;
  // global tier (k_solve_global)
  const int2 *G2;            // entries of G relabelled (slot id, value), same offsets as Gj / Gx
  const int *col2sid;        // [width] column -> slot id: pivotal columns their prio, the others nprio + rank among the non-pivotal
  const int *sid2col;        // [width]
  const PRow *prow;          // [nprio] row of G that eliminates the pivot of that prio
  const unsigned *classbit;  // optional [nwords]: bit set at the first prio of every class of mutually independent pivots
  int nwords;                // words of the pending bitmap
  int gw;                    // slots per row: nprio + number of non-pivotal columns
  long long pop_budget;      // > 0: a row that has merged more entries of G than this gives up (status ST_TABLE: next stage)
  unsigned long long *gsum;  // [slots][width] unreduced sums; all zero between rows
  int *gtouched;             // [slots][width] slot ids touched by the current row
  unsigned *gpending;        // [slots][nwords] pending pivots of the current row; all zero between rows
};

template <int T>
struct Grp {
  __device__ static __forceinline__ void sync() {
    if (T == 32)
      __syncwarp();
    else
      __syncthreads();
  }
  __device__ static __forceinline__ int tid() { return T == 32 ? (threadIdx.x & 31) : threadIdx.x; }
  // minimum of a u64 over the group; scratch: T/32 u64 (unused for warps)
  __device__ static __forceinline__ unsigned long long min64(unsigned long long v, unsigned long long *scratch) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
      v = w < v ? w : v;
    }
    if (T == 32) return v;
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    unsigned long long r = scratch[0];
#pragma unroll
    for (int i = 1; i < T / 32; i++) r = scratch[i] < r ? scratch[i] : r;
    __syncthreads();
    return r;
  }
  // exclusive rank of `pred` among the group + total; scratch: T/32 ints
  __device__ static __forceinline__ int rank(bool pred, int &total, int *scratch) {
    unsigned b = __ballot_sync(0xffffffffu, pred);
    int l = threadIdx.x & 31;
    int pre = __popc(b & ((1u << l) - 1u)), wt = __popc(b);
    if (T == 32) {
      total = wt;
      return pre;
    }
    int w = threadIdx.x >> 5;
    if (l == 0) scratch[w] = wt;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < T / 32; i++) {
      int v = scratch[i];
      if (i < w) base += v;
      tot += v;
    }
    __syncthreads();
    total = tot;
    return base + pre;
  }
  __device__ static __forceinline__ int bcast(int v, int *scratch) {
    if (T == 32) return __shfl_sync(0xffffffffu, v, 0);
    if (threadIdx.x == 0) scratch[0] = v;
    __syncthreads();
    int r = scratch[0];
    __syncthreads();
    return r;
  }
};

__device__ __forceinline__ unsigned hash_col(int c) { return (unsigned)c * 0x9E3779B1u; }

// ---------------------------------------------------------------------------------------------
// shared-memory tiers.  T threads per row, H table slots, PCAP pending entries.
template <int T, int H, int PCAP, bool SMALL>
__global__ void __launch_bounds__(T == 32 ? 256 : T) k_solve_smem(KArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int GROUPS = (T == 32) ? 8 : 1;  // row groups per block
  constexpr int LOGH = (H == 512) ? 9 : (H == 1024) ? 10 : (H == 2048) ? 11 : (H == 4096) ? 12 : (H == 8192) ? 13 : 14;
  static_assert((1 << LOGH) == H, "H must be a supported power of two");
  const int g = (T == 32) ? (threadIdx.x >> 5) : 0;
  const int tid = Grp<T>::tid();
  // layout per group: keys[H] vals[H] pprio[PCAP] pcol[PCAP] pslot[PCAP]
  constexpr size_t PER = (size_t)(2 * H + 3 * PCAP) * 4;
  unsigned char *base = smem_raw + (size_t)g * PER;
  int *keys = (int *)base;
  uint32_t *vals = (uint32_t *)(keys + H);
  int *pprio = (int *)(vals + H);
  int *pcol = pprio + PCAP;
  int *pslot = pcol + PCAP;
  unsigned long long *red64 = (unsigned long long *)(smem_raw + (size_t)GROUPS * PER);
  int *redi = (int *)(red64 + 32);
  const Fp F = a.F;
  const PDesc *__restrict__ pdesc = a.G.pdesc;
  const int *__restrict__ Gj = a.G.Gj;
  const uint32_t *__restrict__ Gx = a.G.Gx;
  constexpr int LIMIT = H - H / 8 - T;  // keep the load factor below 7/8

  for (;;) {
    int t = 0;
    if (tid == 0) t = atomicAdd(a.work_counter, 1);
    t = Grp<T>::bcast(t, redi);
    if (t >= a.ntodo) break;
    const int k = a.todo ? a.todo[t] : t;
    const int brow = a.B.rows ? a.B.rows[k] : k;
    const int maskc = a.B.mask ? a.B.mask[k] : -1;
    for (int s = tid; s < H; s += T) keys[s] = EMPTY;
    Grp<T>::sync();
    int fill = 0, np = 0;
    bool overflow = false;
    unsigned long long bytes = 0, macs = 0;

    // one batch of <= T entries (cj, delta) is merged into the accumulator
    auto merge = [&](bool valid, int cj, uint32_t delta) {
      int slot = -1;
      bool isnew = false;
      if (valid) {
        unsigned h = hash_col(cj) >> (32 - LOGH);
        for (;;) {
          int kk = ((volatile int *)keys)[h];
          if (kk == cj) {
            vals[h] = addmod(vals[h], delta, F);
            break;
          }
          if (kk == EMPTY) {
            int old = atomicCAS(&keys[h], EMPTY, cj);
            if (old == EMPTY) {
              vals[h] = delta;
              isnew = true;
              break;
            }
            if (old == cj) {
              vals[h] = addmod(vals[h], delta, F);
              break;
            }
          }
          h = (h + 1) & (H - 1);
        }
        slot = (int)h;
      }
      int prio = -1;
      bool push = false;
      if (isnew && cj != maskc) {
        PDesc d = pdesc[cj];
        if (d.len >= 0) push = true, prio = d.prio;
      }
      int totnew, totpush;
      Grp<T>::rank(isnew, totnew, redi);
      int pos = Grp<T>::rank(push, totpush, redi);
      if (np + totpush > PCAP) {
        overflow = true;
      } else if (push) {
        pprio[np + pos] = prio;
        pcol[np + pos] = cj;
        pslot[np + pos] = slot;
      }
      np += totpush;
      fill += totnew;
      Grp<T>::sync();
    };

    // ---- load B[k]
    {
      const long long b0 = a.B.Bp[brow], b1 = a.B.Bp[brow + 1];
      bytes += 8 * (b1 - b0) + 8;
      for (long long e0 = b0; e0 < b1 && !overflow; e0 += T) {
        if (fill > LIMIT) {
          overflow = true;
          break;
        }
        long long e = e0 + tid;
        bool valid = e < b1;
        int cj = valid ? a.B.Bj[e] : 0;
        uint32_t v = valid ? a.B.Bx[e] : 0;
        merge(valid, cj, v);
      }
    }
    // ---- eliminate pending pivots by increasing prio
    while (np > 0 && !overflow) {
      unsigned long long best = ~0ULL;
      for (int i = tid; i < np; i += T) {
        unsigned long long key = ((unsigned long long)(unsigned)pprio[i] << 32) | (unsigned)i;
        best = key < best ? key : best;
      }
      best = Grp<T>::min64(best, red64);
      const int idx = (int)(best & 0xffffffffu);
      const int c = pcol[idx];
      const uint32_t mult = vals[pslot[idx]];
      Grp<T>::sync();
      if (tid == 0) {
        pprio[idx] = pprio[np - 1];
        pcol[idx] = pcol[np - 1];
        pslot[idx] = pslot[np - 1];
      }
      np--;
      Grp<T>::sync();
      if (mult == 0 && !a.structural) continue;
      const PDesc d = pdesc[c];
      const uint32_t coef = negmod(mult, F);
      bytes += 8 * (long long)d.len + 8;
      macs += d.len;
      for (int e0 = 0; e0 < d.len; e0 += T) {
        if (fill > LIMIT) {
          overflow = true;
          break;
        }
        int e = e0 + tid;
        bool valid = e < d.len;
        int cj = 0;
        uint32_t delta = 0;
        if (valid) {
          cj = Gj[d.start + e];
          if (cj == c)
            valid = false;  // the unit pivot entry: x[c] keeps the multiplier
          else
            delta = mulmod<SMALL>(coef, Gx[d.start + e], F);
        }
        merge(valid, cj, delta);
      }
    }
    if (overflow) {
      if (tid == 0) {
        a.status[k] = ST_TABLE;
        atomicAdd(&a.stats[2], 1ULL);
      }
      Grp<T>::sync();
      continue;
    }

    // ---- emit.  (1) multipliers -> pend arrays (free now), (2) in-place compaction, (3) sort
    int nl = 0;
    if (a.want_L) {
      for (int s0 = 0; s0 < H && !overflow; s0 += T) {
        int s = s0 + tid;
        int kk = keys[s];
        uint32_t v = vals[s];
        bool pass = false;
        int prio = 0;
        if (kk != EMPTY && v != 0 && kk != maskc) {
          PDesc d = pdesc[kk];
          if (d.len >= 0) pass = true, prio = d.prio;
        }
        int tot;
        int pos = Grp<T>::rank(pass, tot, redi);
        if (nl + tot > PCAP)
          overflow = true;
        else if (pass) {
          pprio[nl + pos] = prio;
          pcol[nl + pos] = (int)v;
        }
        nl += tot;
      }
      Grp<T>::sync();
      if (overflow) {
        if (tid == 0) {
          a.status[k] = ST_TABLE;
          atomicAdd(&a.stats[2], 1ULL);
        }
        Grp<T>::sync();
        continue;
      }
    }
    int nout = 0;
    for (int s0 = 0; s0 < H; s0 += T) {
      int s = s0 + tid;
      int kk = keys[s];
      uint32_t v = vals[s];
      bool pass = kk != EMPTY && (v != 0 || a.structural);
      if (pass && !a.all_columns && !a.structural) pass = (kk == maskc) ? false : (pdesc[kk].len < 0);
      if (pass && a.B.mask && kk == maskc) pass = false;  // own pivot of an rref row goes to the prefix
      int tot;
      int pos = Grp<T>::rank(pass, tot, redi);
      Grp<T>::sync();  // all reads of this batch precede the writes below
      if (pass) {
        keys[nout + pos] = kk;
        vals[nout + pos] = v;
      }
      nout += tot;
      Grp<T>::sync();
    }
    const int npre = a.prefix_col ? 1 : 0;
    if (tid == 0) {
      a.cnt[k] = nout + npre;
      if (a.want_L) a.lcnt[k] = nl;
    }
    if (a.count_only) {
      if (tid == 0) {
        a.status[k] = ST_OK;
        atomicAdd(&a.stats[0], bytes + 8ULL * nout + 8ULL);
        atomicAdd(&a.stats[1], macs);
      }
      Grp<T>::sync();
      continue;
    }
    // bitonic sort of (keys, vals)[0:nout) by key, and of (pprio, pcol)[0:nl) by prio
    auto bitonic = [&](int *kk, uint32_t *vv, int n) {
      int N = 1;
      while (N < n) N <<= 1;
      for (int i = n + tid; i < N; i += T) kk[i] = 0x7fffffff;
      Grp<T>::sync();
      for (int size = 2; size <= N; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int i = tid; i < (N >> 1); i += T) {
            int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
            int hi = lo | stride;
            bool up = ((lo & size) == 0);
            int ka = kk[lo], kb = kk[hi];
            if ((ka > kb) == up) {
              kk[lo] = kb, kk[hi] = ka;
              uint32_t tv = vv[lo];
              vv[lo] = vv[hi], vv[hi] = tv;
            }
          }
          Grp<T>::sync();
        }
    };
    if (nout > 1) bitonic(keys, vals, nout);
    if (nl > 1) bitonic(pprio, (uint32_t *)pcol, nl);
    // slab allocation
    unsigned long long o = 0, lo_ = 0;
    int st = ST_OK;
    if (tid == 0) {
      o = atomicAdd(a.cursor, (unsigned long long)(nout + npre));
      if (o + nout + npre > a.cap) st = ST_SLAB;
      if (a.want_L) {
        lo_ = atomicAdd(a.lcursor, (unsigned long long)nl);
        if (lo_ + nl > a.lcap) st = ST_SLAB;
      }
      if (st == ST_SLAB) {
        atomicAdd(&a.stats[3], 1ULL);
        atomicAdd(&a.stats[4], (unsigned long long)(nout + npre));
        atomicAdd(&a.stats[5], (unsigned long long)nl);
      } else {  // the work is counted once, by the attempt whose output is kept
        atomicAdd(&a.stats[0], bytes + 8ULL * nout + 8ULL);
        atomicAdd(&a.stats[1], macs);
      }
      a.status[k] = st;
      a.off[k] = (a.slab_id << 56) | o;
      if (a.want_L) a.loff[k] = (a.slab_id << 56) | lo_;
    }
    if (T == 32) {
      st = __shfl_sync(0xffffffffu, st, 0);
      o = __shfl_sync(0xffffffffu, o, 0);
      lo_ = __shfl_sync(0xffffffffu, lo_, 0);
    } else {
      __shared__ unsigned long long bc[3];
      if (tid == 0) bc[0] = (unsigned long long)st, bc[1] = o, bc[2] = lo_;
      __syncthreads();
      st = (int)bc[0], o = bc[1], lo_ = bc[2];
      __syncthreads();
    }
    if (st == ST_OK) {
      if (npre && tid == 0) {
        a.oj[o] = a.prefix_col[k];
        a.ox[o] = a.prefix_val;
      }
      for (int i = tid; i < nout; i += T) {
        a.oj[o + npre + i] = keys[i];
        a.ox[o + npre + i] = vals[i];
      }
      for (int i = tid; i < nl; i += T) {
        a.lj[lo_ + i] = pprio[i];
        a.lx[lo_ + i] = (uint32_t)pcol[i];
      }
    }
    Grp<T>::sync();
  }
}

// ---------------------------------------------------------------------------------------------
// heavy tier: CTA per row, direct-indexed accumulator in global memory (slot = column), pending
// pivots in a bitmap over prio.  Output comes out column-sorted by scanning the slots in order.
template <int T, bool SMALL>
__global__ void __launch_bounds__(T) k_solve_heavy(KArgs a, const int *__restrict__ prio2col) {
  __shared__ unsigned long long red64[32];
  __shared__ int redi[40];
  const int tid = threadIdx.x;
  const Fp F = a.F;
  const int W = a.G.width;
  const int nwords = (a.nprio + 31) / 32 + 1;
  int *keys = a.gkeys + (size_t)blockIdx.x * W;          // keys[c] == c when present else EMPTY (kept EMPTY between rows)
  uint32_t *vals = a.gvals + (size_t)blockIdx.x * W;
  unsigned *bitmap = a.gbitmap + (size_t)blockIdx.x * nwords;  // kept all-zero between rows
  const PDesc *__restrict__ pdesc = a.G.pdesc;

  for (;;) {
    int t = 0;
    if (tid == 0) t = atomicAdd(a.work_counter, 1);
    t = Grp<T>::bcast(t, redi);
    if (t >= a.ntodo) break;
    const int k = a.todo ? a.todo[t] : t;
    const int brow = a.B.rows ? a.B.rows[k] : k;
    const int maskc = a.B.mask ? a.B.mask[k] : -1;
    unsigned long long bytes = 0, macs = 0;

    auto merge = [&](bool valid, int cj, uint32_t delta) {
      if (!valid) return;
      if (keys[cj] == cj) {
        vals[cj] = addmod(vals[cj], delta, F);
      } else {
        keys[cj] = cj;
        vals[cj] = delta;
        if (cj != maskc) {
          PDesc d = pdesc[cj];
          if (d.len >= 0) atomicOr(&bitmap[d.prio >> 5], 1u << (d.prio & 31));
        }
      }
    };
    const long long b0 = a.B.Bp[brow], b1 = a.B.Bp[brow + 1];
    bytes += 8 * (b1 - b0) + 8;
    for (long long e = b0 + tid; e < b1; e += T) merge(true, a.B.Bj[e], a.B.Bx[e]);
    __syncthreads();
    int cursor = 0;  // word index
    for (;;) {
      // find the first set bit at or after word `cursor`
      unsigned long long best = ~0ULL;
      for (int w0 = cursor; w0 < nwords; w0 += T) {
        int w = w0 + tid;
        unsigned bits = (w < nwords) ? bitmap[w] : 0u;
        if (bits) best = (unsigned long long)w * 32 + (__ffs(bits) - 1);
        best = Grp<T>::min64(best, red64);
        if (best != ~0ULL) break;
      }
      if (best == ~0ULL) break;
      const int prio = (int)best;
      cursor = prio >> 5;
      const int c = prio2col[prio];
      if (tid == 0) bitmap[prio >> 5] &= ~(1u << (prio & 31));
      const uint32_t mult = vals[c];
      __syncthreads();
      if (mult == 0 && !a.structural) continue;
      const PDesc d = pdesc[c];
      const uint32_t coef = negmod(mult, F);
      bytes += 8 * (long long)d.len + 8;
      macs += d.len;
      for (int e = tid; e < d.len; e += T) {
        int cj = a.G.Gj[d.start + e];
        if (cj != c) merge(true, cj, mulmod<SMALL>(coef, a.G.Gx[d.start + e], F));
      }
      __syncthreads();
    }
    // pass 1: count
    int nout = 0, nl = 0;
    for (int s0 = 0; s0 < W; s0 += T) {
      int s = s0 + tid;
      bool present = s < W && keys[s] == s;
      uint32_t v = present ? vals[s] : 0;
      bool piv = present && s != maskc && pdesc[s].len >= 0;
      bool pass = present && (v != 0 || a.structural) && (a.all_columns || a.structural || !piv) && s != maskc;
      bool lpass = a.want_L && present && v != 0 && piv;
      int tot;
      Grp<T>::rank(pass, tot, redi);
      nout += tot;
      if (a.want_L) {
        Grp<T>::rank(lpass, tot, redi);
        nl += tot;
      }
    }
    const int npre = a.prefix_col ? 1 : 0;
    unsigned long long o = 0, lo_ = 0;
    int st = ST_OK;
    if (tid == 0) {
      a.cnt[k] = nout + npre;
      if (a.want_L) a.lcnt[k] = nl;
      if (!a.count_only) {
        o = atomicAdd(a.cursor, (unsigned long long)(nout + npre));
        if (o + nout + npre > a.cap) st = ST_SLAB;
        if (a.want_L) {
          lo_ = atomicAdd(a.lcursor, (unsigned long long)nl);
          if (lo_ + nl > a.lcap) st = ST_SLAB;
        }
        if (st == ST_SLAB) {
          atomicAdd(&a.stats[3], 1ULL);
          atomicAdd(&a.stats[4], (unsigned long long)(nout + npre));
          atomicAdd(&a.stats[5], (unsigned long long)nl);
        }
        a.off[k] = (a.slab_id << 56) | o;
        if (a.want_L) a.loff[k] = (a.slab_id << 56) | lo_;
      }
      if (st == ST_OK) {  // the work is counted once, by the attempt whose output is kept
        atomicAdd(&a.stats[0], bytes + 8ULL * nout + 8ULL);
        atomicAdd(&a.stats[1], macs);
      }
      a.status[k] = st;
    }
    {
      __shared__ unsigned long long bc[3];
      if (tid == 0) bc[0] = (unsigned long long)st, bc[1] = o, bc[2] = lo_;
      __syncthreads();
      st = (int)bc[0], o = bc[1], lo_ = bc[2];
      __syncthreads();
    }
    const bool write = !a.count_only && st == ST_OK;
    if (write && npre && tid == 0) {
      a.oj[o] = a.prefix_col[k];
      a.ox[o] = a.prefix_val;
    }
    // pass 2: write (column order) + reset the accumulator.  The L stream must be by increasing
    // prio: it is written through the prio2col map in a third pass below.
    int w_ = 0;
    for (int s0 = 0; s0 < W; s0 += T) {
      int s = s0 + tid;
      bool present = s < W && keys[s] == s;
      uint32_t v = present ? vals[s] : 0;
      bool piv = present && s != maskc && pdesc[s].len >= 0;
      bool pass = present && (v != 0 || a.structural) && (a.all_columns || a.structural || !piv) && s != maskc;
      int tot;
      int pos = Grp<T>::rank(pass, tot, redi);
      if (write && pass) {
        a.oj[o + npre + w_ + pos] = s;
        a.ox[o + npre + w_ + pos] = v;
      }
      w_ += tot;
      if (present && !(a.want_L && piv)) keys[s] = EMPTY;
    }
    if (a.want_L) {
      __syncthreads();
      int lw = 0;
      for (int q0 = 0; q0 < a.nprio; q0 += T) {
        int q = q0 + tid;
        int c = q < a.nprio ? prio2col[q] : -1;
        bool present = c >= 0 && keys[c] == c;
        uint32_t v = present ? vals[c] : 0;
        bool lpass = present && v != 0 && c != maskc;
        int tot;
        int pos = Grp<T>::rank(lpass, tot, redi);
        if (write && lpass) {
          a.lj[lo_ + lw + pos] = q;
          a.lx[lo_ + lw + pos] = v;
        }
        lw += tot;
        if (present) keys[c] = EMPTY;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// GLOBAL tier: one warp per row, THOUSANDS of rows in flight.  Rows that do not fit the shared-memory table
// (hundreds to tens of thousands of touched columns, hundreds of pivots reached) are latency-bound: every
// elimination step is a chain of dependent L2 round trips, so throughput comes from concurrency, not from
// a faster step.  The accumulator is therefore a direct-indexed array in global memory — no probing, no
// capacity tiers — at 16 warps per SM (2368 rows in flight on a B200):
//   * slot id: pivotal columns are addressed by their prio, the others by nprio + rank, so the multiplier of a
//     pending pivot and the pending bitmap share one index and G is read as ONE 8-byte (slot, value) stream;
//   * sum[slot] is an UNREDUCED 64-bit sum of residues, updated with atomicAdd (several rows of G are merged
//     in the same warp step, so two lanes may hit one slot); every addend is a non-zero residue, hence
//     "old == 0" detects the first touch exactly once: that lane logs the slot and sets its pending bit;
//   * pending pivots live in a bitmap over prio, scanned 32 words per step from a monotone cursor; with
//     `classbit` all pending pivots of one independence class inside a word are eliminated together
//     (their rows are flattened over the 32 lanes); the sums are reduced mod p only when read;
//   * emission walks the touched log (reduce, filter, write, zero the slot); rows come out unsorted and are
//     sorted by k_sort_rows (normalisation N1).
// Any valid topological order gives bit-identical values (exact arithmetic), so this tier, the shared-memory
// tier and the oracle agree.  Zero multipliers are skipped (no structural mode here).
template <bool SMALL>
__global__ void __launch_bounds__(128) k_solve_global(KArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t slot = (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  unsigned long long *__restrict__ sum = a.gsum + slot * (size_t)a.gw;
  int *__restrict__ touched = a.gtouched + slot * (size_t)a.gw;
  unsigned *__restrict__ pending = a.gpending + slot * (size_t)a.nwords;
  const Fp F = a.F;
  const int nprio = a.nprio;
  const int2 *__restrict__ G2 = a.G2;

  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(a.work_counter, 1);
    t = __shfl_sync(FULL, t, 0);
    if (t >= a.ntodo) break;
    const int k = a.todo ? a.todo[t] : t;
    const int brow = a.B.rows ? a.B.rows[k] : k;
    const int masksid = a.B.mask ? a.col2sid[a.B.mask[k]] : -1;
    int nt = 0, wmin = a.nwords;
    unsigned long long bytes = 0, macs = 0;

    // one warp step: `valid` lanes add delta to sum[sid]; first touches are logged and made pending
    auto merge = [&](bool valid, int sid, uint32_t delta) {
      bool first = false;
      if (valid) first = atomicAdd(&sum[sid], (unsigned long long)delta) == 0ULL;
      const unsigned fb = __ballot_sync(FULL, first);
      int myw = 0x7fffffff;
      if (first) {
        touched[nt + __popc(fb & lt_mask)] = sid;
        if (sid < nprio && sid != masksid) {
          atomicOr(&pending[sid >> 5], 1u << (sid & 31));
          myw = sid >> 5;
        }
      }
      nt += __popc(fb);
#pragma unroll
      for (int o = 16; o; o >>= 1) myw = min(myw, __shfl_xor_sync(FULL, myw, o));
      wmin = min(wmin, myw);
    };

    // ---- B[k]
    {
      const long long b0 = a.B.Bp[brow], b1 = a.B.Bp[brow + 1];
      bytes += 8 * (b1 - b0) + 8;
      for (long long e0 = b0; e0 < b1; e0 += 32) {
        const long long e = e0 + lane;
        bool valid = e < b1;
        int sid = 0;
        uint32_t v = 0;
        if (valid) {
          v = a.B.Bx[e];
          sid = a.col2sid[a.B.Bj[e]];
          valid = v != 0;
        }
        merge(valid, sid, v);
      }
    }
    __syncwarp();
    // ---- eliminate the pending pivots by increasing prio, one class (inside one word) per step
    int cw = wmin;
    bool gave_up = false;
    while (cw < a.nwords) {
      if (a.pop_budget > 0 && (long long)macs > a.pop_budget) {
        gave_up = true;
        break;
      }
      const unsigned bw = (cw + lane < a.nwords) ? __ldcg(&pending[cw + lane]) : 0u;
      const unsigned nz = __ballot_sync(FULL, bw != 0);
      if (nz == 0) {
        cw += 32;
        continue;
      }
      const int f = __ffs(nz) - 1;
      const unsigned w0 = __shfl_sync(FULL, bw, f);
      const int wi = cw + f;
      const int bit0 = __ffs(w0) - 1;
      unsigned mw = 1u << bit0;
      if (a.classbit != nullptr) {
        const unsigned cb = __ldg(&a.classbit[wi]);
        const unsigned above = (bit0 == 31) ? 0u : (cb & ~((2u << bit0) - 1u));  // class starts strictly after bit0
        const int end = above ? (__ffs(above) - 1) : 32;
        const unsigned upto = (end == 32) ? 0xffffffffu : ((1u << end) - 1u);
        mw = w0 & upto & ~((1u << bit0) - 1u);
      }
      const int nm = __popc(mw);
      if (lane == 0) atomicAnd(&pending[wi], ~mw);
      int mprio = -1, mlen = 0;
      long long mstart = 0;
      uint32_t coef = 0;
      if (lane < nm) {
        mprio = wi * 32 + (int)__fns(mw, 0, lane + 1);
        const uint32_t mult = red64(__ldcg(&sum[mprio]), F);
        if (mult != 0) {
          const PRow pr = a.prow[mprio];
          mstart = pr.start, mlen = pr.len, coef = F.p - mult;
        }
      }
      int incl = mlen;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      const int excl_mine = incl - mlen;
      const int total = __shfl_sync(FULL, incl, 31);
      const int nzm = __popc(__ballot_sync(FULL, mlen > 0));
      macs += total, bytes += 8ULL * total + 8ULL * nzm;
      for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        int lo = 0, hi = 31;  // smallest member whose inclusive prefix exceeds idx
#pragma unroll
        for (int s_ = 0; s_ < 5; s_++) {
          const int mid = (lo + hi) >> 1;
          const int v = __shfl_sync(FULL, incl, mid);
          if (v > idx)
            hi = mid;
          else
            lo = mid + 1;
        }
        const int ex = __shfl_sync(FULL, excl_mine, lo);
        const long long st = __shfl_sync(FULL, mstart, lo);
        const uint32_t cf = __shfl_sync(FULL, coef, lo);
        const int mp = __shfl_sync(FULL, mprio, lo);
        bool valid = idx < total;
        int sid = 0;
        uint32_t delta = 0;
        if (valid) {
          const int2 ent = __ldg(&G2[st + (idx - ex)]);
          sid = ent.x;
          if (sid == mp)
            valid = false;  // the unit pivot entry: sum[prio] keeps the multiplier
          else
            delta = mulmod<SMALL>(cf, (uint32_t)ent.y, F);
        }
        merge(valid, sid, delta);
      }
      __syncwarp();
      cw = wi;
    }

    if (gave_up) {
      // a chain this long belongs to the column-major engine: leave the accumulator and the bitmap clean
      for (int t0 = 0; t0 < nt; t0 += 32) {
        const int i = t0 + lane;
        if (i < nt) {
          const int sid = __ldcg(&touched[i]);
          sum[sid] = 0ULL;
          if (sid < nprio) atomicAnd(&pending[sid >> 5], ~(1u << (sid & 31)));
        }
      }
      if (lane == 0) {
        a.status[k] = ST_TABLE;
        atomicAdd(&a.stats[2], 1ULL);
      }
      __syncwarp();
      continue;
    }
    // ---- emit.  pass 1: count
    int nout = 0, nl = 0;
    for (int t0 = 0; t0 < nt; t0 += 32) {
      const int i = t0 + lane;
      const bool valid = i < nt;
      const int sid = valid ? __ldcg(&touched[i]) : 0;
      const uint32_t v = valid ? red64(__ldcg(&sum[sid]), F) : 0u;
      const bool piv = valid && sid < nprio && sid != masksid;
      const bool pass = valid && v != 0 && sid != masksid && (a.all_columns || !piv);
      const bool lpass = a.want_L && v != 0 && piv;
      nout += __popc(__ballot_sync(FULL, pass));
      nl += __popc(__ballot_sync(FULL, lpass));
    }
    const int npre = a.prefix_col ? 1 : 0;
    unsigned long long o = 0, lo_ = 0;
    int st = ST_OK;
    if (lane == 0) {
      a.cnt[k] = nout + npre;
      if (a.want_L) a.lcnt[k] = nl;
      if (!a.count_only) {
        o = atomicAdd(a.cursor, (unsigned long long)(nout + npre));
        if (o + nout + npre > a.cap) st = ST_SLAB;
        if (a.want_L) {
          lo_ = atomicAdd(a.lcursor, (unsigned long long)nl);
          if (lo_ + nl > a.lcap) st = ST_SLAB;
        }
        if (st == ST_SLAB) {
          atomicAdd(&a.stats[3], 1ULL);
          atomicAdd(&a.stats[4], (unsigned long long)(nout + npre));
          atomicAdd(&a.stats[5], (unsigned long long)nl);
        }
        a.off[k] = (a.slab_id << 56) | o;
        if (a.want_L) a.loff[k] = (a.slab_id << 56) | lo_;
      }
      if (st == ST_OK) {  // the work is counted once, by the attempt whose output is kept
        atomicAdd(&a.stats[0], bytes + 8ULL * nout + 8ULL);
        atomicAdd(&a.stats[1], macs);
      }
      a.status[k] = st;
    }
    st = __shfl_sync(FULL, st, 0);
    o = __shfl_sync(FULL, o, 0);
    lo_ = __shfl_sync(FULL, lo_, 0);
    const bool write = !a.count_only && st == ST_OK;
    if (write && npre && lane == 0) {
      a.oj[o] = a.prefix_col[k];
      a.ox[o] = a.prefix_val;
    }
    // pass 2: write (unsorted: k_sort_rows orders them) and give the slots back as zeros
    int wo = 0, wl = 0;
    for (int t0 = 0; t0 < nt; t0 += 32) {
      const int i = t0 + lane;
      const bool valid = i < nt;
      const int sid = valid ? __ldcg(&touched[i]) : 0;
      const uint32_t v = valid ? red64(__ldcg(&sum[sid]), F) : 0u;
      const bool piv = valid && sid < nprio && sid != masksid;
      const bool pass = valid && v != 0 && sid != masksid && (a.all_columns || !piv);
      const bool lpass = a.want_L && v != 0 && piv;
      const unsigned pb = __ballot_sync(FULL, pass), lb = __ballot_sync(FULL, lpass);
      if (valid) sum[sid] = 0ULL;
      if (write && pass) {
        const unsigned long long d = o + npre + wo + __popc(pb & lt_mask);
        a.oj[d] = a.sid2col[sid];
        a.ox[d] = v;
      }
      if (write && lpass) {
        const unsigned long long d = lo_ + wl + __popc(lb & lt_mask);
        a.lj[d] = sid;
        a.lx[d] = v;
      }
      wo += __popc(pb), wl += __popc(lb);
    }
    __syncwarp();
  }
}

// rows written by the global tier: sort (oj, ox)[off+npre, off+cnt) by column and the L stream by prio.
// One CTA per row; rows of at most SORT_CAP entries are sorted in shared memory, longer ones in place.
static constexpr int SORT_CAP = 4096;
// bitonic network for an arbitrary length n: every compare-exchange is ascending (the first step of each merge
// pairs i with its mirror image), so the virtual +inf elements at positions >= n never move
__device__ void bitonic_any(int *kk, uint32_t *vv, int n) {
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
__device__ void sort_segment(int *gj, uint32_t *gx, int n, int *sk, uint32_t *sv) {
  if (n <= 1) return;
  if (n <= SORT_CAP) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = gj[i], sv[i] = gx[i];
    __syncthreads();
    bitonic_any(sk, sv, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) gj[i] = sk[i], gx[i] = sv[i];
    __syncthreads();
  } else {
    __syncthreads();
    bitonic_any(gj, gx, n);  // in place in global memory (rare: rows with more than SORT_CAP entries)
  }
}
__global__ void __launch_bounds__(256) k_sort_rows(const int *__restrict__ todo, int ntodo, const int *__restrict__ status, const int *__restrict__ cnt,
                                                    const unsigned long long *__restrict__ off, int npre, int *__restrict__ oj,
                                                    uint32_t *__restrict__ ox, unsigned long long slab_id, const int *__restrict__ lcnt,
                                                    const unsigned long long *__restrict__ loff, int *__restrict__ lj, uint32_t *__restrict__ lx) {
  __shared__ int sk[SORT_CAP];
  __shared__ uint32_t sv[SORT_CAP];
  for (int t = blockIdx.x; t < ntodo; t += gridDim.x) {
    const int k = todo ? todo[t] : t;
    if (status[k] != ST_OK) continue;
    const unsigned long long o = off[k];
    if ((o >> 56) != slab_id) continue;
    const unsigned long long base = o & ((1ULL << 56) - 1);
    sort_segment(oj + base + npre, ox + base + npre, cnt[k] - npre, sk, sv);
    if (lcnt != nullptr) {
      const unsigned long long lb = loff[k] & ((1ULL << 56) - 1);
      sort_segment(lj + lb, lx + lb, lcnt[k], sk, sv);
    }
  }
}

// every row of a device CSR sorted by column (rows of arbitrary length; used by the batched gesv)
__global__ void __launch_bounds__(256) k_sort_csr_rows(const long long *__restrict__ p, int n, int *__restrict__ j, uint32_t *__restrict__ x) {
  __shared__ int sk[SORT_CAP];
  __shared__ uint32_t sv[SORT_CAP];
  for (int r = blockIdx.x; r < n; r += gridDim.x) sort_segment(j + p[r], x + p[r], (int)(p[r + 1] - p[r]), sk, sv);
}
void sort_csr_rows(const long long *p, int n, int *j, uint32_t *x) {
  if (n > 0) k_sort_csr_rows<<<std::min(n, sm_count() * 8), 256, 0, stream()>>>(p, n, j, x);
  CK(cudaGetLastError());
}

// slot ids of the global tier
__global__ void k_sys_extent(const PDesc *__restrict__ pdesc, int width, unsigned long long *__restrict__ out /* [0] nprio, [1] nnz(G) */,
                             int *__restrict__ isfree) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > width) return;
  if (c == width) {
    isfree[c] = 0;
    return;
  }
  const PDesc d = pdesc[c];
  isfree[c] = d.len < 0;
  if (d.len >= 0) {
    atomicMax(&out[0], (unsigned long long)d.prio + 1ULL);
    atomicMax(&out[1], (unsigned long long)(d.start + d.len));
  }
}
__global__ void k_build_sid(const PDesc *__restrict__ pdesc, const long long *__restrict__ pos, int width, int nprio, int *__restrict__ col2sid,
                            int *__restrict__ sid2col, PRow *__restrict__ prow) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= width) return;
  const PDesc d = pdesc[c];
  int sid;
  if (d.len >= 0) {
    sid = d.prio;
    PRow r;
    r.start = d.start, r.len = d.len, r.pad = c;
    prow[sid] = r;
  } else
    sid = nprio + (int)pos[c];
  col2sid[c] = sid;
  sid2col[sid] = c;
}
__global__ void k_relabel_G2(const int *__restrict__ Gj, const uint32_t *__restrict__ Gx, long long nnz, const int *__restrict__ col2sid,
                             int2 *__restrict__ G2) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) G2[e] = make_int2(col2sid[Gj[e]], (int)Gx[e]);
}

// ---------------------------------------------------------------------------------------------
__global__ void k_collect_status(const int *__restrict__ status, const int *__restrict__ todo, int ntodo, int want,
                                 int *__restrict__ out, int *__restrict__ nout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntodo) return;
  int k = todo ? todo[i] : i;
  if (status[k] == want) out[atomicAdd(nout, 1)] = k;
}
__global__ void k_gather_rows_sel(const int *__restrict__ rows, const int *__restrict__ todo, int n, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rows ? rows[todo[i]] : todo[i];
}
__global__ void k_fill(int *a, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
struct SlabPtrs {
  const int *j[64];
  const uint32_t *x[64];
};
__global__ void k_gather_rows(const int *__restrict__ cnt, const unsigned long long *__restrict__ off,
                              const long long *__restrict__ p, int nrows, SlabPtrs S, int *__restrict__ oj,
                              uint32_t *__restrict__ ox) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nrows) return;
  int c = cnt[w];
  unsigned long long o = off[w];
  int sid = (int)(o >> 56);
  o &= (1ULL << 56) - 1;
  const int *sj = S.j[sid];
  const uint32_t *sx = S.x[sid];
  long long d = p[w];
  for (int i = lane; i < c; i += 32) {
    oj[d + i] = sj[o + i];
    ox[d + i] = sx[o + i];
  }
}
__global__ void k_rows_io_bytes(const long long *__restrict__ Bp, const int *__restrict__ rows_sel, const int *__restrict__ todo, int off, int nr,
                                const int *__restrict__ cnt, unsigned long long *__restrict__ work) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long b = 0;
  if (i < nr) {
    const int row = rows_sel[i];
    b = 8ULL * (unsigned long long)(Bp[row + 1] - Bp[row]) + 8ULL + 8ULL * (unsigned long long)cnt[todo[off + i]] + 8ULL;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&work[0], b);
}
__global__ void k_prio2col(const PDesc *__restrict__ pdesc, int width, int *__restrict__ prio2col) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < width && pdesc[c].len >= 0) prio2col[pdesc[c].prio] = c;
}
__global__ void k_sort_check(const int *todo, int n) {}

// order-preserving compaction of a todo list is not needed: rows are written by k.

void solve_rows_local(const SolveSystem &G, const SolveRows &B, const SolveEmit &E, const Fp &F, SolveResult &R) {
  const int nrows = B.nrows;
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0));
  CK(cudaEventCreate(&ev1));
  CK(cudaEventRecord(ev0, stream()));
  R.cnt.alloc(nrows + 1);
  R.cnt.zero();
  R.p.alloc(nrows + 1);
  if (E.want_L) {
    R.lcnt.alloc(nrows + 1);
    R.lcnt.zero();
    R.lp.alloc(nrows + 1);
  }
  R.nnz = R.lnnz = 0;
  R.stats = WorkStats();
  R.stats.rows = nrows;
  if (nrows == 0) {
    R.p.zero();
    if (E.want_L) R.lp.zero();
    R.j.alloc(0), R.x.alloc(0);
    return;
  }
  DBuf<int> status(nrows);
  DBuf<unsigned long long> off(nrows), loff(E.want_L ? nrows : 1);
  DBuf<unsigned long long> ctrs(8);  // [0..5] stats, [6] cursor, [7] lcursor
  DBuf<int> counter(2);
  ctrs.zero();
  std::vector<DBuf<int>> slabs_j, slabs_lj;
  std::vector<DBuf<uint32_t>> slabs_x, slabs_lx;
  SlabPtrs SP{}, LSP{};

  KArgs a{};
  a.G = G, a.B = B;
  a.count_only = E.count_only, a.all_columns = E.all_columns, a.structural = E.structural, a.want_L = E.want_L;
  a.prefix_col = E.prefix_col, a.prefix_val = E.prefix_val;
  a.cnt = R.cnt.p, a.off = off.p, a.lcnt = R.lcnt.p, a.loff = loff.p;
  a.status = status.p, a.stats = ctrs.p, a.cursor = ctrs.p + 6, a.lcursor = ctrs.p + 7;
  a.F = F;
  a.work_counter = counter.p;

  DBuf<int> todoA, todoB, ntodo_d(1), prio2col;
  const int *todo = nullptr;
  int ntodo = nrows;
  long long guess = 0;
  {
    // initial slab guess: 4x the source rows + 1 per row, refined exactly on overflow
    long long src = 0;
    if (B.rows == nullptr)
      src = 0;  // unknown without a fetch; use a flat guess
    guess = 64LL * nrows + 4096;
  }
  DBuf<int> gkeys;
  DBuf<uint32_t> gvals;
  DBuf<unsigned> gbitmap;
  int nprio = 0;

  // global tier state (built on first use)
  DBuf<int> col2sid, sid2col, isfree, gtouched;
  DBuf<long long> fpos;
  DBuf<PRow> prow;
  DBuf<int2> G2;
  DBuf<unsigned long long> gsum, ext(2);
  DBuf<unsigned> gpending;
  int gslots = 0;
  // rows that overflow tier 0: SPASM_B200_SCHUR_DENSE=0 always the global tier (row at a time); default (1): batches of
  // at least DENSE_MIN rows of an x.U = b solve go through the column-major SpTRSM engine all at once — rows that reach
  // hundreds of pivots are a dependent chain of L2 round trips in any row-at-a-time engine
  static const int dense_policy = getenv("SPASM_B200_SCHUR_DENSE") ? atoi(getenv("SPASM_B200_SCHUR_DENSE")) : 1;
  constexpr int DENSE_MIN = 128;

  // Stages.  SMEM: warp per row, shared-memory hash (rows with at most ~400 touched columns).  Rows that overflow it:
  //  * x.U = b solves (Schur complement, GPLU batches): at least DENSE_MIN rows go to the column-major SpTRSM engine
  //    all at once; fewer rows go through the GLOBAL tier with a step budget, and rows whose elimination chain is
  //    longer than the budget (thousands of dependent steps) are handed to the SpTRSM engine as well;
  //  * other systems (kernel, rref, L solves): GLOBAL tier without a budget;
  //  * structural mode (triangular-solve ABI): the CTA-per-row heavy kernel.
  enum Stage { SMEM, GLOBAL, HEAVY };
  const bool dense_ok = dense_policy == 1 && G.U_dense != nullptr && E.prefix_col == nullptr && B.mask == nullptr && !E.structural && !E.all_columns;
  constexpr long long GLOBAL_BUDGET = 1 << 16;  // entries of G merged into one row (a few milliseconds of dependent steps)

  // one kernel stage over the current todo list, with exact-slab retries; returns with status[] final for the stage
  auto run_stage = [&](Stage stage, long long budget) {
    long long need = guess, lneed = guess;
    if (stage == GLOBAL && !E.count_only) {
      // a row that does not fit the slab is eliminated AGAIN with an exact slab: rows of this stage are the expensive ones,
      // so give the first attempt room for the average case seen so far times a safety factor, within a quarter of the free memory
      const long long cap_entries = (long long)(dev_free_bytes() / 4 / 8);
      need = std::min(std::max<long long>(guess, 16384LL * ntodo), std::max<long long>(cap_entries, guess));
      lneed = need;
    }
    for (int attempt = 0; attempt < 4 && ntodo > 0; attempt++) {
      unsigned long long h_ctrs[8];
      if (!E.count_only) {
        if (slabs_j.size() >= 60) throw Error("solve_rows: too many slabs");
        slabs_j.emplace_back((size_t)need);
        slabs_x.emplace_back((size_t)need);
        a.oj = slabs_j.back().p, a.ox = slabs_x.back().p, a.cap = need;
        a.slab_id = slabs_j.size() - 1;
        SP.j[a.slab_id] = a.oj, SP.x[a.slab_id] = a.ox;
        if (E.want_L) {
          slabs_lj.emplace_back((size_t)lneed);
          slabs_lx.emplace_back((size_t)lneed);
          a.lj = slabs_lj.back().p, a.lx = slabs_lx.back().p, a.lcap = lneed;
          LSP.j[a.slab_id] = a.lj, LSP.x[a.slab_id] = a.lx;
        }
      }
      CK(cudaMemsetAsync(ctrs.p + 2, 0, 6 * sizeof(unsigned long long), stream()));
      counter.zero();
      a.todo = todo, a.ntodo = ntodo;
      const int sms = sm_count();
      if (stage == SMEM) {
        constexpr int H = 512, PC = 256;
        size_t smem = 8 * (size_t)(2 * H + 3 * PC) * 4 + 32 * 8 + 64 * 4;
        int blocks = std::min(cdiv(ntodo, 8), sms * 4);
        if (F.small) {
          CK(cudaFuncSetAttribute(k_solve_smem<32, H, PC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_solve_smem<32, H, PC, true><<<blocks, 256, smem, stream()>>>(a);
        } else {
          CK(cudaFuncSetAttribute(k_solve_smem<32, H, PC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_solve_smem<32, H, PC, false><<<blocks, 256, smem, stream()>>>(a);
        }
        if (attempt == 0) R.stats.light += ntodo;
      } else if (stage == GLOBAL) {
        if (gslots == 0) {
          // slot ids, the relabelled copy of G, and the per-warp accumulators
          ext.zero();
          isfree.alloc(G.width + 1);
          fpos.alloc(G.width + 1);
          k_sys_extent<<<cdiv(G.width + 1, 256), 256, 0, stream()>>>(G.pdesc, G.width, ext.p, isfree.p);
          exclusive_scan_i32_to_i64(isfree.p, fpos.p, G.width + 1);
          unsigned long long hx[2];
          ext.download(hx, 2);
          const long long nfree = fetch(fpos.p + G.width);
          nprio = (int)hx[0];
          const long long gnnz = (long long)hx[1];
          a.nprio = nprio;
          a.gw = nprio + (int)nfree;
          a.nwords = (nprio + 31) / 32 + 1;
          col2sid.alloc(G.width);
          sid2col.alloc(std::max(a.gw, 1));
          sid2col.fill_ff();
          prow.alloc(std::max(nprio, 1));
          k_build_sid<<<cdiv(G.width, 256), 256, 0, stream()>>>(G.pdesc, fpos.p, G.width, nprio, col2sid.p, sid2col.p, prow.p);
          G2.alloc(std::max<long long>(gnnz, 1));
          if (gnnz) k_relabel_G2<<<cdiv(gnnz, 256), 256, 0, stream()>>>(G.Gj, G.Gx, gnnz, col2sid.p, G2.p);
          const size_t per = (size_t)a.gw * 12 + (size_t)a.nwords * 4;
          long long slots = std::min<long long>(((long long)ntodo + 3) / 4 * 4, 16LL * sms);
          const size_t budget_bytes = dev_free_bytes() / 3;
          while (slots > 4 && (size_t)slots * per > budget_bytes) slots = (slots / 2 + 3) / 4 * 4;
          gslots = (int)slots;
          gsum.alloc((size_t)gslots * a.gw);
          gsum.zero();
          gtouched.alloc((size_t)gslots * a.gw);
          gpending.alloc((size_t)gslots * a.nwords);
          gpending.zero();
          a.G2 = G2.p, a.col2sid = col2sid.p, a.sid2col = sid2col.p, a.prow = prow.p, a.classbit = G.classbit;
          a.gsum = gsum.p, a.gtouched = gtouched.p, a.gpending = gpending.p;
          CK(cudaGetLastError());
          g_launches += 4;
        }
        a.pop_budget = budget;
        const int blocks = std::min(gslots / 4, cdiv(ntodo, 4));
        if (F.small)
          k_solve_global<true><<<blocks, 128, 0, stream()>>>(a);
        else
          k_solve_global<false><<<blocks, 128, 0, stream()>>>(a);
        if (!E.count_only)
          k_sort_rows<<<std::min(ntodo, sms * 8), 256, 0, stream()>>>(todo, ntodo, status.p, R.cnt.p, off.p, E.prefix_col ? 1 : 0, a.oj, a.ox, a.slab_id,
                                                                       E.want_L ? R.lcnt.p : nullptr, loff.p, a.lj, a.lx);
        g_launches += 1;
        if (attempt == 0) R.stats.medium += ntodo;
      } else {
        int blocks = std::min(ntodo, sms * 2);
        if (gkeys.p == nullptr) {
          nprio = G.width;  // upper bound on the number of pivots
          prio2col.alloc(nprio + 1);
          prio2col.fill_ff();
          k_prio2col<<<cdiv(G.width, 256), 256, 0, stream()>>>(G.pdesc, G.width, prio2col.p);
          size_t per = (size_t)G.width;
          size_t avail = dev_free_bytes();
          while (blocks > 1 && (size_t)blocks * per * 8 > avail / 2) blocks /= 2;
          gkeys.alloc((size_t)blocks * per);
          gkeys.fill_ff();
          gvals.alloc((size_t)blocks * per);
          gbitmap.alloc((size_t)blocks * ((nprio + 31) / 32 + 1));
          gbitmap.zero();
        } else {
          blocks = std::min<long long>(blocks, (long long)(gkeys.n / (size_t)G.width));
        }
        a.gkeys = gkeys.p, a.gvals = gvals.p, a.gbitmap = gbitmap.p, a.nprio = nprio;
        if (F.small)
          k_solve_heavy<256, true><<<blocks, 256, 0, stream()>>>(a, prio2col.p);
        else
          k_solve_heavy<256, false><<<blocks, 256, 0, stream()>>>(a, prio2col.p);
        if (attempt == 0) R.stats.heavy += ntodo;
      }
      CK(cudaGetLastError());
      g_launches += 1;
      ctrs.download(h_ctrs, 8);
      sync();
      const long long n_slab = (long long)h_ctrs[3];
      if (n_slab == 0) break;
      // rows that did not fit the slab are retried in this stage with an exactly sized slab (rows that overflowed
      // the stage itself keep status ST_TABLE and are collected by the caller)
      DBuf<int> &dst = (todo == todoA.p) ? todoB : todoA;
      dst.alloc(ntodo);
      ntodo_d.zero();
      k_collect_status<<<cdiv(ntodo, 256), 256, 0, stream()>>>(status.p, todo, ntodo, ST_SLAB, dst.p, ntodo_d.p);
      todo = dst.p;
      ntodo = fetch(ntodo_d.p);
      need = (long long)h_ctrs[4] + 16;
      lneed = (long long)h_ctrs[5] + 16;
      if (attempt == 3 && ntodo > 0) throw Error("solve_rows: slab retries exhausted");
    }
  };
  auto collect_overflow = [&]() {
    DBuf<int> &dst = (todo == todoA.p) ? todoB : todoA;
    dst.alloc(nrows);
    ntodo_d.zero();
    k_collect_status<<<cdiv(nrows, 256), 256, 0, stream()>>>(status.p, nullptr, nrows, ST_TABLE, dst.p, ntodo_d.p);
    todo = dst.p;
    ntodo = fetch(ntodo_d.p);
    if (ntodo > 0) guess = std::max<long long>(guess, 1024LL * ntodo);
  };
  auto run_dense = [&]() {
    // ---- rows of an x.U = b solve, all at once through the SpTRSM engine, in chunks that fit
    const DCsr &U = *G.U_dense;
    const int Sm0 = G.width - U.n, r = U.n;
    DBuf<int> rows_sel(ntodo);
    k_gather_rows_sel<<<cdiv(ntodo, 256), 256, 0, stream()>>>(B.rows, todo, ntodo, rows_sel.p);
    const size_t per_row = ((size_t)Sm0 + (size_t)std::max(r, 1)) * 4 + 64;
    long long ch = (long long)(dev_free_bytes() * 2 / 5 / per_row);
    ch = std::max<long long>(256, std::min<long long>(ch / 256 * 256, ntodo));
    for (int off2 = 0; off2 < ntodo; off2 += (int)ch) {
      const int nr = (int)std::min<long long>(ch, ntodo - off2);
      DenseSchur D;
      build_dense_schur_raw(B.Bp, B.Bj, B.Bx, G.width, rows_sel.p + off2, nr, U, G.qinv_dense, F, D, E.want_L, ctrs.p);
      if (E.count_only) {
        DBuf<int> oj0;
        DBuf<uint32_t> ox0;
        dense_rows_to_sparse(D.Dt.p, D.ld, D.Sm0, D.q0.p, nr, todo, off2, R.cnt.p, off.p, 0, oj0, ox0);
      } else {
        if (slabs_j.size() >= 60) throw Error("solve_rows: too many slabs");
        slabs_j.emplace_back();
        slabs_x.emplace_back();
        const unsigned long long sid = slabs_j.size() - 1;
        if (E.want_L) slabs_lj.emplace_back(), slabs_lx.emplace_back();
        dense_rows_to_sparse(D.Dt.p, D.ld, D.Sm0, D.q0.p, nr, todo, off2, R.cnt.p, off.p, sid << 56, slabs_j.back(), slabs_x.back());
        SP.j[sid] = slabs_j.back().p, SP.x[sid] = slabs_x.back().p;
        if (E.want_L) {
          dense_rows_to_sparse(D.Vp.p, D.ldv, r, nullptr, nr, todo, off2, R.lcnt.p, loff.p, sid << 56, slabs_lj.back(), slabs_lx.back());
          LSP.j[sid] = slabs_lj.back().p, LSP.x[sid] = slabs_lx.back().p;
        }
      }
      // input and output rows of these eliminations: 8 nnz(A_k) + 8 and 8 nnz(S_k) + 8 bytes
      k_rows_io_bytes<<<cdiv(nr, 256), 256, 0, stream()>>>(B.Bp, rows_sel.p + off2, todo, off2, nr, R.cnt.p, ctrs.p);
      R.stats.heavy += nr;
      g_launches += 9;
    }
    ntodo = 0;
  };

  run_stage(SMEM, 0);
  collect_overflow();
  if (ntodo > 0) {
    if (E.structural) {
      run_stage(HEAVY, 0);
      collect_overflow();
    } else if (dense_ok && ntodo >= DENSE_MIN && !B.few_pivots) {
      run_dense();
    } else {
      run_stage(GLOBAL, dense_ok ? GLOBAL_BUDGET : 0);
      collect_overflow();
      if (ntodo > 0 && dense_ok) run_dense();
    }
  }
  if (ntodo > 0) throw Error("solve_rows: rows left unsolved after the last tier");

  unsigned long long h_ctrs[8];
  ctrs.download(h_ctrs, 8);
  exclusive_scan_i32_to_i64(R.cnt.p, R.p.p, nrows + 1);
  if (E.want_L) exclusive_scan_i32_to_i64(R.lcnt.p, R.lp.p, nrows + 1);
  R.nnz = fetch(R.p.p + nrows);
  R.stats.bytes = (long long)h_ctrs[0];
  R.stats.macs = (long long)h_ctrs[1];
  if (!E.count_only) {
    R.j.alloc(R.nnz);
    R.x.alloc(R.nnz);
    k_gather_rows<<<cdiv((long long)nrows * 32, 256), 256, 0, stream()>>>(R.cnt.p, off.p, R.p.p, nrows, SP, R.j.p, R.x.p);
    if (E.want_L) {
      R.lnnz = fetch(R.lp.p + nrows);
      R.lj.alloc(R.lnnz);
      R.lx.alloc(R.lnnz);
      k_gather_rows<<<cdiv((long long)nrows * 32, 256), 256, 0, stream()>>>(R.lcnt.p, loff.p, R.lp.p, nrows, LSP, R.lj.p, R.lx.p);
    }
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ev1, stream()));
  sync();
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ev0, ev1));
  R.stats.ms = ms;
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
}

// ------------------------------------------------------------------ rows split over the ranks
// The rows of a Schur complement / kernel basis / rref / batch of right-hand sides are independent given the
// system, which every rank holds.  Rank r solves the contiguous share [r*per, (r+1)*per) of the row list; the
// ranks then exchange the counts (one all-gather) and the entries (one broadcast per rank, straight into the
// final arrays at that rank's offset).  The result is the single-GPU result on every rank, bit for bit:
// rows in input order, each row computed by exactly one rank with the same kernels.
long long g_shard_stats[2] = {0, 0};
__global__ void k_iota_from(int *a, int n, int first) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = first + i;
}
__global__ void k_pick_ll(const long long *__restrict__ p, const int *__restrict__ idx, int n, long long *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = p[idx[i]];
}
// one stream of the result (entries or multipliers): counts -> row pointers -> payload
static long long gather_stream(const DBuf<int> &cnt_loc, const DBuf<int> &j_loc, const DBuf<uint32_t> &x_loc, int nloc, int nrows, int per,
                               DBuf<int> &cnt, DBuf<long long> &p, DBuf<int> &j, DBuf<uint32_t> &x, bool payload) {
  const Dist &dd = dist();
  const int NR = dd.nranks, me = dd.rank;
  cudaStream_t s = stream();
  DBuf<int> send(per), all((size_t)per * NR + 1);
  send.zero();
  if (nloc) CK(cudaMemcpyAsync(send.p, cnt_loc.p, (size_t)nloc * sizeof(int), cudaMemcpyDeviceToDevice, s));
  dist_allgather(send.p, all.p, (size_t)per * sizeof(int));
  cnt.alloc(nrows + 1);
  CK(cudaMemcpyAsync(cnt.p, all.p, (size_t)nrows * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(cnt.p + nrows, 0, sizeof(int), s));
  p.alloc(nrows + 1);
  exclusive_scan_i32_to_i64(cnt.p, p.p, nrows + 1);
  // where each rank's share starts in the payload
  std::vector<int> bidx(NR + 1);
  for (int r = 0; r <= NR; r++) bidx[r] = (int)std::min<long long>((long long)r * per, nrows);
  DBuf<int> dbidx(NR + 1);
  DBuf<long long> dboff(NR + 1);
  dbidx.upload(bidx.data(), NR + 1);
  k_pick_ll<<<1, 64, 0, s>>>(p.p, dbidx.p, NR + 1, dboff.p);
  std::vector<long long> boff(NR + 1);
  dboff.download(boff.data(), NR + 1);
  sync();
  const long long nnz = boff[NR];
  if (!payload) return nnz;
  j.alloc(nnz);
  x.alloc(nnz);
  const long long mine = boff[me + 1] - boff[me];
  if (mine) {
    CK(cudaMemcpyAsync(j.p + boff[me], j_loc.p, (size_t)mine * sizeof(int), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(x.p + boff[me], x_loc.p, (size_t)mine * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  }
  for (int r = 0; r < NR; r++) {
    const long long len = boff[r + 1] - boff[r];
    if (len == 0) continue;
    dist_broadcast(j.p + boff[r], (size_t)len * sizeof(int), r);
    dist_broadcast(x.p + boff[r], (size_t)len * sizeof(uint32_t), r);
  }
  return nnz;
}

void solve_rows(const SolveSystem &G, const SolveRows &B, const SolveEmit &E, const Fp &F, SolveResult &R) {
  const Dist &dd = dist();
  static const bool off = getenv("SPASM_B200_NO_SHARD_ROWS") != nullptr;
  if (!B.collective || dd.nranks <= 1 || off || E.count_only || E.structural || B.nrows < 2 * dd.nranks) {
    solve_rows_local(G, B, E, F, R);
    return;
  }
  const int NR = dd.nranks, me = dd.rank, nrows = B.nrows;
  cudaStream_t s = stream();
  cudaEvent_t ev0, ev1;
  CK(cudaEventCreate(&ev0));
  CK(cudaEventCreate(&ev1));
  CK(cudaEventRecord(ev0, s));
  const int per = (nrows + NR - 1) / NR;
  long long lo_, hi_;
  row_share(nrows, NR, me, &lo_, &hi_);
  const int lo = (int)lo_, hi = (int)hi_, nloc = hi - lo;
  SolveRows Bl = B;
  SolveEmit El = E;
  DBuf<int> iota;
  if (B.rows)
    Bl.rows = B.rows + lo;
  else {
    iota.alloc(std::max(nloc, 1));
    if (nloc) k_iota_from<<<cdiv(nloc, 256), 256, 0, s>>>(iota.p, nloc, lo);
    Bl.rows = iota.p;
  }
  Bl.nrows = nloc;
  Bl.collective = false;
  if (B.mask) Bl.mask = B.mask + lo;
  if (E.prefix_col) El.prefix_col = E.prefix_col + lo;
  SolveResult Rl;
  solve_rows_local(G, Bl, El, F, Rl);
  g_shard_stats[0] += 1, g_shard_stats[1] += nloc;
  R.nnz = gather_stream(Rl.cnt, Rl.j, Rl.x, nloc, nrows, per, R.cnt, R.p, R.j, R.x, true);
  R.lnnz = 0;
  if (E.want_L) R.lnnz = gather_stream(Rl.lcnt, Rl.lj, Rl.lx, nloc, nrows, per, R.lcnt, R.lp, R.lj, R.lx, true);
  // work counters of the whole call: the sum over the ranks
  DBuf<unsigned long long> w(6);
  const unsigned long long hw[6] = {(unsigned long long)Rl.stats.bytes, (unsigned long long)Rl.stats.macs, (unsigned long long)Rl.stats.rows,
                                    (unsigned long long)Rl.stats.light, (unsigned long long)Rl.stats.medium, (unsigned long long)Rl.stats.heavy};
  w.upload(hw, 6);
  dist_allreduce_sum_u64(w.p, 6);
  unsigned long long tot[6];
  w.download(tot, 6);
  CK(cudaEventRecord(ev1, s));
  sync();
  R.stats.bytes = (long long)tot[0], R.stats.macs = (long long)tot[1], R.stats.rows = (long long)tot[2];
  R.stats.light = (long long)tot[3], R.stats.medium = (long long)tot[4], R.stats.heavy = (long long)tot[5];
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ev0, ev1));
  R.stats.ms = ms;
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
}

}  // namespace sb
