// pivots.cu — structural pivot search on the GPU (replaces spasm_pivots_extract_structural,
// prototype src/SpaSM.jl:775-778; phases README.md:21-24; algorithm SURVEY.md A.4).
//
// All three searches reproduce the SEQUENTIAL (single-thread libspasm) result exactly:
//  * Faugère-Lachartre: per leading column arg-min of (weight,row) -> one 64-bit atomicMin per row.
//  * FL on columns: order-dependent greedy, run as deterministic reservation rounds: every
//    undecided row reserves its still-open columns with atomicMin(row id); a row that holds all
//    its reservations cannot be influenced by any smaller undecided row and decides exactly as the
//    sequential scan would.
//  * greedy alternating cycle-free search: windows of rows run their BFS speculatively in parallel
//    (one warp each, private generation-stamped marks), then one warp commits the window in row
//    order, CONTINUING the BFS of a row from the pivots committed before it in the same window
//    (reachability is monotone in the pivot set, so the continued closure equals the sequential
//    one; a row that failed speculatively stays failed).
//  * reorder: heights in the pivot DAG by relaxation, then a stable sort (normalisation N2).
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include "pivots.cuh"

namespace sb {

// ------------------------------------------------------------------ Faugère-Lachartre
__global__ void k_fl_scan(const long long *__restrict__ Ap, const int *__restrict__ Aj, int n,
                          unsigned long long *__restrict__ best) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  long long a = Ap[i], b = Ap[i + 1];
  if (a == b) return;
  int mn = 0x7fffffff;
  for (long long e = a + lane; e < b; e += 32) mn = min(mn, Aj[e]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if (lane == 0) atomicMin(&best[mn], ((unsigned long long)(unsigned)(b - a) << 32) | (unsigned)i);
}
__global__ void k_fl_commit(const unsigned long long *__restrict__ best, int m, int *__restrict__ pinv,
                            int *__restrict__ qinv, int *__restrict__ npiv) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  unsigned long long b = best[j];
  if (b == ~0ULL) return;
  int i = (int)(b & 0xffffffffu);
  pinv[i] = j;
  qinv[j] = i;
  atomicAdd(npiv, 1);
}

// ------------------------------------------------------------------ FL on columns
__global__ void k_flc_close_pivot_rows(const long long *__restrict__ Ap, const int *__restrict__ Aj, int n,
                                       const int *__restrict__ pinv, unsigned char *__restrict__ open) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n || pinv[i] < 0) return;
  for (long long e = Ap[i] + lane; e < Ap[i + 1]; e += 32) open[Aj[e]] = 0;
}
// state[i]: 0 undecided, 1 decided
__global__ void k_flc_reserve(const long long *__restrict__ Ap, const int *__restrict__ Aj, int n,
                              const int *__restrict__ pinv, const int *__restrict__ qinv,
                              const unsigned char *__restrict__ open, unsigned char *__restrict__ state,
                              unsigned long long *__restrict__ res, unsigned round) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n || pinv[i] >= 0 || state[i]) return;
  const unsigned long long key = ((unsigned long long)(~round) << 32) | (unsigned)i;
  int any = 0;
  for (long long e = Ap[i] + lane; e < Ap[i + 1]; e += 32) {
    int j = Aj[e];
    if (open[j] && qinv[j] < 0) {
      atomicMin(&res[j], key);
      any = 1;
    }
  }
  any = __any_sync(0xffffffffu, any);
  if (!any && lane == 0) state[i] = 1;  // nothing open on this row, and columns never reopen
}
__global__ void k_flc_commit(const long long *__restrict__ Ap, const int *__restrict__ Aj, int n, int *__restrict__ pinv,
                             int *__restrict__ qinv, unsigned char *__restrict__ open, unsigned char *__restrict__ state,
                             const unsigned long long *__restrict__ res, unsigned round, int *__restrict__ counters) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n || pinv[i] >= 0 || state[i]) return;
  const unsigned long long key = ((unsigned long long)(~round) << 32) | (unsigned)i;
  const long long a = Ap[i], b = Ap[i + 1];
  int ready = 1;
  long long first = -1;  // position of the first eligible entry in storage order
  for (long long e0 = a; e0 < b; e0 += 32) {
    long long e = e0 + lane;
    int elig = 0, mine = 1;
    if (e < b) {
      int j = Aj[e];
      if (open[j] && qinv[j] < 0) {
        elig = 1;
        mine = (res[j] == key);
      }
    }
    unsigned em = __ballot_sync(0xffffffffu, elig);
    if (!__all_sync(0xffffffffu, mine)) ready = 0;
    if (first < 0 && em) first = e0 + (__ffs(em) - 1);
  }
  if (!ready) {
    if (lane == 0) counters[1] = 1;  // someone is still undecided
    return;
  }
  if (first < 0) {
    if (lane == 0) state[i] = 1;
    return;
  }
  const int jp = Aj[first];
  __syncwarp();
  for (long long e = a + lane; e < b; e += 32) open[Aj[e]] = 0;
  if (lane == 0) {
    pinv[i] = jp;
    qinv[jp] = i;
    state[i] = 1;
    atomicAdd(&counters[0], 1);
  }
}

// ------------------------------------------------------------------ greedy cycle-free search
// Closure marks of one row: TWO BITS per column, 16 columns per 32-bit word, private to the row's warp and
// cleared by it before use (m/4 bytes: 50 KB for 200 000 columns, so the marks of a whole window stay in
// or near L2 instead of being 2 bytes per column of random HBM traffic).
//   MV: the column has been reached by the row's closure;  MC: the column is one of the row's candidates.
// One atomicOr both tests and sets MV and settles duplicates inside a warp step.
static constexpr unsigned MV = 1u, MC = 2u;

struct GreedyArgs {
  const long long *Ap;
  const int *Aj;
  int n, m;
  int *pinv, *qinv;
  const int *cand;  // non-pivotal rows, increasing
  int ncand;
  unsigned *marks;  // [W][mw] packed 2-bit marks
  int mw;           // words per row of marks (multiple of 4)
  int *queue;       // [W][qcap]
  int qcap;
  int *done;        // [ncand] row is decided
  int *ctl;         // one 64-bit word: cursor << 32 | pivots published
  int *next;        // next candidate row to hand out
  int *newcol;      // [ncand] columns of the pivots published, in order
  int *npiv;
  unsigned long long *prof;
};

__device__ __forceinline__ unsigned mark_or(unsigned *marks, int j, unsigned bits) {
  const int sh = (j & 15) * 2;
  return (atomicOr(&marks[j >> 4], bits << sh) >> sh) & 3u;
}
__device__ __forceinline__ unsigned mark_get(const unsigned *marks, int j) { return (__ldcg(&marks[j >> 4]) >> ((j & 15) * 2)) & 3u; }

// the queue holds each pivotal column once, except when a pivot published by another row is met by the
// closure and then applied again from the published list (harmless, rare): slide the live part down
// instead of sizing the queue for that
__device__ __forceinline__ void queue_room(const GreedyArgs &g, int *q, int &head, int &tail, int lane) {
  if (tail + 32 <= g.qcap) return;
  const int live = tail - head;
  for (int k0 = 0; k0 < live; k0 += 32) {
    const int k = k0 + lane;
    int v = 0;
    if (k < live) v = q[head + k];
    __syncwarp();
    if (k < live) q[k] = v;
  }
  __syncwarp();
  head = 0, tail = live;
  if (tail + 32 > g.qcap) __trap();
}

// Expand queued pivotal columns until the queue is empty or no candidate survives.  Up to 32 queue
// entries are expanded together: lane l owns entry head+l, and the entries of those pivot rows are
// walked as one flattened index space, 32 at a time (full lanes even though rows are short).  The
// column index of the NEXT chunk is fetched while the current one is resolved; per chunk the mark
// atomic and the qinv load issue together, so the dependent chain is one round trip.
__device__ __forceinline__ void bfs_run(const GreedyArgs &g, unsigned *marks, int *q, int &head, int &tail, int &surviving, int lane) {
  while (head < tail && surviving > 0) {
    const int nb = min(32, tail - head);
    long long a = 0;
    int len = 0;
    if (lane < nb) {
      const int I = __ldcg(&g.qinv[q[head + lane]]);
      if (I >= 0) {
        a = g.Ap[I];
        len = (int)(g.Ap[I + 1] - a);
      }
    }
    head += nb;
    // inclusive prefix of len over lanes
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    auto fetch_col = [&](int f0) -> int {
      const int f = f0 + lane;
      const int fc = min(f, total - 1);
      int lo = 0;
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        int v = __shfl_sync(0xffffffffu, inc, lo + step - 1);
        if (v <= fc) lo += step;
      }
      const int excl = __shfl_sync(0xffffffffu, inc - len, lo);
      const long long base = __shfl_sync(0xffffffffu, a, lo);
      return (f < total) ? g.Aj[base + (f - excl)] : -1;
    };
    int jj_next = total > 0 ? fetch_col(0) : -1;
    for (int f0 = 0; f0 < total; f0 += 32) {
      const int jj = jj_next;
      if (f0 + 32 < total) jj_next = fetch_col(f0 + 32);  // in flight while this chunk is resolved
      queue_room(g, q, head, tail, lane);
      int push = 0, killed = 0;
      if (jj >= 0) {
        const unsigned old = mark_or(marks, jj, MV);
        const bool piv = __ldcg(&g.qinv[jj]) >= 0;
        if (!(old & MV)) {
          killed = (old & MC) != 0;
          push = piv;
        }
      }
      unsigned pm = __ballot_sync(0xffffffffu, push);
      surviving -= __popc(__ballot_sync(0xffffffffu, killed));
      if (push) q[tail + __popc(pm & ((1u << lane) - 1u))] = jj;
      tail += __popc(pm);
      __syncwarp();
      if (surviving <= 0) break;
    }
    __syncwarp();
  }
}

// ONE cooperative launch, ONE WARP PER ROW AT A TIME.  The grid is W warps (all co-resident); each warp
// takes the next candidate row from a counter, so rows are handed out in increasing order and the W
// rows in flight form a window that slides as rows are decided (no barrier between windows: a slow
// closure delays nobody but its own successors in the commit order).  For its row a warp
//   1. runs the BFS against the pivots known when it starts;
//   2. waits for its turn: rows are decided strictly in row order.  Pivots committed meanwhile are
//      published in an append-only list; every waiting row keeps applying them to its own closure as
//      they appear (continuing its BFS), so when its turn comes only the last few remain.  A row whose
//      candidates are all reached is final at once (reachability only grows) and is skipped by the cursor.
// The row at the cursor is always in flight or the next to be handed out (rows are handed out in order),
// so the spin-waits cannot deadlock.
// ctl64: (cursor = rows decided) << 32 | pivots published     done[]: 1 final without pivot, 2 decided with one
__device__ void greedy_row(const GreedyArgs &g, const int t, const int wn, unsigned *marks, int *q, const int lane) {
  volatile int *done = g.done;
  const int i = g.cand[t];
  unsigned long long *ctl64 = (unsigned long long *)g.ctl;
  auto read_ctl = [&](int &cur, int &nn) {
    unsigned long long w = 0;
    if (lane == 0) w = *(volatile unsigned long long *)ctl64;
    w = __shfl_sync(0xffffffffu, w, 0);
    cur = (int)(w >> 32), nn = (int)(w & 0xffffffffu);
  };
  // pivots published before this point are in qinv for every load below (they are written before the count)
  int applied = 0;
  {
    int cur0;
    read_ctl(cur0, applied);
    __threadfence();
  }
  int head = 0, tail = 0, surviving = 0;
  {
    uint4 *mz = (uint4 *)marks;
    for (int k = lane; k < (g.mw >> 2); k += 32) __stcg(&mz[k], make_uint4(0u, 0u, 0u, 0u));
    __syncwarp();
    const long long a = g.Ap[i], b = g.Ap[i + 1];
    for (long long e0 = a; e0 < b; e0 += 32) {
      long long e = e0 + lane;
      int push = 0, isc = 0, j = 0;
      if (e < b) {
        j = g.Aj[e];
        if (__ldcg(&g.qinv[j]) < 0) {
          mark_or(marks, j, MC);
          isc = 1;
        } else {
          mark_or(marks, j, MV);
          push = 1;
        }
      }
      unsigned pm = __ballot_sync(0xffffffffu, push);
      surviving += __popc(__ballot_sync(0xffffffffu, isc));
      if (push) q[tail + __popc(pm & ((1u << lane) - 1u))] = j;
      tail += __popc(pm);
    }
    __syncwarp();
    unsigned long long tb = 0, te = 0;
    if (g.prof && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tb));
    bfs_run(g, marks, q, head, tail, surviving, lane);
    if (g.prof && lane == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(te));
      atomicMax(g.prof + 0, te - tb);                 // longest speculative BFS (ns)
      atomicAdd(g.prof + 1, te - tb);                 // sum over rows
      atomicAdd(g.prof + 2, (unsigned long long)tail);  // pivot rows expanded
      if (surviving > 0) atomicAdd(g.prof + 3, 1ULL);   // rows alive after speculation
    }
  }
  // first surviving candidate of the row, in storage order
  auto first_candidate = [&]() -> int {
    const long long a = g.Ap[i], b = g.Ap[i + 1];
    int jp = -1;
    for (long long e0 = a; e0 < b && jp < 0; e0 += 32) {
      long long e = e0 + lane;
      int j = e < b ? g.Aj[e] : -1;
      unsigned c = __ballot_sync(0xffffffffu, j >= 0 && mark_get(marks, j) == MC);
      if (c) jp = __shfl_sync(0xffffffffu, j, __ffs(c) - 1);
    }
    return jp;
  };
  int jp_ready = -2;  // -2: not computed for the current closure
  unsigned backoff = 32;
  for (;;) {
    int cur, nn;
    if (surviving <= 0) {  // final: no pivot on this row
      if (lane == 0) {
        done[t] = 1;
        __threadfence();
      }
      read_ctl(cur, nn);  // the cursor may be waiting on us
      if (cur != t) return;
      // fall through to advance the cursor below
    } else {
      read_ctl(cur, nn);
      if (applied < nn) {
        // apply the pivots published since last time: a column this closure has reached (it was not pivotal
        // then) or that is one of its candidates must now be expanded; one it has not reached will be
        // treated as pivotal when the closure gets there
        for (; applied < nn && surviving > 0; applied++) {
          const int jc = __ldcg(&g.newcol[applied]);
          const unsigned st = mark_get(marks, jc);
          if (st == 0) continue;
          if (st == MC) surviving--;
          queue_room(g, q, head, tail, lane);
          if (lane == 0) {
            mark_or(marks, jc, MV);
            q[tail] = jc;
          }
          tail++;
          __syncwarp();
          bfs_run(g, marks, q, head, tail, surviving, lane);
          jp_ready = -2;
        }
        backoff = 32;
        continue;  // re-read the control word
      }
      if (cur != t) {
        if (jp_ready == -2) {  // use the wait: have the answer ready when the turn comes
          jp_ready = first_candidate();
          continue;
        }
        __nanosleep(backoff);
        if (backoff < 512) backoff <<= 1;
        continue;
      }
    }
    // ---- my turn: every earlier row is final and all their pivots are applied (nn is final too,
    // because only the row at the cursor can publish)
    unsigned long long t_turn = 0;
    if (g.prof && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_turn));
    int npub = nn;
    if (surviving > 0) {
      const int jp = jp_ready != -2 ? jp_ready : first_candidate();
      npub = nn + 1;
      if (lane == 0) {
        g.pinv[i] = jp;
        g.qinv[jp] = i;
        g.newcol[nn] = jp;
        atomicAdd(g.npiv, 1);
        done[t] = 2;  // decided WITH a pivot (1: final without)
        __threadfence();
        // publish the count at once (cursor unchanged): the waiting rows start applying it while we look for the next row
        atomicMax(ctl64, ((unsigned long long)t << 32) | (unsigned)npub);
      }
    }
    // advance the cursor over every following row that is already final.  atomicMax keeps the word
    // monotone when a row that just became final advances it concurrently; the re-check after the
    // fence closes the missed-wakeup window (that row does store-done / fence / load-cursor).
    // A row seen with done == 2 has had its turn AFTER ours: the cursor is in other hands and our
    // pivot count is stale, so we stop at once instead of writing it back.
    int nxt = t + 1;
    for (;;) {
      bool stale = false;
      for (;;) {
        const int tt = nxt + lane;
        const int dv = tt < wn ? done[tt] : 0;
        const unsigned fin = __ballot_sync(0xffffffffu, dv != 0);
        const int run = (fin == 0xffffffffu) ? 32 : __ffs(~fin) - 1;
        const unsigned passed = run == 32 ? 0xffffffffu : ((1u << run) - 1u);
        if (__ballot_sync(0xffffffffu, dv == 2) & passed) {
          stale = true;
          break;
        }
        nxt += run;
        if (run < 32 || nxt >= wn) break;
      }
      if (stale) break;
      if (nxt > wn) nxt = wn;
      int cur2 = 0, d = 0;
      if (lane == 0) {
        const unsigned long long mine = ((unsigned long long)nxt << 32) | (unsigned)npub;
        const unsigned long long old = atomicMax(ctl64, mine);
        cur2 = (int)((old > mine ? old : mine) >> 32);
        __threadfence();
        if (cur2 < wn) d = done[cur2];
      }
      cur2 = __shfl_sync(0xffffffffu, cur2, 0);
      d = __shfl_sync(0xffffffffu, d, 0);
      if (cur2 >= wn || d != 1) break;  // handed over (0), or that row already took its turn (2)
      nxt = cur2;
    }
    if (g.prof && lane == 0) {
      unsigned long long t_end;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      atomicAdd(g.prof + 4, t_end - t_turn);  // time spent holding the cursor (serial part)
      atomicAdd(g.prof + 5, 1ULL);
    }
    // a row that became final between our scan and the store re-checks the cursor itself (top of loop)
    return;
  }
}

__global__ void __launch_bounds__(256) k_greedy_stream(GreedyArgs g) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned *marks = g.marks + (size_t)slot * g.mw;
  int *q = g.queue + (size_t)slot * g.qcap;
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(g.next, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= g.ncand) return;
    greedy_row(g, t, g.ncand, marks, q, lane);
    __syncwarp();
  }
}

// ------------------------------------------------------------------ reorder (heights) + extraction
// Heights in the pivot DAG (height = longest chain of pivot rows below a row) by peeling from the sinks: a row is
// final when every pivot row it references is final.  ONE CTA walks the levels; a level is usually a handful of
// rows (the DAG of the banded benchmark matrices is tens of thousands of levels deep, three rows wide), so the cost
// per level is a few dependent loads + a block barrier (~3 us) instead of a launch + host round trip per sweep.
// T = A^T gives, for the pivot column of a finished row, the rows that were waiting for it.
__global__ void k_out_degree(const long long *__restrict__ Ap, const int *__restrict__ Aj, const int *__restrict__ prow, int npiv,
                             const int *__restrict__ pinv, const int *__restrict__ qinv, int *__restrict__ deg, int *__restrict__ queue,
                             int *__restrict__ qtail) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= npiv) return;
  const int i = prow[k], jp = pinv[i];
  int d = 0;
  for (long long e = Ap[i] + lane; e < Ap[i + 1]; e += 32) {
    const int j = Aj[e];
    d += (j != jp && qinv[j] >= 0);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) {
    deg[i] = d;
    if (d == 0) queue[atomicAdd(qtail, 1)] = i;
  }
}
__global__ void __launch_bounds__(1024) k_height_peel(const long long *__restrict__ Tp, const int *__restrict__ Tj, const int *__restrict__ pinv,
                                                       int *__restrict__ deg, int *__restrict__ height, int *__restrict__ queue,
                                                       int *__restrict__ qtail /* in: size of level 0; out: rows finished */) {
  __shared__ int s_begin, s_end, s_tail;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_begin = 0, s_end = *qtail, s_tail = *qtail;
  __syncthreads();
  for (;;) {
    const int b = s_begin, e = s_end;
    if (b == e) break;
    for (int q = b + warp; q < e; q += nwarps) {
      const int i2 = ((volatile int *)queue)[q];
      const int h2 = ((volatile int *)height)[i2] + 1;
      const int c = pinv[i2];
      for (long long t = Tp[c] + lane; t < Tp[c + 1]; t += 32) {
        const int i = Tj[t];
        if (i == i2 || pinv[i] < 0) continue;
        atomicMax(&height[i], h2);
        if (atomicSub(&deg[i], 1) == 1) queue[atomicAdd(&s_tail, 1)] = i;
      }
    }
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x == 0) s_begin = e, s_end = s_tail;
    __syncthreads();
  }
  if (threadIdx.x == 0) *qtail = s_tail;
}
__global__ void k_flag_rows(const int *__restrict__ pinv, int n, int want_pivotal, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = ((pinv[i] >= 0) == (want_pivotal != 0));
  if (i == n) flag[i] = 0;
}
__global__ void k_compact_rows(const int *__restrict__ flag, const long long *__restrict__ pos, int n, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[pos[i]] = i;
}
__global__ void k_sort_keys(const int *__restrict__ rows, int npiv, const int *__restrict__ height, int maxh, unsigned *__restrict__ keys) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < npiv) keys[k] = (unsigned)(maxh - height[rows[k]]);
}
__global__ void k_row_lens(const long long *__restrict__ Ap, const int *__restrict__ p, int npiv, int *__restrict__ len) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < npiv) len[k] = (int)(Ap[p[k] + 1] - Ap[p[k]]);
  if (k == npiv) len[k] = 0;
}
template <bool SMALL>
__global__ void k_extract(const long long *__restrict__ Ap, const int *__restrict__ Aj, const uint32_t *__restrict__ Ax,
                          const int *__restrict__ p, const int *__restrict__ pinv, int npiv, const long long *__restrict__ pos,
                          long long ubase, int urow0, long long *__restrict__ Up, int *__restrict__ Uj,
                          uint32_t *__restrict__ Ux, int *__restrict__ Uqinv, uint32_t *__restrict__ pivval, Fp F) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= npiv) return;
  const int i = p[k], jp = pinv[i];
  const long long a = Ap[i], b = Ap[i + 1], dst = ubase + pos[k];
  // locate the pivot entry (first occurrence of column jp)
  long long pe = -1;
  for (long long e0 = a; e0 < b && pe < 0; e0 += 32) {
    long long e = e0 + lane;
    unsigned hit = __ballot_sync(0xffffffffu, e < b && Aj[e] == jp);
    if (hit) pe = e0 + (__ffs(hit) - 1);
  }
  const uint32_t pv = Ax[pe];
  const uint32_t alpha = dev_inv(pv, F.p);
  if (lane == 0) {
    Uj[dst] = jp;
    Ux[dst] = 1;
    Uqinv[jp] = urow0 + k;
    Up[urow0 + k + 1] = dst + (b - a);
    pivval[k] = pv;
  }
  for (long long e = a + lane; e < b; e += 32) {
    if (e == pe) continue;
    long long d = dst + 1 + (e - a) - (e > pe ? 1 : 0);
    Uj[d] = Aj[e];
    Ux[d] = mulmod<SMALL>(alpha, Ax[e], F);
  }
}

int find_structural_pivots(const DCsr &A, bool greedy, PivotSearch &P, int counts[3]) {
  const int n = A.n, m = A.m;
  cudaStream_t s = stream();
  P.pinv.alloc(n);
  P.qinv.alloc(m);
  P.p.alloc(n);
  P.pinv.fill_ff();
  P.qinv.fill_ff();
  counts[0] = counts[1] = counts[2] = 0;
  if (n == 0 || m == 0 || A.nnz == 0) {
    P.npiv = 0;
    if (n) {
      DBuf<int> flag(n + 1);
      DBuf<long long> pos(n + 1);
      k_flag_rows<<<cdiv(n + 1, 256), 256, 0, s>>>(P.pinv.p, n, 0, flag.p);
      exclusive_scan_i32_to_i64(flag.p, pos.p, n + 1);
      k_compact_rows<<<cdiv(n, 256), 256, 0, s>>>(flag.p, pos.p, n, P.p.p);
    }
    return 0;
  }
  DBuf<int> ctr(4);
  ctr.zero();
  const int rowblocks = cdiv((long long)n * 32, 256);
  double tt = spasm_wtime();
  {  // FL
    DBuf<unsigned long long> best(m);
    best.fill_ff();
    k_fl_scan<<<rowblocks, 256, 0, s>>>(A.p.p, A.j.p, n, best.p);
    k_fl_commit<<<cdiv(m, 256), 256, 0, s>>>(best.p, m, P.pinv.p, P.qinv.p, ctr.p);
    CK(cudaGetLastError());
    counts[0] = fetch(ctr.p);
  }
  P.t_fl = spasm_wtime() - tt, tt = spasm_wtime();
  {  // FL on columns
    DBuf<unsigned char> open(m), state(n);
    DBuf<unsigned long long> res(m);
    CK(cudaMemsetAsync(open.p, 1, m, s));
    state.zero();
    res.fill_ff();
    k_flc_close_pivot_rows<<<rowblocks, 256, 0, s>>>(A.p.p, A.j.p, n, P.pinv.p, open.p);
    ctr.zero();
    for (unsigned round = 0;; round++) {
      CK(cudaMemsetAsync(ctr.p + 1, 0, sizeof(int), s));
      k_flc_reserve<<<rowblocks, 256, 0, s>>>(A.p.p, A.j.p, n, P.pinv.p, P.qinv.p, open.p, state.p, res.p, round);
      k_flc_commit<<<rowblocks, 256, 0, s>>>(A.p.p, A.j.p, n, P.pinv.p, P.qinv.p, open.p, state.p, res.p, round, ctr.p);
      CK(cudaGetLastError());
      P.flcol_rounds = (int)round + 1;
      g_launches += 2;
      if (fetch(ctr.p + 1) == 0) break;
    }
    counts[1] = fetch(ctr.p);
  }
  P.t_flcol = spasm_wtime() - tt, tt = spasm_wtime();
  if (greedy) {
    DBuf<int> flag(n + 1), cand;
    DBuf<long long> pos(n + 1);
    k_flag_rows<<<cdiv(n + 1, 256), 256, 0, s>>>(P.pinv.p, n, 0, flag.p);
    exclusive_scan_i32_to_i64(flag.p, pos.p, n + 1);
    const int ncand = (int)fetch(pos.p + n);
    if (ncand > 0) {
      cand.alloc(ncand);
      k_compact_rows<<<cdiv(n, 256), 256, 0, s>>>(flag.p, pos.p, n, cand.p);
      {
      const int qcap = std::min(n, m) + 1 + 1024;
      const int mw = (((m + 15) >> 4) + 3) & ~3;
      const size_t per = (size_t)mw * 4 + (size_t)qcap * 4;
      size_t budget = std::min<size_t>(dev_free_bytes() / 3, (size_t)24 << 30);
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_greedy_stream, 256, 0));
      const int grid = std::max(1, occ) * sm_count();
      // rows in flight: enough warps to keep the memory system busy, few enough that their marks
      // (mw words each) stay close to the 126 MB L2 — measured at 200 000 columns: 16 warps per SM beat 32 and 48
      int W = std::min(grid * 8, sm_count() * 16);
      W = (int)std::max<size_t>(32, std::min<size_t>((size_t)W, budget / per));
      if (const char *e = getenv("SPASM_B200_GREEDY_W")) W = std::max(32, std::min(grid * 8, atoi(e)));
      W = std::max(8, std::min(W, (ncand + 7) / 8 * 8));
      DBuf<unsigned> marks((size_t)W * mw);
      DBuf<int> queue((size_t)W * qcap), done(ncand), ctl(4), next(1), newcol(ncand);
      ctr.zero(), ctl.zero(), next.zero(), done.zero();
      DBuf<unsigned long long> prof(8);
      prof.zero();
      GreedyArgs g{A.p.p, A.j.p, n, m, P.pinv.p, P.qinv.p, cand.p, ncand, marks.p, mw, queue.p, qcap, done.p, ctl.p, next.p, newcol.p, ctr.p, prof.p};
      {
        void *args[] = {&g};
        CK(cudaLaunchCooperativeKernel((void *)k_greedy_stream, dim3(W / 8), dim3(256), args, 0, s));
        g_launches += 1;
        P.greedy_windows = W;
      }
      CK(cudaGetLastError());
      {
        unsigned long long hp[8];
        prof.download(hp, 8);
        sync();
        if (getenv("SPASM_B200_PROFILE"))
          fprintf(stderr, "[greedy] rows in flight=%d W=%d rows=%d  longest speculative BFS %.1f ms, mean %.3f ms, pivot rows expanded %.3g, alive after speculation %llu; "
                          "cursor held %.3f s over %llu turns (%.1f us / turn)\n",
                  P.greedy_windows, W, ncand, hp[0] * 1e-6, hp[1] * 1e-6 / std::max(ncand, 1), (double)hp[2], hp[3], hp[4] * 1e-9, hp[5],
                  hp[5] ? hp[4] * 1e-3 / hp[5] : 0.0);
      }
      }
      counts[2] = fetch(ctr.p);
    }
  }
  const int npiv = counts[0] + counts[1] + counts[2];
  P.npiv = npiv;
  P.t_greedy = spasm_wtime() - tt, tt = spasm_wtime();
  // ---- reorder: p[0:npiv] by (height desc, row asc), p[npiv:n] the other rows increasing
  {
    DBuf<int> flag(n + 1), prow(std::max(npiv, 1));
    DBuf<long long> pos(n + 1);
    k_flag_rows<<<cdiv(n + 1, 256), 256, 0, s>>>(P.pinv.p, n, 1, flag.p);
    exclusive_scan_i32_to_i64(flag.p, pos.p, n + 1);
    k_compact_rows<<<cdiv(n, 256), 256, 0, s>>>(flag.p, pos.p, n, prow.p);
    k_flag_rows<<<cdiv(n + 1, 256), 256, 0, s>>>(P.pinv.p, n, 0, flag.p);
    exclusive_scan_i32_to_i64(flag.p, pos.p, n + 1);
    k_compact_rows<<<cdiv(n, 256), 256, 0, s>>>(flag.p, pos.p, n, P.p.p + npiv);
    if (npiv > 0) {
      DBuf<int> height(n);
      height.zero();
      {
        DCsr T;
        transpose_csr(A, T);
        DBuf<int> deg(n), queue(npiv), qtail(1);
        qtail.zero();
        k_out_degree<<<cdiv((long long)npiv * 32, 256), 256, 0, s>>>(A.p.p, A.j.p, prow.p, npiv, P.pinv.p, P.qinv.p, deg.p, queue.p, qtail.p);
        k_height_peel<<<1, 1024, 0, s>>>(T.p.p, T.j.p, P.pinv.p, deg.p, height.p, queue.p, qtail.p);
        CK(cudaGetLastError());
        g_launches += 2;
        if (fetch(qtail.p) != npiv) throw Error("structural pivots contain a cycle");
      }
      // max height
      DBuf<int> mx(1);
      size_t tmp = 0;
      cub::DeviceReduce::Max(nullptr, tmp, height.p, mx.p, n, s);
      DBuf<char> t1(tmp);
      cub::DeviceReduce::Max(t1.p, tmp, height.p, mx.p, n, s);
      const int maxh = fetch(mx.p);
      DBuf<unsigned> keys(npiv), keys2(npiv);
      k_sort_keys<<<cdiv(npiv, 256), 256, 0, s>>>(prow.p, npiv, height.p, maxh, keys.p);
      int bits = 1;
      while ((1LL << bits) <= maxh) bits++;
      tmp = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys.p, keys2.p, prow.p, P.p.p, npiv, 0, bits, s);
      DBuf<char> t2(tmp);
      cub::DeviceRadixSort::SortPairs(t2.p, tmp, keys.p, keys2.p, prow.p, P.p.p, npiv, 0, bits, s);
      P.max_height = maxh;
    }
    CK(cudaGetLastError());
  }
  sync();
  P.t_reorder = spasm_wtime() - tt;
  return npiv;
}

// append the normalised pivot rows p[0:npiv] of A to U (pivot entry first, the rest scaled, in
// storage order), set Uqinv, return the pivot values (for L)
void extract_pivot_rows(const DCsr &A, const PivotSearch &P, DCsr &U, DBuf<int> &Uqinv, const Fp &F, DBuf<uint32_t> &pivval) {
  const int npiv = P.npiv;
  pivval.alloc(std::max(npiv, 1));
  if (npiv == 0) return;
  cudaStream_t s = stream();
  DBuf<int> len(npiv + 1);
  DBuf<long long> pos(npiv + 1);
  k_row_lens<<<cdiv(npiv + 1, 256), 256, 0, s>>>(A.p.p, P.p.p, npiv, len.p);
  exclusive_scan_i32_to_i64(len.p, pos.p, npiv + 1);
  const long long add = fetch(pos.p + npiv);
  csr_reserve(U, U.nnz + add, U.n + npiv);
  if (F.small)
    k_extract<true><<<cdiv((long long)npiv * 32, 256), 256, 0, s>>>(A.p.p, A.j.p, A.x.p, P.p.p, P.pinv.p, npiv, pos.p, U.nnz, U.n, U.p.p,
                                                                  U.j.p, U.x.p, Uqinv.p, pivval.p, F);
  else
    k_extract<false><<<cdiv((long long)npiv * 32, 256), 256, 0, s>>>(A.p.p, A.j.p, A.x.p, P.p.p, P.pinv.p, npiv, pos.p, U.nnz, U.n, U.p.p,
                                                                   U.j.p, U.x.p, Uqinv.p, pivval.p, F);
  CK(cudaGetLastError());
  U.nnz += add;
  U.n += npiv;
}

}  // namespace sb
