// echelonize.cu — spasm_echelonize (src/SpaSM.jl:860-866) and its round loop on the GPU:
// structural pivots -> density estimate -> Schur complement, at most max_round times, then the
// dense or GPLU tail (README.md:19-38).  All matrices stay resident in HBM
// between phases; only scalars (counts, densities) and the final factor cross PCIe.
#include <algorithm>
#include <memory>

#include "dense.cuh"
#include "dist.cuh"
#include "pivots.cuh"
#include "sink.cuh"

namespace sb {

extern WorkStats g_last_stats;

static u64 g_seed_state = 0;
static const u64 SEED0 = 0x5a5a5a5a2e6306e0ULL;
static u64 sm64(u64 &s) {
  u64 z = (s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

// phase timings of the last spasm_echelonize call (seconds)
//  0 total  1 upload  2 FL  3 FL-cols  4 greedy  5 reorder+extract  6 density  7 schur  8 tail  9 download
// 10 rounds 11 flcol rounds 12 greedy windows 13 schur bytes 14 schur macs 15 schur kernel ms
double g_timings[16];

struct Echelon {
  Fp F;
  int64_t prime;
  int n0, m;
  DCsr U;
  DBuf<int> Uqinv;
  spasm_triplet *L = nullptr;  // host (only with opts->L)
  int *Lp = nullptr;           // host
  double t_pivots = 0, t_schur = 0, t_tail = 0;
};

// structural pivots of `cur` appended to U.  p (device, [n]) returns the permutation.
static int structural_round(Echelon &E, const DCsr &cur, const std::vector<int> &p_in, bool greedy, PivotSearch &P) {
  int counts[3];
  double t0 = spasm_wtime();
  P.t_fl = P.t_flcol = P.t_greedy = P.t_reorder = 0;
  P.flcol_rounds = P.greedy_windows = 0;
  find_structural_pivots(cur, greedy, P, counts);
  sync();
  (void)t0;
  logf("[pivots] Faugère-Lachartre: %d pivots found [%.1fs]\n", counts[0], P.t_fl);
  logf("[pivots] ``Faugère-Lachartre on columns'': %d pivots found [%.1fs]\n", counts[1], P.t_flcol);
  if (greedy) logf("[pivots] greedy alternating cycle-free search: %d pivots found [%.1fs]\n", counts[2], P.t_greedy);
  g_timings[2] += P.t_fl, g_timings[3] += P.t_flcol, g_timings[4] += P.t_greedy, g_timings[5] += P.t_reorder;
  g_timings[11] += P.flcol_rounds, g_timings[12] += P.greedy_windows;
  logf("[pivots] %d pivots found\n", P.npiv);
  const int urow0 = E.U.n;
  DBuf<uint32_t> pivval;
  {
    double t1 = spasm_wtime();
    extract_pivot_rows(cur, P, E.U, E.Uqinv, E.F, pivval);
    sync();
    g_timings[5] += spasm_wtime() - t1;
  }
  if (E.L != nullptr && P.npiv > 0) {
    std::vector<int> hp(P.npiv);
    std::vector<uint32_t> hv(P.npiv);
    P.p.download(hp.data(), P.npiv);
    pivval.download(hv.data(), P.npiv);
    sync();
    for (int k = 0; k < P.npiv; k++) {
      int i = hp[k], i_orig = p_in.empty() ? i : p_in[i];
      spasm_add_entry(E.L, i_orig, urow0 + k, (i64)hv[k]);
      E.Lp[urow0 + k] = i_orig;
    }
  }
  return P.npiv;
}

__global__ void k_gather_idx(const int *__restrict__ src, const int *__restrict__ idx, int n, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

__global__ void k_count_nonzero(const uint32_t *__restrict__ a, long long n, unsigned long long *__restrict__ out) {
  unsigned long long c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += (a[i] != 0);
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

static double estimate_density(Echelon &E, const DCsr &cur, const int *rows_dev, int nrows, int R_) {
  if (nrows == 0 || E.m == E.U.n) return 0;
  std::vector<int> idx(R_);
  for (int i = 0; i < R_; i++) idx[i] = (int)(sm64(g_seed_state) % (u64)nrows);
  DBuf<int> didx(R_), sample(R_);
  didx.upload(idx.data(), R_);
  k_gather_idx<<<cdiv(R_, 128), 128, 0, stream()>>>(rows_dev, didx.p, R_, sample.p);
  // the sampled rows go through the dense engine (R_ right-hand sides at once): in the dense regime
  // each of them reaches most pivots, which is the worst case of the row-at-a-time engine
  DenseSchur D;
  build_dense_schur(cur, sample.p, R_, E.U, E.Uqinv.p, E.F, D);
  DBuf<unsigned long long> cnt(1);
  cnt.zero();
  const long long tot = (long long)D.Sm0 * D.ld;
  k_count_nonzero<<<std::min<long long>(cdiv(tot, 256), 4096), 256, 0, stream()>>>(D.Dt.p, tot, cnt.p);
  CK(cudaGetLastError());
  const unsigned long long nz = fetch(cnt.p);
  return ((double)nz) / (E.m - E.U.n) / R_;
}

static void add_L_entries(Echelon &E, const SolveResult &R, const std::vector<int> &orig_rows) {
  const int n = (int)orig_rows.size();
  std::vector<long long> lp(n + 1);
  std::vector<int> lj(R.lnnz);
  std::vector<uint32_t> lx(R.lnnz);
  R.lp.download(lp.data(), n + 1);
  if (R.lnnz) R.lj.download(lj.data(), R.lnnz), R.lx.download(lx.data(), R.lnnz);
  sync();
  for (int k = 0; k < n; k++)
    for (long long e = lp[k]; e < lp[k + 1]; e++) spasm_add_entry(E.L, orig_rows[k], lj[e], (i64)lx[e]);
}

// the dense tail's multipliers and row permutation (with-L mode, dense.cuh) go to the host triplet / Lp of the factor
struct HostLSink : LSink {
  Echelon &E;
  const std::vector<int> &orig;  // tail row -> original row
  HostLSink(Echelon &E_, const std::vector<int> &o) : E(E_), orig(o) {}
  void rows(int row0, int nrows, int ubase, const int *cnt, const unsigned long long *offs, const int *oj, const uint32_t *ox) override {
    if (nrows <= 0) return;
    std::vector<int> hc(nrows);
    std::vector<unsigned long long> ho(nrows);
    CK(cudaMemcpyAsync(hc.data(), cnt, (size_t)nrows * sizeof(int), cudaMemcpyDeviceToHost, stream()));
    CK(cudaMemcpyAsync(ho.data(), offs, (size_t)nrows * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream()));
    sync();
    unsigned long long total = 0;
    for (int i = 0; i < nrows; i++) total = std::max(total, ho[i] + (unsigned long long)hc[i]);
    if (total == 0) return;
    std::vector<int> hj(total);
    std::vector<uint32_t> hx(total);
    download_large(hj.data(), oj, (size_t)total * sizeof(int));
    download_large(hx.data(), ox, (size_t)total * sizeof(uint32_t));
    sync();
    for (int i = 0; i < nrows; i++)
      for (unsigned long long e = ho[i]; e < ho[i] + (unsigned long long)hc[i]; e++) spasm_add_entry(E.L, orig[row0 + i], ubase + hj[e], (i64)hx[e]);
  }
  void pivots(int ubase, const int *pivrow, int rr, int row0) override {
    for (int s = 0; s < rr; s++) E.Lp[ubase + s] = orig[row0 + pivrow[s]];
  }
};

// ------------------------------------------------------------------ GPLU tail
// Row-by-row semantics (README.md:34-36) reproduced by speculative batches: a batch of rows is reduced against
// the current U in parallel; one warp then walks the batch IN ORDER and decides every row:
//   * a row that reduced to zero is final whatever happens before it (it stays zero against any larger U);
//   * a non-zero row is accepted as the next pivot row unless it holds a column pivoted earlier in this batch;
//   * from the first such conflict on, every non-zero row is DEFERRED: it is solved again in the next batch, in
//     the same relative order, against the enlarged U (its value, or the order in which pivots enter U, could
//     depend on the rows deferred before it).
// On rank-deficient tails (most rows vanish) almost nothing is deferred and the batches stay large.
// decision[t]: 0 zero row, 1 + k pivot number k of this batch, -1 deferred.
__global__ void k_gplu_commit(const long long *__restrict__ Rp, const int *__restrict__ Rj, int wn, unsigned char *__restrict__ newpiv,
                              int urows, int m, int *__restrict__ out /* [0]=ndeferred [1]=npivots */, int *__restrict__ pivt,
                              int *__restrict__ decision) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  int np = 0, nd = 0;
  bool dirty = false;
  for (int t0 = 0; t0 < wn; t0 += 32) {
    const int tt = t0 + lane;
    const long long ra = tt < wn ? Rp[tt] : 0, rb = tt < wn ? Rp[tt + 1] : 0;
    const bool nonzero = tt < wn && rb > ra;
    if (tt < wn && !nonzero) decision[tt] = 0;
    unsigned todo = __ballot_sync(FULL, nonzero);
    while (todo) {
      const int u = __ffs(todo) - 1;
      todo &= todo - 1;
      const int t = t0 + u;
      const long long a = __shfl_sync(FULL, ra, u), b = __shfl_sync(FULL, rb, u);
      bool defer = dirty;
      if (!defer && urows + np == m) {
        // every column is pivotal: this row cannot be non-zero against the final U; solve it again (it will vanish)
        defer = true;
      }
      if (!defer) {
        int conflict = 0;
        for (long long e = a + lane; e < b; e += 32) conflict |= newpiv[Rj[e]];
        if (__any_sync(FULL, conflict)) defer = true, dirty = true;
      }
      if (defer) {
        if (lane == 0) decision[t] = -1;
        nd++;
      } else {
        if (lane == 0) {
          newpiv[Rj[a]] = 1;  // entries are sorted: the first one is the leftmost column
          pivt[np] = t;
          decision[t] = 1 + np;
        }
        np++;
        __syncwarp();
      }
    }
  }
  if (lane == 0) out[0] = nd, out[1] = np;
}
// carry = the deferred rows of a batch in their REDUCED form (already eliminated against the U of that batch): the
// next batch only has to eliminate them against the pivots that entered U since, instead of solving them again
__global__ void k_carry_lens(const long long *__restrict__ Rp, const int *__restrict__ decision, int wn, int *__restrict__ flag,
                             int *__restrict__ len) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < wn) {
    const bool d = decision[t] == -1;
    flag[t] = d;
    len[t] = d ? (int)(Rp[t + 1] - Rp[t]) : 0;
  }
  if (t == wn) flag[t] = 0, len[t] = 0;
}
__global__ void k_carry_gather(const long long *__restrict__ Rp, const int *__restrict__ Rj, const uint32_t *__restrict__ Rx,
                               const int *__restrict__ decision, int wn, const long long *__restrict__ rowpos,
                               const long long *__restrict__ entpos, long long *__restrict__ Cp, int *__restrict__ Cj, uint32_t *__restrict__ Cx) {
  int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t > wn) return;
  if (t == wn) {
    if (lane == 0) Cp[rowpos[wn]] = entpos[wn];
    return;
  }
  if (decision[t] != -1) return;
  const long long a = Rp[t], b = Rp[t + 1], d = entpos[t];
  if (lane == 0) Cp[rowpos[t]] = d;
  for (long long e = a + lane; e < b; e += 32) Cj[d + (e - a)] = Rj[e], Cx[d + (e - a)] = Rx[e];
}
__global__ void k_shift_ptr(const long long *__restrict__ in, int n_plus_one, long long shift, long long *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_plus_one) out[i] = in[i] + shift;
}
__global__ void k_gplu_lens(const long long *__restrict__ Rp, const int *__restrict__ pivt, int np, int *__restrict__ len) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < np) len[k] = (int)(Rp[pivt[k] + 1] - Rp[pivt[k]]);
  if (k == np) len[k] = 0;
}
template <bool SMALL>
__global__ void k_gplu_append(const long long *__restrict__ Rp, const int *__restrict__ Rj, const uint32_t *__restrict__ Rx,
                              const int *__restrict__ pivt, int np, const long long *__restrict__ pos, long long ubase, int urow0,
                              long long *__restrict__ Up, int *__restrict__ Uj, uint32_t *__restrict__ Ux, int *__restrict__ Uqinv,
                              unsigned char *__restrict__ newpiv, uint32_t *__restrict__ pivval, Fp F) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= np) return;
  const int t = pivt[k];
  const long long a = Rp[t], b = Rp[t + 1], dst = ubase + pos[k];
  const uint32_t pv = Rx[a];
  const uint32_t beta = dev_inv(pv, F.p);
  if (lane == 0) {
    const int jp = Rj[a];
    Uj[dst] = jp;
    Ux[dst] = 1;
    Uqinv[jp] = urow0 + k;
    newpiv[jp] = 0;
    Up[urow0 + k + 1] = dst + (b - a);
    pivval[k] = pv;
  }
  for (long long e = a + 1 + lane; e < b; e += 32) {
    Uj[dst + (e - a)] = Rj[e];
    Ux[dst + (e - a)] = mulmod<SMALL>(beta, Rx[e], F);
  }
}

static void echelonize_GPLU(Echelon &E, const DCsr &cur, const int *rows_dev, int nrows, const std::vector<int> &orig_rows) {
  cudaStream_t s = stream();
  const int m = E.m;
  logf("[echelonize/GPLU] processing matrix of dimension %d x %d\n", nrows, m);
  DBuf<unsigned char> newpiv(m);
  newpiv.zero();
  DBuf<int> out(2), pivt, len, decision, cflag, clen;
  DBuf<long long> pos, crow, cent;
  DBuf<PDesc> pdesc;
  DBuf<uint32_t> pivval;
  DCsr carry;  // deferred rows, reduced against the U of the batch that deferred them
  carry.n = 0, carry.m = m, carry.nnz = 0;
  std::vector<int> carry_orig, hdec;
  // speculation is cheap to undo (deferred rows are carried in reduced form), so the batches start large
  int batch = 4096;
  const bool prof = getenv("SPASM_B200_PROFILE") != nullptr;
  int nbatches = 0;
  long long solved = 0, npivots = 0;
  double t_solve = 0;
  struct ProfGuard {
    const bool on;
    int &nb;
    long long &solved, &np;
    double &ts;
    int nrows;
    ~ProfGuard() {
      if (on) fprintf(stderr, "[GPLU] %d rows: %d batches, %lld row solves, %lld pivots, %.3fs in solve_rows\n", nrows, nb, solved, np, ts);
    }
  } prof_guard{prof, nbatches, solved, npivots, t_solve, nrows};
  for (int done = 0; done < nrows || carry.n > 0;) {
    if (E.U.n == m) {
      logf("\n[echelonize/GPLU] full rank reached\n");
      break;
    }
    const int nc = carry.n;
    const int wf = std::max(0, std::min(batch - nc, nrows - done));  // fresh rows of this batch
    const int wn = nc + wf;
    build_pdesc_U(E.U, E.Uqinv.p, pdesc);
    SolveSystem G{E.U.j.p, E.U.x.p, pdesc.p, m, &E.U, E.Uqinv.p};
    SolveEmit Em;
    Em.want_L = (E.L != nullptr);
    SolveResult R1, R2;
    const double ts0 = spasm_wtime();
    if (nc > 0) {
      SolveRows B1{carry.p.p, carry.j.p, carry.x.p, nullptr, nc, nullptr, true};
      solve_rows(G, B1, Em, E.F, R1);
    }
    if (wf > 0) {
      SolveRows B2{cur.p.p, cur.j.p, cur.x.p, rows_dev + done, wf, nullptr};
      solve_rows(G, B2, Em, E.F, R2);
    }
    t_solve += spasm_wtime() - ts0, nbatches++, solved += wn;
    // one result list: the carried rows first (they come first in row order), then the fresh ones
    SolveResult Rc;
    SolveResult *R = nullptr;
    if (nc > 0 && wf > 0) {
      Rc.nnz = R1.nnz + R2.nnz;
      Rc.p.alloc(wn + 1);
      Rc.j.alloc(std::max<long long>(Rc.nnz, 1));
      Rc.x.alloc(std::max<long long>(Rc.nnz, 1));
      CK(cudaMemcpyAsync(Rc.p.p, R1.p.p, (size_t)nc * sizeof(long long), cudaMemcpyDeviceToDevice, s));
      k_shift_ptr<<<cdiv(wf + 1, 256), 256, 0, s>>>(R2.p.p, wf + 1, R1.nnz, Rc.p.p + nc);
      if (R1.nnz) {
        CK(cudaMemcpyAsync(Rc.j.p, R1.j.p, (size_t)R1.nnz * sizeof(int), cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(Rc.x.p, R1.x.p, (size_t)R1.nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      }
      if (R2.nnz) {
        CK(cudaMemcpyAsync(Rc.j.p + R1.nnz, R2.j.p, (size_t)R2.nnz * sizeof(int), cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(Rc.x.p + R1.nnz, R2.x.p, (size_t)R2.nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      }
      R = &Rc;
    } else
      R = nc > 0 ? &R1 : &R2;
    pivt.alloc(wn);
    decision.alloc(wn);
    k_gplu_commit<<<1, 32, 0, s>>>(R->p.p, R->j.p, wn, newpiv.p, E.U.n, m, out.p, pivt.p, decision.p);
    int h[2];
    out.download(h, 2);
    sync();
    const int nd = h[0], np = h[1];
    npivots += np;
    if (prof && (nbatches <= 40 || nbatches % 50 == 0))
      fprintf(stderr, "[GPLU] batch %d: %d carried + %d fresh rows, %d deferred, %d pivots; fresh: light %lld global %lld dense %lld, %.1f + %.1f ms\n",
              nbatches, nc, wf, nd, np, R2.stats.light, R2.stats.medium, R2.stats.heavy, R1.stats.ms, R2.stats.ms);
    const int urow0 = E.U.n;
    if (np > 0) {
      len.alloc(np + 1);
      pos.alloc(np + 1);
      pivval.alloc(np);
      k_gplu_lens<<<cdiv(np + 1, 256), 256, 0, s>>>(R->p.p, pivt.p, np, len.p);
      exclusive_scan_i32_to_i64(len.p, pos.p, np + 1);
      const long long add = fetch(pos.p + np);
      csr_reserve(E.U, E.U.nnz + add, E.U.n + np);
      if (E.F.small)
        k_gplu_append<true><<<cdiv((long long)np * 32, 256), 256, 0, s>>>(R->p.p, R->j.p, R->x.p, pivt.p, np, pos.p, E.U.nnz, E.U.n, E.U.p.p,
                                                                         E.U.j.p, E.U.x.p, E.Uqinv.p, newpiv.p, pivval.p, E.F);
      else
        k_gplu_append<false><<<cdiv((long long)np * 32, 256), 256, 0, s>>>(R->p.p, R->j.p, R->x.p, pivt.p, np, pos.p, E.U.nnz, E.U.n, E.U.p.p,
                                                                          E.U.j.p, E.U.x.p, E.Uqinv.p, newpiv.p, pivval.p, E.F);
      CK(cudaGetLastError());
      E.U.nnz += add;
      E.U.n += np;
    }
    const bool need_dec = (E.L != nullptr);
    if (need_dec) {
      hdec.resize(wn);
      decision.download(hdec.data(), wn);
      // multipliers found in this batch (final for decided AND deferred rows: later pivots never touch them), then
      // the pivot entries
      std::vector<int> new_carry_orig;
      std::vector<uint32_t> hv(std::max(np, 1));
      if (np) pivval.download(hv.data(), np);
      for (int part = 0; part < 2; part++) {
        SolveResult &Rp_ = part == 0 ? R1 : R2;
        const int cnt = part == 0 ? nc : wf, base = part == 0 ? 0 : nc;
        if (cnt == 0) continue;
        std::vector<long long> lp(cnt + 1);
        std::vector<int> lj(Rp_.lnnz);
        std::vector<uint32_t> lx(Rp_.lnnz);
        Rp_.lp.download(lp.data(), cnt + 1);
        if (Rp_.lnnz) Rp_.lj.download(lj.data(), Rp_.lnnz), Rp_.lx.download(lx.data(), Rp_.lnnz);
        sync();
        for (int t = 0; t < cnt; t++) {
          const int i_orig = part == 0 ? carry_orig[t] : orig_rows[done + t];
          for (long long e = lp[t]; e < lp[t + 1]; e++) spasm_add_entry(E.L, i_orig, lj[e], (i64)lx[e]);
          const int d = hdec[base + t];
          if (d > 0) {
            E.Lp[urow0 + d - 1] = i_orig;
            spasm_add_entry(E.L, i_orig, urow0 + d - 1, (i64)hv[d - 1]);
          } else if (d < 0)
            new_carry_orig.push_back(i_orig);
        }
      }
      carry_orig.swap(new_carry_orig);
    }
    // the deferred rows, reduced, become the carry of the next batch
    DCsr next;
    next.n = nd, next.m = m, next.nnz = 0;
    if (nd > 0) {
      cflag.alloc(wn + 1);
      clen.alloc(wn + 1);
      crow.alloc(wn + 1);
      cent.alloc(wn + 1);
      k_carry_lens<<<cdiv(wn + 1, 256), 256, 0, s>>>(R->p.p, decision.p, wn, cflag.p, clen.p);
      exclusive_scan_i32_to_i64(cflag.p, crow.p, wn + 1);
      exclusive_scan_i32_to_i64(clen.p, cent.p, wn + 1);
      next.nnz = fetch(cent.p + wn);
      next.p.alloc(nd + 1);
      next.j.alloc(std::max<long long>(next.nnz, 1));
      next.x.alloc(std::max<long long>(next.nnz, 1));
      k_carry_gather<<<cdiv((long long)(wn + 1) * 32, 256), 256, 0, s>>>(R->p.p, R->j.p, R->x.p, decision.p, wn, crow.p, cent.p, next.p.p, next.j.p,
                                                                       next.x.p);
      CK(cudaGetLastError());
    }
    sync();
    carry = std::move(next);
    done += wf;
    batch = (2 * nd <= wn) ? std::min(batch * 2, 32768) : std::max(256, std::min(batch, 2 * (wn - nd) + 256));
  }
}

// ------------------------------------------------------------------ the driver
static spasm_lu *echelonize_impl(const spasm_csr *A, echelonize_opts *opts, const DCsr *resident = nullptr, bool download = true,
                                 int *rank_out = nullptr) {
  ApiCall api_scope_;
  echelonize_opts defaults;
  if (opts == nullptr) {
    spasm_echelonize_init_opts(&defaults);
    opts = &defaults;
  }
  const int n0 = A->n, m = A->m;
  const int64_t prime = A->field->p;
  g_seed_state = SEED0;
  const double start = spasm_wtime();
  for (double &t : g_timings) t = 0;
  logf("[echelonize] Start on %d x %d matrix with %lld nnz\n", n0, m, (long long)spasm_nnz(A));
  if (opts->complete) opts->L = 1;

  Echelon E;
  E.F = make_field(prime);
  E.prime = prime, E.n0 = n0, E.m = m;
  E.U.n = 0, E.U.m = m, E.U.nnz = 0;
  E.U.p.alloc(n0 + 2);
  E.U.p.zero();
  E.U.j.alloc(std::max<int64_t>(spasm_nnz(A), 16));
  E.U.x.alloc(std::max<int64_t>(spasm_nnz(A), 16));
  E.Uqinv.alloc(std::max(m, 1));
  E.Uqinv.fill_ff();
  if (opts->L) {
    E.L = spasm_triplet_alloc(n0, n0, std::max<i64>(spasm_nnz(A), 16), prime, true);
    const i64 plen = std::max(n0, m) + 1;
    E.Lp = (int *)spasm_malloc(plen * (i64)sizeof(int));
    for (i64 i = 0; i < plen; i++) E.Lp[i] = -1;
  }

  DCsr A0, S;
  if (resident == nullptr) {
    upload_csr(A, A0, E.F);
    sync();
  }
  g_timings[1] = spasm_wtime() - start;
  const DCsr *cur = resident ? resident : &A0;
  int n = n0, npiv = 0;
  std::vector<int> p_in;  // current row -> original row (empty: identity)
  double density = (n0 > 0 && m > 0) ? (double)spasm_nnz(A) / n0 / m : 0.0;
  bool finished = false, go_dense = false;
  bool tail_distributed = false;  // the dense tail ran with its rows (and, on the other ranks, the rows of U) spread over the ranks
  int *sink_arrays_j = nullptr, *sink_arrays_x = nullptr;
  long long sink_done = 0, sink_cap = 0;
  PivotSearch P;
  P.p.alloc(std::max(n, 1));
  {
    std::vector<int> id(std::max(n, 1));
    for (int i = 0; i < n; i++) id[i] = i;
    P.p.upload(id.data(), n);
    sync();
  }

  for (int round = 0; round < opts->max_round; round++) {
    logf("[echelonize] round %d\n", round);
    g_timings[10] += 1;
    double t0 = spasm_wtime();
    g_phase = "structural pivots";
    npiv = structural_round(E, *cur, p_in, opts->enable_greedy_pivot_search, P);
    E.t_pivots += spasm_wtime() - t0;
    const int rem_rows = n - npiv, rem_cols = m - E.U.n;
    if (rem_rows == 0 || rem_cols == 0) {
      finished = true;
      break;
    }
    const int bound = std::min(n, rem_cols);
    if (npiv < opts->min_pivot_proportion * bound) {
      logf("[echelonize] not enough pivots found; stopping\n");
      break;
    }
    t0 = spasm_wtime();
    g_phase = "density estimate";
    density = estimate_density(E, *cur, P.p.p + npiv, rem_rows, 100);
    g_timings[6] += spasm_wtime() - t0, t0 = spasm_wtime();
    logf("Schur complement is %d x %d, estimated density : %.2f (%lld byte)\n", rem_rows, rem_cols, density,
         (long long)(4.0 * density * rem_rows * rem_cols));
    if (density > opts->sparsity_threshold && opts->enable_dense) {
      logf("[echelonize] Schur complement is dense; stopping\n");
      go_dense = true;
      break;
    }
    // ---- Schur complement on the non-pivotal rows
    g_phase = "schur complement";
    DBuf<PDesc> pdesc;
    build_pdesc_U(E.U, E.Uqinv.p, pdesc);
    SolveSystem G{E.U.j.p, E.U.x.p, pdesc.p, m, &E.U, E.Uqinv.p};
    SolveRows B{cur->p.p, cur->j.p, cur->x.p, P.p.p + npiv, rem_rows, nullptr};
    B.collective = true;  // spasm_echelonize is entered by every rank with the same matrix: the non-pivotal rows are split over the ranks
    SolveEmit Em;
    Em.want_L = (E.L != nullptr);
    SolveResult R;
    solve_rows(G, B, Em, E.F, R);
    g_last_stats = R.stats;
    g_timings[13] += (double)R.stats.bytes, g_timings[14] += (double)R.stats.macs, g_timings[15] += R.stats.ms;
    std::vector<int> hp(rem_rows), p_out(rem_rows);
    CK(cudaMemcpyAsync(hp.data(), P.p.p + npiv, (size_t)rem_rows * sizeof(int), cudaMemcpyDeviceToHost, stream()));
    sync();
    for (int k = 0; k < rem_rows; k++) p_out[k] = p_in.empty() ? hp[k] : p_in[hp[k]];
    if (E.L != nullptr) add_L_entries(E, R, p_out);
    DCsr Snew;
    Snew.n = rem_rows, Snew.m = m, Snew.nnz = R.nnz;
    Snew.p = std::move(R.p);
    Snew.j = std::move(R.j);
    Snew.x = std::move(R.x);
    S = std::move(Snew);
    cur = &S;
    p_in.swap(p_out);
    n = rem_rows;
    npiv = 0;
    density = (n > 0 && rem_cols > 0) ? (double)S.nnz / n / rem_cols : 0.0;
    logf("Schur complement: %d * %d [%lld nz / density= %.3f], %.1fs\n", n, m, (long long)S.nnz, density, spasm_wtime() - t0);
    E.t_schur += spasm_wtime() - t0;
    g_timings[7] += spasm_wtime() - t0;
    // identity permutation for the finish stage if the loop ends here
    if (P.p.n < (size_t)n) P.p.alloc(n);
    {
      std::vector<int> id(n);
      for (int i = 0; i < n; i++) id[i] = i;
      P.p.upload(id.data(), n);
      sync();
    }
  }

  if (!finished) {
    const int rem_rows = n - npiv;
    const double aspect_ratio = (double)rem_rows / m;
    logf("[echelonize] finishing; density = %.3f; aspect ratio = %.1f\n", density, aspect_ratio);
    std::vector<int> orig(rem_rows);
    if (E.L != nullptr) {
      std::vector<int> hp(rem_rows);
      CK(cudaMemcpyAsync(hp.data(), P.p.p + npiv, (size_t)rem_rows * sizeof(int), cudaMemcpyDeviceToHost, stream()));
      sync();
      for (int k = 0; k < rem_rows; k++) orig[k] = p_in.empty() ? hp[k] : p_in[hp[k]];
    }
    const double t0 = spasm_wtime();
    g_phase = "finish (GPLU / dense tail)";
    std::unique_ptr<HostSink> sink_holder;
    if (download && !opts->L && getenv("SPASM_B200_NO_STREAMING") == nullptr) {
      sink_holder.reset(new HostSink());
      g_sink = sink_holder.get();
    }
    struct SinkGuard {
      ~SinkGuard() { g_sink = nullptr; }
    } sink_guard;
    TailOpts topt;
    topt.tall_skinny = opts->enable_tall_and_skinny, topt.low_rank_ratio = opts->low_rank_ratio, topt.start_weight = opts->low_rank_start_weight;
    if (opts->L && opts->enable_dense && (go_dense || density > opts->sparsity_threshold)) {
      // dense tail with L (replaces spasm_ffpack_LU, src/SpaSM.jl:806): row echelon form on the tensor cores
      HostLSink lsink(E, orig);
      topt.lsink = &lsink;
      topt.tall_skinny = false;
      echelonize_dense_device(*cur, P.p.p + npiv, rem_rows, E.U, E.Uqinv, E.F, opts->dense_block_size, topt);
    } else if (opts->L || (!opts->enable_dense && opts->enable_GPLU))
      echelonize_GPLU(E, *cur, P.p.p + npiv, rem_rows, orig);
    else if (opts->enable_tall_and_skinny && aspect_ratio > opts->tall_and_skinny_ratio)
      echelonize_lowrank_device(*cur, P.p.p + npiv, rem_rows, E.U, E.Uqinv, E.F, opts->dense_block_size, opts->low_rank_start_weight);
    else if (opts->enable_dense && (go_dense || density > opts->sparsity_threshold))
      echelonize_dense_device(*cur, P.p.p + npiv, rem_rows, E.U, E.Uqinv, E.F, opts->dense_block_size, topt), tail_distributed = true;
    else if (opts->enable_GPLU)
      echelonize_GPLU(E, *cur, P.p.p + npiv, rem_rows, orig);
    else if (opts->enable_dense)
      echelonize_dense_device(*cur, P.p.p + npiv, rem_rows, E.U, E.Uqinv, E.F, opts->dense_block_size, topt), tail_distributed = true;
    else
      logf("[echelonize] Cannot finish (no valid method enabled). Incomplete echelonization returned\n");
    sync();
    E.t_tail += spasm_wtime() - t0;
    g_timings[8] = E.t_tail;
    if (sink_holder && sink_holder->submitted > 0) {
      sink_done = sink_holder->submitted;
      sink_cap = sink_holder->cap;
      sink_holder->release(&sink_arrays_j, &sink_arrays_x);  // waits for the copies in flight
    }
  }
  const double t_dl = spasm_wtime();
  g_phase = "download";
  if (rank_out) *rank_out = E.U.n;
  if (!download) {  // device-resident timing mode (bench.py `value`): the factor is dropped on the device
    g_timings[0] = spasm_wtime() - start;
    logf("[echelonize] Done in %.1fs. Rank %d, %lld nz in basis\n", spasm_wtime() - start, E.U.n, (long long)E.U.nnz);
    if (E.L) spasm_triplet_free(E.L);
    free(E.Lp);
    return nullptr;
  }

  // ---- the factor goes back to host memory (plain malloc arrays: src/SpaSM.jl:279-304 unsafe_loads them)
  spasm_lu *fact = (spasm_lu *)spasm_malloc(sizeof(spasm_lu));
  E.U.m = m;
  {
    spasm_csr *Uh = nullptr;
    if (sink_arrays_j != nullptr) {
      // the dense tail streamed [0, sink_done) already; fetch what came after (nothing, normally)
      Uh = spasm_csr_alloc(E.U.n, m, 0, prime, true);
      free(Uh->j), free(Uh->x);
      Uh->j = sink_arrays_j, Uh->x = sink_arrays_x;
      Uh->nzmax = sink_cap;
      if (E.U.nnz > sink_done) {
        if (E.U.nnz > Uh->nzmax) spasm_csr_realloc(Uh, E.U.nnz);
        download_large(Uh->j + sink_done, E.U.j.p + sink_done, (size_t)(E.U.nnz - sink_done) * sizeof(int));
        convert_to_balanced(E.U.x.p + sink_done, (int *)E.U.x.p + sink_done, E.U.nnz - sink_done, E.F);
        download_large(Uh->x + sink_done, E.U.x.p + sink_done, (size_t)(E.U.nnz - sink_done) * sizeof(int));
      }
      spasm_csr_realloc(Uh, E.U.nnz);
    } else {
      Uh = spasm_csr_alloc(E.U.n, m, E.U.nnz, prime, true);
      if (E.U.nnz) {
        download_large(Uh->j, E.U.j.p, (size_t)E.U.nnz * sizeof(int));
        // balanced representatives are produced in place (the device copy of U is dropped afterwards)
        convert_to_balanced(E.U.x.p, (int *)E.U.x.p, E.U.nnz, E.F);
        download_large(Uh->x, E.U.x.p, (size_t)E.U.nnz * sizeof(int));
      }
    }
    CK(cudaMemcpyAsync(Uh->p, E.U.p.p, (size_t)(E.U.n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, stream()));
    fact->U = Uh;
  }
  fact->qinv = (int *)spasm_malloc((i64)std::max(m, 1) * (i64)sizeof(int));
  CK(cudaMemcpyAsync(fact->qinv, E.Uqinv.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, stream()));
  sync();
  g_timings[9] = spasm_wtime() - t_dl;
  fact->r = E.U.n;
  fact->complete = 0;
  {
    const Dist &dd = dist();
    // only the sharded dense tail leaves rows of U on other ranks; everything else is computed identically by every rank
    fact->partial = (dd.nranks > 1 && tail_distributed && (dd.shard_factor || dd.rank != 0)) ? 1 : 0;
  }
  fact->L = nullptr;
  fact->p = E.Lp;
  fact->Ltmp = nullptr;
  if (opts->L) {
    // L: n0 x r, every row by increasing U-row index (N1)
    E.L->n = n0;
    E.L->m = std::max(E.U.n, 1);
    spasm_csr *Lc = spasm_compress(E.L);
    Lc->m = E.U.n;
    for (int i = 0; i < Lc->n; i++) {
      const i64 a = Lc->p[i], b = Lc->p[i + 1];
      std::vector<std::pair<int, spasm_ZZp>> row(b - a);
      for (i64 e = a; e < b; e++) row[e - a] = {Lc->j[e], Lc->x[e]};
      std::sort(row.begin(), row.end());
      for (i64 e = a; e < b; e++) Lc->j[e] = row[e - a].first, Lc->x[e] = row[e - a].second;
    }
    fact->L = Lc;
    spasm_triplet_free(E.L);
    fact->complete = 1;
  }
  g_timings[0] = spasm_wtime() - start;
  logf("[echelonize] Done in %.1fs. Rank %d, %lld nz in basis\n", spasm_wtime() - start, fact->r, (long long)E.U.nnz);
  return fact;
}

}  // namespace sb

using namespace sb;

extern "C" {

struct spasm_lu *spasm_echelonize(const struct spasm_csr *A, struct echelonize_opts *opts) {
  try {
    return echelonize_impl(A, opts);
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_echelonize failed during %s: %s\n", g_phase, e.what());
    return nullptr;
  }
}

// ---- device-resident entry points used by bench.py for the kernel-only number: the input CSR is
// uploaded once, each call runs the whole echelonization from HBM and returns the rank only.
struct ResidentMatrix {
  const spasm_csr *host;
  DCsr dev;
};
void *spasm_b200_upload(const struct spasm_csr *A) {
  try {
    ApiCall api_scope_;
    auto *h = new ResidentMatrix{A, {}};
    upload_csr(A, h->dev, make_field(A->field->p));
    sync();
    return h;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_b200_upload failed: %s\n", e.what());
    return nullptr;
  }
}
int spasm_b200_echelonize_resident(void *handle, struct echelonize_opts *opts, double *ms) {
  try {
    ApiCall api_scope_;
    auto *h = (ResidentMatrix *)handle;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, stream()));
    int rank = -1;
    echelonize_impl(h->host, opts, &h->dev, false, &rank);
    CK(cudaEventRecord(e1, stream()));
    CK(cudaEventSynchronize(e1));
    float t = 0;
    CK(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = t;
    return rank;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_b200_echelonize_resident failed: %s\n", e.what());
    return -1;
  }
}
void spasm_b200_release(void *handle) { delete (ResidentMatrix *)handle; }

void spasm_b200_last_timings(double *out) {
  for (int i = 0; i < 16; i++) out[i] = g_timings[i];
}

void spasm_lu_free(struct spasm_lu *N) {
  if (!N) return;
  free(N->qinv);
  free(N->p);
  spasm_csr_free(N->U);
  spasm_csr_free(N->L);
  spasm_triplet_free(N->Ltmp);
  free(N);
}

// prototype src/SpaSM.jl:776-777: host structs in, host structs out (one round of structural pivots)
int spasm_pivots_extract_structural(const struct spasm_csr *A, const int *p_in, struct spasm_lu *fact, int *p,
                                    struct echelonize_opts *opts) {
  try {
    ApiCall api_scope_;
    Echelon E;
    E.prime = A->field->p;
    E.F = make_field(E.prime);
    E.n0 = A->n, E.m = A->m;
    spasm_csr *Uh = fact->U;
    upload_csr(Uh, E.U, E.F);
    E.U.p.grow(Uh->n + A->n + 2);
    E.Uqinv.alloc(std::max(A->m, 1));
    E.Uqinv.upload(fact->qinv, A->m);
    E.L = fact->Ltmp;
    E.Lp = fact->p;
    DCsr dA;
    upload_csr(A, dA, E.F);
    std::vector<int> pin;
    if (p_in) pin.assign(p_in, p_in + A->n);
    PivotSearch P;
    const int urows0 = E.U.n;
    const int npiv = structural_round(E, dA, pin, opts == nullptr || opts->enable_greedy_pivot_search, P);
    if (A->n) P.p.download(p, A->n);
    // write the grown U back into the caller's struct
    const i64 unz = E.U.nnz;
    if (unz > Uh->nzmax) spasm_csr_realloc(Uh, unz);
    std::vector<long long> up(E.U.n + 1);
    E.U.p.download(up.data(), E.U.n + 1);
    sync();
    for (int i = urows0; i <= E.U.n; i++) Uh->p[i] = up[i];
    if (unz) {
      const i64 old = up[urows0];
      CK(cudaMemcpyAsync(Uh->j + old, E.U.j.p + old, (size_t)(unz - old) * sizeof(int), cudaMemcpyDeviceToHost, stream()));
      std::vector<uint32_t> hx(unz - old);
      CK(cudaMemcpyAsync(hx.data(), E.U.x.p + old, (size_t)(unz - old) * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
      sync();
      for (i64 e = old; e < unz; e++) Uh->x[e] = to_bal(hx[e - old], E.F);
    }
    Uh->n = E.U.n;
    E.Uqinv.download(fact->qinv, A->m);
    sync();
    return npiv;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_pivots_extract_structural failed: %s\n", e.what());
    return -1;
  }
}

}  // extern "C"
