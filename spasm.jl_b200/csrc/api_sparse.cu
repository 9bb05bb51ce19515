// api_sparse.cu — C-ABI entry points built on the sparse row-solve engine:
//   spasm_schur, spasm_schur_estimate_density            (prototypes src/SpaSM.jl:761-764)
//   spasm_sparse_triangular_solve                        (src/SpaSM.jl:694-722)
//   spasm_kernel, spasm_rref                             (src/SpaSM.jl:871, :876-882)
#include <algorithm>
#include <numeric>

#include "dist.cuh"
#include "factor.cuh"

namespace sb {

__global__ void k_pdesc_U(const long long *__restrict__ Up, const int *__restrict__ qinv, int m, PDesc *__restrict__ pd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  int i = qinv[c];
  PDesc d;
  if (i >= 0) {
    d.start = Up[i];
    d.len = (int)(Up[i + 1] - Up[i]);
    d.prio = i;
  } else {
    d.start = 0, d.len = -1, d.prio = -1;
  }
  pd[c] = d;
}
void build_pdesc_U(const DCsr &U, const int *qinv, DBuf<PDesc> &pdesc) {
  if (pdesc.n < (size_t)U.m) pdesc.alloc(U.m);
  if (U.m) k_pdesc_U<<<cdiv(U.m, 256), 256, 0, stream()>>>(U.p.p, qinv, U.m, pdesc.p);
  CK(cudaGetLastError());
}

void DevFactor::upload(const spasm_lu *fact) {
  if (fact->partial)
    throw Error("this factor is partial: the rows of U are held by other ranks of the multi-GPU run (rank 0 holds the complete factor "
                "unless spasm_b200_dist_shard_factor(1) was requested)");
  prime = fact->U->field->p;
  F = make_field(prime);
  upload_csr(fact->U, U, F);
  qinv.alloc(U.m);
  qinv.upload(fact->qinv, U.m);
}

void csr_reserve(DCsr &U, int64_t nnz_needed, int rows_needed) {
  U.j.grow(nnz_needed);
  U.x.grow(nnz_needed);
  U.p.grow(rows_needed + 1);
}

__global__ void k_relabel(int *__restrict__ j, long long nnz, const int *__restrict__ pivcol) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) j[e] = pivcol[j[e]];
}
__global__ void k_pivcol(const long long *__restrict__ Up, const int *__restrict__ Uj, int r, int *__restrict__ pivcol) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < r) pivcol[i] = Uj[Up[i]];
}
__global__ void k_pdesc_Ut(const long long *__restrict__ Tp, const int *__restrict__ qinv, int m, int r, PDesc *__restrict__ pd,
                           int *__restrict__ isfree) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  int i = qinv[c];
  PDesc d;
  if (i >= 0) {
    d.start = Tp[c];
    d.len = (int)(Tp[c + 1] - Tp[c]);
    d.prio = r - 1 - i;  // back-substitution: last U row first
  } else {
    d.start = 0, d.len = -1, d.prio = -1;
  }
  pd[c] = d;
  isfree[c] = i < 0;
}
__global__ void k_compact_flags(const int *__restrict__ flag, const long long *__restrict__ pos, int n, int *__restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n && flag[c]) out[pos[c]] = c;
}

void build_kernel_system(const DevFactor &f, KernelSystem &K) {
  const int m = f.U.m, r = f.U.n;
  transpose_csr(f.U, K.Ut);
  DBuf<int> pivcol(r);
  if (r) k_pivcol<<<cdiv(r, 256), 256, 0, stream()>>>(f.U.p.p, f.U.j.p, r, pivcol.p);
  if (K.Ut.nnz) k_relabel<<<cdiv(K.Ut.nnz, 256), 256, 0, stream()>>>(K.Ut.j.p, K.Ut.nnz, pivcol.p);
  K.pdesc.alloc(m);
  DBuf<int> isfree(m + 1);
  DBuf<long long> pos(m + 1);
  if (m) k_pdesc_Ut<<<cdiv(m, 256), 256, 0, stream()>>>(K.Ut.p.p, f.qinv.p, m, r, K.pdesc.p, isfree.p);
  exclusive_scan_i32_to_i64(isfree.p, pos.p, m + 1);
  K.nfree = m - r;
  K.freecols.alloc(K.nfree);
  if (m) k_compact_flags<<<cdiv(m, 256), 256, 0, stream()>>>(isfree.p, pos.p, m, K.freecols.p);
  CK(cudaGetLastError());
}

__global__ void k_to_bal2(const uint32_t *__restrict__ in, int *__restrict__ out, long long n, Fp F) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_bal(in[i], F);
}

spasm_csr *result_to_host_csr(const SolveResult &R, int nrows, int m, int64_t prime, const Fp &F) {
  spasm_csr *S = spasm_csr_alloc(nrows, m, R.nnz, prime, true);
  R.p.download((long long *)S->p, nrows + 1);
  if (R.nnz) {
    R.j.download(S->j, R.nnz);
    DBuf<int> tmp(R.nnz);
    k_to_bal2<<<cdiv(R.nnz, 256), 256, 0, stream()>>>(R.x.p, tmp.p, R.nnz, F);
    tmp.download(S->x, R.nnz);
  }
  sync();
  return S;
}

WorkStats g_last_stats;

}  // namespace sb

using namespace sb;

extern "C" {

// work counters of the last schur / kernel call: {bytes, macs, rows, light, medium, heavy, ms*1000}
void spasm_b200_last_stats(long long *out) {
  out[0] = g_last_stats.bytes, out[1] = g_last_stats.macs, out[2] = g_last_stats.rows;
  out[3] = g_last_stats.light, out[4] = g_last_stats.medium, out[5] = g_last_stats.heavy;
  out[6] = (long long)(g_last_stats.ms * 1000.0);
}

// {row-engine calls that were split over the ranks, rows THIS rank solved in them} since the last reset
void spasm_b200_shard_stats(long long *out, int reset) {
  out[0] = g_shard_stats[0], out[1] = g_shard_stats[1];
  if (reset) g_shard_stats[0] = g_shard_stats[1] = 0;
}

static u64 g_prng = 0x5a5a5a5a2e6306e0ULL;
void spasm_b200_seed(u64 seed) { g_prng = seed; }
static u64 splitmix64_next() {
  u64 z = (g_prng += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

struct spasm_csr *spasm_schur(const struct spasm_csr *A, const int *p, int n, const struct spasm_lu *fact, double est_density,
                              struct spasm_triplet *L, const int *p_in, int *p_out) {
  (void)est_density;
  try {
    ApiCall api_scope_;
    double t0 = spasm_wtime();
    DevFactor f;
    f.upload(fact);
    DCsr dA;
    upload_csr(A, dA, f.F);
    DBuf<PDesc> pdesc;
    build_pdesc_U(f.U, f.qinv.p, pdesc);
    DBuf<int> rows(n);
    rows.upload(p, n);
    SolveSystem G{f.U.j.p, f.U.x.p, pdesc.p, f.U.m, &f.U, f.qinv.p};
    SolveRows B{dA.p.p, dA.j.p, dA.x.p, rows.p, n, nullptr};
    SolveEmit E;
    E.want_L = (L != nullptr);
    SolveResult R;
    solve_rows(G, B, E, f.F, R);
    g_last_stats = R.stats;
    spasm_csr *S = result_to_host_csr(R, n, A->m, f.prime, f.F);
    for (int k = 0; k < n; k++) {
      int i = p[k];
      if (p_out) p_out[k] = p_in ? p_in[i] : i;
    }
    if (L != nullptr) {
      std::vector<long long> lp(n + 1);
      std::vector<int> lj(R.lnnz);
      std::vector<uint32_t> lx(R.lnnz);
      R.lp.download(lp.data(), n + 1);
      if (R.lnnz) R.lj.download(lj.data(), R.lnnz), R.lx.download(lx.data(), R.lnnz);
      sync();
      for (int k = 0; k < n; k++) {
        int i = p[k], i_orig = p_in ? p_in[i] : i;
        for (long long e = lp[k]; e < lp[k + 1]; e++) spasm_add_entry(L, i_orig, lj[e], (i64)lx[e]);
      }
    }
    double dens = (n > 0 && A->m > fact->U->n) ? (double)R.nnz / ((double)n * (A->m - fact->U->n)) : 0.0;
    logf("Schur complement: %d * %d [%lld nz / density= %.3f], %.1fs\n", n, A->m, (long long)R.nnz, dens, spasm_wtime() - t0);
    return S;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_schur failed: %s\n", e.what());
    return nullptr;
  }
}

double spasm_schur_estimate_density(const struct spasm_csr *A, const int *p, int n, const struct spasm_csr *U, const int *qinv,
                                    int R_) {
  try {
    ApiCall api_scope_;
    if (n == 0 || A->m == U->n) return 0;
    Fp F = make_field(A->field->p);
    DCsr dA, dU;
    upload_csr(A, dA, F);
    upload_csr(U, dU, F);
    DBuf<int> dq(U->m);
    dq.upload(qinv, U->m);
    DBuf<PDesc> pdesc;
    build_pdesc_U(dU, dq.p, pdesc);
    std::vector<int> sample(R_);
    for (int i = 0; i < R_; i++) sample[i] = p[splitmix64_next() % (u64)n];
    DBuf<int> rows(R_);
    rows.upload(sample.data(), R_);
    SolveSystem G{dU.j.p, dU.x.p, pdesc.p, dU.m};
    SolveRows B{dA.p.p, dA.j.p, dA.x.p, rows.p, R_, nullptr};
    SolveEmit E;
    E.count_only = true;
    SolveResult R;
    solve_rows(G, B, E, F, R);
    return ((double)R.nnz) / (A->m - U->n) / R_;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_schur_estimate_density failed: %s\n", e.what());
    return -1;
  }
}

// src/SpaSM.jl:694-722.  On output x[j] is valid for every j in xj[top:m]; the pattern is the
// structural reach (as in the reference), listed with pivotal columns by increasing U row and the
// other columns after them — a topological order, which is all the contract promises.
int spasm_sparse_triangular_solve(const struct spasm_csr *U, const struct spasm_csr *B, int k, int *xj, spasm_ZZp *x,
                                  const int *qinv) {
  try {
    ApiCall api_scope_;
    const int m = U->m;
    Fp F = make_field(U->field->p);
    DCsr dU, dB;
    upload_csr(U, dU, F);
    // only row k of B travels
    spasm_csr Bk = *B;
    i64 bp[2] = {0, B->p[k + 1] - B->p[k]};
    Bk.n = 1, Bk.p = bp, Bk.j = B->j + B->p[k], Bk.x = B->x + B->p[k];
    upload_csr(&Bk, dB, F);
    DBuf<int> dq(m);
    dq.upload(qinv, m);
    DBuf<PDesc> pdesc;
    build_pdesc_U(dU, dq.p, pdesc);
    SolveSystem G{dU.j.p, dU.x.p, pdesc.p, m};
    SolveRows Br{dB.p.p, dB.j.p, dB.x.p, nullptr, 1, nullptr};
    SolveEmit E;
    E.structural = true;
    SolveResult R;
    solve_rows(G, Br, E, F, R);
    std::vector<int> cj(R.nnz);
    std::vector<uint32_t> cx(R.nnz);
    if (R.nnz) R.j.download(cj.data(), R.nnz), R.x.download(cx.data(), R.nnz);
    sync();
    int cntp = (int)R.nnz;
    int top = m - cntp;
    // pivotal columns by increasing U row, then the non-pivotal ones by increasing column
    std::vector<int> order(cntp);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      int qa = qinv[cj[a]], qb = qinv[cj[b]];
      if ((qa >= 0) != (qb >= 0)) return qa >= 0;
      if (qa >= 0) return qa < qb;
      return cj[a] < cj[b];
    });
    for (int t = 0; t < cntp; t++) {
      int c = cj[order[t]];
      xj[top + t] = c;
      x[c] = to_bal(cx[order[t]], F);
    }
    return top;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_sparse_triangular_solve failed: %s\n", e.what());
    return -1;
  }
}

// src/SpaSM.jl:876-882; README.md:39-41
struct spasm_csr *spasm_kernel(const struct spasm_lu *fact) {
  try {
    ApiCall api_scope_;
    double t0 = spasm_wtime();
    const int r = fact->U->n, m = fact->U->m;
    logf("[kernel] start. U is %d x %d (%lld nnz). Transposing U\n", r, m, (long long)spasm_nnz(fact->U));
    DevFactor f;
    f.upload(fact);
    KernelSystem K;
    build_kernel_system(f, K);
    SolveSystem G{K.Ut.j.p, K.Ut.x.p, K.pdesc.p, m};
    SolveRows B{K.Ut.p.p, K.Ut.j.p, K.Ut.x.p, K.freecols.p, K.nfree, nullptr};
    B.collective = dist().shard_rows;  // free columns split over the ranks when every rank makes this call (opt-in)
    SolveEmit E;
    E.all_columns = true;
    E.prefix_col = K.freecols.p;
    E.prefix_val = f.F.p - 1;  // -1
    SolveResult R;
    solve_rows(G, B, E, f.F, R);
    g_last_stats = R.stats;
    spasm_csr *Kh = result_to_host_csr(R, K.nfree, m, f.prime, f.F);
    logf("kernel: %d/%d, |K| = %lld\n", K.nfree, K.nfree, (long long)R.nnz);
    logf("[kernel] done in %.1fs. NNZ(K) = %lld\n", spasm_wtime() - t0, (long long)R.nnz);
    return Kh;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_kernel failed: %s\n", e.what());
    return nullptr;
  }
}

// src/SpaSM.jl:871
struct spasm_csr *spasm_rref(const struct spasm_lu *fact, int *Rqinv) {
  try {
    ApiCall api_scope_;
    const int r = fact->U->n, m = fact->U->m;
    DevFactor f;
    f.upload(fact);
    DBuf<PDesc> pdesc;
    build_pdesc_U(f.U, f.qinv.p, pdesc);
    DBuf<int> pivcol(r);
    if (r) k_pivcol<<<cdiv(r, 256), 256, 0, stream()>>>(f.U.p.p, f.U.j.p, r, pivcol.p);
    SolveSystem G{f.U.j.p, f.U.x.p, pdesc.p, m};
    SolveRows B{f.U.p.p, f.U.j.p, f.U.x.p, nullptr, r, pivcol.p};
    B.collective = dist().shard_rows;
    SolveEmit E;
    E.prefix_col = pivcol.p;
    E.prefix_val = 1;
    SolveResult R;
    solve_rows(G, B, E, f.F, R);
    for (int j = 0; j < m; j++) Rqinv[j] = fact->qinv[j];
    return result_to_host_csr(R, r, m, f.prime, f.F);
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_rref failed: %s\n", e.what());
    return nullptr;
  }
}

}  // extern "C"
