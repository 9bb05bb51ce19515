// spmv_solve.cu — SpMV (spasm_xApy / spasm_Axpy, src/SpaSM.jl:640-658) and the dense-RHS solves
// (spasm_dense_forward_solve / spasm_dense_back_solve / spasm_solve / spasm_gesv,
// src/SpaSM.jl:664-692, :895-923).  The solves reuse the row-solve engine: a dense right-hand side
// is just a (long) sparse row, x.U = b leaves the multipliers on the pivotal columns, and x.L = z is
// the same elimination on the rows of L that hold the "diagonal" entries, scaled to unit pivots.
#include <algorithm>

#include "dist.cuh"
#include "factor.cuh"

namespace sb {

template <bool SMALL>
__global__ void k_xApy(const long long *__restrict__ Ap, const int *__restrict__ Aj, const uint32_t *__restrict__ Ax, int n,
                       const uint32_t *__restrict__ x, uint32_t *__restrict__ y, Fp F) {
  // y += x.A : one warp per row scatters x[i]*A[i] with atomics on u32 residues is not exact mod p,
  // so accumulate per output through the transpose instead (called with A^T): y[i] += sum A^T[i,k] x[k]
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  uint32_t acc = 0;
  for (long long e = Ap[i] + lane; e < Ap[i + 1]; e += 32) acc = addmod(acc, mulmod<SMALL>(Ax[e], x[Aj[e]], F), F);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc = addmod(acc, __shfl_xor_sync(0xffffffffu, acc, o), F);
  if (lane == 0) y[i] = addmod(y[i], acc, F);
}

static void spmv(const DCsr &A, const spasm_ZZp *x, spasm_ZZp *y, const Fp &F) {
  // y[A.n] += A . x[A.m]
  DBuf<int> tx(std::max(A.m, 1)), ty(std::max(A.n, 1));
  DBuf<uint32_t> dx(std::max(A.m, 1)), dy(std::max(A.n, 1));
  tx.upload(x, A.m);
  ty.upload(y, A.n);
  convert_to_residues(tx.p, dx.p, A.m, F);
  convert_to_residues(ty.p, dy.p, A.n, F);
  if (A.n) {
    if (F.small)
      k_xApy<true><<<cdiv((long long)A.n * 32, 256), 256, 0, stream()>>>(A.p.p, A.j.p, A.x.p, A.n, dx.p, dy.p, F);
    else
      k_xApy<false><<<cdiv((long long)A.n * 32, 256), 256, 0, stream()>>>(A.p.p, A.j.p, A.x.p, A.n, dx.p, dy.p, F);
    CK(cudaGetLastError());
  }
  convert_to_balanced(dy.p, ty.p, A.n, F);
  ty.download(y, A.n);
  sync();
}

// one dense right-hand side through the engine; returns (column, value) of every nonzero of the solution
static void solve_one(const SolveSystem &G, const std::vector<uint32_t> &b, const Fp &F, std::vector<int> &cols,
                      std::vector<uint32_t> &vals) {
  std::vector<int> bj;
  std::vector<uint32_t> bx;
  for (size_t c = 0; c < b.size(); c++)
    if (b[c]) bj.push_back((int)c), bx.push_back(b[c]);
  long long bp[2] = {0, (long long)bj.size()};
  DBuf<long long> dp(2);
  DBuf<int> dj(std::max<size_t>(bj.size(), 1));
  DBuf<uint32_t> dx(std::max<size_t>(bx.size(), 1));
  dp.upload(bp, 2);
  if (!bj.empty()) dj.upload(bj.data(), bj.size()), dx.upload(bx.data(), bx.size());
  SolveRows B{dp.p, dj.p, dx.p, nullptr, 1, nullptr};
  SolveEmit E;
  E.all_columns = true;
  SolveResult R;
  solve_rows(G, B, E, F, R);
  cols.resize(R.nnz), vals.resize(R.nnz);
  if (R.nnz) R.j.download(cols.data(), R.nnz), R.x.download(vals.data(), R.nnz);
  sync();
}

// device form of "x.L = z": rows of L holding the diagonals, scaled to unit pivots
struct LSystem {
  DCsr L;
  DBuf<uint32_t> scaled;   // values of L with each diagonal row divided by its diagonal
  DBuf<uint32_t> dinv;     // [r] inverse of the diagonal of column k
  DBuf<PDesc> pdesc;       // [r]
  DBuf<int> diagrow;       // [r]
  int r = 0;
};
template <bool SMALL>
__global__ void k_Lsystem(const long long *__restrict__ Lp, const int *__restrict__ Lj, const uint32_t *__restrict__ Lx,
                          const int *__restrict__ diagrow, int r, uint32_t *__restrict__ scaled, uint32_t *__restrict__ dinv,
                          PDesc *__restrict__ pd, Fp F) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= r) return;
  const int i = diagrow[k];
  const long long a = Lp[i], b = Lp[i + 1];
  uint32_t d = 0;
  for (long long e0 = a; e0 < b; e0 += 32) {
    long long e = e0 + lane;
    unsigned hit = __ballot_sync(0xffffffffu, e < b && Lj[e] == k);
    if (hit) {
      d = Lx[e0 + (__ffs(hit) - 1)];
      break;
    }
  }
  const uint32_t inv = dev_inv(d, F.p);
  for (long long e = a + lane; e < b; e += 32) scaled[e] = mulmod<SMALL>(inv, Lx[e], F);
  if (lane == 0) {
    dinv[k] = inv;
    PDesc p;
    p.start = a, p.len = (int)(b - a), p.prio = r - 1 - k;
    pd[k] = p;
  }
}
static void build_Lsystem(const spasm_csr *L, const int *p, const Fp &F, LSystem &S) {
  upload_csr(L, S.L, F);
  S.r = L->m;
  const int r = S.r;
  S.scaled.alloc(std::max<int64_t>(S.L.nnz, 1));
  CK(cudaMemcpyAsync(S.scaled.p, S.L.x.p, (size_t)S.L.nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream()));
  S.dinv.alloc(std::max(r, 1));
  S.pdesc.alloc(std::max(r, 1));
  S.diagrow.alloc(std::max(r, 1));
  std::vector<int> dr(std::max(r, 1));
  for (int k = 0; k < r; k++) dr[k] = p ? p[k] : k;
  S.diagrow.upload(dr.data(), r);
  if (r) {
    if (F.small)
      k_Lsystem<true><<<cdiv((long long)r * 32, 256), 256, 0, stream()>>>(S.L.p.p, S.L.j.p, S.L.x.p, S.diagrow.p, r, S.scaled.p, S.dinv.p, S.pdesc.p, F);
    else
      k_Lsystem<false><<<cdiv((long long)r * 32, 256), 256, 0, stream()>>>(S.L.p.p, S.L.j.p, S.L.x.p, S.diagrow.p, r, S.scaled.p, S.dinv.p, S.pdesc.p, F);
    CK(cudaGetLastError());
  }
}

// z.U = b.  Returns false when b is not in the row space.  z has r entries.
static bool forward(const DevFactor &f, const DBuf<PDesc> &pdesc, const std::vector<int> &qinv_h, const spasm_ZZp *b,
                    std::vector<uint32_t> &z) {
  const int m = f.U.m, r = f.U.n;
  std::vector<uint32_t> bu(m);
  for (int j = 0; j < m; j++) bu[j] = to_u(b[j], f.F);
  SolveSystem G{f.U.j.p, f.U.x.p, pdesc.p, m};
  std::vector<int> cols;
  std::vector<uint32_t> vals;
  solve_one(G, bu, f.F, cols, vals);
  z.assign(std::max(r, 1), 0);
  bool ok = true;
  for (size_t t = 0; t < cols.size(); t++) {
    int i = qinv_h[cols[t]];
    if (i < 0)
      ok = false;
    else
      z[i] = vals[t];
  }
  return ok;
}
// x.L = z
static void backward(const LSystem &S, const std::vector<int> &diagrow_h, const Fp &F, const std::vector<uint32_t> &z, int n, spasm_ZZp *x) {
  SolveSystem G{S.L.j.p, S.scaled.p, S.pdesc.p, S.r};
  std::vector<uint32_t> zz(z.begin(), z.begin() + S.r);
  std::vector<int> cols;
  std::vector<uint32_t> vals;
  solve_one(G, zz, F, cols, vals);
  std::vector<uint32_t> dinv(std::max(S.r, 1));
  if (S.r) S.dinv.download(dinv.data(), S.r);
  sync();
  for (int i = 0; i < n; i++) x[i] = 0;
  for (size_t t = 0; t < cols.size(); t++) {
    const int k = cols[t];
    uint64_t v = (uint64_t)vals[t] * dinv[k] % F.p;
    x[diagrow_h[k]] = to_bal((uint32_t)v, F);
  }
}

// batched gesv helpers.  okflag[k] = 1 iff row k of the forward solve has no entry on a non-pivotal column; its column
// indices are replaced by U row indices on the fly (the second solve reads them as a row over the U rows)
__global__ void k_gesv_check(const long long *__restrict__ Zp, int *__restrict__ Zj, int nb, const int *__restrict__ qinv, int *__restrict__ okflag) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k > nb) return;
  if (k == nb) {
    if (lane == 0) okflag[nb] = 0;
    return;
  }
  int bad = 0;
  for (long long e = Zp[k] + lane; e < Zp[k + 1]; e += 32) {
    const int i = qinv[Zj[e]];
    if (i < 0)
      bad = 1;
    else
      Zj[e] = i;
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) okflag[k] = !bad;
}
__global__ void k_gesv_list(const int *__restrict__ okflag, const long long *__restrict__ pos, int nb, int *__restrict__ list) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nb && okflag[k]) list[pos[k]] = k;
}
template <bool SMALL>
__global__ void k_gesv_finish(int *__restrict__ Xj, uint32_t *__restrict__ Xx, long long nnz, const int *__restrict__ diagrow,
                              const uint32_t *__restrict__ dinv, Fp F) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int k = Xj[e];
  Xx[e] = mulmod<SMALL>(Xx[e], dinv[k], F);
  Xj[e] = diagrow[k];
}

}  // namespace sb

using namespace sb;

extern "C" {

void spasm_xApy(const spasm_ZZp *x, const struct spasm_csr *A, spasm_ZZp *y) {
  try {
    ApiCall api_scope_;
    Fp F = make_field(A->field->p);
    DCsr dA, dT;
    upload_csr(A, dA, F);
    transpose_csr(dA, dT);
    spmv(dT, x, y, F);
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_xApy failed: %s\n", e.what());
  }
}
void spasm_Axpy(const struct spasm_csr *A, const spasm_ZZp *x, spasm_ZZp *y) {
  try {
    ApiCall api_scope_;
    Fp F = make_field(A->field->p);
    DCsr dA;
    upload_csr(A, dA, F);
    spmv(dA, x, y, F);
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_Axpy failed: %s\n", e.what());
  }
}

bool spasm_dense_forward_solve(const struct spasm_csr *U, spasm_ZZp *b, spasm_ZZp *x, const int *q) {
  try {
    ApiCall api_scope_;
    const int m = U->m, r = U->n;
    DevFactor f;
    f.prime = U->field->p;
    f.F = make_field(f.prime);
    upload_csr(U, f.U, f.F);
    std::vector<int> qinv(m, -1);
    for (int i = 0; i < r; i++) qinv[q[i]] = i;
    f.qinv.alloc(std::max(m, 1));
    f.qinv.upload(qinv.data(), m);
    DBuf<PDesc> pdesc;
    build_pdesc_U(f.U, f.qinv.p, pdesc);
    std::vector<uint32_t> z;
    bool ok = forward(f, pdesc, qinv, b, z);
    for (int i = 0; i < r; i++) x[i] = to_bal(z[i], f.F);
    return ok;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_dense_forward_solve failed: %s\n", e.what());
    return false;
  }
}
bool spasm_dense_back_solve(const struct spasm_csr *L, spasm_ZZp *b, spasm_ZZp *x, const int *p) {
  try {
    ApiCall api_scope_;
    Fp F = make_field(L->field->p);
    LSystem S;
    build_Lsystem(L, p, F, S);
    std::vector<int> dr(std::max(S.r, 1));
    for (int k = 0; k < S.r; k++) dr[k] = p ? p[k] : k;
    std::vector<uint32_t> z(std::max(S.r, 1));
    for (int k = 0; k < S.r; k++) z[k] = to_u(b[k], F);
    backward(S, dr, F, z, L->n, x);
    return true;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_dense_back_solve failed: %s\n", e.what());
    return false;
  }
}

bool spasm_solve(const struct spasm_lu *fact, const spasm_ZZp *b, spasm_ZZp *x) {
  try {
    ApiCall api_scope_;
    if (fact->L == nullptr) throw Error("spasm_solve needs a factorisation computed with L=true");
    DevFactor f;
    f.upload(fact);
    const int m = f.U.m, r = f.U.n;
    std::vector<int> qinv(fact->qinv, fact->qinv + m);
    DBuf<PDesc> pdesc;
    build_pdesc_U(f.U, f.qinv.p, pdesc);
    std::vector<uint32_t> z;
    if (!forward(f, pdesc, qinv, b, z)) return false;
    LSystem S;
    build_Lsystem(fact->L, fact->p, f.F, S);
    std::vector<int> dr(fact->p, fact->p + r);
    backward(S, dr, f.F, z, fact->L->n, x);
    return true;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_solve failed: %s\n", e.what());
    return false;
  }
}

// ALL right-hand sides at once (src/SpaSM.jl:915-923): the rows of B go through the row engine in one batch for
// z.U = b, the solvable ones in a second batch for x.L = z; nothing but ok[] and the result crosses PCIe.
struct spasm_csr *spasm_gesv(const struct spasm_lu *fact, const struct spasm_csr *B, bool *ok) {
  try {
    ApiCall api_scope_;
    if (fact->L == nullptr) throw Error("spasm_gesv needs a factorisation computed with L=true");
    DevFactor f;
    f.upload(fact);
    const int m = f.U.m, r = f.U.n, n = fact->L->n, nb = B->n;
    cudaStream_t s = stream();
    DBuf<PDesc> pdesc;
    build_pdesc_U(f.U, f.qinv.p, pdesc);
    LSystem S;
    build_Lsystem(fact->L, fact->p, f.F, S);
    DCsr dB;
    upload_csr(B, dB, f.F);
    // ---- z.U = b for every row of B
    SolveSystem GU{f.U.j.p, f.U.x.p, pdesc.p, m};
    SolveRows RB{dB.p.p, dB.j.p, dB.x.p, nullptr, nb, nullptr};
    RB.collective = dist().shard_rows;  // right-hand sides split over the ranks ("many RHS -> shard RHS")
    SolveEmit Em;
    Em.all_columns = true;
    SolveResult Z;
    solve_rows(GU, RB, Em, f.F, Z);
    // solvable <=> nothing is left on a non-pivotal column; the multipliers become a row over the U rows
    DBuf<int> okflag(std::max(nb, 1) + 1), oklist(std::max(nb, 1));
    DBuf<long long> okpos(std::max(nb, 1) + 1);
    k_gesv_check<<<cdiv((long long)nb * 32 + 32, 256), 256, 0, s>>>(Z.p.p, Z.j.p, nb, f.qinv.p, okflag.p);
    exclusive_scan_i32_to_i64(okflag.p, okpos.p, nb + 1);
    const int nok = (int)fetch(okpos.p + nb);
    std::vector<int> hok(std::max(nb, 1));
    if (nb) okflag.download(hok.data(), nb);
    if (nok) k_gesv_list<<<cdiv(nb, 256), 256, 0, s>>>(okflag.p, okpos.p, nb, oklist.p);
    sync();
    for (int k = 0; k < nb; k++) ok[k] = hok[k] != 0;
    // ---- x.L = z for the solvable ones
    SolveResult X;
    if (nok > 0) {
      SolveSystem GL{S.L.j.p, S.scaled.p, S.pdesc.p, S.r};
      SolveRows RZ{Z.p.p, Z.j.p, Z.x.p, oklist.p, nok, nullptr};
      RZ.collective = dist().shard_rows;
      solve_rows(GL, RZ, Em, f.F, X);
      // (k, v) -> (row of A that holds the k-th diagonal, v / diagonal), then by increasing row index
      if (X.nnz) {
        if (f.F.small)
          k_gesv_finish<true><<<cdiv(X.nnz, 256), 256, 0, s>>>(X.j.p, X.x.p, X.nnz, S.diagrow.p, S.dinv.p, f.F);
        else
          k_gesv_finish<false><<<cdiv(X.nnz, 256), 256, 0, s>>>(X.j.p, X.x.p, X.nnz, S.diagrow.p, S.dinv.p, f.F);
        sort_csr_rows(X.p.p, nok, X.j.p, X.x.p);
      }
    }
    // ---- host CSR: one row per right-hand side (empty when not solvable)
    std::vector<long long> xp(nok + 1, 0);
    if (nok > 0) X.p.download(xp.data(), nok + 1);
    sync();
    const long long xnz = nok > 0 ? xp[nok] : 0;
    spasm_csr *R = spasm_csr_alloc(nb, n, xnz, B->field->p, true);
    if (xnz) {
      X.j.download(R->j, xnz);
      convert_to_balanced(X.x.p, (int *)X.x.p, xnz, f.F);
      CK(cudaMemcpyAsync(R->x, X.x.p, (size_t)xnz * sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    sync();
    // rows of solvable systems are consecutive in X: row k of R starts where its X row starts
    int t = 0;
    for (int k = 0; k < nb; k++) {
      R->p[k] = xp[t];
      if (hok[k]) t++;
    }
    R->p[nb] = xp[nok];
    return R;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_gesv failed: %s\n", e.what());
    return nullptr;
  }
}

}  // extern "C"
