// spmv_solve.cu — SpMV (spasm_xApy / spasm_Axpy, src/SpaSM.jl:640-658) and the dense-RHS solves
// (spasm_dense_forward_solve / spasm_dense_back_solve / spasm_solve / spasm_gesv,
// src/SpaSM.jl:664-692, :895-923).  The solves reuse the row-solve engine: a dense right-hand side
// is just a (long) sparse row, x.U = b leaves the multipliers on the pivotal columns, and x.L = z is
// the same elimination on the rows of L that hold the "diagonal" entries, scaled to unit pivots.
#include <algorithm>

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

struct spasm_csr *spasm_gesv(const struct spasm_lu *fact, const struct spasm_csr *B, bool *ok) {
  try {
    ApiCall api_scope_;
    if (fact->L == nullptr) throw Error("spasm_gesv needs a factorisation computed with L=true");
    DevFactor f;
    f.upload(fact);
    const int m = f.U.m, r = f.U.n, n = fact->L->n;
    std::vector<int> qinv(fact->qinv, fact->qinv + m);
    DBuf<PDesc> pdesc;
    build_pdesc_U(f.U, f.qinv.p, pdesc);
    LSystem S;
    build_Lsystem(fact->L, fact->p, f.F, S);
    std::vector<int> dr(fact->p, fact->p + r);
    std::vector<long long> Xp(B->n + 1, 0);
    std::vector<int> Xj;
    std::vector<spasm_ZZp> Xx;
    std::vector<spasm_ZZp> b(m), x(std::max(n, 1));
    for (int k = 0; k < B->n; k++) {
      std::fill(b.begin(), b.end(), 0);
      for (i64 e = B->p[k]; e < B->p[k + 1]; e++) b[B->j[e]] = B->x[e];
      std::vector<uint32_t> z;
      ok[k] = forward(f, pdesc, qinv, b.data(), z);
      if (ok[k]) {
        backward(S, dr, f.F, z, n, x.data());
        for (int i = 0; i < n; i++)
          if (x[i] != 0) Xj.push_back(i), Xx.push_back(x[i]);
      }
      Xp[k + 1] = (long long)Xj.size();
    }
    spasm_csr *X = spasm_csr_alloc(B->n, n, (i64)Xj.size(), B->field->p, true);
    for (int k = 0; k <= B->n; k++) X->p[k] = Xp[k];
    if (!Xj.empty()) memcpy(X->j, Xj.data(), Xj.size() * sizeof(int)), memcpy(X->x, Xx.data(), Xx.size() * sizeof(spasm_ZZp));
    return X;
  } catch (const std::exception &e) {
    errf("[spasm_b200] spasm_gesv failed: %s\n", e.what());
    return nullptr;
  }
}

}  // extern "C"
