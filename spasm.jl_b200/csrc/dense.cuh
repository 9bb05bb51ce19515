// dense.cuh — dense tail (dense.cu): replaces spasm_schur_dense + spasm_ffpack_rref
// (prototypes src/SpaSM.jl:765-766, :805; algorithm SURVEY.md A.7)
#pragma once
#include "factor.cuh"

namespace sb {

// in-place reduced row echelon form of the row-major n x m matrix S (u32 residues, leading
// dimension ld).  Step s finds pivot column pivcol[s] (increasing: column rank profile) held by
// row pivrow[s].  Returns the rank.
int dense_rref_device(uint32_t *S, int n, int m, long long ld, const Fp &F, DBuf<int> &pivcol, DBuf<int> &pivrow);

// eliminate the rows `rows` of A against U block by block, RREF each block, append to U
void echelonize_dense_device(const DCsr &A, const int *rows, int nrows, DCsr &U, DBuf<int> &Uqinv, const Fp &F, int block_size);

}  // namespace sb
