"""Block-diagonal splitter (src/blocks.jl): decomposition, reassembly, rank and kernel
through the blocks — on the CPU oracle here, on the CUDA library in test_gpu_echelonize.py."""
import numpy as np
import scipy.sparse as sp

import checks
import synth


def blocky_matrix(api, prime=42013, seed=3):
    """three independent random blocks + an empty row + an empty column, rows and columns shuffled"""
    rng = np.random.default_rng(seed)
    parts = []
    for (n, m, k) in [(30, 25, 3), (12, 20, 2), (40, 40, 4)]:
        p, j, x = synth.random_rows(n, m, k, prime, seed + n)
        parts.append(sp.csr_matrix((x.astype(np.int64), j, p), shape=(n, m)))
    M = sp.block_diag(parts + [sp.csr_matrix((1, 1))]).tocsr()
    M = M[rng.permutation(M.shape[0])][:, rng.permutation(M.shape[1])].tocsr()
    return api.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data, prime)


def check_blocks(api):
    A = blocky_matrix(api)
    B = api.Block(A)
    assert len(B) >= 5  # at least the 3 blocks, the empty row and the empty column (unused columns are blocks too)
    assert any(len(r) == 0 for r in B.block2row) and any(len(c) == 0 for c in B.block2col)
    assert sorted(np.concatenate(B.block2row).tolist()) == list(range(A.n))
    assert sorted(np.concatenate(B.block2col).tolist()) == list(range(A.m))
    assert A == api.block_CSR(B)  # reassembly (src/blocks.jl:143-170)
    fact = api.echelonize(A)
    assert api.block_rank(B) == fact.r  # src/blocks.jl:117
    KB = api.block_kernel(B)
    K = api.block_CSR(KB)
    assert K.shape == (A.m - fact.r, A.m)
    Ad, Kd = checks.dense_of(api, A), checks.dense_of(api, K)
    assert not checks.mm(Ad, Kd.T, A.prime).any()
    assert synth.dense_rank_mod_p(Kd, A.prime) == A.m - fact.r


def test_blocks_oracle(oracle):
    check_blocks(oracle)


def test_block_owner_partition(oracle):
    """blocks spread over ranks (second multi-GPU axis): every block has one owner, the partial ranks add up,
    and the heaviest rank carries at most the lightest one's load plus one block"""
    A = blocky_matrix(oracle)
    B = oracle.Block(A)
    full = oracle.echelonize(A).r
    for world in (1, 2, 3, 8):
        owner = oracle.block_owner(B, world)
        assert owner.min() >= 0 and owner.max() < world and len(owner) == len(B)
        assert sum(oracle.block_rank(B, part=(r, world)) for r in range(world)) == full
        w = np.array([max(Bk.nnz(), 1) for Bk in B.blocks])
        load = np.array([w[owner == r].sum() for r in range(world)])
        assert load.max() - load.min() <= w.max()
        fs = oracle.block_echelonize(B, part=(0, world))
        assert [f is not None for f in fs.blocks] == (owner == 0).tolist()
