"""Writes tests/golden/reference_goldens.json from the literal vectors in the reference's own
tests (/root/reference/test/runtests.jl:3-24) and README transcript (/root/reference/README.md:9-48).
Nothing is computed here: the numbers are transcribed, so the file pins the oracle to the reference."""
import json
from pathlib import Path

goldens = {
    "runtests": {  # test/runtests.jl:3,5
        "prime": 42013,
        "I": [1, 1, 3, 3], "J": [1, 2, 3, 4], "V": [1, 2, 3, 4],
        # :21  sparse(k) == sparse([2],[1],ZZp{F}[42012],3,1)
        "kernel": {"I": [2], "J": [1], "V": [42012], "shape": [3, 1]},
        # :23  sparse(kernel(transpose(sm))) == sparse([1,2,3,4],[1,1,2,2],ZZp{F}[2,42012,28010,42012])
        "kernel_transpose": {"I": [1, 2, 3, 4], "J": [1, 1, 2, 2], "V": [2, 42012, 28010, 42012], "shape": [4, 2]},
    },
    "readme": {  # README.md:12-48
        "prime": 42013,
        "I": [1, 1, 2, 2], "J": [1, 2, 1, 2], "V": [1, 2, 3, 6],
        "rank": 1, "nz_in_basis": 2, "nnz_K": 2,
        "kernel": {"I": [1, 2], "J": [1, 1], "V": [3, 42012], "shape": [2, 1]},
        "log_must_contain": [
            "[echelonize] Start on 2 x 2 matrix with 4 nnz",
            "[echelonize] round 0",
            "[pivots] Faugère-Lachartre: 1 pivots found",
            "[pivots] ``Faugère-Lachartre on columns'': 0 pivots found",
            "[pivots] greedy alternating cycle-free search: 0 pivots found",
            "[pivots] 1 pivots found",
            "Schur complement is 1 x 1, estimated density : 0.00",
            "Schur complement: 1 * 2 [0 nz / density= 0.000]",
            "[echelonize] not enough pivots found; stopping",
            "[echelonize] finishing; density = 0.000; aspect ratio = 0.5",
            "[echelonize/GPLU] processing matrix of dimension 1 x 2",
            "Rank 1, 2 nz in basis",
            "[kernel] start. U is 1 x 2 (2 nnz). Transposing U",
            "kernel: 1/1, |K| = 2",
            "NNZ(K) = 2",
        ],
    },
}
Path(__file__).with_name("reference_goldens.json").write_text(json.dumps(goldens, indent=1) + "\n")
