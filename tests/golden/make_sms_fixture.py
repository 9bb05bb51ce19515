"""Writes tests/golden/banded_60x70.sms the way the reference's pure-Julia SMS writer does
(/root/reference/src/SpaSM.jl:531-545: header "rows cols M", one "i j v" line per stored entry of the
SparseMatrixCSC in column-major order, 1-based, values as plain integers, terminator "0 0 0")."""
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT / "tests")]
import synth

n, m, r, prime = 60, 70, 41, 42013
p, j, x = synth.banded_planted(n, m, r, 3.0, 6, prime, 0x5A5A0009, spread=4, colblock=4)
A = sp.csr_matrix((x.astype(np.int64), j, p), shape=(n, m)).tocsc()
A.sort_indices()
lines = [f"{n} {m} M"]
for c in range(m):
    for e in range(A.indptr[c], A.indptr[c + 1]):
        lines.append(f"{A.indices[e] + 1} {c + 1} {int(A.data[e])}")
lines.append("0 0 0")
(ROOT / "tests" / "golden" / "banded_60x70.sms").write_text("\n".join(lines) + "\n")
print(len(lines) - 2, "entries")
