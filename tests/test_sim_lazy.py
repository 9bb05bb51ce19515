"""Orchestration of the sharded dense tail with deferred trailing updates (model in sim_lazy_updates.py, a line-by-line
numpy restatement of dense_tail_core's host logic): every rank count and flush depth gives the same pivots and the same
reduced rows as the eager single-rank elimination."""
import numpy as np
import pytest

import sim_lazy_updates as sim


def make(nrows, Sm0, p, seed, dependent=0):
    rng = np.random.default_rng(seed)
    D = rng.integers(0, p, size=(nrows, Sm0), dtype=np.int64)
    for k in range(dependent):  # rows that are combinations of others: panels of deficient rank
        i, a, b = rng.integers(0, nrows, size=3)
        D[i] = (3 * D[a] + 5 * D[b]) % p
    return D


@pytest.mark.parametrize("NR", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("kcap", [32, 64, 100, 4096])
def test_lazy_equals_eager(NR, kcap):
    p, bs = 251, 8
    for (nrows, Sm0, dep, seed) in [(150, 170, 0, 1), (203, 120, 25, 2), (97, 97, 10, 3)]:
        D = make(nrows, Sm0, p, seed, dep)
        ref, _ = sim.run(D, p, bs, 1, 0, lazy_enabled=False)
        got, st = sim.run(D, p, bs, NR, kcap)
        assert len(ref) == len(got)
        for (pc0, R0), (pc1, R1) in zip(ref, got):
            assert pc0 == pc1 and np.array_equal(R0, R1)
        if kcap == 32 and NR <= 2 and nrows >= 150:
            assert any(s["ncorr"] > 0 and s["nflush"] > 0 for s in st), "the far path was not exercised"


@pytest.mark.parametrize("NR", [1, 2, 4, 8])
def test_far_path_is_exercised_on_every_rank(NR):
    """a case large enough that every rank defers, corrects and flushes (depth 32 = 4 panels of 8 rows)"""
    p, bs = 251, 8
    D = make(400, 300, p, 7, dependent=30)
    ref, _ = sim.run(D, p, bs, 1, 0, lazy_enabled=False)
    got, st = sim.run(D, p, bs, NR, 32)
    for (pc0, R0), (pc1, R1) in zip(ref, got):
        assert pc0 == pc1 and np.array_equal(R0, R1)
    assert all(s["lazy"] and s["ncorr"] > 5 and s["nflush"] > 5 for s in st)
