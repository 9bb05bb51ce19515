"""Run under torchrun with N>=2 GPUs: the sharded dense tail must give rank 0 the SAME factor, bit for
bit, as the single-GPU library gave (file written by rank 0 before the communicator exists) and as
the CPU oracle gives.   torchrun --nproc-per-node 2 tests/dist_gpu_check.py"""
import ctypes as C
import os
import sys

sys.path[:0] = [".", "tests"]
import numpy as np
import torch
import torch.distributed as dist

import __graft_entry__ as e
import bench
import checks
import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = e.load_package()
gpu = pkg.SpaSM()
gpu.log(False)
cases = [(1500, 1500, 5, 42013, 3, {}), (2600, 2400, 6, 65521, 4, dict(dense_block_size=300)), (900, 1000, 4, 4294967291, 5, dict(dense_block_size=128)),
         (5000, 5000, 10, 42013, 6, {}),
         # the dense loop switches to the low-rank mode with the remaining rows spread over the ranks
         (2000, 2000, 3, 42013, 4, dict(dense_block_size=100, sparsity_threshold=0.0, max_round=0), 300),
         (1200, 1000, 4, 65521, 5, dict(dense_block_size=64)),
         (1500, 1400, 3, 4294967291, 8, dict(dense_block_size=50, sparsity_threshold=0.0, max_round=0, low_rank_start_weight=2), 200)]


def make_input(case):
    n, m, k, prime, seed = case[:5]
    if len(case) > 6:
        return synth.planted_rank(n, m, case[6], 0.5, prime, seed)
    return synth.random_rows(n, m, k, prime, seed)


single = []
for case in cases:  # single-GPU results first (no communicator yet)
    n, m, k, prime, seed, kw = case[:6]
    p, j, x = make_input(case)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    single.append(checks.lu_arrays(gpu.echelonize(A, **kw)))
# second axis (src/blocks.jl): independent blocks, one owner rank each, no data-path collective
from test_blocks import blocky_matrix

Ab = blocky_matrix(gpu)
Bb = gpu.Block(Ab)
part = torch.tensor([gpu.block_rank(Bb, part=(rank, world))], device="cuda")
dist.all_reduce(part)
assert int(part.item()) == gpu.echelonize(Ab).r, (rank, int(part.item()))
bench.dist_init(gpu.lib, dist, rank, world)
ok = True
for idx, case in enumerate(cases):
    n, m, k, prime, seed, kw = case[:6]
    p, j, x = make_input(case)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    f = gpu.echelonize(A, **kw)
    got = checks.lu_arrays(f)
    assert got["r"] == single[idx]["r"], (rank, got["r"], single[idx]["r"])
    assert np.array_equal(got["qinv"], single[idx]["qinv"])
    if rank == 0:
        checks.assert_same(single[idx], got, f"case {idx}: ")
        ora = pkg.SpaSM(e.build_oracle())
        checks.assert_same(checks.lu_arrays(ora.echelonize(A, **kw)), got, f"case {idx} vs oracle: ")
        K1, K2 = gpu.kernel(f), ora.kernel(ora.echelonize(A, **kw))
        for a, b in zip(K1.arrays(), K2.arrays()):
            assert np.array_equal(a, b)
# ---- sharded factor: every rank materialises the rows of the dense panels it owns (plus the structural rows);
# together the ranks hold exactly the single-GPU factor, and kernel / solve refuse a partial factor
gpu.lib.spasm_b200_dist_shard_factor.argtypes = [C.c_int]
gpu.lib.spasm_b200_dist_shard_factor(1)
for idx, case in enumerate(cases):
    n, m, k, prime, seed, kw = case[:6]
    p, j, x = make_input(case)
    A = gpu.from_arrays(n, m, p, j, x, prime)
    f = gpu.echelonize(A, **kw)
    assert f.partial and f.r == single[idx]["r"]
    Up, Uj, Ux = f.U.arrays()
    sp, sj, sx = single[idx]["Up"], single[idx]["Uj"], single[idx]["Ux"]
    have = np.diff(Up) > 0
    for i in np.nonzero(have)[0]:
        assert np.array_equal(Uj[Up[i] : Up[i + 1]], sj[sp[i] : sp[i + 1]]) and np.array_equal(Ux[Up[i] : Up[i + 1]], sx[sp[i] : sp[i + 1]]), (rank, idx, i)
    cover = torch.tensor(have.astype(np.int32), device="cuda")
    dist.all_reduce(cover)
    assert int(cover.min().item()) >= 1, f"case {idx}: some row of U is held by no rank"
    try:
        gpu.kernel(f)
        raise SystemExit("kernel() accepted a partial factor")
    except (RuntimeError, ValueError, AssertionError):
        pass
gpu.lib.spasm_b200_dist_shard_factor(0)

# ---- rows of the row engine split over the ranks: the sparse Schur complements inside echelonize
# (always), kernel / rref / gesv after shard_rows(1).  Every rank must end up with the single-GPU result, bit for bit.
gpu.lib.spasm_b200_shard_stats.argtypes = [C.c_void_p, C.c_int]
gpu.lib.spasm_b200_dist_shard_rows.argtypes = [C.c_int]


def shard_stats(reset=True):
    out = (C.c_longlong * 2)()
    gpu.lib.spasm_b200_shard_stats(out, 1 if reset else 0)
    return out[0], out[1]


s_ = 128
ns, ms_, rs = 1911130 // s_, 1955309 // s_, 1033568 // s_
ps, js, xs = synth.banded_planted(ns, ms_, rs, 12.0, 40, 42013, 0x5A5A0003, spread=16, colblock=8)
As = gpu.from_arrays(ns, ms_, ps, js, xs, 42013)
ora = pkg.SpaSM(e.build_oracle())
for kw in (dict(), dict(L=True)):
    shard_stats()
    f = gpu.echelonize(As, **kw)
    calls, mine = shard_stats()
    assert calls >= 1 and 0 < mine, (rank, calls, mine)  # the Schur complement ran split over the ranks
    assert f.r == rs and not f.partial  # no dense tail: every rank holds the complete factor
    fo = ora.echelonize(As, **kw)
    checks.assert_same(checks.lu_arrays(fo), checks.lu_arrays(f), f"rank {rank}, sharded Schur {kw}: ")
    if not kw:
        def rref_(fact):
            q = np.zeros(ms_, dtype=np.int32)
            return gpu.rref(fact, q), q

        K_local, (R_local, q_local) = gpu.kernel(f), rref_(f)
        gpu.lib.spasm_b200_dist_shard_rows(1)
        shard_stats()
        K_sh, (R_sh, q_sh) = gpu.kernel(f), rref_(f)
        calls, mine = shard_stats()
        assert calls == 2 and mine < (ms_ - rs) + rs, (rank, calls, mine)
        gpu.lib.spasm_b200_dist_shard_rows(0)
        for a, b in zip(list(K_local.arrays()) + list(R_local.arrays()) + [q_local], list(K_sh.arrays()) + list(R_sh.arrays()) + [q_sh]):
            assert np.array_equal(a, b), f"rank {rank}: sharded kernel / rref differ from the local ones"
        for a, b in zip(K_sh.arrays(), ora.kernel(fo).arrays()):
            assert np.array_equal(a, b)
    else:
        # 24 right-hand sides: 20 in the row space, 4 random; gesv with the right-hand sides split over the ranks
        rng = np.random.default_rng(7)
        rowsB = []
        Ad = synth.csr_to_dense(ns, ms_, ps, js, xs, 42013).astype(np.int64)
        for t in range(24):
            if t < 20:
                xv = rng.integers(0, 42013, size=ns)
                xv[rng.random(ns) < 0.995] = 0
                rowsB.append((xv @ Ad) % 42013)
            else:
                rowsB.append(rng.integers(0, 42013, size=ms_))
        Bd = np.array(rowsB)
        bp = np.zeros(25, dtype=np.int64)
        bj, bx = [], []
        for t in range(24):
            nzc = np.nonzero(Bd[t])[0]
            bj.append(nzc), bx.append(synth.balanced(Bd[t][nzc], 42013))
            bp[t + 1] = bp[t] + len(nzc)
        Bm = gpu.from_arrays(24, ms_, bp, np.concatenate(bj).astype(np.int32), np.concatenate(bx).astype(np.int32), 42013)
        X_local, ok_local = gpu.gesv(f, Bm)
        gpu.lib.spasm_b200_dist_shard_rows(1)
        shard_stats()
        X_sh, ok_sh = gpu.gesv(f, Bm)
        calls, mine = shard_stats()
        assert calls == 2, (rank, calls, mine)
        gpu.lib.spasm_b200_dist_shard_rows(0)
        assert list(ok_local) == list(ok_sh) and sum(ok_sh) >= 20
        for a, b in zip(X_local.arrays(), X_sh.arrays()):
            assert np.array_equal(a, b), f"rank {rank}: sharded gesv differs from the local one"
dist.barrier()
gpu.lib.spasm_b200_dist_finalize()
dist.destroy_process_group()
if rank == 0:
    print(f"dist check OK on {world} GPUs: {len(cases)} cases bit-exact (single GPU == sharded == oracle); Schur / kernel / rref / gesv rows split over the ranks")
