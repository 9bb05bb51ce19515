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
dist.barrier()
gpu.lib.spasm_b200_dist_finalize()
dist.destroy_process_group()
if rank == 0:
    print(f"dist check OK on {world} GPUs: {len(cases)} cases bit-exact (single GPU == sharded == oracle)")
