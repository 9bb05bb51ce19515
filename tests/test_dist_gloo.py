"""N>1 host logic on CPU: world_size-2 gloo processes check the block-cyclic sharding of the dense
tail (the pure host functions the CUDA library uses, exported through the C ABI) and the bootstrap
plumbing bench.py uses to ship the 128-byte NCCL id (here a fake id, broadcast over gloo)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import ctypes as C, os, sys
sys.path[:0] = [".", "tests"]
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as e
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = C.CDLL(str(e.build_product()), mode=os.RTLD_LOCAL)
lib.spasm_b200_local_positions.restype = C.c_longlong
lib.spasm_b200_local_positions.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
for (n_rem, block) in [(10, 3), (1000, 1000), (2501, 1000), (7, 100), (0, 5), (4096, 64)]:
    buf = np.zeros(max(n_rem, 1), dtype=np.int32)
    cnt = lib.spasm_b200_local_positions(n_rem, block, world, rank, buf.ctypes.data, len(buf))
    mine = buf[:cnt].copy()
    # every position belongs to the owner of its panel, in increasing order
    assert all(lib.spasm_b200_panel_owner(int(k) // block, world) == rank for k in mine)
    assert (np.diff(mine) > 0).all()
    # all ranks together cover every remaining row exactly once
    t = torch.zeros(max(n_rem, 1), dtype=torch.int64)
    t[torch.from_numpy(mine.astype(np.int64))] += 1 if cnt else 0
    if cnt:
        t.zero_(); t[torch.from_numpy(mine.astype(np.int64))] = 1
    dist.all_reduce(t)
    assert (t[:n_rem] == 1).all(), (n_rem, block, t)
    # load balance: sizes differ by at most one panel
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([cnt]))
    assert max(int(s) for s in sizes) - min(int(s) for s in sizes) <= block
# rows of the row engine split over the ranks (Schur complement / kernel / rref / gesv): contiguous shares that tile the row
# list in rank order; "counts first, then the payload" reassembles the single-process result (here with the oracle's Schur
# complement as the per-row work: each rank computes the rows of its share, the pieces are all-gathered in rank order)
lib.spasm_b200_row_share.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
for nrows in (0, 1, 2, 7, 1000, 1001):
    lo, hi = C.c_longlong(0), C.c_longlong(0)
    lib.spasm_b200_row_share(nrows, world, rank, C.byref(lo), C.byref(hi))
    ends = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(ends, torch.tensor([lo.value, hi.value]))
    assert int(ends[0][0]) == 0 and int(ends[-1][1]) == nrows
    assert all(int(ends[r][1]) == int(ends[r + 1][0]) for r in range(world - 1))
    assert max(int(e_[1] - e_[0]) for e_ in ends) - min(int(e_[1] - e_[0]) for e_ in ends) <= (nrows + world - 1) // world
import synth
pkg = e.load_package()
ora = pkg.SpaSM(e.build_oracle())
ora.log(False)
n, m = 600, 700
p_, j_, x_ = synth.random_rows(n, m, 4, 42013, 99)
A = ora.from_arrays(n, m, p_, j_, x_, 42013)
from test_gpu_engine import structural_round
lu, U_, qinv_, perm, npiv = structural_round(pkg, ora, A)
rows = np.ascontiguousarray(perm[npiv:], dtype=np.int32)
def schur_of(sub):
    sub = np.ascontiguousarray(sub, dtype=np.int32)
    pout = np.zeros(max(len(sub), 1), dtype=np.int32)
    return pkg.CSR(ora, ora.lib.spasm_schur(A.data, sub.ctypes.data_as(C.POINTER(C.c_int32)), len(sub), C.byref(lu), 0.0, None, None,
                                           pout.ctypes.data_as(C.POINTER(C.c_int32))))
whole = schur_of(rows)
lo, hi = C.c_longlong(0), C.c_longlong(0)
lib.spasm_b200_row_share(len(rows), world, rank, C.byref(lo), C.byref(hi))
mine = schur_of(rows[lo.value:hi.value])
Sp, Sj, Sx = mine.arrays()
cnt = torch.from_numpy(np.diff(Sp).astype(np.int64))
per = (len(rows) + world - 1) // world
pad = torch.zeros(per, dtype=torch.int64); pad[:len(cnt)] = cnt
allcnt = [torch.zeros(per, dtype=torch.int64) for _ in range(world)]
dist.all_gather(allcnt, pad)
counts = torch.cat(allcnt)[:len(rows)].numpy()
gp = np.concatenate([[0], np.cumsum(counts)])
pieces_j, pieces_x = [None] * world, [None] * world
dist.all_gather_object(pieces_j, Sj.copy()); dist.all_gather_object(pieces_x, Sx.copy())
Wp, Wj, Wx = whole.arrays()
assert np.array_equal(gp, Wp) and np.array_equal(np.concatenate(pieces_j), Wj) and np.array_equal(np.concatenate(pieces_x), Wx)
# independent blocks (src/blocks.jl) spread over the ranks: partial ranks add up to the rank of the whole matrix
from test_blocks import blocky_matrix
pkg = e.load_package()
ora = pkg.SpaSM(e.build_oracle())
A = blocky_matrix(ora)
B = ora.Block(A)
part = torch.tensor([ora.block_rank(B, part=(rank, world))])
dist.all_reduce(part)
assert int(part) == ora.echelonize(A).r
# bootstrap plumbing: rank 0's 128-byte id reaches everyone unchanged
ident = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(ident, src=0)
assert ident.tolist() == list(range(128))
dist.destroy_process_group()
print("ok", rank)
'''


def test_sharding_and_bootstrap_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
