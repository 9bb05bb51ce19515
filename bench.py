#!/usr/bin/env python
"""bench.py — echelonize time-to-rank (s) on BASELINE.json configs[1]:
synthetic 200 000 x 200 000, 10 nnz/row, echelonize + rank mod 42013, on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # CPU oracle restatement of libspasm, all host threads

One "step" is one complete echelonization of the matrix (structural pivots -> dense Schur
complement -> blocked RREF on the tcgen05 tensor cores).  `value` is the time with the input CSR
already resident in HBM and the factor left on the device; `e2e` is the same call through the
reference-facing C ABI (`spasm_echelonize` on host structs: CSR upload and the whole factor U
downloaded into malloc'd host arrays inside the timed region).

`roofline` is the tensor-core kernel of the dense tail (`k_gemm_i8limb`): `achieved` from CUDA events around every launch of
the timed region, `peak` the back-to-back `UTCIMMA` rate measured on this box in this run, `alone` the same kernel by itself at
one mid-elimination shape (with several ranks the in-step launches overlap the main stream's work on a capped number of SMs).
`secondary` (1 GPU): configs[3] through the dense-tail entry point, a GL7d19-shaped instance with its Schur GB/s,
configs[4] at 1/100 (echelonize with L, kernel basis, 20 right-hand sides).

The real reference (SpaSM.jl -> libspasm) cannot run here or on the GPU box (no Julia, no libspasm
sources): the CPU arm is the oracle restatement in oracle/, on a bounded sample of the same
generator (the full 200k case needs ~1e15 scalar modular operations on a CPU); its line carries `value: null` and the
like-for-like numbers in `cpu_baseline` (oracle seconds next to the CUDA library's seconds on the SAME sample).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

PRIME = 42013
SM_COUNT = 148
FULL_N = 200_000
KNOWN_RANK_FULL = 199_993
NNZ_ROW = 10
SEED = 0x5A5A0002
CPU_SAMPLE_N = 20000       # cpu_baseline leg of the product arm: ~20 s of oracle work on 8 cores (blocked OpenMP dense tail)
REF_BUDGET_S = 150.0       # --impl reference: total CPU seconds the K + W sample steps may take
REF_SAMPLE_MAX = 24000     # largest sample the oracle finishes in <= 60 s on 8 cores
TIMING_NAMES = "total upload FL FLcol greedy reorder_extract density schur tail download rounds flcol_rounds greedy_windows schur_bytes schur_macs schur_ms".split()


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_input(n):
    import synth

    return synth.random_rows(n, n, NNZ_ROW, PRIME, SEED)


def csr_bytes(n, nnz):
    return 8 * (n + 1) + 8 * nnz


def gpu_same_sample(pkg, entry, n, p, j, x, expect_rank):
    """the SAME sample through the CUDA library, when a device is present: (resident s, end-to-end s) or (None, None)"""
    try:
        import torch

        if not torch.cuda.is_available():
            return None, None
        gpu = pkg.SpaSM()
        gpu.log(False)
        lib = gpu.lib
        lib.spasm_b200_upload.restype = C.c_void_p
        lib.spasm_b200_upload.argtypes = [C.c_void_p]
        lib.spasm_b200_echelonize_resident.restype = C.c_int
        lib.spasm_b200_echelonize_resident.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        lib.spasm_b200_release.argtypes = [C.c_void_p]
        # fresh caches: the pinned / device blocks of the full-size runs (tens of GB) must not shape the sample's timing
        lib.spasm_b200_trim()
        Ag = gpu.from_arrays(n, n, p, j, x, PRIME)
        opts = gpu.EchelonizeOpts()
        h = lib.spasm_b200_upload(C.cast(Ag.data, C.c_void_p))
        res, e2e = [], []
        for it in range(3):
            ms = C.c_double(0)
            r = lib.spasm_b200_echelonize_resident(h, C.byref(opts), C.byref(ms))
            assert r == expect_rank, (r, expect_rank)
            if it:
                res.append(ms.value / 1e3)
        lib.spasm_b200_release(h)
        for it in range(4):
            t = time.perf_counter()
            f = gpu.echelonize(Ag)
            rk = f.r
            dt = time.perf_counter() - t
            assert rk == expect_rank
            del f
            if it:
                e2e.append(dt)
        return sorted(res)[len(res) // 2], sorted(e2e)[len(e2e) // 2]  # medians (first call of each kind discarded)
    except Exception as exc:  # the reference arm must not die because of the product
        print(f"bench.py: same-sample GPU leg skipped: {exc}", file=sys.stderr)
        return None, None


def run_reference(args):
    """CPU arm: the oracle restatement of libspasm (oracle/), all host threads, on a BOUNDED SAMPLE of the
    workload (same generator, fewer rows).  time-to-rank of a sample is not the time-to-rank of the 200k matrix,
    so this line carries NO top-level value (null): the comparable numbers are cpu_baseline.value (CPU seconds on the
    sample) next to cpu_baseline.gpu_same_sample_s / gpu_same_sample_e2e_s (the CUDA library on the same sample)."""
    import __graft_entry__ as entry

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package()
    ora = pkg.SpaSM(entry.build_oracle())
    ora.log(False)
    ora.lib.spasm_oracle_set_num_threads(0)  # every host core, whatever OMP_NUM_THREADS the launcher exported
    cores = int(ora.lib.spasm_get_num_threads())
    W = 0 if args.warmup is None else args.warmup
    K = args.steps or 1
    n = args.rows
    if not n:
        # calibrate: one 4000-row run, cubic growth, K + W steps inside the budget
        p4, j4, x4 = make_input(4000)
        A4 = ora.from_arrays(4000, 4000, p4, j4, x4, PRIME)
        t = time.perf_counter()
        ora.echelonize(A4)
        t4 = max(time.perf_counter() - t, 1e-3)
        per_step = REF_BUDGET_S / (K + W)
        n = int(4000 * (per_step / t4) ** (1.0 / 3.0)) // 1000 * 1000
        n = max(4000, min(REF_SAMPLE_MAX, n))
    p, j, x = make_input(n)
    A = ora.from_arrays(n, n, p, j, x, PRIME)
    for _ in range(W):
        ora.echelonize(A)
    times = []
    r = None
    for _ in range(K):
        t = time.perf_counter()
        f = ora.echelonize(A)
        times.append(time.perf_counter() - t)
        r = f.r
        del f
    val = sum(times) / len(times)
    g_res, g_e2e = gpu_same_sample(pkg, entry, n, p, j, x, r)
    line = {
        "impl": "reference",
        "metric": "echelonize_time_to_rank", "value": None, "unit": "s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * val, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "int32 residues mod p (fp64 delayed-reduction products)",
        "data": "synthetic",
        "value_note": "null on purpose: the CPU cannot echelonize the 200000-row matrix in bounded time, and the time of a smaller sample is not "
                      "this metric; compare cpu_baseline.value with cpu_baseline.gpu_same_sample_s / gpu_same_sample_e2e_s (same matrix, same box)",
        "config": {"workload": f"BOUNDED SAMPLE of BASELINE configs[1]: random sparse {n}x{n}, {NNZ_ROW} nnz/row, echelonize+rank mod {PRIME} "
                               f"(full workload: {FULL_N}x{FULL_N})", "sample_n": n, "rank": r},
        "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "port", "sample_n": n, "sample_rank": r,
                         "sample": f"oracle restatement of libspasm (blocked dense tail, OpenMP, {cores} threads), {n}x{n} instance of the same generator",
                         "gpu_same_sample_s": g_res, "gpu_same_sample_e2e_s": g_e2e,
                         "speedup_same_sample_e2e": (val / g_e2e) if g_e2e else None},
        "e2e": {"value": None, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def dist_init(lib, dist, rank, world):
    """ship rank 0's NCCL id to every rank (torch.distributed is only the plumbing) and create the
    library's own communicator: the dense panels are broadcast with NCCL inside spasm_echelonize"""
    import torch

    lib.spasm_b200_nccl_unique_id.argtypes = [C.c_void_p]
    lib.spasm_b200_dist_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        assert lib.spasm_b200_nccl_unique_id(buf) == 0
        ident = torch.tensor(list(buf), dtype=torch.uint8)
    ident = ident.cuda()
    dist.broadcast(ident, src=0)
    raw = bytes(ident.cpu().tolist())
    assert lib.spasm_b200_dist_init(rank, world, raw) == 0


def run_ours(args):
    import numpy as np
    import torch

    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    gpu = pkg.SpaSM()
    if not os.environ.get("SPASM_B200_VERBOSE"):
        gpu.log(False)  # progress text off (errors still reach stderr)
    lib = gpu.lib
    lib.spasm_b200_upload.restype = C.c_void_p
    lib.spasm_b200_upload.argtypes = [C.c_void_p]
    lib.spasm_b200_echelonize_resident.restype = C.c_int
    lib.spasm_b200_echelonize_resident.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib.spasm_b200_release.argtypes = [C.c_void_p]
    lib.spasm_b200_last_timings.argtypes = [C.POINTER(C.c_double)]
    lib.spasm_b200_mma_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    lib.spasm_b200_mma_timing.argtypes = [C.c_int]
    lib.spasm_b200_set_cache.argtypes = [C.c_int]
    lib.spasm_b200_utcimma_peak.restype = C.c_double
    lib.spasm_b200_utcimma_peak.argtypes = [C.c_int, C.c_int]
    lib.spasm_b200_mma_timing(1)  # CUDA events around every tcgen05 launch (read once, after the timed region)
    lib.spasm_b200_set_cache(1)   # same-shaped calls in a loop: keep the device blocks between calls (opt-in; default is to return them)
    if world > 1:
        dist_init(lib, dist, rank, world)
        # the factor stays distributed: the owner of every dense panel materialises and downloads ITS rows of U
        # (a complete factor on rank 0 is the library's default; see csrc/dist.cuh)
        lib.spasm_b200_dist_shard_factor.argtypes = [C.c_int]
        lib.spasm_b200_dist_shard_factor(1)

    n = args.rows or FULL_N
    p, j, x = make_input(n)
    A = gpu.from_arrays(n, n, p, j, x, PRIME)
    nnz = int(p[-1])
    opts = gpu.EchelonizeOpts()
    handle = lib.spasm_b200_upload(C.cast(A.data, C.c_void_p))
    assert handle, "upload failed"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        flush.zero_()
        torch.cuda.synchronize()
        ms = C.c_double(0)
        r = lib.spasm_b200_echelonize_resident(handle, C.byref(opts), C.byref(ms))
        assert r >= 0, "echelonize failed (see stderr)"
        return r, ms.value

    W = 3 if args.warmup is None else args.warmup
    K = args.steps or 3
    for _ in range(W):
        resident_step()
    st = (C.c_double * 4)()
    lib.spasm_b200_mma_stats(st, 1)
    barrier()
    t_wall = time.perf_counter()
    dev_ms, ranks_seen = [], set()
    step_ms_all = dev_ms
    with ClockSampler(local) as clk:
        for _ in range(K):
            r, ms = resident_step()
            dev_ms.append(ms)
            ranks_seen.add(r)
        barrier()
    wall = time.perf_counter() - t_wall
    lib.spasm_b200_mma_stats(st, 0)
    mma_ms, mma_macs, mma_calls, launches = st[0], st[1], st[2], st[3]
    T = (C.c_double * 16)()
    lib.spasm_b200_last_timings(T)
    phases = {k: round(v, 6) for k, v in zip(TIMING_NAMES, T)}
    my = sum(dev_ms) / len(dev_ms) / 1e3
    if dist is not None:
        t = torch.tensor([my], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        my = float(t.item())
    assert len(ranks_seen) == 1
    rank_found = ranks_seen.pop()
    # the rank of the fixed-seed matrix is an invariant (found equal for A, A^T and a row permutation, tests/test_gpu_fullsize.py,
    # and on 1, 2 and 8 GPUs): a different value means a wrong elimination, not a different but valid answer
    if n == FULL_N:
        assert rank_found == KNOWN_RANK_FULL, f"rank {rank_found} != {KNOWN_RANK_FULL}"

    # ---- end to end through the C ABI on host structs (upload + download of the factor inside)
    e2e_steps = max(1, args.e2e_steps)
    lu = gpu.echelonize(A)  # warm-up (page cache of the host allocator, pinned bounce buffers)
    u_nnz = lu.U.nnz()  # this rank's share when the factor is distributed
    assert lu.r == rank_found
    del lu
    u_nnz_all = u_nnz
    if dist is not None:
        t = torch.tensor([u_nnz], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        u_nnz_all = int(t.item())
    barrier()
    e2e_t = []
    for _ in range(e2e_steps):
        flush.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        lu = gpu.echelonize(A)
        rk = lu.r  # the device->host read of the step's result
        e2e_t.append(time.perf_counter() - t)
        assert rk == rank_found
        del lu
    e2e = sum(e2e_t) / len(e2e_t)
    if dist is not None:
        t = torch.tensor([e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t.item())
    lib.spasm_b200_release(handle)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    hbm_gbs, bf16_sust, src = measured_peaks()
    # dominant kernel of this workload: k_gemm_i8limb (tcgen05.mma.kind::i8 -> SASS UTCIMMA), timed with
    # CUDA events on the launching stream inside the library.  4 int8 MACs per modular MAC, 2 ops per MAC.
    achieved = (8.0 * mma_macs / (mma_ms * 1e-3) / 1e12) if mma_ms > 0 else None
    # denominator: back-to-back tcgen05.mma.kind::i8 M128 N256 K32 from shared memory on all SMs, measured on THIS box now
    # (SURVEY.md 8d "builder must measure"); MEASURED_PEAKS.json has no int8 figure
    peak = lib.spasm_b200_utcimma_peak(4096, 5)
    peak_source = "measured in this run: back-to-back UTCIMMA M128xN256xK32, operands in shared memory, 148 CTAs (spasm_b200_utcimma_peak)"
    if not peak or peak <= 0:
        peak, peak_source = 2.0 * bf16_sust, f"fallback: 2 x bf16_tflops_sustained of {src}"
    # the same kernel ALONE on the GPU at a mid-elimination shape of the deferred updates (live columns x far rows x depth)
    alone = None
    if world == 1:
        lib.spasm_b200_gemm_probe.restype = C.c_double
        lib.spasm_b200_gemm_probe.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]
        pm, pn, pk = 57344, 49152, 4032
        t_alone = lib.spasm_b200_gemm_probe(PRIME, pm, pn, pk, 3)
        if t_alone and t_alone > 0:
            alone = {"shape_MxNxK": [pm, pn, pk], "kernel_ms": t_alone, "achieved": 8.0 * pm * pn * pk / (t_alone * 1e-3) / 1e12}
            alone["frac"] = alone["achieved"] / peak
    roofline = {
        "kernel": "k_gemm_i8limb (tcgen05.mma.kind::i8, SASS UTCIMMA; TMA-fed, TMEM accumulators)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": (achieved / peak) if achieved else None,
        "in_step_note": ("launch durations inside the step, CUDA events on the launching stream; one stream on a single GPU, so a launch has the "
                         "GPU to itself; `alone` is the same kernel at one mid-elimination shape of the deferred updates" if world == 1 else
                         "launch durations inside the step, CUDA events on the launching stream: with several ranks the deep updates run on a second "
                         "stream restricted to 100 of the 148 SMs and overlap the panel factorisation / broadcast chain of the main stream, so a "
                         "launch takes longer than alone although the step is shorter"),
        "alone": alone,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the ncu --set full capture
        # profiles/r02_ncu_gemm_i8limb_k4096.csv (M=32768 N=16384 K=4096, the depth of the deferred trailing updates)
        "traffic": 4.938201e9 + 2.137045e9,
        "traffic_note": "ncu capture of the kernel alone at M=32768 N=16384 K=4096 (tools/gemm_probe.py): 7.08e9 B of DRAM traffic for "
                        "4.70e9 B algorithmic (4.29e9 B of C read + written, 0.40e9 B of u8 limb planes): x1.5; 6.57 ms, tensor pipe 55 % active; "
                        "the bench's launches have other M and N",
        "peak_source": peak_source,
        "proxy_2x_bf16_sustained": 2.0 * bf16_sust,
        "algorithmic": "8 int8 ops per modular multiply-add (4 limb MMAs x 2), M*N*K per launch (unpadded, live columns only)",
        "kernel_ms_per_step": mma_ms / K, "launches_per_step": mma_calls / K,
        "share_of_step": (mma_ms / K / 1e3) / my if my > 0 else None,
    }

    # ---- secondary figures of BASELINE.json's metric (1 GPU only; outside the timed region of `value`):
    # the dense tail alone on configs[3], and the sparse Schur complement on a GL7d19-shaped instance
    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = {}
        f = lib.spasm_b200_dense_tail_bench
        f.restype = C.c_int
        f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_double)]
        nd, best = 32768, None
        for _ in range(3):
            ms = C.c_double(0)
            lib.spasm_b200_mma_stats(st, 1)
            rd = f(65521, nd, nd, 1000, 0x5A5A0004, C.byref(ms))
            lib.spasm_b200_mma_stats(st, 0)
            assert rd == nd
            if best is None or ms.value < best[0]:
                best = (ms.value, st[0], st[1])
        secondary["dense_tail_configs3"] = {
            "workload": "dense 32768x32768 mod 65521 into the dense-tail entry point (block 1000), generated on the device",
            "seconds": best[0] / 1e3, "rank": nd, "modp_gops": 2 * nd ** 3 / 3 / (best[0] * 1e-3) / 1e9,
            "modp_gops_definition": "2*(n^3/3)/t (SURVEY 8d); the elimination itself performs about n^3 modular MACs (reduced form)",
            "tcgen05_kernel_ms": best[1], "tcgen05_int8_tops": 8 * best[2] / max(best[1], 1e-9) / 1e9}
        import synth

        sc = 64
        ns, ms_, rs = 1911130 // sc, 1955309 // sc, 1033568 // sc
        pjx = synth.banded_planted(ns, ms_, rs, 12.0, 40, 42013, 0x5A5A0003, spread=16, colblock=8)
        As = gpu.from_arrays(ns, ms_, *pjx, 42013)
        gpu.echelonize(As)
        t = time.perf_counter()
        fs = gpu.echelonize(As)
        ts = time.perf_counter() - t
        assert fs.r == rs
        Ls = (C.c_longlong * 7)()
        lib.spasm_b200_last_stats.argtypes = [C.POINTER(C.c_longlong)]
        lib.spasm_b200_last_stats(Ls)
        lib.spasm_b200_last_timings(T)
        ph = dict(zip(TIMING_NAMES, T))
        secondary["sparse_regime_configs2_scaled"] = {
            "workload": f"GL7d19-shaped banded planted-rank generator at 1/{sc} scale ({ns}x{ms_}, rank {rs} by construction), "
                        "rounds of structural pivots + sparse Schur complement + GPLU tail",
            "seconds_e2e": ts, "rank": fs.r,
            "last_schur_algorithmic_gbps": (Ls[0] / (Ls[6] * 1e-6) / 1e9) if Ls[6] > 0 else None,
            "last_schur_gmacs_per_s": (Ls[1] / (Ls[6] * 1e-6) / 1e9) if Ls[6] > 0 else None,
            "schur_bytes_definition": "SURVEY 8d B_schur: 8 B per CSR entry read or written once per use + 8 B per row touched",
            "hbm_peak_gbps": hbm_gbs,
            "phases_s": {k: round(v, 5) for k, v in ph.items() if v}}
        del fs, As
        # the sparse row engine alone (SPASM_B200_SCHUR_DENSE=0 is read once per process: measured by tools/sparse_probe.py, see profiles/)

        # ---- BASELINE configs[4] scaled 1/100 (the full 500k x 1M kernel basis has ~2e11 non-zeros = 1.7 TB): rectangular
        # 5000 x 10000, 8 nnz/row: echelonize(L=true), kernel basis, 16 solvable + 4 random right-hand sides in ONE gesv
        import numpy as np
        import scipy.sparse as sp

        n4, m4 = 5000, 10000
        p4, j4, x4 = synth.random_rows(n4, m4, 8, PRIME, 0x5A5A0005)
        A4 = gpu.from_arrays(n4, m4, p4, j4, x4, PRIME)
        gpu.echelonize(A4, L=True)
        t = time.perf_counter()
        f4 = gpu.echelonize(A4, L=True)
        t_ech = time.perf_counter() - t
        t = time.perf_counter()
        K4 = gpu.kernel(f4)
        t_ker = time.perf_counter() - t
        lib.spasm_b200_last_stats(Ls)
        rng = np.random.default_rng(5)
        Ad = sp.csr_matrix((x4.astype(np.int64) % PRIME, j4, p4), shape=(n4, m4))
        X0 = sp.csr_matrix(rng.integers(0, PRIME, size=(16, n4)))
        Bs = sp.vstack([sp.csr_matrix((X0 @ Ad).toarray() % PRIME), sp.csr_matrix(rng.integers(0, PRIME, size=(4, m4)))]).tocsr()
        B4 = gpu.from_arrays(20, m4, Bs.indptr.astype(np.int64), Bs.indices.astype(np.int32), synth.balanced(Bs.data, PRIME), PRIME)
        gpu.gesv(f4, B4)
        t = time.perf_counter()
        X4, ok4 = gpu.gesv(f4, B4)
        t_sol = time.perf_counter() - t
        assert list(ok4[:16]) == [True] * 16, "a solvable system was rejected"
        Xd = sp.csr_matrix((X4.arrays()[2].astype(np.int64) % PRIME, X4.arrays()[1], X4.arrays()[0]), shape=(20, n4))
        assert not (((Xd[:16] @ Ad).toarray() - Bs[:16].toarray()) % PRIME).any(), "x.A != b"
        secondary["configs4_scaled"] = {
            "workload": f"random {n4}x{m4}, 8 nnz/row mod {PRIME} (BASELINE configs[4] at 1/100): echelonize(L=true) + kernel basis + gesv of 20 right-hand sides "
                        "(16 in the row space, checked x.A == b; 4 random)",
            "rank": f4.r, "echelonize_L_s": t_ech, "kernel_s": t_ker, "kernel_nnz": int(K4.nnz()), "gesv_20_rhs_s": t_sol,
            "solvable": int(sum(bool(v) for v in ok4)),
            "kernel_solve_algorithmic_gbps": (Ls[0] / (Ls[6] * 1e-6) / 1e9) if Ls[6] > 0 else None,
            "kernel_out_gbps_e2e": 8.0 * K4.nnz() / t_ker / 1e9}
        del f4, K4, A4, X4, B4

    # ---- CPU baseline: the oracle on a bounded sample, rank 0 only, with the SAME sample through the CUDA library
    cpu = None
    if not args.no_cpu and world == 1:
        ora = pkg.SpaSM(entry.build_oracle())
        ora.log(False)
        ora.lib.spasm_oracle_set_num_threads(0)
        ns = CPU_SAMPLE_N if n >= CPU_SAMPLE_N else n
        ps, js, xs = make_input(ns)
        As = ora.from_arrays(ns, ns, ps, js, xs, PRIME)
        t = time.perf_counter()
        fo = ora.echelonize(As)
        tc = time.perf_counter() - t
        cores = int(ora.lib.spasm_get_num_threads())
        g_res, g_e2e = gpu_same_sample(pkg, entry, ns, ps, js, xs, fo.r)
        cpu = {"value": tc, "unit": "s", "cores": cores, "kind": "port", "sample_n": ns, "sample_rank": fo.r,
               "sample": f"oracle restatement of libspasm (oracle/: blocked dense tail, OpenMP, {cores} threads), {ns}x{ns} instance of the same generator",
               "gpu_same_sample_s": g_res, "gpu_same_sample_e2e_s": g_e2e,
               "speedup_same_sample_e2e": (tc / g_e2e) if g_e2e else None}

    clocks = clk.summary()
    if achieved and clocks.get("sm_max_mhz"):
        # ncu's sm__ops_path_tensor_op_utcimma_src_int8 peak_sustained: 16384 int8 ops per cycle per SM (profiles/r01_ncu_gemm_i8limb_k4096_*.csv)
        hw = 16384.0 * SM_COUNT * clocks["sm_max_mhz"] * 1e6 / 1e12
        roofline["hw_int8_peak_tops_at_max_clock"] = hw
        roofline["frac_of_hw_int8_peak"] = achieved / hw
    lib.spasm_b200_set_cache(0)
    line = {
        "metric": "echelonize_time_to_rank", "value": my, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * my, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 residues mod p; dense tail on u8 limbs with int32 tensor-core accumulation", "data": "synthetic",
        "config": {"workload": f"random sparse {n}x{n}, {NNZ_ROW} nnz/row, echelonize+rank mod {PRIME} (BASELINE configs[1])",
                   "rank": rank_found, "nnz": nnz, "U_nnz": int(u_nnz_all), "l2": "flushed between iterations (256 MiB write)",
                   "parallelism": "single GPU" if world == 1 else
                   f"structural pivots replicated; dense tail sharded block-cyclically over {world} ranks, one NCCL broadcast per panel; "
                   f"every rank materialises and downloads the rows of U of its own panels (spasm_b200_dist_shard_factor)"},
        "e2e": {"value": e2e, "unit": "s", "steps": e2e_steps, "h2d_bytes_per_step": csr_bytes(n, nnz) * world,
                "d2h_bytes_per_step": int(world * (8 * (rank_found + 1) + 4 * n) + 8 * u_nnz_all)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "secondary": secondary,
        "step_s": [round(v / 1e3, 4) for v in step_ms_all],
        "phases_last_step_s": phases,
        "wall_s_timed_region": wall,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=None, help="debug: override the matrix size (the bench value is only valid at the default)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[3] dense tail / sparse Schur side figures")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
