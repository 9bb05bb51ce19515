"""Host logic of the dense tail's deferred updates (csrc/dense.cu: plan_tail, exported as spasm_b200_tail_plan): for 1, 2, 4 and 8
ranks the accumulators must hold every panel that arrives between two hand-overs (no overflow, ever), and — for the primes the
bench runs on — no flush may be needed before a rank's near rows are used up (such a flush stalls the look-ahead: the main stream
has to wait for the whole second stream).  The loop below replays the bookkeeping of dense_tail_core for full-rank panels."""
import ctypes as C
import os

import pytest

import __graft_entry__ as entry


def tail_plan(lib, Sm0, n_local, bs, NR, kcap, max_k):
    out = (C.c_longlong * 4)()
    lib.spasm_b200_tail_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p]
    lib.spasm_b200_tail_plan(Sm0, n_local, bs, NR, kcap, max_k, 1 << 60, out)
    return bool(out[0]), int(out[1]), int(out[2]), int(out[3])


def replay(lib, nrows, bs, NR, me, kcap, max_k):
    nb = (nrows + bs - 1) // bs
    n_local = sum(min(bs, nrows - b * bs) for b in range(me, nb, NR))
    lazy, group, kdepth, LDK = tail_plan(lib, nrows, n_local, bs, NR, kcap, max_k)
    B16 = (min(bs, nrows) + 15) // 16 * 16
    lb, gend, Kacc = 0, group * bs, 0
    early, handovers, deepest = 0, 0, 0
    for b in range(nb):
        owner = b % NR
        if lazy and owner == me and lb * bs >= gend:  # my near rows are used up: hand-over
            if Kacc > 0 and min(gend, n_local) < n_local:
                handovers += 1
                deepest = max(deepest, Kacc)
            Kacc = 0
            gend = lb * bs + group * bs
        if owner == me:
            if lazy:
                assert lb * bs < gend, "a panel would be factored before its rows were brought up to date"
            lb += 1
        kb = lb * bs
        if lazy and n_local - kb > 0:
            fe = min(max(gend, kb), n_local)
            assert Kacc + B16 <= LDK, "the accumulators overflow"
            Kacc += B16
            assert Kacc <= max_k, "deeper than one tensor-core launch takes"
            if Kacc > kdepth or fe >= n_local:
                if Kacc > kdepth and fe < n_local:
                    early += 1
                    deepest = max(deepest, Kacc)
                Kacc = 0
    return lazy, group, kdepth, early, handovers, deepest


@pytest.mark.parametrize("NR", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(112_000, 1000), (32_768, 1000), (9_000, 500), (5_200, 300), (60_000, 2000)])
def test_no_early_flush_for_two_limb_primes(NR, shape):
    lib = C.CDLL(str(entry.build_product()), mode=os.RTLD_LOCAL)
    nrows, bs = shape
    for me in range(NR):
        lazy, group, kdepth, early, handovers, deepest = replay(lib, nrows, bs, NR, me, 4096, 16384)
        if not lazy:
            continue
        B16 = (min(bs, nrows) + 15) // 16 * 16
        if (group * NR + NR - 1) * B16 > 16384 - B16:
            continue  # (e.g. blocks of 2000 rows on 8 ranks: one interval does not fit one launch; early flushes, still correct)
        assert early == 0, f"rank {me}/{NR}: {early} flushes before the hand-over (group {group}, depth {kdepth})"
        assert handovers >= 1 and deepest <= kdepth
        if NR * bs <= 4096:
            assert group == 4096 // (bs * NR)


@pytest.mark.parametrize("max_k", [10880, 8192])
def test_deep_primes_never_overflow(max_k):
    """3- and 4-limb primes take shallower products: with 8 ranks a flush may come early (slower, still correct) but the
    buffers never overflow and no launch is deeper than the kernel takes — the asserts inside replay()"""
    lib = C.CDLL(str(entry.build_product()), mode=os.RTLD_LOCAL)
    for NR in (1, 2, 4, 8):
        for me in range(NR):
            replay(lib, 112_000, 1000, NR, me, 4096, max_k)


def test_small_tails_stay_eager():
    lib = C.CDLL(str(entry.build_product()), mode=os.RTLD_LOCAL)
    assert tail_plan(lib, 3000, 1800, 1000, 1, 4096, 16384)[0] is False  # not more than two panels of rows: eager updates
    assert tail_plan(lib, 3000, 3000, 1000, 1, 0, 16384)[0] is False     # SPASM_B200_LAZY_K=0
    lazy, group, kdepth, LDK = tail_plan(lib, 112_000, 112_000, 1000, 1, 4096, 16384)
    assert (lazy, group, kdepth, LDK) == (True, 4, 4096, 4096 + 1008)
