"""Host-logic model of dense_tail_core (spasm.jl_b200/csrc/dense.cu): the block-cyclic sharding of the dense
Schur complement over N ranks and the DEFERRED trailing updates (near / far rows, pending factors Rt_acc / Pt_acc,
correction of the multipliers, the two flush triggers), line by line in numpy, against the eager elimination.
It checks the orchestration (indices, flush points, ownership) for rank counts and depths the GPU budget of the
round did not cover; the arithmetic itself is the library's and is tested on the GPU.  tests/test_sim_lazy.py runs it."""
import numpy as np


def rref_panel(rows, cand, p):
    """Gauss-Jordan on `rows` (Sn x Sm0) scanning candidate columns left to right; returns (T-applied reduced pivot rows
    over ALL columns, pivot columns) — any deterministic rule does, eager and lazy use the same one."""
    W = rows.copy() % p
    Sn = W.shape[0]
    used = np.zeros(Sn, dtype=bool)
    piv_rows, piv_cols = [], []
    for c in cand:
        r = next((i for i in range(Sn) if not used[i] and W[i, c] % p), None)
        if r is None:
            continue
        inv = pow(int(W[r, c]), p - 2, p)
        W[r] = W[r] * inv % p
        for i in range(Sn):
            if i != r and W[i, c]:
                W[i] = (W[i] - W[i, c] * W[r]) % p
        used[r] = True
        piv_rows.append(r)
        piv_cols.append(c)
    order = np.argsort(piv_cols)
    return W[[piv_rows[i] for i in order]] if piv_rows else np.zeros((0, rows.shape[1]), dtype=np.int64), [piv_cols[i] for i in order]


def local_positions(nrows, bs, NR, me):
    pos = []
    for b in range((nrows + bs - 1) // bs):
        if b % NR == me:
            pos += list(range(b * bs, min(nrows, (b + 1) * bs)))
    return pos


def run(D, p, bs, NR, kcap, lazy_enabled=True):
    nrows, Sm0 = D.shape
    loc = [D[local_positions(nrows, bs, NR, r)].copy() % p for r in range(NR)]
    n_local = [x.shape[0] for x in loc]
    colpiv = np.zeros(Sm0, dtype=bool)
    Bmax = min(bs, nrows)
    B16 = (Bmax + 15) // 16 * 16
    kc = max(0, min(kcap, 16384 - B16))
    st = []
    for r in range(NR):
        lazy = lazy_enabled and kc >= 2 * B16 and n_local[r] > 2 * bs
        group = max(1, kc // max(bs, 1)) if lazy else 1
        st.append(dict(lazy=lazy, group=group, Kacc=0, gend=group * bs, lb=0, Rt=[], Pt={}, ncorr=0, nflush=0))
    out = []
    nb = (nrows + bs - 1) // bs

    def flush_far(r):
        s = st[r]
        fe = min(s["gend"], n_local[r])
        if s["Kacc"] > 0 and fe < n_local[r]:
            for (Rb, pc, Pb) in s["Rt"]:  # one product of depth Kacc in the library
                loc[r][fe:] = (loc[r][fe:] - Pb[fe:] @ Rb) % p
            s["nflush"] += 1
        s["Kacc"] = 0
        s["Rt"] = []

    for b in range(nb):
        owner = b % NR
        Sn = min(bs, nrows - b * bs)
        so = st[owner]
        if so["lazy"] and so["lb"] * bs >= so["gend"]:
            flush_far(owner)
            so["gend"] = so["lb"] * bs + so["group"] * bs
        k0 = so["lb"] * bs
        cand = [c for c in range(Sm0) if not colpiv[c]]
        R, pc = rref_panel(loc[owner][k0:k0 + Sn], cand, p)
        so["lb"] += 1
        rr = len(pc)
        out.append((pc, R.copy()))
        if rr == 0:
            continue
        colpiv[pc] = True
        for r in range(NR):
            s = st[r]
            kb = s["lb"] * bs
            nk = max(0, n_local[r] - kb)
            if nk == 0:
                continue
            if not s["lazy"]:
                P = loc[r][kb:, pc].copy()
                loc[r][kb:] = (loc[r][kb:] - P @ R) % p
                continue
            rr16 = (rr + 15) // 16 * 16
            fe = min(max(s["gend"], kb), n_local[r])
            Pb = np.zeros((n_local[r], rr), dtype=np.int64)
            Pb[kb:] = loc[r][kb:, pc]
            if s["Kacc"] > 0 and fe < n_local[r]:
                for (Rq, pq, Pq) in s["Rt"]:  # Pt_b -= Pt_acc . Rt_acc[pc_b]^T on the far rows
                    Pb[fe:] = (Pb[fe:] - Pq[fe:] @ Rq[:, pc]) % p
                s["ncorr"] += 1
            if fe > kb:
                loc[r][kb:fe] = (loc[r][kb:fe] - Pb[kb:fe] @ R) % p
            s["Rt"].append((R.copy(), list(pc), Pb))
            s["Kacc"] += rr16
            if s["Kacc"] + B16 > kc or fe >= n_local[r]:
                flush_far(r)
    return out, st
