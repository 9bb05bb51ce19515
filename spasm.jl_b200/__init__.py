"""
spasm.jl_b200 — host-side mirror of SpaSM.jl's API over the C ABI in include/spasm_b200.h.

The reference host language is Julia (not installed here, nor on the GPU box), so this module
replays SpaSM.jl's exact call sequences through `ctypes`; names, argument meaning and error
behaviour follow /root/reference/src/SpaSM.jl:

    Field, ZZp              src/SpaSM.jl:51-121, :383-406
    CSR / csr_alloc         src/SpaSM.jl:144-151, :441, :941-968 (column of the input = SpaSM row)
    sparse                  src/SpaSM.jl:1011-1023 (sorts every row by column)
    transpose               src/SpaSM.jl:589
    EchelonizeOpts          src/SpaSM.jl:325-343, :817-824
    echelonize, LU, rank    src/SpaSM.jl:860-866, :271-305, :1149
    kernel, rref            src/SpaSM.jl:876-882, :871, :1147
    solve, gesv             src/SpaSM.jl:895-923
    sparse_triangular_solve src/SpaSM.jl:714-755
    xapy, axpy, scatter     src/SpaSM.jl:620-658
    log                     src/SpaSM.jl:34-46

The product library is spasm.jl_b200/libspasm_b200.so (CUDA, sm_100a).  There is NO CPU
fallback: `SpaSM()` raises if that library is missing.  tests/ may instantiate
`SpaSM(path_to_oracle)` to drive the CPU oracle through the very same harness.

The Julia source a maintainer would ship is in spasm.jl_b200/julia/ (unexecuted here).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
PRODUCT_LIB = PKG_DIR / "libspasm_b200.so"
PRIME0 = 42013  # src/SpaSM.jl:16


# --------------------------------------------------------------------------- C mirrors
class _Field(C.Structure):  # src/SpaSM.jl:51-56
    _fields_ = [("p", C.c_int64), ("halfp", C.c_int64), ("mhalfp", C.c_int64), ("dinvp", C.c_double)]


class _CSR(C.Structure):  # src/SpaSM.jl:126-134
    _fields_ = [
        ("nzmax", C.c_int64),
        ("n", C.c_int32),
        ("m", C.c_int32),
        ("p", C.POINTER(C.c_int64)),
        ("j", C.POINTER(C.c_int32)),
        ("x", C.POINTER(C.c_int32)),
        ("field", _Field),
    ]


class _Triplet(C.Structure):  # src/SpaSM.jl:234-243
    _fields_ = [
        ("nzmax", C.c_int64),
        ("nz", C.c_int64),
        ("n", C.c_int32),
        ("m", C.c_int32),
        ("i", C.POINTER(C.c_int32)),
        ("j", C.POINTER(C.c_int32)),
        ("x", C.POINTER(C.c_int32)),
        ("field", _Field),
    ]


class _LU(C.Structure):  # src/SpaSM.jl:262-270
    _fields_ = [
        ("r", C.c_int32),
        ("complete", C.c_uint8),
        ("partial", C.c_uint8),  # spasm_b200: padding byte of the reference struct (include/spasm_b200.h)
        ("L", C.POINTER(_CSR)),
        ("U", C.POINTER(_CSR)),
        ("qinv", C.POINTER(C.c_int32)),
        ("p", C.POINTER(C.c_int32)),
        ("Ltmp", C.c_void_p),
    ]


class EchelonizeOpts(C.Structure):  # src/SpaSM.jl:325-343
    _fields_ = [
        ("enable_greedy_pivot_search", C.c_bool),
        ("enable_tall_and_skinny", C.c_bool),
        ("enable_dense", C.c_bool),
        ("enable_GPLU", C.c_bool),
        ("L", C.c_bool),
        ("complete", C.c_bool),
        ("min_pivot_proportion", C.c_double),
        ("max_round", C.c_int32),
        ("sparsity_threshold", C.c_double),
        ("dense_block_size", C.c_int64),  # Julia declares Int (src/SpaSM.jl:339)
        ("low_rank_ratio", C.c_double),
        ("tall_and_skinny_ratio", C.c_double),
        ("low_rank_start_weight", C.c_double),
    ]


assert C.sizeof(_Field) == 32 and C.sizeof(_CSR) == 72 and C.sizeof(_Triplet) == 80
assert C.sizeof(_LU) == 48 and C.sizeof(EchelonizeOpts) == 64

LOGFUNC = C.CFUNCTYPE(C.c_int, C.c_char_p)

# every symbol include/spasm_b200.h declares: (restype, argtypes)
_P = C.POINTER
ABI = {
    "spasm_field_init": (None, [C.c_int64, _P(_Field)]),
    "spasm_ZZp_init": (C.c_int32, [_P(_Field), C.c_int64]),
    "spasm_ZZp_add": (C.c_int32, [_P(_Field), C.c_int32, C.c_int32]),
    "spasm_ZZp_sub": (C.c_int32, [_P(_Field), C.c_int32, C.c_int32]),
    "spasm_ZZp_mul": (C.c_int32, [_P(_Field), C.c_int32, C.c_int32]),
    "spasm_ZZp_inverse": (C.c_int32, [_P(_Field), C.c_int32]),
    "spasm_ZZp_axpy": (C.c_int32, [_P(_Field), C.c_int32, C.c_int32, C.c_int32]),
    "spasm_wtime": (C.c_double, []),
    "spasm_nnz": (C.c_int64, [_P(_CSR)]),
    "spasm_malloc": (C.c_void_p, [C.c_int64]),
    "spasm_calloc": (C.c_void_p, [C.c_int64, C.c_int64]),
    "spasm_realloc": (C.c_void_p, [C.c_void_p, C.c_int64]),
    "spasm_csr_alloc": (_P(_CSR), [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_bool]),
    "spasm_csr_realloc": (None, [_P(_CSR), C.c_int64]),
    "spasm_csr_resize": (None, [_P(_CSR), C.c_int32, C.c_int32]),
    "spasm_csr_free": (None, [_P(_CSR)]),
    "spasm_triplet_alloc": (_P(_Triplet), [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_bool]),
    "spasm_triplet_realloc": (None, [_P(_Triplet), C.c_int64]),
    "spasm_triplet_free": (None, [_P(_Triplet)]),
    "spasm_lu_free": (None, [_P(_LU)]),
    "spasm_get_num_threads": (C.c_int32, []),
    "spasm_get_thread_num": (C.c_int32, []),
    "spasm_add_entry": (None, [_P(_Triplet), C.c_int32, C.c_int32, C.c_int64]),
    "spasm_triplet_transpose": (None, [_P(_Triplet)]),
    "spasm_compress": (_P(_CSR), [_P(_Triplet)]),
    "spasm_triplet_load": (_P(_Triplet), [C.c_void_p, C.c_int64, C.c_void_p]),
    "spasm_triplet_save": (None, [_P(_Triplet), C.c_void_p]),
    "spasm_csr_save": (None, [_P(_CSR), C.c_void_p]),
    "spasm_transpose": (_P(_CSR), [_P(_CSR)]),
    "spasm_scatter": (None, [_P(_CSR), C.c_int32, C.c_int32, _P(C.c_int32)]),
    "spasm_xApy": (None, [_P(C.c_int32), _P(_CSR), _P(C.c_int32)]),
    "spasm_Axpy": (None, [_P(_CSR), _P(C.c_int32), _P(C.c_int32)]),
    "spasm_sparse_triangular_solve": (C.c_int32, [_P(_CSR), _P(_CSR), C.c_int32, _P(C.c_int32), _P(C.c_int32), _P(C.c_int32)]),
    "spasm_dense_back_solve": (C.c_bool, [_P(_CSR), _P(C.c_int32), _P(C.c_int32), _P(C.c_int32)]),
    "spasm_dense_forward_solve": (C.c_bool, [_P(_CSR), _P(C.c_int32), _P(C.c_int32), _P(C.c_int32)]),
    "spasm_pivots_extract_structural": (C.c_int32, [_P(_CSR), _P(C.c_int32), _P(_LU), _P(C.c_int32), _P(EchelonizeOpts)]),
    "spasm_schur_estimate_density": (C.c_double, [_P(_CSR), _P(C.c_int32), C.c_int32, _P(_CSR), _P(C.c_int32), C.c_int32]),
    "spasm_schur": (_P(_CSR), [_P(_CSR), _P(C.c_int32), C.c_int32, _P(_LU), C.c_double, C.c_void_p, _P(C.c_int32), _P(C.c_int32)]),
    "spasm_echelonize_init_opts": (None, [_P(EchelonizeOpts)]),
    "spasm_echelonize": (_P(_LU), [_P(_CSR), _P(EchelonizeOpts)]),
    "spasm_rref": (_P(_CSR), [_P(_LU), _P(C.c_int32)]),
    "spasm_kernel": (_P(_CSR), [_P(_LU)]),
    "spasm_solve": (C.c_bool, [_P(_LU), _P(C.c_int32), _P(C.c_int32)]),
    "spasm_gesv": (_P(_CSR), [_P(_LU), _P(_CSR), _P(C.c_bool)]),
    "spasm_dense_rref": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32, _P(C.c_int32), C.c_int64, _P(C.c_int32)]),
    "spasm_b200_backend": (C.c_char_p, []),
    "spasm_b200_seed": (None, [C.c_uint64]),
}
DATA_SYMBOLS = ["logcallback"]


def _i32ptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P(C.c_int32))


# --------------------------------------------------------------------------- field arithmetic
class Field:
    """Field(p): the finite field F_p (src/SpaSM.jl:73-76)."""

    def __init__(self, p: int):
        assert 2 < p <= 0xFFFFFFFB  # src/SpaSM.jl:74
        self.p, self.halfp, self.mhalfp, self.dinvp = p, p // 2, p // 2 - p + 1, 1.0 / p

    def __eq__(self, o):
        return isinstance(o, Field) and o.p == self.p

    def __hash__(self):
        return hash(("Field", self.p))

    def __repr__(self):
        return f"F_{self.p}"

    def normalize(self, x: int) -> int:  # src/SpaSM.jl:83-88
        if x < self.mhalfp:
            x += self.p
        elif x > self.halfp:
            x -= self.p
        return int(x)

    def __call__(self, x: int) -> "ZZp":  # src/SpaSM.jl:99
        return ZZp(self, x)

    def balanced(self, a):
        """vectorised: any integer array -> balanced int32 residues"""
        a = np.mod(np.asarray(a, dtype=np.int64), self.p)
        return np.where(a > self.halfp, a - self.p, a).astype(np.int32)


class ZZp:
    """Element of F_p stored as the balanced Int32 representative (src/SpaSM.jl:79-121, :383-406)."""

    __slots__ = ("F", "v")

    def __init__(self, F: Field | int, x: int = 0):
        if not isinstance(F, Field):
            F = Field(int(F))
        self.F = F
        self.v = F.normalize(int(x) % F.p)

    def _c(self, o):
        if isinstance(o, ZZp):
            assert o.F == self.F
            return o.v
        return int(o)

    def __add__(self, o):
        return ZZp(self.F, self.v + self._c(o))

    __radd__ = __add__

    def __sub__(self, o):
        return ZZp(self.F, self.v - self._c(o))

    def __rsub__(self, o):
        return ZZp(self.F, self._c(o) - self.v)

    def __neg__(self):
        return ZZp(self.F, -self.v)

    def __mul__(self, o):
        return ZZp(self.F, self.v * self._c(o))

    __rmul__ = __mul__

    def inv(self):  # src/SpaSM.jl:386
        return ZZp(self.F, pow(self.v % self.F.p, -1, self.F.p))

    def __truediv__(self, o):
        return self * (o if isinstance(o, ZZp) else ZZp(self.F, o)).inv()

    def __eq__(self, o):
        return (isinstance(o, ZZp) and o.F == self.F and o.v == self.v) or (isinstance(o, int) and ZZp(self.F, o).v == self.v)

    def __hash__(self):
        return hash((self.v, self.F.p))

    def __int__(self):
        return self.v

    def __repr__(self):
        return str(self.v)


def axpy_zzp(a: ZZp, x: ZZp, y: ZZp) -> ZZp:  # src/SpaSM.jl:387-390
    return ZZp(a.F, a.v * x.v + y.v)


# --------------------------------------------------------------------------- handles
class CSR:
    """Owning / non-owning handle on a C `spasm_csr` (src/SpaSM.jl:144-167)."""

    def __init__(self, api: "SpaSM", ptr, own=True):
        if not ptr:
            raise RuntimeError("spasm: NULL CSR returned by the library")
        self._api, self.data, self._own = api, ptr, own

    def __del__(self):  # finalizer -> spasm_csr_free (src/SpaSM.jl:148, :451)
        try:
            if self._own and self.data:
                self._api.lib.spasm_csr_free(self.data)
                self.data = None
        except Exception:
            pass

    @property
    def n(self):
        return int(self.data.contents.n)

    @property
    def m(self):
        return int(self.data.contents.m)

    @property
    def nzmax(self):
        return int(self.data.contents.nzmax)

    @property
    def prime(self):
        return int(self.data.contents.field.p)

    @property
    def field(self):
        return Field(self.prime)

    @property
    def shape(self):
        return (self.n, self.m)

    # unsafe_wrap views (src/SpaSM.jl:158-163)
    @property
    def p(self):
        return np.ctypeslib.as_array(self.data.contents.p, shape=(self.n + 1,))

    @property
    def j(self):
        return np.ctypeslib.as_array(self.data.contents.j, shape=(max(self.nzmax, 1),))[: self.nzmax]

    @property
    def x(self):
        return np.ctypeslib.as_array(self.data.contents.x, shape=(max(self.nzmax, 1),))[: self.nzmax]

    def nnz(self):
        return int(self._api.lib.spasm_nnz(self.data))

    def arrays(self):
        """(p, j, x) copies trimmed to nnz — the bit-exact comparison key"""
        nz = self.nnz()
        return self.p.copy(), self.j[:nz].copy(), self.x[:nz].copy()

    def row(self, i):
        p = self.p
        return self.j[p[i] : p[i + 1]], self.x[p[i] : p[i + 1]]

    def __repr__(self):  # src/SpaSM.jl:195
        return f"{self.n}x{self.m} CSR matrix % {self.prime} with {self.nnz()} (maximum {self.nzmax}) non-zeros"

    def __eq__(self, o):  # src/SpaSM.jl:1005
        a, b = self._api.sparse(self), self._api.sparse(o)
        return a.shape == b.shape and (a != b).nnz == 0


class LU:
    """Handle on a C `spasm_lu` (src/SpaSM.jl:271-305)."""

    def __init__(self, api: "SpaSM", ptr):
        if not ptr:
            raise RuntimeError("spasm: NULL LU returned by the library (see log)")
        self._api, self.data = api, ptr

    def __del__(self):
        try:
            if self.data:
                self._api.lib.spasm_lu_free(self.data)
                self.data = None
        except Exception:
            pass

    @property
    def r(self):
        return int(self.data.contents.r)

    @property
    def partial(self):
        """multi-GPU runs: True when this process only holds some rows of U (include/spasm_b200.h)"""
        return bool(self.data.contents.partial)

    @property
    def complete(self):
        return bool(self.data.contents.complete)

    @property
    def U(self):
        if not self.data.contents.U:
            raise RuntimeError("M.U is null")  # src/SpaSM.jl:291
        u = CSR(self._api, self.data.contents.U, own=False)
        u._keep = self
        return u

    @property
    def L(self):
        if not self.data.contents.L:
            raise RuntimeError("M.L is null")  # src/SpaSM.jl:288
        l = CSR(self._api, self.data.contents.L, own=False)
        l._keep = self
        return l

    @property
    def qinv(self):
        if not self.data.contents.qinv:
            raise RuntimeError("M.qinv is null")
        return np.ctypeslib.as_array(self.data.contents.qinv, shape=(self.U.m,))

    @property
    def p(self):
        if not self.data.contents.p:
            raise RuntimeError("M.p is null")
        return np.ctypeslib.as_array(self.data.contents.p, shape=(self.r if self.r > 0 else 1,))[: self.r]


# --------------------------------------------------------------------------- the API object
class SpaSM:
    """One loaded library + the SpaSM.jl-shaped functions over it."""

    def __init__(self, libpath: str | os.PathLike | None = None):
        path = Path(libpath) if libpath is not None else PRODUCT_LIB
        if not path.exists():
            raise RuntimeError(
                f"spasm.jl_b200: {path} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " There is no CPU fallback for the product path."
            )
        self.path = path
        self.lib = C.CDLL(str(path), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        for name, (res, args) in ABI.items():
            fn = getattr(self.lib, name)  # AttributeError if the ABI is incomplete
            fn.restype, fn.argtypes = res, args
        self._logcb = None
        self.backend = self.lib.spasm_b200_backend().decode()

    # ---- src/SpaSM.jl:34-46
    def log(self, l=None):
        slot = C.c_void_p.in_dll(self.lib, "logcallback")
        if l is None:
            self._logcb, slot.value = None, None
            return
        if l is True:
            f = lambda s: (print("SPASM: " + s.decode(errors="replace"), end=""), 0)[1]
        elif l is False:
            f = lambda s: 0
        else:
            f = lambda s: int(l(s.decode(errors="replace")) or 0)
        self._logcb = LOGFUNC(f)
        slot.value = C.cast(self._logcb, C.c_void_p).value

    # ---- allocation (src/SpaSM.jl:441-451)
    def csr_alloc(self, n, m, nzmax, prime=PRIME0, with_values=True) -> CSR:
        return CSR(self, self.lib.spasm_csr_alloc(n, m, nzmax, prime, with_values))

    def spzeros(self, F: Field, n, m) -> CSR:
        M = self.csr_alloc(n, m, 0, F.p)
        M.p[:] = 0
        return M

    def nnz(self, A: CSR) -> int:
        return A.nnz()

    def wtime(self) -> float:
        return float(self.lib.spasm_wtime())

    # ---- constructors (src/SpaSM.jl:941-988): a COLUMN of the input becomes a SpaSM ROW
    def CSR(self, A, prime=PRIME0, transpose=True) -> CSR:
        import scipy.sparse as sp

        A = sp.csc_matrix(A)
        A.sum_duplicates()
        rows, cols = A.shape
        F = Field(prime)
        vals = F.balanced(A.data)
        keep = vals != 0
        M = self.csr_alloc(cols, rows, int(A.nnz), prime, True)
        colptr = A.indptr.astype(np.int64)
        if keep.all():
            M.p[:] = colptr
            nz = int(A.nnz)
            M.j[:nz] = A.indices.astype(np.int32)
            M.x[:nz] = vals
        else:
            kept = np.concatenate([[0], np.cumsum(keep)]).astype(np.int64)
            M.p[:] = kept[colptr]
            nz = int(kept[-1])
            M.j[:nz] = A.indices[keep].astype(np.int32)
            M.x[:nz] = vals[keep]
        return M if transpose else self.transpose(M)

    def from_arrays(self, n, m, p, j, x, prime=PRIME0) -> CSR:
        """Build a SpaSM CSR directly from row pointers / columns / values (rows = SpaSM rows)."""
        p = np.asarray(p, dtype=np.int64)
        nz = int(p[-1])
        M = self.csr_alloc(n, m, nz, prime, True)
        M.p[:] = p
        M.j[:nz] = np.asarray(j, dtype=np.int32)[:nz]
        M.x[:nz] = Field(prime).balanced(np.asarray(x)[:nz])
        return M

    # ---- src/SpaSM.jl:1011-1023: CSC matrix (m x n) whose columns are the SpaSM rows, sorted
    def sparse(self, A: CSR, transpose=True):
        import scipy.sparse as sp

        p, j, x = A.arrays()
        mat = sp.csc_matrix((x.astype(np.int64), j.astype(np.int64), p), shape=(A.m, A.n))
        mat.sort_indices()
        return mat if transpose else sp.csc_matrix(mat.T)

    def findnz(self, A: CSR):  # src/SpaSM.jl:1088-1102 (1-based like Julia)
        p, j, x = A.arrays()
        I = np.repeat(np.arange(1, A.n + 1), np.diff(p))
        return I, j.astype(np.int64) + 1, x.copy()

    # ---- src/SpaSM.jl:589
    # ---- SMS files: "n m M" / "i j v" (1-based) / "0 0 0"  (src/SpaSM.jl:498-529)
    @staticmethod
    def _fopen(path, mode):
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        f = libc.fopen(os.fsencode(str(path)), mode.encode())
        if not f:
            raise OSError(f"cannot open {path}")
        libc.fclose.argtypes = [C.c_void_p]
        return libc, f

    def compress(self, T, device=False) -> CSR:
        """compress(T::Triplet) (src/SpaSM.jl:479-493): triplets -> CSR, duplicates summed into the first occurrence, zero sums
        dropped.  device=True runs it on the GPU (spasm_b200_compress, csrc/compress.cu; CUDA library only) — same arrays."""
        if device:
            f = self.lib.spasm_b200_compress
            f.restype, f.argtypes = _P(_CSR), [_P(_Triplet)]
            ptr = f(T)
            if not ptr:
                raise RuntimeError("spasm_b200_compress failed (see stderr)")
            return CSR(self, ptr)
        return CSR(self, self.lib.spasm_compress(T))

    def load(self, path, prime=PRIME0) -> CSR:
        """fileio_load(...; csr = true) (src/SpaSM.jl:506-512): spasm_triplet_load + spasm_compress"""
        libc, f = self._fopen(path, "r")
        try:
            T = self.lib.spasm_triplet_load(f, prime, None)
        finally:
            libc.fclose(f)
        if not T:
            raise ValueError(f"{path}: not an SMS file")
        try:
            return CSR(self, self.lib.spasm_compress(T))
        finally:
            self.lib.spasm_triplet_free(T)

    def save(self, path, A: CSR):
        """fileio_save(f, A::CSR) (src/SpaSM.jl:523-529): spasm_csr_save"""
        libc, f = self._fopen(path, "w")
        try:
            self.lib.spasm_csr_save(A.data, f)
        finally:
            libc.fclose(f)

    def transpose(self, A: CSR) -> CSR:
        return CSR(self, self.lib.spasm_transpose(A.data))

    # ---- src/SpaSM.jl:620-658
    def scatter(self, A: CSR, i: int, beta: int, x: np.ndarray):
        self.lib.spasm_scatter(A.data, i, int(beta), _i32ptr(x))

    def xapy(self, x: np.ndarray, A: CSR, y: np.ndarray):
        assert len(x) == A.n and len(y) == A.m
        self.lib.spasm_xApy(_i32ptr(x), A.data, _i32ptr(y))
        return y

    def axpy(self, A: CSR, x: np.ndarray, y: np.ndarray):
        assert len(x) == A.m and len(y) == A.n
        self.lib.spasm_Axpy(A.data, _i32ptr(x), _i32ptr(y))
        return y

    def vecmat(self, x, A: CSR):  # x * A (src/SpaSM.jl:645)
        return self.xapy(np.ascontiguousarray(x, dtype=np.int32), A, np.zeros(A.m, dtype=np.int32))

    def matvec(self, A: CSR, x):  # A * x (src/SpaSM.jl:658)
        return self.axpy(A, np.ascontiguousarray(x, dtype=np.int32), np.zeros(A.n, dtype=np.int32))

    # ---- src/SpaSM.jl:714-755
    def sparse_triangular_solve_row(self, U: CSR, B: CSR, k: int, xj: np.ndarray, x: np.ndarray, qinv: np.ndarray) -> int:
        m = U.m
        assert m == B.m == len(qinv)
        assert 0 <= k < B.n
        assert not xj.any()
        assert len(xj) >= 3 * m and len(x) >= m
        return int(self.lib.spasm_sparse_triangular_solve(U.data, B.data, k, _i32ptr(xj), _i32ptr(x), _i32ptr(qinv)))

    def sparse_triangular_solve(self, U, B: CSR, qinv: np.ndarray | None = None):
        """X with X*U == B, or None if some row has no solution (src/SpaSM.jl:733-755)."""
        import scipy.sparse as sp

        if isinstance(U, LU):
            qinv, U = U.qinv, U.U
        qinv = np.ascontiguousarray(qinv, dtype=np.int32)
        m = U.m
        xj = np.zeros(3 * m, dtype=np.int32)
        x = np.zeros(m, dtype=np.int32)
        indptr, idx, val = [0], [], []
        for k in range(B.n):
            xj[:] = 0
            top = self.sparse_triangular_solve_row(U, B, k, xj, x, qinv)
            for t in range(top, m):
                j = int(xj[t])
                if x[j] == 0:
                    continue
                if qinv[j] < 0:
                    return None
                idx.append(int(qinv[j]))
                val.append(int(x[j]))
            indptr.append(len(idx))
        Xt = sp.csc_matrix((np.array(val, dtype=np.int64), np.array(idx, dtype=np.int64), np.array(indptr)), shape=(U.n, B.n))
        return self.CSR(Xt, U.prime)

    # ---- src/SpaSM.jl:817-866
    def EchelonizeOpts(self, **kw) -> EchelonizeOpts:
        o = EchelonizeOpts()
        self.lib.spasm_echelonize_init_opts(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(f"type EchelonizeOpts has no field {k}")
            setattr(o, k, v)
        return o

    def echelonize(self, A: CSR, opts: EchelonizeOpts | None = None, verbose=False, **kw) -> LU:
        if opts is None:
            opts = self.EchelonizeOpts()
        for k, v in kw.items():
            if not hasattr(opts, k):
                raise AttributeError(f"type EchelonizeOpts has no field {k}")
            setattr(opts, k, v)
        quiet = not (verbose if isinstance(verbose, bool) else A.nnz() >= verbose)
        with _quiet(self, quiet):
            return LU(self, self.lib.spasm_echelonize(A.data, C.byref(opts)))

    def rank(self, A, **kw) -> int:  # src/SpaSM.jl:305, :1149
        return (A if isinstance(A, LU) else self.echelonize(A, **kw)).r

    def rref(self, fact: LU, Rqinv: np.ndarray) -> CSR:  # src/SpaSM.jl:871
        assert len(Rqinv) >= fact.U.m
        return CSR(self, self.lib.spasm_rref(fact.data, _i32ptr(Rqinv)))

    def kernel(self, fact, verbose=False, **kw) -> CSR:  # src/SpaSM.jl:876-882, :1147
        if isinstance(fact, CSR):
            fact = self.echelonize(fact, **kw)
        quiet = not (verbose if isinstance(verbose, bool) else fact.U.nnz() >= verbose)
        with _quiet(self, quiet):
            return CSR(self, self.lib.spasm_kernel(fact.data))

    def solve(self, fact: LU, b, x=None):  # src/SpaSM.jl:895-905
        fact.L  # force it to be non-null
        b = np.ascontiguousarray(b, dtype=np.int32)
        assert len(b) == fact.U.m
        n = fact.L.n  # rows of A (the reference wrapper allocates U.n entries)
        if x is None:
            x = np.zeros(n, dtype=np.int32)
        else:
            assert len(x) == n
        ok = self.lib.spasm_solve(fact.data, _i32ptr(b), _i32ptr(x))
        return x if ok else None

    def gesv(self, fact: LU, B: CSR, verbose=False):  # src/SpaSM.jl:915-923
        fact.L
        assert B.m == fact.U.m
        ok = np.zeros(B.n, dtype=np.bool_)
        with _quiet(self, not verbose):
            X = CSR(self, self.lib.spasm_gesv(fact.data, B.data, ok.ctypes.data_as(_P(C.c_bool))))
        return X, ok

    # ---- src/blocks.jl
    def Block(self, A: CSR) -> Block:
        """src/blocks.jl:35-105.  Blocks are numbered by their smallest row (row-less blocks last, by column)."""
        import scipy.sparse as sp
        from scipy.sparse.csgraph import connected_components

        n, m = A.shape
        p, j, x = A.arrays()
        rows = np.repeat(np.arange(n), np.diff(p))
        G = sp.coo_matrix((np.ones(len(j), dtype=np.int8), (rows, j.astype(np.int64) + n)), shape=(n + m, n + m))
        ncomp, label = connected_components(G, directed=False)
        first = np.full(ncomp, n + m, dtype=np.int64)
        np.minimum.at(first, label, np.arange(n + m))
        order = np.argsort(first, kind="stable")
        renum = np.empty(ncomp, dtype=np.int64)
        renum[order] = np.arange(ncomp)
        label = renum[label]
        block2row = [np.nonzero(label[:n] == b)[0].astype(np.int32) for b in range(ncomp)]
        block2col = [np.nonzero(label[n:] == b)[0].astype(np.int32) for b in range(ncomp)]
        colpos = np.zeros(m, dtype=np.int32)
        for b in range(ncomp):
            colpos[block2col[b]] = np.arange(len(block2col[b]), dtype=np.int32)
        blocks = []
        for b in range(ncomp):
            rs = block2row[b]
            lens = (p[rs + 1] - p[rs]) if len(rs) else np.zeros(0, dtype=np.int64)
            sp_ = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            idx = np.concatenate([np.arange(p[r], p[r + 1]) for r in rs]) if len(rs) and sp_[-1] else np.zeros(0, dtype=np.int64)
            blocks.append(self.from_arrays(len(rs), len(block2col[b]), sp_, colpos[j[idx]] if len(idx) else np.zeros(0, np.int32),
                                           x[idx] if len(idx) else np.zeros(0, np.int32), A.prime))
        return Block(blocks, block2row, block2col, (n, m))

    @staticmethod
    def block_owner(block: Block, world: int) -> np.ndarray:
        """Owner rank of every block when independent blocks are spread over `world` processes (one GPU each):
        blocks by decreasing number of non-zeros (ties: block number), each to the least loaded rank so far
        (ties: lowest rank).  Pure host arithmetic, identical on every rank; no communication is needed because
        the blocks of src/blocks.jl share neither rows nor columns."""
        w = np.array([Bk.nnz() if isinstance(Bk, CSR) else Bk.U.nnz() for Bk in block.blocks], dtype=np.int64)
        owner = np.zeros(len(w), dtype=np.int32)
        load = np.zeros(max(world, 1), dtype=np.int64)
        for b in sorted(range(len(w)), key=lambda b: (-int(w[b]), b)):
            r = int(np.argmin(load))
            owner[b] = r
            load[r] += max(int(w[b]), 1)
        return owner

    def block_echelonize(self, block: Block, part=None, **kw) -> Block:  # src/blocks.jl:107-115 (blocks run one after the other)
        """part=(rank, world): factor only the blocks block_owner() gives to `rank`; the others stay None."""
        if part is None:
            fs = [self.echelonize(Bk, **kw) for Bk in block.blocks]
        else:
            owner = self.block_owner(block, part[1])
            fs = [self.echelonize(Bk, **kw) if owner[b] == part[0] else None for b, Bk in enumerate(block.blocks)]
        return Block(fs, block.block2row, block.block2col, block.shape)

    def block_rank(self, block: Block, part=None, **kw) -> int:  # src/blocks.jl:117
        """with part=(rank, world): the sum over this rank's blocks only — add the partial ranks up (all_reduce)"""
        owner = self.block_owner(block, part[1]) if part is not None else None
        mine = [Bk for b, Bk in enumerate(block.blocks) if Bk is not None and (owner is None or owner[b] == part[0])]
        return sum((Bk.r if isinstance(Bk, LU) else self.rank(Bk, **kw)) for Bk in mine)

    def block_kernel(self, block: Block, **kw) -> Block:  # src/blocks.jl:119-139
        if block.blocks and isinstance(block.blocks[0], CSR):
            block = self.block_echelonize(block, **kw)
        ks = [self.kernel(f) for f in block.blocks]
        b2r, start = [], 0
        for k in ks:
            b2r.append(np.arange(start, start + k.n, dtype=np.int32))
            start += k.n
        return Block(ks, b2r, block.block2col, (start, block.shape[1]))

    def block_CSR(self, block: Block) -> CSR:  # src/blocks.jl:143-170: reassemble with global column numbers
        n, m = block.shape
        owner = np.zeros((n, 2), dtype=np.int64)
        for b, rs in enumerate(block.block2row):
            owner[rs, 0] = b
            owner[rs, 1] = np.arange(len(rs))
        pp, jj, xx = [0], [], []
        parts = [Bk.arrays() for Bk in block.blocks]
        for i in range(n):
            b, si = owner[i]
            bp, bj, bx = parts[b]
            jj.append(block.block2col[b][bj[bp[si] : bp[si + 1]]])
            xx.append(bx[bp[si] : bp[si + 1]])
            pp.append(pp[-1] + int(bp[si + 1] - bp[si]))
        prime = block.blocks[0].prime if block.blocks else PRIME0
        return self.from_arrays(n, m, np.array(pp, dtype=np.int64), np.concatenate(jj) if jj else np.zeros(0, np.int32),
                                np.concatenate(xx) if xx else np.zeros(0, np.int32), prime)

    def dense_rref(self, prime: int, A: np.ndarray):
        """in-place RREF of a C-contiguous int32 matrix; returns (rank, pivcol[:rank])"""
        assert A.dtype == np.int32 and A.flags["C_CONTIGUOUS"] and A.ndim == 2
        n, m = A.shape
        piv = np.zeros(max(n, 1), dtype=np.int32)
        r = int(self.lib.spasm_dense_rref(prime, n, m, _i32ptr(A.reshape(-1)), m, _i32ptr(piv)))
        return r, piv[:r].copy()


class Block:
    """Block-diagonal decomposition of a matrix by the connected components of its row/column graph
    (src/blocks.jl:1-105).  blocks[b] is a CSR / LU / kernel of component b; block2row / block2col list the
    global 0-based rows / columns of each block in increasing order (a block may have no row or no column)."""

    def __init__(self, blocks, block2row, block2col, shape):
        self.blocks, self.block2row, self.block2col, self.shape = blocks, block2row, block2col, shape

    def __len__(self):
        return len(self.blocks)


class _quiet:
    """Silence the library's progress text unless verbose — the role of capture_stderr
    (src/SpaSM.jl:838-858), done through the log callback instead of redirecting fd 2."""

    def __init__(self, api: SpaSM, active: bool):
        self.api, self.active = api, active and api._logcb is None

    def __enter__(self):
        if self.active:
            self.api.log(False)

    def __exit__(self, *a):
        if self.active:
            self.api.log(None)


_default: SpaSM | None = None


def default() -> SpaSM:
    """The product library (CUDA).  Raises if it has not been built."""
    global _default
    if _default is None:
        _default = SpaSM()
    return _default
