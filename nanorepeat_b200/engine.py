"""ctypes binding of libnanorepeat_b200.so (C ABI in include/nanorepeat_b200.h).

The library is built in-tree by __graft_entry__.build() / `make -C nanorepeat_b200/csrc`.  A missing library or
a missing GPU is an error -- there is no fallback path.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnanorepeat_b200.so")

NR_OK = 0
ERROR_NAMES = {-1: "NR_ERR_CUDA", -2: "NR_ERR_ARG", -3: "NR_ERR_BAD_BASE", -4: "NR_ERR_TOO_LARGE",
               -5: "NR_ERR_NOMEM", -6: "NR_ERR_UNKNOWN_TYPE"}


class Scoring(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("match", "mismatch", "gap_open1", "gap_ext1", "gap_open2", "gap_ext2", "ambiguous", "min_dp_score")]


class Stats(ctypes.Structure):
    _fields_ = [("algorithmic_cells", ctypes.c_int64), ("executed_cells", ctypes.c_int64),
                ("n_tasks", ctypes.c_int64), ("kernel_launches", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


ALN_DTYPE = np.dtype([("score", "<i4"), ("tstart", "<i4"), ("tend", "<i4")])
RUNG_DTYPE = np.dtype([("score", "<i4"), ("starts_in_left", "u1"), ("ends_in_right", "u1"), ("pad", "u1", (2,))])


class NanoRepeatB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


_lib = None

_cpp = ctypes.POINTER(ctypes.c_char_p)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_scp = ctypes.POINTER(Scoring)

# name -> (restype, argtypes): every symbol include/nanorepeat_b200.h declares
SYMBOLS = {
    "nr_get_preset": (ctypes.c_int, [ctypes.c_char_p, _scp]),
    "nr_init": (ctypes.c_int, [ctypes.c_int]),
    "nr_shutdown": (ctypes.c_int, []),
    "nr_last_error": (ctypes.c_char_p, []),
    "nr_device_info": (ctypes.c_int, [_i32p, _i32p, _i32p]),
    "nr_limits": (ctypes.c_int, [_i32p, _i32p]),
    "nr_set_ladder_mode": (ctypes.c_int, [ctypes.c_int]),
    "nr_score_tasks": (ctypes.c_int, [_scp, ctypes.c_int32, _cpp, _i32p, _cpp, _i32p, ctypes.c_void_p]),
    "nr_round2_region": (ctypes.c_int, [_scp, ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_int32, _cpp, _i32p, ctypes.c_void_p]),
    "nr_round3_region": (ctypes.c_int, [_scp, ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                                        ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, _cpp, _i32p, _i32p, _i32p,
                                        _i64p, ctypes.c_void_p, _i64p, _i32p, _i32p]),
    "nr_batch_create_tasks": (ctypes.c_void_p, [_scp, ctypes.c_int32, _cpp, _i32p, _cpp, _i32p]),
    "nr_batch_create_round2": (ctypes.c_void_p, [_scp, ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p,
                                                 ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _cpp, _i32p]),
    "nr_batch_create_round3": (ctypes.c_void_p, [_scp, ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p,
                                                 ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32,
                                                 _cpp, _i32p, _i32p, _i32p]),
    "nr_batch_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "nr_batch_fetch_alns": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "nr_batch_fetch_round3": (ctypes.c_int, [ctypes.c_void_p, _i64p, ctypes.c_void_p, _i64p, _i32p, _i32p]),
    "nr_batch_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Stats)]),
    "nr_batch_destroy": (None, [ctypes.c_void_p]),
    "nr_last_stats": (ctypes.c_int, [ctypes.POINTER(Stats)]),
}


def lib():
    """Load the shared library (no CUDA call happens at load time)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nanorepeat_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != NR_OK:
        raise NanoRepeatB200Error(rc, lib().nr_last_error().decode(errors="replace"))


def get_preset(data_type):
    sc = Scoring()
    rc = lib().nr_get_preset(data_type.encode(), ctypes.byref(sc))
    if rc != NR_OK:
        raise ValueError(lib().nr_last_error().decode())
    return sc


def init(device=-1):
    _check(lib().nr_init(int(device)))


def device_info():
    d, s, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _check(lib().nr_device_info(ctypes.byref(d), ctypes.byref(s), ctypes.byref(c)))
    return dict(device=d.value, sm_count=s.value, clock_khz=c.value)


def set_ladder_mode(mode):
    """1 (default): round 3 shares one backward + one forward sweep across a read's ladder; 0: every rung is its
    own full rectangle.  Same results either way."""
    _check(lib().nr_set_ladder_mode(int(mode)))


def last_stats():
    st = Stats()
    _check(lib().nr_last_stats(ctypes.byref(st)))
    return st.as_dict()


def _cstrs(seqs):
    bs = [s.encode() if isinstance(s, str) else s for s in seqs]
    arr = (ctypes.c_char_p * len(bs))(*bs)
    lens = np.fromiter((len(b) for b in bs), dtype=np.int32, count=len(bs))
    return bs, arr, lens


def _b(s):
    return s.encode() if isinstance(s, str) else s


def score_tasks(queries, targets, sc):
    """Generic engine: (score, tstart, tend) for every (query, target) pair."""
    n = len(queries)
    if n != len(targets):
        raise ValueError("queries and targets differ in length")
    out = np.zeros(n, dtype=ALN_DTYPE)
    _qb, qa, ql = _cstrs(queries)
    _tb, ta, tl = _cstrs(targets)
    _check(lib().nr_score_tasks(ctypes.byref(sc), n, qa, ql.ctypes.data_as(_i32p), ta, tl.ctypes.data_as(_i32p),
                                out.ctypes.data))
    return out


def round2_region(sc, left, motif, T, cores):
    n = len(cores)
    out = np.zeros(n, dtype=ALN_DTYPE)
    _cb, ca, cl = _cstrs(cores)
    lb, mb = _b(left), _b(motif)
    _check(lib().nr_round2_region(ctypes.byref(sc), lb, len(lb), mb, len(mb), int(T), n, ca,
                                  cl.ctypes.data_as(_i32p), out.ctypes.data))
    return out


def rung_offsets(kmin, kmax):
    n_rungs = np.maximum(kmax.astype(np.int64) - kmin.astype(np.int64) + 1, 0)
    off = np.zeros(len(kmin) + 1, dtype=np.int64)
    np.cumsum(n_rungs, out=off[1:])
    return off


def round3_region(sc, left, right, motif, cores, kmin, kmax, want_rungs=False):
    """-> (sum_k, n_k, top_score[, rungs, rung_offset])."""
    n = len(cores)
    kmin = np.ascontiguousarray(kmin, dtype=np.int32)
    kmax = np.ascontiguousarray(kmax, dtype=np.int32)
    sum_k = np.zeros(n, dtype=np.int64)
    n_k = np.zeros(n, dtype=np.int32)
    top = np.zeros(n, dtype=np.int32)
    off = rung_offsets(kmin, kmax)
    rungs = np.zeros(int(off[-1]), dtype=RUNG_DTYPE) if want_rungs else None
    _cb, ca, cl = _cstrs(cores)
    lb, rb, mb = _b(left), _b(right), _b(motif)
    _check(lib().nr_round3_region(ctypes.byref(sc), lb, len(lb), rb, len(rb), mb, len(mb), n, ca,
                                  cl.ctypes.data_as(_i32p), kmin.ctypes.data_as(_i32p), kmax.ctypes.data_as(_i32p),
                                  off.ctypes.data_as(_i64p) if want_rungs else None,
                                  rungs.ctypes.data if want_rungs else None,
                                  sum_k.ctypes.data_as(_i64p), n_k.ctypes.data_as(_i32p), top.ctypes.data_as(_i32p)))
    if want_rungs:
        return sum_k, n_k, top, rungs, off
    return sum_k, n_k, top


class Batch:
    """Device-resident batch: inputs packed and uploaded at construction, run() launches kernels only."""

    def __init__(self, handle, kind, n_items, keep):
        if not handle:
            raise NanoRepeatB200Error(-2, lib().nr_last_error().decode(errors="replace"))
        self._h = ctypes.c_void_p(handle)
        self.kind = kind
        self.n_items = n_items
        self._keep = keep

    @classmethod
    def tasks(cls, sc, queries, targets):
        _qb, qa, ql = _cstrs(queries)
        _tb, ta, tl = _cstrs(targets)
        h = lib().nr_batch_create_tasks(ctypes.byref(sc), len(queries), qa, ql.ctypes.data_as(_i32p), ta,
                                        tl.ctypes.data_as(_i32p))
        return cls(h, "tasks", len(queries), None)

    @classmethod
    def round2(cls, sc, left, motif, T, cores):
        _cb, ca, cl = _cstrs(cores)
        lb, mb = _b(left), _b(motif)
        h = lib().nr_batch_create_round2(ctypes.byref(sc), lb, len(lb), mb, len(mb), int(T), len(cores), ca,
                                         cl.ctypes.data_as(_i32p))
        return cls(h, "round2", len(cores), None)

    @classmethod
    def round3(cls, sc, left, right, motif, cores, kmin, kmax):
        kmin = np.ascontiguousarray(kmin, dtype=np.int32)
        kmax = np.ascontiguousarray(kmax, dtype=np.int32)
        _cb, ca, cl = _cstrs(cores)
        lb, rb, mb = _b(left), _b(right), _b(motif)
        h = lib().nr_batch_create_round3(ctypes.byref(sc), lb, len(lb), rb, len(rb), mb, len(mb), len(cores), ca,
                                         cl.ctypes.data_as(_i32p), kmin.ctypes.data_as(_i32p),
                                         kmax.ctypes.data_as(_i32p))
        return cls(h, "round3", len(cores), (kmin, kmax))

    def run(self, stream=None):
        _check(lib().nr_batch_run(self._h, ctypes.c_void_p(stream) if stream else None))

    def fetch_alns(self):
        st = self.stats()
        out = np.zeros(st["n_tasks"], dtype=ALN_DTYPE)
        _check(lib().nr_batch_fetch_alns(self._h, out.ctypes.data))
        return out

    def fetch_round3(self, want_rungs=False):
        kmin, kmax = self._keep
        n = self.n_items
        sum_k = np.zeros(n, dtype=np.int64)
        n_k = np.zeros(n, dtype=np.int32)
        top = np.zeros(n, dtype=np.int32)
        off = rung_offsets(kmin, kmax)
        rungs = np.zeros(int(off[-1]), dtype=RUNG_DTYPE) if want_rungs else None
        _check(lib().nr_batch_fetch_round3(self._h, off.ctypes.data_as(_i64p) if want_rungs else None,
                                           rungs.ctypes.data if want_rungs else None,
                                           sum_k.ctypes.data_as(_i64p), n_k.ctypes.data_as(_i32p),
                                           top.ctypes.data_as(_i32p)))
        if want_rungs:
            return sum_k, n_k, top, rungs, off
        return sum_k, n_k, top

    def stats(self):
        st = Stats()
        _check(lib().nr_batch_stats(self._h, ctypes.byref(st)))
        return st.as_dict()

    def close(self):
        if self._h:
            lib().nr_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
