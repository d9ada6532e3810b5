"""ctypes binding of oracle/libnr_oracle.so (TEST INFRASTRUCTURE ONLY -- see nr_oracle.c header).

PARITY UNPINNED: the DP restated here is minimap2's published `map-ont` scoring model, the third-party
engine the reference calls at src/NanoRepeat/nanoRepeat_bam.py:362 and :497; it is absent from
/root/reference and from this image, and the reference ships no tests for the path.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnr_oracle.so")


class Scoring(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("match", "mismatch", "gap_open1", "gap_ext1", "gap_open2", "gap_ext2", "ambiguous", "min_dp_score")]


# minimap2 `-x map-ont` scoring (the only preset the reference uses: src/NanoRepeat/tk.py:502-517)
MAP_ONT = dict(match=2, mismatch=4, gap_open1=4, gap_ext1=2, gap_open2=24, gap_ext2=1, ambiguous=1, min_dp_score=80)

ALN_DTYPE = np.dtype([("score", "<i4"), ("tstart", "<i4"), ("tend", "<i4")])


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "nr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        cpp = ctypes.POINTER(ctypes.c_char_p)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.nro_align.argtypes = [ctypes.POINTER(Scoring), ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p,
                                ctypes.c_int32, ctypes.c_void_p]
        L.nro_align_batch.argtypes = [ctypes.POINTER(Scoring), ctypes.c_int32, cpp, i32p, cpp, i32p,
                                      ctypes.c_void_p, ctypes.c_int32]
        L.nro_align_ladders.argtypes = [ctypes.POINTER(Scoring), ctypes.c_int32, cpp, i32p,
                                        ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                                        ctypes.c_char_p, ctypes.c_int32, i32p, i32p,
                                        ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p, ctypes.c_int32]
        L.nro_max_threads.restype = ctypes.c_int
        L.nro_align_window.argtypes = [ctypes.POINTER(Scoring), ctypes.c_char_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
        L.nro_align_window_cigar.argtypes = L.nro_align_window.argtypes + [ctypes.c_char_p, ctypes.c_int32]
        _lib = L
    return _lib


def scoring(**kw):
    d = dict(MAP_ONT)
    d.update(kw)
    return Scoring(**d)


def _cstrs(seqs):
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (ctypes.c_char_p * len(bs))(*bs)
    lens = np.array([len(b) for b in bs], dtype=np.int32)
    return bs, arr, lens


def align(query, target, sc=None):
    """One task -> (score, tstart, tend)."""
    sc = sc or scoring()
    out = np.zeros(1, dtype=ALN_DTYPE)
    q = query.encode() if isinstance(query, str) else query
    t = target.encode() if isinstance(target, str) else target
    rc = lib().nro_align(ctypes.byref(sc), q, len(q), t, len(t), out.ctypes.data)
    if rc != 0:
        raise MemoryError("nro_align failed")
    return int(out["score"][0]), int(out["tstart"][0]), int(out["tend"][0])


def align_batch(queries, targets, sc=None, n_threads=1):
    """Independent (query, target) tasks -> structured array (score, tstart, tend)."""
    sc = sc or scoring()
    n = len(queries)
    assert n == len(targets)
    out = np.zeros(n, dtype=ALN_DTYPE)
    if n == 0:
        return out
    _qb, qa, ql = _cstrs(queries)
    _tb, ta, tl = _cstrs(targets)
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib().nro_align_batch(ctypes.byref(sc), n, qa, ql.ctypes.data_as(i32p), ta, tl.ctypes.data_as(i32p),
                               out.ctypes.data, int(n_threads))
    if rc != 0:
        raise MemoryError("nro_align_batch failed")
    return out


def align_ladders(cores, left, right, motif, kmin, kmax, sc=None, n_threads=1):
    """Round-3 shape: every core against left + motif*k + right for k in [kmin[r], kmax[r]].

    Returns (out, rung_offset): out[rung_offset[r] + (k - kmin[r])] is read r's rung k."""
    sc = sc or scoring()
    n = len(cores)
    kmin = np.ascontiguousarray(kmin, dtype=np.int32)
    kmax = np.ascontiguousarray(kmax, dtype=np.int32)
    n_rungs = np.maximum(kmax.astype(np.int64) - kmin.astype(np.int64) + 1, 0)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(n_rungs, out=off[1:])
    out = np.zeros(int(off[-1]), dtype=ALN_DTYPE)
    if n == 0:
        return out, off
    _cb, ca, cl = _cstrs(cores)
    lb, rb, mb = left.encode(), right.encode(), motif.encode()
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib().nro_align_ladders(ctypes.byref(sc), n, ca, cl.ctypes.data_as(i32p), lb, len(lb), rb, len(rb),
                                 mb, len(mb), kmin.ctypes.data_as(i32p), kmax.ctypes.data_as(i32p),
                                 off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), out.ctypes.data, int(n_threads))
    if rc != 0:
        raise MemoryError("nro_align_ladders failed")
    return out, off


WIN_DTYPE = np.dtype([("score", "<i4"), ("window_score", "<i4"), ("tstart", "<i4"), ("tend", "<i4")])


def align_window(query, target, win_a, win_b, reverse=False, sc=None, want_cigar=False):
    """Joint path (nr_oracle.c, nro_align_window): -> (score, window_score[, tstart, tend, cigar]) of the canonical optimal
    local alignment of query (its reverse complement when reverse) against target, window [win_a, win_b)."""
    sc = sc or scoring()
    out = np.zeros(1, dtype=WIN_DTYPE)
    q = query.encode() if isinstance(query, str) else query
    t = target.encode() if isinstance(target, str) else target
    if not want_cigar:
        rc = lib().nro_align_window(ctypes.byref(sc), q, len(q), t, len(t), int(win_a), int(win_b), int(bool(reverse)), out.ctypes.data)
        if rc != 0:
            raise MemoryError("nro_align_window failed")
        return int(out["score"][0]), int(out["window_score"][0])
    buf = ctypes.create_string_buffer(8 * (len(q) + len(t)) + 64)
    rc = lib().nro_align_window_cigar(ctypes.byref(sc), q, len(q), t, len(t), int(win_a), int(win_b), int(bool(reverse)),
                                      out.ctypes.data, buf, len(buf))
    if rc != 0:
        raise MemoryError(f"nro_align_window_cigar failed ({rc})")
    return int(out["score"][0]), int(out["window_score"][0]), int(out["tstart"][0]), int(out["tend"][0]), buf.value.decode()


def max_threads():
    return int(lib().nro_max_threads())


def align_py(query, target, sc=None):
    """Pure-Python restatement of nro_align for tiny cases (cross-checks the C code; same contract)."""
    p = dict(MAP_ONT) if sc is None else {n: getattr(sc, n) for n, _ in Scoring._fields_}
    a, b = p["match"], p["mismatch"]
    qe1, e1 = p["gap_open1"] + p["gap_ext1"], p["gap_ext1"]
    qe2, e2 = p["gap_open2"] + p["gap_ext2"], p["gap_ext2"]
    n, m = len(query), len(target)
    if n == 0 or m == 0:
        return 0, 0, 0
    acgt = set("ACGT")
    H = [(0, 0)] * (n + 1)
    E1 = [(-qe1, 0)] * (n + 1)
    E2 = [(-qe2, 0)] * (n + 1)
    best = (0, 0, 0)  # score, -tend, start ; python tuples compare lexicographically
    for j in range(1, m + 1):
        tc = target[j - 1].upper()
        fresh = (0, j)
        hdiag = H[0]
        H[0] = fresh
        f1 = (-qe1, j)
        f2 = (-qe2, j)
        for i in range(1, n + 1):
            qc = query[i - 1].upper()
            if tc not in acgt or qc not in acgt:
                s = -p["ambiguous"]
            else:
                s = a if tc == qc else -b
            h = max((hdiag[0] + s, hdiag[1]), E1[i], E2[i], f1, f2, fresh)
            hdiag = H[i]
            H[i] = h
            E1[i] = max((h[0] - qe1, h[1]), (E1[i][0] - e1, E1[i][1]))
            E2[i] = max((h[0] - qe2, h[1]), (E2[i][0] - e2, E2[i][1]))
            f1 = max((h[0] - qe1, h[1]), (f1[0] - e1, f1[1]))
            f2 = max((h[0] - qe2, h[1]), (f2[0] - e2, f2[1]))
            if h[0] > 0:
                best = max(best, (h[0], -j, h[1]))
    if best[0] <= 0:
        return 0, 0, 0
    return best[0], best[2], -best[1]
