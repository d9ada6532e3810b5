"""CPU restatement of the reference's task generation and selection rules (TEST INFRASTRUCTURE ONLY).

Follows reference src/NanoRepeat/nanoRepeat_bam.py line by line (cited per function) with the alignment
engine replaced by oracle/nr_oracle (PARITY UNPINNED for the DP -- see nr_oracle.c).  Pinned against the
reference's own unmodified functions by tests/golden/make_golden.py, which imports them from
/root/reference and feeds them PAF text produced from the same oracle DP.
"""
import numpy as np

from . import nr_oracle


def round1(dist_between_anchors, motif_len, max_dist=None):
    """nanoRepeat_bam.py:338-347 -> (r1 list of float, template_repeat_size T).
    max_dist: the whole region's longest dist_between_anchors when the reads are one piece of a split region."""
    r1 = [float(d) / motif_len for d in dist_between_anchors]
    mx = max(r1)
    if max_dist is not None:
        mx = max(mx, float(max_dist) / motif_len)
    T = int(mx * 1.5) + 1
    if T < mx + 10:
        T = int(mx + 10)
    return r1, T


def round2_select(aln, n_left, motif_len, min_dp_score):
    """nanoRepeat_bam.py:373-384 for one read with its single alignment record.

    aln = (AS, tstart, tend).  minimap2 prints no line below -s min_dp_score, and none for score 0."""
    score, tstart, tend = (int(x) for x in aln)
    if score <= 0 or score < min_dp_score:
        return None
    if tstart <= n_left and tend >= n_left:
        return float(tend - n_left) / motif_len
    return None


def ladder_bounds(r2, fast_mode=False):
    """nanoRepeat_bam.py:463-472 -> (kmin, kmax)."""
    buffer = max(15, int(r2 * 0.05))
    if buffer > 150:
        buffer = 150
    if fast_mode:
        buffer = 15
    kmax = int(r2 + buffer)
    kmin = int(r2 - buffer)
    if kmin < 0:
        kmin = 0
    return kmin, kmax


def round3_select(rungs, kmin, n_left, n_right, motif_len, r2, min_dp_score):
    """nanoRepeat_bam.py:408-434 for one read.  rungs[i] = (AS, tstart, tend) of k = kmin + i.

    Returns r3 (np.float64 mean of the tied best rungs that span both flanks, r2 when none spans, None when
    minimap2 would have printed nothing at all)."""
    recs = []
    for i, (score, tstart, tend) in enumerate(rungs):
        score = int(score)
        if score <= 0 or score < min_dp_score:
            continue
        k = kmin + i
        recs.append((score, int(tstart), int(tend), n_left + motif_len * k + n_right, k))
    if not recs:
        return None
    top = max(r[0] for r in recs)
    ks = [k for (s, ts, te, tlen, k) in recs if s == top and ts < n_left and tlen - te < n_right]
    if ks:
        return np.mean(ks)
    return r2


def estimate_region(left, right, motif, cores, dists, fast_mode=False, sc=None, n_threads=1, max_dist=None):
    """Rounds 1-3 for one region -> dict of per-read lists r1, r2, r3, plus T and the ladders."""
    sc = sc or nr_oracle.scoring()
    n = len(cores)
    if n == 0:
        return dict(r1=[], r2=[], r3=[], T=None, kmin=[], kmax=[])
    m = len(motif)
    r1, T = round1(dists, m, max_dist)
    tpl = left + motif * T
    a2 = nr_oracle.align_batch(cores, [tpl] * n, sc, n_threads)
    r2 = [round2_select(a2[i], len(left), m, sc.min_dp_score) for i in range(n)]
    idx = [i for i in range(n) if r2[i] is not None]
    kmin = [None] * n
    kmax = [None] * n
    r3 = [None] * n
    all_rungs = (None, None)
    if idx:
        kb = [ladder_bounds(r2[i], fast_mode) for i in idx]
        out, off = nr_oracle.align_ladders([cores[i] for i in idx], left, right, motif,
                                           [b[0] for b in kb], [b[1] for b in kb], sc, n_threads)
        all_rungs = (out, off)      # records of read idx[j]'s rungs: out[off[j]:off[j + 1]]
        for j, i in enumerate(idx):
            kmin[i], kmax[i] = kb[j]
            rungs = out[off[j]:off[j + 1]]
            r3[i] = round3_select([(r["score"], r["tstart"], r["tend"]) for r in rungs], kmin[i],
                                  len(left), len(right), m, r2[i], sc.min_dp_score)
    return dict(r1=r1, r2=r2, r3=r3, T=T, kmin=kmin, kmax=kmax, round2_aln=a2, round3_idx=idx, round3_rungs=all_rungs)
