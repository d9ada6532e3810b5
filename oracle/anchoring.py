"""CPU restatement of Step 1 (anchor finding + core extraction) for the tests (TEST INFRASTRUCTURE ONLY).

Follows reference src/NanoRepeat/nanoRepeat_bam.py:165-234 (acceptance rules, distances, core positions) and :308-316
(slicing) with the alignment engine replaced by oracle/nr_oracle (PARITY UNPINNED for the engine: minimap2's seeding,
secondary-hit and mapq heuristics are not modelled; the candidate hits of an anchor are its best exact alignment on each
strand at or above minimap2's -s).  Written independently of nanorepeat_b200/anchoring.py: plain tuples, no shared code.
"""
from . import nr_oracle

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a", "N": "N", "n": "n"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def step1(left_anchor, right_anchor, names, seqs, min_dp_score=80, n_threads=1):
    """-> {name: dict(strand, dist, core_start, core_end, mid_start, mid_end, left_buffer, right_buffer, core, mid)} for
    the reads the reference's rules accept."""
    out = {}
    q, t = [], []
    for s in seqs:
        r = revcomp(s)
        q += [left_anchor, left_anchor, right_anchor, right_anchor]
        t += [s, r, s, r]
    recs = nr_oracle.align_batch(q, t, n_threads=n_threads)
    for i, (name, seq) in enumerate(zip(names, seqs)):
        cands = []
        for a in range(2):
            hits = []
            for s, strand in enumerate("+-"):
                score, ts, te = (int(x) for x in recs[4 * i + 2 * a + s])
                if score > 0 and score >= min_dp_score:
                    hits.append((score, strand, ts, te))
            hits.sort(key=lambda h: -h[0])
            cands.append(hits)

        def good(hits):                                       # :165-179
            if not hits:
                return False
            if len(hits) == 1:
                return True
            if hits[0][3] - hits[0][2] < 10:
                return False
            return hits[0][0] > 1.5 * hits[1][0]

        if not good(cands[0]) or not good(cands[1]):
            continue
        (_ls, lstrand, _lqs, lqe), (_rs, rstrand, rqs, _rqe) = cands[0][0], cands[1][0]
        dist = rqs - lqe if lstrand == rstrand else 0         # :204-207
        if not dist > -10:
            continue
        n = len(seq)
        core_start, core_end = max(lqe - 100, 0), min(rqs + 100, n)
        oriented = revcomp(seq) if lstrand == "-" else seq
        out[name] = dict(strand=lstrand, dist=dist, core_start=core_start, core_end=core_end, mid_start=lqe, mid_end=rqs,
                         left_buffer=lqe - core_start, right_buffer=core_end - rqs, core=oriented[core_start:core_end],
                         mid=oriented[lqe:rqs])
    return out
