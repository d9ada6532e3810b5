"""Host-side mirror of nanoRepeat-joint's grid estimation (reference src/NanoRepeat/nanoRepeat_joint.py), backed by the
CUDA library's joint path (nr_joint_grid: alignment score + window score of the optimal alignment per (read, grid point)).

    round2_grid_points / round3_grid_points   <- which (read, k1, k2) the reference aligns (:397-410, :315-333)
    estimate_two_repeats                      <- estimate_two_repeats_from_paf (:427-478) on binary records
    choose_best_step_size                     <- :345-367

What changed: no FASTQ / FASTA / PAF files and no minimap2 call per grid point (:411-419, :334-343); every distinct
template is packed once and all (read, grid point) tasks of a locus run in one launch; the CIGAR re-scoring of
tk.target_region_alignment_stats_from_cigar is carried through the DP (csrc/nr_window_kernel.cuh).  The arithmetic that
decides results (ranges, steps, windows, the means of the tied grid points) is the reference's, in Python floats / ints.
"""
import numpy as np

from . import engine


def choose_best_step_size(repeat_unit_size, count_ranges):
    """nanoRepeat_joint.py:345-367: the coarse grid's step for one repeat.  count_ranges: iterable of (min, max)."""
    max_len = 50
    max_step_size = int(max_len / repeat_unit_size)
    if max_step_size < 1:
        max_step_size = 1
    l = np.mean([b - a for a, b in count_ranges])
    count_list = []
    for size in range(1, max_step_size + 1):
        count = int(l / size) + 1
        count += size * 2 + 2
        count_list.append((size, count))
    count_list.sort(key=lambda x: x[1])
    return count_list[0][0]


def _ranges_array(ranges):
    """per read (min, max) or None -> (lo, hi, present) arrays"""
    n = len(ranges)
    lo, hi, ok = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64), np.zeros(n, dtype=bool)
    for r, a in enumerate(ranges):
        if a is not None:
            lo[r], hi[r], ok[r] = a[0], a[1], True
    return lo, hi, ok


def _points_of(mask1, mask2, K1, K2):
    """mask1[k1 index, read], mask2[k2 index, read] -> the (read, k1, k2) of every set pair, k1 outermost, then k2, then the
    read: the order of the reference's three nested loops."""
    i1, i2, r = np.nonzero(mask1[:, None, :] & mask2[None, :, :])
    return r.astype(np.int32), K1[i1].astype(np.int32), K2[i2].astype(np.int32)


def round2_grid_points(range1, range2, min1, max1, min2, max2, step1, step2):
    """nanoRepeat_joint.py:397-410.  range1 / range2: per read (min, max) of repeat 1 / 2 from the initial estimate
    (None: the read has none); the grid runs k1 = min1..max1 step step1, k2 = min2..max2 step step2 and a read is aligned
    at a point when min <= k < max for both repeats.  -> (point_read, point_k1, point_k2) in the reference's loop order."""
    K1, K2 = np.arange(min1, max1 + 1, step1, dtype=np.int64), np.arange(min2, max2 + 1, step2, dtype=np.int64)
    lo1, hi1, ok1 = _ranges_array(range1)
    lo2, hi2, ok2 = _ranges_array(range2)
    ok = ok1 & ok2
    mask1 = (lo1[None, :] <= K1[:, None]) & (K1[:, None] < hi1[None, :]) & ok[None, :]
    mask2 = (lo2[None, :] <= K2[:, None]) & (K2[:, None] < hi2[None, :])
    return _points_of(mask1, mask2, K1, K2)


def round3_grid_points(range1, range2, size1, size2, buffer1, buffer2):
    """nanoRepeat_joint.py:296-333: the unit-step grid around every read's coarse estimate (size1 / size2: per read the
    round-2 sizes, None when round 2 gave none), clipped to the read's initial ranges."""
    n = len(size1)
    have = np.array([v is not None and w is not None for v, w in zip(size1, size2)], dtype=bool)
    empty = np.zeros(0, dtype=np.int32)
    if not have.any():
        return empty, empty, empty
    v = np.array([float(x) if h else 0.0 for x, h in zip(size1, have)])
    w = np.array([float(x) if h else 0.0 for x, h in zip(size2, have)])
    min_size1 = max(int(v[have].min() - buffer1), 0)
    max_size1 = int(v[have].max() + buffer1 + 2)
    min_size2 = max(int(w[have].min() - buffer2), 0)
    max_size2 = int(w[have].max() + buffer2 + 2)
    K1, K2 = np.arange(min_size1, max_size1, dtype=np.int64), np.arange(min_size2, max_size2, dtype=np.int64)
    lo1, hi1, _ok1 = _ranges_array([a if h else None for a, h in zip(range1, have)])
    lo2, hi2, _ok2 = _ranges_array([b if h else None for b, h in zip(range2, have)])
    k1c, k2c = K1[:, None], K2[:, None]
    mask1 = (k1c >= (v - buffer1)[None, :]) & (k1c < (v + buffer1)[None, :]) & (k1c >= lo1[None, :]) & (k1c < hi1[None, :]) & have[None, :]
    mask2 = (k2c >= (w - buffer2)[None, :]) & (k2c < (w + buffer2)[None, :]) & (k2c >= lo2[None, :]) & (k2c < hi2[None, :])
    return _points_of(mask1, mask2, K1, K2)


def estimate_two_repeats(n_reads, point_read, point_k1, point_k2, records, min_dp_score=80):
    """nanoRepeat_joint.py:427-478 on binary records: per read the grid point(s) with the highest window score among the
    points that have an alignment at all (minimap2 prints no line below -s); the two sizes are the means of the tied
    points' k1 and of their k2, separately (np.mean, as at :471-472: an exact integer sum over a count).  -> (size1, size2):
    lists with None for reads without any alignment."""
    size1, size2 = [None] * n_reads, [None] * n_reads
    pr = np.asarray(point_read, dtype=np.int64)
    if len(pr) == 0:
        return size1, size2
    score = np.asarray(records["score"], dtype=np.int64)
    win = np.asarray(records["window_score"], dtype=np.int64)
    keep = (score > 0) & (score >= min_dp_score)
    pr, win = pr[keep], win[keep]
    k1, k2 = np.asarray(point_k1, dtype=np.int64)[keep], np.asarray(point_k2, dtype=np.int64)[keep]
    top = np.full(n_reads, np.iinfo(np.int64).min, dtype=np.int64)
    np.maximum.at(top, pr, win)
    tied = win == top[pr]
    cnt = np.bincount(pr[tied], minlength=n_reads)
    sum1 = np.bincount(pr[tied], weights=k1[tied], minlength=n_reads)
    sum2 = np.bincount(pr[tied], weights=k2[tied], minlength=n_reads)
    for r in np.nonzero(cnt)[0]:
        size1[r] = sum1[r] / cnt[r]
        size2[r] = sum2[r] / cnt[r]
    return size1, size2


def quantify_two_repeats(reads, left, mid, right, motif1, motif2, range1, range2, max_size1, max_size2, data_type="ont", align=None):
    """Rounds 2 and 3 of nanoRepeat-joint for one locus (fine_tune_read_count, :234-273) on the GPU.
    reads: raw read sequences (either strand); range1 / range2: per read the (min, max) repeat-count ranges of the initial
    estimate (initial_estimate_repeat_size, host side, :509-649).  -> dict(size1, size2, step1, step2).
    align: stand-in for engine.joint_grid with the same signature (a test seam: the default is the CUDA engine)."""
    sc = engine.get_preset(data_type)
    ok1 = [a for a in range1 if a is not None]
    ok2 = [b for b in range2 if b is not None]
    if not ok1 or not ok2:
        return dict(size1=[None] * len(reads), size2=[None] * len(reads), step1=None, step2=None)
    # :239-260  the grid's extent over all reads (each repeat's own dictionary), clipped to the user's maximum sizes
    r1min, r1max = min([max_size1] + [a[0] for a in ok1]), min(max([0] + [a[1] for a in ok1]), max_size1)
    r2min, r2max = min([max_size2] + [b[0] for b in ok2]), min(max([0] + [b[1] for b in ok2]), max_size2)
    step1 = choose_best_step_size(len(motif1), ok1)
    step2 = choose_best_step_size(len(motif2), ok2)
    pr, p1, p2 = round2_grid_points(range1, range2, r1min, r1max, r2min, r2max, step1, step2)
    rec, _strand = (align or engine.joint_grid)(sc, left, mid, right, motif1, motif2, reads, pr, p1, p2)
    size1, size2 = estimate_two_repeats(len(reads), pr, p1, p2, rec, sc.min_dp_score)
    if step1 > 1 and step2 > 1:                                                    # :268-271
        pr, p1, p2 = round3_grid_points(range1, range2, size1, size2, step1, step2)
        rec, _strand = (align or engine.joint_grid)(sc, left, mid, right, motif1, motif2, reads, pr, p1, p2)
        size1, size2 = estimate_two_repeats(len(reads), pr, p1, p2, rec, sc.min_dp_score)
        step1 = step2 = 1
    return dict(size1=size1, size2=size2, step1=step1, step2=step2)
