"""Sharding the hot path across the GPUs of one box (one process per GPU, torch.distributed for the plumbing).

The path is embarrassingly parallel: every (read, rung) task is independent and the only
dependencies are host-side scalars (round 2 -> ladder bounds per read, T per region).  So regions are dealt to ranks
by predicted DP cells -- the reference stripes regions over worker processes the same way, region i -> worker
i mod P (nanoRepeat_bam.py:604) -- each rank runs rounds 1-3 on its own GPU with no data-path collective, and the
per-read results (three numbers per read) are gathered on the host.  A region with very many reads (config 2: 5 000
reads in 2 regions) is first cut into contiguous read batches; T is a region-wide scalar, so it is computed before the
cut and pinned on every piece.
"""
import copy

import numpy as np


def predicted_cells(rr):
    """Executed-cell estimate for one region from what Step 1 hands over (no alignment needed):
    sum over reads of |core| * (round-2 template + backward |R| + forward |L| + m * k)."""
    m = max(1, len(rr.repeat_unit_seq))
    n_left, n_right = len(rr.left_anchor_seq), len(rr.right_anchor_seq)
    dists = [max(0, rd.dist_between_anchors) for rd in rr.read_dict.values()]
    if not dists:
        return 0
    r1max = max(dists) / m
    T = max(int(r1max * 1.5) + 1, int(r1max + 10))
    total = 0
    for name, rd in rr.read_dict.items():
        q = len(rr.read_core_seq_dict.get(name, ""))
        k = max(0, rd.dist_between_anchors) / m
        total += q * (n_left + m * T + n_right + n_left + int(m * (k + max(15, 0.05 * k))))
    return int(total)


def partition(costs, world_size):
    """Longest-processing-time greedy: item indices per rank, deterministic (ties by index), each rank's list
    in increasing index order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda w: (load[w], w))
        parts[r].append(i)
        load[r] += costs[i]
    return [sorted(p) for p in parts]


def split_region(rr, max_reads):
    """Cut one region into pieces of at most max_reads reads (contiguous in read_dict order).  Round 1's T depends
    on the region's maximum r1 (nanoRepeat_bam.py:344-347), so every piece carries the whole region's longest
    dist_between_anchors as `round1_max_dist` for the operator layer to honour."""
    names = list(rr.read_dict)
    if len(names) <= max_reads:
        return [rr]
    max_dist = max(rr.read_dict[n].dist_between_anchors for n in names)
    pieces = []
    for s in range(0, len(names), max_reads):
        p = copy.copy(rr)
        p.read_dict = {n: rr.read_dict[n] for n in names[s:s + max_reads]}
        p.read_core_seq_dict = {n: rr.read_core_seq_dict[n] for n in p.read_dict if n in rr.read_core_seq_dict}
        p.round1_max_dist = max_dist
        pieces.append(p)
    return pieces


def estimate_regions_sharded(regions, data_type=None, fast_mode=False, max_reads_per_piece=2048, estimate_fn=None,
                             rank=None, world_size=None, gather=True):
    """Rounds 1-3 over `regions` with the work split across the ranks of the default process group.

    Every rank passes the same `regions` list (same order).  Rank r computes the pieces dealt to it on its own GPU
    (estimate_fn defaults to nanorepeat_b200.estimate_regions), then -- gather=True -- the per-read results are
    exchanged on the host (all_gather_object of three float arrays per rank, no NCCL data path) and written into every rank's
    Read objects.  Returns the list of piece indices this rank computed."""
    import torch.distributed as dist
    if estimate_fn is None:
        from .estimation import estimate_regions as estimate_fn
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    pieces = []
    for rr in regions:
        for p in split_region(rr, max_reads_per_piece):
            if data_type is None and not hasattr(p, "data_type"):
                p.data_type = "ont"
            pieces.append(p)
    costs = [predicted_cells(p) for p in pieces]
    parts = partition(costs, world_size)
    mine = parts[rank]
    estimate_fn([pieces[i] for i in mine], data_type, fast_mode)
    if gather and world_size > 1:
        # host-side gather: every rank holds the same pieces in the same order, so three float arrays (NaN = None) and a
        # flag array per rank say everything -- no names, no per-read Python objects on the wire
        reads = [rd for i in mine for rd in pieces[i].read_dict.values()]
        nan = float("nan")
        r1 = np.fromiter((nan if r.round1_repeat_size is None else r.round1_repeat_size for r in reads), np.float64, len(reads))
        r2 = np.fromiter((nan if r.round2_repeat_size is None else r.round2_repeat_size for r in reads), np.float64, len(reads))
        r3 = np.fromiter((nan if r.round3_repeat_size is None else r.round3_repeat_size for r in reads), np.float64, len(reads))
        is_np = np.fromiter((isinstance(r.round3_repeat_size, np.floating) for r in reads), np.bool_, len(reads))
        gathered = [None] * world_size
        dist.all_gather_object(gathered, (mine, r1, r2, r3, is_np))
        for src, (theirs, g1, g2, g3, gnp) in enumerate(gathered):
            if src == rank:
                continue
            v1, v2, v3, fl = g1.tolist(), g2.tolist(), g3.tolist(), gnp.tolist()
            pos = 0
            for i in theirs:
                for read in pieces[i].read_dict.values():
                    a, b, c = v1[pos], v2[pos], v3[pos]
                    read.round1_repeat_size = None if a != a else a
                    read.round2_repeat_size = None if b != b else b
                    read.round3_repeat_size = None if c != c else (np.float64(c) if fl[pos] else c)
                    pos += 1
    return mine
