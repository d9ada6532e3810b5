"""Host-side mirror of the reference's hot-path operators, backed by the CUDA library.

    round1_and_round2_estimation   <- reference src/NanoRepeat/nanoRepeat_bam.py:334-393
    round3_estimation              <- reference src/NanoRepeat/nanoRepeat_bam.py:446-450 (+ :452-500, :408-434)

Same names, arguments, attribute side effects (Read.round{1,2,3}_repeat_size) and degenerate-input behaviour;
`num_cpu` is accepted and ignored (it was minimap2's -t).  What changed: one region-batched call into
libnanorepeat_b200.so per round instead of temp files + one minimap2 process per read + PAF text.
All float arithmetic that decides results (r1, T, r2, ladder bounds with int() truncation, np.mean of the tied
rungs) stays here in Python/numpy float64, written exactly as the reference writes it.
"""
import weakref

import os

import numpy as np

from . import engine
from .presets import get_preset_for_minimap2


# What round1_and_round2_estimation leaves behind for round3_estimation on the same region: the committed round-2 batch
# (its packed reads and kept DP state stay on the device) and what round 2 decided.  Kept HERE, keyed weakly by the
# region object, never as an attribute of it: the reference pickles RepeatRegion objects through a multiprocessing
# queue (nanoRepeat_bam.py:610) and a ctypes handle cannot be pickled.
_ROUND2_CACHE = weakref.WeakKeyDictionary()


def _cache_put(rr, value):
    try:
        _ROUND2_CACHE[rr] = value
    except TypeError:           # not hashable / not weakly referenceable: round 3 packs the reads again
        pass


def _cache_pop(rr):
    try:
        return _ROUND2_CACHE.pop(rr, None)
    except TypeError:
        return None


def _scoring_for(data_type):
    get_preset_for_minimap2(data_type)        # same unknown-type behaviour as tk.py:514-516 (exit 1)
    return engine.get_preset(data_type)


def _cores_of(rr, qnames):
    return list(map(rr.read_core_seq_dict.__getitem__, qnames))


def _round2_launch(data_type, repeat_regions):
    """Round 1 and the launch of round 2 (reference nanoRepeat_bam.py:334-362) for a list of regions, one engine launch
    for all.  Returns the context _round2_finish() needs, or None when there is nothing to align."""
    sc = _scoring_for(data_type)
    min_score = max(1, sc.min_dp_score)
    specs, todo = [], []
    for rr in repeat_regions:
        reads = rr.read_dict
        if len(reads) == 0:
            continue                                                            # :336
        motif_len = len(rr.repeat_unit_seq)
        read_list = list(reads.values())
        # :339-342  r1 = float(dist) / len(motif): the same IEEE division, done on the whole array
        r1 = (np.array([rd.dist_between_anchors for rd in read_list], dtype=np.float64) / np.float64(motif_len)).tolist()
        for rd, v in zip(read_list, r1):
            rd.round1_repeat_size = v
        max_r1 = max(r1)
        if getattr(rr, "round1_max_dist", None) is not None:                    # a piece of a split region
            max_r1 = max(max_r1, float(rr.round1_max_dist) / motif_len)         # (sharding.split_region): T is region-wide
        template_repeat_size = int(max_r1 * 1.5) + 1                            # :344
        if template_repeat_size < max_r1 + 10:                                  # :346-347
            template_repeat_size = int(max_r1 + 10)
        # reads come from read_core_seq_dict, which is what the reference wrote to core_sequences.fastq (:311-321)
        core_dict = rr.read_core_seq_dict
        qnames = list(reads) if core_dict.keys() >= reads.keys() else [n for n in reads if n in core_dict]
        specs.append((rr.left_anchor_seq, rr.repeat_unit_seq, template_repeat_size))
        todo.append((rr, qnames))
    if not specs:
        return None
    b = engine.Batch.begin(sc, "round2_flags")      # the selection reads AS, tend and tstart <= |left| only (:373-384)
    for (left, motif, T), (rr, qnames) in zip(specs, todo):
        # white space around a core is dropped by the library (the reference's FASTQ round trip drops it, :311-321); a
        # base other than ACGT is scored as minimap2 scores N; an anchor with such a base or a core beyond the packed
        # range leaves its reads without a size instead of failing the call (include/nanorepeat_b200.h, "Sequences")
        cores = _cores_of(rr, qnames)
        try:
            b.add_round2(left, motif, T, cores)
        except engine.NanoRepeatB200Error as e:
            if e.code != -3:
                raise
            b.add_round2(left, motif, T, [c.strip() for c in cores], lines=False)      # a core with a newline inside
    b.commit().run()                                                            # was pymm2.main at :362; asynchronous
    return b, todo, specs, min_score


def _round2_finish(ctx):
    """Round-2 selection (reference nanoRepeat_bam.py:364-384).  The committed batch is remembered per region
    (_ROUND2_CACHE) so that round 3 can reuse the packed reads."""
    if ctx is None:
        return
    b, todo, specs, min_score = ctx
    all_score, all_tend, all_inside = b.fetch_round2()
    # selection over every read of every region at once: no PAF line below -s; span test :373; r2 :375 (the same IEEE
    # division per element as the reference's float(tend - |left|) / |motif|)
    counts = np.fromiter((len(q) for _rr, q in todo), dtype=np.int64, count=len(todo))
    n_left_all = np.repeat(np.fromiter((len(left) for left, _m, _T in specs), dtype=np.int64, count=len(specs)), counts)
    m_all = np.repeat(np.fromiter((len(motif) for _l, motif, _T in specs), dtype=np.float64, count=len(specs)), counts)
    ok_all = (all_score >= min_score) & all_inside & (all_tend >= n_left_all)
    r2_all = (all_tend - n_left_all).astype(np.float64) / m_all
    pos = 0
    for idx, (rr, qnames) in enumerate(todo):
        n = len(qnames)
        ok, r2_arr = ok_all[pos:pos + n], r2_all[pos:pos + n]
        pos += n
        reads = rr.read_dict
        for name, good, v in zip(qnames, ok.tolist(), r2_arr.tolist()):
            if good:
                reads[name].round2_repeat_size = v
        # the committed batch (round 3 reuses its packed reads and kept DP state) and what round 2 just decided, so that
        # a round 3 issued right behind it (estimate_regions) need not read 5 000 attributes back
        _cache_put(rr, (b, idx, qnames, ok, r2_arr))


def _round2_many(data_type, repeat_regions):
    _round2_finish(_round2_launch(data_type, repeat_regions))


def round1_and_round2_estimation(data_type, repeat_region, num_cpu=1):
    """Rounds 1 and 2 for one region (reference nanoRepeat_bam.py:334-393)."""
    _round2_many(data_type, [repeat_region])
    return


def ladder_bounds(round2_repeat_size, fast_mode):
    """Reference nanoRepeat_bam.py:463-472, verbatim arithmetic."""
    buffer = max(15, int(round2_repeat_size * 0.05))
    if buffer > 150:
        buffer = 150
    if fast_mode:
        buffer = 15
    max_template_repeat_size = int(round2_repeat_size + buffer)
    min_template_repeat_size = int(round2_repeat_size - buffer)
    if min_template_repeat_size < 0:
        min_template_repeat_size = 0
    return min_template_repeat_size, max_template_repeat_size


def ladder_bounds_array(r2, fast_mode):
    """ladder_bounds over a float64 array.  Same IEEE double arithmetic; np.trunc + cast is Python's int() for these
    magnitudes (tests/test_host.py checks equality with the scalar form)."""
    r2 = np.asarray(r2, dtype=np.float64)
    buf = np.maximum(15, np.trunc(r2 * 0.05).astype(np.int64))
    buf = np.minimum(buf, 150)
    if fast_mode:
        buf = np.full_like(buf, 15)
    kmax = np.trunc(r2 + buf).astype(np.int64)
    kmin = np.maximum(np.trunc(r2 - buf).astype(np.int64), 0)
    return kmin.astype(np.int32), kmax.astype(np.int32)


def round3_estimation_for1read(read, n_k, sum_k, top_score, best_k_list=None):
    """Reference nanoRepeat_bam.py:408-434 on the binary record instead of PAF text."""
    if top_score <= 0:
        return                                                                  # no PAF line at all (:421)
    if n_k > 0:
        if best_k_list is None:
            read.round3_repeat_size = np.float64(sum_k) / np.float64(n_k)
        else:
            read.round3_repeat_size = np.mean(best_k_list)                      # :431
    else:
        read.round3_repeat_size = read.round2_repeat_size                       # :433


def _assign_round3(reads, sum_k, n_k, top):
    # np.mean(list of k) == float64(sum k) / n exactly: the k are small integers, so every partial sum is an
    # exactly representable float64 and numpy's pairwise summation cannot round (tests/test_host.py checks it)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_k = sum_k.astype(np.float64) / n_k.astype(np.float64)
    state = np.where(top <= 0, 0, np.where(n_k > 0, 1, 2)).tolist()
    for read, s, v in zip(reads, state, mean_k):
        if read is None:
            continue
        if s == 1:
            read.round3_repeat_size = v                                         # :431 (np.float64, like np.mean)
        elif s == 2:
            read.round3_repeat_size = read.round2_repeat_size                   # :433
        # s == 0: no PAF line at all (:421) -> untouched


def _round3_reuse_launch(fast_mode, batch, items):
    """Launch of round 3 over the reads a committed round-2 batch already holds on the device.  items: (region, what
    round 2 cached for it).  The sizes are read back from the Read objects: a caller of the two separate operators may
    have edited them in between, as the reference's attributes allow."""
    b3 = engine.Batch.begin_round3_from(batch)
    per_region, valid_parts, r2_parts = [], [], []
    for rr, cached in items:
        _b, idx, qnames, _ok2, _r2_arr = cached
        reads = rr.read_dict
        rl = [reads.get(n) for n in qnames]                                     # a read dropped since round 2 is skipped
        r2 = [None if rd is None else rd.round2_repeat_size for rd in rl]
        valid = np.array([v is not None for v in r2], dtype=bool)               # :460
        r2_valid = np.array([v for v in r2 if v is not None], dtype=np.float64)
        per_region.append((rr, idx, rl, valid))
        valid_parts.append(valid)
        r2_parts.append(r2_valid)
    # ladder bounds of every read of every region in one pass (:463-472)
    valid_all = np.concatenate(valid_parts) if valid_parts else np.zeros(0, dtype=bool)
    kmin_all = np.zeros(len(valid_all), dtype=np.int32)
    kmax_all = np.full(len(valid_all), -1, dtype=np.int32)
    if valid_all.any():
        kmin_all[valid_all], kmax_all[valid_all] = ladder_bounds_array(np.concatenate(r2_parts), fast_mode)
    all_reads, pos = [], 0
    for rr, idx, rl, valid in per_region:
        n = len(rl)
        b3.add_round3_reuse(idx, rr.right_anchor_seq, kmin_all[pos:pos + n], kmax_all[pos:pos + n])
        pos += n
        all_reads += [rd if ok else None for rd, ok in zip(rl, valid.tolist())]
    b3.commit().run()                                                           # was pymm2.main per read at :497
    return b3, all_reads


def _round3_reuse_finish(ctx):
    b3, all_reads = ctx
    with b3:
        sum_k, n_k, top = b3.fetch_round3()
    _assign_round3(all_reads, sum_k, n_k, top)


def _round3_many(data_type, fast_mode, repeat_regions):
    """Round 3 (reference nanoRepeat_bam.py:446-500 + :408-434) for a list of regions, one engine launch for all."""
    sc = _scoring_for(data_type)
    fresh, by_batch = [], {}
    for rr in repeat_regions:
        cached = _cache_pop(rr)
        if cached is not None and engine.ladder_mode() != 0 and cached[0]._h:
            by_batch.setdefault(id(cached[0]), (cached[0], []))[1].append((rr, cached))
        else:
            fresh.append(rr)
    pending = [_round3_reuse_launch(fast_mode, batch, items) for batch, items in by_batch.values()]
    for ctx in pending:
        _round3_reuse_finish(ctx)
    specs, todo = [], []
    for rr in fresh:
        reads, r2, cores = [], [], []
        core_dict = rr.read_core_seq_dict
        for read_name, read in rr.read_dict.items():                            # :457-461
            if read.round2_repeat_size is None:
                continue
            reads.append(read)
            r2.append(read.round2_repeat_size)
            cores.append(core_dict[read_name].strip())                          # :487-491
        if not reads:
            continue
        kmin, kmax = ladder_bounds_array(r2, fast_mode)                         # :463-472
        specs.append((rr.left_anchor_seq, rr.right_anchor_seq, rr.repeat_unit_seq, cores, kmin, kmax))
        todo.extend(reads)
    if not specs:
        return
    sum_k, n_k, top = engine.round3_regions(sc, specs)                          # was pymm2.main per read at :497
    _assign_round3(todo, sum_k, n_k, top)


def round3_estimation(data_type, fast_mode, repeat_region, num_cpu=1):
    """Round 3 for one region (reference nanoRepeat_bam.py:446-450)."""
    _round3_many(data_type, fast_mode, [repeat_region])
    return


def _gather_chunk(rrs):
    """What nr_estimate_regions needs of a list of regions, column by column, plus the Read objects in result order."""
    lefts, rights, motifs, cores, counts, max_dists, todo = [], [], [], [], [], [], []
    any_max = False
    for rr in rrs:
        reads = rr.read_dict
        if len(reads) == 0:
            continue                                                            # :336
        core_dict = rr.read_core_seq_dict
        # reads come from read_core_seq_dict, which is what the reference wrote to core_sequences.fastq (:311-321); a read
        # without a core there gets round 1 only (it is not in the FASTQ the reference aligns)
        max_dist = getattr(rr, "round1_max_dist", None)                         # a piece of a split region (sharding)
        if core_dict.keys() >= reads.keys():
            read_list = list(reads.values())
            cores.extend(map(core_dict.__getitem__, reads))
        else:
            m = len(rr.repeat_unit_seq)
            extra = [rd for n, rd in reads.items() if n not in core_dict]
            for rd in extra:
                rd.round1_repeat_size = float(rd.dist_between_anchors) / m
            far = max(rd.dist_between_anchors for rd in extra)                  # T is over ALL reads of the region (:344)
            max_dist = far if max_dist is None else max(max_dist, far)
            qnames = [n for n in reads if n in core_dict]
            if not qnames:
                continue
            read_list = list(map(reads.__getitem__, qnames))
            cores.extend(map(core_dict.__getitem__, qnames))
        any_max = any_max or max_dist is not None
        lefts.append(rr.left_anchor_seq); rights.append(rr.right_anchor_seq); motifs.append(rr.repeat_unit_seq)
        max_dists.append(max_dist)
        counts.append(len(read_list))
        todo.extend(read_list)
    dists = [rd.dist_between_anchors for rd in todo]
    return (lefts, rights, motifs, cores, dists, max_dists if any_max else None, counts), todo


def _run_chunk(sc, fast_mode, cols, ready=None):
    """One call into the library (ctypes releases the GIL for its duration).  ready: a threading.Event set when this
    thread's Python work is over (the library entered, or the call failed before that)."""
    lefts, rights, motifs, cores, dists, max_dists, counts = cols
    try:
        try:
            return engine.estimate_regions(sc, fast_mode, lefts, rights, motifs, cores, dists, max_dists,
                                           on_ready=ready.set if ready is not None else None, n_reads=counts)
        finally:
            if ready is not None:
                ready.set()
    except engine.NanoRepeatB200Error as e:
        if e.code != -3:
            raise
        # a core with a line break inside or around it (the reads travel as lines): the reference's FASTQ round trip
        # drops white space around a read (:311-321), so do that here and go again
        cores = [c.strip() for c in cores]
        return engine.estimate_regions(sc, fast_mode, lefts, rights, motifs, cores, dists, max_dists, n_reads=counts)


def _assign_chunk(res, todo):
    # (round 3's mean stays an np.float64, like np.mean in the reference, :431; list(array) makes those in one sweep)
    for rd, a, ok, v, s, w in zip(todo, res["r1"].tolist(), res["r2_valid"].tolist(), res["r2"].tolist(),
                                  res["r3_state"].tolist(), list(res["r3"])):
        rd.round1_repeat_size = a                                               # :341
        if ok:
            rd.round2_repeat_size = v                                           # :375-384
            if s == 1:
                rd.round3_repeat_size = w                                       # :431
            elif s == 2:
                rd.round3_repeat_size = v                                       # :433


CHUNK_MIN_READS = 4096      # reads per call into the library when a region list is cut into pipelined chunks
BIG_CHUNK_READS = 16384     # ... for lists of 16 384 reads and more
_POOL = None
_POOL_PID = None


def _chunks(rrs):
    """Contiguous chunks of regions of at least CHUNK_MIN_READS reads each, at most 6."""
    total = sum(len(rr.read_dict) for rr in rrs)
    # A call's Python side (gathering strings before, assigning attributes after) is as long as its kernels; cut in chunks,
    # one chunk's Python side runs while the library and the GPU work on its neighbours.  Chunks of CHUNK_MIN_READS for a
    # mid-sized list (config 2: two regions of 5 000 reads -> two calls in flight), of BIG_CHUNK_READS for a long one
    # (one launch per round and chunk inside the library; smaller launches only cost there).
    per = CHUNK_MIN_READS if total < 4 * CHUNK_MIN_READS else BIG_CHUNK_READS
    n = max(1, min(8, total // per))
    if os.environ.get("NR_PY_CHUNKS"):
        n = max(1, int(os.environ["NR_PY_CHUNKS"]))
    if n > 1:
        # long reads (a kilobase and more on average, judged on a sample): the kernels dwarf the Python side and every
        # launch pays the serial chain of its longest read's stripes, so one call with everything is the fastest
        sample = [len(c) for rr in rrs[::max(1, len(rrs) // 16)] for c in list(rr.read_core_seq_dict.values())[:4]]
        if sample and sum(sample) / len(sample) > 1000:
            n = 1
    if n == 1:
        return [rrs]
    out, cur, acc = [], [], 0
    for rr in rrs:
        cur.append(rr)
        acc += len(rr.read_dict)
        if acc * n >= total * (len(out) + 1) and len(out) < n - 1:
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def _estimate_regions_fused(dt, fast_mode, rrs):
    """Rounds 1-3 of a list of regions through nr_estimate_regions: the library runs round 2, derives every read's ladder
    from it and runs round 3 without coming back here in between; the attributes of every Read are then assigned once
    from the returned arrays.  A long list is cut into a few chunks: while the library (GIL released) and the GPU work
    on one chunk, this thread gathers the next chunk's strings and assigns the previous chunk's results, so that the
    Python side of the boundary -- one attribute read and three attribute writes per Read object -- hides behind the
    kernels."""
    global _POOL, _POOL_PID
    sc = _scoring_for(dt)
    chunks = _chunks(rrs)
    if len(chunks) == 1:
        cols, todo = _gather_chunk(rrs)
        if todo:
            _assign_chunk(_run_chunk(sc, fast_mode, cols), todo)
        return
    if _POOL is None or _POOL_PID != os.getpid():          # (a forked child inherits the object, not its threads)
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=3, thread_name_prefix="nanorepeat_b200")
        _POOL_PID = os.getpid()
    import threading
    pending = []
    for ch in chunks:
        cols, todo = _gather_chunk(ch)
        if todo:
            # hand the chunk over and let the worker build its arguments NOW (it needs the interpreter for that; this
            # thread would otherwise keep it through the next gather and the worker would start half a millisecond late)
            ready = threading.Event()
            pending.append((_POOL.submit(_run_chunk, sc, fast_mode, cols, ready), todo))
            ready.wait()
        while pending and pending[0][0].done():
            fut, td = pending.pop(0)
            _assign_chunk(fut.result(), td)
    for fut, td in pending:
        _assign_chunk(fut.result(), td)


def estimate_regions(regions, data_type=None, fast_mode=False):
    """Rounds 1-3 over a list of RepeatRegion-like objects (the reference runs the two operators per region inside
    quantify1repeat_from_bam, nanoRepeat_bam.py:675-679): one call into the library per data type, which batches the
    regions into a few pipelined launches per round (nr_estimate_regions).  Per-region semantics (T per region, ladder
    per read) are untouched.  Regions may carry their own .data_type; regions of one data type are batched together.
    With nr_set_ladder_mode(0) (every rung its own rectangle: a checking mode) the two operators are run instead."""
    by_type = {}
    for rr in regions:
        by_type.setdefault(data_type or getattr(rr, "data_type", None) or "ont", []).append(rr)
    for dt, rrs in by_type.items():
        if engine.ladder_mode() != 0:
            _estimate_regions_fused(dt, fast_mode, rrs)
            continue
        _round2_many(dt, rrs)
        _round3_many(dt, fast_mode, rrs)
    return regions


def install(nanoRepeat_bam_module, anchoring=False):
    """Patch the reference module in place so the unmodified CLI runs this path:
        import NanoRepeat.nanoRepeat_bam as m; nanorepeat_b200.install(m)
    anchoring=True also replaces Step 1 (find_anchor_locations_in_reads, make_core_seq_fastq; nanoRepeat_bam.py:260-331)
    by the GPU version -- opt-in, because minimap2's secondary-hit / mapq heuristics are replaced by a stated rule there
    (nanorepeat_b200/anchoring.py).
    """
    nanoRepeat_bam_module.round1_and_round2_estimation = round1_and_round2_estimation
    nanoRepeat_bam_module.round3_estimation = round3_estimation
    if anchoring:
        from . import anchoring as _anchoring
        read_class = getattr(nanoRepeat_bam_module, "Read", None)

        def find_anchor_locations_in_reads(data_type, repeat_region, num_cpu):
            return _anchoring.find_anchor_locations_in_reads(data_type, repeat_region, num_cpu, read_class=read_class)

        nanoRepeat_bam_module.find_anchor_locations_in_reads = find_anchor_locations_in_reads
        nanoRepeat_bam_module.make_core_seq_fastq = _anchoring.make_core_seq_fastq
    return nanoRepeat_bam_module
