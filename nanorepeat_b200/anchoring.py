"""Step 1 of the reference's per-region pipeline on the GPU: locate the two anchors in every read and cut the core.

    find_anchor_locations_in_reads   <- reference src/NanoRepeat/nanoRepeat_bam.py:260-286 (+ :165-258)
    make_core_seq_fastq              <- reference src/NanoRepeat/nanoRepeat_bam.py:288-331

The reference aligns every read of the region to `left_anchor` / `right_anchor` (<= 1000 bp each) with
`minimap2 -c -x map-ont` and keeps a read when both anchors are found unambiguously on one strand (:165-219); the core
handed to rounds 1-3 is read[left.qend - 100 : right.qstart + 100] in the read's orientation (:221-234, :308-316).
Here the four alignments per read (two anchors x two strands) are exact local alignments on the CUDA engine with the
ANCHOR as the DP's query and the READ as its target, so the record's (tstart, tend) are the read coordinates the rules
need (paf.qstart / paf.qend in the read's orientation, paf.py:70-74).

What is the reference's and what is not.  The acceptance rules -- one hit: good; several: best AS > 1.5 x second AS;
both anchors good and on one strand; dist = right.qstart - left.qend > -10; the core / mid slicing with its 100-base
buffers and clamps -- are the reference's, line for line.  What minimap2 decides heuristically is replaced by a stated
rule: the candidate hits of an anchor are its best exact alignment on each strand that reaches minimap2's -s (80), so
"second AS" is the other strand's; `mapq > 30` is taken as true whenever the 1.5 x rule holds (mapq is a function of
exactly that score ratio and chain properties this engine does not have); `align_len < 10` cannot occur above -s.
A read longer than the engine's template limit (65 471 bases) is left out, as a read minimap2 printed nothing for.
"""
import os

from . import engine
from .presets import get_preset_for_minimap2

_COMP = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")
BUFFER_LEN = 100                       # nanoRepeat_bam.py:221


def rev_comp(seq):
    """tk.rev_comp (tk.py:346-355) -- which raises on anything but ACGT; here N stays N."""
    return seq.translate(_COMP)[::-1]


def read_fastq(path):
    """-> ([name, ...], [sequence, ...]) exactly as make_core_seq_fastq walks the file (:298-309)."""
    names, seqs = [], []
    with open(path) as f:
        while True:
            l1, l2, l3, l4 = f.readline(), f.readline(), f.readline(), f.readline()
            if not l1 or not l2 or not l3 or not l4:
                break
            names.append(l1.strip()[1:])
            seqs.append(l2.strip())
    return names, seqs


class AnchorHit:
    """The slice of a PAF record Step 1 reads (paf.py:39-74), coordinates in the read's orientation."""
    __slots__ = ("strand", "align_score", "qstart", "qend", "align_len", "mapq")

    def __init__(self, strand, align_score, qstart, qend):
        self.strand, self.align_score, self.qstart, self.qend = strand, align_score, qstart, qend
        self.align_len = qend - qstart
        self.mapq = 60


def check_anchor_mapping(hits):
    """nanoRepeat_bam.py:165-179 on hits sorted by align_score, best first."""
    if len(hits) == 0:
        return False
    if len(hits) == 1:
        return True
    if hits[0].align_len < 10:
        return False
    return hits[0].align_score > 1.5 * hits[1].align_score and hits[0].mapq > 30


def locate_anchors(data_type, left_anchor_seq, right_anchor_seq, read_seqs):
    """-> per read (left hits, right hits), each a list of AnchorHit sorted by score, best first ('+' first on a tie)."""
    get_preset_for_minimap2(data_type)
    sc = engine.get_preset(data_type)
    n = len(read_seqs)
    if n == 0:
        return []
    rc = [rev_comp(s) for s in read_seqs]
    anchors, targets = [], []
    for fwd, rev in zip(read_seqs, rc):
        anchors += [left_anchor_seq, left_anchor_seq, right_anchor_seq, right_anchor_seq]
        targets += [fwd, rev, fwd, rev]
    recs = engine.score_tasks(anchors, targets, sc)          # the anchor is the DP's query: (tstart, tend) are read coordinates
    out = []
    for r in range(n):
        per_anchor = []
        for a in range(2):
            hits = []
            for s, strand in enumerate("+-"):
                rec = recs[4 * r + 2 * a + s]
                if rec["score"] > 0 and rec["score"] >= sc.min_dp_score:
                    hits.append(AnchorHit(strand, int(rec["score"]), int(rec["tstart"]), int(rec["tend"])))
            hits.sort(key=lambda h: h.align_score, reverse=True)           # stable: '+' first on a tie (:190-191)
            per_anchor.append(hits)
        out.append(tuple(per_anchor))
    return out


def find_anchor_locations_in_reads(data_type, repeat_region, num_cpu=1, reads=None, read_class=None):
    """Reference nanoRepeat_bam.py:260-286 + :181-234: fills repeat_region.read_dict with a Read per accepted read
    (dist_between_anchors, strand, core / mid positions, buffer lengths).
    reads: (names, sequences) or {name: sequence}; default: repeat_region.region_fq_file, which is what the reference
    hands minimap2 (:279).  read_class: the Read class to instantiate (default: this package's)."""
    if reads is None:
        names, seqs = read_fastq(repeat_region.region_fq_file)
    elif isinstance(reads, dict):
        names, seqs = list(reads), list(reads.values())
    else:
        names, seqs = reads
    if read_class is None:
        from .repeat_region import Read as read_class
    hits = locate_anchors(data_type, repeat_region.left_anchor_seq, repeat_region.right_anchor_seq, seqs)
    for name, seq, (left_hits, right_hits) in zip(names, seqs, hits):
        if not left_hits and not right_hits:
            continue                                                        # no PAF line for this read (:183)
        read = read_class()
        read.read_name = name
        read.full_read_len = len(seq)
        read.both_anchors_are_good = False
        read.left_anchor_is_good = check_anchor_mapping(left_hits)         # :195-196
        read.right_anchor_is_good = check_anchor_mapping(right_hits)
        if not read.left_anchor_is_good or not read.right_anchor_is_good:
            continue
        left, right = left_hits[0], right_hits[0]
        read.left_anchor_paf, read.right_anchor_paf = left, right
        repeat_region_length = 0
        if left.strand == right.strand:                                     # :206-207
            repeat_region_length = right.qstart - left.qend
        if repeat_region_length > -10:                                      # :209-211
            read.both_anchors_are_good = True
            read.dist_between_anchors = repeat_region_length
        if not read.both_anchors_are_good:
            continue
        repeat_region.read_dict[name] = read
        repeat_region.buffer_len = BUFFER_LEN
        read.core_seq_start_pos = left.qend - BUFFER_LEN                    # :221-234
        read.core_seq_end_pos = right.qstart + BUFFER_LEN
        read.mid_seq_start_pos = left.qend
        read.mid_seq_end_pos = right.qstart
        if read.core_seq_start_pos < 0:
            read.core_seq_start_pos = 0
        if read.core_seq_end_pos > read.full_read_len:
            read.core_seq_end_pos = read.full_read_len
        read.left_buffer_len = left.qend - read.core_seq_start_pos
        read.right_buffer_len = read.core_seq_end_pos - right.qstart
        read.strand = "+" if left.strand == "+" else "-"
    return


def make_core_seq_fastq(repeat_region, reads=None, write_files=None):
    """Reference nanoRepeat_bam.py:288-331: the core (and middle) sequence of every accepted read, in the read's
    orientation, into repeat_region.read_core_seq_dict.  The two FASTQ files the reference also writes (the unmodified
    round1_and_round2_estimation reads core_sequences.fastq) are written when the region has a temp_out_dir, or on
    request; this package's own operators read the dictionary."""
    if reads is None:
        names, seqs = read_fastq(repeat_region.region_fq_file)
    elif isinstance(reads, dict):
        names, seqs = list(reads), list(reads.values())
    else:
        names, seqs = reads
    if write_files is None:
        write_files = bool(getattr(repeat_region, "temp_out_dir", None))
    core_f = mid_f = None
    if write_files:
        repeat_region.core_seq_fq_file = os.path.join(repeat_region.temp_out_dir, "core_sequences.fastq")
        repeat_region.mid_seq_fq_file = os.path.join(repeat_region.temp_out_dir, "middle_sequences.fastq")
        repeat_region.temp_file_list += [repeat_region.core_seq_fq_file, repeat_region.mid_seq_fq_file]
        core_f, mid_f = open(repeat_region.core_seq_fq_file, "w"), open(repeat_region.mid_seq_fq_file, "w")
    for name, seq in zip(names, seqs):
        read = repeat_region.read_dict.get(name)
        if read is None:
            continue
        if read.strand == "-":
            seq = rev_comp(seq)                                             # :311-312
        core = seq[read.core_seq_start_pos:read.core_seq_end_pos]
        mid = seq[read.mid_seq_start_pos:read.mid_seq_end_pos]
        repeat_region.read_core_seq_dict[name] = core
        if core_f:
            core_f.write(f"@{name}\n{core}\n+\n{'0' * len(core)}\n")
            mid_f.write(f"@{name}\n{mid}\n+\n{'0' * len(mid)}\n")
    if core_f:
        core_f.close(); mid_f.close()
    return
