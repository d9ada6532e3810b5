"""Seeded synthetic workloads shaped like BASELINE.json's five configs.

The generator emits exactly what Step 1 of the reference pipeline hands to the hot path
(reference src/NanoRepeat/nanoRepeat_bam.py:205-234 and :308-316): per read the oriented core sequence
(100 read bases of left flank + repeat + 100 read bases of right flank) and `dist_between_anchors`,
plus the region's 1000-bp left/right anchors and repeat unit.  ACGT only (reference tk.py:346-355 cannot
reverse-complement anything else).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)

# per-base (substitution, insertion, deletion) rates and per-unit slip probability.  These are the
# generator's own parameters; the reference only names 0.02-0.07 totals (nanoRepeat_bam.py:694-701).
ERROR_PROFILES = {
    "hifi":    (0.001, 0.002, 0.002, 0.003),
    "ont_q20": (0.010, 0.010, 0.010, 0.010),
    "ont_sup": (0.015, 0.015, 0.015, 0.015),
    "ont":     (0.020, 0.020, 0.020, 0.020),
    "ont_r9":  (0.030, 0.030, 0.030, 0.030),
    "clr":     (0.020, 0.070, 0.030, 0.030),
}


@dataclass
class SynthRegion:
    name: str
    left_anchor_seq: str
    right_anchor_seq: str
    repeat_unit_seq: str
    data_type: str = "ont"
    read_names: List[str] = field(default_factory=list)
    core_seqs: List[str] = field(default_factory=list)
    dist_between_anchors: List[int] = field(default_factory=list)
    true_sizes: List[int] = field(default_factory=list)


def random_seq(rng, n):
    return _BASES[rng.integers(0, 4, size=n)].tobytes().decode()


def mutate(rng, seq, sub, ins, dele):
    """Apply i.i.d. substitutions / insertions / deletions; returns the mutated string."""
    a = np.frombuffer(seq.encode(), dtype=np.uint8)
    n = a.size
    if n == 0:
        return ""
    u = rng.random(n)
    keep = u >= dele
    is_sub = (u >= dele) & (u < dele + sub)
    out = a.copy()
    if is_sub.any():
        # replace by a different base: rotate within ACGT by 1..3
        idx = np.searchsorted(_BASES, out[is_sub], sorter=np.argsort(_BASES))
        order = np.argsort(_BASES)
        cur = order[idx]
        out[is_sub] = _BASES[(cur + rng.integers(1, 4, size=cur.size)) % 4]
    n_ins = rng.random(n) < ins
    # build output: optional inserted base before each kept/deleted position
    pieces = np.empty(2 * n, dtype=np.uint8)
    mask = np.zeros(2 * n, dtype=bool)
    pieces[0::2] = _BASES[rng.integers(0, 4, size=n)]
    mask[0::2] = n_ins
    pieces[1::2] = out
    mask[1::2] = keep
    return pieces[mask].tobytes().decode()


def simulate_core(rng, left, right, motif, k_true, profile, buffer_len=100):
    """One read's (core_seq, dist_between_anchors, k_observed)."""
    sub, ins, dele, slip = ERROR_PROFILES[profile]
    k_obs = int(k_true)
    if slip > 0 and k_true > 0:
        n_slip = rng.binomial(k_true, slip)
        if n_slip:
            k_obs = max(0, k_true + int(rng.choice([-1, 1], size=n_slip).sum()))
    pad = buffer_len + 40
    lf = mutate(rng, left[-pad:], sub, ins, dele)[-buffer_len:]
    rf = mutate(rng, right[:pad], sub, ins, dele)[:buffer_len]
    mid = mutate(rng, motif * k_obs, sub, ins, dele)
    return lf + mid + rf, len(mid), k_obs


def make_region(rng, name, motif, allele_sizes, n_reads, profile, data_type=None, flank=1000,
                allele_weights=None):
    reg = SynthRegion(name=name, left_anchor_seq=random_seq(rng, flank), right_anchor_seq=random_seq(rng, flank),
                      repeat_unit_seq=motif, data_type=data_type or ("ont" if profile == "ont_r9" else profile))
    allele_sizes = list(allele_sizes)
    which = rng.choice(len(allele_sizes), size=n_reads, p=allele_weights)
    for i in range(n_reads):
        k = int(allele_sizes[which[i]])
        core, dist, _ = simulate_core(rng, reg.left_anchor_seq, reg.right_anchor_seq, motif, k, profile)
        reg.read_names.append(f"{name}_read{i}")
        reg.core_seqs.append(core)
        reg.dist_between_anchors.append(dist)
        reg.true_sizes.append(k)
    return reg


def random_motif(rng, m):
    while True:
        s = random_seq(rng, m)
        if len(set(s)) > 1:          # not a homopolymer
            return s


README_MOTIFS = ["TTTAG", "TATTG", "TTCC", "AAAG", "GTTTT"]   # reference README.md:83-88


def config1(seed=1, n_regions=15, reads_per_region=30):
    """15 chr1-style STR regions, 4-5 bp motifs, 30 ont_q20 reads each."""
    rng = np.random.default_rng(seed)
    regs = []
    for r in range(n_regions):
        motif = README_MOTIFS[r] if r < len(README_MOTIFS) else random_motif(rng, int(rng.integers(4, 6)))
        k1 = int(rng.integers(6, 48))
        k2 = int(rng.integers(k1 + 3, 51))
        regs.append(make_region(rng, f"cfg1_r{r}", motif, [k1, k2], reads_per_region, "ont_q20"))
    return regs


def config2(seed=2, n_reads=5000):
    """HTT amplicon: (CAG)n CAACAGCCGCCA (CCG)m as the two BED rows of example_data/HTT_repeat_region.bed,
    ONT reads, CAG alleles 17 and 55 plus a 1% tail up to 150, CCG alleles 7 and 10."""
    rng = np.random.default_rng(seed)
    flank_l, flank_r = random_seq(rng, 1000), random_seq(rng, 1000)
    mid = "CAACAGCCGCCA"
    regs = [SynthRegion("cfg2_HTT_CAG", flank_l, "", "CAG", "ont"), SynthRegion("cfg2_HTT_CCG", "", flank_r, "CCG", "ont")]
    sub, ins, dele, _ = ERROR_PROFILES["ont"]
    # the reference genome carries some allele; anchors are reference sequence around each BED row
    ref_cag, ref_ccg = 19, 7
    regs[0].right_anchor_seq = (mid + "CCG" * ref_ccg + flank_r)[:1000]
    regs[1].left_anchor_seq = (flank_l + "CAG" * ref_cag + mid)[-1000:]
    for i in range(n_reads):
        u = rng.random()
        if u < 0.01:
            cag = int(rng.integers(56, 151))
            ccg = 7
        elif u < 0.5:
            cag, ccg = 17, 10
        else:
            cag, ccg = 55, 7
        for reg, k in ((regs[0], cag), (regs[1], ccg)):
            if reg is regs[0]:
                left, right = flank_l, (mid + "CCG" * ccg + flank_r)
            else:
                left, right = (flank_l + "CAG" * cag + mid), flank_r
            core, dist, _ = simulate_core(rng, left, right, reg.repeat_unit_seq, k, "ont")
            reg.read_names.append(f"htt_read{i}")
            reg.core_seqs.append(core)
            reg.dist_between_anchors.append(dist)
            reg.true_sizes.append(k)
    return regs


def config3(seed=3, n_loci=100000, reads_per_locus=30):
    """Genome-wide catalog: motif length U{2..6}, k in [5,40], 30 HiFi reads per locus."""
    rng = np.random.default_rng(seed)
    regs = []
    for r in range(n_loci):
        motif = random_motif(rng, int(rng.integers(2, 7)))
        k1 = int(rng.integers(5, 41))
        k2 = int(rng.integers(5, 41))
        regs.append(make_region(rng, f"cfg3_l{r}", motif, [k1, k2], reads_per_locus, "hifi"))
    return regs


def config4(seed=4, reads_per_locus=200, scale=1.0):
    """Pathogenic expansions: C9orf72 GGGGCC ~1000 units, FMR1 CGG ~500, ONT R9 reads, plus a normal allele."""
    rng = np.random.default_rng(seed)
    regs = []
    for name, motif, big, spread, normal in (("cfg4_C9orf72", "GGGGCC", 1000, 50, 8), ("cfg4_FMR1", "CGG", 500, 25, 30)):
        big = max(1, int(big * scale)); spread = max(1, int(spread * scale))
        sizes = [normal] + [int(x) for x in rng.integers(big - spread, big + spread + 1, size=8)]
        w = [0.5] + [0.5 / 8] * 8
        regs.append(make_region(rng, name, motif, sizes, reads_per_locus, "ont_r9", data_type="ont", allele_weights=w))
    return regs


def config5(seed=5, n_reads=1000000, reads_per_region=50, k_max=2000):
    """Full-box sweep: motif length U{2..6}, k_true log-uniform on [1, k_max], half ont / half clr."""
    rng = np.random.default_rng(seed)
    regs = []
    n_regions = max(1, n_reads // reads_per_region)
    for r in range(n_regions):
        motif = random_motif(rng, int(rng.integers(2, 7)))
        k = int(np.exp(rng.uniform(0.0, np.log(k_max))))
        prof = "ont" if r % 2 == 0 else "clr"
        regs.append(make_region(rng, f"cfg5_r{r}", motif, [k], reads_per_region, prof))
    return regs


def algorithmic_cells(n_left, n_right, m, core_len, T, kmin, kmax):
    """Full-rectangle cells the reference hands its aligner for one read:
    round 2: q * (|L| + m*T); round 3: sum_k q * (|L| + m*k + |R|)."""
    n = kmax - kmin + 1
    r2 = core_len * (n_left + m * T)
    r3 = core_len * (n * (n_left + n_right) + m * (kmin + kmax) * n // 2) if n > 0 else 0
    return r2, r3


def region_reads(seed=6, n_reads=40, motif="CAG", alleles=(17, 55), profile="ont", flank=1000, outer=2000):
    """What Step 1 of the reference starts from (nanoRepeat_bam.py:577-600): the raw reads overlapping one region, in
    either orientation, some of them ending inside an anchor.  -> (SynthRegion without cores, names, sequences, truth):
    truth[name] = (allele, strand) for reads that contain both anchors in full, None for truncated ones."""
    rng = np.random.default_rng(seed)
    sub, ins, dele, _ = ERROR_PROFILES[profile]
    up, down = random_seq(rng, outer), random_seq(rng, outer)
    reg = SynthRegion(f"reads_{motif}", random_seq(rng, flank), random_seq(rng, flank), motif, "ont")
    comp = str.maketrans("ACGT", "TGCA")
    names, seqs, truth = [], [], {}
    for i in range(n_reads):
        k = int(alleles[int(rng.integers(0, len(alleles)))])
        genome = up + reg.left_anchor_seq + motif * k + reg.right_anchor_seq + down
        lo = int(rng.integers(0, outer))
        hi = len(genome) - int(rng.integers(0, outer))
        full = True
        if i % 7 == 3:                                   # starts inside the left anchor: no left hit worth the name
            lo = outer + flank - int(rng.integers(5, 30)); full = False
        if i % 11 == 5:                                  # ends in the repeat: no right anchor at all
            hi = outer + flank + len(motif) * k // 2; full = False
        read = mutate(rng, genome[lo:hi], sub, ins, dele)
        strand = "+" if rng.random() < 0.5 else "-"
        if strand == "-":
            read = read.translate(comp)[::-1]
        name = f"raw{i}"
        names.append(name); seqs.append(read)
        truth[name] = (k, strand) if full else None
    return reg, names, seqs, truth


def joint_locus(seed=7, n_reads=500, profile="ont", flank=1000, alleles=((17, 10), (55, 7)), amplicon_flank=(150, 600)):
    """nanoRepeat-joint's input shape (reference README.md:167, nanoRepeat_joint.py:509-649): the HTT locus
    (CAG)k1 CAACAGCCGCCA (CCG)k2 with 1000-bp anchors, raw amplicon reads of either strand, and per read the
    (min, max) repeat-count ranges the initial estimate hands to the grid rounds (here: the simulated counts +- what the
    reference's initial estimate typically leaves, :592-649).  -> dict(left, mid, right, motif1, motif2, reads, range1,
    range2, truth)."""
    rng = np.random.default_rng(seed)
    sub, ins, dele, _ = ERROR_PROFILES[profile]
    left, right, mid = random_seq(rng, flank), random_seq(rng, flank), "CAACAGCCGCCA"
    comp = str.maketrans("ACGT", "TGCA")
    reads, range1, range2, truth = [], [], [], []
    for i in range(n_reads):
        k1, k2 = alleles[int(rng.integers(0, len(alleles)))]
        lo, hi = int(rng.integers(*amplicon_flank)), int(rng.integers(*amplicon_flank))
        read = mutate(rng, left[-lo:] + "CAG" * k1 + mid + "CCG" * k2 + right[:hi], sub, ins, dele)
        if rng.random() < 0.5:
            read = read.translate(comp)[::-1]
        reads.append(read)
        truth.append((k1, k2))
        range1.append((max(0, k1 - int(rng.integers(8, 16))), k1 + int(rng.integers(8, 16))))
        range2.append((max(0, k2 - int(rng.integers(4, 7))), k2 + int(rng.integers(4, 7))))
    return dict(left=left, mid=mid, right=right, motif1="CAG", motif2="CCG", reads=reads, range1=range1, range2=range2,
                truth=truth)
