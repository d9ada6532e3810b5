"""Host-side mirror of the reference's hot-path operators, backed by the CUDA library.

    round1_and_round2_estimation   <- reference src/NanoRepeat/nanoRepeat_bam.py:334-393
    round3_estimation              <- reference src/NanoRepeat/nanoRepeat_bam.py:446-450 (+ :452-500, :408-434)

Same names, arguments, attribute side effects (Read.round{1,2,3}_repeat_size) and degenerate-input behaviour;
`num_cpu` is accepted and ignored (it was minimap2's -t).  What changed: one region-batched call into
libnanorepeat_b200.so per round instead of temp files + one minimap2 process per read + PAF text.
All float arithmetic that decides results (r1, T, r2, ladder bounds with int() truncation, np.mean of the tied
rungs) stays here in Python/numpy float64, written exactly as the reference writes it.
"""
import numpy as np

from . import engine
from .presets import get_preset_for_minimap2


def _scoring_for(data_type):
    get_preset_for_minimap2(data_type)        # same unknown-type behaviour as tk.py:514-516 (exit 1)
    return engine.get_preset(data_type)


def round1_and_round2_estimation(data_type, repeat_region, num_cpu=1):
    """Rounds 1 and 2 for one region (reference nanoRepeat_bam.py:334-393)."""
    if len(repeat_region.read_dict) == 0:
        return                                                                  # :336
    sc = _scoring_for(data_type)
    motif = repeat_region.repeat_unit_seq
    left = repeat_region.left_anchor_seq
    names = list(repeat_region.read_dict)

    round1_repeat_size_list = []
    for read_name in names:                                                     # :339-342
        read = repeat_region.read_dict[read_name]
        read.round1_repeat_size = float(read.dist_between_anchors) / len(motif)
        round1_repeat_size_list.append(read.round1_repeat_size)

    template_repeat_size = int(max(round1_repeat_size_list) * 1.5) + 1           # :344
    if template_repeat_size < max(round1_repeat_size_list) + 10:                # :346-347
        template_repeat_size = int(max(round1_repeat_size_list) + 10)

    # one engine call for the region (was pymm2.main at :362); reads come from read_core_seq_dict, which is what
    # the reference wrote to core_sequences.fastq (:311-321)
    qnames = [n for n in names if n in repeat_region.read_core_seq_dict]
    cores = [repeat_region.read_core_seq_dict[n].strip() for n in qnames]
    alns = engine.round2_region(sc, left, motif, template_repeat_size, cores)

    n_left = len(left)
    min_score = max(1, sc.min_dp_score)
    for read_name, a in zip(qnames, alns):
        score, tstart, tend = int(a["score"]), int(a["tstart"]), int(a["tend"])
        if score < min_score:
            continue                                                            # minimap2 prints no line
        if tstart <= n_left and tend >= n_left:                                 # :373
            repeat_region.read_dict[read_name].round2_repeat_size = float(tend - n_left) / len(motif)   # :375
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


def round3_estimation(data_type, fast_mode, repeat_region, num_cpu=1):
    """Round 3 for one region (reference nanoRepeat_bam.py:446-450)."""
    sc = _scoring_for(data_type)
    names, cores, kmin, kmax = [], [], [], []
    for read_name in repeat_region.read_dict:                                   # :457-472
        read = repeat_region.read_dict[read_name]
        if read.round2_repeat_size is None:
            continue
        lo, hi = ladder_bounds(read.round2_repeat_size, fast_mode)
        names.append(read_name)
        cores.append(repeat_region.read_core_seq_dict[read_name].strip())
        kmin.append(lo)
        kmax.append(hi)
    if not names:
        return
    sum_k, n_k, top = engine.round3_region(
        sc, repeat_region.left_anchor_seq, repeat_region.right_anchor_seq, repeat_region.repeat_unit_seq,
        cores, np.asarray(kmin, dtype=np.int32), np.asarray(kmax, dtype=np.int32))
    # np.mean(list of k) == float64(sum k) / n exactly: the k are small integers, so every partial sum is an
    # exactly representable float64 and numpy's pairwise summation cannot round (tests/test_host.py checks it)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_k = sum_k.astype(np.float64) / n_k.astype(np.float64)
    for i, read_name in enumerate(names):
        read = repeat_region.read_dict[read_name]
        if top[i] <= 0:
            continue                                                            # no PAF line at all (:421)
        if n_k[i] > 0:
            read.round3_repeat_size = mean_k[i]                                 # :431
        else:
            read.round3_repeat_size = read.round2_repeat_size                   # :433
    return


def estimate_regions(regions, data_type=None, fast_mode=False):
    """Convenience driver: rounds 1-3 over a list of RepeatRegion-like objects (the reference runs this per
    region inside quantify1repeat_from_bam, nanoRepeat_bam.py:675-679)."""
    for rr in regions:
        dt = data_type or getattr(rr, "data_type", "ont")
        round1_and_round2_estimation(dt, rr, 1)
        round3_estimation(dt, fast_mode, rr, 1)
    return regions


def install(nanoRepeat_bam_module):
    """Patch the reference module in place so the unmodified CLI runs this path:
        import NanoRepeat.nanoRepeat_bam as m; nanorepeat_b200.install(m)
    """
    nanoRepeat_bam_module.round1_and_round2_estimation = round1_and_round2_estimation
    nanoRepeat_bam_module.round3_estimation = round3_estimation
    return nanoRepeat_bam_module
