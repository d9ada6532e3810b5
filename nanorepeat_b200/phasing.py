"""Host-side mirror of the reference's 1-D allele phasing (Step 4 of quantify1repeat_from_bam, reference
src/NanoRepeat/nanoRepeat_bam.py:502-575 and split_alleles.py:52-293), backed by the CUDA library's batched mixture
fits (nr_phase_1d, csrc/nr_gmm.cu) -- for MANY regions per call instead of one sklearn fit sequence per region.

    Allele                       <- split_alleles.py:52-72 (the 1-D fields)
    create_allele_list_1d        <- split_alleles.py:258-293 (labels -> alleles, medians, 2-sd bounds, HIGH / LOW)
    remove_noisy_reads_1d        <- nanoRepeat_bam.py:502-514
    phase_regions_1d             <- split_allele_using_gmm_1d (:515-575) up to the sorted allele list, all regions at once
    split_allele_using_gmm_1d    <- the same for one RepeatRegion-like object (sets results.num_alleles when present)

Trim, bootstrap, the mixture fits and the labels happen in the library; what stays here is the reference's bookkeeping
on the labels.  Files (phased_reads.txt, summary, FASTQ per allele, plots: :380-500) stay with the caller.  The library
draws its random numbers from a seeded counter-based generator (the reference's are unseeded): same `seed`, same result.
"""
import math

import numpy as np

from . import engine

def error_rate_of(data_type, as_written=True):
    """nanoRepeat_bam.py:692-701.  The first test there reads `data_type == 'ont' or 'clr'`, which is always true: the
    reference phases EVERY data type with 0.07.  as_written=True keeps that (results depend on it); False gives the
    table the branches spell out (ont / clr 0.07, ont_sup 0.04, ont_q20 0.03, hifi 0.02)."""
    table = {"ont": 0.07, "clr": 0.07, "ont_sup": 0.04, "ont_q20": 0.03, "hifi": 0.02}
    if data_type not in table:
        raise ValueError(f"unknown data type: {data_type}")
    return 0.07 if as_written else table[data_type]


class Allele:
    def __init__(self):
        self.gmm_mean1 = None
        self.gmm_sd1 = None
        self.readname_list = []
        self.repeat1_size_list = []
        self.repeat1_median_size = None
        self.probability_list = []
        self.num_reads = None
        self.confidence_list = []
        self.gmm_min1 = None
        self.gmm_max1 = None


def create_allele_list_1d(fit, readnames, sizes, probability_cutoff=0.95):
    """split_alleles.py:258-293.  fit: one region's dict from engine.phase_1d; readnames / sizes: all reads of the region in
    the order given to it (label -1 = trimmed as an outlier, as remove_outlier_reads_1d drops them)."""
    alleles = []
    for j in range(fit["n"]):
        a = Allele()
        a.gmm_mean1 = float(fit["means"][j])
        a.gmm_sd1 = math.sqrt(float(fit["variances"][j]))
        alleles.append(a)
    for name, size, lab, pr in zip(readnames, sizes, fit["label"], fit["proba"]):
        if lab < 0:
            continue
        a = alleles[lab]
        a.readname_list.append(name)
        a.repeat1_size_list.append(size)
        a.probability_list.append(float(pr))
    for a in alleles:
        a.num_reads = len(a.readname_list)
        if a.num_reads == 0:
            a.repeat1_median_size = 0
            a.gmm_min1 = 0
            continue
        a.repeat1_median_size = int(np.median(a.repeat1_size_list) + 0.5)
        a.gmm_min1 = a.gmm_mean1 - 2 * a.gmm_sd1
        a.gmm_max1 = a.gmm_mean1 + 2 * a.gmm_sd1
        a.confidence_list = ["LOW" if (p < probability_cutoff or s < a.gmm_min1 or s > a.gmm_max1) else "HIGH"
                             for p, s in zip(a.probability_list, a.repeat1_size_list)]
    alleles.sort(key=lambda a: a.num_reads)
    while alleles and alleles[0].num_reads == 0:
        alleles.pop(0)
    return alleles


def remove_noisy_reads_1d(allele_list, ploidy):
    """nanoRepeat_bam.py:502-514"""
    allele_list.sort(key=lambda a: a.num_reads)
    removed = 0
    while len(allele_list) > ploidy and len(allele_list) >= 2:
        if allele_list[0].num_reads * 1.5 <= allele_list[-ploidy].num_reads:
            removed += allele_list[0].num_reads
            allele_list.pop(0)
        else:
            break
    return allele_list, removed


def phase_regions_1d(read_size_dicts, ploidy=2, error_rate=0.07, max_mutual_overlap=0.15, max_num_components=-1,
                     remove_noisy_reads=False, seed=0, region_id_base=0):
    """read_size_dicts: per region {read name: round-3 size}.  -> per region (allele_list sorted by gmm_mean1,
    num_removed_reads), or None where the reference phases nothing (fewer than two reads, :533-539).
    max_num_components beyond the library's 32 (NR_GMM_MAX_COMPONENTS) is cut to 32: the default is ploidy + 20."""
    if ploidy < 1:
        raise ValueError("ploidy must be >= 1")
    if max_num_components == -1:
        max_num_components = ploidy + 20                                      # nanoRepeat.py:159-160
    params = engine.GmmParams(error_rate=error_rate, max_mutual_overlap=max_mutual_overlap,
                              max_components=min(max_num_components, engine.GMM_MAX_COMPONENTS), seed=seed)
    names = [list(d.keys()) for d in read_size_dicts]
    sizes = [[float(d[k]) for k in ks] for d, ks in zip(read_size_dicts, names)]
    fits = engine.phase_1d(params, sizes, region_id_base)
    out = []
    for fit, ks, xs in zip(fits, names, sizes):
        if fit["n"] == 0:
            out.append(None)
            continue
        alleles = create_allele_list_1d(fit, ks, xs)
        removed = 0
        if remove_noisy_reads and len(alleles) > ploidy:
            alleles, removed = remove_noisy_reads_1d(alleles, ploidy)
        alleles.sort(key=lambda a: a.gmm_mean1)
        out.append((alleles, removed))
    return out


def split_allele_using_gmm_1d(repeat_region, ploidy, error_rate, max_mutual_overlap, max_num_components, remove_noisy_reads, seed=0):
    """nanoRepeat_bam.py:515-575 for one region, without the file outputs: -> (allele_list, num_removed_reads) or None."""
    sizes = {name: read.round3_repeat_size for name, read in repeat_region.read_dict.items() if read.round3_repeat_size is not None}
    res = phase_regions_1d([sizes], ploidy, error_rate, max_mutual_overlap, max_num_components, remove_noisy_reads, seed)[0]
    if res is not None and getattr(repeat_region, "results", None) is not None:
        repeat_region.results.num_alleles = len(res[0])
    return res
