"""Host bookkeeping of the 1-D phasing (nanorepeat_b200/phasing.py) and the checker's trim / labels / overlap rule
(oracle/gmm.py) against what the reference's own functions returned (tests/golden/make_golden_phasing.py).  No GPU."""
import json
import os

import numpy as np
import pytest

from nanorepeat_b200 import phasing
from oracle import gmm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phasing_cases.json")


@pytest.fixture(scope="module")
def doc():
    with open(GOLDEN) as f:
        return json.load(f)


def test_trim_and_labels_equal_reference(doc):
    assert len(doc["cases"]) >= 30
    for c in doc["cases"]:
        names, sizes = list(c["sizes"]), list(c["sizes"].values())
        assert [names[i] for i in gmm.trim(sizes)] == c["kept"]
        kept = [c["sizes"][n] for n in c["kept"]]
        lab, pr = gmm.labels(kept, c["weights"], c["means"], c["variances"])
        assert [int(l) for l in lab] == c["label"]
        assert np.allclose(pr, c["proba"], rtol=1e-9, atol=1e-300)


def test_allele_lists_equal_reference(doc):
    for c in doc["cases"]:
        names, sizes = list(c["sizes"]), list(c["sizes"].values())
        lab_of = dict(zip(c["kept"], zip(c["label"], c["proba"])))
        fit = dict(n=len(c["means"]), means=c["means"], variances=c["variances"], weights=c["weights"],
                   label=[lab_of[n][0] if n in lab_of else -1 for n in names], proba=[lab_of[n][1] if n in lab_of else 0.0 for n in names])
        alleles = phasing.create_allele_list_1d(fit, names, sizes)
        assert len(alleles) == len(c["alleles"])
        for a, e in zip(alleles, c["alleles"]):
            assert a.readname_list == e["reads"] and a.repeat1_size_list == e["sizes"] and a.num_reads == e["num_reads"]
            assert a.repeat1_median_size == e["median"] and a.confidence_list == e["confidence"]
            assert a.gmm_mean1 == e["mean"] and a.gmm_sd1 == e["sd"] and a.gmm_min1 == e["gmm_min"] and a.gmm_max1 == e["gmm_max"]
            assert np.allclose(a.probability_list, e["proba"], rtol=1e-12, atol=0)
        after, removed = phasing.remove_noisy_reads_1d(list(alleles), c["ploidy"])
        assert [a.readname_list for a in after] == c["after_noise_removal"] and removed == c["removed"]


def test_overlap_rule_and_error_rate_quirk(doc):
    for o in doc["overlaps"]:
        assert gmm.interval_has_overlap(o["a"], o["b"]) == o["expected"]
    # nanoRepeat_bam.py:692 `data_type == 'ont' or 'clr'` is always true
    assert {phasing.error_rate_of(t) for t in ("ont", "ont_sup", "ont_q20", "clr", "hifi")} == {0.07}
    assert phasing.error_rate_of("hifi", as_written=False) == 0.02
    with pytest.raises(ValueError):
        phasing.error_rate_of("nanopore")
