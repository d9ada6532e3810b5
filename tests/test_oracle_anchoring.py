"""oracle/anchoring.step1 (the checker of the GPU's Step 1) against what the reference's own,
unmodified find_anchor_locations_in_reads + make_core_seq_fastq decided on the same reads when its pyminimap2 was
replaced by a PAF printer over the oracle's DP (tests/golden/make_golden_anchoring.py).  No GPU."""
import json
import os

from oracle import anchoring, nr_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "anchoring_cases.json")


def test_step1_equals_the_reference_functions():
    nr_oracle.build()
    with open(GOLDEN) as f:
        doc = json.load(f)
    assert len(doc["cases"]) >= 4
    kept_total = minus = 0
    for c in doc["cases"]:
        got = anchoring.step1(c["left"], c["right"], c["names"], c["reads"], n_threads=nr_oracle.max_threads())
        assert list(got) == list(c["kept"]), (c["motif"], list(got), list(c["kept"]))
        for name, e in c["kept"].items():
            assert got[name] == e, (c["motif"], name, {k: (got[name][k], e[k]) for k in e if got[name][k] != e[k]})
            minus += e["strand"] == "-"
        kept_total += len(got)
        assert len(got) < len(c["names"])                      # some reads are rejected in every case
    assert kept_total >= 45 and minus >= 15
