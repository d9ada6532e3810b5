"""Step 1 on the GPU (anchor finding + core extraction, nanoRepeat_bam.py:165-331) against the CPU
restatement, and the whole per-region chain raw reads -> anchors -> cores -> rounds 1-3 against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_step1_equals_oracle_restatement(engine, oracle):
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth, anchoring
    from oracle import anchoring as oanchor
    for seed, motif, alleles in ((6, "CAG", (17, 55)), (7, "GGGGCC", (8, 300)), (8, "AT", (12,))):
        reg, names, seqs, truth = synth.region_reads(seed=seed, n_reads=33, motif=motif, alleles=alleles)
        rr = nrb.RepeatRegion()
        rr.left_anchor_seq, rr.right_anchor_seq, rr.repeat_unit_seq = reg.left_anchor_seq, reg.right_anchor_seq, motif
        anchoring.find_anchor_locations_in_reads("ont", rr, 1, reads=(names, seqs))
        anchoring.make_core_seq_fastq(rr, reads=(names, seqs), write_files=False)
        exp = oanchor.step1(reg.left_anchor_seq, reg.right_anchor_seq, names, seqs, n_threads=oracle.max_threads())
        assert set(rr.read_dict) == set(exp), (seed, set(rr.read_dict) ^ set(exp))
        for name, e in exp.items():
            rd = rr.read_dict[name]
            assert (rd.strand, rd.dist_between_anchors, rd.core_seq_start_pos, rd.core_seq_end_pos, rd.mid_seq_start_pos,
                    rd.mid_seq_end_pos, rd.left_buffer_len, rd.right_buffer_len) == \
                   (e["strand"], e["dist"], e["core_start"], e["core_end"], e["mid_start"], e["mid_end"], e["left_buffer"],
                    e["right_buffer"]), name
            assert rr.read_core_seq_dict[name] == e["core"]
        # every read that holds both anchors in full is kept, on its strand; truncated ones are dropped
        full = {n for n, t in truth.items() if t is not None}
        assert full <= set(rr.read_dict) and all(rr.read_dict[n].strand == truth[n][1] for n in full)
        assert not (set(rr.read_dict) - full)
        for n in full:
            k = truth[n][0]
            assert abs(rr.read_dict[n].dist_between_anchors - len(motif) * k) <= max(12, 0.15 * len(motif) * k), (n, k)


def test_raw_reads_to_repeat_sizes(engine, oracle, tmp_path):
    """The per-region chain of quantify1repeat_from_bam (nanoRepeat_bam.py:669-679) from the region FASTQ on: Step 1 on
    the GPU (reading region_fq_file and writing core_sequences.fastq like the reference), then the two operators; the
    sizes equal the oracle's rounds 1-3 on the same cores, and land on the simulated alleles."""
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth, anchoring
    from oracle import selection
    reg, names, seqs, truth = synth.region_reads(seed=9, n_reads=30, motif="CAG", alleles=(17, 55))
    rr = nrb.RepeatRegion()
    rr.left_anchor_seq, rr.right_anchor_seq, rr.repeat_unit_seq = reg.left_anchor_seq, reg.right_anchor_seq, "CAG"
    rr.temp_out_dir = str(tmp_path)
    rr.region_fq_file = str(tmp_path / "region.fastq")
    with open(rr.region_fq_file, "w") as f:
        for n, s in zip(names, seqs):
            f.write(f"@{n}\n{s}\n+\n{'0' * len(s)}\n")
    anchoring.find_anchor_locations_in_reads("ont", rr, 4)
    anchoring.make_core_seq_fastq(rr)
    assert anchoring.read_fastq(rr.core_seq_fq_file)[0] == [n for n in names if n in rr.read_dict]
    nrb.round1_and_round2_estimation("ont", rr, 4)
    nrb.round3_estimation("ont", False, rr, 4)
    kept = list(rr.read_dict)
    exp = selection.estimate_region(rr.left_anchor_seq, rr.right_anchor_seq, "CAG", [rr.read_core_seq_dict[n] for n in kept],
                                    [rr.read_dict[n].dist_between_anchors for n in kept], n_threads=oracle.max_threads())
    for i, n in enumerate(kept):
        rd = rr.read_dict[n]
        g3, e3 = rd.round3_repeat_size, exp["r3"][i]
        assert rd.round2_repeat_size == exp["r2"][i] and (None if g3 is None else float(g3)) == (None if e3 is None else float(e3)), n
    ok = sum(abs(float(rr.read_dict[n].round3_repeat_size) - truth[n][0]) <= 2 for n in kept if rr.read_dict[n].round3_repeat_size is not None)
    assert len(kept) >= 20 and ok >= len(kept) - 3
