"""The drop-in boundary as the reference would use it: install() on a module of the reference's
shape, the reference's own Read / RepeatRegion classes, pickling through the result queue, fork-after-load with several
workers, and the pymm2.main-shaped shim."""
import inspect
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"


def _import_reference_module():
    """NanoRepeat.nanoRepeat_bam from /root/reference with its absent third-party imports stubbed (build container only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    return make_golden.import_reference()


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree exists in the build container only")
def test_install_matches_the_reference_module_shape():
    """install() patches exactly the names quantify1repeat_from_bam calls, with the reference's parameter lists; the
    reference's own Read / RepeatRegion carry every attribute the operators read or write."""
    import nanorepeat_b200 as nrb
    ref_bam, ref_rr = _import_reference_module()
    own = {n: getattr(ref_bam, n) for n in ("round1_and_round2_estimation", "round3_estimation")}
    try:
        nrb.install(ref_bam)
        for name, fn in own.items():
            assert getattr(ref_bam, name) is getattr(nrb, name)
            assert list(inspect.signature(fn).parameters) == list(inspect.signature(getattr(nrb, name)).parameters), name
        src = inspect.getsource(ref_bam.quantify1repeat_from_bam)
        assert "round1_and_round2_estimation(" in src and "round3_estimation(" in src      # looked up in module globals
    finally:
        for name, fn in own.items():
            setattr(ref_bam, name, fn)
    rd, rr = ref_rr.Read(), ref_rr.RepeatRegion()
    for attr in ("dist_between_anchors", "round1_repeat_size", "round2_repeat_size", "round3_repeat_size"):
        assert hasattr(rd, attr) and getattr(rd, attr) is None
    for attr in ("left_anchor_seq", "right_anchor_seq", "repeat_unit_seq", "read_dict", "read_core_seq_dict"):
        assert hasattr(rr, attr)
    pickle.dumps(rr)


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree exists in the build container only")
def test_shim_feeds_the_unmodified_reference_functions(monkeypatch, tmp_path):
    """The reference's UNMODIFIED round1_and_round2_estimation / round3_estimation on top of pymm2_shim.main, with the
    engine call inside the shim answered by the oracle (no GPU here): the shim's file handling, command recognition and
    PAF text must give the reference's functions what the golden fixtures say."""
    from conftest import load_golden
    from nanorepeat_b200 import pymm2_shim, engine
    from oracle import nr_oracle
    ref_bam, ref_rr = _import_reference_module()
    monkeypatch.setattr(engine, "get_preset", lambda dt: nr_oracle.scoring())
    monkeypatch.setattr(engine, "score_tasks", lambda q, t, sc: nr_oracle.align_batch(q, t, sc, n_threads=nr_oracle.max_threads()))
    monkeypatch.setattr(ref_bam, "pymm2", pymm2_shim)
    doc = load_golden("cfg1_small")
    reg = doc["regions"][0]
    R = ref_rr.RepeatRegion()
    R.left_anchor_seq, R.right_anchor_seq, R.repeat_unit_seq = reg["left"], reg["right"], reg["motif"]
    R.left_anchor_len, R.right_anchor_len = len(reg["left"]), len(reg["right"])
    R.temp_out_dir = str(tmp_path)
    R.core_seq_fq_file = str(tmp_path / "core_sequences.fastq")
    with open(R.core_seq_fq_file, "w") as f:
        for name, core, dist in zip(reg["read_names"], reg["cores"], reg["dists"]):
            rd = ref_rr.Read()
            rd.read_name, rd.dist_between_anchors = name, dist
            R.read_dict[name] = rd
            R.read_core_seq_dict[name] = core
            f.write(f"@{name}\n{core}\n+\n{'0' * len(core)}\n")
    ref_bam.round1_and_round2_estimation(reg["data_type"], R, 1)
    ref_bam.round3_estimation(reg["data_type"], doc["fast_mode"], R, 1)
    for name, exp in zip(reg["read_names"], reg["expected"]):
        rd = R.read_dict[name]
        assert (rd.round1_repeat_size, rd.round2_repeat_size) == (exp["r1"], exp["r2"])
        assert (None if rd.round3_repeat_size is None else float(rd.round3_repeat_size)) == exp["r3"]
    with pytest.raises(NotImplementedError):
        pymm2_shim.main("-c -t 4 -x map-ont anchors.fasta region.fastq")       # Step 1's command is not this library's


@pytest.mark.gpu
def test_installed_operators_on_reference_shaped_objects_pickle(engine):
    """install() on a module of the reference's shape; its driver calls the two names through the module globals on
    objects with the reference's attribute set; the region then pickles (it crosses a multiprocessing queue in the
    reference, nanoRepeat_bam.py:610) after EACH operator, and the results equal the direct calls."""
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth
    from helpers import refshape
    own = (refshape.round1_and_round2_estimation, refshape.round3_estimation)
    nrb.install(refshape)
    try:
        regs = synth.config1(seed=5, n_regions=3, reads_per_region=10)
        for mode in (3, 0):                        # the production ladder and the independent-rectangle checking mode
            engine.set_ladder_mode(mode)
            for reg in regs:
                rr = refshape.region_from_synth(reg)
                refshape.round1_and_round2_estimation("ont", rr, 4)
                mid = pickle.loads(pickle.dumps(rr))                 # between the two operators
                assert [rd.round2_repeat_size for rd in mid.read_dict.values()] == [rd.round2_repeat_size for rd in rr.read_dict.values()]
                refshape.round3_estimation("ont", False, rr, 4)
                back = pickle.loads(pickle.dumps(rr))
                direct = nrb.RepeatRegion.from_synth(reg)
                nrb.estimate_regions([direct], "ont", False)
                for name, rd in direct.read_dict.items():
                    o = back.read_dict[name]
                    assert (rd.round1_repeat_size, rd.round2_repeat_size, rd.round3_repeat_size) == \
                           (o.round1_repeat_size, o.round2_repeat_size, o.round3_repeat_size), (mode, name)
                assert not [k for k in vars(rr) if k.startswith("_nr")]      # nothing of ours is left on the object
                assert sum(rd.round3_repeat_size is not None for rd in rr.read_dict.values()) >= 8
    finally:
        engine.set_ladder_mode(3)
        refshape.round1_and_round2_estimation, refshape.round3_estimation = own


@pytest.mark.gpu
def test_fork_after_load_four_workers():
    """The reference forks up to 16 workers after importing everything (nanoRepeat_bam.py:719-724): the library is loaded
    before the fork, every worker initialises CUDA on its own, results come back pickled and equal the parent's."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "fork_workers.py"), "4"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["exitcodes"] == [0, 0, 0, 0] and res["regions_back"] == 12 and res["equal"], res
    assert res["reads_with_round3"] > 80


@pytest.mark.gpu
def test_pymm2_shim_on_the_gpu(engine, oracle, tmp_path):
    """The two hot-path command shapes through pymm2_shim.main on the CUDA engine: PAF text whose AS / tstart / tend are
    the oracle's, one line per template at or above minimap2's -s, best first."""
    from nanorepeat_b200 import pymm2_shim, synth
    reg = synth.config1(seed=9, n_regions=1, reads_per_region=7)[0]
    m = len(reg.repeat_unit_seq)
    tpl = reg.left_anchor_seq + reg.repeat_unit_seq * 70
    (tmp_path / "round1_ref.fasta").write_text(f">70\n{tpl}\n")
    with open(tmp_path / "core_sequences.fastq", "w") as f:
        for n, c in zip(reg.read_names, reg.core_seqs):
            f.write(f"@{n}\n{c}\n+\n{'0' * len(c)}\n")
    out, err = pymm2_shim.main(f"-c -t 4  -x map-ont  -f 0.0 {tmp_path}/round1_ref.fasta {tmp_path}/core_sequences.fastq")
    ref = oracle.align_batch(reg.core_seqs, [tpl] * len(reg.core_seqs))
    rows = [ln.split("\t") for ln in out.strip().split("\n")]
    assert [r[0] for r in rows] == reg.read_names
    for r, a in zip(rows, ref):
        assert (int(r[7]), int(r[8]), r[12]) == (int(a["tstart"]), int(a["tend"]), f"AS:i:{int(a['score'])}") and r[5] == "70"
    # round 3: one read against a ladder file
    ks = list(range(3, 12))
    with open(tmp_path / "round3_reference.3-11.fasta", "w") as f:
        for k in ks:
            f.write(f">{k}\n{reg.left_anchor_seq}{reg.repeat_unit_seq * k}{reg.right_anchor_seq}\n")
    (tmp_path / "round3_input.read0.fasta").write_text(f">{reg.read_names[0]}\n{reg.core_seqs[0]}\n")
    out, err = pymm2_shim.main(f" -x map-ont  -f 0.0 -N 100 -c --eqx -t 4 {tmp_path}/round3_reference.3-11.fasta {tmp_path}/round3_input.read0.fasta")
    ref, _off = oracle.align_ladders([reg.core_seqs[0]], reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq, [3], [11])
    rows = {int(r[5]): r for r in (ln.split("\t") for ln in out.strip().split("\n"))}
    for k, a in zip(ks, ref):
        if a["score"] >= 80:
            assert (int(rows[k][7]), int(rows[k][8]), rows[k][12]) == (int(a["tstart"]), int(a["tend"]), f"AS:i:{int(a['score'])}")
            assert int(rows[k][6]) == len(reg.left_anchor_seq) + m * k + len(reg.right_anchor_seq)
        else:
            assert k not in rows
    scores = [int(ln.split("\t")[12][5:]) for ln in out.strip().split("\n")]
    assert scores == sorted(scores, reverse=True)
