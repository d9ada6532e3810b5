"""The Python side of estimate_regions (nanorepeat_b200/estimation.py): gathering a region list into chunks, running the
chunks on worker threads, assigning the returned arrays to the Read objects -- with the library call replaced by the CPU
oracle (tests may use it as the checker), so that order, chunk boundaries, reads without a core, empty regions and
errors are exercised without a GPU."""
import numpy as np
import pytest

import nanorepeat_b200 as nrb
from nanorepeat_b200 import engine, estimation, synth


def _fake_engine_estimate(calls):
    from oracle import selection

    def fake(sc, fast_mode, lefts, rights, motifs, cores, dists, max_dists=None, on_ready=None, n_reads=None):
        assert n_reads is not None and sum(n_reads) == len(cores) == len(dists)
        if on_ready is not None:
            on_ready()
        calls.append(list(n_reads))
        r1, r2, r3, ok2, st3 = [], [], [], [], []
        pos = 0
        for g, n in enumerate(n_reads):
            cs, ds = cores[pos:pos + n], dists[pos:pos + n]
            pos += n
            if any(c == "BOOM" for c in cs):
                raise engine.NanoRepeatB200Error(-1, "injected failure")
            res = selection.estimate_region(lefts[g], rights[g], motifs[g], cs, ds, fast_mode=fast_mode, n_threads=2,
                                            max_dist=None if max_dists is None else max_dists[g])
            m = len(motifs[g])
            for i in range(n):
                r1.append(float(ds[i]) / m)
                v2, v3 = res["r2"][i], res["r3"][i]
                ok2.append(v2 is not None)
                r2.append(0.0 if v2 is None else float(v2))
                if v3 is None:
                    st3.append(0); r3.append(0.0)
                elif res["round3_idx"][i] is None:           # fell back to r2
                    st3.append(2); r3.append(float(v3))
                else:
                    st3.append(1); r3.append(float(v3))
        return dict(r1=np.array(r1), r2=np.array(r2), r2_valid=np.array(ok2, bool), r3=np.array(r3),
                    r3_state=np.array(st3, np.uint8), T=np.zeros(len(n_reads), np.int32), stats={})
    return fake


def _regions():
    regs = synth.config1(seed=8, n_regions=7, reads_per_region=6) + synth.config2(seed=9, n_reads=8)
    rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
    rrs.insert(3, nrb.RepeatRegion())                                   # a region without reads (reference :336)
    rrs[3].left_anchor_seq, rrs[3].right_anchor_seq, rrs[3].repeat_unit_seq = "ACGT" * 5, "TTGA" * 5, "CAG"
    name = next(iter(rrs[1].read_dict))
    del rrs[1].read_core_seq_dict[name]                                 # a read that never made it into the core FASTQ
    return rrs, name


def _sizes(rrs):
    return [[(n, rd.round1_repeat_size, rd.round2_repeat_size, None if rd.round3_repeat_size is None else float(rd.round3_repeat_size))
             for n, rd in rr.read_dict.items()] for rr in rrs]


@pytest.mark.parametrize("chunks", ["1", "2", "3", "5"])
def test_chunked_threaded_path_equals_one_call(monkeypatch, chunks):
    monkeypatch.setattr(engine, "ladder_mode", lambda: 3)
    monkeypatch.setattr(estimation, "_scoring_for", lambda dt: None)
    calls = []
    monkeypatch.setattr(engine, "estimate_regions", _fake_engine_estimate(calls))
    monkeypatch.setenv("NR_PY_CHUNKS", "1")
    one, name = _regions()
    nrb.estimate_regions(one, "ont", False)
    assert len(calls) == 1
    calls.clear()
    monkeypatch.setenv("NR_PY_CHUNKS", chunks)
    many, _ = _regions()
    nrb.estimate_regions(many, "ont", False)
    assert len(calls) == min(int(chunks), 5) or int(chunks) == 1
    assert sum(sum(c) for c in calls) == sum(len(rr.read_dict) for rr in many) - 1
    assert _sizes(many) == _sizes(one)
    rd = many[1].read_dict[name]                                         # round 1 only, like the reference
    assert rd.round1_repeat_size is not None and rd.round2_repeat_size is None and rd.round3_repeat_size is None
    assert sum(s[3] is not None for reg in _sizes(many) for s in reg) > 30


def test_a_failing_chunk_raises_and_leaves_no_thread_waiting(monkeypatch):
    monkeypatch.setattr(engine, "ladder_mode", lambda: 3)
    monkeypatch.setattr(estimation, "_scoring_for", lambda dt: None)
    monkeypatch.setattr(engine, "estimate_regions", _fake_engine_estimate([]))
    monkeypatch.setenv("NR_PY_CHUNKS", "3")
    rrs, _ = _regions()
    victim = rrs[5]
    victim.read_core_seq_dict[next(iter(victim.read_core_seq_dict))] = "BOOM"
    with pytest.raises(engine.NanoRepeatB200Error):
        nrb.estimate_regions(rrs, "ont", False)
    good, _ = _regions()                                                  # the pool is still usable afterwards
    nrb.estimate_regions(good, "ont", False)
    assert sum(s[3] is not None for reg in _sizes(good) for s in reg) > 30
