"""CPU suite for the host side: the C-ABI library loads and exports everything the header declares, presets,
the ladder arithmetic of the operator layer, and (without a GPU) compute calls fail loudly instead of falling back."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from nanorepeat_b200 import engine
    L = engine.lib()
    hdr = open(os.path.join(ROOT, "include", "nanorepeat_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(engine.SYMBOLS), declared ^ set(engine.SYMBOLS)
    for name in declared:
        assert getattr(L, name) is not None


def test_presets_match_reference_table():
    import nanorepeat_b200 as nrb
    for dt in ("ont", "ont_sup", "ont_q20", "clr", "hifi"):     # reference tk.py:502-517
        assert nrb.get_preset_for_minimap2(dt) == " -x map-ont "
        sc = nrb.get_scoring(dt)
        assert (sc.match, sc.mismatch, sc.gap_open1, sc.gap_ext1, sc.gap_open2, sc.gap_ext2, sc.min_dp_score) == \
               (2, 4, 4, 2, 24, 1, 80)
    with pytest.raises(SystemExit):
        nrb.get_preset_for_minimap2("pacbio")
    with pytest.raises(ValueError):
        nrb.get_scoring("pacbio")


def test_ladder_bounds_equal_oracle_restatement():
    from nanorepeat_b200.estimation import ladder_bounds
    from oracle import selection
    rng = np.random.default_rng(0)
    vals = list(rng.uniform(-1, 3200, size=2000)) + [0.0, 15.0, 299.99999, 300.0, 3000.0, 7.666666666666667]
    for v in vals:
        for fast in (False, True):
            assert ladder_bounds(float(v), fast) == selection.ladder_bounds(float(v), fast)


def test_vectorised_ladder_bounds_equal_scalar_form():
    from nanorepeat_b200.estimation import ladder_bounds, ladder_bounds_array
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.uniform(-1, 3300, size=20000), np.arange(0, 3200, 1 / 3.0),
                           np.array([0.0, 14.999999999999998, 15.0, 299.99999999999994, 300.0, 3000.0, 3019.9999999999995])])
    for fast in (False, True):
        lo, hi = ladder_bounds_array(vals, fast)
        exp = [ladder_bounds(float(v), fast) for v in vals]
        assert lo.tolist() == [e[0] for e in exp] and hi.tolist() == [e[1] for e in exp]


def test_concat_reads_layout():
    from nanorepeat_b200 import engine
    buf, off = engine._concat(["ACG", "", "TTGA"])
    assert buf == b"ACGTTGA" and off.tolist() == [0, 3, 3, 7]
    buf, off = engine._concat([b"AC", b"G"])
    assert buf == b"ACG" and off.tolist() == [0, 2, 3]
    buf, off = engine._concat([])
    assert buf == b"" and off.tolist() == [0]
    buf, off = engine._concat(["AC\u00e9T"])
    assert buf == b"AC?T" and off.tolist() == [0, 4]


def test_empty_region_is_a_no_op():
    import nanorepeat_b200 as nrb
    rr = nrb.RepeatRegion()
    rr.left_anchor_seq = rr.right_anchor_seq = "ACGT"
    rr.repeat_unit_seq = "CAG"
    nrb.round1_and_round2_estimation("ont", rr, 4)       # reference :336 returns immediately
    nrb.round3_estimation("ont", False, rr, 4)


def test_no_cpu_fallback_without_gpu():
    """On a box without CUDA the compute calls must raise, not quietly compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu suite")
    from nanorepeat_b200 import engine
    sc = engine.get_preset("ont")
    with pytest.raises(engine.NanoRepeatB200Error) as ei:
        engine.score_tasks(["ACGT"], ["ACGT"], sc)
    assert ei.value.code == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nanorepeat_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "libnr_oracle" not in src and "nro_" not in src, f


def test_synth_is_seeded_and_acgt_only():
    from nanorepeat_b200 import synth
    a = synth.config1(seed=1, n_regions=2, reads_per_region=5)
    b = synth.config1(seed=1, n_regions=2, reads_per_region=5)
    assert [r.core_seqs for r in a] == [r.core_seqs for r in b]
    for r in a:
        assert set("".join(r.core_seqs)) <= set("ACGT")
        assert all(d > -10 for d in r.dist_between_anchors)
    r2, r3 = synth.algorithmic_cells(1000, 1000, 5, 285, 27, 2, 32)
    assert r2 == 285 * (1000 + 135)
    assert r3 == 285 * sum(2000 + 5 * k for k in range(2, 33))


def test_mean_of_tied_rungs_is_exact_as_sum_over_count():
    """estimation.round3_estimation returns float64(sum k)/n; the reference calls np.mean(list) (:431)."""
    rng = np.random.default_rng(3)
    for _ in range(2000):
        n = int(rng.integers(1, 302))
        lo = int(rng.integers(0, 3000))
        ks = sorted(rng.choice(np.arange(lo, lo + 301), size=n, replace=False).tolist())
        a = np.mean(ks)
        b = np.float64(np.int64(sum(ks))) / np.float64(np.int32(n))
        assert a == b and type(a) is type(b)
