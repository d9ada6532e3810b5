"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sharded driver, with the per-rank engine replaced by the
CPU oracle (tests may use it as the checker), so partitioning, region splitting, T pinning and the host-side
gather are exercised without a GPU."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _oracle_estimate(regions, data_type=None, fast_mode=False):
    """Same contract as nanorepeat_b200.estimate_regions, computed by the oracle."""
    from oracle import selection
    for rr in regions:
        names = [n for n in rr.read_dict if n in rr.read_core_seq_dict]
        if not names:
            continue
        m = len(rr.repeat_unit_seq)
        dists = [rr.read_dict[n].dist_between_anchors for n in names]
        # honour the region-wide T of a split region by feeding the pinned maximum through round 1
        extra = getattr(rr, "round1_max_dist", None)
        res = selection.estimate_region(rr.left_anchor_seq, rr.right_anchor_seq, rr.repeat_unit_seq,
                                        [rr.read_core_seq_dict[n] for n in names], dists, fast_mode=fast_mode,
                                        n_threads=2, max_dist=extra)
        for i, n in enumerate(names):
            rd = rr.read_dict[n]
            rd.round1_repeat_size = float(dists[i]) / m
            rd.round2_repeat_size = res["r2"][i]
            rd.round3_repeat_size = res["r3"][i]


def _make_regions():
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth
    regs = synth.config1(seed=5, n_regions=5, reads_per_region=7) + synth.config2(seed=6, n_reads=9)
    out = []
    for r in regs:
        rr = nrb.RepeatRegion.from_synth(r)
        rr.data_type = r.data_type
        out.append(rr)
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nanorepeat_b200 import sharding
    regions = _make_regions()
    mine = sharding.estimate_regions_sharded(regions, None, False, max_reads_per_piece=4, estimate_fn=_oracle_estimate)
    res = [[(n, r.round1_repeat_size, r.round2_repeat_size,
             None if r.round3_repeat_size is None else float(r.round3_repeat_size)) for n, r in rr.read_dict.items()]
           for rr in regions]
    q.put((rank, mine, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_run_equals_single_process():
    regions = _make_regions()
    _oracle_estimate(regions)                      # single-process answer, whole regions
    expect = [[(n, r.round1_repeat_size, r.round2_repeat_size,
                None if r.round3_repeat_size is None else float(r.round3_repeat_size)) for n, r in rr.read_dict.items()]
              for rr in regions]
    assert sum(x[3] is not None for reg in expect for x in reg) > 30
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    (r0, mine0, res0), (r1, mine1, res1) = got
    assert set(mine0).isdisjoint(mine1) and mine0 and mine1          # both ranks worked, on different pieces
    assert res0 == expect and res1 == expect                           # every rank ends with every result


def test_partition_is_balanced_and_deterministic():
    from nanorepeat_b200 import sharding
    rng = np.random.default_rng(0)
    costs = [int(x) for x in rng.integers(1, 1000, size=200)] + [50000]
    for world in (1, 2, 4, 8):
        parts = sharding.partition(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        assert parts == sharding.partition(costs, world)
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= max(sum(costs) / world + max(costs), max(costs))
    assert sharding.partition([], 4) == [[], [], [], []]


def test_split_region_pins_the_region_wide_template_size():
    from nanorepeat_b200 import sharding
    rr = _make_regions()[-2]                       # config-2 CAG region, 9 reads
    pieces = sharding.split_region(rr, 4)
    assert [len(p.read_dict) for p in pieces] == [4, 4, 1]
    assert all(p.round1_max_dist == max(r.dist_between_anchors for r in rr.read_dict.values()) for p in pieces)
    assert sharding.split_region(rr, 100) == [rr]
    assert sharding.predicted_cells(rr) > 0
