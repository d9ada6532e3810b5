"""Run as a script by tests/test_boundary.py: the reference's process model (nanoRepeat_bam.py:712-731) on the drop-in.
The parent imports everything and loads the library (no CUDA call), forks P workers, each worker quantifies regions
i = pid, pid + P, ... on the GPU and returns them PICKLED through a multiprocessing queue; the parent then computes the
same regions itself (its first CUDA call comes after the children are done) and compares."""
import json
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import nanorepeat_b200 as nrb                       # noqa: E402
from nanorepeat_b200 import engine, synth           # noqa: E402
from helpers import refshape                        # noqa: E402


def sizes(rr):
    return [(n, rd.round1_repeat_size, rd.round2_repeat_size, None if rd.round3_repeat_size is None else float(rd.round3_repeat_size))
            for n, rd in rr.read_dict.items()]


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    nrb.install(refshape)
    engine.lib()                                    # the .so is mapped before the fork, CUDA is not initialised
    engine.get_preset("ont")
    regs = synth.config1(seed=77, n_regions=10, reads_per_region=9) + synth.config4(seed=78, reads_per_locus=3, scale=0.3)
    regions = [refshape.region_from_synth(r) for r in regs]
    ctx = mp.get_context("fork")                    # the reference's default start method on Linux
    q = ctx.Queue()
    procs = [ctx.Process(target=refshape.worker, args=(pid, P, False, "ont", regions, q)) for pid in range(P)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        for rr in q.get(timeout=600):               # unpickled RepeatRegion objects
            got[rr.left_anchor_seq[:40] + rr.repeat_unit_seq] = sizes(rr)
    for p in procs:
        p.join(timeout=60)
    codes = [p.exitcode for p in procs]
    mine = {}
    for rr in regions:                              # the parent's own objects were not touched by the children
        assert all(rd.round2_repeat_size is None for rd in rr.read_dict.values())
        refshape.quantify1repeat("parent", 1, False, "ont", rr)
        mine[rr.left_anchor_seq[:40] + rr.repeat_unit_seq] = sizes(rr)
    print(json.dumps({"exitcodes": codes, "regions_back": len(got), "equal": got == mine,
                      "reads_with_round3": sum(1 for v in mine.values() for r in v if r[3] is not None)}))


if __name__ == "__main__":
    main()
