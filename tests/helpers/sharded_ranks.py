"""Run under torchrun by tests/test_gpu_multi.py: one workload, split over the ranks by estimate_regions_sharded (each rank
its own GPU), gathered on the host; rank 0 compares with the unsharded run on its GPU and prints a JSON verdict."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch.distributed as dist                     # noqa: E402
import nanorepeat_b200 as nrb                        # noqa: E402
from nanorepeat_b200 import engine, synth, sharding  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")                  # the data path has no collective; the gather is host-side
    engine.init(local)
    assert engine.device_info()["device"] == local
    regs = synth.config1(seed=3, n_regions=6, reads_per_region=20) + synth.config2(seed=4, n_reads=300) + \
        synth.config4(seed=5, reads_per_locus=6, scale=0.2)
    rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
    mine = sharding.estimate_regions_sharded(rrs, "ont", False, max_reads_per_piece=128)
    counts = [None] * dist.get_world_size()
    dist.all_gather_object(counts, len(mine))
    if rank == 0:
        whole = [nrb.RepeatRegion.from_synth(r) for r in regs]
        nrb.estimate_regions(whole, "ont", False)
        same = all((a.read_dict[n].round1_repeat_size, a.read_dict[n].round2_repeat_size, a.read_dict[n].round3_repeat_size,
                    type(a.read_dict[n].round3_repeat_size)) ==
                   (b.read_dict[n].round1_repeat_size, b.read_dict[n].round2_repeat_size, b.read_dict[n].round3_repeat_size,
                    type(b.read_dict[n].round3_repeat_size))
                   for a, b in zip(whole, rrs) for n in a.read_dict)
        n3 = sum(rd.round3_repeat_size is not None for rr in rrs for rd in rr.read_dict.values())
        print(json.dumps({"equal": same, "pieces_per_rank": counts, "reads_with_round3": n3}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
