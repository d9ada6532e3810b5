#!/usr/bin/env python
"""bench.py -- repeat-size estimation hot path (rounds 2 + 3) on BASELINE.json's config 2.

A step = one pass of the hot path over one batch: the HTT amplicon, 5 000 synthetic ONT reads, quantified as the
two BED rows of example_data/HTT_repeat_region.bed (CAG and CCG) -> 10 000 (read, region) units, each aligned
against its round-2 template and its round-3 ladder (31+ rungs).

  value  GCUPS = algorithmic DP cells (full rectangles |core| x |template|) per second, inputs
         resident in HBM, CUDA-event time of the kernel launches only, summed over K steps, max over ranks.
  e2e    same metric through the operator API (round1_and_round2_estimation + round3_estimation) with host
         strings in, Python attributes out: packing, H2D, kernels, D2H, selection all inside the timed region.
  roofline  DPX/integer-pipe bound: executed cells / device time of the dominant kernel (the paired round-3 ladder,
         CUDA events around the launch on its stream) against
         SMs x sm_max_mhz x 64 DPX lanes/clk/SM x 2 cells per lane-instr / 7 DPX instr per cell pair
         (32-bit kernels beside it: 1 cell per lane-instr / 6 DPX instr per cell).
  cpu_baseline / --impl reference  the CPU oracle port (oracle/nr_oracle.c) on the host cores, bounded sample.

N > 1 (torchrun): every rank runs its own batch (seed + rank) -- weak scaling, no data-path collective; NCCL is
used only for the barrier and the max-over-ranks of the times.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DPX_LANES_PER_CLK_PER_SM = 64       # measured: tools/microbench/pipe_rates.cu -> profiles/pipe_rates_r01.jsonl
DPX_INSTR_PER_CELL = 6              # 32-bit word: 2 (five-way max + floor for H) + 4 (E1, E2, F1, F2 updates); the
                                    # running-max op (0.5/cell) counts against the kernel
PAIR_LADDER_TRAFFIC = 74990848      # bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r01_ncu_pair_ladder_kernel_v41_summary.txt); the writes include dirty lines of the L2 flush buffer that the launch evicts
DPX_INSTR_PER_CELL_PAIR = 7         # u16x2 words (two cells per instruction): the floor needs an operand of its own
                                    # (no .RELU on unsigned halves): 3 for H + 4 -> 3.5 per cell


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(seed, n_reads):
    from nanorepeat_b200 import synth
    return synth.config2(seed=seed, n_reads=n_reads)


def cells_of(regs, T_list, kmins, kmaxs, r2_valid):
    from nanorepeat_b200 import synth
    total2 = total3 = 0
    for reg, T, kmin, kmax, ok in zip(regs, T_list, kmins, kmaxs, r2_valid):
        nl, nr_, m = len(reg.left_anchor_seq), len(reg.right_anchor_seq), len(reg.repeat_unit_seq)
        for i, core in enumerate(reg.core_seqs):
            total2 += len(core) * (nl + m * T)
            if ok[i]:
                _, c3 = synth.algorithmic_cells(nl, nr_, m, len(core), T, int(kmin[i]), int(kmax[i]))
                total3 += c3
    return total2, total3


def run_reference_arm(args, rank, world):
    """Oracle port on the host cores (the reference's engine, pyminimap2, is not installable here)."""
    if rank != 0:
        return
    from oracle import nr_oracle, selection
    nr_oracle.build()
    threads = nr_oracle.max_threads()
    n_sample = args.cpu_sample_reads
    regs = make_workload(args.seed, n_sample)
    sc = nr_oracle.scoring()

    def one_step():
        cells = 0
        for reg in regs:
            res = selection.estimate_region(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                                            reg.core_seqs, reg.dist_between_anchors, sc=sc, n_threads=threads)
            ok = [r is not None for r in res["r2"]]
            c2, c3 = cells_of([reg], [res["T"]], [[k if k is not None else 0 for k in res["kmin"]]],
                              [[k if k is not None else -1 for k in res["kmax"]]], [ok])
            cells += c2 + c3
        return cells

    for _ in range(args.warmup if args.warmup < 1 else 1):
        one_step()
    t0 = time.perf_counter()
    cells = 0
    for _ in range(args.steps):
        cells += one_step()
    dt = time.perf_counter() - t0
    gcups = cells / dt / 1e9
    units = 2 * n_sample * args.steps
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS (1e9 DP cells/s)", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "reads_per_s": units / dt,
        "config": {"workload": "config 2: HTT CAG/CCG amplicon, ONT reads, two BED rows, rounds 2+3",
                   "reads": n_sample, "units_per_step": 2 * n_sample},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": f"{n_sample} of the workload's 5000 reads (same seed), both regions, rounds 2+3, "
                                   f"full rectangles, {threads} pthreads"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference engine pyminimap2>=2.30 is absent and not installable offline; this arm times the CPU "
                "oracle port of the same exact DP (oracle/nr_oracle.c)",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=5000, help="reads in the batch (config 2: 5000)")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-sample-reads", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if args.warmup < 3:
        args.warmup = 3
    import torch
    import torch.distributed as dist
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import engine
    from nanorepeat_b200.estimation import ladder_bounds

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    engine.init(local_rank)
    info = engine.device_info()
    peaks, peak_src = load_peaks()

    # ---- workload (weak scaling: every rank its own batch) ----
    regs = make_workload(args.seed + 1000 * rank, args.reads)
    data_type = "ont"
    sc = engine.get_preset(data_type)

    # ---- e2e leg: the operator API, host strings in, attributes out ----
    # The caller's RepeatRegion / Read objects (what Step 1 of the reference hands over) are built outside the timed
    # region, one fresh set per step; the timed call is the two operators over them.
    def fresh_regions():
        return [nrb.RepeatRegion.from_synth(reg) for reg in regs]

    def e2e_step(rrs):
        nrb.estimate_regions(rrs, data_type, False)      # round1_and_round2_estimation + round3_estimation, batched
        return rrs

    rrs = e2e_step(fresh_regions())     # also the first warm-up; gives r2 -> ladders for the resident batches
    h2d = d2h = 0
    T_list, kmins, kmaxs, valid = [], [], [], []
    for reg, rr in zip(regs, rrs):
        m = len(reg.repeat_unit_seq)
        r1max = max(float(d) / m for d in reg.dist_between_anchors)
        T = int(r1max * 1.5) + 1
        if T < r1max + 10:
            T = int(r1max + 10)
        T_list.append(T)
        lo, hi, ok = [], [], []
        for name in reg.read_names:
            r2 = rr.read_dict[name].round2_repeat_size
            ok.append(r2 is not None)
            a, b = ladder_bounds(r2, False) if r2 is not None else (0, -1)
            lo.append(a); hi.append(b)
        kmins.append(np.asarray(lo, np.int32)); kmaxs.append(np.asarray(hi, np.int32)); valid.append(ok)
    cells2, cells3 = cells_of(regs, T_list, kmins, kmaxs, valid)
    cells_step = cells2 + cells3
    units_step = sum(len(r.core_seqs) for r in regs)

    # ---- C-ABI leg: host byte buffers in, numpy records out (pack + H2D + kernels + D2H + selection) ----
    r2_specs = [(reg.left_anchor_seq, reg.repeat_unit_seq, T, reg.core_seqs) for reg, T in zip(regs, T_list)]
    r3_specs = []
    for reg, lo, hi, ok in zip(regs, kmins, kmaxs, valid):
        idx = [i for i, v in enumerate(ok) if v]
        r3_specs.append((reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                         [reg.core_seqs[i] for i in idx], lo[idx], hi[idx]))

    r3_reuse = []       # (right anchor, kmin, kmax over ALL reads of the region; kmax < kmin skips a read)
    for reg, lo, hi, ok in zip(regs, kmins, kmaxs, valid):
        okm = np.asarray(ok, bool)
        r3_reuse.append((reg.right_anchor_seq, np.where(okm, lo, 0).astype(np.int32), np.where(okm, hi, -1).astype(np.int32)))

    def cabi_step():
        b2c = engine.Batch.begin(sc, "round2_flags")
        for spec in r2_specs:
            b2c.add_round2(*spec)
        a = b2c.commit().run().fetch_round2()
        b3c = engine.Batch.begin_round3_from(b2c)        # the reads stay packed in HBM between the rounds
        for i, (right, lo, hi) in enumerate(r3_reuse):
            b3c.add_round3_reuse(i, right, lo, hi)
        s = b3c.commit().run().fetch_round3()
        b3c.close(); b2c.close()
        return a, s

    # ---- resident batches: one per round over both regions ----
    b2 = engine.Batch.begin(sc, "round2_flags")
    for spec in r2_specs:
        b2.add_round2(*spec)
    b2.commit()
    b3 = engine.Batch.begin_round3_from(b2)          # the production flow: round 3 over the reads round 2 left in HBM,
    for i, (right, lo, hi) in enumerate(r3_reuse):   # resuming from the DP state round 2 kept at the end of the left anchor
        b3.add_round3_reuse(i, right, lo, hi)
    batches = [b2, b3.commit()]
    stats = [b.stats() for b in batches]
    executed_step = sum(s["executed_cells"] for s in stats)
    algorithmic_check = sum(s["algorithmic_cells"] for s in stats)
    launches_step = 0
    h2d = sum(s["h2d_bytes"] for s in stats)
    d2h = sum(s["d2h_bytes"] for s in stats)

    stream = torch.cuda.Stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def resident_step():
        for b in batches:
            b.run(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        resident_step()
    torch.cuda.synchronize()
    launches_step = sum(b.stats()["kernel_launches"] for b in batches)

    engine.set_timing(True)             # CUDA events around every kernel, on the stream it is launched on
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, r2_ms, r3_ms = [], [], []
    kern_ms = [dict(paired_ms=0.0, rest_ms=0.0, redo_ms=0.0) for _ in batches]
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                                  # evict L2 between timed iterations (untimed)
        torch.cuda.synchronize()
        e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        with torch.cuda.stream(stream):
            e0.record(stream)
            batches[0].run(stream.cuda_stream)          # round 2: exact_kernel
            em.record(stream)
            batches[1].run(stream.cuda_stream)          # round 3: ladder_kernel
            e1.record(stream)
        e1.synchronize()
        dev_ms.append(e0.elapsed_time(e1)); r2_ms.append(e0.elapsed_time(em)); r3_ms.append(em.elapsed_time(e1))
        for acc, b in zip(kern_ms, batches):
            li = b.launch_info()
            for key in acc:
                acc[key] += li[key]
    barrier()
    engine.set_timing(False)
    linfo = [b.launch_info() for b in batches]
    wall = time.perf_counter() - wall0
    total_ms = float(sum(dev_ms))

    # ---- e2e timed regions ----
    for _ in range(2):
        e2e_step(fresh_regions())
        cabi_step()
    sets = [fresh_regions() for _ in range(args.steps)]
    import gc
    gc.collect()
    barrier()
    t0 = time.perf_counter()
    for rr_set in sets:
        e2e_step(rr_set)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    del sets
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _a2, s3_cabi = cabi_step()
    torch.cuda.synchronize()
    cabi_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()

    # the resident path (valid reads only) and the C-ABI path (all reads, skipped ones zero) must give the same records
    s3 = batches[1].fetch_round3()
    assert all(np.array_equal(x, y) for x, y in zip(s3, s3_cabi)), "resident and C-ABI round-3 results differ"
    # ... and the same as a fresh round-3 batch (full forward sweeps, nothing taken over from round 2)
    with engine.Batch.begin(sc, "round3") as fresh3:
        for spec in r3_specs:
            fresh3.add_round3(*spec)
        s3_fresh = fresh3.commit().run().fetch_round3()
    vmask = np.concatenate([np.asarray(ok, bool) for ok in valid])
    assert all(np.array_equal(x[vmask], y) for x, y in zip(s3, s3_fresh)), "resumed and fresh round-3 results differ"

    if world > 1:
        t = torch.tensor([total_ms, e2e_s, cabi_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, cabi_s = float(t[0]), float(t[1]), float(t[2])
        c = torch.tensor([cells_step, executed_step, units_step, launches_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_all, executed_all, units_all, launches_all = (float(x) for x in c)
    else:
        cells_all, executed_all, units_all, launches_all = cells_step, executed_step, units_step, launches_step

    if rank == 0:
        K = args.steps
        value = cells_all * K / (total_ms * 1e-3) / 1e9
        e2e_val = cells_all * K / e2e_s / 1e9
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        peak32 = info["sm_count"] * sm_max * 1e6 * DPX_LANES_PER_CLK_PER_SM / DPX_INSTR_PER_CELL / 1e9
        peak16 = info["sm_count"] * sm_max * 1e6 * DPX_LANES_PER_CLK_PER_SM * 2 / DPX_INSTR_PER_CELL_PAIR / 1e9

        def kern(cells, ms, peak):
            if not cells or ms <= 0:
                return None
            a = cells / (ms / K * 1e-3) / 1e9
            return {"achieved": a, "peak": peak, "frac": a / peak, "ms_per_launch": ms / K, "executed_cells_per_launch": cells}

        kernels = {
            # the long reads' 32-bit entries run inside the same launches (their cells are < 1 % of config 2's)
            "pair_round2_kernel (round 2, u16x2)": kern(linfo[0]["paired_cells"], kern_ms[0]["paired_ms"], peak16),
            "exact_kernel<fixed scoring> (round 2, separate launch)": kern(linfo[0]["rest_cells"], kern_ms[0]["rest_ms"], peak32),
            "pair_ladder_kernel (round 3, u16x2)": kern(linfo[1]["paired_cells"], kern_ms[1]["paired_ms"], peak16),
            "ladder_kernel<fixed scoring, flag words> (round 3, separate launch)": kern(linfo[1]["rest_cells"], kern_ms[1]["rest_ms"], peak32),
        }
        kernels = {k: v for k, v in kernels.items() if v}
        # the step against the roofline: time the DPX pipe needs for the executed cells of every kernel / time taken
        ideal_ms = sum((li["paired_cells"] / peak16 + li["rest_cells"] / peak32) / 1e9 * 1e3 for li in linfo)
        step_ms = float(sum(dev_ms)) / K
        dom = kernels.get("pair_ladder_kernel (round 3, u16x2)") or next(iter(kernels.values()))
        dom_name = "pair_ladder_kernel (round 3, u16x2)" if "pair_ladder_kernel (round 3, u16x2)" in kernels else next(iter(kernels))
        algo_bytes = h2d + d2h
        line = {
            "metric": "GCUPS", "value": value, "unit": "GCUPS (1e9 DP cells/s, full rectangles)", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "u16x2 (2*score + mark + 64 per half, two reads per word); s32 (score*65536 - span) for reads > 384 bases",
            "data": "synthetic",
            "reads_per_s": units_all * K / (total_ms * 1e-3),
            "config": {"workload": "config 2: HTT CAG/CCG amplicon, 5k ONT reads, two BED rows, rounds 2+3",
                       "reads": args.reads, "units_per_step_per_gpu": units_step,
                       "cells_per_step_per_gpu": cells_step, "l2": "flushed between timed steps (256 MB fill)",
                       "seed": args.seed},
            "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "reads_per_s": units_all * K / e2e_s, "ms_per_step": e2e_s / K * 1e3,
                    "path": "nanorepeat_b200.estimate_regions on RepeatRegion/Read objects (host strings in, "
                            "Read.round{1,2,3}_repeat_size out)",
                    "c_abi": {"value": cells_all * K / cabi_s / 1e9, "unit": "GCUPS",
                              "reads_per_s": units_all * K / cabi_s, "ms_per_step": cabi_s / K * 1e3,
                              "path": "nr_batch_begin/add/commit/run/fetch with host byte buffers in, records out"}},
            "gpu_launches": int(launches_all * K),
            "clocks": clocks,
            "roofline": {"bound": "dpx", "kernel": dom_name,
                         "achieved": dom["achieved"], "peak": dom["peak"], "unit": "GCUPS",
                         "frac": dom["frac"], "ms_per_launch": dom["ms_per_launch"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel, ncu --set full
                         "traffic": PAIR_LADDER_TRAFFIC, "traffic_unit": "bytes per launch (ncu)",
                         "algorithmic_bytes_per_launch": stats[1]["h2d_bytes"] + stats[1]["d2h_bytes"],
                         "kernels": kernels,
                         "step": {"frac": ideal_ms / step_ms, "ms": step_ms, "dpx_ideal_ms": ideal_ms,
                                  "executed_gcups": executed_step / (step_ms * 1e-3) / 1e9,
                                  "def": "sum over kernels of executed cells / that kernel's DPX peak, over the step's device time"},
                         "redo_reads_per_step": linfo[1]["n_redo"],
                         "peak_def": f"{info['sm_count']} SMs x {sm_max:.0f} MHz ({peak_src}) x "
                                     f"{DPX_LANES_PER_CLK_PER_SM} DPX lanes/clk/SM (measured) x 2 cells per lane-instr / "
                                     f"{DPX_INSTR_PER_CELL_PAIR} DPX instr per cell pair (u16x2 kernels); "
                                     f"x 1 / {DPX_INSTR_PER_CELL} for the 32-bit kernels ({peak32:.0f} GCUPS)",
                         "executed_cells_per_step": executed_step, "algorithmic_cells_per_step": cells_step,
                         "hbm_gbs_algorithmic": algo_bytes / (total_ms / K * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "wall_s_timed_region": wall,
        }
        assert algorithmic_check == cells_step, (algorithmic_check, cells_step)
        if not args.no_cpu_baseline:
            from oracle import nr_oracle, selection
            nr_oracle.build()
            threads = nr_oracle.max_threads()
            sample = make_workload(args.seed, args.cpu_sample_reads)
            t0 = time.perf_counter()
            c = 0
            for reg in sample:
                res = selection.estimate_region(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                                                reg.core_seqs, reg.dist_between_anchors, n_threads=threads)
                ok = [r is not None for r in res["r2"]]
                c2, c3 = cells_of([reg], [res["T"]], [[k if k is not None else 0 for k in res["kmin"]]],
                                  [[k if k is not None else -1 for k in res["kmax"]]], [ok])
                c += c2 + c3
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": c / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_sample_reads} reads x 2 regions of the same workload, "
                                              f"rounds 2+3, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
