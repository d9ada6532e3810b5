"""Where quantify_two_repeats spends its time (host vs library).  usage: joint_profile.py [n_reads]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from nanorepeat_b200 import synth, engine, joint
engine.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
loc = synth.joint_locus(seed=7, n_reads=n)
args = (loc["reads"], loc["left"], loc["mid"], loc["right"], "CAG", "CCG", loc["range1"], loc["range2"], 200, 50)
joint.quantify_two_repeats(*args)
t0 = time.perf_counter(); joint.quantify_two_repeats(*args); print(f"{(time.perf_counter() - t0) * 1e3:.1f} ms")
pr = cProfile.Profile(); pr.enable(); joint.quantify_two_repeats(*args); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
