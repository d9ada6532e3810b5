import os, sys, random
sys.path.insert(0, ".")
import torch
from nanorepeat_b200 import engine
engine.init(0)
sc = engine.get_preset("ont")
rng = random.Random(1)
def rs(n): return "".join(rng.choice("ACGT") for _ in range(n))
q = int(os.environ.get("Q", "600")); T = 4000
b = engine.Batch.tasks(sc, [rs(q)], [rs(T)])
for _ in range(3): b.run()
b.fetch_alns()
print("done")
