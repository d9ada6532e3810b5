#!/usr/bin/env python
"""DRAM traffic per launch of one kernel from an .ncu-rep (ncu --set full): dram__bytes_read.sum + dram__bytes_write.sum,
averaged over the captured launches -> profiles/r02_traffic.json, which bench.py quotes as roofline.traffic.
usage: ncu_traffic.py prof.ncu-rep <kernel name substring> <out.json> "<command line that was profiled>" """
import csv, io, json, subprocess, sys
rep, pat, out, cmd = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return i, scale
ir, sr = col("dram__bytes_read.sum"); iw, sw = col("dram__bytes_write.sum"); it = hdr.index("gpu__time_duration.sum")
rd, wr, n, names = 0.0, 0.0, 0, set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat not in name:
        continue
    rd += float(r[ir]) * sr; wr += float(r[iw]) * sw; n += 1; names.add(name.split("(")[0])
doc = {"kernel": sorted(names), "launches": n, "bytes_per_launch": int((rd + wr) / max(n, 1)), "read_bytes_per_launch": int(rd / max(n, 1)),
       "write_bytes_per_launch": int(wr / max(n, 1)), "source": f"ncu --set full --clock-control none, {n} launches of: {cmd}"}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc))
