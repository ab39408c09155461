#!/usr/bin/env python
"""Extract dram bytes per launch of a kernel from an .ncu-rep (ncu --set full) into profiles/attn_traffic.json.
    python tools/ncu_traffic.py gpurun_out/attn_tc3.ncu-rep long_attn_tc2_kernel profiles/attn_traffic.json"""
import csv, io, json, subprocess, sys

rep, kernel, out = sys.argv[1:4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
def col(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return i, scale
ir, sr = col("dram__bytes_read.sum"); iw, sw = col("dram__bytes_write.sum"); it, _ = hdr.index("gpu__time_duration.sum"), None
recs = [r for r in rows[2:] if kernel in r[ik]]
best = max(recs, key=lambda r: float(r[ir]) * sr + float(r[iw]) * sw)
d = {"kernel": kernel, "dram_bytes_per_launch": float(best[ir]) * sr + float(best[iw]) * sw,
     "dram_read": float(best[ir]) * sr, "dram_write": float(best[iw]) * sw,
     "duration_us_under_ncu": float(best[it]), "source": f"ncu --set full capture {rep.split('/')[-1]} (largest launch = c3 long-term layer, T=8, cold L2)"}
json.dump(d, open(out, "w"), indent=1)
print(d)
