"""DRAM traffic per kernel family of one step, from the per-launch ncu metric pass summarised by tools/summarize_ncu.py
(<prefix>_kernels.csv) -> JSON that bench.py reports as roofline.traffic.
usage: python tools/ncu_traffic.py profiles/r01_v17_step_kernels.csv profiles/r01_traffic.json"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
fam = collections.OrderedDict()
for r in csv.DictReader(open(src)):
    name = re.sub(r"<.*", "", r["kernel"])
    e = fam.setdefault(name, {"launches_per_step": 0, "dram_read_bytes_per_step": 0.0, "dram_write_bytes_per_step": 0.0, "time_us_per_step": 0.0})
    e["launches_per_step"] += 1
    e["dram_read_bytes_per_step"] += float(r["dram_read_bytes"])
    e["dram_write_bytes_per_step"] += float(r["dram_write_bytes"])
    e["time_us_per_step"] += float(r["time_us"])
for e in fam.values():
    e["dram_bytes_per_launch"] = round((e["dram_read_bytes_per_step"] + e["dram_write_bytes_per_step"]) / e["launches_per_step"])
    e["time_us_per_step"] = round(e["time_us_per_step"], 1)
out = {"source": src, "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none "
                             "over tools/profile_step.py 2 (MiT-B1 + CFFM, 480x480, T=4, 2 clips), last step; one row per launch in the source CSV",
       "families": fam}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in fam.items()}))
