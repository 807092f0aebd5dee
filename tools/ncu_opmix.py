"""Per-opcode executed warp-instruction counts and top stall lines from an ncu report's source page.
usage: python tools/ncu_opmix.py <report.ncu-rep> [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
print(rows[0][1][:150])
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
cnt, tot, samp = collections.Counter(), 0, []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    n = int(r[ci["Instructions Executed"]] or 0)
    toks = r[ci["Source"]].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    cnt[op.split(".")[0].rstrip(";")] += n
    tot += n
    samp.append((int(r[ci["# Samples"]] or 0), n, r[ci["Source"]].strip()))
print("warp instructions executed:", tot)
print("  ".join(f"{o}:{100 * n / tot:.1f}%" for o, n in cnt.most_common(topn)))
stot = sum(s for s, _, _ in samp) or 1
print("top stall-sample lines:")
for s, n, src in sorted(samp, reverse=True)[:topn]:
    print(f"  {100 * s / stot:5.1f}%  x{n:<9d} {src[:110]}")
