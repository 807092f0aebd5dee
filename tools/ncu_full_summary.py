"""One-line-per-kernel summary of `ncu --set full` reports. usage: python tools/ncu_full_summary.py out.md rep1 rep2 ..."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "DRAM rd MB", 1e-6), ("dram__bytes_write.sum", "DRAM wr MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX %", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %", 1),
        ("smsp__issue_active.avg.pct", "issue %", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1)]
out = ["# ncu --set full --clock-control none: the hot kernels on production shapes (tools/one_kernel.py, third launch)", "",
       "| kernel | " + " | ".join(k[1] for k in KEYS) + " |", "|---|" + "---:|" * len(KEYS)]
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, unit, val = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
    import re
    name = re.sub(r"\((int|bool)\)", "", d["Kernel Name"][0]).replace("void ", "")
    name = re.sub(r"cffm::(<unnamed>|\(anonymous namespace\))::", "", name).split("(")[0]
    cells = []
    for k, _, sc in KEYS:
        if k not in d:
            cells.append("-"); continue
        v, u = d[k]
        x = float(v.replace(",", ""))
        if k == "gpu__time_duration.sum":
            x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
        elif "bytes" in k:
            x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        cells.append(f"{x:.1f}" if x < 1e5 else f"{x:.0f}")
    out.append(f"| `{name[:48]}` ({rep.split('_')[-1].split('.')[0]}) | " + " | ".join(cells) + " |")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
