"""Summarise ncu CSV logs (launch list / metric pass) of tools/profile_step.py into markdown + a compact CSV.
usage: python tools/summarize_ncu.py <launches.csv> <kernels.csv|-> <out_prefix> [title]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    c = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value")}
    k = collections.OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[c["Kernel Name"]]).replace("cffm::<unnamed>::", "").replace("void ", "")
        d = k.setdefault(r[c["ID"]], {"name": name, "grid": r[c["Grid Size"]], "block": r[c["Block Size"]]})
        d[r[c["Metric Name"]]] = float(r[c["Metric Value"]].replace(",", "") or 0)
    return list(k.values())


def last_step(ks):
    starts = [i for i, d in enumerate(ks) if "im2col_nchw" in d["name"] or "patch_embed" in d["name"]]
    return ks[starts[-1]:]


def main():
    launches, kernels, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
    title = sys.argv[4] if len(sys.argv) > 4 else prefix
    step = last_step(load(launches))
    ours = [d for d in step if not d["name"].startswith("at::")]
    tot = sum(d["gpu__time_duration.sum"] for d in step) / 1e3
    per = collections.OrderedDict()
    for d in step:
        e = per.setdefault(d["name"], [0, 0.0])
        e[0] += 1; e[1] += d["gpu__time_duration.sum"] / 1e3
    out = [f"# {title}", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` over `python tools/profile_step.py 2` "
           "(MiT-B1 + CFFM, 480x480, T=4, B=2); last step only. Per-launch times are cold-cache and serialised: "
           "compare SHARES, not absolutes.", "",
           f"launches in the step: {len(step)} ({len(ours)} from libcffm_b200.so); sum of kernel durations: {tot:.1f} us", "",
           "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for n, (cnt, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n[:70]}` | {cnt} | {t:.1f} | {100 * t / tot:.1f}% | {t / cnt:.1f} |")
    if kernels != "-":
        ks = last_step(load(kernels))
        out += ["", "## Per-launch metrics (second ncu pass, same command)", "",
                "| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | DRAM % | SM % | tensor pipe % | warps active % | regs |",
                "|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
        g = lambda d, n: d.get(n, 0.0)
        with open(prefix + "_kernels.csv", "w") as f:
            f.write("idx,kernel,grid,block,time_us,dram_read_bytes,dram_write_bytes,dram_pct,sm_pct,tensor_pct,warps_active_pct,regs\n")
            for i, d in enumerate(ks):
                vals = (g(d, "gpu__time_duration.sum") / 1e3, g(d, "dram__bytes_read.sum"), g(d, "dram__bytes_write.sum"),
                        g(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        g(d, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                        g(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        g(d, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(d, "launch__registers_per_thread"))
                f.write(f"{i},\"{d['name']}\",\"{d['grid']}\",\"{d['block']}\"," + ",".join(f"{v:.2f}" for v in vals) + "\n")
                if not d["name"].startswith("at::"):
                    out.append(f"| {i} | `{d['name'][:40]}` | {d['grid']} | {vals[0]:.1f} | {vals[1] / 1e6:.1f} | {vals[2] / 1e6:.1f} | "
                               f"{vals[3]:.1f} | {vals[4]:.1f} | {vals[5]:.1f} | {vals[6]:.1f} | {vals[7]:.0f} |")
    open(prefix + ".md", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
