"""Kernel timeline of CUDA-graph replays of the bench workload (torch.profiler = CUPTI activity records): per kernel start,
duration and stream inside one replay, plus idle gaps of the main stream.  usage: python tools/timeline.py [out.csv]"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import vss_cffm_b200 as V
from vss_cffm_b200 import synth

torch.set_grad_enabled(False)
out_csv = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv"
m = V.build_segmentor(V.model_cfg("b1"))
synth.fill_module(m, 21)
m = m.cuda().eval()
B, T, H, W = 2, 4, 480, 480
imgs = [t.cuda() for t in synth.synth_clip(B, T, H, W, seed=100)]
g = m.make_graphed(B, T, H, W, synth.img_metas(B, H, W))
g.load(imgs)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        flush.fill_(1)
        g.replay()
        torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
# last replay = events after the last flush fill kernel
fills = [i for i, e in enumerate(ev) if "fill" in e["name"].lower() or "FillFunctor" in e["name"]]
step = ev[fills[-1] + 1:] if fills else ev
t0 = step[0]["ts"]
end = max(e["ts"] + e["dur"] for e in step)
with open(out_csv, "w") as f:
    f.write("start_us,dur_us,stream,kernel\n")
    for e in step:
        name = e["name"].replace("cffm::(anonymous namespace)::", "").replace("cffm::<unnamed>::", "").replace("void ", "").split("(")[0]
        f.write(f"{e['ts'] - t0:.2f},{e['dur']:.2f},{e['args'].get('stream', '?')},\"{name}\"\n")
print(f"kernels in the replay: {len(step)}; span {end - t0:.1f} us; sum of durations {sum(e['dur'] for e in step):.1f} us")
streams = {}
for e in step:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in streams.items():
    busy = sum(e["dur"] for e in es)
    print(f"stream {s}: {len(es)} kernels, busy {busy:.1f} us")
# union of busy intervals over all streams -> GPU idle time inside the step
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in step)
cur_s, cur_e, busy = iv[0][0], iv[0][1], 0.0
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"GPU busy (union) {busy:.1f} us, idle {end - t0 - busy:.1f} us")
# exclusive time: the span between consecutive kernel END times is attributed to the kernel that ends it
import collections
seq = sorted(step, key=lambda e: e["ts"] + e["dur"])
excl, prev = collections.OrderedDict(), t0
for e in seq:
    name = e["name"].replace("cffm::(anonymous namespace)::", "").replace("void ", "").split("(")[0]
    name = name.split("<")[0] + ("<" + name.split("<")[1] if "gemm" in name and "<" in name else "")
    fin = e["ts"] + e["dur"]
    excl.setdefault(name, [0, 0.0, 0.0])
    excl[name][0] += 1
    excl[name][1] += max(0.0, fin - max(prev, e["ts"]))
    excl[name][2] += e["dur"]
    prev = max(prev, fin)
print("exclusive time per kernel (sums to the span):")
for n, (c, x, d) in sorted(excl.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n[:60]:60s} x{c:<3d} exclusive {x:7.1f} us  ({100 * x / (end - t0):4.1f}%)  sum dur {d:7.1f}")
