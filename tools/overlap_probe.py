"""Probe: throughput of CUDA-graph replays of N independent passes (own workspaces) on N streams vs one stream.
usage: [FLUSH=1] python tools/overlap_probe.py [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vss_cffm_b200 as V
from vss_cffm_b200 import synth

torch.set_grad_enabled(False)
B, T, H, W = 2, 4, 480, 480
from vss_cffm_b200.graph import GraphedClips
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = V.build_segmentor(V.model_cfg("b1"))
synth.fill_module(m, 21)
m = m.cuda().eval()
graphs = []
for i in range(NP):
    g = GraphedClips(m, B, T, H, W, synth.img_metas(B, H, W), warmup=2, private_input=True, private_workspace=True)
    g.load([t.cuda() for t in synth.synth_clip(B, T, H, W, seed=100 + i)])
    graphs.append(g)
s = [torch.cuda.Stream() for _ in range(NP)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()


def run(n, two_streams):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for st in s:
        st.wait_event(e0)
    for i in range(n):
        st = s[i % NP] if two_streams else s[0]
        with torch.cuda.stream(st):
            if os.environ.get('FLUSH'):
                flush.fill_(1)
            graphs[i % NP].replay()
    for st in s:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(2):
    run(10, False); run(10, True)
a = min(run(40, False) for _ in range(3))
b = min(run(40, True) for _ in range(3))
print(f"one stream: {a:.4f} ms/step ({8 / a * 1e3:.0f} clip-frames/s); {NP} streams: {b:.4f} ms/step ({8 / b * 1e3:.0f} clip-frames/s); ratio {a / b:.3f} (warm L2, no flush)")
