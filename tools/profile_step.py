"""One or more inference steps of the BASELINE workload (MiT-B1 + CFFM, 480x480, T=4, B=2) for ncu.
usage: python tools/profile_step.py [steps] [variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vss_cffm_b200 as V
from vss_cffm_b200 import synth

torch.set_grad_enabled(False)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
variant = sys.argv[2] if len(sys.argv) > 2 else "b1"
m = V.build_segmentor(V.model_cfg(variant))
synth.fill_module(m, 21)
m = m.cuda().eval()
imgs = [t.cuda() for t in synth.synth_clip(2, 4, 480, 480, seed=100)]
metas = synth.img_metas(2, 480, 480)
for _ in range(steps):
    lab = m.predict_labels(imgs, metas)
torch.cuda.synchronize()
print("launches", V._abi.n_launches, "labels", tuple(lab.shape))
