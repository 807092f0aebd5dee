"""Bring-up probe: time cffm_mha_f16 on the four MiT stage shapes (B1, 8 frames, 480x480)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vss_cffm_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (B, Nq, Nkv, heads, d) in [(8, 14400, 225, 1, 64), (8, 3600, 225, 2, 64), (8, 900, 225, 5, 64), (8, 225, 225, 8, 64)]:
    C = heads * d
    q = torch.randn(B * Nq, C, device="cuda").half()
    kv = torch.randn(B * Nkv, 2 * C, device="cuda").half()
    out = torch.empty(B * Nq, C, device="cuda", dtype=torch.float16)
    ts = []
    for it in range(6):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.mha(q, kv[:, :C], kv[:, C:], out, B, Nq, Nkv, heads, d, d ** -0.5); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    qh = q.float().view(B, Nq, heads, d).transpose(1, 2)
    kh = kv[:, :C].float().view(B, Nkv, heads, d).transpose(1, 2)
    vh = kv[:, C:].float().view(B, Nkv, heads, d).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1) @ vh).transpose(1, 2).reshape(B * Nq, C)
    err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
    fl = 4 * B * heads * Nq * Nkv * d
    print(f"mha B={B} Nq={Nq} heads={heads}: {min(ts[1:]):7.1f} us  {fl / min(ts[1:]) / 1e6:7.1f} TFLOP/s  rel err {err:.2e}", flush=True)
