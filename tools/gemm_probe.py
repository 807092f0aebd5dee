"""Bring-up probe: time cffm_gemm_f16 on the big shapes of the path, optionally with CFFM_GEMM_DEBUG experiments
(bit0: no epilogue stores, bit1: no epilogue work, bit2: no A-tile TMA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vss_cffm_b200 import ops

shapes = [(115200, 256, 64, "f16"), (115200, 64, 64, "f16"), (115200, 64, 64, "res"), (115200, 64, 256, "res"),
          (28800, 512, 128, "f16"), (28800, 128, 512, "res"), (7200, 1024, 256, "f16"), (7938, 768, 256, "f16"),
          (1800, 64, 4096, "f32"), (450, 512, 2880, "f32"), (1800, 512, 512, "f16")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dbg in [0, 1, 2, 4, 6]:
    if dbg:
        os.environ["CFFM_GEMM_DEBUG"] = str(dbg)
    else:
        os.environ.pop("CFFM_GEMM_DEBUG", None)
    line = [f"dbg={dbg}"]
    for M, N, K, mode in shapes:
        a = torch.randn(M, K, device="cuda").half()
        w = torch.randn(N, K, device="cuda").half()
        b = torch.randn(N, device="cuda")
        o16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
        o32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
        kw = dict(out16=o16) if mode == "f16" else (dict(residual=o32, out32=o32) if mode == "res" else dict(out32=o32))
        ts = []
        for it in range(6):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); ops.gemm(a, w, bias=b, **kw); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        line.append(f"{M}x{N}x{K}/{mode}: {min(ts[1:]):.1f}")
    print("  ".join(line), flush=True)
