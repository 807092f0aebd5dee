"""Bring-up timing of a few GEMM shapes of the path (not part of the product). usage: python tools/tune.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vss_cffm_b200 import ops

torch.set_grad_enabled(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1); flush2.max()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
for M, N, K, mode in [(115200, 256, 64, "f16"), (115200, 64, 64, "f16"), (28800, 512, 128, "f16"), (7200, 1280, 320, "f16"),
                      (7938, 768, 256, "f16"), (7200, 1024, 256, "f16"), (1800, 2048, 512, "f16"),
                      (115200, 64, 256, "res"), (28800, 128, 512, "res"), (7200, 320, 1280, "res"), (7200, 256, 1024, "res")]:
    a, w, b = rn(M, K).half(), (rn(N, K) * 0.05).half(), rn(N)
    ref = a.float() @ w.float().t() + b
    if mode == "f16":
        out = torch.empty(M, N, device="cuda", dtype=torch.half)
        fn = lambda: ops.gemm(a, w, bias=b, out16=out)
        fn()
        err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
        by = 2 * M * K + 2 * N * K + 2 * M * N
    else:
        res0 = rn(M, N)
        res = res0.clone()
        fn = lambda: ops.gemm(a, w, bias=b, residual=res, out32=res)
        fn()
        err = ((res - (ref + res0)).abs().max() / ref.abs().max()).item()
        by = 2 * M * K + 2 * N * K + 8 * M * N
    t = timeit(fn)
    print(f"gemm {mode} M={M} N={N} K={K}: {t:.1f} us  {by / t / 1e3:.0f} GB/s  {2 * M * N * K / t / 1e6:.0f} TF/s  err {err:.1e}", flush=True)
