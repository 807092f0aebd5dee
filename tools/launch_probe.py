"""Bring-up probe: steady-state per-kernel cost inside a CUDA graph for tiny launches (fixed overhead floor)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vss_cffm_b200 import ops

dev = "cuda"
M, N, K = 1800, 128, 64
a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); b = torch.randn(N, device=dev)
o16 = torch.empty(M, N, device=dev, dtype=torch.float16); o32 = torch.empty(M, N, device=dev)
g = torch.ones(N, device=dev); be = torch.zeros(N, device=dev)
big_a = torch.randn(115200, 64, device=dev).half(); big_w = torch.randn(64, 64, device=dev).half(); big_o = torch.empty(115200, 64, device=dev, dtype=torch.float16)
big_b = torch.randn(64, device=dev)


def gemm16():
    ops.gemm(a, w, bias=b, out16=o16)


def gemm32():
    ops.gemm(a, w, bias=b, out32=o32)


def ln():
    ops.layernorm(o32, g, be, 1e-5, out16=o16)


def biggemm():
    ops.gemm(big_a, big_w, bias=big_b, out16=big_o)


def bench(name, fns, reps=40):
    for f in fns:
        f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            for f in fns:
                f()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        gr.replay()
    e.record(); torch.cuda.synchronize()
    per = s.elapsed_time(e) * 1e3 / (5 * reps * len(fns))
    print(f"{name:40s} {per:7.2f} us per kernel", flush=True)


bench("tiny gemm f16 x40", [gemm16])
bench("tiny gemm f32 x40", [gemm32])
bench("tiny layernorm x40", [ln])
bench("alternating gemm32 / ln", [gemm32, ln])
bench("big q-proj gemm 115200x64x64", [biggemm])
