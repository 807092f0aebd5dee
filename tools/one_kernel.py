"""Launch ONE hot kernel of the path a few times on production shapes (B=2 clips, T=4, 480x480, MiT-B1), for
`ncu --set full -s 2 -c 1`.  usage: python tools/one_kernel.py <gemm_fc1|gemm_fc2|gemm_q|dwconv|ffn|ffn_s2|patch_embed|head_fuse|argmax|mha|cfm|ln>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vss_cffm_b200 import ops

torch.set_grad_enabled(False)
which = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

if which.startswith("gemm"):
    M, N, K, act = {"gemm_fc1": (115200, 256, 64, 0), "gemm_fc2": (115200, 64, 256, 0), "gemm_q": (115200, 64, 64, 0),
                    "gemm_fc1_s2": (28800, 512, 128, 0), "gemm_qkv": (7938, 768, 256, 0)}[which]
    a, w, b = rn(M, K).half(), (rn(N, K) * 0.05).half(), rn(N)
    if which == "gemm_fc2":
        res = rn(M, N)
        fn = lambda: ops.gemm(a, w, bias=b, residual=res, out32=res)
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.half)
        fn = lambda: ops.gemm(a, w, bias=b, out16=out, act=act)
elif which == "dwconv":
    N, H, W, C = 8, 120, 120, 256
    x, w, b, out = rn(N, H, W, C).half(), (rn(9, C) * 0.3).half(), rn(C), torch.empty(N, H, W, C, device="cuda", dtype=torch.half)
    fn = lambda: ops.dwconv3x3_gelu(x, w, b, out, N, H, W, C)
elif which == "patch_embed":
    N, H, W, Co = 8, 480, 480, 64
    x, wk = rn(N, 3, H, W), ops.patch_embed_s1_weight((rn(Co, 3, 7, 7) * 0.08).half())
    pb, g1, e1, g2, e2 = rn(Co), rn(Co), rn(Co), rn(Co), rn(Co)
    o32, o16 = torch.empty(N * 120 * 120, Co, device="cuda"), torch.empty(N * 120 * 120, Co, device="cuda", dtype=torch.half)
    fn = lambda: ops.patch_embed_s1(x, wk, pb, g1, e1, 1e-5, g2, e2, 1e-6, o32, o16)
elif which in ("ffn", "ffn_s2"):
    N, H, W, C, Co = (8, 120, 120, 256, 64) if which == "ffn" else (8, 60, 60, 512, 128)
    h, dw_w, dw_b = rn(N * H * W, C).half(), (rn(9, C) * 0.3).half(), rn(C)
    w2, b2, res, g1, be1 = (rn(Co, C) * 0.05).half(), rn(Co), rn(N * H * W, Co), rn(Co), rn(Co)
    xn = torch.empty(N * H * W, Co, device="cuda", dtype=torch.half)
    fn = lambda: ops.mixffn_tail(h, N, H, W, dw_w, dw_b, w2, b2, res, res, g1, be1, 1e-6, xn)
elif which == "head_fuse":
    N, C = 8, 256
    sizes = [(120, 120), (60, 60), (30, 30), (15, 15)]
    p = [rn(N * a * b, C).half() for a, b in sizes]
    shift, x32, c16 = rn(C), torch.empty(N * 3600, C, device="cuda"), torch.empty(N * 3600, C, device="cuda", dtype=torch.half)
    fn = lambda: ops.head_fuse(p, sizes, N, C, 0, shift, half32=x32, half16=c16)
elif which == "argmax":
    B, h, w = 2, 60, 60
    x, lab = rn(B * h * w, 128), torch.empty(B, 480, 480, dtype=torch.int64, device="cuda")
    fn = lambda: ops.upsample2_argmax(x, 124, lab, B, h, w, 120, 120, 480, 480)
elif which == "mha":
    Nf, Nq, Nkv, heads, d = 8, 14400, 225, 1, 64
    q, kv, out = rn(Nf * Nq, 64).half(), rn(Nf * Nkv, 128).half(), torch.empty(Nf * Nq, 64, device="cuda", dtype=torch.half)
    fn = lambda: ops.mha(q, kv[:, :64], kv[:, 64:], out, Nf, Nq, Nkv, heads, d, d ** -0.5)
elif which == "cfm":
    B, H, W, E = 2, 60, 60, 256
    nW = 81
    from vss_cffm_b200 import cffm_tables as tb
    lay = ops.cfm_layout()
    qkv_t, kvp = rn(ops.apron_rows(B, H, W), 3 * E).half(), rn(B * 15 * nW, 2 * E).half()
    bias = tb.assemble_bias_tc(rn(169, 8) * 0.1, rn(1, 8, 49, 132) * 0.1, rn(8, 121) * 0.1, [rn(8, 169) * 0.1, rn(8, 121) * 0.1, rn(8, 81) * 0.1],
                               32 ** -0.5, lay)
    out = torch.empty(B * H * W, E, device="cuda", dtype=torch.half)
    fn = lambda: ops.cfm_attention(qkv_t, kvp, bias, out, B, H, W, E, 8, 32 ** -0.5)
elif which == "ln":
    M, C = 115200, 64
    x, gm, bt, out = rn(M, C), rn(C), rn(C), torch.empty(M, C, device="cuda", dtype=torch.half)
    fn = lambda: ops.layernorm(x, gm, bt, 1e-6, out16=out)
else:
    raise SystemExit(f"unknown kernel {which}")

for _ in range(3):
    flush.fill_(1)
    fn()
torch.cuda.synchronize()
print("ok", which)
if os.environ.get("TIME"):
    ts = []
    for _ in range(30):
        flush.fill_(1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); fn(); ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1) * 1e3)
    ts.sort()
    print(f"{which}: median {ts[len(ts) // 2]:.2f} us, min {ts[0]:.2f} us (cold L2, CUDA events)")
