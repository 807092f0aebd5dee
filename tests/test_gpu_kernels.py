"""Per-kernel parity: every C-ABI entry point vs the CPU oracle / a plain fp32 PyTorch restatement of
the same op on the same seeded inputs (inputs are rounded to fp16 first, so the comparison isolates
the kernel: fp32 accumulation order + one fp16 output rounding).

Tolerance (north_star: 1e-3 relative in fp16): max|gpu - ref| <= 1e-3 * max|ref| for fp16 outputs,
1e-4 for fp32 outputs; integer / index work is bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cffm_oracle as O
from vss_cffm_b200 import cffm_tables as tb
from vss_cffm_b200 import synth

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
REL16, REL32 = 1e-3, 1e-4
CFM_TOL = 1e-3


@pytest.fixture(scope="module")
def ops():
    from vss_cffm_b200 import _abi, ops as _ops
    _abi.require_device()
    return _ops


def h16(t):
    """fp32 tensor rounded to fp16-representable values."""
    return t.half().float()


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def check(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"
    return e


# ------------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [
    # (M, N, K)  -- tails in M, N < tile, K < 64, K not a multiple of 64, the real path shapes
    (128, 64, 64), (300, 32, 32), (1000, 256, 152), (4100, 768, 256), (257, 128, 576), (513, 320, 1152),
    (225, 512, 2880), (450, 64, 4096), (3600, 1024, 256), (3600, 256, 1024), (129, 128, 512), (64, 8, 8),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tcgen05_vs_fp32(ops, M, N, K):
    a = h16(synth.synth_array((M, K), 1)).cuda()
    w = h16(synth.synth_array((N, K), 2, scale=K ** -0.5)).cuda()
    bias = synth.synth_array((N,), 3).cuda()
    res = synth.synth_array((M, N), 4).cuda()
    ref = a.double() @ w.double().t() + bias.double()
    # plain: bias only, both outputs
    o16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
    o32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(a.half(), w.half(), bias=bias, out16=o16, out32=o32)
    check(o32, ref, REL32, "gemm f32")
    check(o16, ref, REL16, "gemm f16")
    # cross-check kernel agrees too
    c32 = torch.empty_like(o32)
    ops.gemm(a.half(), w.half(), bias=bias, out32=c32, impl=ops.GEMM_CHECK)
    check(c32, ref, REL32, "gemm check impl")
    # GELU + residual, in place on the residual buffer
    ref2 = F.gelu(ref) + res.double()
    inplace = res.clone()
    ops.gemm(a.half(), w.half(), bias=bias, residual=inplace, out32=inplace, act=ops.ACT_GELU)
    check(inplace, ref2, REL32, "gemm gelu+residual in place")
    # ReLU, no bias
    ops.gemm(a.half(), w.half(), out32=o32, act=ops.ACT_RELU)
    check(o32, F.relu(a.double() @ w.double().t()), REL32, "gemm relu")


def test_gemm_strided_operands(ops):
    """Column-sliced W (ldw > K), row-sliced W and a strided output: the classifier / qkv call sites."""
    M, K = 500, 256
    a = h16(synth.synth_array((M, K), 5)).cuda().half()
    wfull = h16(synth.synth_array((128, 2 * K), 6, scale=0.05)).cuda().half()
    out = torch.zeros(M, 128, dtype=torch.float32, device="cuda")
    ops.gemm(a, wfull[:, :K], out32=out)
    ops.gemm(a, wfull[:, K:], residual=out, out32=out)
    ref = a.double() @ wfull[:, :K].double().t() + a.double() @ wfull[:, K:].double().t()
    check(out, ref, REL32, "two-part classifier")
    w3 = h16(synth.synth_array((768, K), 7, scale=0.05)).cuda().half()
    big = torch.zeros(M, 1024, dtype=torch.float16, device="cuda")
    ops.gemm(a, w3[256:], out16=big[:, 512:])
    check(big[:, 512:], a.double() @ w3[256:].double().t(), REL16, "row-sliced W, strided out")
    assert big[:, :512].abs().max().item() == 0


def test_gemm_rejects_bad_arguments(ops):
    from vss_cffm_b200._abi import CffmError
    a = torch.zeros(16, 12, dtype=torch.float16, device="cuda")
    w = torch.zeros(8, 12, dtype=torch.float16, device="cuda")
    o = torch.zeros(16, 8, dtype=torch.float32, device="cuda")
    with pytest.raises(CffmError, match="UNSUPPORTED"):
        ops.gemm(a, w, out32=o)                                  # K % 8 != 0
    with pytest.raises(CffmError):
        ops.gemm(a.cpu(), w, out32=o)                            # no CPU fallback


# ------------------------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("M,C,f32in", [(1000, 64, True), (77, 320, True), (513, 512, False), (9, 32, True), (300, 256, False)])
def test_layernorm(ops, M, C, f32in):
    x = synth.synth_array((M, C), 8, scale=2.0)
    if not f32in:
        x = h16(x)
    g, b = synth.synth_array((C,), 9) * 0.1 + 1, synth.synth_array((C,), 10) * 0.1
    for eps in (1e-5, 1e-6):
        ref = F.layer_norm(x.double(), (C,), g.double(), b.double(), eps)
        o16 = torch.empty(M, C, dtype=torch.float16, device="cuda")
        o32 = torch.empty(M, C, dtype=torch.float32, device="cuda")
        ops.layernorm(x.cuda() if f32in else x.cuda().half(), g.cuda(), b.cuda(), eps, out16=o16, out32=o32)
        check(o32, ref, REL32, "layernorm f32")
        check(o16, ref, REL16, "layernorm f16")


@pytest.mark.parametrize("S,M,C", [(1, 1000, 64), (4, 300, 320), (8, 77, 512), (2, 129, 128), (1, 9, 32)])
def test_layernorm_chain(ops, S, M, C):
    """LN(sum of split-K partials + bias) -> fp32, then LN2 of that -> fp16 (patch-embed norm + first norm1)."""
    parts = synth.synth_array((S, M, C), 11, scale=2.0)
    bias = synth.synth_array((C,), 12)
    g1, b1 = synth.synth_array((C,), 13) * 0.1 + 1, synth.synth_array((C,), 14) * 0.1
    g2, b2 = synth.synth_array((C,), 15) * 0.1 + 1, synth.synth_array((C,), 16) * 0.1
    y = F.layer_norm(parts.double().sum(0) + bias.double(), (C,), g1.double(), b1.double(), 1e-5)
    z = F.layer_norm(y, (C,), g2.double(), b2.double(), 1e-6)
    o32 = torch.empty(M, C, dtype=torch.float32, device="cuda")
    o16 = torch.empty(M, C, dtype=torch.float16, device="cuda")
    ops.layernorm_chain(parts.cuda(), bias.cuda(), g1.cuda(), b1.cuda(), 1e-5, o32, g2.cuda(), b2.cuda(), 1e-6, o16)
    check(o32, y, REL32, "chain: first norm")
    check(o16, z, REL16, "chain: second norm")
    ops.layernorm_chain(parts[:1].contiguous().cuda(), None, g1.cuda(), b1.cuda(), 1e-5, o32, g2.cuda(), b2.cuda(), 1e-6, o16)
    check(o32, F.layer_norm(parts[0].double(), (C,), g1.double(), b1.double(), 1e-5), REL32, "chain: no bias")


@pytest.mark.parametrize("layout,N,H,W,C,k,s,p", [(0, 2, 64, 96, 3, 7, 4, 3), (0, 1, 30, 50, 3, 7, 4, 3), (0, 2, 480, 480, 3, 7, 4, 3), (1, 2, 16, 24, 64, 3, 2, 1),
                                                  (1, 1, 16, 24, 32, 8, 8, 0), (1, 3, 9, 7, 160, 3, 2, 1),
                                                  (1, 2, 8, 12, 128, 4, 4, 0)])
def test_im2col_bit_exact(ops, layout, N, H, W, C, k, s, p):
    x = h16(synth.synth_array((N, C, H, W), 11))
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    kdim = k * k * C
    kpad = (kdim + 7) // 8 * 8
    out = torch.full((N * Ho * Wo, kpad), 7.0, dtype=torch.float16, device="cuda")
    src = x.cuda().contiguous() if layout == 0 else x.permute(0, 2, 3, 1).contiguous().cuda().half()
    ops.im2col(src, layout, N, H, W, C, k, s, p, out)
    cols = F.unfold(x, k, padding=p, stride=s)                               # (N, C*k*k, L), order (c, ky, kx)
    ref = cols.view(N, C, k, k, Ho * Wo).permute(0, 4, 2, 3, 1).reshape(N * Ho * Wo, kdim)
    assert torch.equal(out[:, :kdim].float().cpu(), ref)
    assert out[:, kdim:].abs().max().item() == 0 if kpad > kdim else True


@pytest.mark.parametrize("N,H,W,C", [(2, 16, 24, 256), (1, 7, 5, 128), (3, 4, 6, 1280)])
def test_dwconv3x3_gelu(ops, N, H, W, C):
    x = h16(synth.synth_array((N, C, H, W), 12))
    w = h16(synth.synth_array((C, 1, 3, 3), 13, scale=0.3))
    b = synth.synth_array((C,), 14, scale=0.1)
    ref = F.gelu(F.conv2d(x.double(), w.double(), b.double(), padding=1, groups=C)).permute(0, 2, 3, 1)
    out = torch.empty(N, H, W, C, dtype=torch.float16, device="cuda")
    ops.dwconv3x3_gelu(x.permute(0, 2, 3, 1).contiguous().cuda().half(), w.view(C, 9).t().contiguous().cuda().half(),
                       b.cuda(), out, N, H, W, C)
    check(out, ref, REL16, "dwconv+gelu")


@pytest.mark.parametrize("n,H,W,N", [(2, 64, 96, 64), (1, 480, 480, 64), (3, 37, 52, 32), (2, 480, 864, 64), (1, 20, 16, 64)])
def test_patch_embed_s1(ops, n, H, W, N):
    """conv 7x7 s4 p3 + bias -> LayerNorm -> LayerNorm in one kernel == the fp64 composition (mix_transformer.py:173-200, :154)."""
    assert ops.patch_embed_s1_supported(W, 3, 7, 4, 3, N) and not ops.patch_embed_s1_supported(W + 1, 3, 7, 4, 3, N)
    x = synth.synth_array((n, 3, H, W), 71)
    w = h16(synth.synth_array((N, 3, 7, 7), 72, scale=147 ** -0.5))
    b = synth.synth_array((N,), 73, scale=0.1)
    g1, e1 = 1 + synth.synth_array((N,), 74, scale=0.1), synth.synth_array((N,), 75, scale=0.1)
    g2, e2 = 1 + synth.synth_array((N,), 76, scale=0.1), synth.synth_array((N,), 77, scale=0.1)
    y = F.conv2d(h16(x).double(), w.double(), b.double(), stride=4, padding=3)      # the kernel rounds the frames to fp16 (MMA operand)
    Ho, Wo = y.shape[2:]
    y = y.permute(0, 2, 3, 1).reshape(n * Ho * Wo, N)
    y1 = F.layer_norm(y, (N,), g1.double(), e1.double(), 1e-5)
    y2 = F.layer_norm(y1, (N,), g2.double(), e2.double(), 1e-6)
    out32 = torch.empty(n * Ho * Wo, N, device="cuda")
    out16 = torch.empty(n * Ho * Wo, N, dtype=torch.float16, device="cuda")
    ops.patch_embed_s1(x.cuda().contiguous(), ops.patch_embed_s1_weight(w.cuda()), b.cuda(), g1.cuda(), e1.cuda(), 1e-5, g2.cuda(), e2.cuda(),
                       1e-6, out32, out16)
    check(out32, y1, 1e-3, "patch_embed LayerNorm 1")
    check(out16, y2, REL16, "patch_embed LayerNorm 2")


@pytest.mark.parametrize("n,H,W,HD,N,ln,alias", [(2, 16, 24, 256, 64, True, True), (1, 7, 5, 128, 64, False, True),
                                                 (3, 30, 33, 512, 128, True, True), (2, 120, 120, 256, 64, True, False),
                                                 (1, 9, 70, 256, 128, True, False), (2, 60, 60, 512, 128, False, True)])
def test_mixffn_tail(ops, n, H, W, HD, N, ln, alias):
    """x + fc2(GELU(dwconv(h))) (+ LayerNorm) in one kernel == the fp64 composition (mix_transformer.py:52-58, :84-88)."""
    assert ops.mixffn_tail_supported(N, HD) and not ops.mixffn_tail_supported(320, 1280)
    M = n * H * W
    h = h16(synth.synth_array((n, HD, H, W), 61))
    dw = h16(synth.synth_array((HD, 1, 3, 3), 62, scale=0.3))
    dwb = synth.synth_array((HD,), 63, scale=0.1)
    w2 = h16(synth.synth_array((N, HD), 64, scale=HD ** -0.5))
    b2 = synth.synth_array((N,), 65, scale=0.1)
    res = synth.synth_array((M, N), 66)
    g, be = 1 + synth.synth_array((N,), 67, scale=0.1), synth.synth_array((N,), 68, scale=0.1)
    act = F.gelu(F.conv2d(h.double(), dw.double(), dwb.double(), padding=1, groups=HD)).permute(0, 2, 3, 1).reshape(M, HD)
    act = act.half().double()                                   # the kernel rounds the activated map to fp16 (the MMA operand)
    x_ref = res.double() + act @ w2.double().t() + b2.double()
    res_d = res.cuda().float().contiguous()
    out32 = res_d if alias else torch.empty_like(res_d)
    lnout = torch.empty(M, N, dtype=torch.float16, device="cuda") if ln else None
    ops.mixffn_tail(h.permute(0, 2, 3, 1).reshape(M, HD).contiguous().cuda().half(), n, H, W, dw.view(HD, 9).t().contiguous().cuda().half(),
                    dwb.cuda(), w2.cuda().half(), b2.cuda(), res_d, out32, g.cuda() if ln else None, be.cuda() if ln else None, 1e-6, lnout)
    check(out32, x_ref, 1e-3, "mixffn_tail x")
    if ln:
        check(lnout, F.layer_norm(x_ref, (N,), g.double(), be.double(), 1e-6), REL16, "mixffn_tail LayerNorm")


@pytest.mark.parametrize("B,Nq,Nkv,heads,d", [(2, 384, 6, 1, 64), (2, 96, 6, 2, 64), (3, 200, 225, 5, 64),
                                              (2, 225, 225, 8, 64), (2, 100, 10, 8, 32), (1, 3600, 64, 8, 32),
                                              (2, 130, 100, 1, 32),
                                              # tcgen05 kernel: many query tiles per resident K/V, key blocks (N_kv > 256:
                                              # four S rounds per tile), head pairs (head_dim 32), odd head counts, tails
                                              (2, 3000, 225, 1, 64), (3, 700, 225, 2, 64), (1, 900, 405, 5, 64),
                                              (2, 300, 512, 2, 64), (1, 257, 257, 1, 64), (2, 150, 320, 3, 64),
                                              (2, 1000, 225, 2, 32), (1, 500, 225, 5, 32), (2, 225, 405, 8, 32),
                                              (1, 2000, 225, 1, 32), (1, 129, 512, 1, 32), (3, 65, 129, 3, 32)])
def test_mha_small_kv(ops, B, Nq, Nkv, heads, d):
    C = heads * d
    q = h16(synth.synth_array((B, Nq, C), 15))
    kv = h16(synth.synth_array((B, Nkv, 2 * C), 16))
    scale = d ** -0.5
    qh = q.double().view(B, Nq, heads, d).transpose(1, 2)
    kh = kv[..., :C].double().view(B, Nkv, heads, d).transpose(1, 2)
    vh = kv[..., C:].double().view(B, Nkv, heads, d).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1) @ vh).transpose(1, 2).reshape(B * Nq, C)
    out = torch.empty(B * Nq, C, dtype=torch.float16, device="cuda")
    kvd = kv.view(B * Nkv, 2 * C).cuda().half()
    ops.mha(q.view(B * Nq, C).cuda().half(), kvd[:, :C], kvd[:, C:], out, B, Nq, Nkv, heads, d, scale)
    check(out, ref, 2e-3, "mha")       # P is rounded to fp16 before P.V: 2 fp16 roundings on the path


# ----------------------------------------------------------------------------------- head decoder
def _head_sd(seed, chans):
    spec = {f"linear_c{i + 1}.proj.weight": (256, c) for i, c in enumerate(chans)}
    spec.update({f"linear_c{i + 1}.proj.bias": (256,) for i in range(4)})
    spec.update({"linear_fuse.conv.weight": (256, 1024, 1, 1), "linear_fuse.bn.weight": (256,),
                 "linear_fuse.bn.bias": (256,), "linear_fuse.bn.running_mean": (256,), "linear_fuse.bn.running_var": (256,)})
    return synth.synth_state_dict(spec, seed)


@pytest.mark.parametrize("t_perm", [0, 4])
def test_head_fuse_vs_oracle(ops, t_perm):
    """Folded MLP decoder + 2x2 mean == oracle head_mlp_decoder + resize(1/2) (cffm_head.py:102-133)."""
    chans, N, h, w = [64, 128, 320, 512], 4, 16, 24
    sd = _head_sd(17, chans)
    feats = [h16(synth.synth_array((N, c, h >> i, w >> i), 18 + i)) for i, c in enumerate(chans)]
    c_ref = O.head_mlp_decoder(sd, "", feats)                                      # (N,256,h,w)
    half_ref = O.resize(c_ref, (h // 2, w // 2))
    # host-side fold (same algebra as CFFMHead._build_plan)
    s = sd["linear_fuse.bn.weight"].double() / torch.sqrt(sd["linear_fuse.bn.running_var"].double() + 1e-5)
    shift = sd["linear_fuse.bn.bias"].double() - sd["linear_fuse.bn.running_mean"].double() * s
    Wf = sd["linear_fuse.conv.weight"].double().view(256, 1024)
    proj, sizes = [], []
    for i in range(4):
        slot = 3 - i
        Wfi = Wf[:, slot * 256:(slot + 1) * 256] * s[:, None]
        pw = (Wfi @ sd[f"linear_c{i + 1}.proj.weight"].double()).float()
        shift = shift + Wfi @ sd[f"linear_c{i + 1}.proj.bias"].double()
        x = feats[i].permute(0, 2, 3, 1).reshape(-1, chans[i]).cuda().half()
        p = torch.empty(x.shape[0], 256, dtype=torch.float16, device="cuda")
        ops.gemm(x, pw.cuda().half(), out16=p)
        proj.append(p)
        sizes.append((h >> i, w >> i))
    full = torch.empty(N * h * w, 256, dtype=torch.float16, device="cuda")
    h32 = torch.empty(N * (h // 2) * (w // 2), 256, dtype=torch.float32, device="cuda")
    h16_ = torch.empty(N * (h // 2) * (w // 2), 256, dtype=torch.float16, device="cuda")
    ops.head_fuse(proj, sizes, N, 256, t_perm, shift.float().cuda(), c_full=full, half32=h32, half16=h16_)
    if t_perm:                                     # clip-major in (n = b*T+t) -> frame-major out (t*B+b); B = 1 here
        order = [(n % t_perm) * (N // t_perm) + n // t_perm for n in range(N)]
        inv = torch.tensor(order).argsort()
        c_ref, half_ref = c_ref[inv], half_ref[inv]
    # fp16 projections feed a sum of 4 terms: 2e-3 of the output scale
    check(full.view(N, h, w, 256), c_ref.permute(0, 2, 3, 1), 2e-3, "_c")
    check(h32.view(N, h // 2, w // 2, 256), half_ref.permute(0, 2, 3, 1), 2e-3, "_c_further f32")
    check(h16_.view(N, h // 2, w // 2, 256), half_ref.permute(0, 2, 3, 1), 2e-3, "_c_further f16")


# ------------------------------------------------------------------------------------------- CFFA
def _apron(x, Hp, Wp, e=3):
    """(B,Hp,Wp,C) -> (B,Hp+2e,Wp+2e,C): position (Y,X) holds x[(Y-e) mod Hp, (X-e) mod Wp]."""
    yi = (torch.arange(Hp + 2 * e) - e) % Hp
    xi = (torch.arange(Wp + 2 * e) - e) % Wp
    return x[:, yi][:, :, xi]


def _block_sd(seed):
    spec = {"norm1.weight": (256,), "norm1.bias": (256,), "pool_layers.0.weight": (1, 49), "pool_layers.0.bias": (1,),
            "pool_layers_clips.0.weight": (1, 49), "pool_layers_clips.0.bias": (1,),
            "pool_layers_clips.1.weight": (1, 9), "pool_layers_clips.1.bias": (1,),
            "pool_layers_clips.2.weight": (1, 4), "pool_layers_clips.2.bias": (1,)}
    return synth.synth_state_dict({"blk." + k: v for k, v in spec.items()}, seed)


@pytest.mark.parametrize("B,H,W", [(1, 20, 25), (2, 14, 21), (1, 60, 60)])
def test_cffa_norm_and_pool_vs_oracle(ops, B, H, W):
    """LN + zero pad + (bilinear o fc-pool) == oracle cffa_assemble (cffm_transformer.py:713-805)."""
    T, C = 4, 256
    sd = _block_sd(21)
    x = synth.synth_array((B, T, H, W, C), 22)                                    # clip-major for the oracle
    xn_ref = F.layer_norm(x, (C,), sd["blk.norm1.weight"], sd["blk.norm1.bias"], 1e-5)
    Hp, Wp = (H + 6) // 7 * 7, (W + 6) // 7 * 7
    xn_pad = F.pad(xn_ref, (0, 0, 0, Wp - W, 0, Hp - H))
    xfm = x.transpose(0, 1).contiguous().cuda()                                   # frame-major for the kernels
    xn = torch.empty(T * B * H * W, C, dtype=torch.float16, device="cuda")
    xt_pad = torch.full((B * (Hp + 6) * (Wp + 6), C), float("nan"), dtype=torch.float16, device="cuda")   # every row must be written
    ops.cffa_norm(xfm, sd["blk.norm1.weight"].cuda(), sd["blk.norm1.bias"].cuda(), 1e-5, xn, xt_pad, B, T, H, W, Hp, Wp, C)
    check(xn.view(T, B, H, W, C), xn_ref.transpose(0, 1), REL16, "norm1")
    # target map in the CFM kernel's layout: zero pad AFTER the norm, then a 3-wide cyclic apron (= torch.roll wrap-around)
    apron = _apron(xn_pad[:, -1], Hp, Wp)
    got = xt_pad.view(B, Hp + 6, Wp + 6, C)
    check(got, apron, REL16, "padded target + apron")
    assert got[:, 3 + H:3 + Hp, 3:3 + Wp].abs().sum().item() == 0 and got[:, 3:3 + Hp, 3 + W:3 + Wp].abs().sum().item() == 0
    assert torch.equal(got[:, :3], got[:, Hp:Hp + 3]) and torch.equal(got[:, :, Wp + 3:], got[:, :, 3:6])   # wrap copies are exact
    # pooling on the kernel's own fp16 LN output (isolates the pool kernel)
    xn16 = xn.view(T, B, H, W, C).float().cpu().transpose(0, 1)
    pooled_ref = O.cffa_assemble(sd, "blk", F.pad(xn16, (0, 0, 0, Wp - W, 0, Hp - H)))
    pools = ["pool_layers.0", "pool_layers_clips.0", "pool_layers_clips.1", "pool_layers_clips.2"]
    pw = torch.cat([sd[f"blk.{p}.weight"].reshape(-1) for p in pools]).cuda()
    pb = torch.cat([sd[f"blk.{p}.bias"].reshape(-1) for p in pools]).cuda()
    nW = (Hp // 7) * (Wp // 7)
    pooled = torch.empty(B * 15 * nW, C, dtype=torch.float16, device="cuda")
    ops.cffa_pool(xn, B, T, H, W, C, pw, pb, pooled)
    got = pooled.view(B, 15 * nW, C)
    off = 0
    for lvl, pr in enumerate(pooled_ref):
        n = pr.shape[1] * pr.shape[2]
        check(got[:, off:off + n], pr.reshape(B, n, C), REL16, f"pooled level {lvl}")
        off += n
    assert off == 15 * nW
    # the two-part form used by the head (reference levels ahead of time, target level per block) fills the same buffer
    split = torch.full_like(pooled, float("nan"))
    xn5 = xn.view(T, B, H, W, C)
    ops.cffa_pool_part(xn5[:T - 1].contiguous(), B, 1, H, W, C, pw, pb, split)
    assert torch.isnan(split.view(B, 15 * nW, C)[:, :nW]).all()                  # target rows untouched by the reference part
    ops.cffa_pool_part(xn5[T - 1].contiguous(), B, 0, H, W, C, pw, pb, split)
    assert torch.equal(split, pooled)                                             # bit-identical to the one-pass kernel


# ------------------------------------------------------------------------------------ CFM attention
def _kernel_row_of_reference_key(lay):
    """Row of the CFM kernel's key tile that holds reference key n (column n of the reference's logits, 0..288):
    own window and ring -> the 13 x 13 halo (a ring key listed twice maps to one row), pooled windows -> their blocks."""
    rows = []
    for n in range(49):
        rows.append((n // 7 + 3) * 13 + n % 7 + 3)
    for dy, dx in O.ring_offsets():
        rows.append((dy + 3) * 13 + dx + 3)
    for blk, cnt in zip(lay["rows"][1:], (25, 49, 25, 9)):
        rows += [blk + m for m in range(cnt)]
    assert len(rows) == 289 and len(set(rows)) == 277
    return torch.tensor(rows)


@pytest.mark.parametrize("B,H,W", [(1, 21, 28), (2, 60, 60), (1, 14, 14), (1, 7, 12)])
def test_cfm_key_assembling_bit_exact(ops, golden_dir, B, H, W):
    """The K and V tiles the CFM kernel's TMA boxes assemble == the key sequence the reference's roll /
    window_partition / valid_ind_rolled / nn.Unfold / cat builds (cffm_transformer.py:378-522), bit for bit:
    every source position carries its own (clip, level, y, x) code, the tiles are dumped and compared with the oracle's
    source table (itself pinned to the reference-generated golden key codes, 63x63 included)."""
    import os
    C = 256
    Hp, Wp = (H + 6) // 7 * 7, (W + 6) // 7 * 7
    nWh, nWw = Hp // 7, Wp // 7
    nW = nWh * nWw
    lay = ops.cfm_layout()
    ch = torch.arange(C)

    def coded(level, b, ys, xs, is_v):
        """[len(ys), C] codes: channel c carries (y, x, 10 level + is_v, 16 b + c // 32) for c % 4 = 0..3; level 0 uses
        1-based y so that a real row is never all zero."""
        out = torch.empty(len(ys), C)
        out[:, ch % 4 == 0] = (ys[:, None] + 1).float().expand(-1, C // 4)
        out[:, ch % 4 == 1] = (xs[:, None] + 1).float().expand(-1, C // 4)
        out[:, ch % 4 == 2] = float(10 * level + is_v + 1)
        out[:, ch % 4 == 3] = (16 * b + ch[ch % 4 == 3] // 32 + 1).float()[None].expand(len(ys), -1)
        return out

    Ha, Wa = Hp + 6, Wp + 6
    qkv_a = torch.zeros(B, Ha, Wa, 3 * C)
    yy, xx = torch.meshgrid(torch.arange(Ha), torch.arange(Wa), indexing="ij")
    ys0, xs0 = ((yy - 3) % Hp).reshape(-1), ((xx - 3) % Wp).reshape(-1)
    sizes = [(nWh, nWw), (nWh, nWw), (2 * nWh, 2 * nWw), (3 * nWh, 3 * nWw)]
    kvp = torch.zeros(B, 15 * nW, 2 * C)
    for b in range(B):
        qkv_a[b, :, :, C:2 * C] = coded(0, b, ys0, xs0, 0).view(Ha, Wa, C)
        qkv_a[b, :, :, 2 * C:] = coded(0, b, ys0, xs0, 1).view(Ha, Wa, C)
        off = 0
        for l, (gh, gw) in enumerate(sizes):
            gy, gx = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
            kvp[b, off:off + gh * gw, :C] = coded(l + 1, b, gy.reshape(-1), gx.reshape(-1), 0)
            kvp[b, off:off + gh * gw, C:] = coded(l + 1, b, gy.reshape(-1), gx.reshape(-1), 1)
            off += gh * gw
    assert qkv_a.max() < 2048 and kvp.max() < 2048                                # exact in fp16
    bias = torch.zeros(8, 49, lay["pitch"], dtype=torch.float16, device="cuda")
    out = torch.empty(B * H * W, C, dtype=torch.float16, device="cuda")
    dump = torch.full((4, B * nW, 2, lay["npad"], 64), float("nan"), dtype=torch.float16, device="cuda")
    ops.cfm_attention(qkv_a.view(-1, 3 * C).cuda().half(), kvp.view(-1, 2 * C).cuda().half(), bias, out, B, H, W, C, 8,
                      32 ** -0.5, dump=dump)
    got = dump.float().cpu()                                                       # [hp, item, K|V, row, 64]
    assert not torch.isnan(got).any()
    lev, ys, xs = O.key_source_table(Hp, Wp)                                       # (nW, 289), reference key order
    rows = _kernel_row_of_reference_key(lay)
    used = torch.zeros(lay["npad"], dtype=torch.bool)
    used[rows] = True
    c64 = torch.arange(64)
    for hp in range(4):
        for b in range(B):
            for is_v in (0, 1):
                tile = got[hp, b * nW:(b + 1) * nW, is_v]                          # (nW, npad, 64)
                ref = torch.zeros(nW, 289, 64)
                ok = (ys >= 0)[..., None].float()
                ref[..., c64 % 4 == 0] = ((ys + 1)[..., None] * ok).expand(-1, -1, 16)
                ref[..., c64 % 4 == 1] = ((xs + 1)[..., None] * ok).expand(-1, -1, 16)
                ref[..., c64 % 4 == 2] = (10 * lev + is_v + 1)[..., None].float().expand(-1, -1, 16) * ok
                ref[..., c64 % 4 == 3] = (16 * b + (hp * 64 + c64[c64 % 4 == 3]) // 32 + 1).float()[None, None] * ok
                assert torch.equal(tile[:, rows], ref), (hp, b, is_v)              # every reference key, duplicates included
                assert tile[:, ~used].abs().sum() == 0                             # rows no reference key maps to stay zero
    t = np.load(os.path.join(golden_dir, "index_tables.npz"))
    if f"key_code_{Hp}x{Wp}" in t:                                                 # the reference's own key codes
        k0 = got[0, :nW, 0][:, rows]                                               # clip 0, head pair 0, K tile
        y1, x1, lv = k0[..., 0].long(), k0[..., 1].long(), (k0[..., 2].long() - 1) // 10
        code = torch.where(y1 == 0, torch.zeros_like(y1), lv * 10000 + (y1 - 1) * 100 + (x1 - 1) + 1)
        assert np.array_equal(code.to(torch.int32).numpy(), t[f"key_code_{Hp}x{Wp}"])


def _attn_sd(golden_dir, seed):
    import json, os
    with open(os.path.join(golden_dir, "state_dict_spec.json")) as f:
        spec = json.load(f)["b1"]
    pre = "decode_head.decoder_focal.blocks.0.attn."
    return {"attn." + k[len(pre):]: synth.synth_tensor("attn." + k[len(pre):], s, seed)
            for k, s in spec.items() if k.startswith(pre) and not synth.is_derived_buffer(k)}


@pytest.mark.parametrize("B,H,W", [(1, 21, 28), (2, 20, 25), (1, 60, 60), (2, 60, 108), (1, 7, 7)])
def test_cfm_attention_vs_oracle(ops, golden_dir, B, H, W):
    """qkv GEMMs + CFM attention kernel == oracle cfm_attention before proj (cffm_transformer.py:364-601)."""
    C, heads = 256, 8
    Hp, Wp = (H + 6) // 7 * 7, (W + 6) // 7 * 7
    nWh, nWw = Hp // 7, Wp // 7
    nW = nWh * nWw
    sd = _attn_sd(golden_dir, 3)
    sd["attn.qkv.weight"], sd["attn.qkv.bias"] = h16(sd["attn.qkv.weight"]), sd["attn.qkv.bias"]
    sd["attn.proj.weight"] = torch.eye(C)                                          # compare pre-proj
    sd["attn.proj.bias"] = torch.zeros(C)
    xt = h16(synth.synth_array((B, Hp, Wp, C), 11))
    xt[:, H:] = 0; xt[:, :, W:] = 0                                               # zero pad after LN
    pooled = [h16(synth.synth_array((B, nWh * l, nWw * l, C), 12 + i)) for i, l in enumerate((1, 1, 2, 3))]
    ref = O.cfm_attention(sd, "attn", xt, pooled)                                  # (B*nW, 49, C)
    ref = ref.view(B, nWh, nWw, 7, 7, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)[:, :H, :W]
    qw, qb = sd["attn.qkv.weight"].cuda().half(), sd["attn.qkv.bias"].cuda()
    xa = _apron(xt, Hp, Wp).reshape(-1, C).cuda().half()                           # what cffa_norm hands to the QKV GEMM
    qkv_a = torch.empty(xa.shape[0], 3 * C, dtype=torch.float16, device="cuda")
    ops.gemm(xa, qw, bias=qb, out16=qkv_a)
    pl = torch.cat([p.reshape(B, -1, C) for p in pooled], 1).reshape(-1, C).cuda().half()
    kvp = torch.empty(B * 15 * nW, 2 * C, dtype=torch.float16, device="cuda")
    ops.gemm(pl, qw[C:], bias=qb[C:], out16=kvp)
    scale = (C // heads) ** -0.5
    bias = tb.assemble_bias_tc(sd["attn.relative_position_bias_table"].cuda(),
                               sd["attn.relative_position_bias_table_to_neighbors"].cuda(),
                               sd["attn.relative_position_bias_table_to_windows.0"].cuda(),
                               [sd[f"attn.relative_position_bias_table_to_windows_clips.{k}"].cuda() for k in range(3)],
                               scale, ops.cfm_layout())
    out = torch.full((B * H * W, C), float("nan"), dtype=torch.float16, device="cuda")
    ops.cfm_attention(qkv_a, kvp, bias, out, B, H, W, C, heads, scale)
    # q, k, v are rounded to fp16 by the GEMM, P by the kernel
    e = check(out.view(B, H, W, C), ref, CFM_TOL, "cfm attention")
    print(f"cfm attention {B}x{H}x{W}: rel err {e:.2e}")


# ------------------------------------------------------------------------------------------- tails
@pytest.mark.parametrize("f32in", [True, False])
@pytest.mark.parametrize("B,h,w,Ho,Wo", [(2, 8, 12, 16, 24), (1, 60, 60, 120, 120), (2, 16, 24, 16, 24), (1, 5, 7, 13, 9)])
def test_resize_nhwc_to_nchw(ops, f32in, B, h, w, Ho, Wo):
    ncls, ldc = 124, 128
    x = synth.synth_array((B, h, w, ldc), 31)
    if not f32in:
        x = h16(x)
    ref = F.interpolate(x[..., :ncls].permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=False)
    out = torch.empty(B, ncls, Ho, Wo, dtype=torch.float32, device="cuda")
    ops.resize_nhwc_to_nchw(x.view(-1, ldc).cuda() if f32in else x.view(-1, ldc).cuda().half(), ncls, out, B, h, w, Ho, Wo)
    check(out, ref, 1e-5, "resize nhwc->nchw")


@pytest.mark.parametrize("B,C,h,w,Ho,Wo", [(2, 124, 16, 24, 64, 96), (1, 124, 120, 120, 480, 480), (1, 7, 9, 5, 20, 33)])
def test_resize_nchw_softmax_argmax(ops, B, C, h, w, Ho, Wo):
    x = synth.synth_array((B, C, h, w), 32)
    up = F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=False)
    out = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device="cuda")
    ops.resize_nchw(x.cuda(), out)
    check(out, up, 1e-5, "resize nchw")
    sm = torch.empty_like(out)
    ops.softmax_nchw(out, sm)
    check(sm, torch.softmax(up, 1), 1e-5, "softmax")
    labels = torch.empty(B, Ho, Wo, dtype=torch.int64, device="cuda")
    ops.resize_argmax(x.cuda(), labels, B, C, h, w, Ho, Wo)
    ref = torch.softmax(up, 1).argmax(1)                                       # reference order: resize, softmax, argmax
    agree = (labels.cpu() == ref).float().mean().item()
    assert agree >= 0.9999, agree                                               # fp32 near-ties may flip an fma rounding
    top2 = up.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])[labels.cpu() != ref]
    assert margin.numel() == 0 or margin.max().item() < 1e-5


@pytest.mark.parametrize("B,h,w,Hm,Wm,Ho,Wo", [(2, 8, 12, 16, 24, 64, 96), (1, 60, 60, 120, 120, 480, 480),
                                               (2, 16, 24, 16, 24, 64, 96), (1, 7, 9, 13, 18, 50, 70)])
def test_upsample2_argmax_vs_two_resizes(ops, B, h, w, Hm, Wm, Ho, Wo):
    """Fused (x2 resize, x4 resize, argmax) == resize -> resize -> softmax -> argmax (cffm_head.py:149,
    encoder_decoder.py:373-377,542,564)."""
    ncls, ldc = 124, 128
    x = synth.synth_array((B, h, w, ldc), 33)
    assert ops.upsample2_argmax_supported(Hm, Wm, Ho, Wo)
    mid = F.interpolate(x[..., :ncls].permute(0, 3, 1, 2), size=(Hm, Wm), mode="bilinear", align_corners=False)
    up = F.interpolate(mid, size=(Ho, Wo), mode="bilinear", align_corners=False)
    ref = torch.softmax(up, 1).argmax(1)
    labels = torch.empty(B, Ho, Wo, dtype=torch.int64, device="cuda")
    ops.upsample2_argmax(x.view(-1, ldc).cuda(), ncls, labels, B, h, w, Hm, Wm, Ho, Wo)
    got = labels.cpu()
    agree = (got == ref).float().mean().item()
    assert agree >= 0.9999, agree
    top2 = up.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])[got != ref]
    assert margin.numel() == 0 or margin.max().item() < 1e-5          # only fp32 near-ties may flip
    assert not ops.upsample2_argmax_supported(120, 120, 130, 130)       # weak upsampling: unfused path is used


def test_gelu_epilogue_accuracy(ops):
    """The one-MUFU erf of the GELU epilogue is exact to 3e-7 (tools/fit_erf.py): checked through a GEMM with
    an identity weight over a dense sweep of pre-activations."""
    M, K = 4096, 64
    xs = torch.linspace(-9, 9, M * K).view(M, K)
    a = h16(xs).cuda().half()
    w = torch.eye(K, dtype=torch.float16, device="cuda")
    out = torch.empty(M, K, dtype=torch.float32, device="cuda")
    ops.gemm(a, w, out32=out, act=ops.ACT_GELU)
    ref = F.gelu(a.double().cpu())
    assert (out.cpu().double() - ref).abs().max().item() < 2e-6


@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (300, 128, 256), (4100, 32, 32), (129, 64, 256), (513, 128, 512), (77, 8, 64)])
def test_gemm_with_fused_layernorm(ops, M, N, K):
    """x = A W^T + bias + residual (in place) and LayerNorm(x) in fp16 from ONE kernel == GEMM then LayerNorm."""
    a = h16(synth.synth_array((M, K), 41)).cuda()
    w = h16(synth.synth_array((N, K), 42, scale=K ** -0.5)).cuda()
    bias = synth.synth_array((N,), 43).cuda()
    res = synth.synth_array((M, N), 44, scale=2.0).cuda()
    g, b = (synth.synth_array((N,), 45) * 0.1 + 1).cuda(), (synth.synth_array((N,), 46) * 0.1).cuda()
    x_ref = a.double() @ w.double().t() + bias.double() + res.double()
    for eps in (1e-5, 1e-6):
        ln_ref = F.layer_norm(x_ref, (N,), g.double(), b.double(), eps)
        x = res.clone()
        ln = torch.empty(M, N, dtype=torch.float16, device="cuda")
        ops.gemm_ln(a.half(), w.half(), bias, x, x, g, b, eps, ln)
        check(x, x_ref, REL32, "fused gemm+ln: x")
        check(ln, ln_ref, REL16, "fused gemm+ln: LayerNorm(x)")
        ln2 = torch.empty_like(ln)
        ops.gemm_ln(a.half(), w.half(), bias, res, None, g, b, eps, ln2)          # no fp32 output requested
        assert torch.equal(ln, ln2)
    assert ops.gemm_ln_supported(128) and not ops.gemm_ln_supported(320)


@pytest.mark.parametrize("M,N,K", [(1000, 64, 152), (300, 128, 576), (4100, 32, 152), (129, 64, 256), (77, 8, 64)])
def test_gemm_with_two_chained_layernorms(ops, M, N, K):
    """y = LN1(A W^T + bias) in fp32 and LN2(y) in fp16 from ONE kernel (patch-embed conv + norm + first norm1)."""
    a = h16(synth.synth_array((M, K), 51)).cuda()
    w = h16(synth.synth_array((N, K), 52, scale=K ** -0.5)).cuda()
    bias = synth.synth_array((N,), 53).cuda()
    g1, b1 = (synth.synth_array((N,), 54) * 0.1 + 1).cuda(), (synth.synth_array((N,), 55) * 0.1).cuda()
    g2, b2 = (synth.synth_array((N,), 56) * 0.1 + 1).cuda(), (synth.synth_array((N,), 57) * 0.1).cuda()
    x = a.double() @ w.double().t() + bias.double()
    y_ref = F.layer_norm(x, (N,), g1.double(), b1.double(), 1e-5)
    z_ref = F.layer_norm(y_ref, (N,), g2.double(), b2.double(), 1e-6)
    y = torch.empty(M, N, dtype=torch.float32, device="cuda")
    z = torch.empty(M, N, dtype=torch.float16, device="cuda")
    ops.gemm_ln_chain(a.half(), w.half(), bias, y, g1, b1, 1e-5, g2, b2, 1e-6, z)
    check(y, y_ref, 2e-4, "chain: first norm")        # LN of fp16-product sums: the GEMM's own rounding passes through rstd
    check(z, z_ref, REL16, "chain: second norm")


@pytest.mark.parametrize("M,N,K", [(1800, 64, 4096), (450, 512, 2880), (1800, 128, 2048), (225, 32, 2048), (300, 64, 512)])
def test_splitk_gemm_and_summing_layernorm(ops, M, N, K):
    """Split-K partial products + LayerNorm(sum + bias) == conv-as-GEMM followed by LayerNorm (mix_transformer.py:101-103)."""
    a = h16(synth.synth_array((M, K), 51)).cuda()
    w = h16(synth.synth_array((N, K), 52, scale=K ** -0.5)).cuda()
    bias = synth.synth_array((N,), 53).cuda()
    g, b = (synth.synth_array((N,), 54) * 0.1 + 1).cuda(), (synth.synth_array((N,), 55) * 0.1).cuda()
    S = max(ops.splitk_plan(M, N, K), 2)
    assert ops.splitk_plan(M, N, K) == ops.splitk_plan(7 * M, N, K)      # batch-size independent
    parts = torch.full((S, M, N), float("nan"), dtype=torch.float32, device="cuda")
    ops.gemm_splitk(a.half(), w.half(), parts)
    x_ref = a.double() @ w.double().t()
    check(parts.sum(0), x_ref, REL32, f"split-K partial sums (S={S})")
    ln_ref = F.layer_norm(x_ref + bias.double(), (N,), g.double(), b.double(), 1e-5)
    o16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
    o32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.layernorm_sum(parts, bias, g, b, 1e-5, out16=o16, out32=o32)
    check(o32, ln_ref, REL32, "layernorm_sum f32")
    check(o16, ln_ref, REL16, "layernorm_sum f16")
    o32b = torch.empty_like(o32)
    ops.layernorm_sum(parts, bias, g, b, 1e-5, out32=o32b)
    assert torch.equal(o32, o32b)                                  # fixed summation order: deterministic


# ------------------------------------------------------------------------------------ implicit-GEMM convolutions
def _ln(x, g, b, eps):
    return F.layer_norm(x, (x.shape[-1],), g, b, eps)


@pytest.mark.parametrize("n,H,W,C,k,s,p,Nout,chain", [
    (2, 24, 30, 64, 3, 2, 1, 128, True),      # OverlapPatchEmbed k3 s2 p1, both LayerNorms chained, short last tile
    (8, 120, 120, 64, 3, 2, 1, 128, True),    # stage-2 patch embed at 480x480
    (2, 16, 24, 128, 4, 4, 0, 128, False),    # Attention.sr, kernel = stride
    (1, 7, 9, 64, 3, 2, 1, 64, False),        # odd sizes, one partial tile
    (1, 60, 216, 64, 3, 2, 1, 128, True),     # 480x864 input: 108-wide output rows, one row per tile
])
def test_conv_gemm_ln_vs_conv2d(ops, n, H, W, C, k, s, p, Nout, chain):
    """Implicit-GEMM convolution (TMA boxes walking the NHWC image with the conv stride, zero fill = zero padding) + bias +
    fused LayerNorm(s) == F.conv2d + F.layer_norm in fp32 on the same fp16-representable operands."""
    x = h16(synth.synth_array((n, H, W, C), 61))
    w = h16(synth.synth_array((Nout, C, k, k), 62, scale=(k * k * C) ** -0.5))
    bias = synth.synth_array((Nout,), 63)
    g1, b1 = 1 + 0.1 * synth.synth_array((Nout,), 64), 0.1 * synth.synth_array((Nout,), 65)
    g2, b2 = 1 + 0.1 * synth.synth_array((Nout,), 66), 0.1 * synth.synth_array((Nout,), 67)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), stride=s, padding=p).permute(0, 2, 3, 1).float()
    Ho, Wo = y.shape[1:3]
    M = n * Ho * Wo
    y = y.reshape(M, Nout)
    wk = w.permute(0, 2, 3, 1).reshape(Nout, k * k * C).contiguous().cuda().half()          # (ky, kx, c) order
    out32 = torch.full((M, Nout), float("nan"), device="cuda")
    ln16 = torch.full((M, Nout), float("nan"), dtype=torch.float16, device="cuda")
    cu = lambda t: t.cuda()
    if chain:
        ops.conv_gemm_ln(x.cuda().half(), n, H, W, C, k, s, p, wk, cu(bias), out32, cu(g1), cu(b1), 1e-5, ln16, cu(g2), cu(b2), 1e-6)
        y1 = _ln(y, g1, b1, 1e-5)
        check(out32, y1, REL32, "LN1(conv) fp32")
        check(ln16, _ln(y1, g2, b2, 1e-6), REL16, "LN2(LN1(conv)) fp16")
    else:
        ops.conv_gemm_ln(x.cuda().half(), n, H, W, C, k, s, p, wk, cu(bias), out32, cu(g1), cu(b1), 1e-5, ln16)
        check(out32, y, REL32, "conv fp32")
        check(ln16, _ln(y, g1, b1, 1e-5), REL16, "LN(conv) fp16")
    assert not torch.isnan(out32).any() and not torch.isnan(ln16.float()).any()            # every row of every (short) tile written


@pytest.mark.parametrize("n,H,W,C,k,s,p,Nout", [
    (8, 120, 120, 64, 8, 8, 0, 64),           # stage-1 Attention.sr at 480x480: K = 4096, 8 splits, 15-row images in 2 tiles
    (2, 30, 30, 320, 3, 2, 1, 512),           # stage-4 patch embed: K = 2880, 5 splits
    (1, 60, 108, 128, 4, 4, 0, 128),          # stage-2 sr of a 480x864 input
])
def test_conv_gemm_splitk_vs_conv2d(ops, n, H, W, C, k, s, p, Nout):
    x = h16(synth.synth_array((n, H, W, C), 71))
    w = h16(synth.synth_array((Nout, C, k, k), 72, scale=(k * k * C) ** -0.5))
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, stride=s, padding=p).permute(0, 2, 3, 1).float()
    M = y.shape[0] * y.shape[1] * y.shape[2]
    wk = w.permute(0, 2, 3, 1).reshape(Nout, k * k * C).contiguous().cuda().half()
    S = ops.splitk_plan(M, Nout, k * k * C)
    assert S > 1
    part = torch.full((S, M, Nout), float("nan"), device="cuda")
    ops.conv_gemm_splitk(x.cuda().half(), n, H, W, C, k, s, p, wk, part)
    check(part.sum(0), y.reshape(M, Nout), REL32, "sum of split-K partials")
