"""Tensor-level wrappers over the C ABI (include/cffm_b200.h).

PyTorch is used for device memory and the current stream only; every function below hands raw
device pointers to libcffm_b200.so.  Each wrapper validates dtype / device / contiguity on the
host so that a bad call fails here, with a Python traceback, instead of inside a kernel.
"""
import torch

from . import _abi
from ._abi import ACT_GELU, ACT_NONE, ACT_RELU, GEMM_CHECK, GEMM_TCGEN05  # noqa: F401

_H, _F = torch.float16, torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name, rows2d=True):
    if not t.is_cuda:
        raise _abi.CffmError(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise _abi.CffmError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise _abi.CffmError(f"{name}: innermost dimension must be contiguous")


def _ld(t):
    """Row stride (elements) of a 2-D row-major view; rows may be strided (column slices)."""
    assert t.dim() == 2
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm(a, w, bias=None, residual=None, out16=None, out32=None, act=ACT_NONE, impl=GEMM_TCGEN05):
    """out = act(a @ w.T + bias) (+ residual).  a [M,K] fp16, w [N,K] fp16, bias fp32 [N],
    residual fp32 [M,N] (may alias out32).  out16 / out32 are preallocated [M,N] views."""
    _chk(a, _H, "gemm.a"); _chk(w, _H, "gemm.w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    for t, d, n in ((bias, _F, "bias"), (residual, _F, "residual"), (out16, _H, "out16"), (out32, _F, "out32")):
        if t is not None:
            _chk(t, d, "gemm." + n)
    if bias is not None:
        assert bias.numel() == N
    for t in (residual, out16, out32):
        if t is not None:
            assert tuple(t.shape) == (M, N), (tuple(t.shape), M, N)
    _abi.call("cffm_gemm_f16", _ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(bias), _ptr(residual),
              _ld(residual) if residual is not None else 0, _ptr(out16), _ld(out16) if out16 is not None else 0,
              _ptr(out32), _ld(out32) if out32 is not None else 0, M, N, K, act, impl, _stream())


def gemm_ln(a, w, bias, residual, out32, ln_gamma, ln_beta, ln_eps, ln_out16):
    """x = a @ w.T + bias (+ residual) -> out32 (optional, may alias residual); LayerNorm(x) -> ln_out16.
    Needs N <= 128 (``gemm_ln_supported``)."""
    _chk(a, _H, "gemm_ln.a"); _chk(w, _H, "gemm_ln.w"); _chk(ln_out16, _H, "gemm_ln.ln_out16")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and tuple(ln_out16.shape) == (M, N) and ln_gamma.numel() == N == ln_beta.numel()
    for t, n in ((bias, "bias"), (residual, "residual"), (out32, "out32"), (ln_gamma, "gamma"), (ln_beta, "beta")):
        if t is not None:
            _chk(t, _F, "gemm_ln." + n)
    for t in (residual, out32):
        if t is not None:
            assert tuple(t.shape) == (M, N)
    _abi.call("cffm_gemm_f16_ln", _ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(bias), _ptr(residual),
              _ld(residual) if residual is not None else 0, _ptr(out32), _ld(out32) if out32 is not None else 0,
              _ptr(ln_gamma), _ptr(ln_beta), float(ln_eps), _ptr(ln_out16), _ld(ln_out16), M, N, K, _stream())


def gemm_ln_chain(a, w, bias, out32, g1, b1, eps1, g2, b2, eps2, ln_out16):
    """y = LayerNorm(a @ w.T + bias; g1, b1) -> out32;  LayerNorm(y; g2, b2) -> ln_out16.  N <= 128."""
    _chk(a, _H, "gemm_ln_chain.a"); _chk(w, _H, "gemm_ln_chain.w"); _chk(out32, _F, "gemm_ln_chain.out32")
    _chk(ln_out16, _H, "gemm_ln_chain.ln_out16")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and tuple(out32.shape) == (M, N) == tuple(ln_out16.shape) and N <= 128
    for t in (bias, g1, b1, g2, b2):
        if t is not None:
            _chk(t, _F, "gemm_ln_chain.vector"); assert t.numel() == N
    _abi.call("cffm_gemm_f16_ln_chain", _ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(bias), _ptr(out32), _ld(out32), _ptr(g1), _ptr(b1),
              float(eps1), _ptr(g2), _ptr(b2), float(eps2), _ptr(ln_out16), _ld(ln_out16), M, N, K, _stream())


def splitk_plan(M, N, K):
    """Number of K splits that fills the SMs for a few-tile / long-K GEMM (1 = do not split)."""
    return int(_abi.load().cffm_splitk_plan(M, N, K))


def gemm_splitk(a, w, partials):
    """partials[s] = a[:, ks] @ w[:, ks].T for the s-th contiguous K range; partials fp32 [S, M, N] contiguous."""
    _chk(a, _H, "gemm_splitk.a"); _chk(w, _H, "gemm_splitk.w"); _chk(partials, _F, "gemm_splitk.partials")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and partials.is_contiguous() and tuple(partials.shape[1:]) == (M, N)
    _abi.call("cffm_gemm_f16_splitk", _ptr(a), _ld(a), _ptr(w), _ld(w), _ptr(partials), M, N, K, partials.shape[0], _stream())


def conv_gemm_supported(C, Wo, stride):
    """The implicit-GEMM convolution takes NHWC images with C % 64 == 0 whose output rows fit one TMA box."""
    return C % 64 == 0 and Wo <= 128 and Wo * stride <= 256


def conv_gemm_ln(x, n, H, W, C, k, stride, pad, w, bias, out32, g1, b1, eps1, ln_out16, g2=None, b2=None, eps2=0.0):
    """y = conv(x) + bias; LayerNorm(y) -> ln_out16 (y -> out32 when given); with g2: out32 = LN1(y), ln_out16 = LN2(out32).
    x fp16 NHWC [n,H,W,C] (any contiguous tensor with n*H*W*C elements), w fp16 [Nout, k*k*C] in (ky,kx,c) order."""
    _chk(x, _H, "conv_gemm_ln.x"); _chk(w, _H, "conv_gemm_ln.w"); _chk(ln_out16, _H, "conv_gemm_ln.ln_out16")
    assert x.is_contiguous() and x.numel() == n * H * W * C and w.shape[1] >= k * k * C
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    Nout = w.shape[0]
    assert tuple(ln_out16.shape) == (n * Ho * Wo, Nout)
    for t in (bias, out32, g1, b1, g2, b2):
        if t is not None:
            _chk(t, _F, "conv_gemm_ln.f32 operand")
    if out32 is not None:
        assert tuple(out32.shape) == (n * Ho * Wo, Nout)
    _abi.call("cffm_conv_gemm_f16_ln", _ptr(x), n, H, W, C, k, stride, pad, _ptr(w), _ld(w), _ptr(bias), _ptr(out32),
              _ld(out32) if out32 is not None else 0, _ptr(g1), _ptr(b1), float(eps1), _ptr(g2), _ptr(b2), float(eps2),
              _ptr(ln_out16), _ld(ln_out16), Nout, _stream())


def conv_gemm_splitk(x, n, H, W, C, k, stride, pad, w, partials):
    """partials[s] = the s-th K range of conv(x); partials fp32 [S, n*Ho*Wo, Nout] contiguous."""
    _chk(x, _H, "conv_gemm_splitk.x"); _chk(w, _H, "conv_gemm_splitk.w"); _chk(partials, _F, "conv_gemm_splitk.partials")
    assert x.is_contiguous() and x.numel() == n * H * W * C and w.shape[1] >= k * k * C and partials.is_contiguous()
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    assert tuple(partials.shape[1:]) == (n * Ho * Wo, w.shape[0])
    _abi.call("cffm_conv_gemm_f16_splitk", _ptr(x), n, H, W, C, k, stride, pad, _ptr(w), _ld(w), _ptr(partials), w.shape[0],
              partials.shape[0], _stream())


def patch_embed_s1_supported(W, cin, k, stride, pad, nout):
    """True when the fused stage-1 patch-embedding kernel takes this convolution."""
    return bool(_abi.load().cffm_patch_embed_s1_supported(int(W), int(cin), int(k), int(stride), int(pad), int(nout)))


def patch_embed_s1_weight(conv_weight):
    """conv.weight [Nout, 3, 7, 7] -> the kernel's fp16 [Nout, 192] layout: column (c*7 + ky)*8 + kx, zero padding."""
    o, c, kh, kw = conv_weight.shape
    assert (c, kh, kw) == (3, 7, 7)
    w = torch.zeros(o, c * kh, 8, dtype=torch.float32, device=conv_weight.device)
    w[:, :, :kw] = conv_weight.detach().float().reshape(o, c * kh, kw)
    out = torch.zeros(o, 192, dtype=_H, device=conv_weight.device)
    out[:, :c * kh * 8] = w.reshape(o, c * kh * 8).to(_H)
    return out


def patch_embed_s1(x, wk, bias, g1, b1, eps1, g2, b2, eps2, out32, ln_out16):
    """LayerNorm(conv7x7 s4 p3 (x) + bias) -> out32; LayerNorm of that -> ln_out16.  x fp32 [n, 3, H, W] contiguous."""
    _chk(x, _F, "patch_embed_s1.x", rows2d=False); _chk(wk, _H, "patch_embed_s1.w"); _chk(out32, _F, "patch_embed_s1.out32")
    _chk(ln_out16, _H, "patch_embed_s1.ln_out16")
    n, c, H, W = x.shape
    N = wk.shape[0]
    Ho, Wo = (H - 1) // 4 + 1, (W - 1) // 4 + 1
    assert x.is_contiguous() and c == 3 and tuple(wk.shape) == (N, 192) and wk.is_contiguous()
    assert out32.is_contiguous() and ln_out16.is_contiguous() and tuple(out32.shape) == (n * Ho * Wo, N) == tuple(ln_out16.shape)
    for t in (bias, g1, b1, g2, b2):
        _chk(t, _F, "patch_embed_s1.vector"); assert t.numel() == N
    _abi.call("cffm_patch_embed_s1", _ptr(x), n, H, W, _ptr(wk), _ptr(bias), _ptr(g1), _ptr(b1), float(eps1), _ptr(g2), _ptr(b2),
              float(eps2), _ptr(out32), _ptr(ln_out16), N, _stream())


def mixffn_tail_supported(N, hidden):
    """True when the fused dwconv + GELU + fc2 (+ residual + LayerNorm) kernel takes this (output width, hidden width)."""
    return bool(_abi.load().cffm_mixffn_tail_supported(int(N), int(hidden)))


def mixffn_tail(h, n, H, W, dw_w9c, dw_b, w2, b2, residual, out32, ln_gamma=None, ln_beta=None, ln_eps=0.0, ln_out16=None):
    """x = residual + (GELU(dwconv3x3(h) + dw_b)) @ w2.T + b2 -> out32 (optional, may alias residual);
    LayerNorm(x) -> ln_out16 (optional).  h fp16 [n*H*W, hidden] (NHWC), residual fp32 [n*H*W, N]."""
    _chk(h, _H, "mixffn_tail.h"); _chk(dw_w9c, _H, "mixffn_tail.dw_w"); _chk(dw_b, _F, "mixffn_tail.dw_b")
    _chk(w2, _H, "mixffn_tail.w2"); _chk(b2, _F, "mixffn_tail.b2"); _chk(residual, _F, "mixffn_tail.residual")
    HD, N = w2.shape[1], w2.shape[0]
    M = n * H * W
    assert h.is_contiguous() and h.numel() == M * HD and tuple(dw_w9c.shape) == (9, HD) and dw_w9c.is_contiguous()
    assert residual.is_contiguous() and tuple(residual.shape) == (M, N)
    for t, dt in ((out32, _F), (ln_out16, _H)):
        if t is not None:
            _chk(t, dt, "mixffn_tail.out"); assert t.is_contiguous() and tuple(t.shape) == (M, N)
    if ln_out16 is not None:
        _chk(ln_gamma, _F, "mixffn_tail.gamma"); _chk(ln_beta, _F, "mixffn_tail.beta")
    _abi.call("cffm_mixffn_tail", _ptr(h), n, H, W, HD, _ptr(dw_w9c), _ptr(dw_b), _ptr(w2), _ld(w2), _ptr(b2), _ptr(residual),
              _ptr(out32), _ptr(ln_gamma), _ptr(ln_beta), float(ln_eps), _ptr(ln_out16), N, _stream())


def layernorm_sum(partials, bias, gamma, beta, eps, out16=None, out32=None):
    """LayerNorm(sum_s partials[s] + bias)."""
    _chk(partials, _F, "layernorm_sum.partials")
    S, M, C = partials.shape
    assert partials.is_contiguous()
    for t in (out16, out32):
        if t is not None:
            assert tuple(t.shape) == (M, C)
    _abi.call("cffm_layernorm_sum", _ptr(partials), S, _ptr(bias), _ptr(gamma), _ptr(beta), float(eps), _ptr(out16),
              _ld(out16) if out16 is not None else 0, _ptr(out32), _ld(out32) if out32 is not None else 0, M, C, _stream())


def layernorm_chain(partials, bias, gamma, beta, eps, out32, gamma2, beta2, eps2, out16):
    """out32 = LayerNorm(sum_s partials[s] + bias); out16 = LayerNorm2(out32).  partials fp32 [S, M, C]."""
    _chk(partials, _F, "layernorm_chain.partials"); _chk(out32, _F, "layernorm_chain.out32"); _chk(out16, _H, "layernorm_chain.out16")
    S, M, C = partials.shape
    assert partials.is_contiguous() and tuple(out32.shape) == (M, C) == tuple(out16.shape)
    _abi.call("cffm_layernorm_chain", _ptr(partials), S, _ptr(bias), _ptr(gamma), _ptr(beta), float(eps), _ptr(out32), _ld(out32),
              _ptr(gamma2), _ptr(beta2), float(eps2), _ptr(out16), _ld(out16), M, C, _stream())


def gemm_ln_supported(N):
    return N <= 128 and N % 8 == 0


def layernorm(x, gamma, beta, eps, out16=None, out32=None):
    """Row LayerNorm of x [M,C] (fp32 or fp16)."""
    assert x.dim() == 2 and x.dtype in (_H, _F)
    _chk(x, x.dtype, "layernorm.x"); _chk(gamma, _F, "layernorm.gamma"); _chk(beta, _F, "layernorm.beta")
    M, C = x.shape
    if out16 is not None:
        _chk(out16, _H, "layernorm.out16"); assert tuple(out16.shape) == (M, C)
    if out32 is not None:
        _chk(out32, _F, "layernorm.out32"); assert tuple(out32.shape) == (M, C)
    _abi.call("cffm_layernorm", _ptr(x), int(x.dtype == _F), _ld(x), _ptr(gamma), _ptr(beta), float(eps),
              _ptr(out16), _ld(out16) if out16 is not None else 0, _ptr(out32),
              _ld(out32) if out32 is not None else 0, M, C, _stream())


def im2col(x, layout, N, H, W, C, k, stride, pad, out):
    """out [N*Ho*Wo, Kpad] fp16 patches, K order (ky,kx,c).  layout 0: fp32 NCHW, 1: fp16 NHWC."""
    _chk(x, _F if layout == 0 else _H, "im2col.x"); _chk(out, _H, "im2col.out")
    assert x.is_contiguous() and out.is_contiguous()
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    assert out.shape[0] == N * Ho * Wo, (out.shape, N, Ho, Wo)
    _abi.call("cffm_im2col", _ptr(x), layout, N, H, W, C, k, stride, pad, _ptr(out), out.shape[1], _stream())


def mha(q, k, v, out, batch, Nq, Nkv, heads, head_dim, scale):
    """softmax(scale q k^T) v.  q [batch*Nq, heads*d], k / v [batch*Nkv, ...] (column slices ok)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _chk(t, _H, "mha." + n)
    assert q.shape[0] == batch * Nq and k.shape[0] == batch * Nkv and v.shape[0] == batch * Nkv
    assert _ld(k) == _ld(v)
    _abi.call("cffm_mha_f16", _ptr(q), _ld(q), _ptr(k), _ptr(v), _ld(k), _ptr(out), _ld(out), batch, Nq, Nkv, heads,
              head_dim, float(scale), _stream())


def dwconv3x3_gelu(x, w9c, bias, out, N, H, W, C):
    _chk(x, _H, "dwconv.x"); _chk(w9c, _H, "dwconv.w"); _chk(bias, _F, "dwconv.bias"); _chk(out, _H, "dwconv.out")
    assert x.is_contiguous() and out.is_contiguous() and x.numel() == N * H * W * C == out.numel()
    assert tuple(w9c.shape) == (9, C)
    _abi.call("cffm_dwconv3x3_gelu", _ptr(x), _ptr(w9c), _ptr(bias), _ptr(out), N, H, W, C, _stream())


def head_fuse(p, sizes, N, C, t_perm, shift, c_full=None, half32=None, half16=None):
    """p: 4 fp16 NHWC maps [N,Hi,Wi,C]; sizes: [(H1,W1),...]."""
    for t in p:
        _chk(t, _H, "head_fuse.p"); assert t.is_contiguous()
    _chk(shift, _F, "head_fuse.shift")
    (H1, W1), (H2, W2), (H3, W3), (H4, W4) = sizes
    if c_full is not None:
        _chk(c_full, _H, "head_fuse.c_full"); assert c_full.is_contiguous()
    if half32 is not None:
        _chk(half32, _F, "head_fuse.half32")
    if half16 is not None:
        _chk(half16, _H, "head_fuse.half16")
    _abi.call("cffm_head_fuse", _ptr(p[0]), _ptr(p[1]), _ptr(p[2]), _ptr(p[3]), N, H1, W1, H2, W2, H3, W3, H4, W4, C,
              t_perm, _ptr(shift), _ptr(c_full), _ptr(half32), _ld(half32) if half32 is not None else 0,
              _ptr(half16), _ld(half16) if half16 is not None else 0, _stream())


def cffa_norm(x, gamma, beta, eps, xn, xt_pad, B, T, H, W, Hp, Wp, C):
    _chk(x, _F, "cffa_norm.x"); _chk(xn, _H, "cffa_norm.xn"); _chk(xt_pad, _H, "cffa_norm.xt_pad")
    assert x.is_contiguous() and xn.is_contiguous() and xt_pad.is_contiguous()
    assert x.numel() == B * T * H * W * C == xn.numel() and xt_pad.numel() == B * (Hp + 6) * (Wp + 6) * C
    _abi.call("cffm_cffa_norm", _ptr(x), _ptr(gamma), _ptr(beta), float(eps), _ptr(xn), _ptr(xt_pad), B, T, H, W, Hp,
              Wp, C, _stream())


def cffa_pool(xn, B, T, H, W, C, pool_w, pool_b, pooled):
    _chk(xn, _H, "cffa_pool.xn"); _chk(pool_w, _F, "cffa_pool.pool_w"); _chk(pool_b, _F, "cffa_pool.pool_b")
    _chk(pooled, _H, "cffa_pool.pooled")
    assert pool_w.numel() == 49 + 49 + 9 + 4 and pool_b.numel() == 4 and pooled.is_contiguous()
    _abi.call("cffm_cffa_pool", _ptr(xn), B, T, H, W, C, _ptr(pool_w), _ptr(pool_b), _ptr(pooled), _stream())


def cffa_pool_part(xn, B, part, H, W, C, pool_w, pool_b, pooled):
    """part 0: target level from xn [B,H,W,C]; part 1: reference levels from xn [3,B,H,W,C]; rows of the other part untouched."""
    _chk(xn, _H, "cffa_pool_part.xn"); _chk(pool_w, _F, "cffa_pool_part.pool_w"); _chk(pool_b, _F, "cffa_pool_part.pool_b")
    _chk(pooled, _H, "cffa_pool_part.pooled")
    assert pool_w.numel() == 49 + 49 + 9 + 4 and pool_b.numel() == 4 and pooled.is_contiguous() and xn.is_contiguous()
    assert xn.numel() == (1 if part == 0 else 3) * B * H * W * C
    _abi.call("cffm_cffa_pool_part", _ptr(xn), B, part, H, W, C, _ptr(pool_w), _ptr(pool_b), _ptr(pooled), _stream())


def cffa_norm_frames(x, gamma, beta, eps, xn, xt_pad, n_frames, first_target, H, W, Hp, Wp, C):
    _chk(x, _F, "cffa_norm_frames.x"); _chk(xn, _H, "cffa_norm_frames.xn")
    assert x.is_contiguous() and xn.is_contiguous() and x.numel() == n_frames * H * W * C == xn.numel()
    if xt_pad is not None:
        _chk(xt_pad, _H, "cffa_norm_frames.xt_pad")
        assert xt_pad.is_contiguous() and xt_pad.numel() == (n_frames - first_target) * (Hp + 6) * (Wp + 6) * C
    _abi.call("cffm_cffa_norm_frames", _ptr(x), _ptr(gamma), _ptr(beta), float(eps), _ptr(xn), _ptr(xt_pad), n_frames,
              first_target, H, W, Hp, Wp, C, _stream())


def cffa_pool_level(xn, n_frames, level, H, W, C, pool_w, pool_b, pooled):
    _chk(xn, _H, "cffa_pool_level.xn"); _chk(pooled, _H, "cffa_pool_level.pooled")
    assert xn.is_contiguous() and pooled.is_contiguous() and xn.numel() == n_frames * H * W * C
    _abi.call("cffm_cffa_pool_level", _ptr(xn), n_frames, level, H, W, C, _ptr(pool_w), _ptr(pool_b), _ptr(pooled), _stream())


_cfm_layout = None


def cfm_layout():
    """Key-row layout of the CFM kernel: {"rows": [halo, pooled target, ref0, ref1, ref2], "npad", "pitch", "apron"}."""
    global _cfm_layout
    if _cfm_layout is None:
        import ctypes
        buf = (ctypes.c_int32 * 8)()
        _abi.check(_abi.load().cffm_cfm_layout(ctypes.cast(buf, ctypes.c_void_p)), "cffm_cfm_layout")
        _cfm_layout = {"rows": list(buf[0:5]), "npad": buf[5], "pitch": buf[6], "apron": buf[7]}
    return _cfm_layout


def apron_rows(B, H, W):
    """Rows of the target map in the CFM kernel's cyclic-apron layout [B, Hp+6, Wp+6]."""
    return B * ((H + 6) // 7 * 7 + 6) * ((W + 6) // 7 * 7 + 6)


def cfm_attention(qkv_a, kv_pooled, bias_tab, out, B, H, W, C, heads, scale, dump=None):
    """qkv_a fp16 [B*(Hp+6)*(Wp+6), 3C] (apron layout), kv_pooled fp16 [B*15*nW, 2C], bias_tab = cffm_tables.assemble_bias_tc."""
    _chk(qkv_a, _H, "cfm.qkv_a"); _chk(kv_pooled, _H, "cfm.kv_pooled"); _chk(bias_tab, _H, "cfm.bias_tab"); _chk(out, _H, "cfm.out")
    assert qkv_a.is_contiguous() and kv_pooled.is_contiguous() and bias_tab.is_contiguous() and out.is_contiguous()
    lay = cfm_layout()
    assert tuple(bias_tab.shape) == (heads, 49, lay["pitch"]), (tuple(bias_tab.shape), lay)
    nW = ((H + 6) // 7) * ((W + 6) // 7)
    assert qkv_a.numel() == apron_rows(B, H, W) * 3 * C and kv_pooled.numel() == B * 15 * nW * 2 * C and out.numel() == B * H * W * C
    if dump is None:
        _abi.call("cffm_cfm_attention", _ptr(qkv_a), _ptr(kv_pooled), _ptr(bias_tab), _ptr(out), B, H, W, C, heads,
                  float(scale), _stream())
    else:
        _chk(dump, _H, "cfm.dump")
        assert dump.is_contiguous() and dump.numel() == 4 * B * nW * 2 * lay["npad"] * 64
        _abi.call("cffm_cfm_attention_dump", _ptr(qkv_a), _ptr(kv_pooled), _ptr(bias_tab), _ptr(out), _ptr(dump), B, H, W, C,
                  heads, float(scale), _stream())


def cfm_attention_slots(qkv_a, kv_tgt, gathered, role_off, role_slots, ref_slot, bias_tab, out, B, H, W, C, heads, scale):
    """CFM attention reading the reference-frame K/V in place from the all-gathered buffer (frame-sharded path).
    kv_tgt fp16 [B, nW, 2C] contiguous; gathered fp16 [n_ranks, flat_tokens, 2C] (a view: dim-0 stride = rank stride) in
    which the maps of reference role k start at token role_off[k], role_slots[k] of them per rank, {1,4,9}[k] nW tokens each;
    ref_slot int32 [B, 3, 2] = (slot, rank) on the device."""
    _chk(qkv_a, _H, "cfm.qkv_a"); _chk(kv_tgt, _H, "cfm.kv_tgt"); _chk(gathered, _H, "cfm.gathered"); _chk(bias_tab, _H, "cfm.bias_tab")
    _chk(out, _H, "cfm.out")
    nW = ((H + 6) // 7) * ((W + 6) // 7)
    assert kv_tgt.is_contiguous() and kv_tgt.numel() == B * nW * 2 * C
    assert gathered.dim() == 3 and gathered.shape[2] == 2 * C and gathered.stride(1) == 2 * C
    per = (1, 4, 9)
    assert all(role_off[k] + role_slots[k] * per[k] * nW <= gathered.shape[1] for k in range(3))
    assert ref_slot.dtype == torch.int32 and ref_slot.is_cuda and ref_slot.is_contiguous() and tuple(ref_slot.shape) == (B, 3, 2)
    assert qkv_a.numel() == apron_rows(B, H, W) * 3 * C and out.numel() == B * H * W * C
    base, es = gathered.data_ptr(), gathered.element_size()
    _abi.call("cffm_cfm_attention_slots", _ptr(qkv_a), _ptr(kv_tgt), nW * 2 * C,
              base + role_off[0] * 2 * C * es, base + role_off[1] * 2 * C * es, base + role_off[2] * 2 * C * es,
              per[0] * nW * 2 * C, per[1] * nW * 2 * C, per[2] * nW * 2 * C, max(role_slots[0], 1), max(role_slots[1], 1),
              max(role_slots[2], 1), gathered.stride(0), gathered.shape[0], _ptr(ref_slot), _ptr(bias_tab), _ptr(out), B, H, W, C,
              heads, float(scale), _stream())


def resize_nhwc_to_nchw(x, ncls, out, B, h, w, Ho, Wo):
    assert x.dtype in (_H, _F)
    _chk(x, x.dtype, "resize.x"); _chk(out, _F, "resize.out")
    assert out.is_contiguous() and out.numel() == B * ncls * Ho * Wo
    _abi.call("cffm_resize_nhwc_to_nchw", _ptr(x), int(x.dtype == _F), x.stride(-2) if x.dim() > 1 else ncls, _ptr(out),
              B, h, w, ncls, Ho, Wo, _stream())


def resize_argmax(logits, labels, B, ncls, h, w, Ho, Wo):
    _chk(logits, _F, "resize_argmax.logits")
    assert logits.is_contiguous() and labels.dtype == torch.int64 and labels.is_cuda and labels.is_contiguous()
    assert logits.numel() == B * ncls * h * w and labels.numel() == B * Ho * Wo
    _abi.call("cffm_resize_argmax", _ptr(logits), _ptr(labels), B, ncls, h, w, Ho, Wo, _stream())


def upsample2_argmax(scores, ncls, labels, B, h, w, Hm, Wm, Ho, Wo):
    """scores fp32 [B*h*w, ldc] NHWC -> two chained bilinear resizes -> argmax labels [B,Ho,Wo], int64 or uint8 (ncls <= 256)."""
    _chk(scores, _F, "upsample2_argmax.scores")
    assert scores.dim() == 2 and scores.shape[0] == B * h * w and scores.shape[1] >= ncls
    assert labels.dtype in (torch.int64, torch.uint8) and labels.is_cuda and labels.is_contiguous() and labels.numel() == B * Ho * Wo
    _abi.call("cffm_upsample2_argmax" if labels.dtype == torch.int64 else "cffm_upsample2_argmax_u8", _ptr(scores), _ld(scores),
              _ptr(labels), B, h, w, ncls, Hm, Wm, Ho, Wo, _stream())


def upsample2_argmax_supported(Hm, Wm, Ho, Wo):
    return (16 * Hm + Ho - 1) // Ho + 3 <= 8 and (16 * Wm + Wo - 1) // Wo + 3 <= 8


def resize_nchw(x, out):
    _chk(x, _F, "resize_nchw.x"); _chk(out, _F, "resize_nchw.out")
    assert x.is_contiguous() and out.is_contiguous() and x.shape[:2] == out.shape[:2]
    B, C, h, w = x.shape
    _abi.call("cffm_resize_nchw", _ptr(x), _ptr(out), B, C, h, w, out.shape[2], out.shape[3], _stream())


def softmax_nchw(x, out):
    _chk(x, _F, "softmax_nchw.x"); _chk(out, _F, "softmax_nchw.out")
    assert x.is_contiguous() and out.is_contiguous() and x.shape == out.shape
    B, C, h, w = x.shape
    _abi.call("cffm_softmax_nchw", _ptr(x), _ptr(out), B, C, h * w, _stream())


_side = {}


class fork:
    """``with ops.fork(key): ...`` enqueues the body on the side stream ``key``, which first waits for everything
    already on the current stream; ``ops.join(key)`` makes the current stream wait for it.  Independent branches of
    the forward (q projection vs the spatial-reduction K/V chain, decoder projections vs the later backbone stages,
    reference-frame assembling vs the target path) then overlap on the GPU; the fork/join pattern is preserved as
    parallel branches when the pass is captured in a CUDA graph.  Every forked stream must be joined (directly or
    through another stream) before the pass ends."""

    def __init__(self, key=0):
        self.key = key

    def __enter__(self):
        k = (torch.cuda.current_device(), self.key)
        if k not in _side:
            _side[k] = torch.cuda.Stream(device=k[0])
        self.side = _side[k]
        self.side.wait_stream(torch.cuda.current_stream())
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self.ctx.__exit__(*exc)
        return False


def join(key=0):
    side = _side.get((torch.cuda.current_device(), key))
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)


class KernelTimer:
    """CUDA-event timing of selected entry points on the launching stream (bench.py's live roofline
    measurement).  ``with KernelTimer({"cffm_cfm_attention"}) as kt: step()`` then ``kt.results()``."""

    def __init__(self, names=None):
        self.names = None if names is None else set(names)      # None: every entry point
        self.records = []                                        # (name, args, start_event, end_event)
        self._open = None

    def _hook(self, name, phase, args):
        if self.names is not None and name not in self.names:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        if phase == 0:
            self._open = ev
        else:
            self.records.append((name, args, self._open, ev))

    def __enter__(self):
        _abi.launch_hook = self._hook
        return self

    def __exit__(self, *exc):
        _abi.launch_hook = None

    def results(self):
        """list of (name, args, milliseconds) per timed launch; synchronises."""
        torch.cuda.synchronize()
        return [(n, a, s.elapsed_time(e)) for n, a, s, e in self.records]
