"""MixVisionTransformer (SegFormer MiT-B0..B5) backbone, B200-native.

Mirrors the reference plugin surface (mmseg/models/backbones/mix_transformer.py:203-423): classes
registered in BACKBONES as ``mit_b0`` .. ``mit_b5``, ``**kwargs`` swallowed (configs pass
``style='pytorch'``), ``init_weights(pretrained)``, ``forward(x) -> list of 4 NCHW maps`` at strides
4/8/16/32, and the SAME state-dict keys and shapes, so reference checkpoints load unchanged.

The nn.Module tree below only HOLDS parameters.  The arithmetic is a fixed sequence of C-ABI
calls (``ops``) on fp16 token-major ("NHWC") activations with an fp32 residual stream:

  patch embed  : im2col -> tcgen05 GEMM whose epilogue applies the patch-embed LayerNorm (eps 1e-5) AND the first
                 block's norm1 (C <= 128), or split-K GEMM + one two-LayerNorm pass over the partial sums  (:173-200)
  attention    : q GEMM on the main stream ; [sr: im2col -> split-K GEMM -> LN(1e-5)] -> kv GEMM on a side stream ;
                 tcgen05 attention kernel (TMA Q/K/V, QK^T and PV on tensor cores, S in TMEM; <= 256 keys, else the
                 mma.sync kernel) ; proj GEMM + residual (in place) + norm2 in its epilogue     (:96-117,:154)
  Mix-FFN      : fc1 GEMM (TMA-store epilogue) -> depthwise 3x3 + GELU -> fc2 GEMM + residual + the next norm1 /
                 the stage norm in its epilogue (separate LayerNorm kernels when C > 128)        (:48-55,:155)
  stage output : fp16 NHWC (also the next stage's im2col input), handed to the decode head's projection at once
                 through ``stage_hook``                                                           (:321-349)
"""
import math
from functools import partial

import torch
import torch.nn as nn

from . import _abi, ops
from .registry import BACKBONES
from .workspace import Workspace

_H, _F = torch.float16, torch.float32
_FFN_FUSED = __import__("os").environ.get("CFFM_FFN_FUSED", "1") != "0"   # 0: separate dwconv and fc2 kernels (A/B runs)
_PE_FUSED = __import__("os").environ.get("CFFM_PE_FUSED", "1") != "0"     # 0: im2col + GEMM for the stage-1 patch embedding


def _round_up(x, m):
    return (x + m - 1) // m * m


class _DWConv(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.dwconv = _DWConv(hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias, sr_ratio):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim, self.num_heads, self.sr_ratio = dim, num_heads, sr_ratio
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)                       # bare LayerNorm: eps 1e-5 (:77)


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer, sr_ratio):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias, sr_ratio)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _OverlapPatchEmbed(nn.Module):
    def __init__(self, patch_size, stride, in_chans, embed_dim):
        super().__init__()
        self.patch_size, self.stride = patch_size, stride
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride, padding=patch_size // 2)
        self.norm = nn.LayerNorm(embed_dim)                     # eps 1e-5 (:175)


class MixVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dims=(64, 128, 256, 512),
                 num_heads=(1, 2, 4, 8), mlp_ratios=(4, 4, 4, 4), qkv_bias=False, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, depths=(3, 4, 6, 3),
                 sr_ratios=(8, 4, 2, 1)):
        super().__init__()
        assert qk_scale is None, "qk_scale is never set by the reference's mit_b* (mix_transformer.py:373-423)"
        self.num_classes = num_classes
        self.depths = list(depths)
        self.embed_dims, self.num_heads, self.sr_ratios = list(embed_dims), list(num_heads), list(sr_ratios)
        self.in_chans = in_chans
        chans = [in_chans] + list(embed_dims)
        for s in range(4):
            k, st = (7, 4) if s == 0 else (3, 2)
            setattr(self, f"patch_embed{s + 1}", _OverlapPatchEmbed(k, st, chans[s], chans[s + 1]))
            setattr(self, f"block{s + 1}", nn.ModuleList([
                _Block(embed_dims[s], num_heads[s], mlp_ratios[s], qkv_bias, norm_layer, sr_ratios[s])
                for _ in range(depths[s])]))
            setattr(self, f"norm{s + 1}", norm_layer(embed_dims[s]))
        self.apply(self._init_weights)
        self._plan = None
        self._ws = Workspace()
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_plan())

    # ------------------------------------------------------------------ reference-compatible API
    @staticmethod
    def _init_weights(m):
        """mix_transformer.py:258-272."""
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                m.bias.data.zero_()

    def init_weights(self, pretrained=None):
        """mix_transformer.py:276-279: ``load_checkpoint(strict=False)`` when a path is given."""
        if isinstance(pretrained, str):
            ckpt = torch.load(pretrained, map_location="cpu")
            sd = ckpt.get("state_dict", ckpt)
            self.load_state_dict({k[9:] if k.startswith("backbone.") else k: v for k, v in sd.items()}, strict=False)

    def invalidate_plan(self):
        self._plan = None

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    def train(self, mode=True):
        if mode:
            raise _abi.CffmError("vss_cffm_b200 implements the inference hot path only (eval mode); "
                                 "training is out of scope (SURVEY.md section 8)")
        return super().train(False)

    # ------------------------------------------------------------------ plan: fp16 operand copies
    def _build_plan(self):
        dev = self.norm1.weight.device
        if dev.type != "cuda":
            raise _abi.CffmError("the backbone runs on a CUDA (sm_100) device only; call .cuda() first")
        _abi.require_device()
        h = lambda t: t.detach().to(dev, _H).contiguous()
        f = lambda t: t.detach().to(dev, _F).contiguous()
        stages = []
        for s in range(4):
            pe = getattr(self, f"patch_embed{s + 1}")
            w = pe.proj.weight.detach()                          # (Cout, Cin, k, k) -> (Cout, k, k, Cin)
            cout, cin, k, _ = w.shape
            kdim = k * k * cin
            kpad = _round_up(kdim, 8)
            wp = torch.zeros(cout, kpad, device=dev, dtype=_H)
            wp[:, :kdim] = w.permute(0, 2, 3, 1).reshape(cout, kdim).to(dev, _H)
            st = dict(k=k, stride=pe.stride, pad=k // 2, cin=cin, cout=cout, kpad=kpad, w=wp, b=f(pe.proj.bias),
                      ng=f(pe.norm.weight), nb=f(pe.norm.bias), eps=pe.norm.eps, blocks=[])
            if (cin, k) == (3, 7):                               # stage 1: weight layout of the fused patch-embedding kernel
                st["wk"] = ops.patch_embed_s1_weight(w.to(dev))
            for blk in getattr(self, f"block{s + 1}"):
                a, m = blk.attn, blk.mlp
                b = dict(n1g=f(blk.norm1.weight), n1b=f(blk.norm1.bias), n1eps=blk.norm1.eps,
                         n2g=f(blk.norm2.weight), n2b=f(blk.norm2.bias), n2eps=blk.norm2.eps,
                         qw=h(a.q.weight), qb=f(a.q.bias) if a.q.bias is not None else None,
                         kvw=h(a.kv.weight), kvb=f(a.kv.bias) if a.kv.bias is not None else None,
                         pw=h(a.proj.weight), pb=f(a.proj.bias), sr=a.sr_ratio,
                         f1w=h(m.fc1.weight), f1b=f(m.fc1.bias), f2w=h(m.fc2.weight), f2b=f(m.fc2.bias),
                         dww=h(m.dwconv.dwconv.weight.detach().reshape(-1, 9).t()), dwb=f(m.dwconv.dwconv.bias))
                if a.sr_ratio > 1:
                    sw = a.sr.weight.detach()                    # (C, C, sr, sr) -> (C, sr, sr, C)
                    b.update(srw=h(sw.permute(0, 2, 3, 1).reshape(sw.shape[0], -1)), srb=f(a.sr.bias),
                             sng=f(a.norm.weight), snb=f(a.norm.bias), seps=a.norm.eps)
                st["blocks"].append(b)
            fn = getattr(self, f"norm{s + 1}")
            st.update(fg=f(fn.weight), fb=f(fn.bias), feps=fn.eps)
            stages.append(st)
        self._plan = stages
        return stages

    # ------------------------------------------------------------------ forward
    def forward_features(self, x, stage_hook=None):
        """mix_transformer.py:313-349.  x (N,3,H,W) fp32 -> 4 fp16 NHWC maps returned as NCHW views.
        ``stage_hook(s, out)`` (optional) is called as soon as stage s is enqueued, so that a consumer can start
        work that only needs that stage (the decode head's per-stage projection) beside the later stages."""
        if x.dim() != 4 or x.shape[1] != self.in_chans:
            raise _abi.CffmError(f"expected (N,{self.in_chans},H,W) input, got {tuple(x.shape)}")
        plan = self._plan or self._build_plan()
        dev = plan[0]["w"].device
        x = x.to(dev, _F).contiguous()
        ws = self._ws
        N, _, H, W = x.shape
        outs = []
        cur, layout = x, 0
        for s, st in enumerate(plan):
            k, stride, pad, C = st["k"], st["stride"], st["pad"], st["cout"]
            Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
            M = N * Ho * Wo
            xres = ws.get(f"s{s}.x", (M, C), _F)                 # fp32 residual stream
            S = ops.splitk_plan(M, C, st["kpad"])                # few tiles x long K (stages 3-4): split K over the SMs
            pe32 = ws.get(f"s{s}.pe32", (S, M, C), _F)
            xn = ws.get(f"s{s}.xn", (M, C), _H)
            b0 = st["blocks"][0]
            # The convolution is an implicit GEMM on the NHWC image itself (TMA boxes walking it with the conv stride, zero
            # fill = zero padding) whenever the image is fp16 NHWC with C % 64 == 0 (stages 2-4); the fp32 NCHW input frames
            # of stage 1 (3 channels) go through a patch matrix.
            implicit = layout == 1 and ops.conv_gemm_supported(st["cin"], Wo, stride) and (S > 1 or C <= 128)
            fused_pe = layout == 0 and _PE_FUSED and "wk" in st and ops.patch_embed_s1_supported(W, st["cin"], k, stride, pad, C)
            if fused_pe:
                # stage 1: convolution of the fp32 NCHW frames, both norms, one kernel, no patch matrix (csrc/patch_embed_sm100.cu)
                ops.patch_embed_s1(cur, st["wk"], st["b"], st["ng"], st["nb"], st["eps"], b0["n1g"], b0["n1b"], b0["n1eps"], xres, xn)
            elif not implicit:
                col = ws.get(f"s{s}.col", (M, st["kpad"]), _H)
                ops.im2col(cur, layout, N, H, W, st["cin"], k, stride, pad, col)
            # patch-embed norm and the first block's norm1 in one pass (the row stays in registers in between)
            if fused_pe:
                pass                                             # done above
            elif S > 1:
                if implicit:
                    ops.conv_gemm_splitk(cur, N, H, W, st["cin"], k, stride, pad, st["w"], pe32)
                else:
                    ops.gemm_splitk(col, st["w"], pe32)
                ops.layernorm_chain(pe32, st["b"], st["ng"], st["nb"], st["eps"], xres, b0["n1g"], b0["n1b"], b0["n1eps"], xn)
            elif C <= 128:                                       # both norms ride in the GEMM epilogue (one tile spans the row)
                if implicit:
                    ops.conv_gemm_ln(cur, N, H, W, st["cin"], k, stride, pad, st["w"], st["b"], xres, st["ng"], st["nb"], st["eps"], xn,
                                     b0["n1g"], b0["n1b"], b0["n1eps"])
                else:
                    ops.gemm_ln_chain(col, st["w"], st["b"], xres, st["ng"], st["nb"], st["eps"], b0["n1g"], b0["n1b"], b0["n1eps"], xn)
            else:
                ops.gemm(col, st["w"], bias=st["b"], out32=pe32[0])
                ops.layernorm_chain(pe32, None, st["ng"], st["nb"], st["eps"], xres, b0["n1g"], b0["n1b"], b0["n1eps"], xn)
            heads = self.num_heads[s]
            d = C // heads
            out = ws.get(f"s{s}.out", (M, C), _H)
            fuse_ln = ops.gemm_ln_supported(C)                   # LayerNorm rides in the GEMM epilogue when N = C <= 128
            nblk = len(st["blocks"])
            for bi, b in enumerate(st["blocks"]):
                # ---- efficient self-attention
                if bi > 0 and not fuse_ln:                       # block 0: chained above; fusable widths: in the previous GEMM
                    ops.layernorm(xres, b["n1g"], b["n1b"], b["n1eps"], out16=xn)
                q = ws.get(f"s{s}.q", (M, C), _H)
                sr = b["sr"]
                with ops.fork():                                 # K/V chain on the side stream, q projection on the main one
                    if sr > 1:
                        Hs, Ws_ = (Ho - sr) // sr + 1, (Wo - sr) // sr + 1
                        Ms = N * Hs * Ws_
                        S = ops.splitk_plan(Ms, C, sr * sr * C)   # 15 tiles x K up to 4096: split K over the SMs
                        s32 = ws.get(f"s{s}.sr32", (S, Ms, C), _F)
                        kvin = ws.get(f"s{s}.srn", (Ms, C), _H)
                        # Attention.sr has kernel = stride: a pure re-tiling of xn, read in place by strided TMA boxes
                        implicit = ops.conv_gemm_supported(C, Ws_, sr) and (S > 1 or C <= 128)
                        if not implicit:
                            scol = ws.get(f"s{s}.srcol", (Ms, sr * sr * C), _H)
                            ops.im2col(xn, 1, N, Ho, Wo, C, sr, sr, 0, scol)
                        if S > 1:
                            if implicit:
                                ops.conv_gemm_splitk(xn, N, Ho, Wo, C, sr, sr, 0, b["srw"], s32)
                            else:
                                ops.gemm_splitk(scol, b["srw"], s32)
                            ops.layernorm_sum(s32, b["srb"], b["sng"], b["snb"], b["seps"], out16=kvin)
                        elif implicit:                           # conv + bias + Attention.norm in one launch
                            ops.conv_gemm_ln(xn, N, Ho, Wo, C, sr, sr, 0, b["srw"], b["srb"], None, b["sng"], b["snb"], b["seps"], kvin)
                        else:
                            ops.gemm(scol, b["srw"], bias=b["srb"], out32=s32[0])
                            ops.layernorm(s32[0], b["sng"], b["snb"], b["seps"], out16=kvin)
                        nkv = Hs * Ws_
                    else:
                        kvin, Ms, nkv = xn, M, Ho * Wo
                    kv = ws.get(f"s{s}.kv", (Ms, 2 * C), _H)
                    ops.gemm(kvin, b["kvw"], bias=b["kvb"], out16=kv)
                ops.gemm(xn, b["qw"], bias=b["qb"], out16=q)
                ops.join()
                ao = ws.get(f"s{s}.ao", (M, C), _H)
                ops.mha(q, kv[:, :C], kv[:, C:], ao, N, Ho * Wo, nkv, heads, d, d ** -0.5)
                # ---- x += proj(attn) ; xn = norm2(x)   (one kernel when fusable)
                if fuse_ln:
                    ops.gemm_ln(ao, b["pw"], b["pb"], xres, xres, b["n2g"], b["n2b"], b["n2eps"], xn)
                else:
                    ops.gemm(ao, b["pw"], bias=b["pb"], residual=xres, out32=xres)
                    ops.layernorm(xres, b["n2g"], b["n2b"], b["n2eps"], out16=xn)
                # ---- Mix-FFN
                Ch = b["f1w"].shape[0]
                h1 = ws.get(f"s{s}.h1", (M, Ch), _H)
                ops.gemm(xn, b["f1w"], bias=b["f1b"], out16=h1)
                # ---- x += fc2(GELU(dwconv(h1))) ; then the next block's norm1 or the stage norm
                if fuse_ln and _FFN_FUSED and ops.mixffn_tail_supported(C, Ch):
                    # one kernel: the convolved / activated hidden map never leaves the SM (csrc/mixffn_sm100.cu)
                    if bi + 1 < nblk:
                        nb = st["blocks"][bi + 1]
                        ops.mixffn_tail(h1, N, Ho, Wo, b["dww"], b["dwb"], b["f2w"], b["f2b"], xres, xres,
                                        nb["n1g"], nb["n1b"], nb["n1eps"], xn)
                    else:
                        ops.mixffn_tail(h1, N, Ho, Wo, b["dww"], b["dwb"], b["f2w"], b["f2b"], xres, None,
                                        st["fg"], st["fb"], st["feps"], out)
                    continue
                h2 = ws.get(f"s{s}.h2", (M, Ch), _H)
                ops.dwconv3x3_gelu(h1, b["dww"], b["dwb"], h2, N, Ho, Wo, Ch)
                if fuse_ln:
                    if bi + 1 < nblk:
                        nb = st["blocks"][bi + 1]
                        ops.gemm_ln(h2, b["f2w"], b["f2b"], xres, xres, nb["n1g"], nb["n1b"], nb["n1eps"], xn)
                    else:
                        ops.gemm_ln(h2, b["f2w"], b["f2b"], xres, None, st["fg"], st["fb"], st["feps"], out)
                else:
                    ops.gemm(h2, b["f2w"], bias=b["f2b"], residual=xres, out32=xres)
            if not fuse_ln or nblk == 0:
                ops.layernorm(xres, st["fg"], st["fb"], st["feps"], out16=out)
            outs.append(out.view(N, Ho, Wo, C).permute(0, 3, 1, 2))   # logical NCHW, channels-last memory
            if stage_hook is not None:
                stage_hook(s, outs[-1])
            cur, layout, H, W = out, 1, Ho, Wo
        return outs

    def forward(self, x, stage_hook=None):
        return self.forward_features(x, stage_hook)


def _variant(embed_dims, depths):
    def ctor(self, **kwargs):                                    # kwargs (style='pytorch') are swallowed (:384-388)
        MixVisionTransformer.__init__(
            self, patch_size=4, embed_dims=embed_dims, num_heads=[1, 2, 5, 8], mlp_ratios=[4, 4, 4, 4],
            qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), depths=depths, sr_ratios=[8, 4, 2, 1],
            drop_rate=0.0, drop_path_rate=0.1)
    return ctor


for _name, _dims, _depths in (("mit_b0", [32, 64, 160, 256], [2, 2, 2, 2]), ("mit_b1", [64, 128, 320, 512], [2, 2, 2, 2]),
                              ("mit_b2", [64, 128, 320, 512], [3, 4, 6, 3]), ("mit_b3", [64, 128, 320, 512], [3, 4, 18, 3]),
                              ("mit_b4", [64, 128, 320, 512], [3, 8, 27, 3]), ("mit_b5", [64, 128, 320, 512], [3, 6, 40, 3])):
    _cls = type(_name, (MixVisionTransformer,), {"__init__": _variant(_dims, _depths), "__module__": __name__})
    globals()[_name] = BACKBONES.register_module()(_cls)
