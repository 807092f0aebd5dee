"""CFFM / CFFM++ decode heads, B200-native.

Plugin surface of the reference (mmseg/models/decode_heads/cffm_head.py:40-157, :161-300, :303-535 on top of
decode_head.py:513-708): registered in HEADS as ``CFFMHead_clips_resize1_8``,
``CFFMHead_clips_resize1_8_finetune_w_prototype3`` and ``CFFMHead_clips_resize1_8_gene_prototype``; same constructor kwargs, attributes
(``num_classes``, ``align_corners``, ``num_clips`` ...), ``init_weights()``, ``forward_test(...)`` ->
(B, num_classes, h, w) logits at 1/4 scale, and the SAME state-dict keys / shapes.

The nn.Module tree only HOLDS parameters.  Eval-mode arithmetic is a fixed sequence of C-ABI calls:

  MLP decoder  : linear_c{i}, linear_fuse and its BatchNorm are folded on the host into four
                 C_i -> 256 projections applied at NATIVE resolution (bilinear upsampling commutes
                 with a 1x1 conv); one kernel then upsamples, sums, ReLUs and emits the 2x2 mean
                 (= resize(_c, 1/2), cffm_head.py:131-133)                     (:102-133)
  CFFM block   : target: norm1 -> pad -> qkv GEMM ; pooled K/V: target-level fc-pool + the three reference
                 levels (norm1 -> pad -> resize -> fc-pool, produced one block ahead on a side stream:
                 reference frames are read-only, :826) -> kv GEMM ; cfm_attention (in-kernel K/V
                 assembling) -> proj GEMM + residual -> LN -> fc1 GELU -> fc2 + residual
                 (cffm_transformer.py:709-832)
  classifier   : linear_pred2 as two accumulating GEMMs over [c_target | cffm_target] (no concat),
                 then the x2 bilinear resize to NCHW fp32                     (:145-155)
  CFFM++       : prototype cross-attention (swin_transformer_2d.py:208-262, :605-665), linear_pred3
                 folded into the same logits buffer with weight 0.5           (:519-532)

Internally frames are FRAME-MAJOR ([T,B,...]: the B target frames are contiguous at the end); the
reference's clip-major order (n = b*T + t) is accepted and permuted inside the fuse kernel.
"""
import glob
import os

import torch
import torch.nn as nn

from . import _abi, ops
from . import cffm_tables as tb
from .registry import HEADS, build_loss
from .workspace import Workspace

_H, _F = torch.float16, torch.float32
HEADS_N, WS = 8, tb.WS
NCLS_ALIGN = 8


def _round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# parameter holders (names = the reference's state-dict keys)
class _MLPEmbed(nn.Module):
    def __init__(self, input_dim, embed_dim):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)


class _ConvBN(nn.Module):
    """mmcv ConvModule(conv -> SyncBN -> ReLU); conv has no bias when a norm follows."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn = nn.BatchNorm2d(cout)                       # eval-mode SyncBN == BN affine (eps 1e-5)
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_out", nonlinearity="relu")


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _WindowAttention3d3(nn.Module):
    """cffm_transformer.py:221-362 (parameters and index buffers only)."""

    def __init__(self, dim, heads):
        super().__init__()
        wa = WS * WS
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * WS - 1) ** 2, heads))   # stays zero (:253)
        self.relative_position_bias_table_to_neighbors = nn.Parameter(torch.zeros(1, heads, wa, tb.N_RING))
        nn.init.trunc_normal_(self.relative_position_bias_table_to_neighbors, std=.02)
        self.relative_position_bias_table_to_windows = nn.ParameterList(
            [nn.Parameter(torch.zeros(heads, (WS + tb.FOCAL_WINDOW - 1) ** 2))])
        self.relative_position_bias_table_to_windows_clips = nn.ParameterList(
            [nn.Parameter(torch.zeros(heads, (WS + kc - 1) ** 2)) for kc in tb.K_CLIPS])
        for p in list(self.relative_position_bias_table_to_windows) + list(self.relative_position_bias_table_to_windows_clips):
            nn.init.trunc_normal_(p, std=.02)
        self.register_buffer("relative_position_index", tb.relative_position_index((WS, WS), (WS, WS)))
        self.register_buffer("valid_ind_rolled", tb.valid_ind_rolled())
        self.register_buffer("relative_position_index_0",
                             tb.relative_position_index((WS, WS), (tb.FOCAL_WINDOW, tb.FOCAL_WINDOW)))
        for k, kc in enumerate(tb.K_CLIPS):
            self.register_buffer(f"relative_position_index_clips_{k}", tb.relative_position_index((WS, WS), (kc, kc)))


class _CffmBlock(nn.Module):
    """cffm_transformer.py:629-707."""

    def __init__(self, dim, heads, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.pool_layers = nn.ModuleList([nn.Linear(WS * WS, 1)])
        self.pool_layers_clips = nn.ModuleList([nn.Linear((WS // l) ** 2, 1) for l in tb.L_CLIPS])
        for m in list(self.pool_layers) + list(self.pool_layers_clips):          # exact mean at init (:678-689)
            m.weight.data.fill_(1.0 / m.in_features)
            m.bias.data.fill_(0)
        self.attn = _WindowAttention3d3(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _BasicLayer3d3(nn.Module):
    def __init__(self, dim, depth, heads=HEADS_N, mlp_ratio=4.):
        super().__init__()
        self.blocks = nn.ModuleList([_CffmBlock(dim, heads, mlp_ratio) for _ in range(depth)])


class _WindowAttentionCluster(nn.Module):
    """swin_transformer_2d.py:157-206."""

    def __init__(self, dim, heads):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * WS - 1) ** 2, heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.register_buffer("relative_position_index", tb.relative_position_index((WS, WS), (WS, WS)))
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        self.qkv_cluster = nn.Linear(dim, dim * 2, bias=True)
        self.proj_cluster = nn.Linear(dim, dim)


class _ClusterBlock(nn.Module):
    def __init__(self, dim, heads, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _WindowAttentionCluster(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _BasicLayerCluster(nn.Module):
    def __init__(self, dim, depth, heads=HEADS_N, mlp_ratio=4.):
        super().__init__()
        self.blocks = nn.ModuleList([_ClusterBlock(dim, heads, mlp_ratio) for _ in range(depth)])


# ------------------------------------------------------------------------------------------------
class BaseDecodeHead_clips_flow(nn.Module):
    """Constructor / attribute contract of decode_head.py:545-589 (only what the CFFM heads use)."""

    def __init__(self, in_channels, channels, *, num_classes, dropout_ratio=0.1, conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type="ReLU"), in_index=-1, input_transform=None,
                 loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0), decoder_params=None,
                 ignore_index=255, sampler=None, align_corners=False, num_clips=5):
        super().__init__()
        self._init_inputs(in_channels, in_index, input_transform)
        self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
        self.conv_cfg, self.norm_cfg, self.act_cfg, self.in_index = conv_cfg, norm_cfg, act_cfg, in_index
        self.loss_decode = build_loss(loss_decode)
        self.ignore_index, self.align_corners, self.num_clips = ignore_index, align_corners, num_clips
        if sampler is not None:
            raise _abi.CffmError("pixel samplers are training-only and out of scope")
        self.sampler = None
        self.conv_seg = nn.Conv2d(channels, num_classes, kernel_size=1)      # unused in forward; kept for the keys
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.fp16_enabled = False

    def _init_inputs(self, in_channels, in_index, input_transform):
        """decode_head.py:598-634."""
        if input_transform is not None:
            assert input_transform in ["resize_concat", "multiple_select"]
        self.input_transform = input_transform
        self.in_index = in_index
        if input_transform is not None:
            assert isinstance(in_channels, (list, tuple))
            assert isinstance(in_index, (list, tuple))
            assert len(in_channels) == len(in_index)
            self.in_channels = sum(in_channels) if input_transform == "resize_concat" else in_channels
        else:
            assert isinstance(in_channels, int)
            assert isinstance(in_index, int)
            self.in_channels = in_channels

    def extra_repr(self):
        return (f"input_transform={self.input_transform}, ignore_index={self.ignore_index}, "
                f"align_corners={self.align_corners}")

    def init_weights(self):
        nn.init.normal_(self.conv_seg.weight, mean=0, std=0.01)
        nn.init.constant_(self.conv_seg.bias, 0)

    def _transform_inputs(self, inputs):
        assert self.input_transform == "multiple_select"
        return [inputs[i] for i in self.in_index]

    def forward_train(self, *a, **k):
        raise NotImplementedError("vss_cffm_b200 covers the inference hot path only (SURVEY.md section 8)")

    def forward_test(self, inputs, img_metas, test_cfg, batch_size=None, num_clips=None, img=None, **kw):
        """decode_head.py:693-708."""
        return self.forward(inputs, batch_size, num_clips, img, **kw)


@HEADS.register_module()
class CFFMHead_clips_resize1_8(BaseDecodeHead_clips_flow):
    WITH_PROTOTYPES = False

    def __init__(self, feature_strides, **kwargs):
        super().__init__(input_transform="multiple_select", **kwargs)
        assert len(feature_strides) == len(self.in_channels)
        assert min(feature_strides) == feature_strides[0]
        self.feature_strides = feature_strides
        c1, c2, c3, c4 = self.in_channels
        decoder_params = kwargs["decoder_params"]
        E = decoder_params["embed_dim"]
        self.embed_dim = E
        self.linear_c4, self.linear_c3 = _MLPEmbed(c4, E), _MLPEmbed(c3, E)
        self.linear_c2, self.linear_c1 = _MLPEmbed(c2, E), _MLPEmbed(c1, E)
        self.linear_fuse = _ConvBN(E * 4, E)
        self.linear_pred = nn.Conv2d(E, self.num_classes, kernel_size=1)
        self.linear_pred2 = nn.Conv2d(E * 2, self.num_classes, kernel_size=1)
        if self.WITH_PROTOTYPES:
            self.linear_pred3 = nn.Conv2d(E, self.num_classes, kernel_size=1)
        self.depths = decoder_params["depths"]
        self.decoder_focal = _BasicLayer3d3(E, self.depths)
        if self.WITH_PROTOTYPES:
            self.n_clusters = 10
            self.save_path = "./cluster_centers/"
            self.dropout3 = nn.Dropout2d(self.dropout_ratio)
            self.decoder_swin = _BasicLayerCluster(E, 1)
            self.finetune = True
        self._plan = None
        self._ws = Workspace()
        self._early, self._early_src = {}, None                  # projections started by project_stage(), and their owner
        self.training = False
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_plan())

    # ------------------------------------------------------------------ module plumbing
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

    # ------------------------------------------------------------------ plan
    def _build_plan(self):
        dev = self.linear_pred.weight.device
        if dev.type != "cuda":
            raise _abi.CffmError("the head runs on a CUDA (sm_100) device only; call .cuda() first")
        _abi.require_device()
        E = self.embed_dim
        if E != 256:
            raise _abi.CffmError(f"the CFFA/CFM kernels are built for embed_dim=256 (all reference configs), got {E}")
        h = lambda t: t.detach().to(dev, _H).contiguous()
        f = lambda t: t.detach().to(dev, _F).contiguous()
        P = {}
        # ---- fold linear_c{i} -> linear_fuse.conv -> BN(eval) into four C_i -> E projections + one shift
        bn = self.linear_fuse.bn
        s = (bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)).to(dev)
        shift = (bn.bias.detach().double().to(dev) - bn.running_mean.detach().double().to(dev) * s)
        Wf = self.linear_fuse.conv.weight.detach().double().to(dev).view(E, 4 * E)
        P["pw"] = []
        for slot, lin in enumerate((self.linear_c4, self.linear_c3, self.linear_c2, self.linear_c1)):   # cat order (:119)
            Wfi = Wf[:, slot * E:(slot + 1) * E] * s[:, None]
            P["pw"].append(h(Wfi @ lin.proj.weight.detach().double().to(dev)))
            shift = shift + Wfi @ lin.proj.bias.detach().double().to(dev)
        P["pw"] = P["pw"][::-1]                                  # index 0 = c1 ... 3 = c4
        P["shift"] = f(shift)
        ncp = _round_up(self.num_classes, NCLS_ALIGN)
        P["ncp"] = ncp

        def cls_w(conv, scale=1.0):
            w = torch.zeros(ncp, conv.weight.shape[1], device=dev, dtype=_H)
            w[:self.num_classes] = (conv.weight.detach().view(self.num_classes, -1) * scale).to(dev, _H)
            b = torch.zeros(ncp, device=dev, dtype=_F)
            b[:self.num_classes] = conv.bias.detach().to(dev, _F) * scale
            return w, b
        P["pred_w"], P["pred_b"] = cls_w(self.linear_pred)
        P["pred2_w"], P["pred2_b"] = cls_w(self.linear_pred2)
        P["blocks"] = []
        for blk in self.decoder_focal.blocks:
            a = blk.attn
            pools = [blk.pool_layers[0]] + list(blk.pool_layers_clips)
            P["blocks"].append(dict(
                n1g=f(blk.norm1.weight), n1b=f(blk.norm1.bias), n1eps=blk.norm1.eps,
                n2g=f(blk.norm2.weight), n2b=f(blk.norm2.bias), n2eps=blk.norm2.eps,
                pool_w=f(torch.cat([p.weight.detach().reshape(-1) for p in pools])),
                pool_b=f(torch.cat([p.bias.detach().reshape(-1) for p in pools])),
                qkv_w=h(a.qkv.weight), qkv_b=f(a.qkv.bias),
                proj_w=h(a.proj.weight), proj_b=f(a.proj.bias),
                bias=tb.assemble_bias_tc(f(a.relative_position_bias_table), f(a.relative_position_bias_table_to_neighbors),
                                         f(a.relative_position_bias_table_to_windows[0]),
                                         [f(t) for t in a.relative_position_bias_table_to_windows_clips],
                                         (E // HEADS_N) ** -0.5, ops.cfm_layout(), HEADS_N),
                f1w=h(blk.mlp.fc1.weight), f1b=f(blk.mlp.fc1.bias), f2w=h(blk.mlp.fc2.weight), f2b=f(blk.mlp.fc2.bias)))
        if self.WITH_PROTOTYPES:
            blk = self.decoder_swin.blocks[0]
            a = blk.attn
            P["swin"] = dict(
                n1g=f(blk.norm1.weight), n1b=f(blk.norm1.bias), n1eps=blk.norm1.eps,
                n2g=f(blk.norm2.weight), n2b=f(blk.norm2.bias), n2eps=blk.norm2.eps,
                q_w=h(a.qkv.weight[:E]), q_b=f(a.qkv.bias[:E]),            # K,V thirds of qkv are dead (:216-227)
                kvc_w=h(a.qkv_cluster.weight), kvc_b=f(a.qkv_cluster.bias),
                pc_w=h(a.proj_cluster.weight), pc_b=f(a.proj_cluster.bias),
                f1w=h(blk.mlp.fc1.weight), f1b=f(blk.mlp.fc1.bias), f2w=h(blk.mlp.fc2.weight), f2b=f(blk.mlp.fc2.bias))
            P["pred3_w"], P["pred3_b"] = cls_w(self.linear_pred3, 0.5)   # x2 + 0.5*x3 (:532), resize is linear
        self._plan = P
        return P

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _as_nhwc16(t):
        """(N,C,H,W) tensor -> fp16 [N,H,W,C] contiguous.  The backbone already hands out NHWC fp16
        memory (zero-copy); anything else (a foreign backbone) is converted with a plain copy."""
        v = t.permute(0, 2, 3, 1)
        if t.dtype == _H and v.is_contiguous():
            return v
        return v.to(_H).contiguous()

    def _load_centers(self, img_metas, batch_size, device):
        """cffm_head.py:429-455: prototypes of each sample's video from <save_path>/<video>/centers.pt."""
        assert batch_size == len(img_metas)
        centers = []
        for meta in img_metas:
            video = meta["filename"].split("/")[-3]
            path = self.save_path + video + "/centers.pt"
            if os.path.isfile(path):
                centers.append(torch.load(path, map_location="cpu"))
            else:
                parts = [torch.load(p, map_location="cpu") for p in sorted(glob.glob(self.save_path + video + "/*.pt"))]
                if not parts:
                    raise FileNotFoundError(f"no prototype file for video '{video}' under {self.save_path}")
                ci = torch.cat(parts, dim=1)
                assert ci.dim() == 3 and ci.shape[0] == 1, ci.shape
                ci = ci.squeeze(0)
                keep = torch.topk(torch.rand(ci.shape[0]), int(ci.shape[0] * 0.8))[1].sort()[0]
                centers.append(ci[keep].unsqueeze(0))
        return torch.cat(centers, dim=0).to(device)

    def forward(self, inputs, batch_size=None, num_clips=None, imgs=None, img_metas=None, *, centers=None,
                frame_major=False):
        """Eval forward (cffm_head.py:99-157 / :423-535).

        inputs: 4 feature maps (N, C_i, H_i, W_i), N = batch_size*num_clips frames in the reference's
        clip-major order, or frame-major when ``frame_major=True`` (what EncoderDecoder_clips of this
        package feeds).  Returns (batch_size, num_classes, H_1, W_1) fp32 logits."""
        lg, (hs, ws_), (h, w) = self.forward_scores(inputs, batch_size, num_clips, img_metas, centers=centers,
                                                    frame_major=frame_major)
        out = torch.empty(batch_size, self.num_classes, h, w, dtype=_F, device=lg.device)
        ops.resize_nhwc_to_nchw(lg, self.num_classes, out, batch_size, hs, ws_, h, w)    # (:149); identity when early
        return out

    def project_stage(self, i, feat):
        """Folded ``linear_c{i+1}`` projection of ONE backbone stage (cffm_head.py:108-117), enqueued on a side stream
        the moment that stage exists: it then overlaps the later backbone stages, which leave most SMs idle.
        ``forward_scores`` picks the result up if it is handed the very same feature map."""
        if feat.device.type != "cuda":
            return
        P = self._plan or self._build_plan()
        if i not in self.in_index:
            return
        slot = list(self.in_index).index(i)
        t = self._as_nhwc16(feat)
        n, h, w, c = t.shape
        out = self._ws.get(f"p{slot}", (n * h * w, self.embed_dim), _H)
        with ops.fork("proj"):
            ops.gemm(t.reshape(-1, c), P["pw"][slot], out16=out)
        self._early[slot] = (t.data_ptr(), tuple(t.shape), out)

    def early_projections_belong_to(self, feature_list):
        """The caller of ``project_stage`` names the list of feature maps the projections were computed from;
        ``forward_scores`` re-uses them only when it is handed that very list object (anything else recomputes)."""
        self._early_src = feature_list

    def forward_scores(self, inputs, batch_size, num_clips, img_metas=None, *, centers=None, frame_major=False):
        """Everything of ``forward`` up to (not including) the last bilinear resize: class scores of the target
        frames, NHWC fp32 ``[B*hs*ws, ncp]`` (first num_classes columns valid), their size (hs, ws) and the size
        (h, w) the reference resizes them to.  EncoderDecoder_clips fuses that resize with its own."""
        assert batch_size is not None and num_clips is not None
        P = self._plan or self._build_plan()
        ws = self._ws
        feats = [self._as_nhwc16(t) for t in self._transform_inputs(inputs)]
        N = feats[0].shape[0]
        if N != batch_size * num_clips:
            raise _abi.CffmError(f"got {N} frames for batch_size={batch_size} x num_clips={num_clips}")
        E, ncls, ncp = self.embed_dim, self.num_classes, P["ncp"]
        sizes = [(t.shape[1], t.shape[2]) for t in feats]
        for t, cin in zip(feats, self.in_channels):
            assert t.shape[3] == cin, (tuple(t.shape), cin)
        B, T = batch_size, num_clips
        h, w = sizes[0]
        t_perm = 0 if (frame_major or T == 1) else T
        # ---- per-frame MLP decoder, folded (:108-119)
        proj = [ws.get(f"p{i}", (N * sizes[i][0] * sizes[i][1], E), _H) for i in range(4)]
        owned = inputs is self._early_src and self._early_src is not None
        early_done = {i for i, (ptr, shp, buf) in self._early.items()
                      if owned and ptr == feats[i].data_ptr() and shp == tuple(feats[i].shape) and buf.data_ptr() == proj[i].data_ptr()}
        if self._early:
            ops.join("proj")                                     # whatever project_stage() started must finish before p{i} is re-used
        self._early, self._early_src = {}, None
        todo = [i for i in range(4) if i not in early_done]
        if todo:
            with ops.fork():                                     # the small projections beside the big one
                for i in todo[1:]:
                    ops.gemm(feats[i].reshape(-1, feats[i].shape[3]), P["pw"][i], out16=proj[i])
            ops.gemm(feats[todo[0]].reshape(-1, feats[todo[0]].shape[3]), P["pw"][todo[0]], out16=proj[todo[0]])
            ops.join()
        early = num_clips != self.num_clips                      # eval-mode early return (:127-129)
        if early:
            c_full = ws.get("c_full", (N * h * w, E), _H)
            ops.head_fuse(proj, sizes, N, E, t_perm, P["shift"], c_full=c_full)
            ct = c_full[(T - 1) * B * h * w:]                    # frame-major: target frames are last
            lg = ws.get("lg_full", (B * h * w, ncp), _F)
            ops.gemm(ct, P["pred_w"], bias=P["pred_b"], out32=lg)
            return lg, (h, w), (h, w)
        if num_clips != 4:
            raise _abi.CffmError("CFFM needs exactly 3 reference frames + 1 target (focal_l_clips=[1,2,3], "
                                 f"cffm_head.py:93-94); got num_clips={num_clips}")
        if h % 2 or w % 2:
            raise _abi.CffmError(f"1/4-scale feature size must be even (got {h}x{w}): the fused half-resolution "
                                 "resize is an exact 2x2 mean only then")
        h2, w2 = h // 2, w // 2
        HW = h2 * w2
        x32 = ws.get("x32", (N * HW, E), _F)                     # [T,B,h2,w2,E] fp32 = _c_further
        c16 = ws.get("c16", (N * HW, E), _H)
        ops.head_fuse(proj, sizes, N, E, t_perm, P["shift"], half32=x32, half16=c16)
        xt = x32[(T - 1) * B * HW:]                              # target residual stream, updated in place
        ct16 = c16[(T - 1) * B * HW:]
        xt0 = None
        if self.WITH_PROTOTYPES:
            xt0 = ws.get("xt0", (B * HW, E), _F)
            xt0.copy_(xt)                                        # decoder_swin starts from the pre-CFFM target (:519)
        # ---- CFFM blocks (:138)
        Hp, Wp = _round_up(h2, WS), _round_up(w2, WS)
        nW = (Hp // WS) * (Wp // WS)
        Pp = 15 * nW
        nref = (T - 1) * B                                       # frame-major: reference frames first, targets last
        xn_t = ws.get("xn_t", (B * HW, E), _H)
        na = ops.apron_rows(B, h2, w2)                           # target map with its cyclic apron [B, Hp+6, Wp+6]
        xt_pad = ws.get("xt_pad", (na, E), _H)                   # every row (pad zeros included) is rewritten per call
        qkv_t = ws.get("qkv_t", (na, 3 * E), _H)
        ao = ws.get("ao", (B * HW, E), _H)
        xn2 = ws.get("xn2", (B * HW, E), _H)
        hid = ws.get("hid", (B * HW, 4 * E), _H)
        xt16 = ws.get("xt16", (B * HW, E), _H)
        depth = len(P["blocks"])

        def assemble_refs(i):
            """Reference-frame side of CFFA for block i (norm1 -> pad -> resize -> fc-pool, :713-805).  A block never
            modifies the reference frames (:826), so this does not depend on the previous block: it runs on its own
            stream beside the target path and is joined just before the pooled K/V projection."""
            b = P["blocks"][i]
            xn_r = ws.get(f"xn_r{i}", (nref * HW, E), _H)
            pooled = ws.get(f"pooled{i}", (B * Pp, E), _H)
            with ops.fork("refs"):
                ops.cffa_norm_frames(x32[:nref * HW], b["n1g"], b["n1b"], b["n1eps"], xn_r, None, nref, nref, h2, w2, Hp, Wp, E)
                ops.cffa_pool_part(xn_r, B, 1, h2, w2, E, b["pool_w"], b["pool_b"], pooled)

        assemble_refs(0)
        for i, b in enumerate(P["blocks"]):
            pooled = ws.get(f"pooled{i}", (B * Pp, E), _H)
            kvp = ws.get(f"kvp{i}", (B * Pp, 2 * E), _H)
            ops.cffa_norm_frames(xt, b["n1g"], b["n1b"], b["n1eps"], xn_t, xt_pad, B, 0, h2, w2, Hp, Wp, E)
            with ops.fork():                                     # target-level pooling + pooled K/V projection beside the target QKV
                ops.cffa_pool_part(xn_t, B, 0, h2, w2, E, b["pool_w"], b["pool_b"], pooled)
                ops.join("refs")                                 # reference levels of this block (started one block earlier)
                ops.gemm(pooled, b["qkv_w"][E:], bias=b["qkv_b"][E:], out16=kvp)   # Q third is dead work (:449)
            if i + 1 < depth:
                assemble_refs(i + 1)
            ops.gemm(xt_pad, b["qkv_w"], bias=b["qkv_b"], out16=qkv_t)
            ops.join()
            ops.cfm_attention(qkv_t, kvp, b["bias"], ao, B, h2, w2, E, HEADS_N, (E // HEADS_N) ** -0.5)
            ops.gemm(ao, b["proj_w"], bias=b["proj_b"], residual=xt, out32=xt)
            ops.layernorm(xt, b["n2g"], b["n2b"], b["n2eps"], out16=xn2)
            ops.gemm(xn2, b["f1w"], bias=b["f1b"], out16=hid, act=ops.ACT_GELU)
            last = i == depth - 1
            ops.gemm(hid, b["f2w"], bias=b["f2b"], residual=xt, out32=xt, out16=xt16 if last else None)
        # ---- linear_pred2 on cat([_c_further[:,-1], _c2[:,-1]]) without the concat (:145-148)
        lg = ws.get("lg", (B * HW, ncp), _F)
        ops.gemm(ct16, P["pred2_w"][:, :E], bias=P["pred2_b"], out32=lg)
        ops.gemm(xt16, P["pred2_w"][:, E:], residual=lg, out32=lg)
        if self.WITH_PROTOTYPES:
            if centers is None:
                centers = self._load_centers(img_metas, batch_size, lg.device)
            self._cluster_branch(P, xt0, centers, lg, B, HW)
        return lg, (h2, w2), (h, w)

    def _cluster_branch(self, P, xt0, centers, lg, B, HW):
        """decoder_swin + linear_pred3, accumulated into ``lg`` with weight 0.5 (cffm_head.py:519-532;
        swin_transformer_2d.py:208-262, :605-665).  The window partition around the attention is a
        per-token no-op (no positional bias, no mask on this branch)."""
        ws, S, E = self._ws, P["swin"], self.embed_dim
        assert centers.dim() == 3 and centers.shape[0] == B and centers.shape[2] == E, tuple(centers.shape)
        K = centers.shape[1]
        cen = centers.to(lg.device, _F).contiguous().view(B * K, E)
        tn = ws.get("sw.tn", (B * HW, E), _H)
        cn = ws.get("sw.cn", (B * K, E), _H)
        ops.layernorm(xt0, S["n1g"], S["n1b"], S["n1eps"], out16=tn)
        ops.layernorm(cen, S["n1g"], S["n1b"], S["n1eps"], out16=cn)         # same norm on the prototypes (:622)
        q = ws.get("sw.q", (B * HW, E), _H)
        ops.gemm(tn, S["q_w"], bias=S["q_b"], out16=q)
        kv = ws.get("sw.kv", (B * K, 2 * E), _H)
        ops.gemm(cn, S["kvc_w"], bias=S["kvc_b"], out16=kv)
        ao = ws.get("sw.ao", (B * HW, E), _H)
        d = E // HEADS_N
        ops.mha(q, kv[:, :E], kv[:, E:], ao, B, HW, K, HEADS_N, d, d ** -0.5)
        ops.gemm(ao, S["pc_w"], bias=S["pc_b"], residual=xt0, out32=xt0)
        ops.layernorm(xt0, S["n2g"], S["n2b"], S["n2eps"], out16=tn)
        hid = ws.get("sw.hid", (B * HW, 4 * E), _H)
        ops.gemm(tn, S["f1w"], bias=S["f1b"], out16=hid, act=ops.ACT_GELU)
        c3 = ws.get("sw.c3", (B * HW, E), _H)
        ops.gemm(hid, S["f2w"], bias=S["f2b"], residual=xt0, out16=c3)
        ops.gemm(c3, P["pred3_w"], bias=P["pred3_b"], residual=lg, out32=lg)


@HEADS.register_module()
class CFFMHead_clips_resize1_8_finetune_w_prototype3(CFFMHead_clips_resize1_8):
    """CFFM++ (cffm_head.py:303-535): CFFM + cross-attention to k-means prototypes of the video."""
    WITH_PROTOTYPES = True

    def forward_test(self, inputs, img_metas, test_cfg, batch_size=None, num_clips=None, img=None, **kw):
        return self.forward(inputs, batch_size, num_clips, img, img_metas, **kw)


@HEADS.register_module()
class CFFMHead_clips_resize1_8_gene_prototype(CFFMHead_clips_resize1_8):
    """Prototype generation for CFFM++ (cffm_head.py:161-300): the MLP decoder on every frame, ``linear_pred`` logits of
    the last frame as the return value, and -- the point of this head -- k-means (K = 100, 10 Lloyd iterations) on the
    1/8-scale decoder features of the clip, saved as ``<save_path>/<video>/centers.pt`` (a (1, K, 256) fp32 tensor),
    exactly where ``CFFMHead_clips_resize1_8_finetune_w_prototype3`` reads them.  Same parameters / state-dict keys as
    the CFFM head (``decoder_focal`` and ``linear_pred2`` exist but are not used by this forward, as in the reference)."""

    EARLY_RETURN_LAST_FRAME = False          # every frame handed in is decoded and clustered (no eval early return, :239-300)
    FUSED_TAIL = False                       # the segmentor calls forward_test (k-means + file output), not forward_scores

    def __init__(self, feature_strides, **kwargs):
        super().__init__(feature_strides, **kwargs)
        self.n_clusters = 100                                     # cffm_head.py:217
        self.save_path = "./cluster_centers/"
        self.kmeans_max_iter = 10                                 # cffm_head.py:280

    def forward_test(self, inputs, img_metas, test_cfg, batch_size=None, num_clips=None, img=None, **kw):
        return self.forward(inputs, batch_size, num_clips, img, img_metas, **kw)

    def cluster_features(self, inputs, batch_size, num_clips, frame_major=False):
        """(logits of the last frame NHWC fp32 [B*h*w, ncp], 1/8-scale features fp16 [T, B, h2*w2, E], (h, w))."""
        P = self._plan or self._build_plan()
        ws = self._ws
        feats = [self._as_nhwc16(t) for t in self._transform_inputs(inputs)]
        N = feats[0].shape[0]
        if N != batch_size * num_clips:
            raise _abi.CffmError(f"got {N} frames for batch_size={batch_size} x num_clips={num_clips}")
        E, ncp = self.embed_dim, P["ncp"]
        sizes = [(t.shape[1], t.shape[2]) for t in feats]
        h, w = sizes[0]
        if h % 2 or w % 2:
            raise _abi.CffmError(f"1/4-scale feature size must be even (got {h}x{w})")
        B, T = batch_size, num_clips
        if self._early:
            ops.join("proj")
        self._early, self._early_src = {}, None
        proj = [ws.get(f"p{i}", (N * sizes[i][0] * sizes[i][1], E), _H) for i in range(4)]
        for i in range(4):
            ops.gemm(feats[i].reshape(-1, feats[i].shape[3]), P["pw"][i], out16=proj[i])
        t_perm = 0 if (frame_major or T == 1) else T
        c_full = ws.get("c_full", (N * h * w, E), _H)
        HW = (h // 2) * (w // 2)
        c16 = ws.get("c16", (N * HW, E), _H)
        ops.head_fuse(proj, sizes, N, E, t_perm, P["shift"], c_full=c_full, half16=c16)
        lg = ws.get("lg_full", (B * h * w, ncp), _F)
        ops.gemm(c_full[(T - 1) * B * h * w:], P["pred_w"], bias=P["pred_b"], out32=lg)   # x[:, -1] is all that is returned (:299)
        return lg, c16.view(T, B, HW, E), (h, w)

    def forward(self, inputs, batch_size=None, num_clips=None, imgs=None, img_metas=None, *, frame_major=False, save=True):
        from .kmeans import KMeans
        assert batch_size == 1, "prototype generation runs one video clip at a time (cffm_head.py:269)"
        lg, feats, (h, w) = self.cluster_features(inputs, batch_size, num_clips, frame_major)
        centers = []
        for ii in range(batch_size):                             # clip ii: all its frames' 1/8-scale pixels (:273-283)
            km = KMeans(n_clusters=self.n_clusters, max_iter=self.kmeans_max_iter, mode="euclidean", verbose=0)
            km.fit_predict(feats[:, ii].reshape(-1, self.embed_dim))
            centers.append(km.centroids)
        self.centers = torch.stack(centers, dim=0)               # (B, K, E) fp32
        if save and img_metas is not None:
            video = img_metas[0]["filename"].split("/")[-3]
            path = self.save_path + video
            os.makedirs(path, exist_ok=True)
            torch.save(self.centers.cpu(), path + "/centers.pt")
        out = torch.empty(batch_size, self.num_classes, h, w, dtype=_F, device=lg.device)
        ops.resize_nhwc_to_nchw(lg, self.num_classes, out, batch_size, h, w, h, w)
        return out
