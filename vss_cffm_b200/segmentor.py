"""``EncoderDecoder_clips`` -- the caller of the hot path, same surface as the reference
(mmseg/models/segmentors/encoder_decoder.py:295-591 on base.py:14-303), inference only.

``model(img=[[T x (B,3,H,W)]], img_metas=[[dict x B]], return_loss=False)`` -> list of B
``np.ndarray (H,W) int64`` label maps.  Differences in HOW (not what) it computes:

  * frames are stacked frame-major ((T,B,...) instead of (B,T,...), :556-560) so that the B target
    frames are contiguous for the CFFM kernels; per-clip results are unchanged;
  * when the head early-returns (eval and T != head.num_clips, cffm_head.py:127-129) only the last
    frame influences the output, so only the last frame is pushed through the backbone;
  * final bilinear resize + softmax + argmax (:373-377, :542, :564) is one kernel (softmax is
    monotone, so the argmax is taken on the interpolated logits).
"""
import os

import torch
import torch.nn as nn

from . import _abi, ops
from . import registry as builder
from .registry import SEGMENTORS
from .workspace import Workspace

_F = torch.float32


@SEGMENTORS.register_module()
class EncoderDecoder_clips(nn.Module):
    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None):
        super().__init__()
        self.backbone = builder.build_backbone(backbone)
        if neck is not None:
            raise _abi.CffmError("necks are not used by any CFFM config and are out of scope")
        if auxiliary_head is not None:
            raise _abi.CffmError("auxiliary heads are training-only and out of scope")
        self.decode_head = builder.build_head(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        if self.align_corners:
            raise _abi.CffmError("align_corners=True is not used by any CFFM config and is not implemented")
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.fp16_enabled = False
        self.init_weights(pretrained=pretrained)
        assert self.with_decode_head
        self._ws = Workspace()
        for m in self.modules():
            m.training = False
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_graphs())

    def _apply(self, fn, *a, **k):
        self.invalidate_graphs()                                 # .cuda() / .to(): the captured passes read the old tensors
        return super()._apply(fn, *a, **k)

    def workspaces(self):
        """Every Workspace a forward pass of this model touches (CUDA-graph owners pin them)."""
        out = [self._ws]
        for m in (self.backbone, self.decode_head):
            if hasattr(m, "_ws"):
                out.append(m._ws)
        return out

    # ------------------------------------------------------------------ BaseSegmentor surface
    @property
    def with_neck(self):
        return False

    @property
    def with_auxiliary_head(self):
        return False

    @property
    def with_decode_head(self):
        return hasattr(self, "decode_head") and self.decode_head is not None

    def init_weights(self, pretrained=None):
        self.backbone.init_weights(pretrained=pretrained)
        self.decode_head.init_weights()

    def train(self, mode=True):
        if mode:
            raise _abi.CffmError("vss_cffm_b200 implements the inference hot path only (eval mode); "
                                 "training is out of scope (SURVEY.md section 8)")
        return super().train(False)

    def forward_train(self, *a, **k):
        raise NotImplementedError("vss_cffm_b200 covers the inference hot path only (SURVEY.md section 8)")

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        """base.py:135-149."""
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def forward_test(self, imgs, img_metas, **kwargs):
        """base.py:76-111."""
        for var, name in [(imgs, "imgs"), (img_metas, "img_metas")]:
            if not isinstance(var, list):
                raise TypeError(f"{name} must be a list, but got {type(var)}")
        if len(imgs) != len(img_metas):
            raise ValueError(f"num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})")
        for img_meta in img_metas:
            for key in ("ori_shape", "img_shape", "pad_shape"):
                shapes = [m[key] for m in img_meta]
                assert all(s == shapes[0] for s in shapes)
        if len(imgs) == 1:
            return self.simple_test(imgs[0], img_metas[0], **kwargs)
        return self.aug_test(imgs, img_metas, **kwargs)

    def aug_test(self, imgs, img_metas, rescale=True):
        raise NotImplementedError("aug_test cannot run in the reference either (wrong arity, "
                                  "encoder_decoder.py:582 vs :518)")

    # ------------------------------------------------------------------ the hot path
    def extract_feat(self, img):
        """encoder_decoder.py:323-327.  When both sides are this package's modules the head projects every stage
        output as soon as the backbone has produced it (beside the later, smaller stages)."""
        hook = getattr(self.decode_head, "project_stage", None)
        if hook is not None and hasattr(self.backbone, "forward_features"):
            outs = self.backbone(img, stage_hook=hook)
            self.decode_head.early_projections_belong_to(outs)  # valid for exactly this list object
            return outs
        return self.backbone(img)

    def _device(self):
        return next(self.parameters()).device

    def _stack(self, img):
        """list of T (B,3,H,W) -> (frames (T,B,3,H,W) frame-major on the model's device, B, T)."""
        if not isinstance(img, (list, tuple)):
            if img.dim() != 5:
                raise _abi.CffmError("expected a list of T (B,3,H,W) tensors or a (B,T,3,H,W) tensor")
            img = list(img.unbind(1))                            # (B,T,3,H,W) as in forward_train
        T, B = len(img), img[0].shape[0]
        dev = self._device()
        if dev.type != "cuda":
            raise _abi.CffmError("the model runs on a CUDA (sm_100) device only; call .cuda() first (no CPU fallback)")
        frames = self._ws.get("frames", (T, B) + tuple(img[0].shape[1:]), _F, device=dev)
        for t, f in enumerate(img):                              # one (async, if pinned) copy per frame stack
            frames[t].copy_(f, non_blocking=True)
        return frames, B, T

    def _features(self, frames, B, T):
        """Backbone on frame-major frames; on the head's early-return path (eval and T != num_clips,
        cffm_head.py:127-129) only the target frame influences the output, so only it is encoded."""
        if T != self.decode_head.num_clips and getattr(self.decode_head, "EARLY_RETURN_LAST_FRAME", True):
            return self.extract_feat(frames[-1]), 1
        return self.extract_feat(frames.reshape(T * B, *frames.shape[2:])), T

    def encode_decode_frames(self, frames, img_metas, B, T, **head_kw):
        """Backbone + head on frame-major frames (T,B,3,H,W) -> (B, num_classes, H/4, W/4) logits."""
        x, t = self._features(frames, B, T)
        return self.decode_head.forward_test(x, img_metas, self.test_cfg, B, t, frame_major=True, **head_kw)

    def encode_decode(self, img, img_metas, batch_size, num_clips):
        """Reference signature (:367-378): img (B*T,3,H,W) clip-major -> logits resized to (H,W)."""
        H, W = img.shape[2:]
        frames = img.reshape(batch_size, num_clips, *img.shape[1:]).transpose(0, 1).contiguous()
        out = self.encode_decode_frames(frames.to(self._device(), _F), img_metas, batch_size, num_clips)
        full = torch.empty(batch_size, self.num_classes, H, W, dtype=_F, device=out.device)
        ops.resize_nchw(out, full)
        return full

    def whole_inference(self, img, img_meta, rescale, batch_size, num_clips):
        seg_logit = self.encode_decode(img, img_meta, batch_size, num_clips)
        ori = tuple(img_meta[0]["ori_shape"][:2])
        if rescale and ori != tuple(seg_logit.shape[2:]):
            out = torch.empty(*seg_logit.shape[:2], *ori, dtype=_F, device=seg_logit.device)
            ops.resize_nchw(seg_logit, out)
            seg_logit = out
        return seg_logit

    def inference(self, img, img_meta, rescale, batch_size, num_clips):
        """Reference signature (:518-552): softmax probabilities (B, num_classes, H, W)."""
        assert self.test_cfg["mode"] in ["slide", "whole"]
        if self.test_cfg["mode"] == "slide":
            raise _abi.CffmError("slide inference is not used by any CFFM config (test_cfg.mode='whole')")
        seg_logit = self.whole_inference(img, img_meta, rescale, batch_size, num_clips)
        out = torch.empty_like(seg_logit)
        ops.softmax_nchw(seg_logit, out)
        return self._flip(out, img_meta)

    @staticmethod
    def _flip(t, img_meta):
        if img_meta[0].get("flip", False):
            d = img_meta[0]["flip_direction"]
            assert d in ["horizontal", "vertical"]
            t = t.flip(dims=(-1,) if d == "horizontal" else (-2,))
        return t

    def predict_labels(self, img, img_meta, rescale=True, **head_kw):
        """Device-side part of simple_test: int64 label maps (B,H,W) on the GPU."""
        frames, B, T = self._stack(img)
        return self.labels_from_frames(frames, img_meta, rescale, **head_kw)

    def labels_from_frames(self, frames, img_meta, rescale=True, label_dtype=torch.int64, **head_kw):
        """frames: (T,B,3,H,W) fp32 on the device, frame-major -> labels (B,H,W), int64 like the reference's argmax or
        (``label_dtype=torch.uint8``, at most 256 classes) 8-bit: an eighth of the bytes a serving loop reads back.  Pure
        sequence of C-ABI launches on the current stream (no host sync, stable workspace addresses): CUDA-graph capturable."""
        assert label_dtype in (torch.int64, torch.uint8) and (label_dtype == torch.int64 or self.num_classes <= 256)
        T, B = frames.shape[:2]
        H, W = frames.shape[-2:]
        ori = tuple(img_meta[0]["ori_shape"][:2])
        assert all(tuple(m["ori_shape"][:2]) == ori for m in img_meta)
        head = self.decode_head
        fused = (not (rescale and ori != (H, W))) and hasattr(head, "forward_scores") and getattr(head, "FUSED_TAIL", True)
        if fused:
            x, t = self._features(frames, B, T)
            if "img_metas" not in head_kw:
                head_kw = dict(head_kw, img_metas=img_meta)
            scores, (hs, ws_), (h, w) = head.forward_scores(x, B, t, frame_major=True, **head_kw)
            if ops.upsample2_argmax_supported(h, w, H, W):       # head's x2 resize + the x4 resize + argmax: one kernel
                labels = torch.empty(B, H, W, dtype=label_dtype, device=scores.device)
                ops.upsample2_argmax(scores, self.num_classes, labels, B, hs, ws_, h, w, H, W)
                return self._flip(labels, img_meta)
            logits = torch.empty(B, self.num_classes, h, w, dtype=_F, device=scores.device)
            ops.resize_nhwc_to_nchw(scores, self.num_classes, logits, B, hs, ws_, h, w)
        else:
            logits = self.encode_decode_frames(frames, img_meta, B, T, **head_kw)     # (B,ncls,h,w)
            h, w = logits.shape[2:]
        if rescale and ori != (H, W):                            # two chained resizes (:373-377 then :507-514)
            mid = torch.empty(B, self.num_classes, H, W, dtype=_F, device=logits.device)
            ops.resize_nchw(logits, mid)
            logits, h, w, H, W = mid, H, W, ori[0], ori[1]
        labels = torch.empty(B, H, W, dtype=torch.int64, device=logits.device)
        ops.resize_argmax(logits, labels, B, self.num_classes, h, w, H, W)
        return self._flip(labels, img_meta).to(label_dtype)

    def make_graphed(self, B, T, H, W, img_meta=None, rescale=True, **head_kw):
        """Capture one inference pass for fixed (B,T,H,W) in a CUDA graph: the ~130 kernel launches of a step are
        replayed with one cudaGraphLaunch, removing the host launch overhead.  Returns a ``GraphedClips``."""
        from .graph import GraphedClips
        return GraphedClips(self, B, T, H, W, img_meta, rescale, head_kw)

    # ------------------------------------------------------------------ CUDA-graph cache behind the mmseg call
    GRAPH_CACHE_SIZE = 4                                          # geometries kept captured (least recently used is dropped)
    GRAPH_CACHE_AFTER = 2                                         # capture on the n-th call with the same geometry

    def _cached_graph(self, img, img_meta, rescale, head_kw):
        """The captured pass for this call's geometry, or None (eager).  ``model(img=..., return_loss=False)`` launches
        ~100 kernels; from the second call with the same (B, T, H, W, ori_shape, flip) on, the pass is captured once
        (``GraphedClips``) and every later call is one H2D copy per frame stack + one cudaGraphLaunch.  Heads that read
        per-call host state (CFFM++ prototypes from files, prototype generation) always run eagerly."""
        head = self.decode_head
        if (not getattr(self, "graph_cache", True) or head_kw or getattr(head, "WITH_PROTOTYPES", False) or
                not getattr(head, "FUSED_TAIL", True) or os.environ.get("CFFM_GRAPH_CACHE", "1") == "0"):
            return None
        frames = img if isinstance(img, (list, tuple)) else list(img.unbind(1))
        T, (B, C, H, W) = len(frames), tuple(frames[0].shape)
        m0 = img_meta[0]
        key = (B, T, C, H, W, bool(rescale), tuple(m0["ori_shape"][:2]), bool(m0.get("flip", False)), m0.get("flip_direction"))
        cache = self.__dict__.setdefault("_graphs", {})
        seen = self.__dict__.setdefault("_graph_seen", {})
        g = cache.get(key)
        if g is None:
            seen[key] = seen.get(key, 0) + 1
            if seen[key] < self.GRAPH_CACHE_AFTER:
                return None
            if len(seen) > 64:
                seen.clear()
            if len(cache) >= self.GRAPH_CACHE_SIZE:
                cache.pop(next(iter(cache)))
            g = self.make_graphed(B, T, H, W, [dict(m) for m in img_meta], rescale)
        else:
            cache.pop(key)                                       # re-insert: most recently used last
        cache[key] = g
        return g

    def invalidate_graphs(self):
        """Drop every captured pass (new weights change the folded plans the graphs read)."""
        self.__dict__.pop("_graphs", None)
        self.__dict__.pop("_graph_seen", None)

    def simple_test(self, img, img_meta, rescale=True, **head_kw):
        """encoder_decoder.py:554-572: list of B (H,W) int64 numpy label maps."""
        g = self._cached_graph(img, img_meta, rescale, head_kw)
        if g is not None:
            labels = g(img if isinstance(img, (list, tuple)) else list(img.unbind(1)))
        else:
            labels = self.predict_labels(img, img_meta, rescale, **head_kw)
        return list(labels.cpu().numpy())
