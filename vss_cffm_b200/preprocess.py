"""Test-time clip preprocessing on the GPU, bit-exact with the reference's CPU pipeline
(``AlignedResize_clips(keep_ratio=True, size_divisor=32)`` -> ``Normalize_clips`` -> ``ImageToTensor_clips``,
mmseg/datasets/pipelines/transforms.py:382-421, :1277-1297; config ``test_pipeline`` in local_configs/cffm/**).

Input: decoded frames, uint8 BGR HWC (what ``LoadImageFromFile`` / cv2.imread produce).  Output: the frame-major
(T, B, 3, H, W) fp32 device tensor ``EncoderDecoder_clips.labels_from_frames`` consumes, and the ``img_metas`` dicts.
Uploading uint8 frames moves a quarter of the bytes of the fp32 tensors the reference builds on the CPU.
"""
import ctypes

import numpy as np
import torch

from . import _abi, ops


def rescale_size(old_size, scale):
    """mmcv.rescale_size: (w, h), (long, short) -> (new_w, new_h)."""
    w, h = old_size
    f = min(max(scale) / max(h, w), min(scale) / min(h, w))
    return int(w * float(f) + 0.5), int(h * float(f) + 0.5)


def _as_u8_batch(imgs, dev):
    """list of HWC uint8 arrays / tensors (same size) or one (N,h,w,3) tensor -> contiguous CUDA uint8 (N,h,w,3)."""
    if isinstance(imgs, (list, tuple)):
        ts = [torch.from_numpy(np.ascontiguousarray(i)) if isinstance(i, np.ndarray) else i for i in imgs]
        x = torch.stack([t.to(dev, non_blocking=True) for t in ts])
    else:
        x = imgs
    if not x.is_cuda:
        raise _abi.CffmError("preprocess: expected CUDA tensors (or host frames plus a device); no CPU fallback exists")
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3:
        raise _abi.CffmError(f"preprocess: expected uint8 (N,h,w,3) frames, got {x.dtype} {tuple(x.shape)}")
    return x.contiguous()


def resize_u8(imgs, H, W):
    """cv2.resize(INTER_LINEAR) of uint8 (N,h,w,3) CUDA frames -> (N,H,W,3), bit-exact."""
    x = _as_u8_batch(imgs, None)
    N, h, w, _ = x.shape
    out = torch.empty(N, H, W, 3, dtype=torch.uint8, device=x.device)
    _abi.call("cffm_resize_u8", x.data_ptr(), N, h, w, out.data_ptr(), H, W, ops._stream())
    return out


def resize_normalize(imgs, H, W, mean, std, to_rgb=True, out=None):
    """resize (identity for equal sizes) + mmcv.imnormalize + HWC->CHW: uint8 (N,h,w,3) -> fp32 (N,3,H,W)."""
    x = _as_u8_batch(imgs, None)
    N, h, w, _ = x.shape
    if out is None:
        out = torch.empty(N, 3, H, W, dtype=torch.float32, device=x.device)
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape[-3:]) == (3, H, W) and out.shape[0] == N
    assert out.stride(-1) == 1 and out.stride(-2) == W and out.stride(-3) == H * W
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    _abi.call("cffm_resize_normalize_u8", x.data_ptr(), N, h, w, out.data_ptr(), out.stride(0), H, W,
              ctypes.cast(m, ctypes.c_void_p), ctypes.cast(s, ctypes.c_void_p), int(bool(to_rgb)), ops._stream())
    return out


class ClipPreprocessor:
    """``pp(clips)``: clips = list of B clips, each a list of T decoded frames (uint8 BGR HWC, numpy or torch, all of one
    size) -> ((T, B, 3, H, W) fp32 CUDA tensor, list of B img_meta dicts)."""

    def __init__(self, img_scale=(853, 480), size_divisor=32, mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375),
                 to_rgb=True, device=None):
        self.img_scale, self.size_divisor = tuple(img_scale), size_divisor
        self.mean, self.std, self.to_rgb = tuple(mean), tuple(std), to_rgb
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def output_size(self, h, w):
        rw, rh = rescale_size((w, h), self.img_scale)
        d = self.size_divisor
        return (rh, rw), (int(np.ceil(rh / d)) * d, int(np.ceil(rw / d)) * d)

    def run(self, x, T, B, out=None):
        """Device part (no host work, CUDA-graph capturable): x uint8 (T*B, h, w, 3) CUDA, frame-major -> (T,B,3,H,W) fp32."""
        h, w = x.shape[1:3]
        (rh, rw), (H, W) = self.output_size(h, w)
        if (rh, rw) != (h, w):
            x = resize_u8(x, rh, rw)                             # mmcv.imrescale (transforms.py:399-400)
        if out is None:
            out = torch.empty(T, B, 3, H, W, dtype=torch.float32, device=x.device)
        resize_normalize(x, H, W, self.mean, self.std, self.to_rgb, out=out.view(T * B, 3, H, W))   # _align + Normalize_clips
        return out

    def metas(self, B, h, w, filenames=None):
        (_, _), (H, W) = self.output_size(h, w)
        sf = np.array([W / w, H / h, W / w, H / h], dtype=np.float32)
        return [dict(ori_shape=(h, w, 3), img_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=sf, flip=False,
                     keep_ratio=True, filename=(filenames[b] if filenames else f"data/video{b}/origin/00000000.jpg"),
                     img_norm_cfg=dict(mean=np.array(self.mean, np.float32), std=np.array(self.std, np.float32), to_rgb=self.to_rgb))
                for b in range(B)]

    def __call__(self, clips, filenames=None, out=None):
        _abi.require_device()
        B, T = len(clips), len(clips[0])
        h, w = clips[0][0].shape[:2]
        x = _as_u8_batch([clips[b][t] for t in range(T) for b in range(B)], self.device)   # frame-major
        return self.run(x, T, B, out), self.metas(B, h, w, filenames)
