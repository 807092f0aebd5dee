"""CPU restatement (test infrastructure only) of the test-time clip preprocessing that feeds the hot path
(everything after JPEG decoding):

  AlignedResize_clips(keep_ratio=True, size_divisor=32)   mmseg/datasets/pipelines/transforms.py:382-421
      mmcv.imrescale(img, scale)  ->  _align: mmcv.imresize(img, (align_w, align_h))     (both cv2.resize INTER_LINEAR on uint8)
  Normalize_clips(mean, std, to_rgb=True)                 transforms.py:1277-1297 -> mmcv.imnormalize
  ImageToTensor_clips                                     HWC -> CHW

The arithmetic lives in third-party code: mmcv 1.3.0 (pinned by README.md:34 / mmseg/__init__.py:5-6, absent here) whose
imrescale / imresize / imnormalize are thin wrappers over OpenCV, and OpenCV itself (cv2 4.13 IS installed in this image).
``resize_u8`` restates cv2::resize INTER_LINEAR for 8-bit images (11-bit fixed-point coefficients, horizontal pass to int32,
vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2) and ``normalize`` restates mmcv.imnormalize_
(cvtColor BGR2RGB, cv2.subtract with the fp32 mean, cv2.multiply with the float64 1/std).  Pinned:
tests/test_preprocess.py checks both bit for bit against cv2 itself on random images (the oracle is exact, not approximate).
"""
import numpy as np


def rescale_size(old_size, scale):
    """mmcv.image.geometric.rescale_size for a (long, short) tuple scale.  old_size = (w, h) -> (new_w, new_h)."""
    w, h = old_size
    max_long_edge, max_short_edge = max(scale), min(scale)
    f = min(max_long_edge / max(h, w), max_short_edge / min(h, w))
    return int(w * float(f) + 0.5), int(h * float(f) + 0.5)


def aligned_size(w, h, size_divisor=32):
    """AlignedResize_clips._align (transforms.py:382-384)."""
    return int(np.ceil(w / size_divisor)) * size_divisor, int(np.ceil(h / size_divisor)) * size_divisor


def _axis_x(src, dst):
    idx = np.zeros(dst, np.int32); a0 = np.zeros(dst, np.int32); a1 = np.zeros(dst, np.int32)
    scale = src / dst
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f)); f = np.float32(f - s)
        if s < 0:
            f, s = np.float32(0), 0
        if s >= src - 1:
            f, s = np.float32(0), src - 1
        a0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * 2048)))
        a1[d] = int(np.rint(np.float32(f * 2048)))
        idx[d] = s
    return idx, a0, a1


def _axis_y(src, dst):
    i0 = np.zeros(dst, np.int32); i1 = np.zeros(dst, np.int32); b0 = np.zeros(dst, np.int32); b1 = np.zeros(dst, np.int32)
    scale = src / dst
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f)); f = np.float32(f - s)
        b0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * 2048)))
        b1[d] = int(np.rint(np.float32(f * 2048)))
        i0[d], i1[d] = min(max(s, 0), src - 1), min(max(s + 1, 0), src - 1)
    return i0, i1, b0, b1


def resize_u8(img, W, H):
    """cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, bit-exact."""
    h, w, _ = img.shape
    xi, a0, a1 = _axis_x(w, W)
    y0, y1, b0, b1 = _axis_y(h, H)
    x1 = np.minimum(xi + 1, w - 1)
    src = img.astype(np.int32)
    hor = src[:, xi, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]
    s0, s1 = hor[y0], hor[y1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def normalize(img_u8, mean, std, to_rgb=True):
    """mmcv.imnormalize on a uint8 HWC image -> float32 HWC:  float32(double(float32(x) - mean32) * (1 / double(std32)))."""
    x = img_u8[..., ::-1] if to_rgb else img_u8
    d = x.astype(np.float32) - np.asarray(mean, np.float32)
    return (d.astype(np.float64) * (1.0 / np.asarray(std, np.float32).astype(np.float64))).astype(np.float32)


def preprocess_frame(img_u8, img_scale=(853, 480), size_divisor=32, mean=(123.675, 116.28, 103.53),
                     std=(58.395, 57.12, 57.375), to_rgb=True):
    """One frame of the test pipeline: uint8 BGR HWC -> float32 CHW, plus (img_shape, scale_factor)."""
    h, w = img_u8.shape[:2]
    rw, rh = rescale_size((w, h), img_scale)
    x = resize_u8(img_u8, rw, rh)
    aw, ah = aligned_size(rw, rh, size_divisor)
    x = resize_u8(x, aw, ah)
    out = normalize(x, mean, std, to_rgb).transpose(2, 0, 1).copy()
    scale_factor = np.array([aw / w, ah / h, aw / w, ah / h], dtype=np.float32)
    return out, (ah, aw, 3), scale_factor
