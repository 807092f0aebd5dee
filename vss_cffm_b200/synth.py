"""Deterministic synthetic weights and inputs (there is no network for checkpoints/datasets).

Values are drawn from ``numpy.random.RandomState`` seeded per state-dict key, so the same
tensors are reproduced bit-for-bit in the build container (golden generation from the real
reference), on the GPU box (parity tests, bench) and inside the CPU oracle.

The zero-/mean-initialised tables of the reference (``relative_position_bias_table`` stays all
zero and the fc-pools are exact means at default init, reference
mmseg/models/decode_heads/cffm_module/cffm_transformer.py:253-254,678-680) are randomised on
purpose so that every bias / pooling path is exercised.
"""
import zlib

import numpy as np
import torch

_INT_BUFFERS = ("relative_position_index", "valid_ind_rolled", "num_batches_tracked", "attn_mask",
                "valid_ind_unfold")


def is_derived_buffer(key):
    return any(s in key for s in _INT_BUFFERS)


def _rs(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_tensor(key, shape, seed=0):
    """fp32 tensor for state-dict entry ``key`` of ``shape``."""
    shape = tuple(shape)
    r = _rs(key, seed)
    n = r.standard_normal(shape).astype(np.float32) if len(shape) else np.float32(r.standard_normal())
    leaf = key.rsplit(".", 1)[-1]
    if "relative_position_bias_table" in key:
        v = 0.3 * n
    elif "pool_layers" in key:
        if leaf == "weight":
            v = (1.0 / shape[-1]) * (1.0 + 0.3 * n)
        else:
            v = 0.05 * n
    elif leaf == "running_mean":
        v = 0.1 * n
    elif leaf == "running_var":
        v = r.uniform(0.5, 1.5, size=shape).astype(np.float32)
    elif leaf == "bias":
        v = 0.05 * n
    elif leaf == "weight" and len(shape) == 1:            # LayerNorm / BatchNorm scale
        v = 1.0 + 0.1 * n
    elif leaf == "weight":
        fan_in = int(np.prod(shape[1:]))
        v = n / np.sqrt(fan_in)
    else:
        v = 0.1 * n
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def synth_state_dict(spec, seed=0):
    """spec: mapping key -> shape (derived integer buffers are skipped)."""
    return {k: synth_tensor(k, s, seed) for k, s in spec.items() if not is_derived_buffer(k)}


def fill_module(module, seed=0):
    """Overwrite every float parameter/buffer of ``module`` in place, keyed by state-dict name."""
    with torch.no_grad():
        for k, t in module.state_dict().items():
            if is_derived_buffer(k) or not t.is_floating_point():
                continue
            t.copy_(synth_tensor(k, t.shape, seed).to(t.device, t.dtype))
    return module


def synth_clip(batch, frames, height, width, seed=0):
    """list of ``frames`` tensors (B,3,H,W) fp32 ~ N(0,1): the post-Normalize_clips statistics
    (reference local_configs/_base_/datasets/vspw_repeat2.py:4-5)."""
    r = np.random.RandomState(1000 + seed)
    return [torch.from_numpy(r.standard_normal((batch, 3, height, width)).astype(np.float32)) for _ in range(frames)]


def synth_array(shape, seed, scale=1.0):
    r = np.random.RandomState(seed)
    return torch.from_numpy((scale * r.standard_normal(tuple(shape))).astype(np.float32))


def img_metas(batch, height, width, video="vid0"):
    return [dict(ori_shape=(height, width, 3), img_shape=(height, width, 3), pad_shape=(height, width, 3),
                 flip=False, filename=f"data/{video}/origin/{i:08d}.jpg") for i in range(batch)]
