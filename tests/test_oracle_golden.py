"""Pins the CPU oracle (oracle/cffm_oracle.py) against golden vectors produced by the
UNMODIFIED reference (oracle/make_goldens.py).  fp32 CPU on both sides: tolerances only
cover summation-order differences (gather formulation vs roll/unfold/cat)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cffm_oracle as O
from vss_cffm_b200 import synth

torch.set_grad_enabled(False)


def _spec(golden_dir, tag):
    with open(os.path.join(golden_dir, "state_dict_spec.json")) as f:
        return json.load(f)[tag]


def _close(a, b, rtol=2e-5, atol=2e-5):
    a = torch.as_tensor(a); b = torch.as_tensor(b)
    assert a.shape == b.shape
    scale = b.abs().max().item()
    err = (a - b).abs().max().item()
    assert err <= atol + rtol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


def test_relative_position_index_buffers(golden_dir):
    t = np.load(os.path.join(golden_dir, "index_tables.npz"))
    assert np.array_equal(O.relative_position_index((7, 7), (7, 7)).numpy(), t["relative_position_index"])
    assert np.array_equal(O.relative_position_index((7, 7), (5, 5)).numpy(), t["relative_position_index_0"])
    for k, kc in enumerate(O.CFFM_K_CLIPS):
        assert np.array_equal(O.relative_position_index((7, 7), (kc, kc)).numpy(), t[f"relative_position_index_clips_{k}"])
    assert t["valid_ind_rolled"].shape == (132,) and len(O.ring_offsets()) == 132


@pytest.mark.parametrize("Hp,Wp", [(21, 28), (63, 63)])
def test_key_source_table_bit_exact(golden_dir, Hp, Wp):
    """Integer K/V source coordinates == what the reference's roll/partition/unfold/cat builds."""
    t = np.load(os.path.join(golden_dir, "index_tables.npz"))
    lev, ys, xs = O.key_source_table(Hp, Wp)
    code = torch.where(ys < 0, torch.zeros_like(ys), lev * 10000 + ys * 100 + xs + 1).to(torch.int32)
    assert code.shape == (Hp // 7 * (Wp // 7), 289)
    assert np.array_equal(code.numpy(), t[f"key_code_{Hp}x{Wp}"])
    mask = torch.where(ys[:, 181:] < 0, -100.0, 0.0)
    assert np.array_equal(mask.numpy(), t[f"mask_{Hp}x{Wp}"])
    # 120 unique ring positions + 12 duplicates
    ring = {(int(a), int(b)) for a, b in zip(ys[0, 49:181], xs[0, 49:181])}
    assert len(ring) == 120


def test_cfm_attention(golden_dir):
    g = np.load(os.path.join(golden_dir, "cfm_attention.npz"))
    C, Hp, Wp, B = 256, 21, 28, 1
    nWh, nWw = Hp // 7, Wp // 7
    spec = {k[len("decode_head.decoder_focal.blocks.0.attn."):]: v for k, v in _spec(golden_dir, "b1").items()
            if k.startswith("decode_head.decoder_focal.blocks.0.attn.")}
    sd = {"attn." + k: synth.synth_tensor("attn." + k, s, 3) for k, s in spec.items() if not synth.is_derived_buffer(k)}
    xt = synth.synth_array((B, Hp, Wp, C), 11)
    pooled = [synth.synth_array((B, nWh, nWw, C), 12), synth.synth_array((B, nWh, nWw, C), 13),
              synth.synth_array((B, 2 * nWh, 2 * nWw, C), 14), synth.synth_array((B, 3 * nWh, 3 * nWw, C), 15)]
    out, probs = O.cfm_attention(sd, "attn", xt, pooled, return_probs=True)
    _close(out, g["out"])
    _close(probs[0, 5], g["probs_w5"], atol=1e-6)


def test_basic_layer3d3(golden_dir):
    g = np.load(os.path.join(golden_dir, "basic_layer3d3.npz"))
    spec = {k[len("decode_head."):]: v for k, v in _spec(golden_dir, "b1").items()
            if k.startswith("decode_head.decoder_focal.")}
    sd = synth.synth_state_dict(spec, 4)
    x = synth.synth_array((1, 4, 256, 20, 25), 21)
    y = O.basic_layer3d3(sd, "decoder_focal", x, 2)
    assert torch.equal(y[:, :-1], x[:, :-1])           # reference frames pass through unchanged
    _close(y[0, -1], g["target"], rtol=5e-5)


@pytest.mark.parametrize("tag,seed", [("b0", 5), ("b1", 6)])
def test_mit_backbone(golden_dir, tag, seed):
    g = np.load(os.path.join(golden_dir, f"mit_{tag}.npz"))
    spec = {k[len("backbone."):]: v for k, v in _spec(golden_dir, tag).items() if k.startswith("backbone.")}
    sd = synth.synth_state_dict(spec, seed)
    outs = O.mit_forward(sd, "", synth.synth_array((2, 3, 64, 96), 31), "mit_" + tag)
    for i, o in enumerate(outs):
        _close(o, g[f"out{i}"], rtol=5e-5)


@pytest.mark.parametrize("tag,T,B,seed,depth", [("b0", 2, 1, 7, 1), ("b0", 4, 1, 7, 1), ("b1", 4, 2, 8, 2), ("b2", 4, 1, 10, 2)])
def test_end_to_end_segmentor(golden_dir, tag, T, B, seed, depth):
    g = np.load(os.path.join(golden_dir, f"e2e_{tag}_T{T}.npz"))
    sd = synth.synth_state_dict(_spec(golden_dir, tag), seed)
    imgs = synth.synth_clip(B, T, 64, 96, seed=seed)
    pred, logits = O.segmentor_simple_test(sd, imgs, "mit_" + tag, depth, return_logits=True)
    _close(logits, g["logits"], rtol=1e-4)
    agree = (pred.numpy() == g["pred"]).mean()
    assert agree >= 0.9995, agree


def test_cffmpp_cluster_layer(golden_dir):
    g = np.load(os.path.join(golden_dir, "cffmpp_cluster_layer.npz"))
    spec = {k[len("decode_head."):]: v for k, v in _spec(golden_dir, "b1pp").items()
            if k.startswith("decode_head.decoder_swin.")}
    sd = synth.synth_state_dict({"decode_head." + k: v for k, v in spec.items()}, 9)
    tok = synth.synth_array((2, 8 * 12, 256), 41)
    centers = synth.synth_array((2, 10, 256), 42)
    out = O.cluster_layer(sd, "decode_head.decoder_swin", tok, centers, 1)
    _close(out, g["out"], rtol=5e-5)


# ------------------------------------------------------------------ round 2: production grid, MiT-B5 + depth 4, CFFM++ at 60 x 60
def test_basic_layer3d3_production_grid(golden_dir):
    """BasicLayer3d3 on (1,4,256,60,60): 81 windows, every ring / pooled border case, vs the unmodified reference."""
    g = np.load(os.path.join(golden_dir, "basic_layer3d3_60x60.npz"))
    spec = {k[len("decode_head."):]: v for k, v in _spec(golden_dir, "b1").items()
            if k.startswith("decode_head.decoder_focal.")}
    sd = synth.synth_state_dict(spec, 4)
    x = synth.synth_array((1, 4, 256, 60, 60), 23)
    y = O.basic_layer3d3(sd, "decoder_focal", x, 2)
    assert torch.equal(y[:, :-1], x[:, :-1])
    _close(y[0, -1, :, ::3, ::3], g["target_s3"], rtol=5e-5)


def test_end_to_end_mit_b5_depth4(golden_dir):
    """MiT-B5 ([3,6,40,3] blocks) + CFFM head of depth 4 (local_configs/cffm/B5) vs the unmodified reference."""
    g = np.load(os.path.join(golden_dir, "e2e_b5_T4.npz"))
    with open(os.path.join(golden_dir, "state_dict_spec_b5.json")) as f:
        spec = json.load(f)["b5"]
    assert sum(1 for k in spec if k.startswith("decode_head.decoder_focal.blocks.") and k.endswith("norm1.weight")) == 4
    sd = synth.synth_state_dict(spec, 11)
    imgs = synth.synth_clip(1, 4, 64, 96, seed=11)
    pred, logits = O.segmentor_simple_test(sd, imgs, "mit_b5", 4, return_logits=True)
    _close(logits, g["logits"], rtol=2e-4)
    assert (pred.numpy() == g["pred"]).mean() >= 0.9995


def test_cffmpp_cluster_layer_production_grid(golden_dir):
    g = np.load(os.path.join(golden_dir, "cffmpp_cluster_layer_60x60.npz"))
    spec = {k[len("decode_head."):]: v for k, v in _spec(golden_dir, "b1pp").items()
            if k.startswith("decode_head.decoder_swin.")}
    sd = synth.synth_state_dict({"decode_head." + k: v for k, v in spec.items()}, 9)
    tok = synth.synth_array((1, 3600, 256), 43)
    for K in (64, 100):
        centers = synth.synth_array((1, K, 256), 44 + K)
        out = O.cluster_layer(sd, "decode_head.decoder_swin", tok, centers, 1)
        _close(out[:, ::9], g[f"out_k{K}_s9"], rtol=5e-5)
