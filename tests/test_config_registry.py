"""Host logic of the drop-in boundary: registries, config loader, state-dict contract, index tables."""
import json
import os

import numpy as np
import pytest
import torch

import vss_cffm_b200 as V
from oracle import cffm_oracle as O
from vss_cffm_b200 import cffm_tables as tb
from vss_cffm_b200 import configs, synth
from vss_cffm_b200.config import Config
from vss_cffm_b200.registry import Registry, build_from_cfg

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")


def test_registered_names_match_the_reference():
    for n in ("mit_b0", "mit_b1", "mit_b2", "mit_b3", "mit_b4", "mit_b5"):
        assert n in V.BACKBONES
    assert "CFFMHead_clips_resize1_8" in V.HEADS
    assert "CFFMHead_clips_resize1_8_finetune_w_prototype3" in V.HEADS
    assert "EncoderDecoder_clips" in V.SEGMENTORS
    assert "CrossEntropyLoss" in V.LOSSES


def test_registry_semantics():
    """Fake-plugin pattern of the reference's tests/test_models/test_segmentor.py:47-78."""
    reg = Registry("thing")

    @reg.register_module()
    class A:
        def __init__(self, x, y=2):
            self.x, self.y = x, y
    with pytest.raises(KeyError):
        reg.register_module()(A)
    reg.register_module(force=True)(A)
    obj = build_from_cfg(dict(type="A", x=1), reg, default_args=dict(y=5))
    assert (obj.x, obj.y) == (1, 5)
    with pytest.raises(KeyError):
        build_from_cfg(dict(type="B"), reg)
    with pytest.raises(KeyError):
        build_from_cfg(dict(x=1), reg)
    with pytest.raises(TypeError):
        build_from_cfg([1], reg)
    assert build_from_cfg(dict(type=A, x=3), reg).x == 3


def test_config_base_merge_and_delete(tmp_path):
    (tmp_path / "base.py").write_text("optimizer = dict(type='SGD', lr=0.1, momentum=0.9)\nmodel = dict(a=1, b=dict(c=2, d=3))\n")
    (tmp_path / "child.py").write_text("_base_ = ['./base.py']\noptimizer = dict(_delete_=True, type='AdamW', lr=1e-4)\n"
                                       "model = dict(b=dict(c=5))\n")
    cfg = Config.fromfile(str(tmp_path / "child.py"))
    assert cfg.optimizer == dict(type="AdamW", lr=1e-4)
    assert cfg.model == dict(a=1, b=dict(c=5, d=3))
    assert cfg.model.b.c == 5
    with pytest.raises(FileNotFoundError):
        Config.fromfile(str(tmp_path / "nope.py"))


@needs_ref
@pytest.mark.parametrize("key", sorted(configs.REFERENCE_FILES))
def test_reference_configs_load_unchanged_and_build(key):
    """local_configs/cffm/** load with this package's loader, equal the built-in dicts, and build."""
    variant, kind = key
    cfg = Config.fromfile(os.path.join(REF, configs.REFERENCE_FILES[key]))
    model = dict(cfg.model)
    assert isinstance(model.pop("pretrained"), str)
    builtin = configs.model_cfg(variant, kind)
    builtin.pop("pretrained")
    assert model == builtin
    assert cfg.optimizer["type"] == "AdamW" and "momentum" not in cfg.optimizer          # _delete_ honoured
    if variant in ("b0", "b1"):
        cfg.model["pretrained"] = None
        m = V.build_segmentor(cfg.model)
        assert type(m).__name__ == "EncoderDecoder_clips" and m.num_classes == 124 and not m.training


@pytest.mark.parametrize("tag,kind,spec_tag", [("b0", "cffm", "b0"), ("b1", "cffm", "b1"), ("b2", "cffm", "b2"),
                                               ("b1", "cffmpp", "b1pp")])
def test_state_dict_contract(golden_dir, tag, kind, spec_tag):
    """Same keys and shapes as the reference's modules (checkpoints load unchanged)."""
    with open(os.path.join(golden_dir, "state_dict_spec.json")) as f:
        ref = json.load(f)[spec_tag]
    m = V.build_segmentor(V.model_cfg(tag, kind))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == ref
    sd = synth.synth_state_dict(ref, 1)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(synth.is_derived_buffer(k) for k in missing)


def test_index_tables_bit_exact(golden_dir):
    t = np.load(os.path.join(golden_dir, "index_tables.npz"))
    assert np.array_equal(tb.relative_position_index((7, 7), (7, 7)).numpy(), t["relative_position_index"])
    assert np.array_equal(tb.relative_position_index((7, 7), (5, 5)).numpy(), t["relative_position_index_0"])
    for k, kc in enumerate(tb.K_CLIPS):
        assert np.array_equal(tb.relative_position_index((7, 7), (kc, kc)).numpy(), t[f"relative_position_index_clips_{k}"])
    assert np.array_equal(tb.valid_ind_rolled().numpy(), t["valid_ind_rolled"])
    head = V.build_head(V.model_cfg("b1")["decode_head"])
    a = head.decoder_focal.blocks[0].attn
    for name in ("relative_position_index", "relative_position_index_0", "relative_position_index_clips_0",
                 "relative_position_index_clips_1", "relative_position_index_clips_2", "valid_ind_rolled"):
        assert np.array_equal(getattr(a, name).numpy(), t[name]) and getattr(a, name).dtype == torch.int64


def test_bias_table_assembly_matches_oracle():
    spec = {"a.relative_position_bias_table": (169, 8), "a.relative_position_bias_table_to_neighbors": (1, 8, 49, 132),
            "a.relative_position_bias_table_to_windows.0": (8, 121), "a.relative_position_bias_table_to_windows_clips.0": (8, 169),
            "a.relative_position_bias_table_to_windows_clips.1": (8, 121), "a.relative_position_bias_table_to_windows_clips.2": (8, 81)}
    sd = synth.synth_state_dict(spec, 2)
    from vss_cffm_b200 import ops
    lay = ops.cfm_layout()                                     # host-only query of the kernel's key-row layout
    scale = 32 ** -0.5
    got = tb.assemble_bias_tc(sd["a.relative_position_bias_table"], sd["a.relative_position_bias_table_to_neighbors"],
                              sd["a.relative_position_bias_table_to_windows.0"],
                              [sd[f"a.relative_position_bias_table_to_windows_clips.{k}"] for k in range(3)], scale, lay)
    assert got.shape == (8, 49, lay["pitch"]) and got.dtype == torch.float16
    ref = O.cfm_bias_table(sd, "a").double()                   # (8, 49, 289) in the reference's key order
    # fold the reference columns that name the same key (the 12 ring keys listed twice): exp-sum, exactly what the
    # softmax does with two equal logits + different biases
    rows = [(n // 7 + 3) * 13 + n % 7 + 3 for n in range(49)] + [(dy + 3) * 13 + dx + 3 for dy, dx in O.ring_offsets()]
    for blk, cnt in zip(lay["rows"][1:], (25, 49, 25, 9)):
        rows += [blk + m for m in range(cnt)]
    want = torch.full((8, 49, lay["npad"]), float("-inf"), dtype=torch.float64)
    for n, r in enumerate(rows):
        want[:, :, r] = torch.logaddexp(want[:, :, r], ref[:, :, n])
    assert len(set(rows)) == 277
    g = got[:, :, :lay["npad"]].double() * scale
    used = torch.isfinite(want)
    assert torch.equal(torch.isfinite(g), used)                # unused key rows carry -inf
    assert (g[used] - want[used]).abs().max() <= 1e-3 * want[used].abs().max()      # fp16 table


def test_modules_refuse_training_and_cpu_execution():
    m = V.build_segmentor(V.model_cfg("b0"))
    with pytest.raises(V._abi.CffmError):
        m.train()
    assert m.eval() is m
    if not torch.cuda.is_available():
        with pytest.raises(V._abi.CffmError):       # no CPU fallback
            m(img=[synth.synth_clip(1, 4, 64, 64)], img_metas=[synth.img_metas(1, 64, 64)], return_loss=False)
    with pytest.raises(NotImplementedError):
        m(img=None, img_metas=None, return_loss=True)
    with pytest.raises(TypeError):
        m.forward_test(torch.zeros(1), [[]])


def test_cffmpp_load_centers_both_branches(tmp_path):
    """``_load_centers`` (reference cffm_head.py:429-455): <save_path>/<video>/centers.pt when it exists; otherwise every
    *.pt of the video concatenated along the prototype axis and a random 80 % of them kept, in their original order."""
    head = V.build_head(V.model_cfg("b1", "cffmpp")["decode_head"])
    head.save_path = str(tmp_path) + "/"
    (tmp_path / "vidA").mkdir(); (tmp_path / "vidB").mkdir()
    ca = torch.arange(1 * 64 * 256, dtype=torch.float32).view(1, 64, 256)
    torch.save(ca, tmp_path / "vidA" / "centers.pt")
    parts = [torch.full((1, 10, 256), float(i)) + torch.arange(10).view(1, 10, 1) * 0.01 for i in range(8)]   # 80 prototypes
    for i, p in enumerate(parts):
        torch.save(p, tmp_path / "vidB" / f"part{i:02d}.pt")
    metas = [dict(filename="data/vidA/origin/00000001.jpg")]
    got = head._load_centers(metas, 1, "cpu")
    assert torch.equal(got, ca)
    metas = [dict(filename="data/vidB/origin/00000007.jpg")]
    torch.manual_seed(0)
    got = head._load_centers(metas, 1, "cpu")
    allp = torch.cat(parts, dim=1)[0]
    assert got.shape == (1, 64, 256)                           # int(80 * 0.8) kept
    # every kept row is one of the 80 prototypes, each at most once, original order preserved
    idx = [int((allp == r).all(dim=1).nonzero()[0]) for r in got[0]]
    assert idx == sorted(set(idx)) and len(idx) == 64
    with pytest.raises(FileNotFoundError):
        head._load_centers([dict(filename="data/vidC/origin/00000001.jpg")], 1, "cpu")
    with pytest.raises(AssertionError):
        head._load_centers(metas, 2, "cpu")                    # batch_size must equal len(img_metas) (:430)
