"""Generate tests/golden/* by executing the UNMODIFIED reference on CPU (fp32).

Run in the build container only (needs /root/reference):

    python -m oracle.make_goldens

The reference's own tests hold no vector for this path, so the
oracle is pinned against outputs of the reference itself.  Inputs and weights are NOT
stored: they are regenerated from ``vss_cffm_b200.synth`` (numpy RandomState, stable
across machines); only reference OUTPUTS, integer tables and state-dict specs are stored.
Absent third-party packages (mmcv, timm, IPython, fast_pytorch_kmeans) are replaced by
the plumbing-only stand-ins in ``oracle/ref_shims`` -- all arithmetic is executed by the
reference's files.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shims.install import install, REFERENCE_ROOT  # noqa: E402
from vss_cffm_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG = {
    "b0": "local_configs/cffm/B0/cffm.b0.480x480.vspw2.160k.py",
    "b1": "local_configs/cffm/B1/cffm.b1.480x480.vspw2.160k.py",
    "b2": "local_configs/cffm/B2/cffm.b2.480x480.vspw2.160k.py",
    "b1pp": "local_configs/cffm/B1/cffm.b1.480x480.vspw2_fine_w_proto.40k.py",
}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def load_synth(module, seed):
    sd = module.state_dict()
    new = {k: (synth.synth_tensor(k, v.shape, seed) if (v.is_floating_point() and not synth.is_derived_buffer(k)) else v)
           for k, v in sd.items()}
    module.load_state_dict(new, strict=True)
    return module.eval()


def capture_locals(code, names, fn):
    store = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code is code:
            for n in names:
                if n in frame.f_locals:
                    store[n] = frame.f_locals[n]
            store["__return__"] = arg
    sys.setprofile(prof)
    try:
        out = fn()
    finally:
        sys.setprofile(None)
    return out, store


def build_ref_segmentor(tag):
    import mmcv
    from mmseg.models import build_segmentor
    cfg = mmcv.Config.fromfile(os.path.join(REFERENCE_ROOT, CFG[tag]))
    cfg.model["pretrained"] = None
    return quiet(build_segmentor, cfg.model), cfg


def main():
    install()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    import mmseg.models  # noqa: F401  (registers everything)
    from mmseg.models.decode_heads.cffm_module import cffm_transformer as ct
    from mmseg.models.decode_heads.pvt import swin_transformer_2d as st

    # ---------------------------------------------------------------- 1. state-dict contract
    spec = {}
    models = {}
    for tag in CFG:
        m, cfg = build_ref_segmentor(tag)
        models[tag] = m
        spec[tag] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f, indent=0, sort_keys=True)

    # ---------------------------------------------------------------- 2. integer tables
    attn = quiet(ct.WindowAttention3d3, 8, expand_size=3, window_size=(7, 7), focal_window=5, focal_level=2,
                 num_heads=8, pool_method="fc", focal_l_clips=[1, 2, 3], focal_kernel_clips=[7, 5, 3])
    tables = {k: v.numpy() for k, v in attn.state_dict().items() if "index" in k or "valid_ind" in k}
    # K projection := identity on channel 0 so that k_all carries the source code of every key slot
    attn.qkv.weight.zero_(); attn.qkv.bias.zero_()
    for c in range(8):
        attn.qkv.weight[8 + c, 0] = 1.0          # every head's (d=1) K channel = input channel 0
    for Hp, Wp in ((21, 28), (63, 63)):
        nWh, nWw = Hp // 7, Wp // 7

        def code(level, h, w):
            y = torch.arange(h).view(h, 1).expand(h, w)
            x = torch.arange(w).view(1, w).expand(h, w)
            t = torch.zeros(1, h, w, 8)
            t[0, :, :, 0] = (level * 10000 + y * 100 + x + 1).float()
            return t
        x_all = [[code(0, Hp, Wp), code(1, nWh, nWw)], code(2, nWh, nWw), code(3, 2 * nWh, 2 * nWw),
                 code(4, 3 * nWh, 3 * nWw)]
        masks = [[None, None], None, None, None]
        _, loc = capture_locals(ct.WindowAttention3d3.forward.__code__, ["k_all"],
                                lambda: attn(x_all, mask_all=masks, batch_size=1, num_clips=4))
        k_all = loc["k_all"]                                    # (nW, 8, 289, 1)
        assert k_all.shape == (nWh * nWw, 8, 289, 1)
        assert (k_all[:, 0] == k_all[:, 7]).all()
        tables[f"key_code_{Hp}x{Wp}"] = k_all[:, 0, :, 0].round().to(torch.int32).numpy()
        # the -100 masks the reference derives for the pooled segments
        tables[f"mask_{Hp}x{Wp}"] = torch.cat([masks[0][1].reshape(nWh * nWw, -1)] +
                                              [masks[i].reshape(nWh * nWw, -1) for i in (1, 2, 3)], 1).numpy()
    np.savez_compressed(os.path.join(GOLD, "index_tables.npz"), **tables)

    # ---------------------------------------------------------------- 3. CFM attention (dim 256)
    C = 256
    attn = quiet(ct.WindowAttention3d3, C, expand_size=3, window_size=(7, 7), focal_window=5, focal_level=2,
                 num_heads=8, pool_method="fc", focal_l_clips=[1, 2, 3], focal_kernel_clips=[7, 5, 3])
    sd = attn.state_dict()
    attn.load_state_dict({k: (synth.synth_tensor("attn." + k, v.shape, 3) if v.is_floating_point() else v)
                          for k, v in sd.items()})
    Hp, Wp, B = 21, 28, 1
    nWh, nWw = Hp // 7, Wp // 7
    x_all = [[synth.synth_array((B, Hp, Wp, C), 11), synth.synth_array((B, nWh, nWw, C), 12)],
             synth.synth_array((B, nWh, nWw, C), 13), synth.synth_array((B, 2 * nWh, 2 * nWw, C), 14),
             synth.synth_array((B, 3 * nWh, 3 * nWw, C), 15)]
    out, loc = capture_locals(ct.WindowAttention3d3.forward.__code__, ["attn"],
                              lambda: attn(x_all, mask_all=[[None, None], None, None, None], batch_size=B, num_clips=4))
    np.savez_compressed(os.path.join(GOLD, "cfm_attention.npz"), out=out.numpy(),
                        probs_w5=loc["attn"][5].numpy())       # post-softmax probs of window 5 (all heads)

    # ---------------------------------------------------------------- 4. BasicLayer3d3 (2 blocks)
    layer = quiet(ct.BasicLayer3d3, dim=C, depth=2, num_heads=8, window_size=7, mlp_ratio=4., qkv_bias=True,
                  pool_method="fc", focal_level=2, focal_window=5, expand_size=3, focal_l_clips=[1, 2, 3],
                  focal_kernel_clips=[7, 5, 3])
    sd = layer.state_dict()
    layer.load_state_dict({k: (synth.synth_tensor("decoder_focal." + k, v.shape, 4)
                               if (v.is_floating_point() and not synth.is_derived_buffer(k)) else v)
                           for k, v in sd.items()})
    layer.eval()
    x = synth.synth_array((1, 4, C, 20, 25), 21)
    y = layer(x)
    assert torch.equal(y[:, :-1], x[:, :-1])
    np.savez_compressed(os.path.join(GOLD, "basic_layer3d3.npz"), target=y[0, -1].numpy())

    # ---------------------------------------------------------------- 5. MiT backbones
    for tag, seed in (("b0", 5), ("b1", 6)):
        bb = load_synth(models[tag].backbone, seed)
        img = synth.synth_array((2, 3, 64, 96), 31)
        outs = bb(img)
        np.savez_compressed(os.path.join(GOLD, f"mit_{tag}.npz"), **{f"out{i}": o.numpy() for i, o in enumerate(outs)})

    # ---------------------------------------------------------------- 6. end-to-end segmentor
    for tag, T, B, seed in (("b0", 2, 1, 7), ("b0", 4, 1, 7), ("b1", 4, 2, 8), ("b2", 4, 1, 10)):
        m = load_synth(models[tag], seed)
        H, W = 64, 96
        imgs = synth.synth_clip(B, T, H, W, seed=seed)
        metas = [synth.img_metas(B, H, W)]
        pred, cap = capture_locals(type(m.decode_head).forward.__code__, [],
                                   lambda: m(img=[imgs], img_metas=metas, return_loss=False))
        np.savez_compressed(os.path.join(GOLD, f"e2e_{tag}_T{T}.npz"), logits=cap["__return__"].numpy(),
                            pred=np.stack(pred).astype(np.int16))

    # ---------------------------------------------------------------- 7. CFFM++ prototype branch
    m = load_synth(models["b1pp"], 9)
    head = m.decode_head
    B, h2, w2 = 2, 8, 12
    tok = synth.synth_array((B, h2 * w2, C), 41)
    centers = synth.synth_array((B, 10, C), 42)
    c3 = head.decoder_swin(tok, h2, w2, centers)[0]
    np.savez_compressed(os.path.join(GOLD, "cffmpp_cluster_layer.npz"), out=c3.numpy())
    print("goldens written to", GOLD)
    for fn in sorted(os.listdir(GOLD)):
        print(f"  {fn}: {os.path.getsize(os.path.join(GOLD, fn)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
