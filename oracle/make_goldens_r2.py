"""Round-2 additions to tests/golden/*, generated like oracle/make_goldens.py by executing the UNMODIFIED reference on CPU
(fp32) -- run in the build container only (needs /root/reference):

    python -m oracle.make_goldens_r2

  basic_layer3d3_60x60.npz     BasicLayer3d3 (depth 2) at the PRODUCTION grid (1,4,256,60,60): 81 windows, every ring / pooled
                               border case; the target frame's output sampled every 3rd pixel (20 x 20 x 256 fp32)
  e2e_b5_T4.npz                EncoderDecoder_clips, MiT-B5 + CFFM head depth 4 (local_configs/cffm/B5), 64x96, T=4
  cffmpp_cluster_layer_60x60.npz  CFFM++ prototype layer (decoder_swin) on 3600 tokens with K = 64 and K = 100 prototypes,
                               every 9th token

Inputs and weights are regenerated from vss_cffm_b200.synth; only reference OUTPUTS are stored.  Kept apart from
make_goldens.py so that re-running that script still reproduces the round-1 files bit for bit.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_goldens import CFG, GOLD, build_ref_segmentor, capture_locals, load_synth, quiet  # noqa: E402
from oracle.ref_shims.install import install  # noqa: E402
from vss_cffm_b200 import synth  # noqa: E402

CFG["b5"] = "local_configs/cffm/B5/cffm.b5.480x480.vspw2.160k.py"


def main():
    install()
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    import mmseg.models  # noqa: F401
    from mmseg.models.decode_heads.cffm_module import cffm_transformer as ct
    C = 256
    # ---- BasicLayer3d3 at 60 x 60
    layer = quiet(ct.BasicLayer3d3, dim=C, depth=2, num_heads=8, window_size=7, mlp_ratio=4., qkv_bias=True,
                  pool_method="fc", focal_level=2, focal_window=5, expand_size=3, focal_l_clips=[1, 2, 3],
                  focal_kernel_clips=[7, 5, 3])
    sd = layer.state_dict()
    layer.load_state_dict({k: (synth.synth_tensor("decoder_focal." + k, v.shape, 4)
                               if (v.is_floating_point() and not synth.is_derived_buffer(k)) else v) for k, v in sd.items()})
    layer.eval()
    x = synth.synth_array((1, 4, C, 60, 60), 23)
    y = layer(x)
    assert torch.equal(y[:, :-1], x[:, :-1])
    np.savez_compressed(os.path.join(GOLD, "basic_layer3d3_60x60.npz"), target_s3=y[0, -1, :, ::3, ::3].numpy())
    # ---- MiT-B5 + depth-4 head end to end
    m, _ = build_ref_segmentor("b5")
    m = load_synth(m, 11)
    B, T, H, W = 1, 4, 64, 96
    imgs = synth.synth_clip(B, T, H, W, seed=11)
    metas = [synth.img_metas(B, H, W)]
    pred, cap = capture_locals(type(m.decode_head).forward.__code__, [], lambda: m(img=[imgs], img_metas=metas, return_loss=False))
    np.savez_compressed(os.path.join(GOLD, "e2e_b5_T4.npz"), logits=cap["__return__"].numpy(), pred=np.stack(pred).astype(np.int16))
    spec_path = os.path.join(GOLD, "state_dict_spec_b5.json")
    import json
    with open(spec_path, "w") as f:
        json.dump({"b5": {k: list(v.shape) for k, v in m.state_dict().items()}}, f, indent=0, sort_keys=True)
    # ---- CFFM++ prototype layer at 60 x 60, K = 64 and K = 100
    mpp, _ = build_ref_segmentor("b1pp")
    head = load_synth(mpp, 9).decode_head
    tok = synth.synth_array((1, 3600, C), 43)
    out = {}
    for K in (64, 100):
        centers = synth.synth_array((1, K, C), 44 + K)
        out[f"out_k{K}_s9"] = head.decoder_swin(tok, 60, 60, centers)[0][:, ::9].numpy()
    np.savez_compressed(os.path.join(GOLD, "cffmpp_cluster_layer_60x60.npz"), **out)
    for fn in ("basic_layer3d3_60x60.npz", "e2e_b5_T4.npz", "cffmpp_cluster_layer_60x60.npz", "state_dict_spec_b5.json"):
        print(f"  {fn}: {os.path.getsize(os.path.join(GOLD, fn)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
