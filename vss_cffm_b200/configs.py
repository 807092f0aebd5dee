"""``model=`` sections equal to the reference's local_configs/cffm/B{0,1,2,5}/*.py.

/root/reference does not exist on the GPU box, so bench.py / smoke() / the GPU tests build their
models from here; tests/test_config_registry.py checks (in the build container, where the
reference is mounted) that each dict equals what ``Config.fromfile`` reads from the reference's
own file, so the two cannot drift.
"""
IN_CHANNELS = {"b0": [32, 64, 160, 256], "b1": [64, 128, 320, 512], "b2": [64, 128, 320, 512],
               "b5": [64, 128, 320, 512]}
HEAD_DEPTHS = {"b0": 1, "b1": 2, "b2": 2, "b5": 4}
REFERENCE_FILES = {
    (v, "cffm"): f"local_configs/cffm/{v.upper()}/cffm.{v}.480x480.vspw2.160k.py" for v in IN_CHANNELS
}
REFERENCE_FILES.update({
    (v, "cffmpp"): f"local_configs/cffm/{v.upper()}/cffm.{v}.480x480.vspw2_fine_w_proto.40k.py" for v in IN_CHANNELS
})
REFERENCE_FILES.update({
    (v, "proto"): f"local_configs/cffm/{v.upper()}/cffm.{v}.480x480.vspw2_gene_prototype.py" for v in IN_CHANNELS
})
_HEAD_TYPE = {"cffm": "CFFMHead_clips_resize1_8", "cffmpp": "CFFMHead_clips_resize1_8_finetune_w_prototype3",
              "proto": "CFFMHead_clips_resize1_8_gene_prototype"}


def model_cfg(variant="b1", kind="cffm", num_classes=124, num_clips=4, depths=None):
    """The ``model`` dict of the reference config for MiT-``variant`` + CFFM (``kind='cffm'``) or
    CFFM++ (``kind='cffmpp'``); ``pretrained`` is None (no checkpoint is available offline)."""
    return dict(
        type="EncoderDecoder_clips",
        pretrained=None,
        backbone=dict(type=f"mit_{variant}", style="pytorch"),
        decode_head=dict(
            type=_HEAD_TYPE[kind],
            in_channels=list(IN_CHANNELS[variant]),
            in_index=[0, 1, 2, 3],
            feature_strides=[4, 8, 16, 32],
            channels=128,
            dropout_ratio=0.1,
            num_classes=num_classes,
            norm_cfg=dict(type="SyncBN", requires_grad=True),
            align_corners=False,
            decoder_params=dict(embed_dim=256, depths=HEAD_DEPTHS[variant] if depths is None else depths),
            loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0),
            num_clips=num_clips),
        train_cfg=dict(),
        test_cfg=dict(mode="whole"))
