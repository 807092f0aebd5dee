"""``CrossEntropyLoss`` registered so that ``build_loss(loss_decode)`` in the head constructor works
with the reference configs (decode_head.py:574).  Training is out of scope: calling it raises."""
import torch.nn as nn

from .registry import LOSSES


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    def __init__(self, use_sigmoid=False, use_mask=False, reduction="mean", class_weight=None, loss_weight=1.0):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        self.use_sigmoid, self.use_mask = use_sigmoid, use_mask
        self.reduction, self.loss_weight, self.class_weight = reduction, loss_weight, class_weight

    def forward(self, *args, **kwargs):
        raise NotImplementedError("vss_cffm_b200 covers the inference hot path only; losses are out of scope")
