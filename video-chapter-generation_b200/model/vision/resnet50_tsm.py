"""Mirror of the reference's model/vision/resnet50_tsm.py: ResNet-50 + Temporal Shift Module vision stream.

Same constructor, attributes (``base_model``, ``feature_dim``, ``segments_size``, ``head``) and checkpoint keys as
video_chapter_generation/model/vision/resnet50_tsm.py:10-24.  The reference downloads ImageNet weights in the
constructor (:15); there is no network here, so parameters start from torchvision's default initialisation and are
expected to come from ``load_state_dict``.
"""
import torch
import torch.nn as nn

from ops.basic_ops import Identity
from ops.temporal_shift import make_temporal_shift

from ._resnet_params import ResNet50Params, _unimodal_vision_forward


class Resnet50TSM(torch.nn.Module):
    def __init__(self, segments_size=8, shift_div=8, pretrain_stage=True):
        super().__init__()
        self.pretrain_stage = pretrain_stage
        self.base_model = ResNet50Params()
        make_temporal_shift(self.base_model, n_segment=segments_size, n_div=shift_div)
        self.segments_size = segments_size
        self.shift_div = shift_div
        self.feature_dim = self.base_model.fc.in_features
        self.base_model.fc = Identity()   # discard the classifier
        self.head = None

    def build_chapter_head(self):
        self.head = nn.Linear(self.segments_size * self.feature_dim, 2)

    def forward(self, x):
        """x [B,T,3,224,224] -> (logits [B,2], prob [B,2]) (reference :68-77).  Runs inside libvcg_b200.so."""
        return _unimodal_vision_forward(self, x, self.shift_div)
