"""Mirror of the reference's model/vision/resnet50.py (plain ResNet-50, no temporal shift; importable because the
named callers import it, video_chapter_generation/test_video_segment_point.py:25)."""
import torch
import torch.nn as nn

from ops.basic_ops import Identity

from ._resnet_params import ResNet50Params, _unimodal_vision_forward


class Resnet50(torch.nn.Module):
    def __init__(self, segments_size, pretrain_stage=True):
        super().__init__()
        self.pretrain_stage = pretrain_stage
        self.segments_size = segments_size
        self.base_model = ResNet50Params()
        self.feature_dim = self.base_model.fc.in_features
        self.base_model.fc = Identity()
        self.head = None

    def build_chapter_head(self):
        self.head = nn.Linear(self.segments_size * self.feature_dim, 2)

    def forward(self, x):
        """x [B,T,3,224,224] -> (logits [B,2], prob [B,2]) (reference :64-73), plain ResNet-50 (no temporal shift)."""
        return _unimodal_vision_forward(self, x, 0)
