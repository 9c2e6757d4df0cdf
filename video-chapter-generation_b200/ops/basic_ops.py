"""Mirror of the reference's ops/basic_ops.py: the ``Identity`` module the backbones use to drop their classifier
(video_chapter_generation/ops/basic_ops.py, used at model/vision/resnet50_tsm.py:19 and as the vision model of the
precomputed-embedding configuration)."""
import torch


class Identity(torch.nn.Module):
    def forward(self, input):
        return input
