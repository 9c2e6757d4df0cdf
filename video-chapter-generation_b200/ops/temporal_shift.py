"""Mirror of the reference's ops/temporal_shift.py API (TSM, arXiv:1811.08383).

``TemporalShift`` wraps a conv (as ``.net``, which is what puts ``conv1.net.weight`` into the checkpoint keys,
reference ops/temporal_shift.py:138) and records ``n_segment`` / ``fold_div``.  In this implementation the shift is
never materialised: the engine folds it into the A-operand load of the wrapped 1x1 conv (csrc/conv_gemm.cuh) and
into the epilogue of the previous block.  ``TemporalShift.shift`` is kept as a static helper with the reference's
semantics (reference :34-51) for CPU-side checks; ``forward`` of a wrapped conv is not a supported entry point —
the whole backbone runs through ``TwoStream.forward``.
"""
import torch
import torch.nn as nn


class TemporalShift(nn.Module):
    def __init__(self, net, n_segment=3, n_div=8, inplace=False):
        super().__init__()
        self.net = net
        self.n_segment = n_segment
        self.fold_div = n_div
        self.inplace = inplace

    def forward(self, x):
        raise RuntimeError("TemporalShift.forward: the shift is fused into the sm_100a conv kernels; "
                           "run the model through TwoStream.forward (there is no eager fallback)")

    @staticmethod
    def shift(x, n_segment, fold_div=3, inplace=False):
        nt, c, h, w = x.size()
        x = x.view(nt // n_segment, n_segment, c, h, w)
        fold = c // fold_div
        out = torch.zeros_like(x)
        out[:, :-1, :fold] = x[:, 1:, :fold]                    # channels [0,fold): from frame t+1
        out[:, 1:, fold:2 * fold] = x[:, :-1, fold:2 * fold]    # channels [fold,2fold): from frame t-1
        out[:, :, 2 * fold:] = x[:, :, 2 * fold:]
        return out.view(nt, c, h, w)


def make_temporal_shift(net, n_segment, n_div=8, place='blockres', temporal_pool=False):
    """Wrap conv1 of every bottleneck (place='blockres', the only placement the reference callers use)."""
    if temporal_pool:
        raise NotImplementedError("temporal_pool is unused by the reference's scoring path")
    if 'blockres' not in place:
        raise NotImplementedError(place)
    for name in ("layer1", "layer2", "layer3", "layer4"):
        for block in getattr(net, name).children():
            block.conv1 = TemporalShift(block.conv1, n_segment=n_segment, n_div=n_div)
