"""Batch-statistics BatchNorm mode of the vision stream (opt-in).

Reference caller #1 nulls ``running_mean`` / ``running_var`` of every ``nn.BatchNorm2d`` after ``.eval()``
(test_video_segment_point.py:116-122).  ``F.batch_norm`` then normalises every BatchNorm of the ResNet-50 with the
statistics of the batch it is handed, i.e. of the ``B*T`` frames of ONE ``model(...)`` call: clips of a DataLoader batch
are coupled, nothing can be folded into the convolutions, and results depend on the batch composition.  The engine's
fused pipeline (eval-mode statistics folded into the weights) cannot express that, so this mode runs the vision stream
layer by layer from the stand-alone C-ABI operators of libvcg_b200.so:

    raw conv (tcgen05 implicit GEMM, no bias / activation)  ->  vcg_op_bn_batch_stats  ->  vcg_op_bn_apply
    (+ residual, ReLU)  ->  vcg_op_tsm_shift in front of every conv1 (ops/temporal_shift.py:34-51)

and hands the ``[B,T,2048]`` embeddings to the engine's text stream + head.  Slower than the fused path (every
activation makes three extra trips through HBM) and meant for reproducing caller #1's numbers only; switch it on with
``TwoStream.bn_batch_stats = True`` or ``VCG_BN_BATCH_STATS=1``.  torch is device memory + the current stream here, as
everywhere in this package: there is no eager arithmetic on the activations.
"""
import contextlib

import torch

from . import binding as _b
from . import ops


def _device_guard(device):
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


def _nhwc_weight(w, dtype):
    """[Cout,Cin,k,k] fp32 -> [Cout,k,k,Cin] in the activation dtype."""
    return w.detach().permute(0, 2, 3, 1).contiguous().to(dtype)


def pack_stem_weight(w, dtype):
    """[64,3,7,7] fp32 -> the stem operand of vcg_op_stem_conv: bf16 [64][4 row pairs][8 px][2 rows][4 ch] (kh = 7,
    kw = 7 and channel 3 are zero padding), fp32 [64][7][8 px][4 ch]."""
    w = w.detach().float()
    if dtype == torch.bfloat16:
        w8 = torch.zeros(64, 8, 8, 4, device=w.device)
        w8[:, :7, :7, :3] = w.permute(0, 2, 3, 1)
        return w8.view(64, 4, 2, 8, 4).permute(0, 1, 3, 2, 4).contiguous().to(dtype)
    w7 = torch.zeros(64, 7, 8, 4, device=w.device)
    w7[:, :, :7, :3] = w.permute(0, 2, 3, 1)
    return w7.contiguous()


class BatchStatVision:
    """ResNet-50 (+ TSM) with batch-statistics BatchNorm over the frames of one call.  Holds re-laid-out copies of the
    convolution weights and the BatchNorm affine parameters of ``state_dict`` (keys below ``prefix``)."""

    BLOCKS = (3, 4, 6, 3)

    def __init__(self, state_dict, clip_frames, shift_div, precision, device, prefix="vision_model.", eps=1e-5):
        if precision not in ("bf16", "fp32"):
            raise RuntimeError(f"unknown precision {precision}")
        self.T, self.shift_div, self.eps = clip_frames, shift_div, eps
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.device = torch.device(device)       # the operators refuse anything but CUDA tensors (no CPU fallback)
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        dev, dt = self.device, self.dtype

        def conv(key):
            return _nhwc_weight(sd[key].to(dev), dt)

        def bn(key):
            return (sd[key + ".weight"].detach().to(dev).float().contiguous(),
                    sd[key + ".bias"].detach().to(dev).float().contiguous())

        self.stem_w = pack_stem_weight(sd["conv1.weight"].to(dev), dt)
        self.stem_bn = bn("bn1")
        self.blocks = []
        for stage, n_blocks in enumerate(self.BLOCKS, start=1):
            for i in range(n_blocks):
                b = f"layer{stage}.{i}."
                tsm = b + "conv1.net.weight" in sd        # TemporalShift wrapper (ops/temporal_shift.py:127-144)
                blk = {
                    "tsm": tsm,
                    "stride": 2 if (i == 0 and stage > 1) else 1,
                    "w1": conv(b + ("conv1.net.weight" if tsm else "conv1.weight")), "bn1": bn(b + "bn1"),
                    "w2": conv(b + "conv2.weight"), "bn2": bn(b + "bn2"),
                    "w3": conv(b + "conv3.weight"), "bn3": bn(b + "bn3"),
                    "wd": None, "bnd": None,
                }
                if b + "downsample.0.weight" in sd:
                    blk["wd"], blk["bnd"] = conv(b + "downsample.0.weight"), bn(b + "downsample.1")
                if tsm and shift_div <= 0:
                    raise RuntimeError("TemporalShift keys in the state dict but shift_div == 0")
                self.blocks.append(blk)

    def _bn(self, x, gb, residual=None, relu=True):
        mean, rstd = ops.bn_batch_stats(x, self.eps)
        return ops.bn_apply(x, mean, rstd, gb[0], gb[1], residual, relu)

    def embed(self, img_clip):
        """img_clip [B,T,3,224,224] fp32 (normalised, CUDA) -> vision_emb [B,T,2048] fp32; BatchNorm statistics are
        those of these B*T frames (the reference's F.batch_norm with running statistics set to None)."""
        B, T = img_clip.shape[0], img_clip.shape[1]
        if T != self.T or tuple(img_clip.shape[2:]) != (3, 224, 224):
            raise RuntimeError(f"vcg_b200: img_clip must be [B,{self.T},3,224,224], got {tuple(img_clip.shape)}")
        if B == 0:
            return torch.empty(0, T, 2048, dtype=torch.float32, device=img_clip.device)
        n = B * T
        none = _b.ACT_NONE
        with _device_guard(img_clip.device):
            xp = ops.nchw_to_stem(img_clip.float().contiguous().view(n, 3, 224, 224), self.dtype)
            x = self._bn(ops.stem_conv_act(xp, self.stem_w, None, none), self.stem_bn)
            del xp
            x, _ = ops.maxpool_tsm(x, T, 0)
            for blk in self.blocks:
                xin = ops.tsm_shift(x, T, x.shape[-1] // self.shift_div) if blk["tsm"] else x
                a = self._bn(ops.conv2d_nhwc(xin, blk["w1"], act=none), blk["bn1"])
                del xin
                a = self._bn(ops.conv2d_nhwc(a, blk["w2"], stride=blk["stride"], act=none), blk["bn2"])
                a = ops.conv2d_nhwc(a, blk["w3"], act=none)
                identity = x
                if blk["wd"] is not None:
                    identity = self._bn(ops.conv2d_nhwc(x, blk["wd"], stride=blk["stride"], act=none), blk["bnd"],
                                        relu=False)
                x = self._bn(a, blk["bn3"], residual=identity, relu=True)
                del a, identity
            return ops.avgpool(x).view(B, T, 2048)
