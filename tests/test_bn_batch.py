"""CPU tests of the batch-statistics BatchNorm mode (reference caller #1, test_video_segment_point.py:116-122).

1. The oracle restatement (two_stream_oracle.batch_stat_bn) against the goldens oracle/make_golden_bn_batch.py wrote from
   the unmodified reference.
2. The ORCHESTRATION of vcg_b200.bn_batch.BatchStatVision (weight re-layout, stem packing, shift placement, residual and
   downsample order, which BatchNorm gets a ReLU) against the same golden, with the C-ABI operators replaced by torch
   stand-ins defined HERE (test doubles: the product has no CPU path, vcg_b200.ops refuses CPU tensors).
"""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def golden_case(name):
    from oracle import two_stream_oracle as orc
    from oracle import weights as W
    g = np.load(os.path.join(GOLDEN, f"bn_batch_{name}.npz"))
    T, L, B, seed = [int(x) for x in g["meta"][:4]]
    ids, mask = W.make_text(B, L, seed=seed)
    frames = W.make_frames_u8(4 * (B - 1) + T, seed=seed)
    img = orc.gather_clips(orc.preprocess_u8(frames), [int(s) for s in g["clip_starts"]], T)
    return g, T, L, B, ids, mask, img


def test_oracle_batch_stat_mode_matches_reference_golden():
    from oracle import two_stream_oracle as orc
    from oracle import weights as W
    g, T, L, B, ids, mask, img = golden_case("attn_T8_L32_B2")
    sd = W.make_state_dict(T, "attn", seed=123)
    with torch.no_grad(), orc.batch_stat_bn():
        logits, probs, vis, lang = orc.two_stream_forward(sd, img, ids, mask, T, 128, "attn", 8)
    assert rel(logits, g["logits"]) <= 1e-5 and rel(vis, g["vision_emb"]) <= 1e-5 and rel(probs, g["probs"]) <= 1e-5
    # the clips of a call are coupled: clip 0 alone is a different function value, pinned by the reference too
    with torch.no_grad(), orc.batch_stat_bn():
        alone = orc.two_stream_forward(sd, img[:1], ids[:1], mask[:1], T, 128, "attn", 8)[0]
    assert rel(alone, g["logits_clip0_alone"]) <= 1e-5
    assert rel(alone, g["logits"][:1]) > 1e-3
    # and the flag does not leak out of the context manager
    assert orc.BN_BATCH_STATS is False


def _standin_ops():
    """torch restatements of the operators BatchStatVision composes (fp32, NHWC) — test doubles only."""
    m = types.SimpleNamespace()

    def nchw_to_stem(img, dtype):
        assert dtype == torch.float32
        n = img.shape[0]
        out = torch.zeros(n, 230, 240, 4)
        out[:, 3:227, 3:227, :3] = img.permute(0, 2, 3, 1)
        return out

    def stem_conv_act(xp, w_packed, bias=None, act=0):
        assert act == 0 and bias is None and tuple(w_packed.shape) == (64, 7, 8, 4)
        w = w_packed[:, :, :7, :3].permute(0, 3, 1, 2)
        x = xp[:, 3:227, 3:227, :3].permute(0, 3, 1, 2)
        return F.conv2d(x, w, stride=2, padding=3).permute(0, 2, 3, 1).contiguous()

    def maxpool_tsm(x, T, shift_div):
        assert shift_div == 0
        return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous(), None

    def tsm_shift(x, T, fold):
        n, H, W, C = x.shape
        v = x.view(n // T, T, H, W, C)
        out = torch.zeros_like(v)
        out[:, :-1, ..., :fold] = v[:, 1:, ..., :fold]
        out[:, 1:, ..., fold:2 * fold] = v[:, :-1, ..., fold:2 * fold]
        out[..., 2 * fold:] = v[..., 2 * fold:]
        return out.view(n, H, W, C)

    def conv2d_nhwc(x, w, bias=None, residual=None, stride=1, act=0, **kw):
        assert act == 0 and bias is None and residual is None and not kw
        k = w.shape[1]
        y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), stride=stride, padding=k // 2)
        return y.permute(0, 2, 3, 1).contiguous()

    def bn_batch_stats(x, eps=1e-5):
        flat = x.reshape(-1, x.shape[-1]).double()
        mean = flat.mean(0)
        var = (flat * flat).mean(0) - mean * mean
        return mean.float(), (1.0 / torch.sqrt(var.clamp_min(0) + eps)).float()

    def bn_apply(x, mean, rstd, gamma, beta, residual=None, relu=True):
        y = (x - mean) * rstd * gamma + beta
        if residual is not None:
            y = y + residual
        return F.relu(y) if relu else y

    def avgpool(x):
        return x.mean(dim=(1, 2))

    for f in (nchw_to_stem, stem_conv_act, maxpool_tsm, tsm_shift, conv2d_nhwc, bn_batch_stats, bn_apply, avgpool):
        setattr(m, f.__name__, f)
    return m


def test_batch_stat_vision_orchestration_against_reference_golden(monkeypatch):
    from oracle import weights as W
    from vcg_b200 import bn_batch
    g, T, L, B, ids, mask, img = golden_case("attn_T8_L32_B2")
    sd = W.make_state_dict(T, "attn", seed=123)
    monkeypatch.setattr(bn_batch, "ops", _standin_ops())
    with torch.no_grad():
        emb = bn_batch.BatchStatVision(sd, T, 8, "fp32", "cpu").embed(img)
    assert emb.shape == (B, T, 2048) and emb.dtype == torch.float32
    assert rel(emb, g["vision_emb"]) <= 1e-4


def test_real_operators_refuse_cpu_tensors():
    from oracle import weights as W
    from vcg_b200 import bn_batch
    T = 8
    sd = W.make_state_dict(T, "attn", seed=123)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bn_batch.BatchStatVision(sd, T, 8, "fp32", "cpu").embed(torch.zeros(1, T, 3, 224, 224))


def test_stem_weight_packing_round_trip():
    from vcg_b200.bn_batch import pack_stem_weight
    w = torch.randn(64, 3, 7, 7)
    p32 = pack_stem_weight(w, torch.float32)
    assert tuple(p32.shape) == (64, 7, 8, 4)
    assert torch.equal(p32[:, :, :7, :3].permute(0, 3, 1, 2), w) and float(p32[:, :, 7].abs().max()) == 0 and float(p32[..., 3].abs().max()) == 0
    p16 = pack_stem_weight(w, torch.bfloat16)
    assert tuple(p16.shape) == (64, 4, 8, 2, 4)
    back = p16.permute(0, 1, 3, 2, 4).reshape(64, 8, 8, 4)[:, :7, :7, :3].permute(0, 3, 1, 2)
    assert torch.equal(back, w.to(torch.bfloat16))
